#!/bin/bash
# A/B of the warp-specialised STFT kernel against the single-role one (GPU box): parity table, kernel time, tests
mkdir -p gpurun_out
for m in 0 1 2; do echo "== time v4 lockstep=$m"; SONAR_V4_LOCKSTEP=$m timeout 300 python scripts/variant_bench.py 64 2>&1 | tail -1; done
echo "== time v3"; SONAR_STFT_V3=1 timeout 300 python scripts/variant_bench.py 64 2>&1 | tail -1
echo "== parity (v4)"; timeout 600 python scripts/dev_check_fp.py > gpurun_out/v4_check.txt 2>&1; echo rc=$?
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_fingerprint.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -4
