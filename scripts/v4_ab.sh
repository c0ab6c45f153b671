#!/bin/bash
# A/B of the STFT kernel generations (GPU box): kernel time, parity table, tests
mkdir -p gpurun_out
echo "== time v5 (kernel pair)"; timeout 300 python scripts/variant_bench.py 64 2>&1 | tail -1
echo "== time v3"; SONAR_STFT_V3=1 timeout 300 python scripts/variant_bench.py 64 2>&1 | tail -1
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_fingerprint.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_v5 -c 2 -o gpurun_out/r2_v5b -f python scripts/profile_fp.py 16 300 44100 1 > gpurun_out/r2_v5b_ncu.log 2>&1; tail -2 gpurun_out/r2_v5b_ncu.log
