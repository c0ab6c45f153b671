#!/bin/bash
# Round-2 measurement pass (GPU box): full -m gpu suite, the default bench line, the reference arm, the ncu launch list of
# the same bench command (small batch) and one ncu --set full capture of the step's main kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/r02_final_tests.txt
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-s16 --no-legs --pairs 8 > gpurun_out/r02_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"yin32_kernel|frame_walk|stft_v5|spectral_exact_kernel|yin_track" -c 6 -o gpurun_out/r02_kernels -f python scripts/profile_fp.py 16 300 44100 1 > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv
ncu --set full --clock-control none --import-source on -k regex:"dtw_fill_warp|znorm_kernel|ncc_tiled" -c 3 -o gpurun_out/r02_align_kernels -f python bench.py --steps 1 --warmup 1 --pairs 8 --no-cpu --no-s16 --no-legs --no-profile --no-clocks > gpurun_out/r02_ncu_full2.log 2>&1
ncu -i gpurun_out/r02_align_kernels.ncu-rep --page raw --csv > gpurun_out/r02_align_kernels_raw.csv
cat gpurun_out/r02_final_tests.txt
tail -c 600 gpurun_out/r02_bench.err
ls -la gpurun_out
