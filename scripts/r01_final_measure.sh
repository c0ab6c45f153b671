set -x
python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/final_tests.txt
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference_arm.json 2> gpurun_out/r01_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-s16 --pairs 8 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"yin_frame_fft|ncc_tiled|dtw_fill_warp|frame_walk|stft_v2|rms_from_parts|znorm|xs_spectrum|xs_curve" -s 18 -c 9 -o gpurun_out/r01_kernels -f python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu --no-s16 --no-profile --no-clocks > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"xs_pass" -s 48 -c 2 -o gpurun_out/r01_xs_pass -f python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu --no-s16 --no-profile --no-clocks > gpurun_out/ncu_full2.log 2>&1
cat gpurun_out/final_tests.txt
ls -la gpurun_out/*.ncu-rep
