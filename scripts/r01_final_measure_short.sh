python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/final_tests.txt
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-s16 --pairs 8 > gpurun_out/ncu_list.log 2>&1
cat gpurun_out/final_tests.txt
python -c "
import json; d=json.load(open('gpurun_out/r01_bench.json')); print(round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), round(d['e2e_s16_ingest']['value']), round(d['fingerprint_only']['ms_per_step'],2), d['gpu_launches'], d['roofline']['frac'])"
