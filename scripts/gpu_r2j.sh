mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_rowops.py tests/test_gpu_configs.py -q -m gpu -x --tb=short > gpurun_out/t_rowops.log 2>&1; echo "rowops+configs tests exit $?" >> gpurun_out/summary.txt
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench exit $?" >> gpurun_out/summary.txt
python - <<PY
import json
b=json.loads(open('gpurun_out/bench_c3.json').read().strip().splitlines()[-1])
print('ms/step', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['ms_per_step'],3), 'long', round(b['long_window']['ms_per_step'],3), 'det', b['detections_check']['match'])
print(json.dumps(b['config3_split'], indent=1))
PY
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_rowops.log; tail -5 gpurun_out/bench_c3.err
