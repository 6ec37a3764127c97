mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_rowops.py tests/test_gpu_modules.py -q -m gpu -x --tb=short > gpurun_out/t_rowops.log 2>&1; echo "rowops+modules tests exit $?" >> gpurun_out/summary.txt
timeout 300 python scripts/dwconv_probe.py > gpurun_out/dwconv_probe.log 2>&1; echo "dwconv probe exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_dw.json 2> gpurun_out/bench_dw.err; echo "bench exit $?" >> gpurun_out/summary.txt
python - <<PY
import json
b=json.loads(open('gpurun_out/bench_dw.json').read().strip().splitlines()[-1])
print('ms/step', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['ms_per_step'],3), 'long', round(b['long_window']['ms_per_step'],3), 'det', b['detections_check']['match'], b['detections_check']['timed_loop_crc32'])
for k,v in b['roofline']['per_kernel'].items(): print('   ', k, v)
PY
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_rowops.log; cat gpurun_out/dwconv_probe.log
