mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for v in 0 default; do
if [ $v = default ]; then unset UNAV_TC_PPAIR; else export UNAV_TC_PPAIR=$v; fi
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pp_$v.json 2> gpurun_out/bench_pp_$v.err; echo "bench ppair=$v exit $?" >> gpurun_out/summary.txt
python - <<PY
import json
b=json.loads(open('gpurun_out/bench_pp_$v.json').read().strip().splitlines()[-1])
print('ppair=$v', 'ms/step', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['ms_per_step'],3), 'long', round(b['long_window']['ms_per_step'],3), 'det', b['detections_check']['match'], b['detections_check']['timed_loop_crc32'])
for k,v in b['roofline']['per_kernel'].items():
    if 'gemm' in k: print('   ', k, v)
PY
done
unset UNAV_TC_PPAIR
timeout 900 python -m pytest tests -q -m gpu -x --tb=short > gpurun_out/test_all.log 2>&1; echo "pytest -m gpu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/test_all.log
