mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short > gpurun_out/test_all.log 2>&1; echo "pytest -m gpu exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/test_all.log; tail -1 gpurun_out/smoke.log
python - <<PY
import json
b=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', round(b['value'],1), 'ms/step', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value'],1), 'long', round(b['long_window']['ms_per_step'],3), 'det', b['detections_check']['match'])
print('cpu', b['cpu_baseline']['value'], b['parity_vs_reference'])
print(json.dumps(b['config3_split'])[:700])
print(json.dumps(b['config1_batch1'])[:600])
PY
