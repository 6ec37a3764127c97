mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nms.py tests/test_gpu_model.py -q -m gpu --tb=short > gpurun_out/test_quick.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/test_quick.log
python bench.py --steps 30 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value',round(b['value'],1),'ms/step',round(b['ms_per_step'],3),'e2e',round(b['e2e']['value'],1),round(b['e2e']['ms_per_step'],3),'launches',b['launches_per_step'], b['step_ms_min_med_max'], b['gather_ms'])
print(b['roofline']['kernel_time_shares'])
PY
