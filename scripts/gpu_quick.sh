mkdir -p gpurun_out
python bench.py --steps 30 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value',round(b['value'],1),'ms/step',round(b['ms_per_step'],3),'e2e',round(b['e2e']['value'],1),round(b['e2e']['ms_per_step'],3),'launches',b['launches_per_step'], b.get('loop_ms'), b['gather_ms'], b['clocks'])
PY
