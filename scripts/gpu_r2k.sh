mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n=$N exit $?"
python - <<PY
import json
b=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('value', round(b['value'],1), 'ms/step', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value'],1), 'det', b['detections_check']['match'])
print(json.dumps(b['config3_split'], indent=1))
PY
tail -3 gpurun_out/bench_n$N.err
