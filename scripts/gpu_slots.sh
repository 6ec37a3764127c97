# throughput vs number of batches in flight (engine plans / forward streams used alternately)
for n in 1 2 3 4; do
UNAV_BENCH_SLOTS=$n python bench.py --steps 30 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slots $n value', round(b['value'],1), round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value'],1), round(b['e2e']['ms_per_step'],3), b['loop_ms'], b['gather_ms'])"
done
