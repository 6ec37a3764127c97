for i in 1 2; do
for d in . _ab; do
( cd $d; python bench.py --steps 30 --warmup 10 --no-cpu-baseline --mode bf16x3 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$d', round(b['ms_per_step'],3), round(b['e2e']['ms_per_step'],3))" )
done
done
