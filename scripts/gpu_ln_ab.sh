# LayerNorm kernel A/B on one box: exact-width instantiations (default) vs the generic kernel (UNAV_LN_GENERIC=1)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rowops.py tests/test_gpu_model.py -q -m gpu -x --tb=short > gpurun_out/t_ln.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t_ln.log
for i in 1 2; do
for v in default generic; do
if [ $v = default ]; then unset UNAV_LN_GENERIC; else export UNAV_LN_GENERIC=1; fi
timeout 120 python bench.py --steps 30 --warmup 10 --no-cpu-baseline 2>gpurun_out/ab_ln_$v.err | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); pk=b['roofline']['per_kernel']['layernorm_rows']; print('ln=$v', round(b['ms_per_step'],3), round(b['e2e']['ms_per_step'],3), pk)"
done
done
