# A/B of the GEMM tile policy on one box: UNAV_TC_PAIR unset (default policy) vs 2 (256 x 128 CTA-pair tiles on every full grid)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x --tb=short -k "pair" > gpurun_out/t_pair.log 2>&1; echo "pair tests exit $?"; tail -5 gpurun_out/t_pair.log
for i in 1 2; do
for v in default 2; do
if [ $v = default ]; then unset UNAV_TC_PAIR; else export UNAV_TC_PAIR=$v; fi
python bench.py --steps 30 --warmup 10 --no-cpu-baseline 2>gpurun_out/ab_$v.err | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pair=$v', round(b['ms_per_step'],3), round(b['e2e']['ms_per_step'],3), {k:round(v,4) for k,v in b['roofline']['kernel_time_shares'].items() if 'gemm' in k}, b['roofline']['traced_step_us'])"
done
done
