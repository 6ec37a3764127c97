# First validation + A/B of the 128 x 256 GEMM tiles (UNAV_TC_BN=256).  Every step under its own timeout (a hang must not eat the box).
mkdir -p gpurun_out
UNAV_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x --tb=short -k bn256 > gpurun_out/t_bn256.log 2>&1; echo "bn256 parity exit $?"; tail -4 gpurun_out/t_bn256.log
for i in 1 2; do
for v in default 256; do
if [ $v = default ]; then unset UNAV_TC_BN; else export UNAV_TC_BN=$v; fi
timeout 120 python bench.py --steps 30 --warmup 10 --no-cpu-baseline 2>gpurun_out/ab_bn_$v.err | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bn=$v', round(b['ms_per_step'],3), round(b['e2e']['ms_per_step'],3), {k:(v['us_per_step'],v['tflops']) for k,v in b['roofline']['per_kernel'].items() if 'gemm' in k})"
echo "bench bn=$v exit $?"
done
done
