mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x --tb=short > gpurun_out/t_gemm.log 2>&1; echo "gemm tests exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_configs.py -q -m gpu -x --tb=short -k "not config3 and not config4" > gpurun_out/t_model.log 2>&1; echo "model/config tests exit $?" >> gpurun_out/summary.txt
timeout 400 python scripts/gemm_ab.py x3 > gpurun_out/ab_x3.log 2>&1; echo "gemm_ab exit $?" >> gpurun_out/summary.txt
timeout 100 python scripts/gemm_phases_pp.py 2 3600 2048 512 2 op > gpurun_out/phases_pp.log 2>&1
timeout 100 python scripts/gemm_phases_pp.py 1 7200 1536 512 0 both >> gpurun_out/phases_pp.log 2>&1
UNAV_PP_EW=8 timeout 100 python scripts/gemm_phases_pp.py 2 3600 2048 512 2 op >> gpurun_out/phases_pp.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_gemm.log; tail -3 gpurun_out/t_model.log; tail -32 gpurun_out/ab_x3.log; python - <<PY
import json
b=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print('value', round(b['value'],1), 'ms/step', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value'],1), 'det', b['detections_check']['match'])
print(json.dumps(b['roofline']['kernel_time_shares']))
PY
