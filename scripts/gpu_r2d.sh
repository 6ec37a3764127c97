mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_modules.py -q -m gpu -x --tb=short > gpurun_out/t_gemm.log 2>&1; echo "gemm+modules tests exit $?" >> gpurun_out/summary.txt
for spec in "6 3584 512 512 0 both" "2 3600 2048 512 2 op" "1 3584 256 256 0 both"; do
  timeout 120 python scripts/gemm_phases_pp.py $spec
done > gpurun_out/phases_pp.log 2>&1
timeout 300 python scripts/gemm_ab.py x3 > gpurun_out/ab_x3.log 2>&1; echo "ab x3 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_gemm.log; grep -A12 "tile 1" gpurun_out/phases_pp.log | head -60; cat gpurun_out/ab_x3.log
