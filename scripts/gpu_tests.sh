mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for f in gemm attention rowops nms model; do
  timeout 600 python -m pytest tests/test_gpu_$f.py -q -m gpu -s --tb=short > gpurun_out/test_$f.log 2>&1
  echo "test_gpu_$f exit $?" >> gpurun_out/summary.txt
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/test_gemm.log
