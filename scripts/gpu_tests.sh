mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1800 python -m pytest tests -q -m gpu -x --tb=short > gpurun_out/test_all.log 2>&1
echo "pytest -m gpu exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -4 gpurun_out/test_all.log; tail -2 gpurun_out/smoke.log
