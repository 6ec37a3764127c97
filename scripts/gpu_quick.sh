mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -q -m gpu --tb=short > gpurun_out/test_quick.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_quick.log
