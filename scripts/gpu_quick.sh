mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py -q -m gpu --tb=short > gpurun_out/test_modules.log 2>&1; echo "modules exit $?"; tail -25 gpurun_out/test_modules.log
