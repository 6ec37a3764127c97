mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_configs.py -q -m gpu --tb=short -s > gpurun_out/test_cfg.log 2>&1; echo "tests exit $?"; grep -E "rel err|passed|failed|Error|error" gpurun_out/test_cfg.log | head -20; tail -5 gpurun_out/test_cfg.log
