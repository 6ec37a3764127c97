mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_configs.py -q -m gpu -x --tb=short > gpurun_out/t_ingest.log 2>&1; echo "ingest+configs tests exit $?"; tail -3 gpurun_out/t_ingest.log
timeout 300 python scripts/config3_run.py > gpurun_out/config3_n1.json 2> gpurun_out/config3_n1.err; echo "config3 n=1 exit $?"; tail -1 gpurun_out/config3_n1.json | cut -c1-600
