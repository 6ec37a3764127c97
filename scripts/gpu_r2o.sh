mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ingest.py -q -m gpu -x --tb=short > gpurun_out/t_ingest.log 2>&1; echo "ingest tests exit $?"; tail -3 gpurun_out/t_ingest.log
for BS in 16 32; do
UNAV_CONFIG3_BATCH=$BS timeout 300 python scripts/config3_run.py > gpurun_out/config3_n1_b$BS.json 2> gpurun_out/config3_n1_b$BS.err; echo "config3 n=1 b=$BS exit $?"; tail -1 gpurun_out/config3_n1_b$BS.json | cut -c1-520
done
