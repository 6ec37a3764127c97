mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu -x --tb=short > gpurun_out/t_attn.log 2>&1; echo "attention tests exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_configs.py -q -m gpu -x --tb=short -s -k "config4 or config2" > gpurun_out/t_cfg4.log 2>&1; echo "config4/2 tests exit $?" >> gpurun_out/summary.txt
timeout 300 python scripts/config4_run.py 2 > gpurun_out/config4_tc.json 2> gpurun_out/config4_tc.err; echo "config4 tc exit $?" >> gpurun_out/summary.txt
UNAV_ATTN_LONG_SIMT=1 timeout 300 python scripts/config4_run.py 2 > gpurun_out/config4_simt.json 2> gpurun_out/config4_simt.err; echo "config4 simt exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_attn.log; grep "T=2304\|passed\|failed\|Error" gpurun_out/t_cfg4.log | tail -5; tail -1 gpurun_out/config4_tc.json; tail -1 gpurun_out/config4_simt.json; tail -3 gpurun_out/config4_tc.err
