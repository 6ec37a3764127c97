# round 2, call B: first validation of the persistent CTA-pair GEMM (parity, two-stream interleave, A/B per shape); every step under its own timeout
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x --tb=short -k "persistent" > gpurun_out/t_ppair.log 2>&1; echo "ppair tests exit $?" >> gpurun_out/summary.txt
timeout 300 python scripts/gemm_ab.py x3 > gpurun_out/ab_x3.log 2>&1; echo "ab x3 exit $?" >> gpurun_out/summary.txt
AB_PASSES=1 timeout 300 python scripts/gemm_ab.py x3 > gpurun_out/ab_x3_p1.log 2>&1; echo "ab x3 1-pass exit $?" >> gpurun_out/summary.txt
UNAV_TC_STAGES=4 timeout 300 python scripts/gemm_ab.py x3 > gpurun_out/ab_x3_s4.log 2>&1; echo "ab x3 4 stages exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_ppair.log; cat gpurun_out/ab_x3.log
