# round 2, call A: phase-A validation (tests, smoke, bench with the real-reference CPU arm, reference arm), BN=256 experiment
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short -s > gpurun_out/test_all.log 2>&1
echo "pytest -m gpu exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?" >> gpurun_out/summary.txt
UNAV_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x --tb=short -k bn256 > gpurun_out/t_bn256.log 2>&1; echo "bn256 parity exit $?" >> gpurun_out/summary.txt
timeout 200 python scripts/gemm_probe.py all x3 > gpurun_out/probe_default.log 2>&1; echo "probe default exit $?" >> gpurun_out/summary.txt
UNAV_TC_BN=256 timeout 200 python scripts/gemm_probe.py all x3 > gpurun_out/probe_bn256.log 2>&1; echo "probe bn256 exit $?" >> gpurun_out/summary.txt
UNAV_TC_PAIR=1 timeout 200 python scripts/gemm_probe.py all x3 > gpurun_out/probe_pair1.log 2>&1; echo "probe pair1 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/test_all.log; tail -2 gpurun_out/smoke.log; tail -c 600 gpurun_out/bench.err
