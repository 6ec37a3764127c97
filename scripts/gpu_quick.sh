mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -q -m gpu --tb=short > gpurun_out/test_quick.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_quick.log
python scripts/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1; cat gpurun_out/gemm_probe.log
python bench.py --steps 20 --warmup 5 --trace-out gpurun_out/trace.json --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
REGEX='regex:^(gemm_|attention_|ln_rows|dwconv|rowcopy|maxsig|softnms|merge_|decode_|transpose_cast|align_embed|build_masks|pool_match)'
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$REGEX" -c 520 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu list exit $?"
