mkdir -p gpurun_out
REGEX='regex:^(gemm_|attention_|ln_rows|dwconv|rowcopy|maxsig|softnms|merge_|decode_|transpose_cast|align_embed|build_masks|pool_match)'
python scripts/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1; cat gpurun_out/gemm_probe.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$REGEX" -c 520 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu list exit $?"
python scripts/gemm_probe.py 0 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 5 -c 2 -o gpurun_out/prof_gemm -f python scripts/gemm_probe.py 0 > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"
