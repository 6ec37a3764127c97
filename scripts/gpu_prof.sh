mkdir -p gpurun_out
REGEX='regex:^(gemm_|attention_|ln_rows|dwconv|rowcopy|maxsig|softnms|merge_|decode_|transpose_cast|align_embed|build_masks|pool_match)'
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$REGEX" -c 520 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu list exit $?"
