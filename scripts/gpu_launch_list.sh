# the launch-list part of gpu_profile_r2.sh alone (plain run with the CUDA-event trace, then the ncu list with DRAM bytes, grid and
# SM-active cycles per launch); make_profile_summary_r2.py keeps the --set full captures already in gpurun_out/
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-run"
$CMD --trace-out gpurun_out/trace.json > gpurun_out/plain.json 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
REGEX='regex:^(gemm_|attention_|ln_rows|dwconv|rowcopy|maxsig|softnms|merge_|decode_|transpose_cast|align_embed|build_masks|pool_match|collate_pad|pack_operand)'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size,sm__cycles_active.avg,sm__cycles_elapsed.avg --clock-control none -k "$REGEX" -c 1400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
