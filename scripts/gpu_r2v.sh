# SM-occupancy-weighted launch list: which kernels consume GPU capacity (active SM fraction x duration)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-run"
$CMD > gpurun_out/plain.json 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
REGEX='regex:^(gemm_|attention_|ln_rows|dwconv|rowcopy|maxsig|softnms|merge_|decode_|transpose_cast|align_embed|build_masks|pool_match|collate_pad)'
timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size,sm__cycles_active.avg,sm__cycles_elapsed.avg,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k "$REGEX" -c 1300 --csv --log-file gpurun_out/launches_occ.csv $CMD > gpurun_out/ncu_occ.log 2>&1
echo "launch list exit $?"
wc -l gpurun_out/launches_occ.csv
