# round-2 profile set (final build): plain run with the CUDA-event trace, ncu launch list with DRAM bytes per launch, and one
# `--set full` capture of each kernel the review named; every ncu pass runs the same command after it exited 0 without ncu
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-run"
$CMD --trace-out gpurun_out/trace.json > gpurun_out/plain.json 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
REGEX='regex:^(gemm_|attention_|ln_rows|dwconv|rowcopy|maxsig|softnms|merge_|decode_|transpose_cast|align_embed|build_masks|pool_match|collate_pad|pack_operand)'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size,sm__cycles_active.avg,sm__cycles_elapsed.avg --clock-control none -k "$REGEX" -c 1400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -o gpurun_out/prof_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "capture $1 exit $?"
}
# skip counts: the weight packing and the eager warm-up pass precede the replays; any launch of the class is representative
cap ppair 'gemm_tcgen05_ppair_kernel' 30
cap gemm6464 'gemm_tcgen05_kernel<\(int\)64, \(int\)64>' 100
cap gemm12832 'gemm_tcgen05_kernel<\(int\)128, \(int\)32>' 20
cap attention 'attention_tcgen05_kernel' 40
cap maxsig 'maxsig_tcgen05_kernel' 12
cap softnms 'softnms_lazy_kernel' 3
cap decode 'decode_kernel' 3
cap dwstream 'dwconv_ln_stream_kernel' 4
cap dwtiled 'dwconv_ln_kernel' 60
cap lnrows 'ln_rows_exact_kernel' 30
cap rowcopy 'rowcopy_kernel' 30
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
