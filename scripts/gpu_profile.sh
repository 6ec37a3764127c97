# round-1 final profile set: launch list (durations), DRAM bytes of every GEMM launch, full capture of the heaviest GEMM
mkdir -p gpurun_out
REGEX='regex:^(gemm_|attention_|ln_rows|dwconv|rowcopy|maxsig|softnms|merge_|decode_|transpose_cast|align_embed|build_masks|pool_match|collate_pad|map_match)'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD --trace-out gpurun_out/trace.json > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$REGEX" -c 520 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "launch list exit $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^gemm_tcgen05' -c 200 --csv --log-file gpurun_out/gemm_dram.csv $CMD > gpurun_out/ncu2.log 2>&1
echo "gemm dram exit $?"
IDX=$(python - <<'PY'
import json
tr = json.load(open('gpurun_out/trace.json'))
g = [t for t in tr if t['kernel'].startswith('gemm_tcgen05')]
k = max(range(len(g)), key=lambda i: g[i]['us'])
print(k)
import sys
print(len(g), g[k], file=sys.stderr)
PY
)
echo "heaviest gemm index $IDX"
ncu --set full --clock-control none --import-source on -k 'regex:^gemm_tcgen05' -s $IDX -c 1 -o gpurun_out/prof_gemm_top -f $CMD > gpurun_out/ncu3.log 2>&1
echo "full capture exit $?"
ncu --set full --clock-control none --import-source on -k "regex:^dwconv_ln" -c 3 -o gpurun_out/prof_attn_dw -f $CMD > gpurun_out/ncu4.log 2>&1
echo "attn/dw capture exit $?"
# the kernel bench.py's roofline names as dominant (gemm_tcgen05_kernel<128, 32>): its heaviest launch, and one LayerNorm launch
IDX2=$(python - <<'PY'
import json
tr = json.load(open('gpurun_out/trace.json'))
g = [t for t in tr if t['kernel'].startswith('gemm_tcgen05')]
c = [i for i in range(len(g)) if g[i]['kernel'] == 'gemm_tcgen05_kernel<128, 32>']
print(max(c, key=lambda i: g[i]['us']) if c else 0)
PY
)
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:^gemm_tcgen05' -s $IDX2 -c 1 -o gpurun_out/prof_gemm_dom -f $CMD > gpurun_out/ncu5.log 2>&1
echo "dominant-kernel capture (gemm launch $IDX2) exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:^ln_rows' -s 4 -c 1 -o gpurun_out/prof_ln -f $CMD > gpurun_out/ncu6.log 2>&1
echo "ln_rows capture exit $?"
