mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-run"
$CMD --trace-out gpurun_out/trace.json > gpurun_out/plain.json 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -o gpurun_out/prof_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "capture $1 exit $?"
}
cap gemm6464 'gemm_tcgen05_kernel<\(int\)64, \(int\)64>' 100
cap gemm12832 'gemm_tcgen05_kernel<\(int\)128, \(int\)32>' 20
cap ppair 'gemm_tcgen05_ppair_kernel' 30
ls -la gpurun_out/prof_gemm*.ncu-rep gpurun_out/prof_ppair.ncu-rep | awk '{print $5, $9}'
