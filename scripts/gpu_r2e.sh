mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppair -s 3 -c 1 -o gpurun_out/prof_ppair -f python scripts/gemm_phases_pp.py 6 3584 512 512 0 both > gpurun_out/ncu_ppair.log 2>&1
echo "ncu ppair exit $?"; tail -3 gpurun_out/ncu_ppair.log; ls -la gpurun_out/*.ncu-rep
