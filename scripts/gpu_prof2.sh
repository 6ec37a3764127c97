mkdir -p gpurun_out
python scripts/gemm_probe.py 1 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 5 -c 2 -o gpurun_out/prof_gemm_tiny -f python scripts/gemm_probe.py 1 > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"; cat gpurun_out/plain2.log
