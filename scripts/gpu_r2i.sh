mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_ln_stream -c 40 -o gpurun_out/prof_dws -f python scripts/dwconv_probe.py > gpurun_out/ncu_dws.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_dws.log
