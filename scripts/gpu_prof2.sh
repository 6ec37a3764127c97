mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:^(softnms|merge_|decode_|maxsig|dwconv)' -c 9 -o gpurun_out/prof_misc -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
echo "ncu exit $?"
