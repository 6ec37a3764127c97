mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:^dwconv' -c 2 -o gpurun_out/prof_dw -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:^(softnms|merge_|decode_)' -c 3 -o gpurun_out/prof_nms -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
echo done
