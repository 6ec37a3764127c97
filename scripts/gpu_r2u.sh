mkdir -p gpurun_out
for cfg in "bf16x3 8" "bf16x3 16" "fast 8" "f16 8" "f16x3 8"; do
  set -- $cfg
  UNAV_PP_EW=$2 timeout 300 python bench.py --mode $1 --no-cpu-baseline > gpurun_out/bench_$1_ew$2.json 2> gpurun_out/bench_$1_ew$2.err
  python - "$1" "$2" <<PY
import json,sys
b=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}_ew{sys.argv[2]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'ew', sys.argv[2], 'value', round(b['value'],1), 'ms/step', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value'],1), 'long', round(b['long_window']['ms_per_step'],3), 'det', b['detections_check']['match'], 'traced', round(b['roofline']['traced_step_us']))
PY
done
