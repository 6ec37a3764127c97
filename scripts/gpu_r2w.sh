mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
  python - "$tag" <<PY
import json,sys
b=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(b['value'],1), 'ms/step', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value'],1), 'long', round(b['long_window']['ms_per_step'],3), 'det', b['detections_check']['match'], 'traced', round(b['roofline']['traced_step_us']), 'dwconv', round(b['roofline']['per_kernel']['dwconv_ln']['us_per_step']))
PY
}
run strip2 UNAV_DWCONV_STRIP=2
run strip3 UNAV_DWCONV_STRIP=3
run strip8 UNAV_DWCONV_STRIP=8
