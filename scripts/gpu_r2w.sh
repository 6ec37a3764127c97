mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
  python - "$tag" <<PY
import json,sys
b=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(b['value'],1), 'long', round(b['long_window']['ms_per_step'],3), 'dwconv', round(b['roofline']['per_kernel']['dwconv_ln']['us_per_step']), 'traced', round(b['roofline']['traced_step_us']))
PY
}
run base X=1
run b296 UNAV_DWCONV_BLOCKS=296
run b148 UNAV_DWCONV_BLOCKS=148
run b296s2 UNAV_DWCONV_BLOCKS=296 UNAV_DWCONV_STRIP=2
