mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
  python - "$tag" <<PY
import json,sys
b=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(b['value'],1), 'long', round(b['long_window']['ms_per_step'],3), 'e2e', round(b['e2e']['value']), 'b1 sync ms', round(b['config1_batch1']['gpu_ms_per_video_sync'],3), 'traced', round(b['roofline']['traced_step_us']))
PY
}
run hack1 UNAV_ATTN_HACK_HEADS=1
run base1 X=1
run hack2 UNAV_ATTN_HACK_HEADS=1
run base2 X=1
