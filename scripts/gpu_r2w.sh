mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
  python - "$tag" <<PY
import json,sys
b=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(b['value'],1), 'long', round(b['long_window']['ms_per_step'],3), 'det', b['detections_check']['match'], 'softnms us', round(b['roofline']['per_kernel']['softnms']['us_per_step']), 'b1 sync ms', round(b['config1_batch1']['gpu_ms_per_video_sync'],3))
PY
}
run w0 UNAV_NMS_WARP_MAX=0
run w32 UNAV_NMS_WARP_MAX=32
run w64 UNAV_NMS_WARP_MAX=64
run w128 UNAV_NMS_WARP_MAX=128
