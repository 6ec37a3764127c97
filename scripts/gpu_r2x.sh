mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -q -m gpu -x --tb=short > gpurun_out/t_gemm.log 2>&1; echo "gemm/model tests exit $?"; tail -2 gpurun_out/t_gemm.log
timeout 200 python scripts/launch_gap_probe.py 2>&1 | grep gemm
UNAV_PP_NO_DRY=1 timeout 200 python scripts/launch_gap_probe.py 2>&1 | grep gemm
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
  python - "$tag" <<PY
import json,sys
b=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
pk=b['roofline']['per_kernel']
print(sys.argv[1], 'value', round(b['value'],1), 'long', round(b['long_window']['ms_per_step'],3), 'det', b['detections_check']['match'], 'traced', round(b['roofline']['traced_step_us']), {k[13:]:round(v['us_per_step']) for k,v in pk.items() if 'gemm' in k}, 'b1 sync ms', round(b['config1_batch1']['gpu_ms_per_video_sync'],3))
PY
}
run dry X=1
run nodry UNAV_PP_NO_DRY=1
