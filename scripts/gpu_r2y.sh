mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rowops.py tests/test_gpu_modules.py tests/test_gpu_model.py tests/test_gpu_configs.py -q -m gpu -x --tb=short -k "not config3" > gpurun_out/t_quick.log 2>&1; echo "rowops/modules/model/configs tests exit $?"
tail -5 gpurun_out/t_quick.log
timeout 300 python scripts/dwconv_probe.py > gpurun_out/dwconv_probe.log 2>&1; head -4 gpurun_out/dwconv_probe.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
  python - "$tag" <<PY
import json,sys
b=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(b['value'],1), 'ms/step', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value'],1), 'long', round(b['long_window']['ms_per_step'],3), 'det', b['detections_check']['match'], 'traced', round(b['roofline']['traced_step_us']), 'b1 sync ms', round(b['config1_batch1']['gpu_ms_per_video_sync'],3))
pk=b['roofline']['per_kernel']
print({k:round(v['us_per_step']) for k,v in pk.items()})
PY
}
run new X=1
