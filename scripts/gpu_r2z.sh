mkdir -p gpurun_out
c3() { python - "$1" <<PY
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c=b["config3_split"]; print(sys.argv[1], "value", round(b["value"]), "config3", round(c["videos_per_s"]), c["rank0_breakdown"]["host_submit_s"])
PY
}
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_after_smoke.json 2>/dev/null; c3 gpurun_out/b_after_smoke.json
timeout 600 python -m pytest tests/test_gpu_eval_dropin.py -q -m gpu -x > gpurun_out/t_a.log 2>&1; echo "dropin exit $?"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_after_dropin.json 2>/dev/null; c3 gpurun_out/b_after_dropin.json
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_model.py tests/test_gpu_ingest.py tests/test_gpu_metrics.py -q -m gpu -x > gpurun_out/t_b.log 2>&1; echo "configs exit $?"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_after_configs.json 2>/dev/null; c3 gpurun_out/b_after_configs.json
free -g | head -2
