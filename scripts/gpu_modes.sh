timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_model.py -q -m gpu -x --tb=short -s 2>&1 | grep -E "rel err|passed|failed|Error|assert" | head -30
for m in f16x3 fast f16 bf16x3; do
python bench.py --steps 30 --warmup 10 --no-cpu-baseline --mode $m > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; tail -1 gpurun_out/bench_$m.err
python - <<PY
import json
b=json.loads(open('gpurun_out/bench_$m.json').read().strip().splitlines()[-1])
print('$m value',round(b['value'],1),'ms/step',round(b['ms_per_step'],3),'e2e',round(b['e2e']['value'],1), 'gemm frac', round(b['roofline']['frac'],4), round(b['roofline']['achieved'],1))
PY
done
