UNAV_TC_PAIR=1 timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x --tb=short 2>&1 | tail -8
for cfg in "UNAV_TC_PAIR=0" "UNAV_TC_PAIR=1"; do
  echo "== $cfg"
  for i in 3 0 6 2 14 15 16 17 18 19 5 21; do env $cfg timeout 120 python scripts/gemm_probe.py $i x3 2>&1 | tail -1; done
done
