mkdir -p gpurun_out
for spec in "6 3584 512 512 0 both" "6 3584 512 512 0 f32" "6 3584 512 512 0 op" "1 3584 256 256 0 both" "2 3600 2048 512 2 op" "1 7056 1024 3072 0 f32"; do
  timeout 120 python scripts/gemm_phases_pp.py $spec
done > gpurun_out/phases_pp.log 2>&1
echo "phases exit $?"
cat gpurun_out/phases_pp.log
