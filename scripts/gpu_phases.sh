timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_model.py -q -m gpu -x --tb=short 2>&1 | tail -5
python scripts/gemm_phases.py 2 3584 2048 512 2
python scripts/gemm_phases.py 1 7056 1024 3072
python scripts/gemm_phases.py 1 448 256 256
for i in 0 2 6 3 10 4 5 1 9; do python scripts/gemm_probe.py $i x3 2>&1 | tail -1; done
bash scripts/gpu_quick.sh
