mkdir -p gpurun_out
( UNAV_PP_FINE=1 timeout 100 python scripts/gemm_phases_pp.py 2 3584 512 512 0 f32; UNAV_PP_FINE=1 timeout 100 python scripts/gemm_phases_pp.py 6 3584 512 512 0 f32; UNAV_PP_FINE=1 timeout 100 python scripts/gemm_phases_pp.py 6 3584 512 512 0 both;  UNAV_PP_FINE=1 timeout 100 python scripts/gemm_phases_pp.py 2 3600 2048 512 2 op ) > gpurun_out/phases_fine.log 2>&1
cat gpurun_out/phases_fine.log
