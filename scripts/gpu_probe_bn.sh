for st in 2 3 4; do echo "== stages $st (BN=64 ONCE=1)"; UNAV_TC_BN=64 UNAV_TC_ONCE=1 UNAV_TC_STAGES=$st python scripts/gemm_probe.py 9 x3 | grep TF; done
for st in 2 4 8; do echo "== stages $st (BN=64 ONCE=2)"; UNAV_TC_BN=64 UNAV_TC_ONCE=2 UNAV_TC_STAGES=$st python scripts/gemm_probe.py 9 x3 | grep TF; done
UNAV_TC_BN=64 UNAV_TC_ONCE=1 python scripts/gemm_phases.py 1 448 512 1536
