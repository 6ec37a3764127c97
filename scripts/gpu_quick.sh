for pdl in 0 1; do echo "== PDL=$pdl"; UNAV_PDL=$pdl python scripts/gemm_probe.py all | grep -E "\[448|\[128|\[896"; done
