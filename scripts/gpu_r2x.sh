mkdir -p gpurun_out
( for s in "32 224 4 128" "32 224 4 64" "16 225 8 64" "32 112 4 64" "32 56 4 64"; do timeout 120 python scripts/attn_phases.py $s; done ) > gpurun_out/attn_phases.log 2>&1
cat gpurun_out/attn_phases.log
