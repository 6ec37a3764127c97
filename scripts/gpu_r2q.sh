mkdir -p gpurun_out
for BS in 16 32 16; do
UNAV_CONFIG3_BATCH=$BS timeout 300 python scripts/config3_run.py > gpurun_out/config3_n1_b$BS.json 2> gpurun_out/config3_n1_b$BS.err; echo "config3 n=1 b=$BS exit $?"; tail -1 gpurun_out/config3_n1_b$BS.json | cut -c1-620
done
