mkdir -p gpurun_out
N=${1:-8}
for BS in ${2:-32}; do
if [ $N = 1 ]; then
UNAV_CONFIG3_BATCH=$BS timeout 600 python scripts/config3_run.py > gpurun_out/config3_n${N}_b$BS.json 2> gpurun_out/config3_n${N}_b$BS.err
else
UNAV_CONFIG3_BATCH=$BS timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/config3_run.py > gpurun_out/config3_n${N}_b$BS.json 2> gpurun_out/config3_n${N}_b$BS.err
fi
echo "config3 n=$N b=$BS exit $?"
tail -1 gpurun_out/config3_n${N}_b$BS.json | cut -c1-520
done
