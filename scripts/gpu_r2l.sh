mkdir -p gpurun_out
N=${1:-8}
nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/config3_run.py > gpurun_out/config3_n$N.json 2> gpurun_out/config3_n$N.err
echo "config3 n=$N exit $?"
tail -1 gpurun_out/config3_n$N.json; tail -3 gpurun_out/config3_n$N.err
