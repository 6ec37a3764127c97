# Run on the B200 box: bench (ours + reference arm), then the ncu launch list of the same short command.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --trace-out gpurun_out/trace.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/bench.json
if [ "$1" = "ref" ]; then
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
fi
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"_kernel$" -c 520 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
tail -3 gpurun_out/bench.err
