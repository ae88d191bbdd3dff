mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
CMD="python bench.py --steps 2 --warmup 1 --no-roofline --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
