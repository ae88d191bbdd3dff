nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 > gpurun_out/parity.log 2>&1; echo "parity exit $?"; tail -25 gpurun_out/parity.log
timeout 600 python -m pytest tests/test_gpu_tc.py -q -s -m gpu --timeout 120 > gpurun_out/tc.log 2>&1; echo "tc exit $?"; tail -40 gpurun_out/tc.log
