mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "roi_align or temporal" 2>&1 | tail -3
python bench.py --kernels-only 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d.items(): print('%-22s %8.1f us  %8.1f %s  frac %.3f'%(k, v['seconds']*1e6, v['achieved'], v['unit'], v['frac']))
"
