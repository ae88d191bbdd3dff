# per-kernel roofline table of bench.py --kernels-only (one line per kernel)
python bench.py --kernels-only 2>/dev/null | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())
for k, v in d.items(): print('%-32s %8.1f us  frac %.3f' % (k, v['seconds'] * 1e6, v['frac']))
print('norms+rescore %.1f us' % ((d['msra_topk_sample']['seconds'] - d['msra_gemm_topk_kernel']['seconds']) * 1e6))
"
