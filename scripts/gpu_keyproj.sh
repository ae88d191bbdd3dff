mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "tafa or temporal_roi_align_golden or keyproj" 2>&1 | tail -8
echo "persist w14 d2"; timeout 300 python scripts/probe_keyproj.py 16 300 2>&1 | tail -2
for dbg in 1 6 7; do echo "persist dbg=$dbg"; VOD_KP_DBG=$dbg timeout 300 python scripts/probe_keyproj.py 16 300 2>&1 | tail -1 | sed 's/.*logits/logits/;s/| apply.*//'; done
echo "persist w12 d2"; VOD_KP_WARPS=12 timeout 300 python scripts/probe_keyproj.py 16 300 2>&1 | tail -1| sed 's/.*logits/logits/;s/| apply.*//'
echo "persist w10 d3"; VOD_KP_WARPS=10 VOD_KP_DEPTH=3 timeout 300 python scripts/probe_keyproj.py 16 300 2>&1 | tail -1| sed 's/.*logits/logits/;s/| apply.*//'
timeout 300 python scripts/probe_keyproj.py 32 300 2>&1 | tail -2
timeout 300 python scripts/probe_keyproj.py 8 300 2>&1 | tail -2
