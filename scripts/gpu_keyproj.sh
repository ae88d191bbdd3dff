mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -q -m gpu --timeout 300 -x -k "msra or temporal or candidate" 2>&1 | tail -3
python scripts/probe_msra.py 2>&1 | tail -1
VOD_KP_PERSIST=1 VOD_KP_TB=16 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "tafa or temporal_roi_align_golden or keyproj or dff" 2>&1 | tail -3
echo "shipped (one CTA per 16-frame tile)"; timeout 300 python scripts/probe_keyproj.py 16 300 2>&1 | tail -1 | sed 's/.*logits/logits/;s/| apply.*//'
for cfg in "14 2" "7 4" "15 1" "10 3"; do set -- $cfg; echo "persist TB16 w$1 d$2"; VOD_KP_PERSIST=1 VOD_KP_TB=16 VOD_KP_WARPS=$1 VOD_KP_DEPTH=$2 timeout 300 python scripts/probe_keyproj.py 16 300 2>&1 | tail -1 | sed 's/.*logits/logits/;s/| apply.*//'; done
echo "persist TB16 w14 d2, T1=32"; VOD_KP_PERSIST=1 VOD_KP_TB=16 timeout 300 python scripts/probe_keyproj.py 32 300 2>&1 | tail -1 | sed 's/.*logits/logits/;s/| apply.*//'
