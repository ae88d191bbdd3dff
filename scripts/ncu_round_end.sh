#!/bin/bash
# End-of-round ncu evidence (one GPU): (1) launch list of the three eager phases of the cfg-3 step, (2) --set full of every
# hand-written kernel at bench shapes, (3) --set full of the denoising aggregator's kernels.  Each target first runs without ncu.
mkdir -p gpurun_out
TAG=${1:-r02b}
python scripts/ncu_target.py --eager > gpurun_out/${TAG}_target_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
    python scripts/ncu_target.py --eager > gpurun_out/${TAG}_target_ncu.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/${TAG}_launches_raw.csv
export VOD_PROFILE=1
python bench.py --kernels-only > gpurun_out/${TAG}_kernels_plain.log 2>&1 && \
ncu --set full --clock-control none \
    -k regex:"msra_gemm_topk|msra_rescore|msra_overflow_scan|roi_align_kernel|tafa_kernel|tafa_keyproj|selsa_tc|embed_|flow_warp" \
    -c 28 -o gpurun_out/${TAG}_full python bench.py --kernels-only > gpurun_out/${TAG}_full.log 2>&1
echo "full exit $?"
# gpurun_out/ travels back only below 64 MiB: summarise on the box, keep the report of the logits kernel alone (with source)
python scripts/ncu_summary.py gpurun_out/${TAG}_full.ncu-rep gpurun_out/${TAG}_ncu_full_summary.csv && rm -f gpurun_out/${TAG}_full.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"tafa_keyproj_persist" -c 2 -o gpurun_out/${TAG}_keyproj \
    python bench.py --kernels-only > gpurun_out/${TAG}_keyproj.log 2>&1
D="python bench.py --config denoise --steps 1 --warmup 1 --no-cpu-baseline --no-eager-reference"
$D > gpurun_out/${TAG}_denoise_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"mdcn_im2col|temporal_softmax" -s 8 -c 3 -o gpurun_out/${TAG}_denoise $D \
    > gpurun_out/${TAG}_denoise.log 2>&1
echo "denoise exit $?"; python scripts/ncu_summary.py gpurun_out/${TAG}_denoise.ncu-rep gpurun_out/${TAG}_ncu_denoise_summary.csv
ls -la gpurun_out/; du -sh gpurun_out
