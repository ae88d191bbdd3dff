# ncu --set full of the hand-written kernels at bench shapes (python bench.py --kernels-only, one launch each)
mkdir -p gpurun_out
export VOD_PROFILE=1
TAG=${1:-r1e}
CMD="python bench.py --kernels-only"
$CMD > gpurun_out/kernels_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"msra_gemm_topk|msra_rescore|roi_align_kernel|tafa_kernel|tafa_keyproj|selsa_tc|embed_|flow_warp|nchw_to_nhwc|rows_l2norm" -c 30 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
