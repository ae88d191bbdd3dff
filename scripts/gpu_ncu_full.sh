mkdir -p gpurun_out
CMD="python bench.py --kernels-only"
$CMD > gpurun_out/kernels_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"msra_gemm_topk|msra_rescore|roi_align_kernel|selsa_tc_kernel|nms_mask|nms_rank|nms_sweep" -s 7 -c 14 -o gpurun_out/prof_r1a $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log; cat gpurun_out/kernels_plain.log | tail -2; ls -la gpurun_out/*.ncu-rep
