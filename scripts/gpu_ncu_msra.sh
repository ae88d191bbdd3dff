mkdir -p gpurun_out
export VOD_PROFILE=1
CMD="python bench.py --kernels-only"
$CMD > gpurun_out/kernels_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"msra_gemm_topk|msra_rescore|rows_l2norm" -c 6 -o gpurun_out/prof_r1d $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_full.log
