# ncu --set full of the key-projected logits kernel (probe_keyproj.py at cfg-3 size), one launch per configuration
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"tafa_keyproj" -s 3 -c 1 -o gpurun_out/prof_kp_p0 -f python scripts/probe_keyproj.py 16 300 > gpurun_out/ncu_kp.log 2>&1
VOD_KP_DBG=7 ncu --set full --clock-control none --import-source on -k regex:"tafa_keyproj" -s 3 -c 1 -o gpurun_out/prof_kp_p7 -f python scripts/probe_keyproj.py 16 300 >> gpurun_out/ncu_kp.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_kp.log; ls -la gpurun_out/prof_kp*
