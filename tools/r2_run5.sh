mkdir -p gpurun_out
YQ_NET=yolov3 YQ_WARM=0 YQ_NO_PROFILE_FORWARD=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_u8_tc_flat2_kernel -s 5 -c 2 -o gpurun_out/v3_flat2 -f python tools/prof_forward.py > gpurun_out/ncu_v3_full.log 2>&1
tail -3 gpurun_out/ncu_v3_full.log; ls -la gpurun_out/*.ncu-rep
