mkdir -p gpurun_out
YQ_NET=yolov3 YQ_WARM=0 YQ_NO_PROFILE_FORWARD=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_u8_tc_flat2_kernel -s 6 -c 1 -o gpurun_out/v3_flat2b -f python tools/prof_forward.py > gpurun_out/ncu_v3_full.log 2>&1
tail -2 gpurun_out/ncu_v3_full.log | cut -c1-200; ls -la gpurun_out/v3_flat2b.ncu-rep
