mkdir -p gpurun_out
YQ_NET=tiny YQ_WARM=1 YQ_NO_PROFILE_FORWARD=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_u8_tc_l0 -s 1 -c 1 -o gpurun_out/l0 -f python tools/prof_forward.py > gpurun_out/ncu_l0.log 2>&1
tail -3 gpurun_out/ncu_l0.log | cut -c1-200; ls -la gpurun_out/l0.ncu-rep
