mkdir -p gpurun_out
YQ_NET=tiny YQ_WARM=2 YQ_NO_PROFILE_FORWARD=1 timeout 900 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"maxpool|route_rows" -s 6 -c 3 -o gpurun_out/pool -f python tools/prof_forward.py > gpurun_out/ncu_pool.log 2>&1
tail -2 gpurun_out/ncu_pool.log | cut -c1-200
