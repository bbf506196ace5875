# ncu --set full of every launch of one forward (yolov3-tiny, batch 128) and of the first launches of the full yolov3 (batch 64),
# reduced on the box to the key-metric tables committed under profiles/ (the reports themselves are ~50 MB each)
mkdir -p gpurun_out
YQ_NET=tiny YQ_WARM=1 YQ_NO_PROFILE_FORWARD=1 timeout 1500 ncu --set full --clock-control none -s 17 -c 17 -o /tmp/r2_forward_full -f python tools/prof_forward.py > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/r2_forward_full.ncu-rep gpurun_out/r2_forward_full.csv
YQ_NET=yolov3 YQ_WARM=1 YQ_NO_PROFILE_FORWARD=1 timeout 1500 ncu --set full --clock-control none -s 78 -c 16 -o /tmp/r2_yolov3_head_full -f python tools/prof_forward.py > gpurun_out/ncu_full_v3.log 2>&1
python tools/ncu_summary.py /tmp/r2_yolov3_head_full.ncu-rep gpurun_out/r2_yolov3_head_full.csv
