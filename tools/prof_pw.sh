# ncu launch list (key metrics) of the pointwise-flavour launches of one forward of each network, plus one full capture
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed
for NET in tiny yolov3; do
  YQ_NET=$NET YQ_WARM=1 YQ_NO_PROFILE_FORWARD=1 timeout 600 ncu --metrics $M --clock-control none -k regex:conv_u8_tc_pw --csv --log-file gpurun_out/pw_$NET.csv python tools/prof_forward.py > gpurun_out/pw_prof_$NET.log 2>&1
done
YQ_NET=tiny YQ_WARM=1 YQ_NO_PROFILE_FORWARD=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_u8_tc_pw -s 3 -c 3 -o gpurun_out/pw_tiny_full python tools/prof_forward.py > gpurun_out/pw_full.log 2>&1
ls -la gpurun_out/pw_*
