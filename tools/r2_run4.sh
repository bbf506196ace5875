mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_yolov3.py -q -m gpu --timeout 600 2>&1 | tail -40 ) > gpurun_out/t_yolov3.log 2>&1
( timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/t_parity.log 2>&1
timeout 600 python bench.py --net yolov3 --steps 20 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
YQ_NET=yolov3 YQ_WARM=0 YQ_NO_PROFILE_FORWARD=1 timeout 900 ncu --metrics $M --clock-control none -c 110 --csv --log-file gpurun_out/v3_metrics.csv python tools/prof_forward.py > gpurun_out/ncu_v3.log 2>&1
tail -30 gpurun_out/t_yolov3.log; tail -4 gpurun_out/t_parity.log
for f in bench_v3 bench_tiny; do cut -c1-220 gpurun_out/$f.json; done; tail -3 gpurun_out/bench_v3.err; tail -3 gpurun_out/ncu_v3.log; wc -l gpurun_out/v3_metrics.csv
