mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_yolov3.py -q -m gpu --timeout 600 -x 2>&1 | tail -30 ) > gpurun_out/t_yolov3.log 2>&1
( timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -30 ) > gpurun_out/t_parity.log 2>&1
timeout 600 python bench.py --net yolov3 --steps 20 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
timeout 300 python bench.py --steps 100 --warmup 5 --streams 1 --no-cpu-baseline > gpurun_out/bench_tiny_s1.json 2> gpurun_out/bench_tiny.err
tail -12 gpurun_out/t_yolov3.log; tail -6 gpurun_out/t_parity.log
for f in bench_v3 bench_tiny_s1; do cut -c1-220 gpurun_out/$f.json; done; tail -3 gpurun_out/bench_v3.err
