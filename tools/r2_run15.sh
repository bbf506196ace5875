mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_yolov3.py -q -m gpu --timeout 600 2>&1 | tail -30 ) > gpurun_out/t_v3.log 2>&1
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_input.py tests/test_gpu_dropin.py tests/test_gpu_dp.py -q -m gpu --timeout 600 -x 2>&1 | tail -8 ) > gpurun_out/t_rest.log 2>&1
timeout 600 python bench.py --net yolov3 --steps 20 --warmup 3 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
YQ_NO_FUSE_SHORTCUT=1 timeout 600 python bench.py --net yolov3 --steps 20 --warmup 3 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/bench_v3_nofuse.json 2> /dev/null
YQ_FLAT2X_WIDE=0 timeout 600 python bench.py --net yolov3 --steps 20 --warmup 3 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/bench_v3_nowide.json 2> /dev/null
YQ_FLAT2X_WIDE=0 timeout 300 python bench.py --steps 100 --warmup 5 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/bench_tiny_nowide.json 2> /dev/null
timeout 300 python bench.py --steps 100 --warmup 5 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err
tail -25 gpurun_out/t_v3.log; tail -5 gpurun_out/t_rest.log
for f in bench_v3 bench_v3_nofuse bench_v3_nowide bench_tiny_nowide bench_tiny; do cut -c1-200 gpurun_out/$f.json; done; tail -3 gpurun_out/bench_v3.err
