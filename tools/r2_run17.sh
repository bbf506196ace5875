mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -12 ) > gpurun_out/t_all.log 2>&1
timeout 600 python bench.py --net yolov3 --steps 20 --warmup 3 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
timeout 600 python bench.py --net yolov3 --steps 20 --warmup 3 --streams 2 --no-cpu-baseline --no-extras > gpurun_out/bench_v3_s2.json 2> /dev/null
timeout 300 python bench.py --steps 100 --warmup 5 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/bench_tiny_s1.json 2> gpurun_out/bench_tiny.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_tiny.json 2>> gpurun_out/bench_tiny.err
tail -8 gpurun_out/t_all.log
for f in bench_v3 bench_v3_s2 bench_tiny_s1 bench_tiny; do cut -c1-200 gpurun_out/$f.json; done; tail -3 gpurun_out/bench_v3.err
