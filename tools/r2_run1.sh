# round 2, GPU call 1: new yolov3 tests, regression suite, int8 peak probe, first yolov3 bench line, route_rows A/B
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_yolov3.py -q -m gpu --timeout 600 2>&1 | tail -60 ) > gpurun_out/t_yolov3.log 2>&1
( timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/t_parity.log 2>&1
( timeout 60 tools/probes/probe_i8_peak ) > gpurun_out/i8_peak.json 2> gpurun_out/i8_peak.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err
timeout 600 python bench.py --net yolov3 --steps 20 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
YQ_ROUTE_ROWS=0 timeout 200 python bench.py --steps 200 --warmup 5 --streams 1 --no-cpu-baseline > gpurun_out/bench_rr0.json 2>/dev/null
YQ_ROUTE_ROWS=1 timeout 200 python bench.py --steps 200 --warmup 5 --streams 1 --no-cpu-baseline > gpurun_out/bench_rr1.json 2>/dev/null
tail -25 gpurun_out/t_yolov3.log; tail -5 gpurun_out/t_parity.log; cat gpurun_out/i8_peak.json gpurun_out/i8_peak.err
cut -c1-300 gpurun_out/bench_tiny.json; tail -3 gpurun_out/bench_tiny.err
cut -c1-300 gpurun_out/bench_v3.json; tail -5 gpurun_out/bench_v3.err
cut -c1-200 gpurun_out/bench_rr0.json; cut -c1-200 gpurun_out/bench_rr1.json
