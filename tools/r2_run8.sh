mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_input.py tests/test_gpu_dropin.py -q -m gpu --timeout 600 2>&1 | tail -40 ) > gpurun_out/t_new.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_yolov3.py tests/test_gpu_parity.py -q -m gpu --timeout 600 -x 2>&1 | tail -15 ) > gpurun_out/t_all.log 2>&1
( timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > gpurun_out/smoke.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err
tail -40 gpurun_out/t_new.log; tail -6 gpurun_out/t_all.log; cat gpurun_out/smoke.log; cut -c1-200 gpurun_out/bench_tiny.json
