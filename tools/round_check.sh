mkdir -p gpurun_out
( timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/tests.log 2>&1
timeout 200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 120 python bench.py --steps 200 --warmup 10 --streams 1 --no-cpu-baseline > gpurun_out/bench_s1.json 2>> gpurun_out/bench.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 51 -c 85 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline --streams 1 > gpurun_out/ncu_bench.log 2>&1
( timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > gpurun_out/smoke.log
cat gpurun_out/tests.log gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench.json; cut -c1-200 gpurun_out/bench_s1.json; tail -2 gpurun_out/bench.err; wc -l gpurun_out/launches.csv
