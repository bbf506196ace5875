# tools/round_check.sh -- the command set behind profiles/r2_* (run under gpurun on one B200)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
( timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 ) > gpurun_out/tests.log 2>&1
( timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > gpurun_out/smoke.log
( tools/probes/probe_i8_peak ) > gpurun_out/i8_peak.json 2> gpurun_out/i8_peak.err
timeout 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 200 python bench.py --steps 200 --warmup 10 --streams 1 --no-cpu-baseline > gpurun_out/bench_s1.json 2>> gpurun_out/bench.err
timeout 400 python bench.py --net yolov3 --steps 50 --warmup 5 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
# launch lists of one forward (WARM forwards first so that lazily built tensor maps exist), DRAM bytes per launch
for NET in tiny yolov3; do
  YQ_NET=$NET YQ_WARM=0 YQ_NO_PROFILE_FORWARD=1 timeout 900 ncu --metrics $M --clock-control none -c 130 --csv --log-file gpurun_out/launches_$NET.csv python tools/prof_forward.py > gpurun_out/prof_$NET.log 2>&1
  B=$( [ $NET = tiny ] && echo 128 || echo 64 )
  python tools/make_traffic.py gpurun_out/launches_$NET.csv gpurun_out/prof_$NET.log $NET $B gpurun_out/r2_traffic_$NET.json
done
# kernel timelines of the forward in steady state (graph replays, one stream; CUPTI records through torch.profiler)
for NET in tiny yolov3; do YQ_NET=$NET timeout 300 python tools/timeline.py gpurun_out/r2_timeline_$NET.json > gpurun_out/r2_timeline_$NET.txt 2>&1; done
cat gpurun_out/tests.log gpurun_out/smoke.log gpurun_out/i8_peak.json; cut -c1-300 gpurun_out/bench.json; cut -c1-200 gpurun_out/bench_s1.json; cut -c1-300 gpurun_out/bench_v3.json; tail -n 2 gpurun_out/bench.err; tail -n 2 gpurun_out/bench_v3.err
# evidence for profiles/ (copy after the LAST kernel change: bench.py trusts r2_traffic_*.json only for the same kernel sources)
python tools/ncu_summary.py gpurun_out/launches_tiny.csv /dev/null > /dev/null 2>&1 || true
