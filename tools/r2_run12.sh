mkdir -p gpurun_out
timeout 300 python bench.py --steps 200 --warmup 5 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/ab_base.json 2>/dev/null
YQ_FLAT2_1X1=1 timeout 300 python bench.py --steps 200 --warmup 5 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/ab_1x1.json 2>/dev/null
YQ_FLAT2X=2 timeout 300 python bench.py --steps 200 --warmup 5 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/ab_2x.json 2>/dev/null
timeout 300 python bench.py --steps 200 --warmup 5 --streams 2 --no-cpu-baseline > gpurun_out/ab_s2.json 2>gpurun_out/ab_s2.err
timeout 300 python bench.py --steps 200 --warmup 5 --streams 3 --no-cpu-baseline --no-extras > gpurun_out/ab_s3.json 2>/dev/null
timeout 300 python bench.py --steps 200 --warmup 5 --streams 2 --no-cpu-baseline --host-mem wc > gpurun_out/ab_wc.json 2>/dev/null
for f in ab_base ab_1x1 ab_2x ab_s2 ab_s3 ab_wc; do python - <<PY
import json
d=json.load(open("gpurun_out/$f.json"))
print("$f", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["e2e"].get("h2d_ceiling_gbs_rank0"), d.get("batch1"), d.get("int8_tops_measured_peak"), [(r["layer"], r["ms"]) for r in d["layers"] if r["layer"] in (13,15,18,20,22)])
PY
done; tail -3 gpurun_out/ab_s2.err
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
YQ_NET=yolov3 YQ_WARM=0 YQ_NO_PROFILE_FORWARD=1 timeout 900 ncu --metrics $M --clock-control none -c 110 --csv --log-file gpurun_out/v3_metrics2.csv python tools/prof_forward.py > gpurun_out/ncu_v3.log 2>&1
tail -2 gpurun_out/ncu_v3.log | cut -c1-300
