mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 ) > gpurun_out/tests.log 2>&1; tail -3 gpurun_out/tests.log
for V in "YQ_FLAT2_PCQ=0" "YQ_FLAT2_PCQ=1"; do
env $V timeout 300 python bench.py --net yolov3 --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_pcq.json 2> gpurun_out/bench_pcq.err
python - "$V" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_pcq.json").read().strip().splitlines()[-1])
print(sys.argv[1], "yolov3", round(d["value"]), d["ms_per_step"], [ (r["layer"], r["ms"]) for r in d["layers"] if r["layer"] in (2,3,6,7,13,14,38,39,63,64)])
PY
env $V timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/bench_pcq.json 2>> gpurun_out/bench_pcq.err
python - "$V" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_pcq.json").read().strip().splitlines()[-1])
print(sys.argv[1], "tiny", round(d["value"]), d["ms_per_step"], [ (r["layer"], r["ms"]) for r in d["layers"] if r["layer"] in (8,)])
PY
done
tail -2 gpurun_out/bench_pcq.err
