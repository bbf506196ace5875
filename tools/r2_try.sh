mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 ) > gpurun_out/tests.log 2>&1; tail -3 gpurun_out/tests.log
timeout 300 python bench.py --steps 200 --warmup 10 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/bench_try_s1.json 2> gpurun_out/bench_try.err
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/bench_try.json 2>> gpurun_out/bench_try.err
timeout 300 python bench.py --net yolov3 --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_try_v3.json 2>> gpurun_out/bench_try.err
python - <<'PY'
import json
for f in ("bench_try_s1","bench_try","bench_try_v3"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), d["ms_per_step"], [ (r["layer"], r["ms"]) for r in d["layers"][:18]])
PY
tail -3 gpurun_out/bench_try.err
