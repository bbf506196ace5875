mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 ) > gpurun_out/tests.log 2>&1; tail -4 gpurun_out/tests.log
timeout 300 python bench.py --net yolov3 --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_c32_v3.json 2> gpurun_out/bench_c32.err
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/bench_c32_tiny.json 2>> gpurun_out/bench_c32.err
YQ_POOL_LEAN=0 timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/bench_c32_tiny_nolean.json 2>> gpurun_out/bench_c32.err
python - <<'PY'
import json
for f in ("bench_c32_v3","bench_c32_tiny","bench_c32_tiny_nolean"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), d["ms_per_step"], [ (r["layer"], r["ms"], r["kernel"], r["fused"]) for r in d["layers"][:11]])
PY
tail -3 gpurun_out/bench_c32.err
