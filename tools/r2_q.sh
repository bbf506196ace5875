mkdir -p gpurun_out
for V in "YQ_X=1" "YQ_ROWS_TWO=0" "YQ_ROWS_TWO=0 YQ_ROWS_DB=1" "YQ_ROWS_TWO=0 YQ_ROWS_SPLIT=1"; do
env $V timeout 300 python bench.py --steps 200 --warmup 10 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - "$V" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"]), d["ms_per_step"], [ (r["layer"], r["ms"]) for r in d["layers"][:4]])
PY
done
