mkdir -p gpurun_out
for V in "YQ_L0_GROUPS=2" "YQ_L0_GROUPS=3 YQ_L0_BULK=0" "YQ_L0_GROUPS=2 YQ_L0_BULK=0" "YQ_L0_GROUPS=2" "YQ_L0_GROUPS=3 YQ_L0_BULK=0"; do
env $V timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - "$V" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"]), d["ms_per_step"], [ (r["layer"], r["ms"]) for r in d["layers"][:2]])
PY
done
