mkdir -p gpurun_out
for K in 0 1 2 4 8 3 6 5 7 15; do
  YQ_L0_KNOBS=$K timeout 300 python bench.py --steps 50 --warmup 5 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/bench_l0k.json 2> gpurun_out/bench_l0.err
  python - "$K" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_l0k.json").read().strip().splitlines()[-1])
print("knobs", sys.argv[1], [ (r["layer"], r["ms"]) for r in d["layers"][:2]])
PY
done
