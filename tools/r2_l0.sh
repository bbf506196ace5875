mkdir -p gpurun_out
for V in "YQ_L0_GROUPS=2" "YQ_L0_GROUPS=3" "YQ_L0_GROUPS=1"; do
( env $V timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "nchw or network or tiny or full_size" 2>&1 | tail -4 ) > gpurun_out/t_l0.log 2>&1
echo "$V"; tail -2 gpurun_out/t_l0.log
done
for V in "YQ_L0_TSTORE=0" "YQ_L0_TSTORE=1" "YQ_L0_TSTORE=1 YQ_L0_GROUPS=3" "YQ_L0_TSTORE=0" "YQ_L0_TSTORE=1"; do
  env $V timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/bench_l0.json 2> gpurun_out/bench_l0.err
  python - "$V" <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/bench_l0.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"]), d["ms_per_step"], [ (r["layer"], r["ms"]) for r in d["layers"][:4]])
PY
done
