# A/B of bench.py's device-resident arm under schedule switches: "UPROUTE BRANCH_STREAM STREAMS" per line of $CASES
mkdir -p gpurun_out
: > gpurun_out/ab.log
echo "${CASES:=1 1 1;1 1 2;1 1 3;1 1 1;1 1 2}" | tr ';' '\n' | while read u b s; do
  YQ_UPROUTE=$u YQ_BRANCH_STREAM=$b timeout 120 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --streams $s 2>gpurun_out/ab.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$u $b $s', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'], d['clocks']['sm_mhz'])
" >> gpurun_out/ab.log 2>&1
done
cat gpurun_out/ab.log; tail -3 gpurun_out/ab.err
