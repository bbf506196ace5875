# A/B of bench.py's device-resident arm under schedule switches: "UPROUTE BRANCH_STREAM EARLY_ROUTE STREAMS" per entry of $CASES
mkdir -p gpurun_out
: > gpurun_out/ab.log
if [ -n "$AB_TESTS" ]; then ( timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$AB_TESTS" 2>&1 | tail -5 ) > gpurun_out/t1.log 2>&1; cat gpurun_out/t1.log; fi
echo "${CASES:=1 1 0 1;1 1 1 1;1 1 0 2;1 1 1 2;1 1 0 1;1 1 1 1;1 1 1 2}" | tr ';' '\n' | while read u b e s; do
  YQ_UPROUTE=$u YQ_BRANCH_STREAM=$b YQ_EARLY_ROUTE=$e timeout 120 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --streams $s 2>gpurun_out/ab.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$u $b $e $s', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'], d['clocks']['sm_mhz'])
" >> gpurun_out/ab.log 2>&1
done
cat gpurun_out/ab.log; tail -3 gpurun_out/ab.err
