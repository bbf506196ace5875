mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "route or upsample or schedule_switches or fused_network or graph_replay or pipelined" 2>&1 | tail -5 ) > gpurun_out/t1.log 2>&1
for c in "0 0" "1 0" "0 1" "1 1" "0 0" "1 1"; do set -- $c
  YQ_UPROUTE=$1 YQ_BRANCH_STREAM=$2 timeout 90 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1 $2', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'], d['clocks']['sm_mhz'])
" >> gpurun_out/ab.log 2>&1
done
cat gpurun_out/t1.log gpurun_out/ab.log
