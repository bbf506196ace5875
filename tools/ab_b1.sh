# batch-1 latency (BASELINE configs[1]) with the round's switches on / off
for V in "1 0 0" "0 0 0" "1 1 1" ; do
set -- $V
YQ_PW=$1 YQ_NO_CAT=$2 YQ_NO_UP2=$3 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PW NO_CAT NO_UP2 = $V', 'batch1', d['batch1']['latency_ms'], 'value', round(d['value']))"
done
