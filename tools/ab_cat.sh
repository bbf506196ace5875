# A/B of the route-elimination switches on the tiny network: YQ_NO_CAT (materialised route), YQ_NO_UP2 (route's own launch upsamples layer 18)
mkdir -p gpurun_out
for V in "0 0" "0 1" "1 1"; do
set -- $V
YQ_NO_CAT=$1 YQ_NO_UP2=$2 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/ab_tiny_nocat$1_noup$2.json 2>gpurun_out/ab.err
YQ_NO_CAT=$1 YQ_NO_UP2=$2 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras --streams 1 > gpurun_out/ab_tiny1_nocat$1_noup$2.json 2>>gpurun_out/ab.err
done
tail -3 gpurun_out/ab.err
