mkdir -p gpurun_out
for NC in 0 1; do
YQ_NO_CAT=$NC python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/ab_tiny_nocat$NC.json 2>gpurun_out/ab.err
YQ_NO_CAT=$NC python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras --streams 1 > gpurun_out/ab_tiny1_nocat$NC.json 2>>gpurun_out/ab.err
done
tail -3 gpurun_out/ab.err
