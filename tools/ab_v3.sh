mkdir -p gpurun_out
for NC in 0 1; do
YQ_NO_CAT=$NC python bench.py --net yolov3 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ab_v3_nocat$NC.json 2>gpurun_out/ab.err
done
tail -3 gpurun_out/ab.err
