mkdir -p gpurun_out
for V in 0 1; do
YQ_FLAT2X_1X1=$V python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extras --streams 1 > gpurun_out/ab_tiny1_f2x$V.json 2>gpurun_out/ab.err
YQ_FLAT2X_1X1=$V python bench.py --net yolov3 --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ab_v3_f2x$V.json 2>>gpurun_out/ab.err
done
tail -3 gpurun_out/ab.err
