mkdir -p gpurun_out
for V in 1 0; do
YQ_NO_PWT=$((1-V)) python bench.py --net yolov3 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ab_v3_pwt$V.json 2>gpurun_out/ab.err
done
tail -3 gpurun_out/ab.err
