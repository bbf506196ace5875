mkdir -p gpurun_out
python bench.py --net yolov3 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ab_v3_new.json 2>gpurun_out/ab.err
tail -3 gpurun_out/ab.err
