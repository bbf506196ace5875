mkdir -p gpurun_out
for PW in 1 0; do
YQ_PW=$PW python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/ab_tiny_pw$PW.json 2>gpurun_out/ab.err
YQ_PW=$PW python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras --streams 1 > gpurun_out/ab_tiny1_pw$PW.json 2>>gpurun_out/ab.err
YQ_PW=$PW python bench.py --net yolov3 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ab_v3_pw$PW.json 2>>gpurun_out/ab.err
done
tail -3 gpurun_out/ab.err
