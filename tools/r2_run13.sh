mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
( timeout 600 python -m pytest tests/test_gpu_dp.py tests/test_gpu_input.py -q -m gpu 2>&1 | tail -15 ) > gpurun_out/t_dp.log 2>&1
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from yolo_quantization_b200 import synth
L = synth.yolov3_tiny_quant()
synth.write_cfg("/tmp/tiny.cfg", L, batch=128)
synth.write_weights("/tmp/tiny.weights", L)
PY
( timeout 300 tools/dp_demo /tmp/tiny.cfg /tmp/tiny.weights 128 100 ) > gpurun_out/dp_demo.json 2> gpurun_out/dp_demo.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 --host-mem wc > gpurun_out/bench_2gpu_wc.json 2>> gpurun_out/bench_2gpu.err
cat gpurun_out/t_dp.log; cat gpurun_out/dp_demo.json; tail -3 gpurun_out/dp_demo.err; cut -c1-900 gpurun_out/bench_2gpu.json; echo; python - <<'PY'
import json
for f in ("bench_2gpu","bench_2gpu_wc"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), d["e2e"])
PY
head -20 gpurun_out/topo.txt
