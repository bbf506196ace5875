mkdir -p gpurun_out
( timeout 300 python -m pytest "tests/test_gpu_yolov3.py::test_full_yolov3_conv_shapes_at_real_sizes" -x -q -m gpu 2>&1 | tail -60 ) > gpurun_out/t_real.log 2>&1
( timeout 600 python tools/dbg_v3.py 416 1 2>&1 | tail -40 ) > gpurun_out/dbg416.log
( timeout 600 python tools/dbg_v3.py 416 3 2>&1 | tail -40 ) > gpurun_out/dbg416b3.log
head -70 gpurun_out/t_real.log; cat gpurun_out/dbg416.log gpurun_out/dbg416b3.log
