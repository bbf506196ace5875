# compute-sanitizer evidence (profiles/r2_sanitizer_*.log): memcheck on smoke() (debug + production plan of yolov3-tiny at 96x96) and
# on the full-yolov3 network test (shortcut fusion, stride-2 per-tap, flat2 1x1 / wide rows), racecheck on smoke()
mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 ) > gpurun_out/r2_sanitizer_memcheck_smoke.log
( timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "tests/test_gpu_yolov3.py::test_yolov3_network_96_vs_oracle" "tests/test_gpu_yolov3.py::test_conv_with_fused_shortcut_vs_oracle" "tests/test_gpu_input.py" "tests/test_gpu_dp.py" -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2_sanitizer_memcheck_yolov3.log
( timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "pointwise or two_tensors or fused_network" 2>&1 | tail -15 ) > gpurun_out/r2_sanitizer_memcheck_pointwise.log
( timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_yolov3.py -q -m gpu -k "patch_mode or network_416" 2>&1 | tail -15 ) > gpurun_out/r2_sanitizer_memcheck_patch.log
( timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -25 ) > gpurun_out/r2_sanitizer_racecheck_smoke.log
for f in gpurun_out/r2_sanitizer_*.log; do echo "== $f"; tail -n 4 $f; done
