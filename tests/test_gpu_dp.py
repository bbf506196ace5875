"""GPU: the C-level data-parallel entry (yq_dp_*): replicas on every visible GPU (1 on the single-GPU test box, 2+
under `gpurun --gpus N`), ONE ncclBroadcast of the packed arena, per-device results equal the single-replica results."""
import ctypes as C

import numpy as np
import pytest

from yolo_quantization_b200 import _lib, darknet, synth

pytestmark = pytest.mark.gpu


def test_dp_load_and_predict(built, tmp_path):
    lib = _lib.load()
    ndev = min(lib.yq_device_count(), 4)
    layers = synth.yolov3_tiny_quant()
    cfg, wts = str(tmp_path / "t.cfg"), str(tmp_path / "t.weights")
    synth.write_cfg(cfg, layers, batch=2, width=96, height=96)
    synth.write_weights(wts, layers, width=96, height=96)
    devs = (C.c_int * ndev)(*range(ndev))
    dp = lib.yq_dp_load_network(cfg.encode(), wts.encode(), 2, devs, ndev)
    assert dp, _lib.last_error()
    assert lib.yq_dp_num_devices(dp) == ndev
    if ndev > 1:
        # the replicas took their big filter images out of the broadcast blob, device to device
        assert lib.yq_dp_arena_bytes(dp) > 8_000_000 and lib.yq_dp_images_from_arena(dp) >= 13 * (ndev - 1)
    x = np.stack([synth.synthetic_image(100 + i, 3, 96, 96) for i in range(2 * ndev)])
    nout = lib.yq_network_output_floats(lib.yq_dp_replica(dp, 0))
    out = np.empty((ndev, nout), np.float32)
    assert lib.yq_dp_network_predict_u8(dp, x.ctypes.data, out.ctypes.data) == 0, _lib.last_error()
    lib.yq_dp_free_network(dp)
    one = darknet.load_network(cfg, wts, batch=2)
    for i in range(ndev):
        assert np.array_equal(one.predict_u8(x[2 * i:2 * i + 2]), out[i]), f"device {i}"
    one.free()


def test_dp_errors(built, tmp_path):
    lib = _lib.load()
    devs = (C.c_int * 1)(0)
    assert not lib.yq_dp_load_network(b"/nonexistent.cfg", b"/nonexistent.weights", 1, devs, 1) and "cannot open" in _lib.last_error()
    many = (C.c_int * 64)(*range(64))
    assert not lib.yq_dp_load_network(b"x", b"y", 1, many, 64) and "devices requested" in _lib.last_error()
