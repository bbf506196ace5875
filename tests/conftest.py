import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu() -> bool:
    try:
        from yolo_quantization_b200 import _lib
        return _lib.load().yq_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must not silently pass: the tests run and fail loudly.
    pass


@pytest.fixture(scope="session")
def built():
    """Build the shared library and the oracle once per session (idempotent make)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def tiny_net_files(tmp_path_factory):
    """Seeded yolov3-tiny cfg/weights (416x416) + the SynthLayer truth."""
    from yolo_quantization_b200 import synth
    d = tmp_path_factory.mktemp("tiny")
    layers = synth.yolov3_tiny_quant()
    cfg, wts = str(d / "tiny.cfg"), str(d / "tiny.weights")
    synth.write_cfg(cfg, layers, batch=1)
    info = synth.write_weights(wts, layers)
    return cfg, wts, info, layers
