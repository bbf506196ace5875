"""CPU: the C-ABI library loads, exports every symbol include/yq_b200.h declares, and fails loudly (no
fallback) when no CUDA device is present.  No compute is called here."""
import ctypes
import os
import re

import numpy as np
import pytest

from yolo_quantization_b200 import _lib, darknet, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "yq_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"YQ_API[^;(]*?\b(yq_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 45
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/yq_b200.h but not exported by libyq_b200.so"
    # and the ctypes table binds exactly the declared surface
    assert sorted(_lib.SIGNATURES) == syms


def test_library_has_sm100a_code_only(built):
    out = os.popen(f"cuobjdump --list-elf {_lib.LIB_PATH} 2>/dev/null").read()
    if not out:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_struct_layouts_match_header(built):
    # yq_conv_desc: 12 ints/float (48 B) + 5 pointers + int (+pad) ; yq_layer_info: 22 4-byte fields; yq_act_geom: 3 ints
    assert ctypes.sizeof(_lib.ConvDesc) == 48 + 5 * 8 + 8
    assert ctypes.sizeof(_lib.LayerInfo) == 22 * 4
    assert ctypes.sizeof(_lib.ActGeom) == 12


def test_channel_stride(built):
    lib = _lib.load()
    assert [lib.yq_channel_stride(c) for c in (1, 3, 4, 5, 16, 30, 384, 1024)] == [4, 4, 4, 16, 16, 32, 384, 1024]


def test_no_gpu_fails_loudly(built, tiny_net_files):
    lib = _lib.load()
    if lib.yq_device_count() > 0:
        pytest.skip("a GPU is present")
    cfg, wts, _, _ = tiny_net_files
    with pytest.raises(_lib.YqError, match="no CUDA device"):
        darknet.load_network(cfg, wts)
    with pytest.raises(_lib.YqError, match="no CUDA device"):
        darknet.DeviceBuffer(16)
    with pytest.raises(_lib.YqError):
        darknet.ConvolutionalLayerQuant(4, 4, 4, 16, 1, 1, 0, 3, np.zeros(64, np.uint8), np.zeros(16, np.uint8),
                                        np.zeros(16, np.int32), np.ones(16), np.ones(16), 0, 0, 1.0)


def test_product_never_imports_oracle():
    """the product path may not import, call or link anything under oracle/."""
    pkg = os.path.join(ROOT, "yolo_quantization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|oracle/|liboracle|yq_oracle", txt), \
                    f"{f} reaches into oracle/"


def test_synthetic_weights_stream_layout(tiny_net_files):
    cfg, wts, info, layers = tiny_net_files
    assert os.path.getsize(wts) == 43_431_034
    convs = [l for l in info if l.kind == "conv"]
    assert len(convs) == 13 and sum(int(np.prod(l.w_u8.shape)) for l in convs) == 8_672_688
    macs = sum(l.out_h * l.out_w * l.out_c * l.c * l.spec.size ** 2 for l in convs)
    assert macs == 2_724_074_496
    txt = open(cfg).read()
    assert txt.count("[convolutional]") == 13 and txt.count("[maxpool]") == 6 and txt.count("[yolo]") == 2


def test_pack_arena_file_errors_and_empty_round_trip(built, tmp_path):
    """yq_pack.cu host logic (no GPU): a missing or foreign file is an error with a message, an empty arena round-trips."""
    from yolo_quantization_b200 import _lib
    lib = _lib.load()
    assert lib.yq_pack_arena_clear() == 0
    assert lib.yq_pack_arena_load(str(tmp_path / "missing.yqpk").encode()) < 0 and "cannot open" in _lib.last_error()
    junk = tmp_path / "junk.yqpk"
    junk.write_bytes(b"not an arena at all")
    assert lib.yq_pack_arena_load(str(junk).encode()) < 0 and "not a packed-weight arena" in _lib.last_error()
    empty = tmp_path / "empty.yqpk"
    assert lib.yq_pack_arena_save(str(empty).encode()) == 0
    assert empty.read_bytes()[:4] == b"YQPK" and lib.yq_pack_arena_load(str(empty).encode()) == 0


def test_pack_arena_drops_damaged_entries(built, tmp_path):
    """yq_pack.cu: every entry carries an FNV-1a checksum of its data; an entry whose data no longer matches is dropped at load
    (a miss: the image is rebuilt), the others are kept; collecting is off until enabled."""
    import struct
    from yolo_quantization_b200 import _lib
    lib = _lib.load()

    def fnv(data, h=14695981039346656037):
        for b in data:
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h

    good, bad = bytes(range(64)), bytes(range(64, 128))
    blob = struct.pack("<III", 0x4B505159, 5, 2)
    for key, tag, data, damage in ((1, b"ohwi.128", good, False), (2, b"rows", bad, True)):
        blob += struct.pack("<Q", key) + tag.ljust(24, b"\0") + struct.pack("<QQ", len(data), fnv(data))
        blob += (data[:-1] + bytes([data[-1] ^ 1])) if damage else data
    path = tmp_path / "a.yqpk"
    path.write_bytes(blob)
    assert lib.yq_pack_arena_clear() == 0
    assert lib.yq_pack_arena_load(str(path).encode()) == 1             # one of two entries survives
    e, h, m, d = (ctypes.c_int() for _ in range(4))
    assert lib.yq_pack_arena_stats(ctypes.byref(e), ctypes.byref(h), ctypes.byref(m), ctypes.byref(d)) == 0
    assert e.value == 1 and d.value == 1                               # dirty: the file gets rewritten without the damaged entry
    assert lib.yq_pack_arena_save(str(path).encode()) == 1
    assert lib.yq_pack_arena_clear() == 0 and lib.yq_pack_arena_load(str(path).encode()) == 1
    assert lib.yq_pack_arena_clear() == 0


def test_leaky_divide_by_ten_constant():
    """yq_epilogue.cuh: LEAKY's round(q / 10) for q < 0 is (h * 52429 + 262145) >> 19 with h = |q| -- floor((h + 5) / 10) -- exact for
    every h the epilogue lets through (make_epi lowers xlim so that h <= 81914), and the product stays inside 32 bits."""
    import numpy as np
    h = np.arange(0, 81915, dtype=np.uint64)
    assert np.array_equal((h * 52429 + 262145) >> 19, (h + 5) // 10)
    assert 81914 * 52429 + 262145 < 2 ** 32
    # ... and it is the reference's double arithmetic: round-half-away of -h * 0.1 (convolutional_layer.c:737)
    q = -h.astype(np.int64)
    ref = np.where(q * 0.1 - np.trunc(q * 0.1) <= -0.5, np.trunc(q * 0.1) - 1, np.trunc(q * 0.1)).astype(np.int64)
    assert np.array_equal(-((h.astype(np.int64) + 5) // 10), ref)
