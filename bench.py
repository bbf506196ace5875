#!/usr/bin/env python
"""bench.py -- images/sec of yolov3-tiny INT8 416x416 on B200 (the metric BASELINE.json names).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = one forward pass of the quantized hot path (input layout transform, 13 fused conv
layers, pools, upsample, routes, yolo heads) over one batch of B=128 synthetic images per GPU
(BASELINE.json configs[2]; configs[3] is the same 128 images/GPU on 8 GPUs -> weak scaling).

  value  device-resident throughput: inputs already in HBM, CUDA-graph replay, CUDA events on the
         networks' streams, max over ranks.  --streams S (default 2) keeps S forwards in flight per GPU, each on
         its own network instance (own activation tensors and stream); K steps = K forwards in all, timed from
         one start mark every stream waits for to the last stream's end mark.  --streams 1: one after the other.
  e2e    the same metric through the public host-buffer API (yq_network_submit_u8 / yq_network_collect, the
         2-deep pipelined form of network_predict): every step does the H2D of its uint8 batch from pinned
         memory, the forward and the D2H of both yolo heads inside the timed region.
  roofline      the dominant kernel launch (per-layer CUDA events measured live after the timed loop).
  cpu_baseline  the UNMODIFIED reference compiled from /root/reference (oracle/_ref) timed on this box's
                host cores on a bounded sample.

--impl reference times the reference's own CPU implementation (OpenMP build, all host threads).
Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec yolov3-tiny INT8 416x416"
UNIT = "images/s"
MACS_PER_IMAGE = 2_724_074_496
NOMINAL_INT8_TOPS = 4500.0
# --net yolov3: BASELINE configs[4] (full yolov3, 75 conv layers, INT8 per-channel, batch 64, HBM GB/s vs peak)
NETS = {
    "tiny": {"metric": METRIC, "macs": MACS_PER_IMAGE, "batch": 128,
             "what": "yolov3-tiny INT8 per-channel (24 layers, relu6, 5 classes)", "configs": "BASELINE configs[2]; x{w} GPUs = configs[3] sharding"},
    "yolov3": {"metric": "images/sec yolov3 INT8 416x416", "macs": 32_932_037_632, "batch": 64,
               "what": "full yolov3 INT8 per-channel (107 layers: 75 conv, 23 quantized shortcut, leaky, 80 classes)",
               "configs": "BASELINE configs[4]"},
}


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update(hbm_gbs=float(m["hbm_gbs"]), bf16_tflops=float(m["bf16_tflops"]),
                 bf16_tflops_sustained=float(m.get("bf16_tflops_sustained", m["bf16_tflops"])), source="measured")
    except Exception:
        pass
    return p


def source_hash():
    """sha256 over the kernel sources: profiles/*_traffic.json is only trusted for the build it was captured on"""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "yolo_quantization_b200", "csrc", "*.cu")) + glob.glob(os.path.join(ROOT, "yolo_quantization_b200", "csrc", "*.cuh")) +
                    glob.glob(os.path.join(ROOT, "yolo_quantization_b200", "csrc", "*.h"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def int8_peak(pk):
    """the chip's tcgen05 kind::i8 rate: measured live by tools/probes/probe_i8_peak (all SMs, N = 256 MMAs back to back, CUDA events),
    else the figure committed under profiles/, else 2 x the measured bf16 rate"""
    probe = os.path.join(ROOT, "tools", "probes", "probe_i8_peak")
    try:
        out = subprocess.run([probe], capture_output=True, text=True, timeout=60).stdout
        d = json.loads(out.strip().splitlines()[-1])
        return float(d["int8_tops_cta1"]), "measured live: tools/probes/probe_i8_peak (tcgen05.mma kind::i8, M=128 N=256 K=32, all SMs)", d
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "r2_i8_peak.json")) as f:
            d = json.load(f)
        return float(d["int8_tops_cta1"]), "profiles/r2_i8_peak.json (tools/probes/probe_i8_peak on this pool's B200)", d
    except Exception:
        return 2.0 * pk["bf16_tflops_sustained"], "2 x sustained bf16 of MEASURED_PEAKS.json (probe unavailable)", None


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 50 ms while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.samples:
            if ts < t0 or ts > t1 + 0.25:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def layer_work(li, batch, conv_written=True):
    """Algorithmic ops / bytes of one layer launch (SURVEY Appendix B2 accounting: activations in + out, weights once).
    A conv launch that fuses the following 2x2 max-pool is charged its input plus what it WRITES: the pooled tensor,
    and the conv tensor only when that is materialised too (SURVEY 8(d) 'fully fused' figure)."""
    t = li.type
    if t == 0:
        macs = li.out_h * li.out_w * li.n * li.c * li.size * li.size
        out = li.out_h * li.out_w * li.n * (5 if li.quant_stop_flag else 1)
        if li.fused in (1, 2):
            out = (li.out_h // 2) * (li.out_w // 2) * li.n + (out if conv_written else 0)
        byts = batch * (li.h * li.w * li.c + out) + li.c * li.n * li.size ** 2
        return 2 * macs * batch, byts
    if t == 4:
        return 0, 2 * 4 * batch * li.out_c * li.out_h * li.out_w
    if t == 5:      # quantized shortcut: two tensors read, one written
        return 0, 3 * batch * li.out_h * li.out_w * li.out_c
    return 0, batch * (li.h * li.w * li.c + li.out_h * li.out_w * li.out_c)


def run_reference(args):
    """--impl reference: the reference's CPU QUANTIZATION=1 path (oracle/_ref, OpenMP, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import yq_oracle as O
    from yolo_quantization_b200 import synth
    cores = os.cpu_count() or 1
    layers = synth.yolov3_tiny_quant()
    with tempfile.TemporaryDirectory() as d:
        cfg, wts, img = os.path.join(d, "t.cfg"), os.path.join(d, "t.weights"), os.path.join(d, "img.f32")
        synth.write_cfg(cfg, layers, batch=1)
        info = synth.write_weights(wts, layers)
        im = synth.synthetic_image(1)
        synth.image_to_float(im).tofile(img)
        n = args.warmup + args.steps
        if os.path.exists(O.REF_HARNESS_OMP):
            kind = "reference"
            err = O.run_reference("time", cfg, wts, img, str(n), omp=True, threads=cores)
            times = O.ref_times(err)
        else:
            kind = "port"
            O.build()
            times = []
            for _ in range(n):
                t0 = time.perf_counter()
                O.forward_network(info, im)
                times.append(time.perf_counter() - t0)
    timed = times[args.warmup:] or times[-1:]
    total = sum(timed)
    val = len(timed) / total
    sample = (f"{len(timed)} steps of 1 image (batch 1, network_predict window, examples/detector.c:922-924), "
              f"{args.warmup} warm-up; OMP_NUM_THREADS={cores}; stdout of the reference's per-layer printf -> /dev/null")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "yolov3-tiny INT8 per-channel 416x416, reference CPU QUANTIZATION=1 path, 1 image per step"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "cpu_model": cpu_model()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def cpu_baseline(cfg1, wts, img_f32, info, im, net="tiny"):
    from oracle import yq_oracle as O
    cores = os.cpu_count() or 1
    iters = 8
    if net != "tiny":
        # the reference has no quantized shortcut, so it cannot run this network: the oracle port (OpenMP over output channels) is timed
        times = []
        for _ in range(2):
            t0 = time.perf_counter()
            O.forward_network(info, im)
            times.append(time.perf_counter() - t0)
        return {"value": 1.0 / min(times), "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"2 images at batch 1 (best), full yolov3 416x416, oracle port (exact-integer C restatement, OpenMP, {cores} threads); "
                          "the reference itself cannot run a quantized shortcut"}
    if os.path.exists(O.REF_HARNESS_OMP):
        times = O.ref_times(O.run_reference("time", cfg1, wts, img_f32, str(iters), omp=True, threads=cores))
        kind = "reference"
    else:
        kind, times = "port", []
        for _ in range(iters):
            t0 = time.perf_counter()
            O.forward_network(info, im)
            times.append(time.perf_counter() - t0)
    times = times[1:]
    res = {"value": 1.0 / statistics.median(times), "unit": UNIT, "cores": cores, "kind": kind, "cpu_model": cpu_model(),
           "sample": f"{len(times)} images at batch 1 (median network_predict time, 1 warm-up), yolov3-tiny 416x416, "
                     f"{'oracle/_ref OpenMP build' if kind == 'reference' else 'oracle port'}, {cores} threads"}
    if kind == "reference" and os.path.exists(O.REF_HARNESS):
        # the reference's default build (MULTI_CORE=0), one thread
        t1 = O.ref_times(O.run_reference("time", cfg1, wts, img_f32, "3"))[1:]
        res["single_thread"] = {"value": 1.0 / statistics.median(t1), "unit": UNIT, "cores": 1,
                                "sample": f"{len(t1)} images, MULTI_CORE=0 build (the reference's default), 1 warm-up"}
    return res


_REAL_STDOUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner at the first collective, the
    oracle's make): from here on fd 1 points at stderr and emit() writes the line to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: 128 for --net tiny, 64 for --net yolov3)")
    ap.add_argument("--net", default="tiny", choices=sorted(NETS), help="tiny = the headline metric (BASELINE configs[2]); yolov3 = configs[4]")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", type=int, default=-1, help="-1 auto, 0 SIMT only, 1 tcgen05 where available")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--streams", type=int, default=2,
                    help="device-resident arm: forwards in flight per GPU (each on its own network instance and stream)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--host-mem", default="pinned", choices=["pinned", "wc"],
                    help="e2e arm: input batches in plain page-locked memory or write-combined page-locked memory (yq_host_alloc)")
    ap.add_argument("--no-extras", action="store_true", help="skip the batch-1 latency, H2D ceiling and int8 peak probe legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    NET = NETS[args.net]
    args.batch = args.batch or NET["batch"]
    _claim_stdout()
    if args.impl == "reference":
        if args.net != "tiny":
            # the reference has no quantized shortcut (src/shortcut_layer.c:62-67 is float only): it cannot run this network
            emit({"impl": "reference", "unavailable": "the reference cannot run the full yolov3 in its QUANTIZATION=1 path (no quantized shortcut)"})
            return 0
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from yolo_quantization_b200 import darknet, dp, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path is CUDA-only (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B = args.batch
    layers = synth.yolov3_tiny_quant() if args.net == "tiny" else synth.yolov3_quant()
    tmp = tempfile.TemporaryDirectory()
    cfg = os.path.join(tmp.name, "tiny.cfg")
    cfg1 = os.path.join(tmp.name, "tiny_b1.cfg")
    wts = os.path.join(tmp.name, "tiny.weights")
    synth.write_cfg(cfg, layers, batch=B)
    synth.write_cfg(cfg1, layers, batch=1)
    info = None
    # weights: rank 0 synthesises the .weights stream; the replicas receive it by ONE NCCL broadcast
    if rank == 0:
        info = synth.write_weights(wts, layers)
    dp.broadcast_weights_file(wts, rank, world, torch.device("cuda", local), dist if world > 1 else None)
    net = darknet.load_network(cfg, wts, batch=B, device=local)
    if args.kernel >= 0:
        net.set_conv_kernel(args.kernel)
    net.use_graph(not args.no_graph)
    stream = torch.cuda.ExternalStream(net.stream, device=local)
    # --streams S > 1: S network instances (own activation tensors, own stream, same weights file) take the batches in turn,
    # so the thin tail layers of one forward share the SMs with the wide first layers of the next
    nets = [net]
    # (forced flavours run one forward at a time: the multicast-cluster per-tap kernel of --kernel 1 is not sized for sharing
    # its SMs' TMEM columns with another launch's clusters)
    for _ in range(1, max(1, args.streams if args.kernel < 0 else 1)):
        other = darknet.load_network(cfg, wts, batch=B, device=local)
        if args.kernel >= 0:
            other.set_conv_kernel(args.kernel)
        other.use_graph(not args.no_graph)
        nets.append(other)
    streams = [torch.cuda.ExternalStream(n_.stream, device=local) for n_ in nets]
    S = len(nets)

    # inputs: R distinct batches so consecutive steps never re-read the same input from L2
    R = 4
    rng = np.random.default_rng(1000 + rank)
    host = torch.empty((R, B, 3, 416, 416), dtype=torch.uint8).pin_memory()
    host.numpy()[...] = rng.integers(0, 256, size=host.shape, dtype=np.uint8)
    dev = host.cuda(non_blocking=False)
    host_ptrs = [host[i].data_ptr() for i in range(R)]
    wc_block = None
    if args.host_mem == "wc":      # the e2e arm's input batches in write-combined page-locked memory (DMA reads skip the CPU caches)
        import ctypes as C
        from yolo_quantization_b200 import _lib
        wc_block = _lib.load().yq_host_alloc(host.numel(), 1)
        if not wc_block:
            raise SystemExit("yq_host_alloc failed: " + _lib.last_error())
        C.memmove(wc_block, host.data_ptr(), host.numel())
        host_ptrs = [wc_block + i * B * 3 * 416 * 416 for i in range(R)]
    out_host = torch.empty(net.output_floats, dtype=torch.float32).pin_memory()
    in_bytes = B * 3 * 416 * 416
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    # ---- device-resident throughput ------------------------------------------------------------
    for i in range(args.warmup * S):
        nets[i % S].forward_device(dev[i % R].data_ptr())
    for n_ in nets:
        n_.synchronize()
    barrier()
    t_start = time.time()
    e0 = torch.cuda.Event(enable_timing=True)
    e1s = [torch.cuda.Event(enable_timing=True) for _ in nets]
    e0.record(stream)
    for k in range(1, S):
        streams[k].wait_event(e0)   # every stream starts behind the one start mark
    for i in range(args.steps):
        nets[i % S].forward_device(dev[i % R].data_ptr())
    for k in range(S):
        e1s[k].record(streams[k])
    for n_ in nets:
        n_.synchronize()
    barrier()
    ms = max(e0.elapsed_time(e) for e in e1s)
    # ---- end to end through the host-buffer API (2-deep submit/collect pipeline) -----------------
    out_hosts = [out_host, torch.empty_like(out_host).pin_memory()]

    def e2e_loop(n):
        slots = []
        for i in range(n):
            slots.append((net.submit_raw(host_ptrs[i % R]), i))
            if len(slots) == 2:
                sl, j = slots.pop(0)
                net.collect_raw(sl, out_hosts[j % 2].data_ptr())
        for sl, j in slots:
            net.collect_raw(sl, out_hosts[j % 2].data_ptr())

    e2e_loop(3)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_e2e0 = time.perf_counter()
    e2.record(stream)
    e2e_loop(args.steps)
    e3.record(stream)
    net.synchronize()
    t_e2e1 = time.perf_counter()
    barrier()
    # the last D2H completes on the host after the compute stream's last event: take the longer of the two clocks
    ms_e2e = max(e2.elapsed_time(e3), 1e3 * (t_e2e1 - t_e2e0))
    t_end = time.time()
    # ---- extras: the box's H2D ceiling for this batch, batch-1 latency -------------------------------------------------
    h2d_gbs = None
    batch1 = None
    if not args.no_extras:
        ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cstream = torch.cuda.Stream(device=local)
        from yolo_quantization_b200 import _lib as _l
        lib_ = _l.load()
        with torch.cuda.stream(cstream):
            lib_.yq_cuda_push(dev[0].data_ptr(), host_ptrs[0], in_bytes, cstream.cuda_stream)
            ce0.record(cstream)
            for i in range(10):
                lib_.yq_cuda_push(dev[i % R].data_ptr(), host_ptrs[i % R], in_bytes, cstream.cuda_stream)
            ce1.record(cstream)
        cstream.synchronize()
        h2d_gbs = 10 * in_bytes / (ce0.elapsed_time(ce1) * 1e-3) / 1e9
        if world == 1 and args.net == "tiny":
            cfgb1 = os.path.join(tmp.name, "b1.cfg")
            synth.write_cfg(cfgb1, layers, batch=1)
            n1 = darknet.load_network(cfgb1, wts, batch=1, device=local)
            n1.use_graph(True)
            s1 = torch.cuda.ExternalStream(n1.stream, device=local)
            # a ring of 8 single-image inputs: the forward is one CUDA graph per input pointer (captured on first use, at most 16 kept),
            # so a serving loop rotates over a few device buffers -- 200 DIFFERENT pointers would time 200 captures, not 200 forwards
            ring = [dev[r][k].data_ptr() for r in range(min(R, 4)) for k in range(2)]
            for i in range(3 * len(ring)):
                n1.forward_device(ring[i % len(ring)])
            n1.synchronize()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record(s1)
            for i in range(200):
                n1.forward_device(ring[i % len(ring)])
            b1.record(s1)
            n1.synchronize()
            lat = b0.elapsed_time(b1) / 200
            batch1 = {"latency_ms": lat, "images_per_s": 1e3 / lat,
                      "what": "BASELINE configs[1]: batch 1, device-resident input (ring of 8 images), CUDA-graph replay, 200 forwards back to back on one stream"}
            n1.free()
    # ---- per-layer CUDA events (un-graphed forwards on the same stream) -------------------------
    prof_iters = max(3, min(20, args.steps))
    lm = np.zeros(net.n + 1, np.float64)
    for i in range(prof_iters):
        lm += net.profile_forward(dev[i % R].data_ptr())
    lm /= prof_iters
    clocks = sampler.stop(t_start, time.time())

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        pk = peaks()
        if args.no_extras:
            p_int8, p_src, p_raw = 2.0 * pk["bf16_tflops_sustained"], "2 x sustained bf16 of MEASURED_PEAKS.json (--no-extras)", None
        else:
            p_int8, p_src, p_raw = int8_peak(pk)
        infos = net.layers()
        rows = []
        for i, li in enumerate(infos):
            ops, byts = layer_work(li, B, conv_written=False)   # production plan: fused convs write the pooled tensor only
            if li.type == 2 and li.fused == 5:
                # a route the next conv reads in two parts: its launch (if any) only brings the upsampled input into shape
                up = infos[i - 1] if i > 0 and infos[i - 1].type == 3 and infos[i - 1].fused else None
                byts = B * (up.h * up.w * up.c + up.out_h * up.out_w * up.out_c) if up else 0
            t_ms = float(lm[i + 1])
            if t_ms <= 0 or (li.type in (1, 3, 4, 5) and li.fused):    # a fused-away max-pool / upsample / yolo / shortcut layer has no launch of its own
                continue
            t_tc = ops / (p_int8 * 1e12) * 1e3
            t_mem = byts / (pk["hbm_gbs"] * 1e9) * 1e3
            bound = "tensor" if t_tc > t_mem else "hbm"
            rows.append({"layer": i, "type": darknet.LAYER_TYPES[li.type], "ms": round(t_ms, 4), "bound": bound,
                         "frac": round(max(t_tc, t_mem) / t_ms, 4), "kernel": int(li.kernel) if li.type == 0 else None,
                         "fused": int(li.fused),
                         "ops": ops, "bytes": byts})
        # the NCHW -> NHWC(4) layout transform in front of layer 0 (lm[0]): in + out bytes
        c0 = infos[0]
        tb = B * c0.h * c0.w * (c0.c + darknet.channel_stride(c0.c))
        if float(lm[0]) > 0.005:     # (not launched when layer 0 reads the CHW planes itself: the interval is two back-to-back events)
            rows.insert(0, {"layer": -1, "type": "nchw_to_nhwc", "ms": round(float(lm[0]), 4), "bound": "hbm",
                            "frac": round(tb / (pk["hbm_gbs"] * 1e9) * 1e3 / float(lm[0]), 4), "kernel": None, "fused": 0, "ops": 0, "bytes": tb})
        top = max(rows, key=lambda r: r["ms"])
        if top["bound"] == "tensor":
            ach = top["ops"] / (top["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": p_int8, "unit": "TFLOP/s", "frac": ach / p_int8}
        else:
            ach = top["bytes"] / (top["ms"] * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"]}
        # dram__bytes_read + dram__bytes_write of the same launch from the ncu capture tools/round_check.sh made -- trusted only for
        # the build it was captured on (hash of the kernel sources) and for this net / batch
        traffic, traffic_note = None, "no capture for this build"
        try:
            with open(os.path.join(ROOT, "profiles", f"r2_traffic_{args.net}.json")) as f:
                tj = json.load(f)
            if tj.get("source_hash") == source_hash() and tj.get("batch") == B:
                traffic = tj["layers"].get(str(top["layer"]))
                traffic_note = f"profiles/r2_traffic_{args.net}.json (ncu --set full, same kernel sources {tj['source_hash']})"
            else:
                traffic_note = f"profiles/r2_traffic_{args.net}.json is from other kernel sources ({tj.get('source_hash')} != {source_hash()}) or another batch: not used"
        except Exception:
            pass
        roof.update(traffic=traffic, traffic_source=traffic_note, kernel=f"layer {top['layer']} ({top['type']}, flavour {top['kernel']})",
                    share_of_step=top["ms"] / float(lm.sum()),
                    peak_source=(f"HBM {pk['hbm_gbs']} GB/s ({pk['source']} MEASURED_PEAKS.json); int8 tensor peak {p_int8:.0f} TOP/s = {p_src} "
                                 f"(2 x sustained bf16 would be {2 * pk['bf16_tflops_sustained']:.0f}, nominal dense int8 {NOMINAL_INT8_TOPS:.0f}); "
                                 f"'TFLOP/s' counts int8 ops"))
        ips = world * B * args.steps / (ms * 1e-3)
        ips_e2e = world * B * args.steps / (ms_e2e * 1e-3)
        total_ops = 2 * NET["macs"]
        step_bytes = sum(r["bytes"] for r in rows)
        line = {
            "metric": NET["metric"], "value": ips, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{NET['what']}, 416x416, batch {B} per GPU ({NET['configs'].format(w=world)})",
                       "global_batch": B * world, "parallelism": f"dp{world}", "cuda_graph": not args.no_graph, "streams": S,
                       "l2": f"{R} rotating input batches ({R * in_bytes >> 20} MiB) > 126 MB L2 and several hundred MiB of "
                             f"activations written per step; no explicit flush"},
            "e2e": {"value": ips_e2e, "unit": UNIT, "h2d_bytes_per_step": in_bytes * world,
                    "d2h_bytes_per_step": int(net.output_floats) * 4 * world, "ms_per_step": ms_e2e / args.steps, "host_memory": args.host_mem,
                    # what this rank's link delivers for one batch-sized pinned copy on its own (10 back to back): the ceiling of e2e
                    "h2d_ceiling_gbs_rank0": h2d_gbs,
                    "h2d_achieved_gbs_per_gpu": in_bytes / (ms_e2e / args.steps * 1e-3) / 1e9,
                    "frac_of_h2d_ceiling": (in_bytes / (ms_e2e / args.steps * 1e-3) / 1e9 / h2d_gbs) if h2d_gbs else None},
            "gpu_launches": net.launches_per_forward * args.steps * world,
            "clocks": clocks,
            "roofline": roof,
            # algorithmic activation + weight bytes of every launch of one step (unfused accounting per launch) over the step time
            "hbm_gbs_algorithmic_whole_net": step_bytes / (ms / args.steps * 1e-3) / 1e9,
            "hbm_frac_whole_net": step_bytes / (ms / args.steps * 1e-3) / 1e9 / pk["hbm_gbs"],
            "int8_tops_measured_peak": p_int8, "int8_peak_probe": p_raw,
            "batch1": batch1,
            "int8_tops_whole_net": ips / world * total_ops / 1e12,
            "frac_int8_peak_whole_net": ips / world * total_ops / 1e12 / p_int8,
            "layers": [{k: r[k] for k in ("layer", "type", "ms", "bound", "frac", "kernel", "fused")} for r in rows],
        }
        if world == 1 and not args.no_cpu_baseline:
            im = synth.synthetic_image(1)
            img_f32 = os.path.join(tmp.name, "img.f32")
            synth.image_to_float(im).tofile(img_f32)
            line["cpu_baseline"] = cpu_baseline(cfg1, wts, img_f32, info, im, args.net)
        emit(line)
    for n_ in nets:
        n_.free()
    if wc_block:
        from yolo_quantization_b200 import _lib
        _lib.load().yq_host_free(wc_block)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
