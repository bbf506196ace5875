"""ncu launch list of ONE forward (csv: --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum[,...]) + the
LAUNCH_ORDER line tools/prof_forward.py printed -> profiles/r2_traffic_<net>.json: measured DRAM bytes per layer launch, keyed by a
hash of the kernel sources so that bench.py only trusts it for the build it was captured on.

    python tools/make_traffic.py <ncu.csv> <prof_forward.log> <net> <batch> <out.json> [first_launch_index]
"""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (source_hash)

src, log, net, batch, dst = sys.argv[1:6]
first = int(sys.argv[6]) if len(sys.argv) > 6 else 0
order = None
for line in open(log):
    if line.startswith("LAUNCH_ORDER"):
        order = json.loads(line.split(" ", 1)[1])
assert order is not None, "no LAUNCH_ORDER line in " + log
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
per = {}
names = {}
for r in rows[hi + 1:]:
    if len(r) < len(rows[hi]):
        continue
    k = int(r[ix["ID"]])
    per.setdefault(k, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    names[k] = r[ix["Kernel Name"]].replace("void <unnamed>::", "").split("(")[0]
ids = sorted(per)[first:first + len(order)]
assert len(ids) == len(order), f"{len(ids)} launches in the capture, {len(order)} in one forward"
layers, kernels, times = {}, {}, {}
for k, layer in zip(ids, order):
    m = per[k]
    layers[str(layer)] = layers.get(str(layer), 0) + m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
    times[str(layer)] = times.get(str(layer), 0) + m["gpu__time_duration.sum"] / 1e3
    kernels[str(layer)] = names[k]
out = {"net": net, "batch": int(batch), "source_hash": bench.source_hash(), "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
       "how": "ncu --clock-control none, one forward, per-launch (cold cache, serialised)", "layers": layers, "kernel": kernels, "ncu_us": times}
json.dump(out, open(dst, "w"), indent=1)
print(dst, len(order), "launches", sum(layers.values()) / 1e6, "MB DRAM per forward")
