"""Summarises an ncu launch list of bench.py (gpu__time_duration.sum [, dram__bytes_read.sum, dram__bytes_write.sum]):
per-kernel totals of one full step, the per-layer table of the conv-only replay (the last 75 conv launches), and
profiles/<round>/traffic.json (DRAM bytes of those conv launches) that bench.py's roofline.traffic reads.

    python tools/summarize_launches.py gpurun_out/launches.csv profiles/round1 r1h
"""
import csv
import json
import os
import sys
from collections import OrderedDict

src, outdir, tag = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(lines))
by_id = OrderedDict()
for r in rows:
    e = by_id.setdefault(r["ID"], dict(name=r["Kernel Name"]))
    e[r["Metric Name"]] = float(r["Metric Value"])
launches = list(by_id.values())
names = [e["name"] for e in launches]
first = [i for i, n in enumerate(names) if "conv_first_tc_kernel" in n]
# conv-only replay = last conv_first launch up to the end; one full step = between the two conv_first launches before it
a, b = first[-1], len(launches)
conv = [e for e in launches[a:b] if "conv_" in e["name"] and "pack" not in e["name"]]
unit = {r["Metric Name"]: r["Metric Unit"] for r in rows}
scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3}[unit["gpu__time_duration.sum"]]
bscale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = []
# a full step: the last stretch between two first-conv launches that also holds the NMS kernels
step = []
for lo, hi in zip(first[:-1], first[1:]):
    if any("nms_select" in e["name"] for e in launches[lo:hi]):
        step = launches[lo:hi]
agg = OrderedDict()
for e in step:
    k = e["name"].split("(")[0][-64:]
    v = agg.setdefault(k, [0, 0.0])
    v[0] += 1
    v[1] += e["gpu__time_duration.sum"] * scale
out.append("one full step (forward graph + post-processing), kernel durations under ncu (cold caches, serialised):")
for k, v in agg.items():
    out.append("  %-66s n=%3d  us=%9.1f" % (k, v[0], v[1]))
out.append("  total %.1f us in %d launches" % (sum(v[1] for v in agg.values()), len(step)))
tot = sum(e["gpu__time_duration.sum"] * scale for e in conv)
out.append("conv-only replay: %d launches, %.1f us" % (len(conv), tot))
if "dram__bytes_read.sum" in unit:
    rd = sum(e["dram__bytes_read.sum"] * bscale[unit["dram__bytes_read.sum"]] for e in conv)
    wr = sum(e["dram__bytes_write.sum"] * bscale[unit["dram__bytes_write.sum"]] for e in conv)
    out.append("  DRAM read %.1f MB, write %.1f MB" % (rd / 1e6, wr / 1e6))
    with open(os.path.join(outdir, "traffic.json"), "w") as fh:
        json.dump(dict(conv_launches=len(conv), dram_read_bytes=rd, dram_write_bytes=wr,
                       sum_kernel_us_under_ncu=tot,
                       source="profiles/%s/launches_%s.csv: last conv-only graph replay of bench.py --steps 2 (%d launches), "
                              "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
                              % (os.path.basename(outdir), tag, len(conv))), fh, indent=1)
with open(os.path.join(outdir, "launches_%s_summary.txt" % tag), "w") as fh:
    fh.write("\n".join(out) + "\n")
print("\n".join(out))
