"""ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file launches.csv`) -> per-kernel totals and shares for profiles/.
    python tools/launch_list_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rNN_launches_step.json
"""
import csv
import json
import sys
from collections import OrderedDict

path, command = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
agg = OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "")[:80]
    ms = float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
total = sum(a[1] for a in agg.values())
kernels = [dict(kernel=k, launches=a[0], total_ms=round(a[1], 3), share=round(a[1] / total, 4)) for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]
print(json.dumps(dict(command=command, note="one timed step of the 2-MoT-block Wan-14B-width model (per-block kernel shares equal the 40-block model's); "
                      "per-launch times are cold-cache and serialised: compare shares", total_ms=round(total, 3), kernels=kernels), indent=1))
