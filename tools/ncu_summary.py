"""`ncu -i X.ncu-rep --page raw --csv` -> markdown table (one row per captured launch) and, with --json, the per-kernel mean
DRAM traffic that bench.py reports as roofline.traffic.
usage: python tools/ncu_summary.py raw.csv [--json profiles/ncu_traffic.json] > profiles/rNN_ncu_fused.md"""
import csv
import json
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = list(csv.reader(open(path, errors="replace")))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
WANT = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
units = rows[1]


def val(r, name):
    if name not in idx:
        return None
    try:
        v = float(r[idx[name]].replace(",", ""))
    except ValueError:
        return None
    u = units[idx[name]]
    if name.startswith("gpu__time"):
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    if name.startswith("dram__bytes"):
        v = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0) * v / 1e6
    return v


print("| kernel | " + " | ".join(w[1] for w in WANT) + " | DRAM GB/s |\n|---|" + "---:|" * (len(WANT) + 1))
traffic = defaultdict(list)
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("gb200::enc::", "").replace("gb200::", "").strip()[:48]
    vals = [val(r, w[0]) for w in WANT]
    t, rd, wr = vals[0], vals[1], vals[2]
    gbs = (rd + wr) * 1e6 / (t * 1e-6) / 1e9 if t and rd is not None and wr is not None else None
    print(f"| `{name}` | " + " | ".join("" if v is None else (f"{v:.1f}" if abs(v) < 1e5 else f"{v:.3g}") for v in vals) +
          f" | {'' if gbs is None else f'{gbs:.0f}'} |")
    if rd is not None and wr is not None:
        traffic[name].append((rd + wr) * 1e6)
if "--json" in sys.argv:
    out = {k: dict(dram_bytes_per_launch=sum(v) / len(v), launches_captured=len(v),
                   note="mean of dram__bytes_read.sum + dram__bytes_write.sum over the launches of this kernel in the "
                        "`ncu --set full` capture of tools/run_layer_once.py (tools/final_profile.sh)")
           for k, v in traffic.items()}
    with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
        json.dump(out, f, indent=1)
