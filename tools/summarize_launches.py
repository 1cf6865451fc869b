"""ncu `--metrics gpu__time_duration.sum --csv` launch list -> markdown table per kernel family.
usage: python tools/summarize_launches.py launches.csv "<command that was profiled>" > profiles/rNN_launches.md"""
import csv, re, sys
from collections import defaultdict

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= iv or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)     # -> us
    if "spin_kernel" in r[ik]:          # torch.cuda._sleep of bench.py's attribution pass, not part of a step
        continue
    name = re.sub(r"<unnamed>::", "", r[ik])
    name = re.sub(r"<.*", "", name)
    name = re.sub(r"\(.*", "", name).strip()
    if name.startswith("void "):
        name = name[5:]
    name = name[:70]
    agg[name][0] += 1
    agg[name][1] += v
tot_n = sum(a[0] for a in agg.values())
tot_t = sum(a[1] for a in agg.values())
print(f"Command: `{cmd}`\n")
print("(per-launch times under ncu are cold-cache and serialised -- compare SHARES only; the headline `value` is measured "
      "from CUDA-graph replay, never under a profiler)\n")
print(f"{tot_n} launches, {tot_t / 1e3:.2f} ms of kernel time\n")
print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if t / tot_t < 0.002:
        continue
    print(f"| `{name}` | {n} | {t:.1f} | {100 * t / tot_t:.1f}% | {t / n:.1f} |")
