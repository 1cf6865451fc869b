"""`ncu -i X.ncu-rep --page raw --csv` -> per launch: duration, throughputs, occupancy and the dominant warp-stall reasons.
usage: python tools/ncu_stalls.py raw.csv"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
if not stall:
    stall = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")]
KEYS = [("gpu__time_duration.sum", "ns"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("l1tex__data_bank_conflicts_pipe_lsu.sum", "bankconf"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def f(r, k):
    try:
        return float(r[idx[k]].replace(",", ""))
    except (KeyError, ValueError):
        return None


for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("gb200::enc::", "").replace("gb200::", "").strip()[:44]
    vals = " ".join(f"{lab}={f(r, k):.4g}" for k, lab in KEYS if f(r, k) is not None)
    st = sorted(((f(r, h) or 0.0, h) for h in stall), reverse=True)[:4]
    sts = ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0].replace('.ratio', '')}={v:.2f}" for v, h in st)
    print(f"{name}\n    {vals}\n    stalls: {sts}")
