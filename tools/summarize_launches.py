"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel:
    python tools/summarize_launches.py gpurun_out/launches.csv "header comment" > profiles/launches_summary.csv"""
import csv
import re
import sys

OURS = re.compile(r"\bk_[a-z0-9_]+")
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows:
    name = r[ki]
    m = OURS.search(name)
    t = re.search(r"<[^>]*>", name[m.end():m.end() + 24]) if m else None
    short = (m.group(0) + (t.group(0) if t and name[m.end()] == "<" else "")) if m else re.sub(r"^void ", "", name)[:80]
    ns = float(r[vi].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[ui], 1)
    a = agg.setdefault(short, [0.0, 0, bool(m)])
    a[0] += ns
    a[1] += 1
tot = sum(a[0] for a in agg.values())
ours = sum(a[0] for a in agg.values() if a[2])
for c in sys.argv[2:]:
    print("# " + c)
print(f"# total {tot / 1e6:.2f} ms over {sum(a[1] for a in agg.values())} launches; libst_b200 kernels: {ours / tot * 100:.1f}% of device time")
print("share_pct,total_us,launches,us_per_launch,kernel,ours")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{a[0] / tot * 100:.2f},{a[0] / 1e3:.1f},{a[1]},{a[0] / 1e3 / a[1]:.2f},{k},{'yes' if a[2] else 'no'}")
