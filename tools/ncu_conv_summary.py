"""Summarise an `ncu --set full` capture of the gather-conv kernels.  Here (no GPU needed):
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_conv_summary.py /tmp/raw.csv "header comment" > profiles/ncu_conv_rX.csv
Also writes profiles/ncu_traffic.json ({kernel<template>: mean DRAM bytes per launch}) next to the csv when
--traffic-json PATH is given; bench.py reports it as roofline.traffic."""
import csv
import json
import re
import sys

COLS = [("time_us", "gpu__time_duration.sum"), ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue_pct", "sm__inst_issued.avg.pct_of_peak_sustained_active"), ("warps_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("l1_hit", "l1tex__t_sector_hit_rate.pct"), ("l2_hit", "lts__t_sector_hit_rate.pct"), ("regs", "launch__registers_per_thread"),
        ("smem_KB", "launch__shared_mem_per_block_dynamic"), ("lsu_wavefronts_M", "l1tex__data_pipe_lsu_wavefronts.sum")]
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tj = None
    if "--traffic-json" in sys.argv:
        tj = sys.argv[sys.argv.index("--traffic-json") + 1]
        args = [a for a in args if a != tj]
    rows = list(csv.reader(l for l in open(args[0]) if l.startswith('"')))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for c in args[1:]:
        print("# " + c)
    print("kernel,grid," + ",".join(n for n, _ in COLS))
    traffic = {}
    for r in rows:
        name = r[idx["Kernel Name"]]
        m = re.search(r"k_[a-z0-9_]+(<[^>]*>)?", name)
        short = m.group(0).replace(" ", "").replace(",", "x") if m else name[:40].replace(",", ";")
        vals = []
        for n, metric in COLS:
            if metric not in idx or r[idx[metric]] in ("", "n/a"):
                vals.append("")
                continue
            v = float(r[idx[metric]].replace(",", ""))
            u = units[idx[metric]]
            if n == "time_us" or n.endswith("_MB"):
                v *= UNIT.get(u, 1.0)
            elif n == "smem_KB":
                v *= {"byte": 1 / 1024, "Kbyte": 1.0}.get(u, 1.0)
            elif n.endswith("_M"):
                v *= 1e-6
            vals.append(f"{v:.3f}")
        grid = r[idx["Grid Size"]].strip("()").split(",")[0] if "Grid Size" in idx else ""
        print(f"{short},{grid}," + ",".join(vals))
        if vals[1] and vals[2]:
            traffic.setdefault(short, []).append((float(vals[1]) + float(vals[2])) * 1e6)
    if tj:
        json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(tj, "w"), indent=1)


if __name__ == "__main__":
    main()
