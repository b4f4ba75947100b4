#!/usr/bin/env python
"""Aggregate an ncu per-launch metrics CSV of one denoising step (gpu__time_duration, dram bytes, lts bytes, tensor pipe)
per kernel; optionally write the GEMM family's DRAM traffic as JSON for bench.py's roofline.traffic."""
import collections
import csv
import json
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    src = sys.argv[1]
    out_json = sys.argv[2] if len(sys.argv) > 2 else None
    rows = list(csv.DictReader([l for l in open(src) if l.startswith('"')]))
    per = collections.defaultdict(dict)
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        u, m = r["Metric Unit"], r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        elif u in UNIT:
            v *= UNIT[u]
        per[r["ID"]][m] = v
        per[r["ID"]]["k"] = r["Kernel Name"].split("(")[0][:44]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    for d in per.values():
        a = agg[d["k"]]
        t = d.get("gpu__time_duration.sum", 0.0)
        a[0] += 1
        a[1] += t
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
        a[4] += d.get("lts__t_bytes.sum", 0.0)
        a[5] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * t
    print(f"{'kernel':46s} {'n':>4s} {'us':>9s} {'dramR MB':>9s} {'dramW MB':>9s} {'L2 MB':>9s} {'tensor%':>7s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:46s} {a[0]:4d} {a[1]:9.1f} {a[2] / 1e6:9.1f} {a[3] / 1e6:9.1f} {a[4] / 1e6:9.1f} {a[5] / max(a[1], 1e-9):7.1f}")
    if out_json:
        g = [a for k, a in agg.items() if "gemm_tcgen05" in k]
        n = sum(a[0] for a in g)
        dram = sum(a[2] + a[3] for a in g)
        json.dump({"source": src, "kernel": "gemm_tcgen05_kernel (all BN)", "launches": n, "dram_bytes_per_step": dram,
                   "dram_bytes_per_launch": dram / n, "l2_bytes_per_step": sum(a[4] for a in g),
                   "note": "ncu replays each launch cold-cache: activations that are L2 hits inside the real step are "
                           "counted as DRAM reads here"}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
