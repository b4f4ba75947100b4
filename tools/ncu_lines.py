#!/usr/bin/env python
"""Join the per-SASS-instruction samples of an ncu report (`ncu -i rep --page source --csv`) with nvdisasm -gi line
info of the matching cubin and aggregate stall samples per CUDA source line.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep uni_renderer_b200/build/gemm_sm100.o 'gemm_tcgen05_kernelILi160'
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(obj, func_pat):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    out, cur, active = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            active = re.search(func_pat, ln) is not None
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            # attribute inlined helpers to the call site in our own file
            cur = (m.group(3), int(m.group(4))) if m.group(3) else (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2))
    return out


def main():
    rep, obj, pat = sys.argv[1:4]
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    csv_txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(csv_txt.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    base = int(data[0][ix["Address"]], 16)
    lines = sass_lines(obj, pat)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: [0.0, collections.Counter(), 0.0])
    total = 0.0
    for r in data:
        off = int(r[ix["Address"]], 16) - base
        n = float(r[ix["# Samples"]] or 0)
        key = lines.get(off, (None, ""))[0]
        a = agg[key]
        a[0] += n
        a[2] += float(r[ix["Instructions Executed"]] or 0)
        total += n
        for h in stall_cols:
            v = float(r[ix[h]] or 0)
            if v:
                a[1][h[6:]] += v
    src_cache = {}
    print(f"total samples {total:.0f}")
    for key, (n, st, ie) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        text = ""
        if key and key[0] and os.path.exists(key[0]):
            src_cache.setdefault(key[0], open(key[0]).read().splitlines())
            text = src_cache[key[0]][key[1] - 1].strip()[:80]
        where = f"{os.path.basename(key[0])}:{key[1]}" if key and key[0] else "?"
        top = ", ".join(f"{k}={v:.0f}" for k, v in st.most_common(3))
        print(f"{n:7.0f} {100 * n / total:5.1f}% {ie:9.0f} inst  {where:24s} {text:80s} [{top}]")


if __name__ == "__main__":
    main()
