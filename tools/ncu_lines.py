#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS export with `nvdisasm -g` line info and aggregate
executed instructions / stall samples per CUDA source line.

  cuobjdump -xelf all libsubzero_b200.so; nvdisasm -g -c sz_kernels.sm_100a.cubin > kern.sass
  ncu -i prof.ncu-rep --page source --csv --kernel-name regex:k_narrow > src.csv
  python tools/ncu_lines.py kern.sass src.csv _Z8k_narrow 40
"""
import collections
import csv
import re
import sys


def main():
    sass, src, func, top = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lines = {}
    cur, infn = None, False
    for ln in open(sass):
        if ln.startswith(".text."):
            infn = func in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m:
            lines[int(m.group(1), 16)] = (cur, m.group(2).strip())
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    H = rows[hi]
    ie, isamp = H.index("Instructions Executed"), H.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    tot_i = tot_s = 0.0
    for r in rows[hi + 1:]:
        if r and r[0] in ("Address", "Kernel Name"):
            break  # only the first captured launch of the kernel
        if len(r) != len(H):
            continue
        a = int(r[0], 16)
        if base is None:
            base = a
        key = lines.get(a - base, (None, ""))[0]
        n, s = float(r[ie] or 0), float(r[isamp] or 0)
        agg[key][0] += n
        agg[key][1] += s
        for i, h in stall_cols:
            v = float(r[i] or 0)
            if v:
                agg[key][2][h] += v
        tot_i += n
        tot_s += s
    print("total warp instructions %.0f, samples %.0f" % (tot_i, tot_s))
    srcs = {}
    for key, (n, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        text = ""
        if key:
            f = key[0]
            if f not in srcs:
                try:
                    srcs[f] = open("/root/repo/subzero.jl_b200/csrc/" + f).read().split("\n")
                except OSError:
                    srcs[f] = []
            if 0 < key[1] <= len(srcs[f]):
                text = srcs[f][key[1] - 1].strip()[:70]
        top3 = ",".join("%s=%.0f" % (k.replace("stall_", ""), v) for k, v in st.most_common(3))
        print("%-22s instr %5.1f%% samp %5.1f%% [%s] %s" % ("%s:%d" % key if key else "?", 100 * n / tot_i, 100 * s / tot_s, top3, text))


if __name__ == "__main__":
    main()
