#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS export with `nvdisasm -g` line info and aggregate
executed instructions / stall samples per CUDA source line — including __noinline__ device
functions, which nvdisasm lists as separate .text sections: each section is located inside the
ncu listing by matching its opcode sequence.

  cuobjdump -xelf all libsubzero_b200.so; nvdisasm -g -c sz_kernels.sm_100a.cubin > kern.sass
  ncu -i prof.ncu-rep --page source --csv --kernel-name regex:k_narrow_ab > src.csv
  python tools/ncu_lines.py kern.sass src.csv 40 [launch_index]
"""
import collections
import csv
import re
import sys


def opcode(s):
    s = s.strip()
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    return s.split()[0] if s else ""


def main():
    sass, src, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    funcs, cur, name = collections.OrderedDict(), None, None
    for ln in open(sass):
        if ln.startswith(".text."):
            name = ln.strip()[6:-1]
            funcs[name] = []
            cur = None
            continue
        if name is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m:
            funcs[name].append((opcode(m.group(2)), cur))
    rows = list(csv.reader(open(src)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hi = starts[which]
    H = rows[hi]
    end = len(rows)
    for i in range(hi + 1, len(rows)):
        if rows[i] and rows[i][0] in ("Address", "Kernel Name"):
            end = i
            break
    body = [r for r in rows[hi + 1:end] if len(r) == len(H)]
    ops = [opcode(r[1]) for r in body]
    ie, isamp = H.index("Instructions Executed"), H.index("# Samples")
    ith = H.index("Thread Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
    # locate every nvdisasm function inside the ncu listing
    line_of = [None] * len(body)
    func_of = [None] * len(body)
    for fname, ins in funcs.items():
        if len(ins) < 4:
            continue
        key = [o for o, _ in ins]
        n = len(key)
        for pos in range(0, len(ops) - n + 1):
            if ops[pos] == key[0] and ops[pos:pos + min(n, 24)] == key[:min(n, 24)] and ops[pos + n - 1] == key[-1]:
                if all(f is None for f in func_of[pos:pos + n]):
                    for k in range(n):
                        line_of[pos + k] = ins[k][1]
                        func_of[pos + k] = fname
                    break
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, collections.Counter()])
    fagg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    tot_i = tot_s = 0.0
    for r, key, fn in zip(body, line_of, func_of):
        n, s, th = float(r[ie] or 0), float(r[isamp] or 0), float(r[ith] or 0)
        a = agg[key]
        a[0] += n
        a[1] += s
        a[2] += th
        for i, h in stall_cols:
            v = float(r[i] or 0)
            if v:
                a[3][h] += v
        fa = fagg[fn]
        fa[0] += n
        fa[1] += s
        fa[2] += th
        tot_i += n
        tot_s += s
    print("total warp instructions %.0f, samples %.0f" % (tot_i, tot_s))
    print("-- per function")
    for fn, (n, s, th) in sorted(fagg.items(), key=lambda kv: -kv[1][1]):
        print("  %-60s instr %5.1f%% samp %5.1f%% lanes %4.1f" % ((fn or "?")[:60], 100 * n / tot_i, 100 * s / tot_s, th / n if n else 0))
    print("-- per line")
    srcs = {}
    for key, (n, s, th, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        text = ""
        if key:
            f = key[0]
            if f not in srcs:
                try:
                    srcs[f] = open("/root/repo/subzero.jl_b200/csrc/" + f).read().split("\n")
                except OSError:
                    srcs[f] = []
            if 0 < key[1] <= len(srcs[f]):
                text = srcs[f][key[1] - 1].strip()[:64]
        top3 = ",".join("%s=%.0f" % (k.replace("stall_", ""), v) for k, v in st.most_common(3))
        print("%-26s instr %5.1f%% samp %5.1f%% lanes %4.1f [%s] %s" % ("%s:%d" % key if key else "?", 100 * n / tot_i, 100 * s / tot_s,
                                                                    th / n if n else 0, top3, text))


if __name__ == "__main__":
    main()
