#!/usr/bin/env python3
"""Attribute an ncu SASS profile to CUDA source lines (ncu's csv source page carries no line numbers).

Joins  `ncu -i REP --page source --csv`  (per-instruction executed counts + stall samples, by address)
with   `nvdisasm -gi` of the same kernel from the shipped .so (offset -> file:line, with the inlining chain)
by instruction offset.  The .so must be the build that was profiled.

usage: ncu_lines.py REP.ncu-rep KERNEL_SUBSTRING [--so PATH] [--warps N] [--top K] [--by inner|outer]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(so, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    out = ""
    for cubin in sorted(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")):  # one per translation unit
        text = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
        if kernel_sub in text:
            out = text
            break
    table = {}
    active = False
    chain = []
    in_run = False
    for ln in out.split("\n"):
        if ln.startswith("//-----") and ".text." in ln:
            active = kernel_sub in ln
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:  # a run of these lines is one inlining chain, innermost frame first
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            if not in_run:
                chain = []
                in_run = True
            chain.append(loc)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            in_run = False
            table[int(m.group(1), 16)] = (list(chain), m.group(2).strip())
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("--so", default=os.path.join(ROOT, "sad_monte_carlo_b200", "libsadmc_gpu.so"))
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--by", default="inner")
    ap.add_argument("--moves", type=float, default=0.0, help="moves per launch: prints instructions per warp-move")
    a = ap.parse_args()
    table = sass_lines(a.so, a.kernel)
    src = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    nwarps = None
    tot_s = tot_i = 0
    mismatch = 0
    for r in rows[2:]:
        try:
            addr = int(r[ix["Address"]], 16)
            n = int(r[ix["Instructions Executed"]])
        except Exception:
            continue
        if base is None:
            base, nwarps = addr, max(n, 1)
        off = addr - base
        chain, text = table.get(off, ([("?", 0)], "?"))
        if text.split()[0].split(".")[0] not in r[ix["Source"]]:
            mismatch += 1
        key = chain[0] if a.by == "inner" else chain[-1]
        s = int(r[ix["# Samples"]] or 0)
        agg[key][0] += n
        agg[key][1] += s
        for c in stall_cols:
            v = int(r[ix[c]] or 0)
            if v:
                agg[key][2][c[6:]] += v
        tot_s += s
        tot_i += n
    print("kernel instructions matched to nvdisasm: %d mismatching opcodes (0 = the .so is the profiled build)" % mismatch)
    print("warps %d, warp instructions %d%s, samples %d" % (nwarps, tot_i, (" (%.1f per warp-move)" % (tot_i / nwarps / a.moves)) if a.moves else "", tot_s))
    srcs = {}
    for (f, l), (n, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
        if f not in srcs:
            p = [os.path.join(dp, f) for dp, _, fs in os.walk(ROOT) if f in fs and ".git" not in dp]
            srcs[f] = open(p[0]).read().split("\n") if p else []
        text = srcs[f][l - 1].strip()[:80] if 0 < l <= len(srcs[f]) else ""
        top = ", ".join("%s %.0f%%" % (k, 100.0 * v / max(1, sum(st.values()))) for k, v in st.most_common(2))
        per = (" %7.1f" % (n / nwarps / a.moves)) if a.moves else ""
        print("%5.2f%% %s %-22s %-26s | %s" % (100.0 * s / tot_s, per, "%s:%d" % (f, l), top, text))


if __name__ == "__main__":
    main()
