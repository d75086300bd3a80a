#!/usr/bin/env python3
"""Compare a dump of the reference's random streams (docs/PIN_WITH_REFERENCE.md, `golden-dump.rs`) or a pair of
checkpoints with the CPU oracle.  Test infrastructure: uses oracle/ (never the product library).

    python tools/check_reference_dump.py dump.jsonl
    python tools/check_reference_dump.py --checkpoints ref.json ours.json [--rtol 1e-12]
    python tools/check_reference_dump.py --write-own-dump out.jsonl      # the oracle's own streams, same format
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.oracle_lib import load_oracle, rng_stream  # noqa: E402

KINDS = {"u64": 0, "f64_bits": 1, "gen_range": 2, "uniform": 3, "normal_bits": 4, "uniform_f64_bits": 5, "gen_range_f64_bits": 7}
SEEDS = [0, 1, 10137, 123456789]


def oracle_values(kind, seed, arg, count):
    L = load_oracle()
    st = np.zeros(2, np.uint64)
    L.oracle_rng_seed(int(seed), st.ctypes.data_as(C.POINTER(C.c_uint64)))
    if kind in ("gen_range", "uniform"):
        out = rng_stream(st, KINDS[kind], count, n_arg=int(arg))
    elif kind in ("uniform_f64_bits", "gen_range_f64_bits"):
        out = rng_stream(st, KINDS[kind], count, lo=float(arg[0]), hi=float(arg[1]))
    else:
        out = rng_stream(st, KINDS[kind], count)
    return [int(x) for x in np.asarray(out).view(np.uint64)]


def check_dump(path, out=print):
    bad = 0
    n = 0
    for line in open(path):
        line = line.strip()
        if not line:
            continue
        d = json.loads(line)
        want = [int(x) for x in d["values"]]
        got = oracle_values(d["kind"], d["seed"], d["arg"], len(want))
        n += 1
        if got != want:
            k = next(i for i, (a, b) in enumerate(zip(got, want)) if a != b)
            out("MISMATCH %s seed %s arg %s: first difference at draw %d (oracle %d, reference %d)" % (d["kind"], d["seed"], d["arg"], k, got[k], want[k]))
            bad += 1
    out("%d stream(s) checked, %d mismatch(es)" % (n, bad))
    return bad


def write_own_dump(path):
    with open(path, "w") as f:
        for seed in SEEDS:
            rows = [("u64", 0, 2000), ("f64_bits", 0, 2000)]
            for m in (2, 3, 31, 32, 38, 100, 256, 1000):
                rows += [("gen_range", m, 2000), ("uniform", m, 2000)]
            rows += [("normal_bits", 0, 20000), ("uniform_f64_bits", [-1.0, 1.0], 2000), ("gen_range_f64_bits", [-1.0, 1.0], 2000)]
            for kind, arg, count in rows:
                f.write(json.dumps({"kind": kind, "seed": seed, "arg": arg, "values": oracle_values(kind, seed, arg, count)}) + "\n")


def _close(a, b, rtol):
    if isinstance(a, float) or isinstance(b, float):
        if a is None or b is None:
            return a is b
        return a == b or abs(a - b) <= rtol * max(abs(a), abs(b), 1.0)
    if isinstance(a, dict) and isinstance(b, dict):
        return set(a) == set(b) and all(_close(a[k], b[k], rtol) for k in a)
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(_close(x, y, rtol) for x, y in zip(a, b))
    return a == b


def check_checkpoints(ref, ours, rtol=0.0, out=print):
    from sad_monte_carlo_b200 import checkpoint  # the codecs only (yaml / json / cbor readers)
    a, b = checkpoint.load(ref), checkpoint.load(ours)
    bad = 0
    for k in ("moves", "accepted_moves", "rng", "bins", "method", "system", "round_trips", "have_visited_since_maxentropy", "max_S",
              "max_S_index", "translation_scale", "acceptance_rate"):
        if k == "system":
            # derived tables / pending changes are not state of the trajectory
            sa, sb = json.loads(json.dumps(a[k])), json.loads(json.dumps(b[k]))
            for s in (sa, sb):
                body = next(iter(s.values()))
                for drop in ("possible_change", "change", "invcdf"):
                    body.pop(drop, None)
            ok = _close(sa, sb, rtol)
        else:
            ok = _close(a.get(k), b.get(k), rtol)
        if not ok:
            out("DIFFERENT: %s" % k)
            bad += 1
    out("checkpoints %s" % ("agree" if bad == 0 else "differ in %d field(s)" % bad))
    return bad


def main(argv):
    if argv and argv[0] == "--write-own-dump":
        write_own_dump(argv[1])
        return 0
    if argv and argv[0] == "--checkpoints":
        rtol = float(argv[argv.index("--rtol") + 1]) if "--rtol" in argv else 0.0
        return 1 if check_checkpoints(argv[1], argv[2], rtol) else 0
    return 1 if check_dump(argv[0]) else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
