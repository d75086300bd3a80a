#!/usr/bin/env python3
"""DRAM traffic of the LJ31 move kernel from an `ncu --set full` capture -> profiles/r02_lj31_traffic.json (read by bench.py).

    ncu --set full --clock-control none --import-source on -k regex:move_kernel --launch-skip 2 --launch-count 1 \\
        -o gpurun_out/r02_lj31 python tools/profile_lj.py 75776 1 20000 4 3 1000000
    python tools/ncu_traffic.py gpurun_out/r02_lj31.ncu-rep 75776 20000 "after 1e6 burn-in moves per walker"

The capture is ONE launch of the bench workload's kernel (same walkers per GPU, same engine configuration, a shorter
launch than the bench's: ncu replays the kernel ~40 times); bench.py scales bytes per move to its own launch length.
"""
import csv
import io
import json
import os
import subprocess
import sys

rep, walkers, moves = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
note = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def metric(name):
    v, u = float(vals[col[name]].replace(",", "")), units[col[name]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u)
    if scale is None:
        raise SystemExit("unexpected unit %r for %s" % (u, name))
    return v * scale


rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
out = {"kernel": vals[col["Kernel Name"]], "walkers": walkers, "moves_per_launch": moves, "dram_bytes_read": rd, "dram_bytes_write": wr,
       "dram_bytes_per_move": (rd + wr) / (walkers * moves), "algorithmic_bytes_per_move": 80.0,
       "source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (%s: %d walkers x %d moves%s)" % (
           os.path.basename(rep), walkers, moves, ", " + note if note else "")}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_lj31_traffic.json")
json.dump(out, open(p, "w"), indent=1)
print(json.dumps(out))
