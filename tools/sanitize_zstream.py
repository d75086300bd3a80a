#!/usr/bin/env python3
"""The LJ31 / LJ38 z-stream move kernels (cp.async ring, L2-resident stream) under compute-sanitizer, with awkward walker counts
(partial last CTA, ghost threads taking part in the warp-cooperative steps) and enough moves for accepted moves and an energy
re-summation.   compute-sanitizer --tool memcheck|racecheck python tools/sanitize_zstream.py [moves]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi  # noqa: E402

FM, R, Z = _abi.FLAG_FAST_MATH, _abi.INIT_RANDOMIZE, _abi.FLAG_LJ_STREAM_Z
moves = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
for name, method, kw in (
    ("lj31 sad z-stream", "sad", dict(N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01, n_walkers=150,
                                      bin_window_lo=-133.62, bin_window_hi=0.02)),
    ("lj31 wl z-stream", "wl", dict(N=31, lj_radius=2.5, max_allowed_energy=0.0, min_allowed_energy=-110.0, energy_bin=0.5, n_walkers=45,
                                    bin_window_lo=-133.62, bin_window_hi=0.6)),
    ("lj38 1/t-wl z-stream", "inv-t-wl", dict(N=38, lj_radius=3.0, max_allowed_energy=0.0, min_allowed_energy=-150.0, energy_bin=0.5, n_walkers=350,
                                              bin_window_lo=-174.0, bin_window_hi=0.6)),
):
    eng = WalkerEngine(make_config("lj", method, init_mode=R, lanes_per_walker=1, flags=FM | Z, **kw))
    assert eng.streams_z()
    eng.run(moves)
    eng.run(7)
    ok = all(eng.walker(w).status == 0 for w in range(eng.n_walkers))
    acc = sum(eng.walker(w).accepted_moves for w in range(0, eng.n_walkers, 7))
    print("%-24s walkers %4d moves %d accepted (every 7th walker) %d ok=%s" % (name, eng.n_walkers, eng.num_moves(), acc, ok), flush=True)
    eng.close()
