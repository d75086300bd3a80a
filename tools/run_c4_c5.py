#!/usr/bin/env python3
"""BASELINE.json configs 4 and 5 sharded over the GPUs of one node, with the NCCL merge of the report.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/run_c4_c5.py --out gpurun_out/r02_c4_c5

One process per GPU (ShardedEngine: rank r holds a contiguous block of global walkers, walker w is `--seed seed + w`
whatever the GPU count), no traffic while moving; after the moves ONE all-gather merges the folded report.

C4  LJ38, R = 3 sigma, translation scale 0.05, energy bin 0.01 (run-decahedra-clusters.sh:15,43-47):
      1/t-WL on [-173, -100]  vs  SAD (min_T 0.05, max_allowed_energy 0),
    compared through the walker-mean entropy on the bins both cover.
C5  WCA N = 256, reduced density 0.8 (wca/run-wca.py:104-107 volumes), SAMC t0 = 1e7, energy bin 1, fast tier.

Rank 0 prints one JSON line per run and writes the merged arrays to <out>_<name>.npz.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sad_monte_carlo_b200 import make_config, _abi  # noqa: E402
from sad_monte_carlo_b200.parallel import ShardedEngine, unpack_merged  # noqa: E402

FM, R = _abi.FLAG_FAST_MATH, _abi.INIT_RANDOMIZE


def run_one(name, cfg, walkers_per_gpu, moves, chunk, rank, world, local, sad_mode, out):
    import torch
    import torch.distributed as dist
    se = ShardedEngine(cfg, walkers_per_gpu * world, rank=rank, world=world, device=local)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms, done = 0.0, 0
    t0 = time.time()
    while done < moves:
        n = min(chunk, moves - done)
        se.run(n)
        ms += se.engine.last_run_ms()
        done += n
    t = torch.tensor([ms], dtype=torch.float64, device=se.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    se.engine.fold_select(0, 1, sad_mode)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    merged = unpack_merged(se.merged_device())
    e1.record()
    torch.cuda.synchronize()
    halted = torch.tensor(list(se.engine.num_halted()), dtype=torch.int64, device=se.device)
    if world > 1:
        dist.all_reduce(halted)
    lo, width, nb = se.engine.window()
    line = None
    if rank == 0:
        total = world * walkers_per_gpu
        line = {"run": name, "n_gpus": world, "walkers_total": total, "moves_per_walker": moves,
                "moves_per_s": total * moves / (float(t.item()) * 1e-3), "kernel_ms_max_over_ranks": float(t.item()),
                "merge_ms": e0.elapsed_time(e1), "merged_histogram_total": int(merged["histogram"].sum()),
                "expected_histogram_total": total * (moves + 1), "halted": halted.tolist(), "wall_s": round(time.time() - t0, 1)}
        np.savez_compressed("%s_%s.npz" % (out, name), window_lo=lo, width=width, walkers=total, moves=moves, **merged)
        print(json.dumps(line), flush=True)
    se.close()
    return (lo, width, merged) if rank == 0 else None


def mean_entropy(merged, min_walkers):
    cnt = merged["lnw_count"].astype(np.float64)
    ok = cnt >= min_walkers
    S = np.zeros_like(cnt)
    S[ok] = merged["lnw_sum"][ok] / cnt[ok]
    return S, ok


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_c4_c5"))
    ap.add_argument("--lj-walkers", type=int, default=18944)
    ap.add_argument("--lj-moves", type=float, default=5e6)
    ap.add_argument("--wca-walkers", type=int, default=9472)
    ap.add_argument("--wca-moves", type=float, default=2e5)
    ap.add_argument("--only", default="c4,c5")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lj = dict(N=38, lj_radius=3.0, energy_bin=0.01, move_value=0.05, n_walkers=1, init_mode=R, lanes_per_walker=1, flags=FM, seed=0)
    if "c4" in a.only:
        wl = run_one("c4_lj38_inv_t_wl", make_config("lj", "inv-t-wl", min_allowed_energy=-173.0, max_allowed_energy=-100.0,
                                                    bin_window_lo=-173.1, bin_window_hi=-99.9, **lj),
                     a.lj_walkers, int(a.lj_moves), 1000000, rank, world, local, 0, a.out)
        sad = run_one("c4_lj38_sad", make_config("lj", "sad", max_allowed_energy=0.0, sad_min_T=0.05, bin_window_lo=-174.0,
                                                 bin_window_hi=0.02, **lj),
                      a.lj_walkers, int(a.lj_moves), 1000000, rank, world, local, 2, a.out)
        if rank == 0:
            (lo1, w1, m1), (lo2, w2, m2) = wl, sad
            total = world * a.lj_walkers
            S1, ok1 = mean_entropy(m1, 0.5 * total)
            S2, ok2 = mean_entropy(m2, 0.5 * total)
            # the two windows share the bin grid (multiples of the bin width): align by energy
            E1 = lo1 + (np.arange(len(S1)) + 0.5) * w1
            E2 = lo2 + (np.arange(len(S2)) + 0.5) * w2
            shift = int(round((lo1 - lo2) / w1))
            idx2 = np.arange(len(S1)) + shift
            inside = (idx2 >= 0) & (idx2 < len(S2))
            both = inside & ok1
            both[inside] &= ok2[idx2[inside]]
            d = S1[both] - S2[idx2[both]]
            d = d - d.mean() if d.size else d
            print(json.dumps({"comparison": "C4 LJ38: walker-mean entropy, 1/t-WL [-173, -100] minus SAD (min_T 0.05), constant removed",
                              "bins_compared": int(both.sum()), "energy_range": [float(E1[both].min()), float(E1[both].max())] if both.any() else None,
                              "rms_difference": float(np.sqrt(np.mean(d * d))) if d.size else None,
                              "max_abs_difference": float(np.abs(d).max()) if d.size else None,
                              "wl_bins_visited_by_half_the_walkers": int(ok1.sum()), "sad_bins_in_range_of_half_the_walkers": int(ok2.sum()),
                              "lowest_energy_wl": float(E1[ok1].min()) if ok1.any() else None,
                              "lowest_energy_sad": float(E2[ok2].min()) if ok2.any() else None}), flush=True)
    if "c5" in a.only:
        N = 256
        run_one("c5_wca256_samc", make_config("wca", "samc", N=N, reduced_density=0.8, samc_t0=1e7, energy_bin=1.0, max_allowed_energy=10.0 * N,
                                              n_walkers=1, init_mode=R, bin_window_lo=0.0, bin_window_hi=10.0 * N + 40.0, lanes_per_walker=8,
                                              flags=FM, seed=0),
                a.wca_walkers, int(a.wca_moves), 100000, rank, world, local, 0, a.out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
