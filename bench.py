#!/usr/bin/env python3
"""bench.py -- LJ31 SAD Monte-Carlo moves/sec on N B200s (BASELINE.json's metric, config 3).

A "step" is one launch of the hot path: `moves_per_step` x `EnergyMC::move_once` (reference
src/mc/energy.rs:904-974) for every walker of every GPU.  Workload = the reference's LJ31 SAD run
(`--lj-N 31 --lj-radius 2.5 --max-allowed-energy 0 --sad-min-T 0.01 --energy-bin 0.01
--translation-scale 0.05`, run-lj-clusters.sh:53), synthetic random-start walkers, f64.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA engine
    python bench.py --impl reference [...]                        # the CPU restatement on all host cores

Multi-GPU: launched by torchrun, one rank per GPU; walkers are sharded (weak scaling: fixed walkers
per GPU), no collective on the move path; one fold + NCCL all-reduce of the merged histogram after the
timed region (reporting-interval semantics), timed separately.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_MOVE = 900.0  # 30 (N - 1) FP64 flops for N = 31 (a divide counted as one flop)
BYTES_PER_MOVE = 80.0   # 2 lnw reads + RMW of histogram, energy_total, energy_squared_total, lnw
TIMED_MOVES_PER_WALKER = 10_000_000  # the timed window covers at least this many moves of every walker


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the move kernel from the `ncu --set full` capture committed under
    profiles/ (written by tools/ncu_traffic.py from the capture's raw page): bytes per move and what was captured."""
    p = os.path.join(ROOT, "profiles", "r02_lj31_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def lj31_config(n_walkers, walker_offset=0, device=0, lanes=0, flags=0):
    from sad_monte_carlo_b200 import make_config, _abi
    return make_config("lj", "sad", N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01,
                       move_value=0.05, n_walkers=n_walkers, walker_offset=walker_offset, device=device,
                       init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=lanes, seed=0, flags=flags,
                       bin_window_lo=-133.62, bin_window_hi=0.02)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def load_oracle():
    """The CPU restatement (oracle/): used ONLY as cpu_baseline / reference arm, never by the product path."""
    so = os.path.join(ROOT, "oracle", "liboracle_sadmc.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    from sad_monte_carlo_b200._abi import Config
    L = C.CDLL(so)
    L.oracle_bench.restype = C.c_double
    L.oracle_bench.argtypes = [C.POINTER(Config), C.c_uint32, C.c_uint64, C.c_uint64]
    return L


def cpu_moves_per_sec(threads, target_seconds, warmup_moves=200000):
    """One independent LJ31 SAD walker per host thread; returns (moves/s, sample description)."""
    L = load_oracle()
    cfg = lj31_config(threads)
    probe = 200000
    t = L.oracle_bench(C.byref(cfg), threads, warmup_moves, probe)
    if t <= 0:
        raise RuntimeError("oracle_bench failed")
    n = max(probe, int(probe * target_seconds / t))
    t = L.oracle_bench(C.byref(cfg), threads, warmup_moves, n)
    return threads * n / t, "%d threads x %d moves (after %d warm-up moves each), %.1f s" % (threads, n, warmup_moves, t)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    L = load_oracle()
    cfg = lj31_config(threads)
    per_step = args.cpu_moves_per_step
    for _ in range(args.warmup):
        L.oracle_bench(C.byref(cfg), threads, 0, per_step // 4)
    t_total = 0.0
    for _ in range(args.steps):
        t_total += L.oracle_bench(C.byref(cfg), threads, 100000, per_step)
    value = threads * per_step * args.steps / t_total
    sample = "%d host threads x %d moves per step, one walker per thread, construction + 1e5 warm-up moves untimed" % (threads, per_step)
    print(json.dumps({
        "impl": "reference", "metric": "LJ31 SAD MC moves/sec", "value": value, "unit": "moves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # our arm's config; what this arm actually ran on the CPU is the bounded sample described in cpu_baseline.sample
        "config": our_config(args),
        "cpu_baseline": {"value": value, "unit": "moves/s", "cores": threads, "kind": "port", "sample": sample},
        # what THIS arm ran (the `config` above names the shared workload as the GPU arm runs it)
        "ran": {"walkers": threads, "moves_per_walker_per_step": per_step, "untimed_moves_per_walker_per_step": 100000,
                "arithmetic": "reference operation order (oracle/, g++ -O3 -ffp-contract=off)", "host_threads": threads,
                "round_trip_diagnostics": True, "timed_s": t_total},
        "e2e": {"value": value, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(walkers_per_gpu, moves_per_step, **kw):
    c = {"workload": "LJ31 SAD f64: N=31 R=2.5 max_allowed_energy=0 min_T=0.01 energy_bin=0.01 translation_scale=0.05 "
                     "(reference run-lj-clusters.sh:53), random-start walkers",
         "walkers_per_gpu": walkers_per_gpu, "moves_per_walker_per_step": moves_per_step,
         "l2": "inputs larger than L2 (per-walker bin windows: tens of GB)"}
    c.update(kw)
    return c


def our_config(args):
    """The `config` block of the GPU arm (the reference arm repeats it and adds what it ran itself under `ran`)."""
    return workload_config(args.walkers, args.moves_per_step, lanes_per_walker=args.lanes,
                           arithmetic="exact (reference operation order)" if args.exact else "fast-math (<= 1e-12 rel. per move)",
                           burn_in_moves=args.burn_in, round_trip_diagnostics=not args.no_round_trips,
                           timed_window="moves %d to %d of every walker" % (
                               args.burn_in + args.warmup * args.moves_per_step,
                               args.burn_in + (args.warmup + args.steps) * args.moves_per_step))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from sad_monte_carlo_b200 import WalkerEngine, load_library

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the walker engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load_library()
    W = args.walkers
    cfg = lj31_config(W, walker_offset=rank * W, device=local, lanes=args.lanes,
                      flags=(1 if args.no_round_trips else 0) | (0 if args.exact else 4))
    eng = WalkerEngine(cfg)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # un-timed burn-in: past SAD's initial range-discovery transient
    eng.run(args.burn_in)
    launches0 = eng.launch_count()
    for _ in range(args.warmup):
        eng.run(args.moves_per_step)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches1 = eng.launch_count()
    ev0.record(stream)
    for _ in range(args.steps):
        eng.run_async(args.moves_per_step)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    gpu_launches = eng.launch_count() - launches1
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_moves = float(world) * W * args.moves_per_step * args.steps
    value = total_moves / (ms_max * 1e-3)

    # ---- end to end through the C ABI with HOST buffers: resume-from-host -> run -> merged bins to host ----
    sys_host = torch.empty((W, eng.system_len), dtype=torch.float64).pin_memory().numpy()
    rng_host = torch.empty((W, 2), dtype=torch.int64).pin_memory().numpy().view(np.uint64)
    sys_host[:] = eng.systems()
    rng_host[:] = eng.rngs()
    _, _, nb = eng.window()
    h2d = sys_host.nbytes + rng_host.nbytes
    d2h = 6 * nb * 8 + W * 8
    for _ in range(2):
        eng.set_systems(sys_host)
        eng.set_rngs(rng_host)
        eng.run(args.moves_per_step)
        eng.fold()
        eng.energies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.set_systems(sys_host)
        eng.set_rngs(rng_host)
        eng.run(args.moves_per_step)
        merged = eng.fold()
        energies = eng.energies()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_moves / float(te.item())

    # ---- reporting-interval merge: one-pass device fold into ONE packed buffer + ONE collective (rank-ordered) ----
    from sad_monte_carlo_b200.parallel import PACKED_FIELDS, merge_packed, unpack_merged
    dev = torch.device("cuda", local)
    packed = torch.zeros((PACKED_FIELDS, nb), dtype=torch.float64, device=dev)
    merge_packed(packed)  # warm the communicator
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    eng.fold_packed_device(packed.data_ptr())
    merged_packed = merge_packed(packed)
    f1.record(stream)
    torch.cuda.synchronize()
    fold_ms = f0.elapsed_time(f1)
    moves_now = eng.num_moves()
    hist_total = int(unpack_merged(merged_packed)["histogram"].sum())
    halted = eng.num_halted()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        traffic = measured_traffic()
        fp64 = C.c_double(0.0)
        lib.sadmc_measure_fp64_peak(local, 5, C.byref(fp64))
        per_gpu_moves_s = value / world
        line = {
            "metric": "LJ31 SAD MC moves/sec", "value": value, "unit": "moves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": our_config(args),
            "clocks": clocks, "gpu_launches": int(gpu_launches),
            "e2e": {"value": e2e_value, "unit": "moves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "sadmc_set_systems + sadmc_set_rngs (pinned host) -> sadmc_run -> sadmc_fold + sadmc_get_energies (host)"},
            "roofline": {"bound": "fp64", "achieved": per_gpu_moves_s * FLOPS_PER_MOVE / 1e12, "peak": fp64.value,
                         "unit": "TFLOP/s", "frac": (per_gpu_moves_s * FLOPS_PER_MOVE / 1e12) / fp64.value if fp64.value else None,
                         "traffic": traffic["dram_bytes_per_move"] * W * args.moves_per_step if traffic else None,
                         "traffic_unit": "DRAM bytes per launch (measured bytes per move of the committed ncu capture x moves per launch; "
                                         "algorithmic: 80 B per move)",
                         "per_unit": "900 FP64 flop per move (30 per pair x 30 pairs), divide = 1 flop; "
                                     "`achieved` = 900 x moves per launch / average launch duration",
                         "peak_source": "DFMA microkernel in this library, measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "kernel": "move_kernel<LjThreadSys<%s, 31, %d%s>, SAD>" % ("exact" if args.exact else "fast", args.lanes, ", z streamed from L2" if eng.streams_z() else ""), "ms_per_launch": ms_max / args.steps},
            "roofline_hbm": {"bound": "hbm", "achieved": per_gpu_moves_s * BYTES_PER_MOVE / 1e9, "peak": peaks.get("hbm_gbs"),
                             "unit": "GB/s", "frac": per_gpu_moves_s * BYTES_PER_MOVE / 1e9 / peaks.get("hbm_gbs"),
                             "traffic": traffic["dram_bytes_per_move"] * W * args.moves_per_step if traffic else None,
                             "traffic_source": traffic["source"] if traffic else "no ncu capture committed for this kernel",
                             "per_unit": "80 B of bin traffic per move (algorithmic: 2 ln w reads + read-modify-write of 4 fields); "
                                         "traffic = measured DRAM bytes per move of the ncu capture x moves per launch",
                             "peak_source": peak_src},
            "fold_merge_ms": fold_ms,
            "fold_merge": "one-pass device fold into one packed buffer + one all-gather, shards added in rank order",
            "checks": {"merged_histogram_total": hist_total, "expected": int(world * W * (moves_now + 1)),
                       "walkers_halted": {"left_window": halted[0], "failed_verify": halted[1]}},
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                threads = os.cpu_count() or 1
                v, sample = cpu_moves_per_sec(threads, args.cpu_seconds)
                line["cpu_baseline"] = {"value": v, "unit": "moves/s", "cores": threads, "kind": "port", "sample": sample}
            except Exception as ex:  # the baseline is a reported number, never a reason to lose the GPU line
                line["cpu_baseline"] = {"value": None, "unit": "moves/s", "cores": 0, "kind": "port", "sample": "failed: %s" % ex}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--walkers", type=int, default=int(os.environ.get("SADMC_BENCH_WALKERS", 0)),
                    help="walkers per GPU (default: whole waves of resident CTAs: 56 832 = 148 SMs x 3 CTAs x 128 walkers, one lane per walker with z streamed from L2; 75 776 = two waves of 2 CTAs for --exact; 85 248 at two lanes)")
    ap.add_argument("--moves-per-step", type=int, default=int(os.environ.get("SADMC_BENCH_MOVES", 0)),
                    help="moves per walker and launch (default: so that the K timed steps cover 1e7 moves of every walker)")
    ap.add_argument("--burn-in", type=int, default=int(os.environ.get("SADMC_BENCH_BURN_IN", 1000000)),
                    help="un-timed moves per walker before the warm-up steps (BASELINE.md section 3: 1e6)")
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("SADMC_BENCH_LANES", 1)))
    ap.add_argument("--no-round-trips", action="store_true")
    ap.add_argument("--exact", action="store_true", help="reference operation order (bit-exact vs the oracle) instead of fast-math")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--cpu-moves-per-step", type=int, default=2000000)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.steps < 1:
        args.steps = 1
    if args.moves_per_step <= 0:
        args.moves_per_step = -(-TIMED_MOVES_PER_WALKER // args.steps)
    if args.exact and args.lanes != 1:
        args.lanes = 1  # the reference's sequential pair sum cannot be split across lanes
    if args.walkers == 0:
        # whole waves of resident CTAs: 148 SMs x 3 CTAs x 128 walkers (one lane per walker, z streamed from L2: the engine picks that
        # layout for this count), 148 x 2 x 128 x 2 waves (--exact: all coordinates in shared memory), 148 x 3 x 64 x 3 (two lanes)
        args.walkers = 85248 if args.lanes == 2 else (75776 if args.exact else 56832)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
