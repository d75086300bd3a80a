"""Host-side mirror of the reference's energy-ceiling replica Monte Carlo (`MC<S>` of src/mc/energy_replicas.rs, the
`replicas` binary; fake/run-fake.py:16-23 is a user) for a batch of independent simulations on one GPU.

`ReplicasMC` wraps the `sadmc_replicas_*` entry points of include/sadmc_gpu.h; names follow the reference (`run_once`,
`moves`, `replicas`, `median`).  Simulation k of the batch is the reference process run with `--seed seed + k`.
"""
import ctypes as C

import numpy as np

from . import load_library
from ._abi import Config, ZenoReplicaState
from ._capi import f64p, u64p
from .engine import SadmcError


class ReplicasMC:
    """`energy_replicas::MC<Any>` x n_sim (cfg.n_walkers) on one GPU."""

    def __init__(self, cfg: Config, min_T=0.2, independent_systems_before_new_bin=64, max_replicas=64, max_init=0):
        self.L = load_library()
        self.cfg = cfg
        self.n_sim, self.max_replicas = int(cfg.n_walkers), int(max_replicas)
        self.h = C.c_void_p()
        self._check(self.L.sadmc_replicas_create(C.byref(cfg), float(min_T), int(independent_systems_before_new_bin), self.max_replicas,
                                                 int(max_init), C.byref(self.h)))

    def _check(self, rc):
        if rc != 0:
            raise SadmcError(rc, self.L.sadmc_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.sadmc_replicas_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_once(self, n_rounds=1):
        """n_rounds x `MC::run_once` (energy_replicas.rs:504-642) for every simulation."""
        self._check(self.L.sadmc_replicas_run(self.h, int(n_rounds)))

    def moves(self, sim=0):
        n = C.c_uint64()
        self._check(self.L.sadmc_replicas_num_moves(self.h, sim, C.byref(n)))
        return n.value

    def num_replicas(self, sim=0):
        n = C.c_uint32()
        self._check(self.L.sadmc_replicas_num_replicas(self.h, sim, C.byref(n)))
        return n.value

    def replicas(self, sim=0):
        n = self.num_replicas(sim)
        out = (ZenoReplicaState * n)()
        self._check(self.L.sadmc_replicas_get_replicas(self.h, sim, n, out))
        return list(out)

    def rng(self, sim=0):
        s = np.zeros(2, np.uint64)
        self._check(self.L.sadmc_replicas_get_rng(self.h, sim, s.ctypes.data_as(u64p)))
        return int(s[0]), int(s[1])

    def median(self, sim=0):
        """The energies the MedianEstimator holds (energy_replicas.rs:45-99), in its order."""
        n = C.c_uint32()
        self._check(self.L.sadmc_replicas_get_median(self.h, sim, 0, None, C.byref(n)))
        e = np.zeros(n.value)
        self._check(self.L.sadmc_replicas_get_median(self.h, sim, n.value, e.ctypes.data_as(f64p), C.byref(n)))
        return e

    def system(self, sim, replica):
        n = C.c_size_t()
        self._check(self.L.sadmc_replicas_system_len(self.h, C.byref(n)))
        buf = np.zeros(n.value)
        self._check(self.L.sadmc_replicas_get_system(self.h, sim, replica, buf.ctypes.data_as(f64p), buf.size))
        return buf

    def last_run_ms(self):
        ms = C.c_float()
        self._check(self.L.sadmc_replicas_last_run_ms(self.h, C.byref(ms)))
        return ms.value


# ---- the `replicas` command line (src/bin/replicas.rs: `MC::<Any>::from_args`, `loop { mc.run_once() }`) -------------------
HELP = """python -m sad_monte_carlo_b200.replicas <system flags> --min-T t [--independent-systems-before-new-bin n] [--seed s]
        [--max-iter n] [--save-time hours] [--movie-time x] [--save-as file.{yaml,cbor}] [--num-walkers n_sim] [--max-replicas r]

The reference's `replicas` binary (MCParams, src/mc/energy_replicas.rs:15-42; fake/run-fake.py:16-23) for `--num-walkers`
independent simulations on one GPU (simulation k = `--seed seed + k`).  Checkpoints: one document per simulation in the
reference's serde schema (`MC`: min_T, rng, save_as, moves, independent_systems_before_new_bin, median, replicas[], save,
movie, report; energy_replicas.rs:307-333), which plotting/parse-replicas.py reads.  Resuming is not built."""


def simulation_document(mc, sim, save_as, min_T, indep, report=None, movie=None, save=None):
    """`MC<S>` of simulation `sim` as the reference serialises it (energy_replicas.rs:307-333, Replica 103-145)."""
    from . import _abi
    from .checkpoint import EXTRA_LABEL, _system_document
    cfg = mc.cfg
    cell = None
    reps = []
    for r, q in enumerate(mc.replicas(sim)):
        extra = {}
        if cfg.system in EXTRA_LABEL and q.above_extra_count:
            extra[EXTRA_LABEL[cfg.system]] = [q.above_extra_total, int(q.above_extra_count)]
        reps.append({"max_energy": q.max_energy, "cutoff_energy": q.cutoff_energy, "rejected_count": int(q.rejected_count),
                     "accepted_count": int(q.accepted_count), "above_count": int(q.above_count), "below_count": int(q.below_count),
                     "upwelling_count": int(q.upwelling_count), "above_total": q.above_total, "below_total": q.below_total,
                     "above_total_squared": q.above_total_squared, "below_total_squared": q.below_total_squared, "above_extra": extra,
                     "lowest_max_energy": q.lowest_max_energy, "system": _system_document(cfg, mc.system(sim, r), cell),
                     "unique_visitors": int(q.unique_visitors), "collecting_data": bool(q.collecting_data),
                     "rng": {"s0": int(q.rng_s0), "s1": int(q.rng_s1)}, "translation_scale": q.translation_scale})
    s0, s1 = mc.rng(sim)
    return {"min_T": min_T, "rng": {"s0": s0, "s1": s1}, "save_as": str(save_as), "moves": int(mc.moves(sim)),
            "independent_systems_before_new_bin": int(indep), "median": {"energies": [float(x) for x in mc.median(sim)]}, "replicas": reps,
            "save": save if save is not None else {"save_time_seconds": 3600.0},
            "movie": movie if movie is not None else {"movie_time": None, "which_frame": 0, "period": "Never"},
            "report": report if report is not None else {"max_iter": "Never", "max_independent_samples": None, "quiet": True}}


def save_checkpoint(mc, save_as, min_T, indep, **docs):
    """`MC::checkpoint` (energy_replicas.rs:453-501): one file per simulation, written atomically."""
    import os
    from . import checkpoint as ck
    ext = os.path.splitext(str(save_as))[1].lstrip(".")
    out = []
    for sim in range(mc.n_sim):
        p = ck.walker_path(save_as, sim, mc.n_sim)
        ck.write_atomic(p, ck.dumps(simulation_document(mc, sim, p, min_T, indep, **docs), ext))
        out.append(p)
    return out


def main(argv=None, out=print):
    import json
    import os
    import sys
    from . import histogram as H
    argv = list(sys.argv[1:] if argv is None else argv)
    H.ALL_FLAGS.setdefault("min-T", H.F64)
    H.ALL_FLAGS.setdefault("independent-systems-before-new-bin", H.INT)
    H.ALL_FLAGS.setdefault("max-replicas", H.INT)
    flags = H.parse_flags(argv)
    if flags.get("help"):
        out(HELP)
        return 0
    for bad in ("sad-min-T", "samc-t0", "wl", "wl-min-gamma", "inv-t-wl", "Inv-t-WL", "energy-bin", "min-allowed-energy", "max-allowed-energy",
                "translation-scale", "acceptance-rate", "resume-from", "T", "canonical-T"):
        if bad in flags:
            raise H.UsageError("--%s is not a flag of `replicas` (MCParams, energy_replicas.rs:15-29)" % bad)
    min_T = flags.get("min-T", 0.2)  # MCParams::default (energy_replicas.rs:34)
    indep = flags.get("independent-systems-before-new-bin", 64)  # 390
    f = dict(flags)
    for k in ("min-T", "independent-systems-before-new-bin", "max-replicas"):
        f.pop(k, None)
    f["sad-min-T"] = 1.0  # a method is required by the shared parser; replicas ignores it
    cfg = H.config_from_flags(f)
    save_as = flags.get("save-as", "resume.yaml")
    if os.path.splitext(save_as)[1].lstrip(".") not in ("yaml", "json", "cbor"):
        raise H.UsageError("I don't know how to create file %r" % save_as)
    from . import checkpoint as ck
    if "save-as" in flags and os.path.exists(ck.walker_path(save_as, 0, cfg.n_walkers)):
        raise H.UsageError("%s exists: resuming a `replicas` checkpoint is not built (remove the file to start over)" % save_as)
    pp = H.plugin_params(flags)
    r_max = flags.get("max-replicas", 64)
    if flags.get("dry-run"):
        out(json.dumps({"config": H.config_summary(cfg), "min_T": min_T, "independent_systems_before_new_bin": indep, "max_replicas": r_max,
                        "plugins": pp, "save_as": save_as}))
        return 0
    from . import plugins
    mc = ReplicasMC(cfg, min_T, indep, r_max)
    report = plugins.Report(pp["max_iter"], pp["max_independent_samples"], pp["quiet"], out=out)
    saver = plugins.Save(pp["save_time"])
    movie = plugins.Movie(pp["movie_time"])
    docs = lambda: dict(report=report.document(), save=saver.document(), movie=movie.document())  # noqa: E731
    max_iter = report.max_iter[1] if report.max_iter[0] == "TotalMoves" else None
    # the reference ticks movie / report / save once per move of the round, after the round (energy_replicas.rs:595-641); with
    # several simulations the schedule follows simulation 0 (their move counts differ once the ladders have grown differently)
    while True:
        before = mc.moves(0)
        n_rounds = 1000
        if max_iter is not None:
            left = max_iter - before
            per = r_max * max(1, _steps(cfg))  # the most one round can add: never run past the round that reaches max_iter
            n_rounds = max(1, min(1000, left // per))
        mc.run_once(n_rounds)
        after = mc.moves(0)
        frame_at = None
        while movie.period[0] == "TotalMoves" and before < movie.period[1] <= after:
            m = movie.period[1]
            if not movie.shall_i_save(m):
                break
            frame_at = m
        if frame_at is not None:
            d = os.path.splitext(save_as)[0]
            save_checkpoint(mc, os.path.join(d, "%014d.cbor" % frame_at), min_T, indep, **docs())
        if max_iter is not None and after >= max_iter:
            save_checkpoint(mc, save_as, min_T, indep, **docs())
            out("All done!")
            break
        if saver.shall_i_save(after) or frame_at is not None:
            save_checkpoint(mc, save_as, min_T, indep, **docs())
    mc.close()
    return 0


def _steps(cfg):
    """System::min_moves_to_randomize (ising.rs:86, fake.rs:113, lj.rs:280, wca.rs:268, erfinv.rs:86)."""
    from . import _abi
    if cfg.system == _abi.SYS_ISING:
        return cfg.N * cfg.N
    if cfg.system == _abi.SYS_FAKE:
        return 1 if cfg.fake_function == _abi.FAKE_LINEAR else (cfg.N if cfg.fake_function == _abi.FAKE_QUADRATIC else 3)
    return cfg.N


if __name__ == "__main__":
    import sys
    sys.exit(main())
