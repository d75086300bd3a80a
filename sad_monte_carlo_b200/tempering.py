"""Host-side mirror of the reference's replica-exchange Monte Carlo (`MC<S>` of src/mc/tempering.rs, the `tempering`
binary) for a batch of independent simulations on one GPU.

`TemperingMC` wraps the `sadmc_tempering_*` entry points of include/sadmc_gpu.h; the names follow the reference
(`run_once`, `moves`, `replicas`, `canonical_steps`).  Simulation k of the batch is the reference process run with
`--seed seed + k`.  All compute happens in libsadmc_gpu.so.
"""
import ctypes as C

import numpy as np

from . import load_library
from ._abi import Config, ReplicaState
from ._capi import f64p, u64p
from .engine import SadmcError


def geometric_spacing(min_T, max_T, num_T):
    """two-wells/run-two-wells.py:36-43: the temperature ladder the reference's job scripts hand to `--T`."""
    r = (max_T / min_T) ** (1.0 / (num_T - 1))
    return [min_T * r ** i for i in range(num_T)]


class TemperingMC:
    """`tempering::MC<Any>` x n_sim (cfg.n_walkers) on one GPU."""

    def __init__(self, cfg: Config, T, canonical_steps=1):
        self.L = load_library()
        self.cfg = cfg
        self.T = np.ascontiguousarray(T, dtype=np.float64)
        self.n_sim, self.n_T = int(cfg.n_walkers), int(self.T.size)
        self.canonical_steps = int(canonical_steps)
        self.h = C.c_void_p()
        self._check(self.L.sadmc_tempering_create(C.byref(cfg), self.T.ctypes.data_as(f64p), self.n_T, self.canonical_steps, C.byref(self.h)))

    def _check(self, rc):
        if rc != 0:
            raise SadmcError(rc, self.L.sadmc_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.sadmc_tempering_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_once(self, n_rounds=1):
        """n_rounds x `MC::run_once` (tempering.rs:272-342) for every simulation."""
        self._check(self.L.sadmc_tempering_run(self.h, int(n_rounds)))

    @property
    def moves(self):
        n = C.c_uint64()
        self._check(self.L.sadmc_tempering_num_moves(self.h, C.byref(n)))
        return n.value

    @property
    def steps_per_round(self):
        n = C.c_uint64()
        self._check(self.L.sadmc_tempering_steps_per_round(self.h, C.byref(n)))
        return n.value

    def last_run_ms(self):
        ms = C.c_float()
        self._check(self.L.sadmc_tempering_last_run_ms(self.h, C.byref(ms)))
        return ms.value

    def replicas(self, sim=0):
        out = (ReplicaState * self.n_T)()
        self._check(self.L.sadmc_tempering_get_replicas(self.h, sim, out))
        return list(out)

    def rng(self, sim=0):
        s = np.zeros(2, np.uint64)
        self._check(self.L.sadmc_tempering_get_rng(self.h, sim, s.ctypes.data_as(u64p)))
        return int(s[0]), int(s[1])

    def system(self, sim, replica):
        n = C.c_size_t()
        self._check(self.L.sadmc_tempering_system_len(self.h, C.byref(n)))
        buf = np.zeros(n.value)
        self._check(self.L.sadmc_tempering_get_system(self.h, sim, replica, buf.ctypes.data_as(f64p), buf.size))
        return buf

    def set_system(self, sim, replica, buf):
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        self._check(self.L.sadmc_tempering_set_system(self.h, sim, replica, buf.ctypes.data_as(f64p), buf.size))

    def restore(self, sim, doc):
        """Put a simulation's `MC` document (simulation_document) back: tempering.rs:196-213 deserialises the whole MC."""
        from .checkpoint import _system_image
        if [float(t) for t in doc["T"]] != [float(t) for t in self.T] or doc["canonical_steps"] != self.canonical_steps:
            raise ValueError("checkpoint ladder / canonical_steps differ from this simulation's")
        reps = (ReplicaState * self.n_T)()
        n = C.c_size_t()
        self._check(self.L.sadmc_tempering_system_len(self.h, C.byref(n)))
        for r, q in enumerate(doc["replicas"]):
            self.set_system(sim, r, _system_image(self.cfg, q["system"], n.value))
            reps[r].T = q["T"]
            for k in ("rejected_count", "accepted_count", "rejected_swap_count", "accepted_swap_count", "ignored_count",
                      "total_energy", "total_energy_squared", "translation_scale"):
                setattr(reps[r], k, q[k])
            reps[r].rng_s0, reps[r].rng_s1 = q["rng"]["s0"], q["rng"]["s1"]
        self._check(self.L.sadmc_tempering_set_replicas(self.h, sim, reps))
        s = np.array([doc["rng"]["s0"], doc["rng"]["s1"]], np.uint64)
        self._check(self.L.sadmc_tempering_set_rng(self.h, sim, s.ctypes.data_as(u64p)))

    def set_moves(self, moves):
        self._check(self.L.sadmc_tempering_set_num_moves(self.h, int(moves)))

    def set_translation_scales(self, scales):
        """`Replica::translation_scale` per temperature (the reference's constructor fixes 1.0, tempering.rs:88)."""
        a = np.ascontiguousarray(scales, dtype=np.float64)
        assert a.size == self.n_T
        self._check(self.L.sadmc_tempering_set_translation_scales(self.h, a.ctypes.data_as(f64p)))

    def all_replicas(self):
        """Counters and moments of every replica as arrays [n_sim, n_T] (one device read per simulation)."""
        keys = ["accepted_count", "rejected_count", "accepted_swap_count", "rejected_swap_count", "ignored_count",
                "total_energy", "total_energy_squared", "energy"]
        out = {k: np.zeros((self.n_sim, self.n_T)) for k in keys}
        for s in range(self.n_sim):
            for r, q in enumerate(self.replicas(s)):
                for k in keys:
                    out[k][s, r] = getattr(q, k)
        return out

    def cell_box(self):
        box, rc = (C.c_double * 3)(), C.c_double()
        self._check(self.L.sadmc_tempering_cell_box(self.h, box, C.byref(rc)))
        return [box[0], box[1], box[2]], rc.value

    def mean_energy(self, sim=0):
        """plotting/parse-tempering.py:57-73: <E> and <E^2> per temperature from the accumulated moments (the number of
        samples is moves + swap attempts - ignored, as there)."""
        reps = self.replicas(sim)
        n = np.array([r.accepted_count + r.rejected_count + r.accepted_swap_count + r.rejected_swap_count - r.ignored_count
                      for r in reps], float)
        e = np.array([r.total_energy for r in reps]) / n
        e2 = np.array([r.total_energy_squared for r in reps]) / n
        return e, e2


# ---- the `tempering` command line (src/bin/tempering.rs: `MC::<Any>::from_args`, `loop { mc.run_once() }`) ----------------
HELP = """python -m sad_monte_carlo_b200.tempering <system flags> --T t0 --T t1 ... [--canonical-steps k] [--seed s]
        [--max-iter n] [--save-time hours] [--movie-time x] [--save-as file.{yaml,json,cbor}] [--num-walkers n_sim] [--fast-math]

The reference's `tempering` binary (MCParams, src/mc/tempering.rs:14-41; two-wells/run-two-wells.py:45-61) for
`--num-walkers` independent simulations on one GPU (simulation k = `--seed seed + k`).  Checkpoints: one document per
simulation in the reference's serde schema (`MC`: T, rng, save_as, moves, replicas[], canonical_steps, save, movie, report;
tempering.rs:123-145), which plotting/parse-tempering.py reads; `--save-as` on an existing set resumes it (195-213)."""


def simulation_document(mc, sim, save_as, report=None, movie=None, save=None):
    """`MC<S>` of simulation `sim` as the reference serialises it (tempering.rs:123-145, Replica 46-73)."""
    from . import _abi
    from .checkpoint import _system_document
    cfg = mc.cfg
    reps = []
    for r, q in enumerate(mc.replicas(sim)):
        reps.append({"T": q.T, "rejected_count": int(q.rejected_count), "accepted_count": int(q.accepted_count),
                     "rejected_swap_count": int(q.rejected_swap_count), "accepted_swap_count": int(q.accepted_swap_count),
                     "ignored_count": int(q.ignored_count),
                     "system": _system_document(cfg, mc.system(sim, r), mc.cell_box() if cfg.system in (_abi.SYS_WCA, _abi.SYS_SW) else None),
                     "rng": {"s0": int(q.rng_s0), "s1": int(q.rng_s1)}, "total_energy": q.total_energy,
                     "total_energy_squared": q.total_energy_squared, "translation_scale": q.translation_scale})
    s0, s1 = mc.rng(sim)
    return {"T": [float(t) for t in mc.T], "rng": {"s0": s0, "s1": s1}, "save_as": str(save_as), "moves": int(mc.moves), "replicas": reps,
            "canonical_steps": mc.canonical_steps,
            "save": save if save is not None else {"save_time_seconds": 3600.0},
            "movie": movie if movie is not None else {"movie_time": None, "which_frame": 0, "period": "Never"},
            "report": report if report is not None else {"max_iter": "Never", "max_independent_samples": None, "quiet": True}}


def save_checkpoint(mc, save_as, **docs):
    """`MC::checkpoint` (tempering.rs:236-269): one file per simulation, written atomically."""
    import os
    from . import checkpoint as ck
    ext = os.path.splitext(str(save_as))[1].lstrip(".")
    out = []
    for sim in range(mc.n_sim):
        p = ck.walker_path(save_as, sim, mc.n_sim)
        ck.write_atomic(p, ck.dumps(simulation_document(mc, sim, p, **docs), ext))
        out.append(p)
    return out


def main(argv=None, out=print):
    import json
    import os
    import sys
    from . import _abi
    from . import histogram as H
    argv = list(sys.argv[1:] if argv is None else argv)
    # `T: Vec<f64>` (tempering.rs:17): --T may be given any number of times
    T, rest, i = [], [], 0
    while i < len(argv):
        a = argv[i]
        if a == "--T" and i + 1 < len(argv):
            T.append(H.evaluate(argv[i + 1]))
            i += 2
        elif a.startswith("--T="):
            T.append(H.evaluate(a[4:]))
            i += 1
        else:
            rest.append(a)
            i += 1
    H.ALL_FLAGS.setdefault("canonical-steps", H.INT)
    flags = H.parse_flags(rest)
    if flags.get("help"):
        out(HELP)
        return 0
    for bad in ("sad-min-T", "samc-t0", "wl", "wl-min-gamma", "inv-t-wl", "Inv-t-WL", "energy-bin", "min-allowed-energy", "max-allowed-energy",
                "translation-scale", "acceptance-rate", "resume-from"):
        if bad in flags:
            raise H.UsageError("--%s is not a flag of `tempering` (MCParams, tempering.rs:14-28)" % bad)
    if not T:  # MCParams::default (tempering.rs:31-33)
        T = [0.001, 0.002, 0.004, 0.008, 0.016, 0.032, 0.064, 0.128, 0.256, 0.512, 1.024]
    f = dict(flags)
    f["sad-min-T"] = 1.0  # a method is required by the shared parser; tempering ignores it
    cfg = H.config_from_flags(f)
    save_as = flags.get("save-as", "resume.yaml")
    if os.path.splitext(save_as)[1].lstrip(".") not in ("yaml", "json", "cbor"):
        raise H.UsageError("I don't know how to create file %r" % save_as)
    from . import checkpoint as ck
    resuming = "save-as" in flags and os.path.exists(ck.walker_path(save_as, 0, cfg.n_walkers))  # tempering.rs:195-213
    pp = H.plugin_params(flags)
    steps = flags.get("canonical-steps", 1)
    docs0 = None
    if resuming:  # the whole MC comes from the file (T, canonical_steps included); report and save parameters from the flags
        docs0 = [ck.load(ck.walker_path(save_as, k, cfg.n_walkers)) for k in range(cfg.n_walkers)]
        T, steps = [float(t) for t in docs0[0]["T"]], int(docs0[0]["canonical_steps"])
        if len({d["moves"] for d in docs0}) != 1:
            raise H.UsageError("the simulations' checkpoints disagree on moves")
    if flags.get("dry-run"):
        out(json.dumps({"config": H.config_summary(cfg), "T": T, "canonical_steps": steps, "plugins": pp, "save_as": save_as, "resuming": resuming}))
        return 0
    from . import plugins
    if resuming:
        cfg.init_mode = _abi.INIT_EXTERNAL
    mc = TemperingMC(cfg, T, steps)
    if resuming:
        try:
            for k, d in enumerate(docs0):
                mc.restore(k, d)
            mc.set_moves(docs0[0]["moves"])
        except ValueError as ex:
            raise H.UsageError(str(ex))
        out("Resuming from file %r" % save_as)
    report = plugins.Report(pp["max_iter"], pp["max_independent_samples"], pp["quiet"], out=out, resumed=resuming)
    saver = plugins.Save(pp["save_time"], resumed=resuming)
    movie = plugins.Movie(pp["movie_time"])
    if resuming and docs0[0].get("movie"):
        movie.restore(docs0[0]["movie"])
    docs = lambda: dict(report=report.document(), save=saver.document(), movie=movie.document())  # noqa: E731
    per_round = mc.steps_per_round * mc.n_T
    # MC::run_once ticks movie / report / save once per move of the round, AFTER the round (tempering.rs:321-341): a frame
    # or the final checkpoint carries the tick's move number and the end-of-round state.  Rounds are launched in batches
    # that end with the round in which the next scheduled event falls.
    max_iter = report.max_iter[1] if report.max_iter[0] == "TotalMoves" else None
    while True:
        before = mc.moves
        due = [m for m in (max_iter, movie.period[1] if movie.period[0] == "TotalMoves" else None, saver.next_output) if m is not None and m > before]
        target = min(due) if due else before + per_round * 1000
        n_rounds = max(1, min(1000, -(-(target - before) // per_round)))
        if "max-launch" in flags:
            n_rounds = max(1, min(n_rounds, flags["max-launch"] // max(1, per_round)))
        mc.run_once(n_rounds)
        after = mc.moves
        frame_at = None
        while movie.period[0] == "TotalMoves" and before < movie.period[1] <= after:
            m = movie.period[1]
            if not movie.shall_i_save(m):
                break
            frame_at = m
        if frame_at is not None:
            d = os.path.splitext(save_as)[0]
            save_checkpoint(mc, os.path.join(d, "%014d.cbor" % frame_at), **docs())
        if max_iter is not None and after >= max_iter:
            save_checkpoint(mc, save_as, **docs())
            out("All done!")
            break
        if saver.shall_i_save(after) or frame_at is not None:
            save_checkpoint(mc, save_as, **docs())
    mc.close()
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(main())
