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

    def mean_energy(self, sim=0):
        """plotting/parse-tempering.py:57-73: <E> and <E^2> per temperature from the accumulated moments (the number of
        samples is moves + swap attempts - ignored, as there)."""
        reps = self.replicas(sim)
        n = np.array([r.accepted_count + r.rejected_count + r.accepted_swap_count + r.rejected_swap_count - r.ignored_count
                      for r in reps], float)
        e = np.array([r.total_energy for r in reps]) / n
        e2 = np.array([r.total_energy_squared for r in reps]) / n
        return e, e2
