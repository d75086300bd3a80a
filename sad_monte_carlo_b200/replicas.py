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
