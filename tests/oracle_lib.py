"""ctypes face of oracle/liboracle_sadmc.so -- the CPU oracle (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

from sad_monte_carlo_b200._abi import BinningState, Config, ReplicaState, WalkerState, ZenoReplicaState

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def load_oracle():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(ROOT, "oracle", "liboracle_sadmc.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    L = C.CDLL(so)
    L.oracle_last_error.restype = C.c_char_p
    L.oracle_create.restype = C.c_void_p
    L.oracle_create.argtypes = [C.POINTER(Config), C.c_uint32, f64p, C.c_size_t, C.c_uint64]
    L.oracle_destroy.argtypes = [C.c_void_p]
    L.oracle_run.argtypes = [C.c_void_p, C.c_uint64]
    L.oracle_get_walker.argtypes = [C.c_void_p, C.POINTER(WalkerState)]
    L.oracle_get_bins.argtypes = [C.c_void_p, C.c_uint32, u64p, u64p, f64p, f64p, f64p, u64p, u8p, u64p, f64p, u64p]
    L.oracle_system_len.restype = C.c_size_t
    L.oracle_system_len.argtypes = [C.c_void_p]
    L.oracle_get_system.argtypes = [C.c_void_p, f64p, C.c_size_t]
    L.oracle_set_system.argtypes = [C.c_void_p, f64p, C.c_size_t]
    L.oracle_set_rng.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    L.oracle_set_lj_tree_lanes.argtypes = [C.c_void_p, C.c_int]
    L.oracle_sys_energy.restype = C.c_double
    L.oracle_sys_energy.argtypes = [C.c_void_p]
    L.oracle_sys_compute_energy.restype = C.c_double
    L.oracle_sys_compute_energy.argtypes = [C.c_void_p]
    L.oracle_sw_compute_energy_slowly.restype = C.c_double
    L.oracle_sw_compute_energy_slowly.argtypes = [C.c_void_p]
    L.oracle_sys_plan_move.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_int), f64p]
    L.oracle_sys_confirm.argtypes = [C.c_void_p]
    L.oracle_sys_randomize.argtypes = [C.c_void_p, f64p]
    L.oracle_sys_verify_energy.argtypes = [C.c_void_p]
    L.oracle_rng_seed.argtypes = [C.c_uint64, u64p]
    L.oracle_rng_stream.argtypes = [u64p, C.c_int, C.c_uint64, C.c_double, C.c_double, C.c_uint64, u64p]
    L.oracle_exp.restype = C.c_double
    L.oracle_exp.argtypes = [C.c_double]
    L.oracle_log.restype = C.c_double
    L.oracle_log.argtypes = [C.c_double]
    L.oracle_erf_inv.restype = C.c_double
    L.oracle_erf_inv.argtypes = [C.c_double]
    L.oracle_zig_tables.argtypes = [f64p, f64p]
    L.oracle_bench.restype = C.c_double
    L.oracle_bench.argtypes = [C.POINTER(Config), C.c_uint32, C.c_uint64, C.c_uint64]
    L.oracle_set_math_mode.argtypes = [C.c_int]
    L.oracle_replicas_create.restype = C.c_void_p
    L.oracle_replicas_create.argtypes = [C.POINTER(Config), C.c_uint32, C.c_double, C.c_uint64, C.c_uint32, C.c_uint64]
    L.oracle_replicas_destroy.argtypes = [C.c_void_p]
    L.oracle_replicas_run.argtypes = [C.c_void_p, C.c_uint64]
    L.oracle_replicas_num_moves.restype = C.c_uint64
    L.oracle_replicas_num_moves.argtypes = [C.c_void_p]
    L.oracle_replicas_num_replicas.restype = C.c_uint32
    L.oracle_replicas_num_replicas.argtypes = [C.c_void_p]
    L.oracle_replicas_get_rng.argtypes = [C.c_void_p, u64p]
    L.oracle_replicas_get_median.restype = C.c_uint32
    L.oracle_replicas_get_median.argtypes = [C.c_void_p, C.c_uint32, f64p]
    L.oracle_replicas_get_replicas.argtypes = [C.c_void_p, C.POINTER(ZenoReplicaState)]
    L.oracle_replicas_system_len.restype = C.c_size_t
    L.oracle_replicas_system_len.argtypes = [C.c_void_p]
    L.oracle_replicas_get_system.argtypes = [C.c_void_p, C.c_uint32, f64p, C.c_size_t]
    L.oracle_tempering_create.restype = C.c_void_p
    L.oracle_tempering_create.argtypes = [C.POINTER(Config), C.c_uint32, f64p, C.c_uint32, C.c_uint64, f64p, C.c_size_t, C.c_uint64]
    L.oracle_tempering_destroy.argtypes = [C.c_void_p]
    L.oracle_tempering_run.argtypes = [C.c_void_p, C.c_uint64]
    L.oracle_tempering_num_moves.restype = C.c_uint64
    L.oracle_tempering_num_moves.argtypes = [C.c_void_p]
    L.oracle_tempering_get_rng.argtypes = [C.c_void_p, u64p]
    L.oracle_tempering_get_replicas.argtypes = [C.c_void_p, C.POINTER(ReplicaState)]
    L.oracle_tempering_system_len.restype = C.c_size_t
    L.oracle_tempering_system_len.argtypes = [C.c_void_p]
    L.oracle_tempering_get_system.argtypes = [C.c_void_p, C.c_uint32, f64p, C.c_size_t]
    L.oracle_rng_jump.argtypes = [u64p]
    L.oracle_tempering_set_translation_scales.argtypes = [C.c_void_p, f64p]
    L.oracle_binning_create.restype = C.c_void_p
    L.oracle_binning_create.argtypes = [C.POINTER(Config), C.c_uint32, f64p, C.c_size_t, C.c_uint64]
    L.oracle_binning_destroy.argtypes = [C.c_void_p]
    L.oracle_binning_run.argtypes = [C.c_void_p, C.c_uint64]
    L.oracle_binning_get_walker.argtypes = [C.c_void_p, C.POINTER(BinningState)]
    L.oracle_binning_get_bins.argtypes = [C.c_void_p, C.c_uint32, f64p, u64p, f64p, u64p, f64p, u64p, u64p, f64p, u64p]
    L.oracle_binning_get_bins_f64.argtypes = [C.c_void_p, C.c_uint32, f64p, f64p, f64p, f64p, f64p, f64p, f64p, f64p, f64p]
    L.oracle_binning_get_aggregates.argtypes = [C.c_void_p, C.c_char_p, f64p]
    L.oracle_binning_get_high_resolution.argtypes = [C.c_void_p, C.c_uint32, f64p, C.POINTER(C.c_uint32), u64p]
    L.oracle_binning_system_len.restype = C.c_size_t
    L.oracle_binning_system_len.argtypes = [C.c_void_p]
    L.oracle_binning_get_system.argtypes = [C.c_void_p, f64p, C.c_size_t]
    _LIB = L
    return L


class OracleMC:
    """One reference walker: `EnergyMC<Any>` of the reference, restated on the CPU."""

    def __init__(self, cfg, walker=0, system_state=None, attempts_override=0):
        self.L = load_oracle()
        self.cfg = cfg
        st = None
        n = 0
        if system_state is not None:
            st = np.ascontiguousarray(system_state, dtype=np.float64)
            n = st.size
        self.h = self.L.oracle_create(C.byref(cfg), walker, _ptr(st, f64p), n, attempts_override)
        if not self.h:
            raise RuntimeError("oracle_create: " + self.L.oracle_last_error().decode())

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, n):
        if self.L.oracle_run(self.h, int(n)) != 0:
            raise RuntimeError("oracle_run: " + self.L.oracle_last_error().decode())

    def walker(self):
        s = WalkerState()
        self.L.oracle_get_walker(self.h, C.byref(s))
        return s

    def bins(self):
        n = self.walker().bins_len
        out = {
            "histogram": np.zeros(n, np.uint64), "t_found": np.zeros(n, np.uint64), "lnw": np.zeros(n),
            "energy_total": np.zeros(n), "energy_squared_total": np.zeros(n),
            "round_trips": np.zeros(n, np.uint64), "have_visited": np.zeros(n, np.uint8),
            "wl_hist": np.zeros(n, np.uint64), "extra_total": np.zeros(n), "extra_count": np.zeros(n, np.uint64),
        }
        rc = self.L.oracle_get_bins(self.h, n, _ptr(out["histogram"], u64p), _ptr(out["t_found"], u64p),
                                    _ptr(out["lnw"], f64p), _ptr(out["energy_total"], f64p),
                                    _ptr(out["energy_squared_total"], f64p), _ptr(out["round_trips"], u64p),
                                    _ptr(out["have_visited"], u8p), _ptr(out["wl_hist"], u64p),
                                    _ptr(out["extra_total"], f64p), _ptr(out["extra_count"], u64p))
        assert rc == 0
        return out

    def system(self):
        n = self.L.oracle_system_len(self.h)
        buf = np.zeros(n)
        assert self.L.oracle_get_system(self.h, _ptr(buf, f64p), n) == 0
        return buf

    def set_system(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        self.L.oracle_set_system(self.h, _ptr(buf, f64p), buf.size)

    def set_rng(self, s0, s1):
        self.L.oracle_set_rng(self.h, int(s0), int(s1))

    def energy(self):
        return self.L.oracle_sys_energy(self.h)

    def compute_energy(self):
        return self.L.oracle_sys_compute_energy(self.h)

    def plan_move(self, mean_distance):
        some = C.c_int(0)
        e = C.c_double(0)
        self.L.oracle_sys_plan_move(self.h, mean_distance, C.byref(some), C.byref(e))
        return (e.value if some.value else None)

    def confirm(self):
        self.L.oracle_sys_confirm(self.h)

    def verify_energy(self):
        return self.L.oracle_sys_verify_energy(self.h) == 0

    def randomize(self):
        e = C.c_double(0)
        if self.L.oracle_sys_randomize(self.h, C.byref(e)) != 0:
            raise RuntimeError("oracle randomize: " + self.L.oracle_last_error().decode())
        return e.value


def rng_stream(state, kind, count, n_arg=0, lo=0.0, hi=1.0):
    """Draw `count` values; `state` is a 2-element uint64 array, updated in place."""
    L = load_oracle()
    out = np.zeros(count, np.uint64)
    L.oracle_rng_stream(_ptr(state, u64p), kind, n_arg, lo, hi, count, _ptr(out, u64p))
    return out


class OracleBinningMC:
    """One reference walker of the `binning` binary: energy_binning.rs `EnergyMC<Any>` over binning::histogram."""

    def __init__(self, cfg, walker=0, system_state=None, attempts_override=0):
        self.L = load_oracle()
        self.cfg = cfg
        st, n = None, 0
        if system_state is not None:
            st = np.ascontiguousarray(system_state, dtype=np.float64)
            n = st.size
        self.h = self.L.oracle_binning_create(C.byref(cfg), walker, _ptr(st, f64p), n, attempts_override)
        if not self.h:
            raise RuntimeError("oracle_binning_create: " + self.L.oracle_last_error().decode())

    def close(self):
        if self.h:
            self.L.oracle_binning_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, n):
        if self.L.oracle_binning_run(self.h, int(n)) != 0:
            raise RuntimeError("oracle_binning_run: " + self.L.oracle_last_error().decode())

    def walker(self):
        s = BinningState()
        self.L.oracle_binning_get_walker(self.h, C.byref(s))
        return s

    def bins(self):
        n = self.walker().bins_len
        out = {"lnw_total": np.zeros(n), "lnw_count": np.zeros(n, np.uint64), "energy_total": np.zeros(n),
               "energy_count": np.zeros(n, np.uint64), "t_found_total": np.zeros(n), "t_found_count": np.zeros(n, np.uint64),
               "hist_count": np.zeros(n, np.uint64), "extra_total": np.zeros(n), "extra_count": np.zeros(n, np.uint64)}
        rc = self.L.oracle_binning_get_bins(
            self.h, n, _ptr(out["lnw_total"], f64p), _ptr(out["lnw_count"], u64p), _ptr(out["energy_total"], f64p),
            _ptr(out["energy_count"], u64p), _ptr(out["t_found_total"], f64p), _ptr(out["t_found_count"], u64p),
            _ptr(out["hist_count"], u64p), _ptr(out["extra_total"], f64p), _ptr(out["extra_count"], u64p))
        assert rc == 0
        return out

    def bins_f64(self):
        n = self.walker().bins_len
        keys = ["lnw_total", "lnw_count", "energy_total", "energy_count", "t_found_total", "t_found_count", "hist_count", "extra_total", "extra_count"]
        out = {k: np.zeros(n) for k in keys}
        assert self.L.oracle_binning_get_bins_f64(self.h, n, *[_ptr(out[k], f64p) for k in keys]) == 0
        return out

    def high_resolution(self):
        mn, n = C.c_double(), C.c_uint32()
        assert self.L.oracle_binning_get_high_resolution(self.h, 0, C.byref(mn), C.byref(n), None) == 0
        cnt = np.zeros(n.value, np.uint64)
        assert self.L.oracle_binning_get_high_resolution(self.h, n.value, C.byref(mn), C.byref(n), _ptr(cnt, u64p)) == 0
        return mn.value, cnt

    def aggregates(self, name=""):
        """(min_total, max_total, e_max_total, min_count, max_count, e_max_count, total_count) of bins.lnw ("") or an extra."""
        out = np.zeros(7)
        if self.L.oracle_binning_get_aggregates(self.h, name.encode(), _ptr(out, f64p)) != 0:
            return None
        return out

    def system(self):
        n = self.L.oracle_binning_system_len(self.h)
        buf = np.zeros(n)
        assert self.L.oracle_binning_get_system(self.h, _ptr(buf, f64p), n) == 0
        return buf


class OracleTempering:
    """One reference `tempering` process: `MC<Any>` of src/mc/tempering.rs, restated on the CPU."""

    def __init__(self, cfg, T, canonical_steps=1, sim=0, system_state=None, attempts_override=0):
        self.L = load_oracle()
        self.T = np.ascontiguousarray(T, dtype=np.float64)
        st, n = None, 0
        if system_state is not None:
            st = np.ascontiguousarray(system_state, dtype=np.float64)
            n = st.size
        self.h = self.L.oracle_tempering_create(C.byref(cfg), sim, _ptr(self.T, f64p), self.T.size, int(canonical_steps),
                                                _ptr(st, f64p), n, attempts_override)
        if not self.h:
            raise RuntimeError("oracle_tempering_create: " + self.L.oracle_last_error().decode())

    def close(self):
        if self.h:
            self.L.oracle_tempering_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_once(self, n_rounds=1):
        if self.L.oracle_tempering_run(self.h, int(n_rounds)) != 0:
            raise RuntimeError("oracle_tempering_run: " + self.L.oracle_last_error().decode())

    @property
    def moves(self):
        return self.L.oracle_tempering_num_moves(self.h)

    def set_translation_scales(self, scales):
        a = np.ascontiguousarray(scales, dtype=np.float64)
        self.L.oracle_tempering_set_translation_scales(self.h, _ptr(a, f64p))

    def rng(self):
        s = np.zeros(2, np.uint64)
        self.L.oracle_tempering_get_rng(self.h, _ptr(s, u64p))
        return int(s[0]), int(s[1])

    def replicas(self):
        out = (ReplicaState * self.T.size)()
        self.L.oracle_tempering_get_replicas(self.h, out)
        return list(out)

    def system(self, replica):
        n = self.L.oracle_tempering_system_len(self.h)
        buf = np.zeros(n)
        assert self.L.oracle_tempering_get_system(self.h, replica, _ptr(buf, f64p), n) == 0
        return buf


class OracleReplicas:
    """One reference `replicas` process: `MC<Any>` of src/mc/energy_replicas.rs, restated on the CPU."""

    def __init__(self, cfg, min_T=0.2, independent_systems_before_new_bin=64, sim=0, max_init=0, attempts_override=0):
        self.L = load_oracle()
        self.h = self.L.oracle_replicas_create(C.byref(cfg), sim, float(min_T), int(independent_systems_before_new_bin), int(max_init), attempts_override)
        if not self.h:
            raise RuntimeError("oracle_replicas_create: " + self.L.oracle_last_error().decode())

    def close(self):
        if self.h:
            self.L.oracle_replicas_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_once(self, n_rounds=1):
        if self.L.oracle_replicas_run(self.h, int(n_rounds)) != 0:
            raise RuntimeError("oracle_replicas_run: " + self.L.oracle_last_error().decode())

    def moves(self):
        return self.L.oracle_replicas_num_moves(self.h)

    def num_replicas(self):
        return self.L.oracle_replicas_num_replicas(self.h)

    def rng(self):
        s = np.zeros(2, np.uint64)
        self.L.oracle_replicas_get_rng(self.h, _ptr(s, u64p))
        return int(s[0]), int(s[1])

    def median(self):
        e = np.zeros(4096)
        n = self.L.oracle_replicas_get_median(self.h, 4096, _ptr(e, f64p))
        return e[:n].copy()

    def replicas(self):
        out = (ZenoReplicaState * self.num_replicas())()
        self.L.oracle_replicas_get_replicas(self.h, out)
        return list(out)

    def system(self, replica):
        n = self.L.oracle_replicas_system_len(self.h)
        buf = np.zeros(n)
        assert self.L.oracle_replicas_get_system(self.h, replica, _ptr(buf, f64p), n) == 0
        return buf
