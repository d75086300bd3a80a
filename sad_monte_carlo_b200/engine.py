"""Host-side mirror of the reference's `EnergyMC` for a batch of walkers.

`WalkerEngine` is a thin, typed wrapper over the C ABI (include/sadmc_gpu.h):
the names follow the reference (`move_once` -> `run(n)`, `num_moves`,
`num_accepted_moves`, `system`, `bins`, `plan_move` / `confirm` / `energy` /
`compute_energy` / `verify_energy` of src/system/mod.rs:54-120).  All compute
happens in libsadmc_gpu.so; numpy is only used for host buffers.
"""
import ctypes as C

import numpy as np

from . import _abi, load_library
from ._abi import BinningState, Config, WalkerState
from ._capi import f64p, u64p, u8p


class SadmcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sadmc error %d: %s" % (code, msg))
        self.code = code


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class WalkerEngine:
    """`EnergyMC<Any>` x n_walkers on one GPU (reference: src/mc/energy.rs:167-210)."""

    def __init__(self, cfg: Config):
        self.L = load_library()
        self.cfg = cfg
        self.h = C.c_void_p()
        self._check(self.L.sadmc_create(C.byref(cfg), C.byref(self.h)))
        n = C.c_size_t()
        self._check(self.L.sadmc_system_len(self.h, C.byref(n)))
        self.system_len = n.value
        self.n_walkers = cfg.n_walkers

    # -- plumbing ---------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise SadmcError(rc, self.L.sadmc_last_error().decode())

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.L.sadmc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- the hot path -------------------------------------------------------
    def run(self, n_moves):
        """n_moves x move_once (energy.rs:904-974) for every walker; blocking."""
        self._check(self.L.sadmc_run(self.h, int(n_moves)))

    def run_async(self, n_moves):
        self._check(self.L.sadmc_run_async(self.h, int(n_moves)))

    def sync(self):
        self._check(self.L.sadmc_sync(self.h))

    def start(self):
        self._check(self.L.sadmc_start(self.h))

    def set_stream(self, cuda_stream_ptr):
        self._check(self.L.sadmc_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def last_run_ms(self):
        ms = C.c_float()
        self._check(self.L.sadmc_last_run_ms(self.h, C.byref(ms)))
        return ms.value

    def move_launch_shape(self):
        """(threads per CTA, threads per walker, shared bytes per CTA, bytes per walker streamed from L2) of the move kernel"""
        b, t, z = C.c_uint32(), C.c_uint32(), C.c_uint32()
        sm = C.c_uint64()
        self._check(self.L.sadmc_move_launch_shape(self.h, C.byref(b), C.byref(t), C.byref(sm), C.byref(z)))
        return b.value, t.value, sm.value, z.value

    def streams_z(self):
        return self.move_launch_shape()[3] != 0

    def launch_count(self):
        n = C.c_uint64()
        self._check(self.L.sadmc_launch_count(self.h, C.byref(n)))
        return n.value

    # -- state --------------------------------------------------------------
    def num_moves(self):
        n = C.c_uint64()
        self._check(self.L.sadmc_num_moves(self.h, C.byref(n)))
        return n.value

    def num_accepted_moves(self):
        n = C.c_uint64()
        self._check(self.L.sadmc_num_accepted_moves(self.h, C.byref(n)))
        return n.value

    def accepted_moves_range(self):
        """(min, max) of the per-walker accepted-move counts."""
        lo, hi = C.c_uint64(), C.c_uint64()
        self._check(self.L.sadmc_accepted_moves_range(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def num_halted(self):
        """(left the bin window, failed verify_energy) walker counts since creation."""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self.L.sadmc_num_halted(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def walker(self, w=0) -> WalkerState:
        s = WalkerState()
        self._check(self.L.sadmc_get_walker(self.h, w, C.byref(s)))
        return s

    def binning_walker(self, w=0) -> BinningState:
        """FLAG_BINNING engines: scalars of energy_binning.rs's EnergyMC / histogram::Bins of walker w."""
        s = BinningState()
        self._check(self.L.sadmc_get_binning_walker(self.h, w, C.byref(s)))
        return s

    def binning_bins(self, w=0):
        """FLAG_BINNING engines: bins.lnw and the `extra` accumulators of walker w, element 0 = bin at bins_min."""
        n = self.binning_walker(w).bins_len
        out = {"lnw_total": np.zeros(n), "lnw_count": np.zeros(n, np.uint64), "energy_total": np.zeros(n),
               "energy_count": np.zeros(n, np.uint64), "t_found_total": np.zeros(n), "t_found_count": np.zeros(n, np.uint64),
               "hist_count": np.zeros(n, np.uint64), "extra_total": np.zeros(n), "extra_count": np.zeros(n, np.uint64)}
        self._check(self.L.sadmc_get_binning_bins(
            self.h, w, n, _p(out["lnw_total"], f64p), _p(out["lnw_count"], u64p), _p(out["energy_total"], f64p),
            _p(out["energy_count"], u64p), _p(out["t_found_total"], f64p), _p(out["t_found_count"], u64p),
            _p(out["hist_count"], u64p), _p(out["extra_total"], f64p), _p(out["extra_count"], u64p)))
        return out

    def energies(self):
        e = np.zeros(self.n_walkers)
        self._check(self.L.sadmc_get_energies(self.h, _p(e, f64p)))
        return e

    def bins(self, w=0):
        n = self.walker(w).bins_len
        out = {
            "histogram": np.zeros(n, np.uint64), "t_found": np.zeros(n, np.uint64), "lnw": np.zeros(n),
            "energy_total": np.zeros(n), "energy_squared_total": np.zeros(n),
            "round_trips": np.zeros(n, np.uint64), "have_visited": np.zeros(n, np.uint8),
            "wl_hist": np.zeros(n, np.uint64), "extra_total": np.zeros(n), "extra_count": np.zeros(n, np.uint64),
        }
        self._check(self.L.sadmc_get_bins(
            self.h, w, n, _p(out["histogram"], u64p), _p(out["t_found"], u64p), _p(out["lnw"], f64p),
            _p(out["energy_total"], f64p), _p(out["energy_squared_total"], f64p), _p(out["round_trips"], u64p),
            _p(out["have_visited"], u8p), _p(out["wl_hist"], u64p), _p(out["extra_total"], f64p),
            _p(out["extra_count"], u64p)))
        return out

    def system(self, w=0):
        buf = np.zeros(self.system_len)
        self._check(self.L.sadmc_get_system(self.h, w, _p(buf, f64p), buf.size))
        return buf

    def set_system(self, w, buf):
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        self._check(self.L.sadmc_set_system(self.h, w, _p(buf, f64p), buf.size))

    def systems(self):
        buf = np.zeros((self.n_walkers, self.system_len))
        self._check(self.L.sadmc_get_systems(self.h, _p(buf, f64p), buf.size))
        return buf

    def set_systems(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        assert buf.size == self.n_walkers * self.system_len
        self._check(self.L.sadmc_set_systems(self.h, _p(buf, f64p), buf.size))

    def rngs(self):
        s = np.zeros((self.n_walkers, 2), np.uint64)
        self._check(self.L.sadmc_get_rngs(self.h, _p(s, u64p)))
        return s

    def set_rngs(self, s):
        s = np.ascontiguousarray(s, dtype=np.uint64)
        assert s.shape == (self.n_walkers, 2)
        self._check(self.L.sadmc_set_rngs(self.h, _p(s, u64p)))

    # -- resume (mc/mod.rs:70-84) ---------------------------------------------
    def set_walker_bins(self, w, state: WalkerState, bins):
        """Inverse of walker(w) + bins(w) on an engine created with INIT_EXTERNAL."""
        def arr(k, dt):
            a = bins.get(k)
            return None if a is None else np.ascontiguousarray(a, dtype=dt)
        keep = [arr("histogram", np.uint64), arr("t_found", np.uint64), arr("lnw", np.float64), arr("energy_total", np.float64),
                arr("energy_squared_total", np.float64), arr("round_trips", np.uint64), arr("have_visited", np.uint8),
                arr("wl_hist", np.uint64), arr("extra_total", np.float64), arr("extra_count", np.uint64)]
        types = [u64p, u64p, f64p, f64p, f64p, u64p, u8p, u64p, f64p, u64p]
        self._check(self.L.sadmc_set_walker_bins(self.h, w, C.byref(state), *[_p(a, t) for a, t in zip(keep, types)]))

    def binning_bins_f64(self, w=0):
        """As binning_bins with every count as f64 (FLAG_BINNING_LINEAR engines keep fractional counts)."""
        n = self.binning_walker(w).bins_len
        keys = ["lnw_total", "lnw_count", "energy_total", "energy_count", "t_found_total", "t_found_count", "hist_count", "extra_total", "extra_count"]
        out = {k: np.zeros(n) for k in keys}
        self._check(self.L.sadmc_get_binning_bins_f64(self.h, w, n, *[_p(out[k], f64p) for k in keys]))
        return out

    def high_resolution(self, w=0):
        """FLAG_BINNING engines created with high_resolution_de: (Bins::min, counts) of the finer histogram of walker w."""
        mn, n = C.c_double(), C.c_uint32()
        self._check(self.L.sadmc_get_high_resolution(self.h, w, 0, C.byref(mn), C.byref(n), None))
        cnt = np.zeros(n.value, np.uint64)
        self._check(self.L.sadmc_get_high_resolution(self.h, w, n.value, C.byref(mn), C.byref(n), _p(cnt, u64p)))
        return mn.value, cnt

    def set_high_resolution(self, w, bins_min, counts):
        cnt = np.ascontiguousarray(counts, dtype=np.uint64)
        self._check(self.L.sadmc_set_high_resolution(self.h, w, float(bins_min), cnt.size, _p(cnt, u64p) if cnt.size else None))

    def set_binning_walker(self, w, state: BinningState, bins):
        """FLAG_BINNING engines: inverse of binning_walker(w) + binning_bins(w) on an engine created with INIT_EXTERNAL."""
        def arr(k, dt):
            a = bins.get(k)
            return None if a is None else np.ascontiguousarray(a, dtype=dt)
        keys = [("lnw_total", np.float64, f64p), ("lnw_count", np.uint64, u64p), ("energy_total", np.float64, f64p),
                ("energy_count", np.uint64, u64p), ("t_found_total", np.float64, f64p), ("t_found_count", np.uint64, u64p),
                ("hist_count", np.uint64, u64p), ("extra_total", np.float64, f64p), ("extra_count", np.uint64, u64p)]
        keep = [arr(k, dt) for k, dt, _ in keys]
        self._check(self.L.sadmc_set_binning_walker(self.h, w, C.byref(state), *[_p(a, t) for a, (_, _, t) in zip(keep, keys)]))

    def resume(self, moves):
        """Instead of start(): continue from restored walkers at move count `moves`."""
        self._check(self.L.sadmc_resume(self.h, int(moves)))

    def set_lnw(self, lnw_window):
        """Every walker's ln w := the window-aligned array (fixed weights: use with method "samc", samc_t0 = 0)."""
        a = np.ascontiguousarray(lnw_window, dtype=np.float64)
        self._check(self.L.sadmc_set_lnw(self.h, _p(a, f64p), a.size))

    def window(self):
        lo, width, n = C.c_double(), C.c_double(), C.c_uint32()
        self._check(self.L.sadmc_window(self.h, C.byref(lo), C.byref(width), C.byref(n)))
        return lo.value, width.value, n.value

    def cell_box(self):
        """(box_diagonal[3], r_cutoff) of a periodic fluid: `Cell` of optcell.rs:27-33 as the engine derived it."""
        box, rc = (C.c_double * 3)(), C.c_double()
        self._check(self.L.sadmc_cell_box(self.h, box, C.byref(rc)))
        return [box[0], box[1], box[2]], rc.value

    def fold_select(self, first_walker=0, walker_stride=1, sad_range_only=False, walker_count=0):
        """Which walkers the following folds merge (interleaved groups: ensemble error bars; stride 1 + walker_count: a
        contiguous shard) and whether SAD walkers contribute ln w only inside their own [too_lo, too_hi]
        (True / 1) or only strictly inside it (2: without the two half-updated end bins)."""
        self._check(self.L.sadmc_fold_select_ex(self.h, first_walker, walker_stride, walker_count, int(sad_range_only)))

    def fold_settled(self, tl_max=0):
        """SAD walkers contribute ln w to the following folds only if their range has not changed since move tl_max
        (Sad::tL <= tl_max); 0 = every walker."""
        self._check(self.L.sadmc_fold_settled(self.h, int(tl_max)))

    def fold(self):
        """Window-aligned sums over the local walkers (device fold kernel), as host arrays."""
        _, _, n = self.window()
        out = {"histogram": np.zeros(n, np.uint64), "energy_total": np.zeros(n), "energy_squared_total": np.zeros(n),
               "lnw_sum": np.zeros(n), "lnw_sq_sum": np.zeros(n), "lnw_count": np.zeros(n, np.uint64)}
        self._check(self.L.sadmc_fold(self.h, _p(out["histogram"], u64p), _p(out["energy_total"], f64p),
                                      _p(out["energy_squared_total"], f64p), _p(out["lnw_sum"], f64p),
                                      _p(out["lnw_sq_sum"], f64p), _p(out["lnw_count"], u64p)))
        return out

    def fold_device(self, hist, etot, e2tot, lnw_sum, lnw_sq, lnw_cnt):
        """Same, into caller-owned DEVICE buffers given as raw pointers (e.g. torch tensors' data_ptr())."""
        self._check(self.L.sadmc_fold_device(self.h, *[C.c_void_p(p) for p in (hist, etot, e2tot, lnw_sum, lnw_sq, lnw_cnt)]))

    def fold_packed_device(self, packed):
        """The fold as one DEVICE buffer of 7 x nbins doubles (raw pointer), for a single collective."""
        self._check(self.L.sadmc_fold_packed_device(self.h, C.c_void_p(packed)))

    # -- trait-shaped shims (src/system/mod.rs:54-120) -------------------------
    def randomize(self, w=0):
        """System::randomize driven by walker w's generator; returns the new energy."""
        e = C.c_double()
        self._check(self.L.sadmc_sys_randomize(self.h, w, C.byref(e)))
        return e.value

    def energy(self, w=0):
        e = C.c_double()
        self._check(self.L.sadmc_sys_energy(self.h, w, C.byref(e)))
        return e.value

    def compute_energy(self, w=0):
        e = C.c_double()
        self._check(self.L.sadmc_sys_compute_energy(self.h, w, C.byref(e)))
        return e.value

    def plan_move(self, w, mean_distance):
        some, e = C.c_int(), C.c_double()
        self._check(self.L.sadmc_sys_plan_move(self.h, w, mean_distance, C.byref(some), C.byref(e)))
        return e.value if some.value else None

    def confirm(self, w=0):
        self._check(self.L.sadmc_sys_confirm(self.h, w))

    def verify_energy(self, w=0):
        rc = self.L.sadmc_sys_verify_energy(self.h, w)
        if rc == _abi.ERR_VERIFY:
            return False
        self._check(rc)
        return True


EnergyMC = WalkerEngine
