"""ctypes mirror of include/sadmc_gpu.h (structs, enums, prototypes).

Kept field-for-field in step with the header; tests/test_abi.py checks the
struct sizes against the compiled library (`sadmc_sizeof_*`).
"""
import ctypes as C
import math

ABI_VERSION = 1

# sadmc_system_kind  (AnyParams variants, reference src/system/any.rs:10-27)
SYS_FAKE, SYS_FAKE_ERFINV, SYS_WCA, SYS_LJ, SYS_ISING, SYS_SW, SYS_TWO_WELLS = 1, 2, 3, 4, 6, 7, 8
SYSTEM_NAMES = {"fake": SYS_FAKE, "fake-erfinv": SYS_FAKE_ERFINV, "wca": SYS_WCA, "lj": SYS_LJ,
                "ising": SYS_ISING, "sw": SYS_SW, "two-wells": SYS_TWO_WELLS}
# sadmc_fake_function (reference src/system/fake.rs:12-36)
FAKE_LINEAR, FAKE_QUADRATIC, FAKE_PIECES, FAKE_GAUSSIAN = 0, 1, 2, 3
# sadmc_method_kind (reference src/mc/energy.rs:44-69)
METHOD_SAD, METHOD_SAMC, METHOD_WL, METHOD_INV_T_WL, METHOD_CANONICAL = 1, 2, 3, 4, 5
METHOD_NAMES = {"sad": METHOD_SAD, "samc": METHOD_SAMC, "wl": METHOD_WL, "inv-t-wl": METHOD_INV_T_WL,
                "canonical": METHOD_CANONICAL}
MOVE_TRANSLATION_SCALE, MOVE_ACCEPTANCE_RATE = 0, 1
INIT_REFERENCE, INIT_RANDOMIZE, INIT_EXTERNAL = 0, 1, 2
FLAG_NO_ROUND_TRIPS = 1
FLAG_SUM_TREE = 2
FLAG_FAST_MATH = 4
FLAG_BINNING = 16  # energy_binning.rs bookkeeping over binning::histogram (the `binning` binary)
FLAG_BINNING_LINEAR = 32  # with FLAG_BINNING: binning::linear (interpolated ln w, f64 counts)
FLAG_LJ_SMEM_Z = 64  # LJ31 fast tier: force all coordinates in shared memory (2 CTAs per SM) ...
FLAG_LJ_STREAM_Z = 128  # ... or force z streamed from L2 (3 CTAs per SM); default: by walker count.  Same results bit for bit.
FLAG_HELPER_WARPS = 8  # experiment: helper warps for the LJ pair loop (with FLAG_FAST_MATH, lanes_per_walker = 1)

OK, ERR_INVALID, ERR_CUDA, ERR_WINDOW, ERR_UNSUPPORTED, ERR_VERIFY = 0, -1, -2, -3, -4, -5

NAN = float("nan")


class Config(C.Structure):
    """struct sadmc_config"""
    _fields_ = [
        ("abi_version", C.c_uint32), ("system", C.c_int32),
        ("N", C.c_uint32),
        ("lj_radius", C.c_double), ("reduced_density", C.c_double), ("filling_fraction", C.c_double),
        ("cell_width", C.c_double * 3), ("sw_well_width", C.c_double),
        ("fake_function", C.c_int32), ("_pad0", C.c_int32),
        ("fake_a", C.c_double), ("fake_b", C.c_double), ("fake_e1", C.c_double), ("fake_e2", C.c_double),
        ("fake_sigma", C.c_double),
        ("tw_h2_to_h1", C.c_double), ("tw_barrier_over_h1", C.c_double), ("tw_r2", C.c_double),
        ("erfinv_mean_energy", C.c_double),
        ("method", C.c_int32), ("move_plan", C.c_int32),
        ("sad_min_T", C.c_double), ("samc_t0", C.c_double), ("wl_min_gamma", C.c_double),
        ("canonical_T", C.c_double),
        ("seed", C.c_uint64),
        ("energy_bin", C.c_double), ("min_allowed_energy", C.c_double), ("max_allowed_energy", C.c_double),
        ("move_value", C.c_double),
        ("n_walkers", C.c_uint32), ("walker_offset", C.c_uint32), ("device", C.c_int32), ("init_mode", C.c_int32),
        ("bin_window_lo", C.c_double), ("bin_window_hi", C.c_double),
        ("lanes_per_walker", C.c_int32), ("flags", C.c_uint32),
        ("high_resolution_de", C.c_double),
    ]


class WalkerState(C.Structure):
    """struct sadmc_walker_state"""
    _fields_ = [
        ("moves", C.c_uint64), ("accepted_moves", C.c_uint64),
        ("acceptance_rate", C.c_double), ("translation_scale", C.c_double),
        ("rng_s0", C.c_uint64), ("rng_s1", C.c_uint64),
        ("energy", C.c_double), ("bins_min", C.c_double), ("bins_width", C.c_double),
        ("bins_len", C.c_uint32), ("window_first", C.c_uint32),
        ("method", C.c_int32), ("status", C.c_int32),
        ("too_lo", C.c_double), ("too_hi", C.c_double), ("latest_parameter", C.c_double),
        ("tL", C.c_uint64), ("tF", C.c_uint64), ("num_states", C.c_uint64), ("highest_hist", C.c_uint64),
        ("samc_t0", C.c_double),
        ("wl_gamma", C.c_double), ("wl_num_states", C.c_double), ("wl_min_energy", C.c_double),
        ("wl_lowest_hist", C.c_uint64), ("wl_highest_hist", C.c_uint64), ("wl_total_hist", C.c_uint64),
        ("wl_hist_len", C.c_uint32), ("wl_inv_t", C.c_int32),
        ("max_S", C.c_double), ("max_S_index", C.c_uint32), ("_pad", C.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("_")}


class BinningState(C.Structure):
    """struct sadmc_binning_state (FLAG_BINNING engines)"""
    _fields_ = [
        ("moves", C.c_uint64), ("accepted_moves", C.c_uint64),
        ("acceptance_rate", C.c_double), ("translation_scale", C.c_double),
        ("rng_s0", C.c_uint64), ("rng_s1", C.c_uint64),
        ("energy", C.c_double), ("bins_min", C.c_double), ("bins_width", C.c_double),
        ("bins_min_e", C.c_double), ("bins_max_e", C.c_double),
        ("bins_len", C.c_uint32), ("window_first", C.c_uint32),
        ("method", C.c_int32), ("status", C.c_int32),
        ("too_lo", C.c_double), ("too_hi", C.c_double), ("latest_parameter", C.c_double), ("tF", C.c_double),
        ("tL", C.c_uint64), ("num_states", C.c_uint64),
        ("samc_t0", C.c_double), ("wl_gamma", C.c_double),
        ("wl_inv_t", C.c_int32), ("_pad", C.c_int32),
        ("lnw_max_count", C.c_uint64), ("lnw_total_count", C.c_uint64),
        ("t_found_max_total", C.c_double),
        ("hist_min_count", C.c_uint64), ("hist_total_count", C.c_uint64),
        ("lnw_max_count_f64", C.c_double), ("hist_min_count_f64", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("_")}


class ReplicaState(C.Structure):
    """struct sadmc_replica_state (`Replica` of src/mc/tempering.rs:46-73 without its system)"""
    _fields_ = [
        ("T", C.c_double),
        ("rejected_count", C.c_uint64), ("accepted_count", C.c_uint64), ("rejected_swap_count", C.c_uint64),
        ("accepted_swap_count", C.c_uint64), ("ignored_count", C.c_uint64),
        ("total_energy", C.c_double), ("total_energy_squared", C.c_double), ("translation_scale", C.c_double),
        ("rng_s0", C.c_uint64), ("rng_s1", C.c_uint64), ("energy", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ZenoReplicaState(C.Structure):
    """struct sadmc_zeno_replica_state (`Replica` of src/mc/energy_replicas.rs:103-145 without its system)"""
    _fields_ = [
        ("max_energy", C.c_double), ("cutoff_energy", C.c_double), ("lowest_max_energy", C.c_double), ("translation_scale", C.c_double),
        ("rejected_count", C.c_uint64), ("accepted_count", C.c_uint64), ("above_count", C.c_uint64), ("below_count", C.c_uint64),
        ("upwelling_count", C.c_uint64), ("unique_visitors", C.c_uint64),
        ("above_total", C.c_double), ("below_total", C.c_double), ("above_total_squared", C.c_double), ("below_total_squared", C.c_double),
        ("above_extra_total", C.c_double), ("above_extra_count", C.c_uint64),
        ("collecting_data", C.c_int32), ("_pad", C.c_int32),
        ("rng_s0", C.c_uint64), ("rng_s1", C.c_uint64), ("energy", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("_")}


def make_config(system, method="sad", **kw):
    """Build a Config with the reference's defaults (EnergyMCParams::default, energy.rs:99-115)."""
    c = Config()
    c.abi_version = ABI_VERSION
    c.system = SYSTEM_NAMES[system] if isinstance(system, str) else int(system)
    c.method = METHOD_NAMES[method] if isinstance(method, str) else int(method)
    c.sad_min_T = 0.2
    c.wl_min_gamma = NAN
    c.energy_bin = NAN
    c.min_allowed_energy = NAN
    c.max_allowed_energy = NAN
    c.move_plan = MOVE_TRANSLATION_SCALE
    c.move_value = 0.05
    c.n_walkers = 1
    c.bin_window_lo = NAN
    c.bin_window_hi = NAN
    c.init_mode = INIT_REFERENCE
    c.high_resolution_de = NAN
    # per-system defaults of the reference
    c.reduced_density = 1.0      # WcaNParams::default, wca.rs:380-388
    c.filling_fraction = 0.3     # SquareWellNParams::default, optsquare.rs:347-355
    c.sw_well_width = 1.3
    if c.system in (SYS_WCA, SYS_SW):
        c.N = 100
    for k, v in kw.items():
        if k == "cell_width":
            for i in range(3):
                c.cell_width[i] = float(v[i])
        elif not hasattr(c, k):
            raise AttributeError("sadmc_config has no field %r" % k)
        else:
            setattr(c, k, v)
    return c


def isnan(x):
    return isinstance(x, float) and math.isnan(x)
