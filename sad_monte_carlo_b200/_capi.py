"""ctypes prototypes for every symbol include/sadmc_gpu.h declares."""
import ctypes as C

from ._abi import BinningState, Config, ReplicaState, WalkerState, ZenoReplicaState

u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p

# name -> (restype, argtypes); the list is also what tests/test_abi.py checks against the header
PROTOTYPES = {
    "sadmc_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
    "sadmc_destroy": (None, [vp]),
    "sadmc_last_error": (C.c_char_p, []),
    "sadmc_abi_version": (C.c_int, []),
    "sadmc_reference_system": (C.c_int, [C.POINTER(Config), f64p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sadmc_start": (C.c_int, [vp]),
    "sadmc_set_stream": (C.c_int, [vp, vp]),
    "sadmc_get_stream": (vp, [vp]),
    "sadmc_run": (C.c_int, [vp, C.c_uint64]),
    "sadmc_run_async": (C.c_int, [vp, C.c_uint64]),
    "sadmc_sync": (C.c_int, [vp]),
    "sadmc_last_run_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "sadmc_launch_count": (C.c_int, [vp, u64p]),
    "sadmc_move_launch_shape": (C.c_int, [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), u64p, C.POINTER(C.c_uint32)]),
    "sadmc_num_moves": (C.c_int, [vp, u64p]),
    "sadmc_num_accepted_moves": (C.c_int, [vp, u64p]),
    "sadmc_num_halted": (C.c_int, [vp, u64p, u64p]),
    "sadmc_accepted_moves_range": (C.c_int, [vp, u64p, u64p]),
    "sadmc_get_walker": (C.c_int, [vp, C.c_uint32, C.POINTER(WalkerState)]),
    "sadmc_get_energies": (C.c_int, [vp, f64p]),
    "sadmc_get_bins": (C.c_int, [vp, C.c_uint32, C.c_uint32, u64p, u64p, f64p, f64p, f64p, u64p, u8p, u64p, f64p, u64p]),
    "sadmc_get_binning_walker": (C.c_int, [vp, C.c_uint32, C.POINTER(BinningState)]),
    "sadmc_get_binning_bins": (C.c_int, [vp, C.c_uint32, C.c_uint32, f64p, u64p, f64p, u64p, f64p, u64p, u64p, f64p, u64p]),
    "sadmc_get_binning_bins_f64": (C.c_int, [vp, C.c_uint32, C.c_uint32, f64p, f64p, f64p, f64p, f64p, f64p, f64p, f64p, f64p]),
    "sadmc_get_high_resolution": (C.c_int, [vp, C.c_uint32, C.c_uint32, f64p, C.POINTER(C.c_uint32), u64p]),
    "sadmc_set_high_resolution": (C.c_int, [vp, C.c_uint32, C.c_double, C.c_uint32, u64p]),
    "sadmc_set_binning_walker": (C.c_int, [vp, C.c_uint32, C.POINTER(BinningState), f64p, u64p, f64p, u64p, f64p, u64p, u64p, f64p, u64p]),
    "sadmc_system_len": (C.c_int, [vp, C.POINTER(C.c_size_t)]),
    "sadmc_get_system": (C.c_int, [vp, C.c_uint32, f64p, C.c_size_t]),
    "sadmc_set_system": (C.c_int, [vp, C.c_uint32, f64p, C.c_size_t]),
    "sadmc_get_systems": (C.c_int, [vp, f64p, C.c_size_t]),
    "sadmc_set_systems": (C.c_int, [vp, f64p, C.c_size_t]),
    "sadmc_get_rngs": (C.c_int, [vp, u64p]),
    "sadmc_set_rngs": (C.c_int, [vp, u64p]),
    "sadmc_set_walker_bins": (C.c_int, [vp, C.c_uint32, C.POINTER(WalkerState), u64p, u64p, f64p, f64p, f64p, u64p, u8p, u64p, f64p, u64p]),
    "sadmc_resume": (C.c_int, [vp, C.c_uint64]),
    "sadmc_set_lnw": (C.c_int, [vp, f64p, C.c_uint32]),
    "sadmc_window": (C.c_int, [vp, f64p, f64p, C.POINTER(C.c_uint32)]),
    "sadmc_cell_box": (C.c_int, [vp, f64p, f64p]),
    "sadmc_fold_select": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_int]),
    "sadmc_fold_select_ex": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
    "sadmc_fold_settled": (C.c_int, [vp, C.c_uint64]),
    "sadmc_fold_device": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "sadmc_fold_packed_device": (C.c_int, [vp, vp]),
    "sadmc_fold": (C.c_int, [vp, u64p, f64p, f64p, f64p, f64p, u64p]),
    "sadmc_sys_energy": (C.c_int, [vp, C.c_uint32, f64p]),
    "sadmc_sys_compute_energy": (C.c_int, [vp, C.c_uint32, f64p]),
    "sadmc_sys_plan_move": (C.c_int, [vp, C.c_uint32, C.c_double, C.POINTER(C.c_int), f64p]),
    "sadmc_sys_confirm": (C.c_int, [vp, C.c_uint32]),
    "sadmc_sys_randomize": (C.c_int, [vp, C.c_uint32, f64p]),
    "sadmc_sys_verify_energy": (C.c_int, [vp, C.c_uint32]),
    "sadmc_tempering_create": (C.c_int, [C.POINTER(Config), f64p, C.c_uint32, C.c_uint64, C.POINTER(vp)]),
    "sadmc_tempering_destroy": (None, [vp]),
    "sadmc_tempering_run": (C.c_int, [vp, C.c_uint64]),
    "sadmc_tempering_num_moves": (C.c_int, [vp, u64p]),
    "sadmc_tempering_steps_per_round": (C.c_int, [vp, u64p]),
    "sadmc_tempering_get_replicas": (C.c_int, [vp, C.c_uint32, C.POINTER(ReplicaState)]),
    "sadmc_tempering_get_rng": (C.c_int, [vp, C.c_uint32, u64p]),
    "sadmc_tempering_set_replicas": (C.c_int, [vp, C.c_uint32, C.POINTER(ReplicaState)]),
    "sadmc_tempering_set_rng": (C.c_int, [vp, C.c_uint32, u64p]),
    "sadmc_tempering_set_num_moves": (C.c_int, [vp, C.c_uint64]),
    "sadmc_tempering_set_translation_scales": (C.c_int, [vp, f64p]),
    "sadmc_tempering_system_len": (C.c_int, [vp, C.POINTER(C.c_size_t)]),
    "sadmc_tempering_get_system": (C.c_int, [vp, C.c_uint32, C.c_uint32, f64p, C.c_size_t]),
    "sadmc_tempering_set_system": (C.c_int, [vp, C.c_uint32, C.c_uint32, f64p, C.c_size_t]),
    "sadmc_tempering_cell_box": (C.c_int, [vp, f64p, f64p]),
    "sadmc_tempering_last_run_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "sadmc_replicas_create": (C.c_int, [C.POINTER(Config), C.c_double, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    "sadmc_replicas_destroy": (None, [vp]),
    "sadmc_replicas_run": (C.c_int, [vp, C.c_uint64]),
    "sadmc_replicas_num_moves": (C.c_int, [vp, C.c_uint32, u64p]),
    "sadmc_replicas_num_replicas": (C.c_int, [vp, C.c_uint32, C.POINTER(C.c_uint32)]),
    "sadmc_replicas_get_replicas": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.POINTER(ZenoReplicaState)]),
    "sadmc_replicas_get_rng": (C.c_int, [vp, C.c_uint32, u64p]),
    "sadmc_replicas_get_median": (C.c_int, [vp, C.c_uint32, C.c_uint32, f64p, C.POINTER(C.c_uint32)]),
    "sadmc_replicas_system_len": (C.c_int, [vp, C.POINTER(C.c_size_t)]),
    "sadmc_replicas_get_system": (C.c_int, [vp, C.c_uint32, C.c_uint32, f64p, C.c_size_t]),
    "sadmc_replicas_last_run_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "sadmc_measure_fp64_peak": (C.c_int, [C.c_int, C.c_int, f64p]),
    "sadmc_selftest_exp_cmp": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, u64p, u64p]),
}


def bind(lib):
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    lib.sadmc_sizeof_config.restype = C.c_size_t
    lib.sadmc_sizeof_walker_state.restype = C.c_size_t
    lib.sadmc_sizeof_binning_state.restype = C.c_size_t
    lib.sadmc_sizeof_replica_state.restype = C.c_size_t
    lib.sadmc_sizeof_zeno_replica_state.restype = C.c_size_t
    return lib
