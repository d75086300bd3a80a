"""Helpers for the -m gpu parity tests: compare a GPU walker with the CPU oracle walker."""
import ctypes as C

import numpy as np

from sad_monte_carlo_b200 import _abi

SCALARS_EXACT = ["moves", "accepted_moves", "acceptance_rate", "translation_scale", "rng_s0", "rng_s1", "bins_min",
                 "bins_width", "bins_len", "method", "too_lo", "too_hi", "latest_parameter", "tL", "tF", "num_states",
                 "highest_hist", "samc_t0", "wl_gamma", "wl_num_states", "wl_min_energy", "wl_lowest_hist",
                 "wl_highest_hist", "wl_total_hist", "wl_hist_len", "max_S", "max_S_index"]
BINS_EXACT = ["histogram", "t_found", "lnw", "energy_total", "energy_squared_total", "round_trips", "have_visited",
              "wl_hist", "extra_total", "extra_count"]


def clone_config(cfg, **kw):
    c = _abi.Config()
    C.memmove(C.byref(c), C.byref(cfg), C.sizeof(cfg))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def method_fields(method):
    """Which sadmc_walker_state fields are meaningful for a given run-time method."""
    common = ["moves", "accepted_moves", "acceptance_rate", "translation_scale", "rng_s0", "rng_s1", "bins_min",
              "bins_width", "bins_len", "method", "max_S", "max_S_index"]
    if method == _abi.METHOD_SAD:
        return common + ["too_lo", "too_hi", "latest_parameter", "tL", "tF", "num_states", "highest_hist"]
    if method == _abi.METHOD_SAMC:
        return common + ["samc_t0"]
    if method in (_abi.METHOD_WL, _abi.METHOD_INV_T_WL):
        return common + ["wl_gamma", "wl_num_states", "wl_min_energy", "wl_lowest_hist", "wl_highest_hist",
                         "wl_total_hist", "wl_hist_len"]
    return common


def assert_walker_equal(eng, w, omc, exact=True, rtol=1e-12, check_system=True, context=""):
    """GPU walker w == oracle walker: every scalar and every per-bin vector, bit for bit when exact."""
    g, o = eng.walker(w), omc.walker()
    assert g.status == 0, "%s walker %d status %d" % (context, w, g.status)
    fields = method_fields(o.method)
    for f in fields:
        a, b = getattr(g, f), getattr(o, f)
        if exact or isinstance(a, int):
            assert a == b, "%s walker %d: %s gpu=%r oracle=%r" % (context, w, f, a, b)
        else:
            assert abs(a - b) <= rtol * max(1.0, abs(b)), "%s walker %d: %s gpu=%r oracle=%r" % (context, w, f, a, b)
    if exact:
        assert g.energy == o.energy, "%s walker %d energy %r vs %r" % (context, w, g.energy, o.energy)
    else:
        assert abs(g.energy - o.energy) <= rtol * max(1.0, abs(o.energy))
    gb, ob = eng.bins(w), omc.bins()
    for k in BINS_EXACT:
        if k == "wl_hist" and o.method not in (_abi.METHOD_WL, _abi.METHOD_INV_T_WL):
            continue
        if exact or gb[k].dtype != np.float64:
            assert np.array_equal(gb[k], ob[k]), "%s walker %d: bins.%s differ at %s" % (
                context, w, k, np.nonzero(gb[k] != ob[k])[0][:5])
        else:
            assert np.allclose(gb[k], ob[k], rtol=rtol, atol=rtol), "%s walker %d: bins.%s" % (context, w, k)
    if check_system:
        gs, osys = eng.system(w), omc.system()
        if exact:
            assert np.array_equal(gs, osys), "%s walker %d: system differs" % (context, w)
        else:
            assert np.allclose(gs, osys, rtol=rtol, atol=1e-300)
