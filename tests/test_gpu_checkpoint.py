"""Checkpoint files in the reference's serde schema (yaml / json / cbor) written from GPU walkers, read back, and
resumed: the continuation equals the uninterrupted run bit for bit; the documents carry what the reference's Python
tools index (plotting/parse-binning.py:103-170); the plugin loop drives the engine with one launch per period."""
import os

import numpy as np
import pytest

from sad_monte_carlo_b200 import WalkerEngine, analysis, checkpoint, make_config, plugins, _abi
from tests.gpu_common import BINS_EXACT, method_fields

pytestmark = pytest.mark.gpu


def lj_cfg(**kw):
    base = dict(N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01, n_walkers=6,
                init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=1, bin_window_lo=-133.62, bin_window_hi=0.02)
    base.update(kw)
    return make_config("lj", "sad", **base)


def same(a, b, ctx):
    assert a.num_moves() == b.num_moves()
    assert np.array_equal(a.rngs(), b.rngs()) and np.array_equal(a.systems(), b.systems()), ctx
    for w in range(a.n_walkers):
        wa, wb = a.walker(w), b.walker(w)
        for f in method_fields(wa.method) + ["energy"]:
            assert getattr(wa, f) == getattr(wb, f), (ctx, w, f)
        ba, bb = a.bins(w), b.bins(w)
        for k in BINS_EXACT:
            assert np.array_equal(ba[k], bb[k]), (ctx, w, k)


@pytest.mark.parametrize("ext", ["yaml", "json", "cbor"])
@pytest.mark.parametrize("case", ["lj31", "ising_wl", "fake"])
def test_checkpoint_file_resume_equals_continuous(ext, case, tmp_path):
    if case == "lj31":
        cfg = lj_cfg()
    elif case == "ising_wl":
        cfg = make_config("ising", "inv-t-wl", N=8, min_allowed_energy=-128.0, max_allowed_energy=50.0, n_walkers=3, seed=2)
    else:
        cfg = make_config("fake", "sad", fake_function=_abi.FAKE_QUADRATIC, N=3, sad_min_T=0.001, energy_bin=0.01,
                          n_walkers=4, seed=3, bin_window_lo=-2.5, bin_window_hi=4.0)
    full = WalkerEngine(cfg)
    full.run(15000)
    full.run(15000)
    first = WalkerEngine(cfg)
    first.run(15000)
    save_as = str(tmp_path / ("run." + ext))
    paths = checkpoint.save(first, save_as)
    assert len(paths) == cfg.n_walkers and all(os.path.exists(p) for p in paths)
    first.close()
    second = checkpoint.resume(cfg, save_as)
    second.run(15000)
    same(full, second, "%s %s" % (case, ext))


def test_document_has_the_reference_schema_and_feeds_the_reference_post_processing(tmp_path):
    eng = WalkerEngine(lj_cfg(n_walkers=2))
    eng.run(40000)
    doc = checkpoint.walker_document(eng, 1, save_as="sad-lj31.yaml")
    # EnergyMC fields, src/mc/energy.rs:167-210
    assert set(doc) == {"system", "method", "moves", "time_L", "accepted_moves", "min_allowed_energy", "max_allowed_energy",
                        "move_plan", "translation_scale", "acceptance_rate", "rng", "save_as", "report", "movies", "save",
                        "manager", "bins", "have_visited_since_maxentropy", "round_trips", "max_S", "max_S_index"}
    assert set(doc["bins"]) == {"min", "width", "histogram", "t_found", "lnw", "energy_total", "energy_squared_total", "extra"}
    assert set(doc["method"]["Sad"]) == {"min_T", "too_lo", "too_hi", "tL", "tF", "num_states", "highest_hist", "version",
                                         "latest_parameter"}
    lj = doc["system"]["Lj"]
    assert set(lj) == {"E", "error", "possible_change", "positions", "max_radius_squared", "max_radius"}
    assert len(lj["positions"]) == 31 and set(lj["positions"][0]) == {"x", "y", "z"}
    assert doc["moves"] == 40000 and doc["move_plan"] == {"TranslationScale": 0.05} and doc["max_allowed_energy"] == 0.0
    assert doc["min_allowed_energy"] is None and set(doc["rng"]) == {"s0", "s1"}
    # what plotting/parse-binning.py:150-169 does with such a file
    p = tmp_path / "sad-lj31.yaml"
    checkpoint.write_atomic(str(p), checkpoint.dumps(doc, "yaml"))
    data = checkpoint.load(str(p))
    b, m = data["bins"], data["method"]["Sad"]
    E = b["min"] + (np.arange(len(b["lnw"])) + 0.5) * b["width"]
    s = analysis.sad_excess_entropy(b["lnw"], b["histogram"], E, m["too_lo"], m["too_hi"], m["min_T"])
    assert np.isfinite(s).all() and s.max() == 0.0 and sum(b["histogram"]) == 40001
    st = eng.walker(1)
    assert (m["too_lo"], m["too_hi"], m["tF"]) == (st.too_lo, st.too_hi, st.tF)


def test_plugin_loop_runs_one_launch_per_period_and_checkpoints_on_schedule(tmp_path):
    cfg = make_config("ising", "sad", N=8, sad_min_T=1.0, n_walkers=4, seed=1)
    eng = WalkerEngine(cfg)
    save_as = str(tmp_path / "ising.json")
    launches0 = eng.launch_count()
    n = plugins.run_simulation(eng, plugins.Report(max_iter=20000), plugins.Save(save_time_hours=None),
                               plugins.Movie(movie_time=4.0), save_as=save_as, checkpoint_walkers=[0, 3])
    assert eng.num_moves() == 20000
    # Save doubles (1, 2, 4, ...), movie frames at powers of 4, one launch per period: a few dozen launches for 2e4 moves
    assert n <= 40
    frames = sorted(os.listdir(str(tmp_path / "ising")))
    assert [f for f in frames if f.endswith("-w000000.cbor")] == ["%014d-w000000.cbor" % (4 ** k) for k in range(8)]
    final = checkpoint.load(checkpoint.walker_path(save_as, 3, 4))
    assert final["moves"] == 20000 and final["report"]["max_iter"] == {"TotalMoves": 20000}
    # the plugin loop changes nothing about the trajectory
    ref = WalkerEngine(cfg)
    ref.run(20000)
    same(ref, eng, "plugin loop")
    assert eng.launch_count() - launches0 >= n
