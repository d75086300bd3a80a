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
@pytest.mark.parametrize("case", ["lj31", "ising_wl", "fake", "sw", "wca", "two_wells", "erfinv"])
def test_checkpoint_file_resume_equals_continuous(ext, case, tmp_path):
    if case == "lj31":
        cfg = lj_cfg()
    elif case == "ising_wl":
        cfg = make_config("ising", "inv-t-wl", N=8, min_allowed_energy=-128.0, max_allowed_energy=50.0, n_walkers=3, seed=2)
    elif case == "sw":  # tests/resume-sad.rs
        cfg = make_config("sw", "sad", N=64, filling_fraction=0.25, sad_min_T=0.5, n_walkers=3, seed=1,
                          move_plan=_abi.MOVE_ACCEPTANCE_RATE, move_value=0.5)
    elif case == "wca":  # the N*N-attempt start, relaxed below max_allowed_energy by from_params; `pressure` extras
        cfg = make_config("wca", "samc", N=20, reduced_density=0.4, energy_bin=1.0, n_walkers=3, seed=9, samc_t0=1e3,
                          max_allowed_energy=200.0)
    elif case == "two_wells":
        cfg = make_config("two-wells", "sad", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, sad_min_T=0.001,
                          energy_bin=1e-3, move_value=1e-2, n_walkers=3, seed=1)
    elif case == "erfinv":
        cfg = make_config("fake-erfinv", "samc", N=3, erfinv_mean_energy=0.0, samc_t0=1e3, energy_bin=0.05, n_walkers=3, seed=5,
                          bin_window_lo=-30.0, bin_window_hi=30.0)
    else:
        cfg = make_config("fake", "sad", fake_function=_abi.FAKE_QUADRATIC, N=3, sad_min_T=0.001, energy_bin=0.01,
                          n_walkers=4, seed=3, bin_window_lo=-2.5, bin_window_hi=4.0)
    if ext != "yaml" and case in ("sw", "wca", "two_wells", "erfinv"):
        pytest.skip("the three codecs are covered by the first three systems")
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
    # --resume-from: the engine configuration is rebuilt from the document alone
    doc = checkpoint.load(paths[0])
    tag = next(iter(doc["system"]))
    assert tag == checkpoint.SYSTEM_TAGS[cfg.system]
    over = {}
    if not _abi.isnan(cfg.bin_window_lo):
        over = dict(bin_window_lo=cfg.bin_window_lo, bin_window_hi=cfg.bin_window_hi)
    if cfg.system == _abi.SYS_LJ:
        over["lanes_per_walker"] = cfg.lanes_per_walker
    cfg2 = checkpoint.config_from_document(doc, n_walkers=cfg.n_walkers, **over)
    third = checkpoint.resume(cfg2, save_as)
    third.run(15000)
    same(full, third, "%s %s from the document" % (case, ext))


def test_system_documents_have_the_reference_field_names():
    # wca.rs:23-33 + optcell.rs:27-40 (subcells are #[serde(skip)]), optsquare.rs:24-31, two_wells.rs:219-232, erfinv.rs:29-38
    e = WalkerEngine(make_config("wca", "samc", N=20, reduced_density=0.4, samc_t0=1e3, max_allowed_energy=200.0, seed=9))
    d = checkpoint.walker_document(e, 0)["system"]["Wca"]
    assert set(d) == {"E", "error", "cell", "possible_change"} and d["possible_change"] == "None"
    assert set(d["cell"]) == {"box_diagonal", "r_cutoff", "positions"} and len(d["cell"]["positions"]) == 20
    assert d["cell"]["r_cutoff"] == 2.0 ** (1.0 / 6.0) and abs(d["cell"]["box_diagonal"]["x"] ** 3 - 50.0) < 1e-9
    e = WalkerEngine(make_config("sw", "sad", N=50, filling_fraction=0.3, sad_min_T=0.5))
    d = checkpoint.walker_document(e, 0)["system"]["Sw"]
    assert set(d) == {"E", "cell", "possible_change"} and d["cell"]["r_cutoff"] == 1.3
    e = WalkerEngine(make_config("two-wells", "sad", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, sad_min_T=0.001,
                                 energy_bin=1e-3, move_value=1e-2))
    d = checkpoint.walker_document(e, 0)["system"]["TwoWells"]
    assert set(d) == {"position", "d_squared", "parameters", "change", "well_position", "invcdf"}
    assert set(d["invcdf"]) == {"num_points", "dim", "r1", "r2", "dx1_ball1", "stencils"} and len(d["invcdf"]["stencils"]) == 120000
    st = np.array(d["invcdf"]["stencils"]).reshape(12, 10000)
    assert (st[:, 0] == 0).all() and np.allclose(st[:, -1], 1.0) and (np.diff(st, axis=1) >= 0).all()
    assert abs(st[5, 5000] - 0.5) < 1e-3  # the symmetric sphere coordinates: half the weight below 0
    e = WalkerEngine(make_config("fake-erfinv", "samc", N=3, erfinv_mean_energy=0.0, samc_t0=1e3, energy_bin=0.05,
                                 bin_window_lo=-30.0, bin_window_hi=30.0))
    d = checkpoint.walker_document(e, 0)["system"]["FakeErfinv"]
    assert set(d) == {"position", "parameters", "possible_change"} and d["parameters"] == {"mean_energy": 0.0}


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
