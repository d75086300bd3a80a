"""The `histogram` command line: flag spelling and expression-valued numbers as auto_args
derives them from the reference's parameter structs, checked on the command lines the reference's own job scripts and
tests use.  `--dry-run` needs no GPU."""
import io
import json
import math

import pytest

from sad_monte_carlo_b200 import _abi, histogram


def dry(argv):
    lines = []
    assert histogram.main(argv + ["--dry-run"], out=lines.append) == 0
    return json.loads(lines[-1])


def test_expressions():
    ev = histogram.evaluate
    assert ev("10^(1/8)") == 10 ** (1 / 8)          # lj/run-lj.py:14 --movie-time '10^(1/8)'
    assert ev("1/3") == 1 / 3
    assert ev("1e9") == 1e9 and ev("2^-11") == 2.0 ** -11 and ev("-133.53") == -133.53
    assert ev("2^3^2") == 2.0 ** 9                   # right-associative
    assert ev("sqrt(2)*pi") == math.sqrt(2) * math.pi
    assert ev(" ( 1 + 2 ) * 3 - 4 / 8 ") == 8.5
    for bad in ("", "1/0", "2*", "foo", "3)", "1 2"):
        with pytest.raises(histogram.ExprError):
            ev(bad)


def test_lj31_sad_job_of_run_lj_clusters_sh():
    # run-lj-clusters.sh:53 (the `lj-cluster` binary spells --N/--radius; `histogram` prefixes the Any variant: lj/run-lj.py:71)
    d = dry("--lj-N 31 --max-allowed-energy=0 --sad-min-T 0.01 --translation-scale 0.05 --energy-bin 0.01 --save-as lj-sad-31-bin001.yaml "
            "--movie-time 10^(1/8) --save-time 0.5 --lj-radius 2.5 --seed=3".split())
    c = d["config"]
    assert (c["system"], c["N"], c["lj_radius"], c["method"]) == (_abi.SYS_LJ, 31, 2.5, _abi.METHOD_SAD)
    assert (c["sad_min_T"], c["energy_bin"], c["max_allowed_energy"], c["min_allowed_energy"]) == (0.01, 0.01, 0.0, None)
    assert (c["move_plan"], c["move_value"], c["seed"], c["n_walkers"]) == (_abi.MOVE_TRANSLATION_SCALE, 0.05, 3, 1)
    assert d["plugins"] == dict(max_iter=None, max_independent_samples=None, quiet=False, save_time=0.5, movie_time=10 ** 0.125)
    assert d["save_as"] == "lj-sad-31-bin001.yaml" and d["resuming"] is False


def test_square_well_flags_of_the_resume_test():
    # tests/resume-sad.rs:29-43
    d = dry(["--sw-N=100", "--sw-filling-fraction=0.3", "--sw-well-width=1.3", "--sad-min-T=0.5", "--acceptance-rate=0.5",
             "--max-iter=1000", "--save-as=big-guy.yaml"])
    c = d["config"]
    assert (c["system"], c["N"], c["filling_fraction"], c["sw_well_width"]) == (_abi.SYS_SW, 100, 0.3, 1.3)
    assert (c["move_plan"], c["move_value"], c["sad_min_T"]) == (_abi.MOVE_ACCEPTANCE_RATE, 0.5, 0.5)
    assert d["plugins"]["max_iter"] == 1000 and d["plugins"]["save_time"] == 1.0  # SaveParams::default: one hour


@pytest.mark.parametrize("argv,expect", [
    ("--ising-N 32 --wl --wl-min-gamma=1e-4 --min-allowed-energy=-2048 --max-allowed-energy=50",      # ising-wl-min-gamma.sh:8
     dict(system=_abi.SYS_ISING, N=32, method=_abi.METHOD_WL, wl_min_gamma=1e-4, min_allowed_energy=-2048.0, max_allowed_energy=50.0)),
    ("--lj-N 31 --lj-radius 2.5 --Inv-t-WL --min-allowed-energy=-133.53 --max-allowed-energy=-110 --energy-bin 0.001",  # run-lj-clusters.sh:9
     dict(system=_abi.SYS_LJ, method=_abi.METHOD_INV_T_WL, energy_bin=0.001)),
    ("--lj-N 38 --lj-radius 3 --inv-t-wl", dict(N=38, method=_abi.METHOD_INV_T_WL)),
    ("--lj-N 31 --lj-radius 2.5 --samc-t0 1e5", dict(method=_abi.METHOD_SAMC, samc_t0=1e5)),              # run-lj-clusters.sh:32
    ("--wca-reduced-density 0.8 --wca-N 256 --samc-t0 1e7", dict(system=_abi.SYS_WCA, N=256, reduced_density=0.8)),  # wca/run-wca.py:107
    ("--wca-cell-volume 1000 --wca-N 256 --sad-min-T 1", dict(cell_width=[10.0, 10.0, 10.0])),
    ("--sw-cell-width 6 7 8 --sw-N 10 --sw-well-width 1.5 --sad-min-T 1", dict(system=_abi.SYS_SW, cell_width=[6.0, 7.0, 8.0], sw_well_width=1.5)),
    ("--fake-linear --sad-min-T 0.001 --energy-bin 0.01", dict(system=_abi.SYS_FAKE, fake_function=_abi.FAKE_LINEAR, N=1)),  # fake/run-fake.py:61
    ("--fake-quadratic-dimensions 3 --sad-min-T 0.001", dict(fake_function=_abi.FAKE_QUADRATIC, N=3)),
    ("--fake-pieces-a 0.1 --fake-pieces-b 0.2 --fake-pieces-e1 1.0 --fake-pieces-e2 0.5 --sad-min-T 0.1",
     dict(fake_function=_abi.FAKE_PIECES, fake_a=0.1, fake_b=0.2, fake_e1=1.0, fake_e2=0.5)),
    ("--fake-erfinv-mean-energy 0 --fake-erfinv-N 3 --sad-min-T 0.1", dict(system=_abi.SYS_FAKE_ERFINV, N=3, erfinv_mean_energy=0.0)),
    ("--two-wells-N 12 --two-wells-h2-to-h1 1.1 --two-wells-barrier-over-h1 0.1 --two-wells-r2 0.5 --sad-min-T 0.001 --seed 7",
     dict(system=_abi.SYS_TWO_WELLS, N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, seed=7)),  # two-wells/run-two-wells.py:140-164
    ("--ising-N 16 --T 2.5", dict(method=_abi.METHOD_CANONICAL, canonical_T=2.5)),
    ("--ising-N 16 --sad-min-T 1 --num-walkers 4096 --gpu-device 3 --bin-window-lo -600 --bin-window-hi 600",
     dict(n_walkers=4096, device=3, bin_window_lo=-600.0, bin_window_hi=600.0)),
])
def test_flag_spellings(argv, expect):
    c = dry(argv.split())["config"]
    for k, v in expect.items():
        assert c[k] == v, (k, c[k], v)


@pytest.mark.parametrize("argv,msg", [
    ("--lj-N 31 --lj-radius 2.5", "no method"),
    ("--sad-min-T 1", "no system"),
    ("--lj-N 31 --lj-radius 2.5 --ising-N 4 --sad-min-T 1", "more than one system"),
    ("--lj-N 31 --lj-radius 2.5 --sad-min-T 1 --samc-t0 3", "more than one method"),
    ("--lj-N 31 --lj-radius 2.5 --sad-min-T 1 --translation-scale 0.1 --acceptance-rate 0.5", "more than one move plan"),
    ("--lj-N 31 --sad-min-T 1", "--lj-radius is required"),
    ("--lj-N 3.5 --lj-radius 2 --sad-min-T 1", "integer"),
    ("--lj-N 31 --lj-radius 2.5 --sad-min-T 1 --frobnicate", "unknown flag"),
    ("--lj-N 31 --lj-radius 2.5 --sad-min-T", "needs 1 value"),
    ("--water-N 10 --sad-min-T 1", "no device kernel"),
    ("--wca-N 10 --wca-reduced-density 1 --wca-fcc --sad-min-T 1", "fcc"),
    ("--ising-N 8 --sad-min-T 1 --save-as run.txt", "I don't know how to create file"),     # mc/mod.rs:118
])
def test_usage_errors(argv, msg):
    with pytest.raises(SystemExit) as ei:
        dry(argv.split())
    assert msg in str(ei.value)


def test_help_lists_every_flag():
    lines = []
    histogram.main(["--help"], out=lines.append)
    for f in ("--lj-N", "--sad-min-T", "--Inv-t-WL", "--save-as", "--resume-from", "--num-walkers", "--movie-time"):
        assert f in lines[0]


def test_resume_from_rebuilds_the_configuration_from_the_document(tmp_path):
    from sad_monte_carlo_b200 import checkpoint
    doc = {"system": {"Wca": {"E": 3.0, "error": 0.0, "possible_change": "None",
                              "cell": {"box_diagonal": {"x": 5.0, "y": 5.0, "z": 6.0}, "r_cutoff": 2 ** (1 / 6),
                                       "positions": [{"x": 1.0, "y": 1.0, "z": 1.0}, {"x": 3.0, "y": 3.0, "z": 3.0}]}}},
           "method": {"WL": {"gamma": 0.5, "lowest_hist": 0, "highest_hist": 3, "total_hist": 9, "num_states": 4.0, "hist": [1, 2],
                             "min_energy": 1.0, "inv_t": True, "min_gamma": None}},
           "moves": 1000, "accepted_moves": 500, "min_allowed_energy": 0.0, "max_allowed_energy": 20.0,
           "move_plan": {"AcceptanceRate": 0.4}, "translation_scale": 0.07, "acceptance_rate": 0.5, "rng": {"s0": 1, "s1": 2},
           "save_as": "x.json", "report": {"max_iter": {"TotalMoves": 5000}, "max_independent_samples": None, "quiet": True},
           "movies": {"movie_time": None, "which_frame": 0, "period": "Never"}, "save": {"save_time_seconds": 1800.0}, "manager": {},
           "bins": {"min": 2.5, "width": 1.0, "histogram": [1, 2], "t_found": [0, 1], "lnw": [0.0, 1.0], "energy_total": [3.0, 7.0],
                    "energy_squared_total": [9.0, 25.0], "extra": {}},
           "have_visited_since_maxentropy": [False, True], "round_trips": [1, 1], "max_S": 0.0, "max_S_index": 0}
    for ext in ("json", "yaml", "cbor"):
        p = tmp_path / ("x." + ext)
        checkpoint.write_atomic(str(p), checkpoint.dumps(doc, ext))
        d = dry(["--resume-from", str(p)])
        c = d["config"]
        assert (c["system"], c["N"], c["cell_width"], c["method"]) == (_abi.SYS_WCA, 2, [5.0, 5.0, 6.0], _abi.METHOD_INV_T_WL)
        assert (c["move_plan"], c["move_value"], c["energy_bin"], c["init_mode"]) == (_abi.MOVE_ACCEPTANCE_RATE, 0.4, 1.0, _abi.INIT_EXTERNAL)
        assert (c["min_allowed_energy"], c["max_allowed_energy"]) == (0.0, 20.0)
        assert d["plugins"]["max_iter"] == 5000 and d["plugins"]["save_time"] == 0.5


def test_main_glue_runs_with_a_stand_in_engine(monkeypatch, tmp_path):
    """The part of main() behind --dry-run (engine creation, plugins, the run loop, shard placement) with a stand-in
    engine: no GPU here, the real path is tests/test_gpu_cli.py."""
    from sad_monte_carlo_b200 import engine as engine_mod, plugins
    made = []

    class Engine:
        def __init__(self, cfg):
            self.cfg, self.n_walkers, self.moves = cfg, cfg.n_walkers, 0
            made.append(self)

        def run(self, n):
            self.moves += n

        def num_moves(self):
            return self.moves

        def num_accepted_moves(self):  # summed over the walkers, as the C ABI reports it
            return self.n_walkers * (self.moves // 2)

        def accepted_moves_range(self):
            return self.moves // 2, self.moves // 2

        def verify_energy(self, w=0):
            return True

        def close(self):
            self.closed = True

    saved = []
    monkeypatch.setattr(engine_mod, "WalkerEngine", Engine)
    monkeypatch.setattr(plugins.EngineMC, "checkpoint", lambda self: saved.append((self.save_as, self.engine.moves)))
    monkeypatch.chdir(tmp_path)
    lines = []
    assert histogram.main("--ising-N 16 --sad-min-T 1 --max-iter 1000 --num-walkers 8 --save-as a.json".split(), out=lines.append) == 0
    e = made[-1]
    assert (e.moves, e.cfg.n_walkers, e.cfg.walker_offset, e.closed) == (1000, 8, 0, True)
    assert saved[-1] == ("a.json", 1000) and "1000 moves per walker, 8 walkers" in lines[-1]
    for k, v in (("WORLD_SIZE", "4"), ("RANK", "2"), ("LOCAL_RANK", "2")):
        monkeypatch.setenv(k, v)
    assert histogram.main("--ising-N 16 --sad-min-T 1 --max-iter 50 --num-walkers 8 --save-as a.json --quiet".split(), out=lines.append) == 0
    e = made[-1]
    assert (e.moves, e.cfg.n_walkers, e.cfg.walker_offset, e.cfg.device) == (50, 2, 4, 2) and saved[-1] == ("a.rank2of4.json", 50)
