"""CPU checks of the `binning` and `tempering` command lines (flag handling and --dry-run: no device needed) on the
command lines of the reference's job scripts."""
import json

import pytest

from sad_monte_carlo_b200 import _abi, binning, replicas, tempering


def _dry(mod, argv):
    lines = []
    assert mod.main(argv + ["--dry-run"], out=lines.append) == 0
    return json.loads(lines[0])


def test_binning_job_script_line_of_run_fake():
    # fake/run-fake.py:25-36: binning --save-time 0.5 --histogram-bin de --translation-scale 0.05 --movie-time 10^(1/4) --fake-linear ...
    d = _dry(binning, "--save-time 0.5 --histogram-bin 0.01 --translation-scale 0.05 --movie-time 10^(1/4) --fake-linear "
                      "--save-as sad-linear-0.01.yaml --max-iter 1e11 --sad-min-T 0.001".split())
    assert d["binning"] == {"Histogram": {"bin": 0.01}}
    assert d["plugins"]["max_iter"] == 10 ** 11 and abs(d["plugins"]["movie_time"] - 10 ** 0.25) < 1e-15
    assert d["save_as"] == "sad-linear-0.01.yaml"


def test_binning_wl_line_with_bounds_and_default_bin():
    d = _dry(binning, "--fake-quadratic-dimensions 3 --wl --min-allowed-energy 0 --max-allowed-energy 0.99 --translation-scale 0.05".split())
    assert d["binning"] == {"Histogram": {"bin": 1.0}}  # BinningParams::default (binning.rs:63-69)


def test_binning_wca_line():
    # wca-transposed/run.py:52
    d = _dry(binning, "--wca-reduced-density 0.8 --wca-N 32 --sad-min-T 0.5 --max-allowed-energy 640 --translation-scale 0.005 "
                      "--histogram-bin 1 --lanes-per-walker 32".split())
    assert d["binning"]["Histogram"]["bin"] == 1.0


@pytest.mark.parametrize("argv,msg", [
    ("--fake-linear --linear-bin 0.01 --histogram-bin 0.01 --sad-min-T 0.001", "more than one binning"),
    ("--fake-linear --energy-bin 0.01 --sad-min-T 0.001", "--histogram-bin"),
    ("--fake-linear --histogram-bin 0.01 --T 0.5", "no canonical method"),
])
def test_binning_refuses_what_is_not_built(argv, msg):
    with pytest.raises(SystemExit) as ei:
        binning.main(argv.split() + ["--dry-run"], out=lambda s: None)
    assert msg in str(ei.value)


def test_binning_high_resolution_flag():
    cfg = binning.config_from_flags({"fake-linear": True, "histogram-bin": 0.01, "high-resolution-de": 0.001, "sad-min-T": 0.1})
    assert cfg.high_resolution_de == 0.001
    cfg = binning.config_from_flags({"fake-linear": True, "histogram-bin": 0.01, "sad-min-T": 0.1})
    assert cfg.high_resolution_de != cfg.high_resolution_de  # None


def test_binning_linear_flag():
    cfg = binning.config_from_flags({"fake-linear": True, "linear-bin": 0.02, "sad-min-T": 0.1})
    assert cfg.flags & _abi.FLAG_BINNING_LINEAR and cfg.flags & _abi.FLAG_BINNING and cfg.energy_bin == 0.02


def test_binning_config_carries_the_flag():
    cfg = binning.config_from_flags({"fake-linear": True, "histogram-bin": 0.125, "sad-min-T": 0.1})
    assert cfg.flags & _abi.FLAG_BINNING and cfg.energy_bin == 0.125


def test_tempering_job_script_line_of_run_two_wells():
    # two-wells/run-two-wells.py:45-61 with geometric_spacing(0.001, 1, 20) and --canonical-steps 10
    T = tempering.geometric_spacing(0.001, 1.0, 20)
    argv = ("--two-wells-N 12 --two-wells-h2-to-h1 1.1 --two-wells-barrier-over-h1 0.1 --two-wells-r2 0.5 --movie-time 10^(1/8) "
            "--save-time 0.5 --save-as tem+x.cbor --max-iter 1e12 --canonical-steps 10 --seed 3").split()
    for t in T:
        argv += ["--T", str(t)]
    d = _dry(tempering, argv)
    assert d["T"] == T and d["canonical_steps"] == 10 and d["save_as"] == "tem+x.cbor"
    assert abs(T[-1] - 1.0) < 1e-12 and abs(T[1] / T[0] - T[2] / T[1]) < 1e-12


def test_tempering_default_ladder_and_refused_flags():
    d = _dry(tempering, "--fake-linear".split())
    assert d["T"][0] == 0.001 and d["T"][-1] == 1.024 and len(d["T"]) == 11 and d["canonical_steps"] == 1  # tempering.rs:31-35
    with pytest.raises(SystemExit):
        tempering.main("--fake-linear --sad-min-T 0.1 --dry-run".split(), out=lambda s: None)


def test_replicas_job_script_line_of_run_fake():
    # fake/run-fake.py:16-23: replicas <system> --movie-time 10^(1/4) --save-time 0.5 --save-as r-linear.yaml --max-iter 1e11 --min-T 0.001
    d = _dry(replicas, "--fake-linear --movie-time 10^(1/4) --save-time 0.5 --save-as r-linear.yaml --max-iter 1e11 --min-T 0.001".split())
    assert d["min_T"] == 0.001 and d["independent_systems_before_new_bin"] == 64 and d["save_as"] == "r-linear.yaml"
    d = _dry(replicas, "--fake-erfinv-mean-energy 0 --fake-erfinv-N 3 --min-T 0.1 --independent-systems-before-new-bin 16".split())
    assert d["min_T"] == 0.1 and d["independent_systems_before_new_bin"] == 16
    with pytest.raises(SystemExit):
        replicas.main("--fake-linear --sad-min-T 0.1 --dry-run".split(), out=lambda s: None)
