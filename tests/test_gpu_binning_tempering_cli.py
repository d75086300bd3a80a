"""The `binning` and `tempering` command lines end to end on the GPU: a short run of each, and the checkpoints they
write read back in the reference's serde schema (what plotting/parse-binning.py / parse-tempering.py index)."""
import os

import numpy as np
import pytest

from sad_monte_carlo_b200 import _abi, binning, checkpoint, make_config, replicas, tempering
from tests.oracle_lib import OracleBinningMC, OracleReplicas, OracleTempering

pytestmark = pytest.mark.gpu


def _run(mod, args, cwd):
    old = os.getcwd()
    os.chdir(cwd)
    try:
        lines = []
        assert mod.main(list(args), out=lines.append) == 0
        return lines
    finally:
        os.chdir(old)


@pytest.mark.parametrize("ext", ["yaml", "cbor"])
def test_binning_run_writes_the_reference_schema_and_matches_the_oracle(tmp_path, ext):
    args = ("--fake-quadratic-dimensions 3 --histogram-bin 0.01 --translation-scale 0.05 --sad-min-T 0.001 --seed 4 --max-iter 30000 "
            "--quiet --save-as sad." + ext).split()
    _run(binning, args, tmp_path)
    doc = checkpoint.load(str(tmp_path / ("sad." + ext)))
    assert doc["moves"] == 30000 and set(doc["method"]) == {"Sad"} and doc["high_resolution"] is None
    h = doc["bins"]["Histogram"]
    assert set(h) == {"min", "min_e", "max_e", "width", "lnw", "extra"} and h["width"] == 0.01
    assert set(h["lnw"]) == {"total", "min_total", "max_total", "e_max_total", "count", "min_count", "max_count", "e_max_count", "total_count"}
    assert {"energy", "t_found"} <= set(h["extra"]) and sum(h["extra"]["energy"]["count"]) == 30000
    o = OracleBinningMC(make_config("fake", "sad", fake_function=_abi.FAKE_QUADRATIC, N=3, sad_min_T=0.001, energy_bin=0.01, move_value=0.05, seed=4))
    o.run(30000)
    s, b = o.walker(), o.bins()
    assert (doc["rng"]["s0"], doc["rng"]["s1"]) == (s.rng_s0, s.rng_s1) and doc["accepted_moves"] == s.accepted_moves
    assert doc["method"]["Sad"]["too_lo"] == s.too_lo and doc["method"]["Sad"]["tF"] == s.tF and doc["method"]["Sad"]["num_states"] == s.num_states
    assert np.array_equal(np.array(h["lnw"]["total"]), b["lnw_total"]) and h["lnw"]["count"] == [int(x) for x in b["lnw_count"]]
    assert h["lnw"]["max_count"] == s.lnw_max_count and h["extra"]["t_found"]["max_total"] == s.t_found_max_total


@pytest.mark.parametrize("sysargs,method", [
    ("--fake-quadratic-dimensions 3 --histogram-bin 0.01 --translation-scale 0.05", "--sad-min-T 0.001"),
    ("--ising-N 16 --histogram-bin 4", "--wl --min-allowed-energy -400 --max-allowed-energy 0"),
    ("--fake-linear --histogram-bin 0.01 --high-resolution-de 0.0007 --translation-scale 0.05", "--sad-min-T 0.001"),
    ("--two-wells-N 12 --two-wells-h2-to-h1 1.1 --two-wells-barrier-over-h1 0.1 --two-wells-r2 0.5 --histogram-bin 0.001 --translation-scale 0.01", "--samc-t0 1e4"),
])
def test_binning_resume_continues_bit_for_bit(tmp_path, sysargs, method):
    """tests/resume-sad.rs for the `binning` binary: 2e4 moves + resume to 5e4 == 5e4 moves in one go (only save_as differs)."""
    base = (sysargs + " " + method + " --seed 7 --num-walkers 3 --quiet").split()
    _run(binning, base + ["--max-iter", "2e4", "--save-as", "a.yaml"], tmp_path)
    lines = _run(binning, base + ["--max-iter", "5e4", "--save-as", "a.yaml"], tmp_path)
    assert any("Resuming" in x for x in lines)
    _run(binning, base + ["--max-iter", "5e4", "--save-as", "b.yaml"], tmp_path)
    for w in range(3):
        a = open(tmp_path / ("a-w%06d.yaml" % w)).read().splitlines()
        b = open(tmp_path / ("b-w%06d.yaml" % w)).read().splitlines()
        diff = [(x, y) for x, y in zip(a, b) if x != y]
        assert len(a) == len(b) and all(x.startswith("save_as") for x, _ in diff), diff[:3]


def test_binning_linear_run_writes_the_linear_variant(tmp_path):
    args = "--fake-quadratic-dimensions 3 --linear-bin 0.02 --translation-scale 0.05 --sad-min-T 0.001 --seed 4 --max-iter 20000 --quiet --save-as lin.json".split()
    _run(binning, args, tmp_path)
    doc = checkpoint.load(str(tmp_path / "lin.json"))
    h = doc["bins"]["Linear"]
    assert h["width"] == 0.02 and abs(sum(h["extra"]["energy"]["count"]) - 20000) < 1e-6 and isinstance(h["lnw"]["count"][0], float)
    o = OracleBinningMC(make_config("fake", "sad", fake_function=_abi.FAKE_QUADRATIC, N=3, sad_min_T=0.001, energy_bin=0.02, move_value=0.05, seed=4,
                                    flags=_abi.FLAG_BINNING | _abi.FLAG_BINNING_LINEAR))
    o.run(20000)
    b = o.bins_f64()
    assert np.array_equal(np.array(h["lnw"]["total"]), b["lnw_total"]) and np.array_equal(np.array(h["lnw"]["count"]), b["lnw_count"])
    assert h["lnw"]["max_count"] == o.walker().lnw_max_count_f64


def test_tempering_run_writes_one_document_per_simulation(tmp_path):
    T = tempering.geometric_spacing(0.01, 1.0, 6)
    args = ("--two-wells-N 12 --two-wells-h2-to-h1 1.1 --two-wells-barrier-over-h1 0.1 --two-wells-r2 0.5 --canonical-steps 10 --seed 2 "
            "--num-walkers 3 --max-iter 200000 --quiet --movie-time 10 --save-as tem.cbor").split()
    for t in T:
        args += ["--T", repr(t)]
    _run(tempering, args, tmp_path)
    files = sorted(os.listdir(tmp_path))
    assert [f for f in files if f.endswith(".cbor")] == ["tem-w%06d.cbor" % k for k in range(3)] and "tem" in files
    doc = checkpoint.load(str(tmp_path / "tem-w000001.cbor"))
    assert set(doc) == {"T", "rng", "save_as", "moves", "replicas", "canonical_steps", "save", "movie", "report"}
    per_round = 12 * 10 * 6
    assert doc["moves"] % per_round == 0 and 200000 <= doc["moves"] < 200000 + per_round and doc["T"] == T
    o = OracleTempering(make_config("two-wells", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, seed=2), T, 10, sim=1)
    o.run_once(doc["moves"] // per_round)
    assert (doc["rng"]["s0"], doc["rng"]["s1"]) == o.rng()
    for r, q in zip(doc["replicas"], o.replicas()):
        assert set(r) == {"T", "rejected_count", "accepted_count", "rejected_swap_count", "accepted_swap_count", "ignored_count", "system",
                          "rng", "total_energy", "total_energy_squared", "translation_scale"}
        assert (r["accepted_count"], r["rejected_count"], r["accepted_swap_count"], r["total_energy"], r["rng"]["s0"]) == (
            q.accepted_count, q.rejected_count, q.accepted_swap_count, q.total_energy, q.rng_s0)
        assert set(r["system"]) == {"TwoWells"}
    frames = sorted(os.listdir(tmp_path / "tem"))
    assert frames and all(f.endswith(".cbor") for f in frames)  # movie frames at 10^k, labelled with the tick's move number


def test_tempering_resume_continues_bit_for_bit(tmp_path):
    T = tempering.geometric_spacing(0.01, 1.0, 5)
    base = ("--two-wells-N 12 --two-wells-h2-to-h1 1.1 --two-wells-barrier-over-h1 0.1 --two-wells-r2 0.5 --canonical-steps 4 --seed 5 "
            "--num-walkers 2 --quiet").split()
    for t in T:
        base += ["--T", repr(t)]
    _run(tempering, base + ["--max-iter", "48000", "--save-as", "a.json"], tmp_path)
    lines = _run(tempering, base + ["--max-iter", "120000", "--save-as", "a.json"], tmp_path)
    assert any("Resuming" in x for x in lines)
    _run(tempering, base + ["--max-iter", "120000", "--save-as", "b.json"], tmp_path)
    for k in range(2):
        a = checkpoint.load(str(tmp_path / ("a-w%06d.json" % k)))
        b = checkpoint.load(str(tmp_path / ("b-w%06d.json" % k)))
        a.pop("save_as"), b.pop("save_as")
        assert a == b


def test_replicas_run_writes_one_document_per_simulation(tmp_path):
    args = "--fake-quadratic-dimensions 3 --min-T 0.001 --independent-systems-before-new-bin 16 --seed 2 --num-walkers 2 --max-iter 300000 --quiet --save-as r.yaml".split()
    _run(replicas, args, tmp_path)
    doc = checkpoint.load(str(tmp_path / "r-w000001.yaml"))
    assert set(doc) == {"min_T", "rng", "save_as", "moves", "independent_systems_before_new_bin", "median", "replicas", "save", "movie", "report"}
    o = OracleReplicas(make_config("fake", fake_function=_abi.FAKE_QUADRATIC, N=3, seed=2), 0.001, 16, sim=1)
    while o.moves() < doc["moves"]:
        o.run_once(1)
    assert o.moves() == doc["moves"] and (doc["rng"]["s0"], doc["rng"]["s1"]) == o.rng() and len(doc["replicas"]) == o.num_replicas() > 3
    assert doc["median"]["energies"] == list(o.median())
    for r, q in zip(doc["replicas"], o.replicas()):
        assert (r["cutoff_energy"], r["above_count"], r["below_total"], r["unique_visitors"], r["translation_scale"]) == (
            q.cutoff_energy, q.above_count, q.below_total, q.unique_visitors, q.translation_scale)
    assert doc["replicas"][0]["max_energy"] == float("inf")
