"""The compiled host (host/*.hpp, host/histogram.cpp -> sad_monte_carlo_b200/bin/histogram): the same command line,
the same documents and the same three codecs as the Python host, checked against it on the CPU (`--dry-run` and
`--convert` need no GPU)."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

from sad_monte_carlo_b200 import build, checkpoint, histogram

BIN = build.build_host()

COMMAND_LINES = [
    "--lj-N 31 --max-allowed-energy=0 --sad-min-T 0.01 --translation-scale 0.05 --energy-bin 0.01 --save-as lj-sad-31-bin001.yaml "
    "--movie-time 10^(1/8) --save-time 0.5 --lj-radius 2.5 --seed=3",                                    # run-lj-clusters.sh:53
    "--sw-N=100 --sw-filling-fraction=0.3 --sw-well-width=1.3 --sad-min-T=0.5 --acceptance-rate=0.5 --max-iter=1000 --save-as=big-guy.yaml",
    "--ising-N 32 --wl --wl-min-gamma=1e-4 --min-allowed-energy=-2048 --max-allowed-energy=50",
    "--lj-N 31 --lj-radius 2.5 --Inv-t-WL --min-allowed-energy=-133.53 --max-allowed-energy=-110 --energy-bin 0.001",
    "--lj-N 38 --lj-radius 3 --inv-t-wl --max-iter 1e9 --quiet",
    "--lj-N 31 --lj-radius 2.5 --samc-t0 1e5 --max-independent-samples 12",
    "--wca-reduced-density 0.8 --wca-N 256 --samc-t0 1e7 --max-allowed-energy 2560",
    "--wca-cell-volume 1000 --wca-N 256 --sad-min-T 1",
    "--wca-cell-width=5,6,7 --wca-N 20 --sad-min-T 1",
    "--sw-cell-width 6 7 8 --sw-N 10 --sw-well-width 1.5 --sad-min-T 1",
    "--fake-linear --sad-min-T 0.001 --energy-bin 0.01",
    "--fake-quadratic-dimensions 3 --sad-min-T 0.001",
    "--fake-pieces-a 0.1 --fake-pieces-b 0.2 --fake-pieces-e1 1.0 --fake-pieces-e2 0.5 --sad-min-T 0.1",
    "--fake-gaussian-sigma 0.3 --sad-min-T 0.1",
    "--fake-erfinv-mean-energy 0 --fake-erfinv-N 3 --sad-min-T 0.1",
    "--two-wells-N 12 --two-wells-h2-to-h1 1.1 --two-wells-barrier-over-h1 0.1 --two-wells-r2 1/2 --sad-min-T 0.001 --seed 7",
    "--ising-N 16 --T 2.5",
    "--ising-N 16 --sad-min-T sqrt(2)*pi --num-walkers 4096 --gpu-device 3 --bin-window-lo -600 --bin-window-hi 600 --fast-math --lanes-per-walker 1",
    "--lj-N 31 --lj-radius 2.5 --max-allowed-energy=0 --sad-min-T 0.01 --energy-bin 0.01 --translation-scale 0.05 --num-walkers 56832 --fast-math --lj-stream-z",
    "--lj-N 38 --lj-radius 3 --max-allowed-energy=0 --sad-min-T 0.01 --energy-bin 0.01 --num-walkers 300 --fast-math --lj-smem-z --lanes-per-walker 1",
]


def run(argv, check=True, cwd=None):
    r = subprocess.run([BIN] + argv, capture_output=True, text=True, cwd=cwd)
    if check:
        assert r.returncode == 0, r.stderr
    return r


def same(a, b):
    if isinstance(a, float) and isinstance(b, float):
        return a == b or (math.isnan(a) and math.isnan(b))
    if isinstance(a, dict) and isinstance(b, dict):
        return list(a) == list(b) and all(same(a[k], b[k]) for k in a)
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    return type(a) is type(b) and a == b


@pytest.mark.parametrize("line", COMMAND_LINES)
def test_dry_run_is_what_the_python_host_parses(line):
    argv = line.split()
    got = json.loads(run(argv + ["--dry-run"]).stdout)
    out = []
    histogram.main(argv + ["--dry-run"], out=out.append)
    want = json.loads(out[-1])
    assert same(got, want), (got, want)


@pytest.mark.parametrize("argv,msg", [
    ("--lj-N 31 --lj-radius 2.5", "no method"), ("--sad-min-T 1", "no system"),
    ("--lj-N 31 --lj-radius 2.5 --ising-N 4 --sad-min-T 1", "more than one system"),
    ("--lj-N 31 --sad-min-T 1", "--lj-radius is required"), ("--lj-N 3.5 --lj-radius 2 --sad-min-T 1", "integer"),
    ("--lj-N 31 --lj-radius 2.5 --sad-min-T 1 --frobnicate", "unknown flag"), ("--lj-N 31 --lj-radius 2.5 --sad-min-T", "needs 1 value"),
    ("--lj-N 31 --lj-radius 1/0 --sad-min-T 1", "division by zero"), ("--water-N 10 --sad-min-T 1", "no device kernel"),
    ("--ising-N 8 --sad-min-T 1 --save-as run.txt", "I don't know how to create file"),
])
def test_usage_errors_exit_2(argv, msg):
    r = run(argv.split() + ["--dry-run"], check=False)
    assert r.returncode == 2 and msg in r.stderr


def test_engine_errors_are_messages_not_crashes(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run(["--ising-N", "8", "--sad-min-T", "1", "--max-iter", "10", "--save-as", str(tmp_path / "x.json")], check=False)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and not (tmp_path / "x.json").exists()


DOC = {
    "system": {"Wca": {"E": 3.25, "error": 1e-300, "possible_change": "None",
                       "cell": {"box_diagonal": {"x": 5.0, "y": 5.5, "z": 6.0}, "r_cutoff": 2 ** (1 / 6),
                                "positions": [{"x": 0.1, "y": -1e-7, "z": 1.0}, {"x": 3.0000000000000004, "y": 3.0, "z": 1e21}]}}},
    "method": {"WL": {"gamma": 0.5, "lowest_hist": 0, "highest_hist": 3, "total_hist": 9, "num_states": 4.0, "hist": [1, 2, 0],
                      "min_energy": -1.0, "inv_t": True, "min_gamma": None}},
    "moves": 12345678901234, "rng": {"s0": 18446744073709551615, "s1": 9223372036854775808}, "which_frame": -3,
    "move_plan": {"AcceptanceRate": 0.4}, "save_as": "dir with space/x: y.yaml", "odd strings": ["", "null", "1.5", "true", "a, b", "- x", "#c", "é"],
    "report": {"max_iter": "Never", "max_independent_samples": None, "quiet": False}, "manager": {}, "empty": [],
    "bins": {"lnw": [0.0, -0.0, 1.0 / 3.0, 1e100, float("inf")] + [0.1 * k for k in range(60)], "histogram": list(range(70))},
    "nested": [[1, 2], [3, [4, {"k": [5.5]}]], {"a": {"b": {"c": [1, 2, 3]}}}],
}


@pytest.mark.parametrize("src", ["yaml", "json", "cbor"])
@pytest.mark.parametrize("dst", ["yaml", "json", "cbor"])
def test_codecs_agree_with_the_python_host(src, dst, tmp_path):
    a, b = tmp_path / ("in." + src), tmp_path / ("out." + dst)
    checkpoint.write_atomic(str(a), checkpoint.dumps(DOC, src))       # written by Python (PyYAML flow style, wrapped lines)
    run(["--convert", str(a), "--convert-to", str(b)])                  # read and re-written by the compiled host
    got = checkpoint.load(str(b))                                      # read by Python
    assert same(got, DOC), (src, dst)


def test_block_style_yaml_as_serde_writes_it(tmp_path):
    # serde_yaml emits block sequences and never flow collections
    text = """---
system:
  Ising:
    E: -4.0
    N: 2
    S:
      - 1
      - -1
      - 1
      - -1
    possible_change: ~
method:
  Sad:
    min_T: 1.0
    too_lo: -8.0
moves: 7
bins:
  extra: {}
  lnw:
    - 0.0
    - 1.5e-3
positions:
  - x: 1.0
    y: 2.0
  - x: 3.0
    y: 4.0
save_as: "a b.yaml"
"""
    a, b = tmp_path / "in.yaml", tmp_path / "out.json"
    a.write_text(text)
    run(["--convert", str(a), "--convert-to", str(b)])
    import yaml
    assert same(json.loads(b.read_text()), yaml.safe_load(text))


def test_two_wells_invcdf_tables_match_the_python_host():
    # derived data of the TwoWells document (two_wells.rs:46-137); the compiled host sums sequentially like the reference
    src = os.path.join(os.path.dirname(BIN), "..", "..", "host", "checkpoint.hpp")
    assert "two_wells_invcdf" in open(src).read()
    st = np.array(checkpoint.two_wells_invcdf(12, 0.5)["stencils"]).reshape(12, 10000)
    assert (st[:, 0] == 0).all() and np.allclose(st[:, -1], 1.0)


def test_codecs_fuzz_against_the_python_host(tmp_path):
    """Random documents (nested maps / lists, awkward strings and numbers) written by Python in one format, re-encoded by
    the compiled host into another, read back by Python."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    import re
    # the one class of plain scalars the yaml generations disagree on: `1e5` is a float for serde_yaml (1.2 core schema,
    # what the reference writes) and a string for PyYAML (1.1); the compiled host reads it as the reference means it
    numberish = re.compile(r"[-+]?(\d[\d_]*\.?[\d_]*|\.\d+)([eE][-+]?\d+)?")
    ok = lambda s: s == s.strip() and not numberish.fullmatch(s)
    keys = st.text(alphabet="abcXYZ_ -:#'\"09", min_size=1, max_size=8).filter(ok)
    scalars = st.one_of(st.none(), st.booleans(), st.integers(min_value=-2 ** 63, max_value=2 ** 64 - 1),
                        st.floats(allow_nan=False), st.text(alphabet="ab -:,#[]{}'\"\\~!&*|>%@`09.eE+\t", max_size=10).filter(ok),
                        st.sampled_from(["null", "~", "true", "No", "0x1f", ".inf", "-", "? x", "1_000", "a: b", "- a", "a #b", "x" * 100 + " y" * 40]))
    docs = st.recursive(scalars, lambda c: st.one_of(st.lists(c, max_size=5), st.dictionaries(keys, c, max_size=5)), max_leaves=30)
    n = [0]

    @settings(max_examples=300, derandomize=True, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.dictionaries(keys, docs, min_size=1, max_size=6), st.sampled_from(["yaml", "json", "cbor"]), st.sampled_from(["yaml", "json", "cbor"]))
    def check(doc, src, dst):
        n[0] += 1
        a, b = tmp_path / ("f%d.%s" % (n[0], src)), tmp_path / ("g%d.%s" % (n[0], dst))
        checkpoint.write_atomic(str(a), checkpoint.dumps(doc, src))
        run(["--convert", str(a), "--convert-to", str(b)])
        got = checkpoint.load(str(b))
        assert same(got, doc), (src, dst, doc, got)

    check()


def _doc(system, method, **top):
    d = {"system": system, "method": method, "moves": 1000, "time_L": 0, "accepted_moves": 500, "min_allowed_energy": None,
         "max_allowed_energy": None, "move_plan": {"TranslationScale": 0.05}, "translation_scale": 0.05, "acceptance_rate": 0.5,
         "rng": {"s0": 1, "s1": 2}, "save_as": "x.json", "report": {"max_iter": {"TotalMoves": 5000}, "max_independent_samples": 77, "quiet": False},
         "movies": {"movie_time": 2.0, "which_frame": 9, "period": {"TotalMoves": 1024}}, "save": {"save_time_seconds": 900.0}, "manager": {},
         "bins": {"min": -2.5, "width": 0.5, "histogram": [1, 2, 3], "t_found": [0, 1, 2], "lnw": [0.0, 1.0, 2.0], "energy_total": [1.0, 2.0, 3.0],
                  "energy_squared_total": [1.0, 4.0, 9.0], "extra": {}},
         "have_visited_since_maxentropy": [False, True, True], "round_trips": [1, 1, 1], "max_S": 0.0, "max_S_index": 0}
    d.update(top)
    return d


P3 = [{"x": 0.1, "y": 0.2, "z": 0.3}, {"x": 1.0, "y": 1.1, "z": 1.2}, {"x": -1.0, "y": 0.0, "z": 0.5}]
SAD = {"Sad": {"min_T": 0.01, "too_lo": -2.0, "too_hi": -1.0, "tL": 10, "tF": 20, "num_states": 3, "highest_hist": 3, "version": "Sad",
               "latest_parameter": 100.0}}
WLM = {"WL": {"gamma": 0.25, "lowest_hist": 1, "highest_hist": 3, "total_hist": 6, "num_states": 3.0, "hist": [1, 2, 3], "min_energy": -2.0,
              "inv_t": False, "min_gamma": 1e-4}}
RESUME_DOCS = {
    "lj": _doc({"Lj": {"E": -2.0, "error": 0.0, "possible_change": "None", "positions": P3, "max_radius_squared": 6.25, "max_radius": 2.5}}, SAD,
               max_allowed_energy=0.0),
    "lj_open": _doc({"Lj": {"E": -2.0, "error": 0.0, "possible_change": "None", "positions": P3, "max_radius_squared": 6.25, "max_radius": 2.5}},
                    {"Samc": {"t0": 1e5}}),
    "ising": _doc({"Ising": {"E": -4.0, "N": 2, "S": [1, -1, 1, -1], "possible_change": None}}, WLM, min_allowed_energy=-8.0, max_allowed_energy=2.0,
                  move_plan={"AcceptanceRate": 0.3}),
    "fake_linear": _doc({"Fake": {"position": [0.5], "function": "Linear", "possible_change": [0.0]}}, SAD),
    "fake_quadratic": _doc({"Fake": {"position": [0.1, 0.2, 0.3, 0.4], "function": {"Quadratic": {"dimensions": 4}}, "possible_change": [0.0] * 4}}, SAD),
    "fake_pieces": _doc({"Fake": {"position": [0.1, 0.2, 0.3], "function": {"Pieces": {"a": 0.1, "b": 0.2, "e1": 1.0, "e2": 0.5}},
                                  "possible_change": [0.0] * 3}}, {"Canonical": {"temperature": 1.5}}),
    "fake_gaussian": _doc({"Fake": {"position": [0.1, 0.2, 0.3], "function": {"Gaussian": {"sigma": 0.3}}, "possible_change": [0.0] * 3}}, SAD),
    "erfinv": _doc({"FakeErfinv": {"position": [0.5, 0.5, 0.5], "parameters": {"mean_energy": 0.25}, "possible_change": []}}, {"Samc": {"t0": 10.0}}),
    "wca": _doc({"Wca": {"E": 3.0, "error": 1e-9, "possible_change": "None",
                         "cell": {"box_diagonal": {"x": 5.0, "y": 5.5, "z": 6.0}, "r_cutoff": 2 ** (1 / 6), "positions": P3}}},
                {"WL": dict(WLM["WL"], inv_t=True, min_gamma=None)}),
    "sw": _doc({"Sw": {"E": -3.0, "possible_change": "None",
                       "cell": {"box_diagonal": {"x": 5.0, "y": 5.0, "z": 5.0}, "r_cutoff": 1.3, "positions": P3}}}, SAD),
    "two_wells": _doc({"TwoWells": {"position": [-0.99] + [0.0] * 5, "d_squared": 0.9801, "parameters": {"N": 6, "h2_to_h1": 1.1, "barrier_over_h1": 0.1,
                                                                                                    "r2": 0.5},
                                    "change": {"index": 0, "values": {"x": 0.0, "y": 0.0, "z": 0.0}}, "well_position": 0.66, "invcdf": {}}}, SAD),
}


@pytest.mark.parametrize("name", sorted(RESUME_DOCS))
@pytest.mark.parametrize("ext", ["json", "yaml", "cbor"])
def test_resume_from_configuration_is_what_the_python_host_derives(name, ext, tmp_path):
    # --resume-from reads nothing but the file (mc/mod.rs:92-106): system, method, bounds, move plan, bin width and the
    # plugin parameters all come out of the document -- the same way in both hosts
    p = tmp_path / ("ck." + ext)
    checkpoint.write_atomic(str(p), checkpoint.dumps(RESUME_DOCS[name], ext))
    argv = ["--resume-from", str(p), "--num-walkers", "1", "--dry-run"]
    got = json.loads(run(argv).stdout)
    out = []
    histogram.main(argv, out=out.append)
    want = json.loads(out[-1])
    assert same(got, want), (got, want)
    assert got["config"]["init_mode"] == 2 and got["plugins"]["max_iter"] == 5000 and got["plugins"]["save_time"] == 0.25


@pytest.fixture(scope="module")
def selftest_exe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("selftest") / "selftest_plugins")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(root, "host", "selftest_plugins.cpp")])
    return exe


def _python_schedule(max_iter, movie_time, doubling, accept_every, max_samples=None):
    """The same scripted Monte Carlo through the Python host's plugins (held against a literal per-move restatement of
    plugin.rs in tests/test_host_plugins.py)."""
    from sad_monte_carlo_b200 import plugins
    lines = []

    class MC:
        moves = 0

        def num_moves(self):
            return self.moves

        def num_accepted_moves(self):
            return self.moves // accept_every

        def independent_samples(self):
            return self.num_accepted_moves()

        def verify_energy(self):
            lines.append("verify %d" % self.moves)

        def checkpoint(self):
            lines.append("checkpoint %d" % self.moves)

        def save_movie_frame(self, m):
            lines.append("frame %d" % m)

    mc = MC()
    report = plugins.Report(max_iter=max_iter or None, max_independent_samples=max_samples, quiet=True)
    save = plugins.Save(save_time_hours=None if doubling else 1e-12 / 3600.0)
    movies = plugins.Movie(movie_time=movie_time)
    manager = plugins.PluginManager()
    for _ in range(100000):
        n = manager.moves_until_next_action()
        mc.moves += n
        a = manager.run(mc, [report, save, movies], moves_made=n)
        lines.append("tick %d action %d period %d frame %d" % (mc.moves, int(a), manager.period, movies.which_frame))
        if a == plugins.Action.EXIT:
            return lines
    raise AssertionError("the scripted run never ended")


@pytest.mark.parametrize("max_iter,movie_time,doubling,accept_every,max_samples", [
    (1000, 2.0, 1, 3, None), (12345, 10 ** 0.125, 1, 2, None), (5000, None, 1, 5, None), (3000, 1.5, 0, 2, None),
    (0, 3.0, 1, 4, 777), (10 ** 7, 10.0, 1, 3, None), (1, 2.0, 1, 1, None),
])
def test_plugin_schedule_equals_the_python_host(max_iter, movie_time, doubling, accept_every, max_samples, selftest_exe):
    exe = selftest_exe
    argv = [exe, str(max_iter), "none" if movie_time is None else repr(movie_time), str(doubling), str(accept_every)]
    if max_samples is not None:
        argv.append(str(max_samples))
    got = subprocess.run(argv, capture_output=True, text=True, check=True).stdout.splitlines()
    want = _python_schedule(max_iter, movie_time, doubling, accept_every, max_samples)
    if not doubling:
        # with a save_time the next checkpoint depends on the measured time per move: only the stop and the frames are comparable
        keep = lambda ls: [l for l in ls if l.startswith("frame")] + ls[-1:]
        got, want = [l.split(" period")[0] for l in keep(got)], [l.split(" period")[0] for l in keep(want)]
    assert got == want


def test_resumed_run_restarts_its_clock_in_both_hosts(selftest_exe):
    """plugin.rs:315-320, 373-376: a resumed Save has no start and next_output = 0 -- it saves at its first tick, 2^20 moves
    later, and from then on about every save_time of THIS process's run time (not of run time / all moves since move 0)."""
    from sad_monte_carlo_b200 import plugins
    resumed_at, rate, save_time = 5_000_000_000, 1e6, 1800.0
    got = subprocess.run([selftest_exe, "resumed", str(resumed_at), repr(rate), repr(save_time), "6"], capture_output=True, text=True, check=True)
    cpp = [int(l.split()[1]) for l in got.stdout.splitlines()]

    class Clock:
        t = 0.0

        def __call__(self):
            return self.t

    class MC:
        moves = resumed_at
        saved = []

        def num_moves(self):
            return self.moves

        def num_accepted_moves(self):
            return self.moves // 2

        def independent_samples(self):
            return self.moves // 2

        def verify_energy(self):
            pass

        def checkpoint(self):
            self.saved.append(self.moves)

    clock, mc = Clock(), MC()
    plugs = [plugins.Report(quiet=True, resumed=True, clock=clock), plugins.Save(save_time_hours=save_time / 3600.0, clock=clock, resumed=True)]
    m = plugins.PluginManager()
    while len(mc.saved) < 6:
        n = m.moves_until_next_action()
        mc.moves += n
        clock.t += n / rate
        m.run(mc, plugs, moves_made=n)
    assert cpp == mc.saved
    assert cpp[0] == resumed_at + 1 and cpp[1] == resumed_at + 1 + (1 << 20)
    assert all(abs((b - a) / rate - save_time) < 0.05 * save_time for a, b in zip(cpp[2:], cpp[3:]))


@pytest.mark.parametrize("rank,world,local", [(0, 2, 0), (3, 4, 3), (9, 16, 1)])
def test_one_process_per_gpu_sharding_is_the_same_in_both_hosts(rank, world, local, monkeypatch):
    # torchrun / mpirun export WORLD_SIZE, RANK, LOCAL_RANK: --num-walkers is the total, every rank gets a contiguous block
    # of global walkers (walker w == the reference run with --seed seed+w, whatever the GPU count), its own device and files
    argv = "--lj-N 31 --lj-radius 2.5 --sad-min-T 0.01 --max-allowed-energy 0 --num-walkers 8192 --seed 5 --save-as out/run.cbor --dry-run".split()
    env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(local))
    r = subprocess.run([BIN] + argv, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    for k, v in (("WORLD_SIZE", world), ("RANK", rank), ("LOCAL_RANK", local)):
        monkeypatch.setenv(k, str(v))
    out = []
    histogram.main(argv, out=out.append)
    want = json.loads(out[-1])
    assert same(got, want)
    c = got["config"]
    assert (c["n_walkers"], c["walker_offset"], c["device"], c["seed"]) == (8192 // world, rank * (8192 // world), local, 5)
    assert got["save_as"] == "out/run.rank%dof%d.cbor" % (rank, world)
    # an explicit device wins; a walker count that does not divide is an error
    r = subprocess.run([BIN] + argv + ["--gpu-device", "0"], capture_output=True, text=True, env=env)
    assert json.loads(r.stdout)["config"]["device"] == 0
    r = subprocess.run([BIN] + [a if a != "8192" else "8191" for a in argv], capture_output=True, text=True, env=env)
    assert r.returncode == 2 and "does not divide" in r.stderr
    with pytest.raises(SystemExit):
        histogram.main([a if a != "8192" else "8191" for a in argv], out=out.append)
