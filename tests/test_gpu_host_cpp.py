"""The compiled host on a GPU: `sad_monte_carlo_b200/bin/histogram` (host/histogram.cpp) drives the same engine through
the same C ABI as the Python host, so the two must leave the same checkpoints, resume each other's files, and pass
the reference's tests/resume-sad.rs."""
import os
import subprocess

import numpy as np
import pytest

from sad_monte_carlo_b200 import build, checkpoint, histogram

pytestmark = pytest.mark.gpu
BIN = build.build_host()


def cpp(args, cwd):
    r = subprocess.run([BIN] + list(args), capture_output=True, text=True, cwd=str(cwd))
    assert r.returncode == 0, r.stderr
    return r.stdout


def py(args, cwd):
    old = os.getcwd()
    os.chdir(str(cwd))
    try:
        out = []
        assert histogram.main(list(args), out=out.append) == 0
        return out
    finally:
        os.chdir(old)


def same_value(a, b, rtol=0.0):
    if isinstance(a, float) or isinstance(b, float):
        if a is None or b is None:
            return a is b
        return a == b or abs(a - b) <= rtol * max(abs(a), abs(b), 1e-300)
    if isinstance(a, dict) and isinstance(b, dict):
        return list(a) == list(b) and all(same_value(a[k], b[k], rtol) for k in a)
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(same_value(x, y, rtol) for x, y in zip(a, b))
    return a == b


def assert_same_documents(a, b, skip=("save_as",), rtol_for=()):
    assert list(a) == list(b)
    for k in a:
        if k in skip:
            continue
        assert same_value(a[k], b[k], 1e-9 if k in rtol_for else 0.0), k


CASES = {
    "ising_sad": ["--ising-N", "16", "--sad-min-T", "1", "--seed", "5", "--max-iter", "3e4", "--quiet", "--num-walkers", "2"],
    "sw_sad": ["--sw-N=64", "--sw-filling-fraction=0.25", "--sw-well-width=1.3", "--sad-min-T=0.5", "--acceptance-rate=0.5", "--max-iter=20000", "--quiet"],
    "lj13_exact_wl": ["--lj-N", "13", "--lj-radius", "2", "--min-allowed-energy=-44", "--max-allowed-energy=0", "--inv-t-wl", "--energy-bin", "0.05",
                      "--max-iter", "2e4", "--quiet", "--lanes-per-walker", "1", "--movie-time", "10"],
    "wca_samc": ["--wca-N", "20", "--wca-reduced-density", "0.4", "--samc-t0", "1e3", "--max-allowed-energy", "200", "--seed", "9", "--max-iter", "6000",
                 "--quiet"],
    "fake_pieces": ["--fake-pieces-a", "0.1", "--fake-pieces-b", "0.2", "--fake-pieces-e1", "1.0", "--fake-pieces-e2", "0.5", "--sad-min-T", "0.1",
                    "--energy-bin", "0.01", "--max-iter", "2e4", "--quiet"],
    "two_wells": ["--two-wells-N", "12", "--two-wells-h2-to-h1", "1.1", "--two-wells-barrier-over-h1", "0.1", "--two-wells-r2", "0.5", "--sad-min-T", "0.001",
                  "--energy-bin", "1e-3", "--translation-scale", "1e-2", "--max-iter", "2e4", "--quiet"],
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("ext", ["yaml", "cbor"])
def test_compiled_and_python_hosts_leave_the_same_checkpoint(case, ext, tmp_path):
    if ext == "cbor" and case not in ("ising_sad", "lj13_exact_wl"):
        pytest.skip("codecs are compared on the CPU (tests/test_host_cpp.py)")
    cpp(CASES[case] + ["--save-as", "c." + ext], tmp_path)
    py(CASES[case] + ["--save-as", "p." + ext], tmp_path)
    n = 2 if "--num-walkers" in CASES[case] else 1
    for w in range(n):
        a = checkpoint.load(checkpoint.walker_path(str(tmp_path / ("c." + ext)), w, n))
        b = checkpoint.load(checkpoint.walker_path(str(tmp_path / ("p." + ext)), w, n))
        if case == "two_wells":  # derived tables: sequential sums here, vectorised sums there
            sa, sb = a["system"]["TwoWells"].pop("invcdf"), b["system"]["TwoWells"].pop("invcdf")
            assert np.allclose(sa.pop("stencils"), sb.pop("stencils"), rtol=0, atol=1e-9) and same_value(sa, sb, 1e-12)
        assert_same_documents(a, b)
    if case == "lj13_exact_wl":  # movie frames at powers of 10 (plugin.rs:434-463), same names from both hosts
        assert sorted(os.listdir(tmp_path / "c")) == sorted(os.listdir(tmp_path / "p")) == ["%014d.cbor" % 10 ** k for k in range(5)]


COMMON = ["--sw-N=100", "--sw-filling-fraction=0.3", "--sw-well-width=1.3", "--sad-min-T=0.5", "--acceptance-rate=0.5", "--quiet"]


@pytest.mark.parametrize("total,first", [(2, 1), (1000, 999), (1000, 500), (100000, 40000)])
def test_resume_sad_through_the_compiled_host(total, first, tmp_path):
    # tests/resume-sad.rs: only the save_as line of the final yaml may differ
    cpp(COMMON + ["--max-iter=%d" % total, "--save-as=big-guy.yaml"], tmp_path)
    cpp(COMMON + ["--max-iter=%d" % first, "--save-as=small-guy.yaml"], tmp_path)
    out = cpp(COMMON + ["--max-iter=%d" % total, "--save-as=small-guy.yaml"], tmp_path)
    assert "Resuming from file" in out
    s1, s2 = open(tmp_path / "big-guy.yaml").read().splitlines(), open(tmp_path / "small-guy.yaml").read().splitlines()
    assert len(s1) == len(s2)
    diff = [(x, y) for x, y in zip(s1, s2) if x != y]
    assert len(diff) == 1 and diff[0][0].startswith("save_as")


def test_hosts_resume_each_others_files(tmp_path):
    args = ["--ising-N", "16", "--sad-min-T", "1", "--seed", "3", "--quiet"]
    cpp(args + ["--max-iter", "30000", "--save-as", "full.json"], tmp_path)
    # Python starts, the compiled host finishes
    py(args + ["--max-iter", "12000", "--save-as", "a.yaml"], tmp_path)
    cpp(args + ["--max-iter", "30000", "--save-as", "a.yaml"], tmp_path)
    # the compiled host starts, Python finishes
    cpp(args + ["--max-iter", "12000", "--save-as", "b.cbor"], tmp_path)
    py(args + ["--max-iter", "30000", "--save-as", "b.cbor"], tmp_path)
    full = checkpoint.load(str(tmp_path / "full.json"))
    for name in ("a.yaml", "b.cbor"):
        assert_same_documents(full, checkpoint.load(str(tmp_path / name)))
    # --resume-from: the configuration comes from the document alone
    cpp(args + ["--max-iter", "12000", "--save-as", "c.json"], tmp_path)
    doc = checkpoint.load(str(tmp_path / "c.json"))
    doc["report"]["max_iter"] = {"TotalMoves": 30000}
    checkpoint.write_atomic(str(tmp_path / "c.json"), checkpoint.dumps(doc, "json"))
    cpp(["--resume-from", "c.json"], tmp_path)
    assert_same_documents(full, checkpoint.load(str(tmp_path / "c.json")))


def test_engine_errors_become_exit_status_1(tmp_path):
    r = subprocess.run([BIN, "--sw-N", "10", "--sw-cell-width", "1.2", "5", "5", "--sw-well-width", "1.3", "--sad-min-T", "1", "--max-iter", "10"],
                       capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 1 and "not large enough" in r.stderr  # wca.rs:186-191 / optsquare.rs panic -> status, not a crash
