"""The `histogram` command line on the device engine: the reference's tests/resume-sad.rs run through it (a run to
`total` moves, and a run to `first` moves continued by the same command with a larger --max-iter, must leave the same
checkpoint), plus --resume-from and many walkers per command."""
import os

import numpy as np
import pytest

from sad_monte_carlo_b200 import checkpoint, histogram

pytestmark = pytest.mark.gpu

COMMON = ["--sw-N=100", "--sw-filling-fraction=0.3", "--sw-well-width=1.3", "--sad-min-T=0.5", "--acceptance-rate=0.5", "--quiet"]


def run(args, cwd):
    old = os.getcwd()
    os.chdir(cwd)
    try:
        lines = []
        assert histogram.main(list(args), out=lines.append) == 0
        return lines
    finally:
        os.chdir(old)


def differing_lines(a, b):
    la, lb = a.splitlines(), b.splitlines()
    assert len(la) == len(lb)
    return [(x, y) for x, y in zip(la, lb) if x != y]


@pytest.mark.parametrize("total,first", [(2, 1), (5, 4), (1000, 1), (1000, 999), (1000, 500), (200000, 100000)])
def test_resume_sad_through_the_command_line(total, first, tmp_path):
    run(COMMON + ["--max-iter=%d" % total, "--save-as=big-guy.yaml"], tmp_path)
    run(COMMON + ["--max-iter=%d" % first, "--save-as=small-guy.yaml"], tmp_path)
    assert checkpoint.load(str(tmp_path / "small-guy.yaml"))["moves"] == first
    out = run(COMMON + ["--max-iter=%d" % total, "--save-as=small-guy.yaml"], tmp_path)
    assert any("Resuming from file" in l for l in out)
    s1 = open(tmp_path / "big-guy.yaml").read()
    s2 = open(tmp_path / "small-guy.yaml").read()
    diff = differing_lines(s1, s2)
    assert len(diff) == 1 and "save_as" in diff[0][0]  # resume-sad.rs:84: only save_as differs


def test_resume_from_continues_the_checkpoint_as_it_is(tmp_path):
    args = ["--ising-N", "16", "--sad-min-T", "1", "--seed", "5", "--quiet", "--save-time", "1/3600"]
    run(args + ["--max-iter", "3e4", "--save-as", "a.json"], tmp_path)
    run(args + ["--max-iter", "1e4", "--save-as", "b.json"], tmp_path)
    # raise max_iter inside the checkpoint, as a user editing the yaml would; --resume-from reads nothing else
    doc = checkpoint.load(str(tmp_path / "b.json"))
    doc["report"]["max_iter"] = {"TotalMoves": 30000}
    checkpoint.write_atomic(str(tmp_path / "b.json"), checkpoint.dumps(doc, "json"))
    run(["--resume-from", "b.json"], tmp_path)
    a, b = checkpoint.load(str(tmp_path / "a.json")), checkpoint.load(str(tmp_path / "b.json"))
    assert a["moves"] == b["moves"] == 30000
    for k in ("bins", "method", "rng", "system", "accepted_moves", "round_trips", "have_visited_since_maxentropy", "max_S"):
        assert a[k] == b[k], k


def test_many_walkers_one_file_each_and_walker_w_is_seed_plus_w(tmp_path):
    base = ["--lj-N", "13", "--lj-radius", "2", "--max-allowed-energy=0", "--sad-min-T", "0.05", "--energy-bin", "0.05",
            "--translation-scale", "0.05", "--max-iter", "2e4", "--quiet", "--lanes-per-walker", "1"]
    run(base + ["--seed", "3", "--num-walkers", "4", "--checkpoint-walkers", "3", "--save-as", "many.cbor"], tmp_path)
    # three of four walkers written: the set is marked as partial (and a later --save-as on it refuses to resume)
    assert sorted(os.listdir(tmp_path)) == ["many-w%06d.cbor" % w for w in range(3)] + ["many.partial"]
    assert (tmp_path / "many.partial").read_text() == "3 of 4 walkers\n"
    with pytest.raises(SystemExit) as ei:  # histogram.UsageError
        run(base + ["--seed", "3", "--num-walkers", "4", "--save-as", "many.cbor"], tmp_path)
    assert "cannot be resumed" in str(ei.value)
    run(base + ["--seed", "5", "--save-as", "one.cbor"], tmp_path)
    many, one = checkpoint.load(str(tmp_path / "many-w000002.cbor")), checkpoint.load(str(tmp_path / "one.cbor"))
    for k in ("bins", "method", "rng", "system", "accepted_moves"):
        assert many[k] == one[k], k
    assert np.sum(one["bins"]["histogram"]) == 20001
