"""tools/check_reference_dump.py (docs/PIN_WITH_REFERENCE.md): the tool that compares a dump from the reference's own
binary with the oracle must itself be known to work -- run it on a dump written by the oracle, on a corrupted copy,
and on checkpoint pairs."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import check_reference_dump as crd  # noqa: E402


def test_own_dump_round_trips_and_a_corrupted_one_is_caught(tmp_path):
    p = tmp_path / "dump.jsonl"
    crd.write_own_dump(str(p))
    lines = open(p).read().splitlines()
    assert len(lines) == 4 * 21
    assert crd.check_dump(str(p), out=lambda s: None) == 0
    # 20 000 normals per seed leave the ziggurat's fast path a few hundred times: the dump exercises wedge and tail
    d = json.loads(lines[18])
    assert d["kind"] == "normal_bits"
    x = np.array(d["values"], np.uint64).view(np.float64)
    assert abs(x.mean()) < 0.03 and abs(x.std() - 1.0) < 0.02 and np.abs(x).max() > 3.6541528853610088
    d["values"][777] ^= 1
    lines[18] = json.dumps(d)
    q = tmp_path / "bad.jsonl"
    q.write_text("\n".join(lines) + "\n")
    msgs = []
    assert crd.check_dump(str(q), out=msgs.append) == 1
    assert "first difference at draw 777" in msgs[0]


def test_checkpoint_comparison(tmp_path):
    from sad_monte_carlo_b200 import checkpoint
    doc = {"system": {"Ising": {"E": -4.0, "N": 2, "S": [1, -1, 1, -1], "possible_change": None}},
           "method": {"Samc": {"t0": 10.0}}, "moves": 5, "accepted_moves": 3, "rng": {"s0": 1, "s1": 2},
           "bins": {"min": -6.0, "width": 4.0, "histogram": [2, 4], "t_found": [0, 1], "lnw": [0.5, 1.25],
                    "energy_total": [-8.0, 0.0], "energy_squared_total": [32.0, 0.0], "extra": {}},
           "round_trips": [1, 1], "have_visited_since_maxentropy": [False, True], "max_S": 0.0, "max_S_index": 0,
           "translation_scale": 0.05, "acceptance_rate": 0.5, "save_as": "a.json"}
    a, b = tmp_path / "a.json", tmp_path / "b.yaml"
    checkpoint.write_atomic(str(a), checkpoint.dumps(doc, "json"))
    other = json.loads(json.dumps(doc))
    other["save_as"] = "b.yaml"                               # may differ (tests/resume-sad.rs:84)
    other["system"]["Ising"]["possible_change"] = [0, -4.0]   # pending change: not trajectory state
    checkpoint.write_atomic(str(b), checkpoint.dumps(other, "yaml"))
    assert crd.check_checkpoints(str(a), str(b), out=lambda s: None) == 0
    other["bins"]["lnw"][1] = 1.25 * (1 + 1e-13)
    checkpoint.write_atomic(str(b), checkpoint.dumps(other, "yaml"))
    assert crd.check_checkpoints(str(a), str(b), out=lambda s: None) == 1
    assert crd.check_checkpoints(str(a), str(b), rtol=1e-12, out=lambda s: None) == 0
    other["rng"]["s1"] = 3
    checkpoint.write_atomic(str(b), checkpoint.dumps(other, "yaml"))
    msgs = []
    assert crd.check_checkpoints(str(a), str(b), rtol=1e-12, out=msgs.append) == 1 and msgs[0] == "DIFFERENT: rng"
