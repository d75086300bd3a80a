"""Host-side scheduling (reference src/mc/plugin.rs) and checkpoint encodings, without a GPU."""
import numpy as np
import pytest

from sad_monte_carlo_b200 import checkpoint, plugins
from sad_monte_carlo_b200.plugins import Action, Movie, PluginManager, Report, Save


class FakeMC:
    def __init__(self):
        self.moves = 0
        self.saved, self.frames, self.verified = [], [], 0

    def run(self, n):
        self.moves += n

    def num_moves(self):
        return self.moves

    def num_accepted_moves(self):
        return self.moves // 2

    def independent_samples(self):
        return self.num_accepted_moves()

    def verify_energy(self):
        self.verified += 1

    def checkpoint(self):
        self.saved.append(self.moves)

    def save_movie_frame(self, moves):
        self.frames.append(moves)


def drive(mc, plugs, limit=10**6):
    m = PluginManager()
    launches = []
    while len(launches) < limit:
        n = m.moves_until_next_action()
        mc.run(n)
        launches.append(n)
        if m.run(mc, plugs) == Action.EXIT:
            break
    return launches


def reference_per_move(plugs_factory, max_moves):
    """The reference's literal loop: PluginManager::run after EVERY move (energy.rs:967-973, plugin.rs:93-144)."""
    mc = FakeMC()
    plugs = plugs_factory()
    period, moves = 1, 0
    while mc.moves < max_moves + 5:
        mc.run(1)
        moves += 1
        if moves >= period:
            moves = 0
            todo = max([p.run(mc) for p in plugs] + [Action.NONE])
            if todo >= Action.LOG:
                mc.verify_energy()
            if todo >= Action.SAVE:
                mc.checkpoint()
                for p in plugs:
                    p.save(mc)
            if todo >= Action.EXIT:
                break
            new_period = 1 << 40
            for p in plugs:
                kind, n = p.run_period()
                if kind == "TotalMoves" and n > mc.num_moves() and n - mc.num_moves() < new_period:
                    new_period = n - mc.num_moves()
                elif kind == "Period" and n < new_period:
                    new_period = n
            period = new_period
    return mc


def test_batched_manager_reproduces_the_per_move_manager():
    factory = lambda: [Report(max_iter=5000), Save(save_time_hours=None), Movie(movie_time=1.5)]  # noqa: E731
    ref = reference_per_move(factory, 5000)
    mc = FakeMC()
    launches = drive(mc, factory())
    assert mc.moves == ref.moves == 5000
    assert mc.saved == ref.saved and mc.frames == ref.frames
    assert sum(launches) == 5000 and len(launches) < 60  # a few dozen launches instead of 5000 plugin calls


def test_save_doubles_without_a_time_budget_and_report_exits_at_max_iter():
    mc = FakeMC()
    drive(mc, [Report(max_iter=1000), Save(save_time_hours=None)])
    # Save without save_time: next_output doubles (plugin.rs:380-382); the final Exit also saves (Action ordering)
    assert mc.saved == [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1000]
    assert mc.moves == 1000


def test_movie_frames_at_powers_of_movie_time():
    mc = FakeMC()
    drive(mc, [Report(max_iter=300), Movie(movie_time=2.0)])
    assert mc.frames == [1, 2, 4, 8, 16, 32, 64, 128, 256]
    mv = Movie(movie_time=1.3335)
    mv.shall_i_save(1)
    assert mv.period == plugins.total_moves(2) and mv.which_frame == 2  # 1.3335^1 rounds to 1 again: skipped, frames never repeat


def test_save_schedule_follows_the_clock():
    t = [0.0]
    s = Save(save_time_hours=1.0 / 3600.0, clock=lambda: t[0])  # one second between saves
    t[0] = 0.5
    assert s.shall_i_save(1)           # 0.5 s per move -> 1 + 2 moves per period, not < moves -> 1/time_per_move = 2
    assert s.next_output == 2
    t[0] = 0.0
    s2 = Save(save_time_hours=1.0 / 3600.0, clock=lambda: t[0])
    t[0] = 1e-3
    s2.shall_i_save(1)                 # 1 ms per move: moves + 1 < 1000 -> next after one second's worth of moves
    assert s2.next_output == 1000
    t[0] = 10.0
    assert s2.shall_i_save(10000)      # 1 ms per move, 1001 moves per period < moves
    assert s2.next_output == 10000 + 1001


def test_independent_samples_stop():
    mc = FakeMC()
    drive(mc, [Report(max_iter=None, max_independent_samples=100), Save(save_time_hours=None)])
    assert mc.num_accepted_moves() >= 100 and mc.moves <= 512


@pytest.mark.parametrize("ext", ["yaml", "json", "cbor"])
def test_document_round_trips_through_every_format(ext, tmp_path):
    doc = {"system": {"Lj": {"E": -101.25, "error": 1.5e-12, "possible_change": "None",
                             "positions": [{"x": 0.1, "y": -2.0, "z": 3e-7}], "max_radius_squared": 6.25, "max_radius": 2.5}},
           "method": {"Sad": {"min_T": 0.01, "too_lo": -131.2, "too_hi": -70.4, "tL": 123, "tF": 120, "num_states": 5000,
                              "highest_hist": 77, "version": "Sad", "latest_parameter": 6080.0}},
           "moves": 2**40 + 7, "rng": {"s0": 2**64 - 1, "s1": 0x6e789e6aa1b965f4}, "min_allowed_energy": None,
           "bins": {"min": -133.605, "width": 0.01, "histogram": [0, 1, 2**33], "lnw": [0.0, -1.5, 1e-300], "extra": {}},
           "have_visited_since_maxentropy": [True, False, True], "save_as": "x." + ext, "manager": {}}
    p = tmp_path / ("ck." + ext)
    checkpoint.write_atomic(str(p), checkpoint.dumps(doc, ext))
    assert checkpoint.load(str(p)) == doc
    assert [f.name for f in tmp_path.iterdir()] == ["ck." + ext]  # no temporary file left behind


def test_unknown_extension_is_an_error_like_the_reference_panic():
    with pytest.raises(ValueError):
        checkpoint.dumps({}, "toml")
    with pytest.raises(ValueError):
        checkpoint.loads(b"", "dat")


def test_cbor_is_standard_cbor():
    # RFC 8949 appendix A examples
    enc = lambda x: checkpoint.dumps(x, "cbor")  # noqa: E731
    assert enc(1000) == bytes.fromhex("1903e8")
    assert enc(-1000) == bytes.fromhex("3903e7")
    assert enc(1.1) == bytes.fromhex("fb3ff199999999999a")
    assert enc([1, [2, 3]]) == bytes.fromhex("8201820203")
    assert enc({"a": 1, "b": [2, 3]}) == bytes.fromhex("a26161016162820203")
    assert enc(18446744073709551615) == bytes.fromhex("1bffffffffffffffff")
    assert checkpoint.loads(bytes.fromhex("f97c00"), "cbor") == np.inf  # half-precision floats decode too


# ---- resumed runs: plugin.rs:315-320 (`next_output`, `start` are serde(skip)), 373-376, 262-264 ---------------------

class FakeClock:
    def __init__(self):
        self.t = 0.0

    def __call__(self):
        return self.t


def test_resumed_save_measures_time_per_move_over_this_process_only():
    """A run resumed at 5e9 moves that makes 1e6 moves per second must checkpoint about every save_time, not after
    another 5e9 moves: the reference restarts its clock at the resume point (start = None -> (now, moves)) and saves
    again 2^20 moves after the first post-resume save."""
    clock = FakeClock()
    resumed_at = 5_000_000_000
    rate = 1e6  # moves per second of this process
    mc = FakeMC()
    mc.moves = resumed_at
    save = Save(save_time_hours=0.5, clock=clock, resumed=True)
    report = Report(max_iter=resumed_at + 20_000_000_000, quiet=False, out=lambda s: lines.append(s), resumed=True, clock=clock)
    lines = []
    m = PluginManager()
    for _ in range(40):
        n = m.moves_until_next_action()
        mc.run(n)
        clock.t += n / rate
        m.run(mc, [report, save])
        if len(mc.saved) >= 6:
            break
    assert mc.saved[0] == resumed_at + 1            # first tick of a resumed run saves (next_output = 0)
    assert mc.saved[1] == resumed_at + 1 + (1 << 20)  # plugin.rs:375
    gaps = np.diff(mc.saved[2:])
    assert len(gaps) >= 2 and np.all(np.abs(gaps / rate - 1800.0) < 0.05 * 1800.0), gaps  # then every half hour
    # the report's time per move is this process's (1 us), not (run time) / (all moves since move 0)
    progress = [ln for ln in lines if "per move" in ln]
    assert progress and all("1 us per move" in ln for ln in progress), lines


def test_fresh_save_schedule_is_unchanged():
    clock = FakeClock()
    mc = FakeMC()
    save = Save(save_time_hours=None, clock=clock)
    drive(mc, [save, Report(max_iter=100)])
    assert mc.saved[:7] == [1, 2, 4, 8, 16, 32, 64]


def test_incomplete_or_partial_checkpoint_sets_are_refused_before_any_engine_exists(tmp_path):
    base = str(tmp_path / "run.json")
    for w in (0, 1, 3):  # walker 2 of 4 is missing
        open(checkpoint.walker_path(base, w, 4), "w").write("{}")
    with pytest.raises(ValueError) as ei:
        checkpoint.check_resumable(None, base, 4)
    assert "incomplete" in str(ei.value) and "1 of 4" in str(ei.value) and "run-w000002.json" in str(ei.value)
    open(str(tmp_path / "run.partial"), "w").write("2 of 4 walkers\n")
    with pytest.raises(ValueError) as ei:
        checkpoint.check_resumable(None, base, 4)
    assert "--checkpoint-walkers" in str(ei.value) and "2 of 4 walkers" in str(ei.value)


def test_checkpoint_set_is_written_all_or_nothing(tmp_path):
    """A halted walker stops the save before any file is touched; a failure while documents are being staged leaves the
    previous set in place and no temporaries behind."""
    class Eng:
        n_walkers = 3
        halted = (0, 0)
        fail_at = None

        def num_halted(self):
            return self.halted

        def walker(self, w):
            class S:
                status = 0
            return S()

    eng = Eng()
    base = str(tmp_path / "set.json")
    real = checkpoint.walker_document
    try:
        def fake_document(engine, w, save_as="x", **kw):
            if engine.fail_at == w:
                raise RuntimeError("device read failed")
            return {"moves": engine.moves, "walker": w}
        checkpoint.walker_document = fake_document
        eng.moves = 100
        checkpoint.save(eng, base)
        first = [open(checkpoint.walker_path(base, w, 3)).read() for w in range(3)]
        eng.moves, eng.fail_at = 200, 2
        with pytest.raises(RuntimeError):
            checkpoint.save(eng, base)
        assert [open(checkpoint.walker_path(base, w, 3)).read() for w in range(3)] == first  # still the complete old set
        assert sorted(p.name for p in tmp_path.iterdir()) == ["set-w000000.json", "set-w000001.json", "set-w000002.json"]
        eng.fail_at, eng.halted = None, (1, 0)
        with pytest.raises(RuntimeError) as ei:
            checkpoint.save(eng, base)
        assert "no checkpoint written" in str(ei.value)
        assert [open(checkpoint.walker_path(base, w, 3)).read() for w in range(3)] == first
        eng.halted = (0, 0)
        checkpoint.save(eng, base, walkers=range(2))
        assert (tmp_path / "set.partial").read_text() == "2 of 3 walkers\n"
        checkpoint.save(eng, base)
        assert not (tmp_path / "set.partial").exists()
    finally:
        checkpoint.walker_document = real
