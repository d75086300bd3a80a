"""Report / Save / Movie scheduling around the device engine: the reference's `PluginManager`.

In the reference `move_once` calls `PluginManager::run` after EVERY move (src/mc/energy.rs:967-973); the manager only
counts until `period` moves have passed, then asks every plugin what to do and recomputes the period as the minimum
over the plugins' `run_period()` (src/mc/plugin.rs:93-144).  Nothing observable happens in between, so the device
engine runs exactly `period` moves per launch and the plugins see the state they would have seen:

    manager = PluginManager()
    while True:
        engine.run(manager.moves_until_next_action())      # one kernel launch
        if manager.run(mc, [report, save, movies]) == Action.EXIT: break

The plugins restate plugin.rs: `Report` 149-310 (stop at max_iter / max_independent_samples, progress line),
`Save` 313-400 (checkpoint schedule: doubling, or wall-clock based), `Movie` 402-477 (frames at powers of
movie_time).  `mc` is anything with num_moves(), num_accepted_moves(), independent_samples(), checkpoint(),
verify_energy() and save_movie_frame(moves) -- `EngineMC` adapts a WalkerEngine.
"""
import enum
import os
import time


class Action(enum.IntEnum):  # plugin.rs:54-64, ordered: max() combines
    NONE = 0
    LOG = 1
    SAVE = 2
    EXIT = 3


NEVER = ("Never", None)  # TimeToRun, plugin.rs:43-51


def total_moves(n):
    return ("TotalMoves", int(n))


def period(n):
    return ("Period", int(n))


class Plugin:
    def run(self, mc):
        return Action.NONE

    def run_period(self):
        return NEVER

    def log(self, mc):
        pass

    def save(self, mc):
        pass


class Report(Plugin):
    """plugin.rs:149-310"""

    def __init__(self, max_iter=None, max_independent_samples=None, quiet=True, out=print, resumed=False, clock=time.monotonic):
        self.max_iter = total_moves(max_iter) if max_iter is not None else NEVER
        self.max_independent_samples = max_independent_samples
        self.quiet = quiet
        self.clock = clock
        # `start` is #[serde(skip, default)] (plugin.rs:153-155): a deserialised Report has None and takes (now, moves)
        # at its first log instead of printing (262-264), so the time per move is measured over THIS process only
        self.start = None if resumed else (clock(), 0)
        self.out = out

    def document(self):
        mi = "Never" if self.max_iter == NEVER else {"TotalMoves": self.max_iter[1]}
        return {"max_iter": mi, "max_independent_samples": self.max_independent_samples, "quiet": self.quiet}

    def am_all_done(self, moves, independent_samples):  # plugin.rs:271-283
        if self.max_iter[0] == "TotalMoves" and moves >= self.max_iter[1]:
            return True
        if self.max_independent_samples is not None:
            return independent_samples >= self.max_independent_samples
        return False

    def run(self, mc):
        return Action.EXIT if self.am_all_done(mc.num_moves(), mc.independent_samples()) else Action.NONE

    def run_period(self):
        return self.max_iter

    def log(self, mc):  # Report::print, 206-269 (wording kept, durations in seconds)
        if self.quiet:
            return
        moves = mc.num_moves()
        if self.start is None:  # plugin.rs:262-264
            self.start = (self.clock(), moves)
            return
        t0, it0 = self.start
        runtime = self.clock() - t0
        per_move = runtime / max(1, moves - it0)
        if self.max_iter[0] == "TotalMoves":
            mx = self.max_iter[1]
            left = max(0, mx - moves)
            self.out("[%.3g] %d%% complete after %.0f s (%.0f s left, %.3g us per move)" % (
                moves, int(100.0 * moves / mx), runtime, per_move * left, per_move * 1e6))
        else:
            self.out("[%.3g] after %.0f s (%.3g us per move)" % (moves, runtime, per_move * 1e6))

    def save(self, mc):  # 295-309; with many walkers: the mean per walker (what one reference process would print)
        if self.quiet:
            return
        acc, moves = mc.num_accepted_moves(), mc.num_moves()
        self.out("        Accepted %.3g/%.3g = %.0f%% of the moves" % (acc, moves, 100.0 * acc / max(1, moves)))


class Save(Plugin):
    """plugin.rs:313-400.  save_time_seconds None: checkpoints at moves 1, 2, 4, 8, ...; otherwise the schedule adapts
    to the measured time per move so that a checkpoint happens about every save_time_seconds."""

    def __init__(self, save_time_hours=1.0, clock=time.monotonic, resumed=False):
        self.clock = clock
        # next_output and start are #[serde(skip, default)] (plugin.rs:315-320): a resumed run saves at its first tick
        # (next_output = 0), takes (now, moves) as its start there and saves next 2^20 moves later (373-376) -- it
        # never divides this process's run time by the move count of the earlier processes
        self.next_output = 0 if resumed else 1
        self.start = None if resumed else (clock(), 0)
        self.save_time_seconds = None if save_time_hours is None else 3600.0 * save_time_hours

    def document(self):
        return {"save_time_seconds": self.save_time_seconds}

    def shall_i_save(self, moves):  # 353-383
        if moves < self.next_output:
            return False
        if self.save_time_seconds is not None:
            if self.start is None:  # plugin.rs:373-376
                self.start = (self.clock(), moves)
                self.next_output = moves + (1 << 20)
                return True
            t0, it0 = self.start
            per_move = (self.clock() - t0) / max(1, moves - it0)
            per_move = max(per_move, 1e-30)
            moves_per_period = 1 + int(self.save_time_seconds / per_move)
            if moves_per_period < moves:
                self.next_output = moves + moves_per_period
            elif moves + 1.0 < 1.0 / per_move:
                self.next_output = int(1.0 / per_move)
            else:
                self.next_output = moves * 2
        else:
            self.next_output *= 2
        return True

    def run(self, mc):
        return Action.SAVE if mc.num_moves() >= self.next_output else Action.NONE

    def run_period(self):
        return total_moves(self.next_output)

    def save(self, mc):
        self.shall_i_save(mc.num_moves())


class Movie(Plugin):
    """plugin.rs:402-477: frame k is due at move round(movie_time ** k)."""

    def __init__(self, movie_time=None):
        self.movie_time = movie_time
        self.which_frame = 0
        self.period = total_moves(1) if movie_time is not None else NEVER

    def document(self):
        p = "Never" if self.period == NEVER else {"TotalMoves": self.period[1]}
        return {"movie_time": self.movie_time, "which_frame": self.which_frame, "period": p}

    def restore(self, doc):
        """`movies` of a checkpoint: all three fields are serialised (plugin.rs:403-408), a resumed run keeps its schedule."""
        self.movie_time = doc.get("movie_time")
        self.which_frame = int(doc.get("which_frame", 0))
        p = doc.get("period", "Never")
        self.period = NEVER if p == "Never" else total_moves(p["TotalMoves"])

    def shall_i_save(self, moves):  # 446-463
        if self.movie_time is not None and self.period == total_moves(moves):
            which = self.which_frame + 1
            nxt = int(self.movie_time ** which + 0.5)
            while nxt <= moves:
                which += 1
                nxt = int(self.movie_time ** which + 0.5)
            self.which_frame = which
            self.period = total_moves(nxt)
            return True
        return False

    def run(self, mc):
        if self.shall_i_save(mc.num_moves()):
            mc.save_movie_frame(mc.num_moves())
            return Action.SAVE
        return Action.NONE

    def run_period(self):
        return self.period


class PluginManager:
    """plugin.rs:74-144.  `period` and `moves` are not serialised (74-80): a resumed run ticks after its first move."""

    def __init__(self):
        self.period = 1
        self.moves = 0

    def moves_until_next_action(self):
        return max(1, self.period - self.moves)

    def run(self, mc, plugins, moves_made=None):
        """Call after the engine advanced by `moves_made` moves (default: moves_until_next_action())."""
        self.moves += self.moves_until_next_action() if moves_made is None else moves_made
        if self.moves < self.period:
            return Action.NONE
        self.moves = 0
        todo = Action.NONE
        for p in plugins:
            todo = max(todo, p.run(mc))
        if todo >= Action.LOG:
            mc.verify_energy()
            for p in plugins:
                p.log(mc)
        if todo >= Action.SAVE:
            mc.checkpoint()
            for p in plugins:
                p.save(mc)
        if todo >= Action.EXIT:
            return Action.EXIT
        new_period = 1 << 40  # run plugins every trillion iterations minimum
        now = mc.num_moves()
        for p in plugins:
            kind, n = p.run_period()
            if kind == "TotalMoves":
                if n > now and n - now < new_period:
                    new_period = n - now
            elif kind == "Period":
                if n < new_period:
                    new_period = n
        self.period = new_period
        return todo


class EngineMC:
    """Adapts a WalkerEngine to what the plugins call on `MonteCarlo` (src/mc/mod.rs:37-143)."""

    def __init__(self, engine, save_as="resume.yaml", checkpoint_walkers=None, report=None, save=None, movies=None):
        self.engine = engine
        self.save_as = str(save_as)
        self.checkpoint_walkers = checkpoint_walkers
        self.report, self.save_plugin, self.movies = report, save, movies

    def num_moves(self):
        return self.engine.num_moves()

    # Every walker is one reference run (`--seed seed + w`), so the per-run quantities the plugins ask for are
    # per-walker quantities: the progress line shows the mean over the walkers, and --max-independent-samples is
    # reached when EVERY walker has that many (the slowest walker decides; all walkers make the same number of moves).
    def num_accepted_moves(self):
        return self.engine.num_accepted_moves() // max(1, self.engine.n_walkers)

    def independent_samples(self):  # mc/mod.rs:134-136
        return self.engine.accepted_moves_range()[0]

    def verify_energy(self):  # PluginManager::run calls sys.verify_energy() before logging (plugin.rs:102-103)
        if not self.engine.verify_energy(0):
            raise RuntimeError("verify_energy failed for walker 0")

    def _docs(self):
        d = {}
        if self.report is not None:
            d["report"] = self.report.document()
        if self.save_plugin is not None:
            d["save"] = self.save_plugin.document()
        if self.movies is not None:
            d["movies"] = self.movies.document()
        return d

    def checkpoint(self):
        from . import checkpoint
        return checkpoint.save(self.engine, self.save_as, walkers=self.checkpoint_walkers, **self._docs())

    def save_movie_frame(self, moves):  # Movie::save_frame, plugin.rs:434-444: <save_as stem>/<moves:014>.cbor
        from . import checkpoint
        d = os.path.splitext(self.save_as)[0]
        return checkpoint.save(self.engine, os.path.join(d, "%014d.cbor" % moves), walkers=self.checkpoint_walkers, **self._docs())


def run_simulation(engine, report, save=None, movies=None, save_as="resume.yaml", checkpoint_walkers=None, max_launch=None):
    """`loop { mc.move_once() }` of src/bin/histogram.rs for a device engine: returns the number of launches."""
    plugins = [p for p in (report, save, movies) if p is not None]
    mc = EngineMC(engine, save_as, checkpoint_walkers, report, save, movies)
    manager = PluginManager()
    launches = 0
    while True:
        n = manager.moves_until_next_action()
        if max_launch is not None:
            n = min(n, max_launch)
        engine.run(n)
        launches += 1
        if manager.run(mc, plugins, moves_made=n) == Action.EXIT:
            return launches
