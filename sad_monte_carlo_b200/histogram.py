"""`histogram`: the reference's main binary (src/bin/histogram.rs = `EnergyMC::<Any>::from_args::<AnyParams>()` +
`loop { mc.move_once() }`) over the device engine -- SURVEY.md section 8 f2.

    python -m sad_monte_carlo_b200.histogram --lj-N 31 --lj-radius 2.5 --max-allowed-energy=0 --sad-min-T 0.01 \\
        --translation-scale 0.05 --energy-bin 0.01 --save-as lj-sad-31-bin001.yaml --movie-time '10^(1/8)' \\
        --save-time 0.5 --seed=3 --max-iter 1e9 --num-walkers 4096

Flags are the ones `auto_args` derives from the reference's parameter structs (struct fields -> `--kebab-case`, enum
variants -> prefixes, `_fields` flattened; mc/mod.rs:22-32, mc/energy.rs:41-97, mc/plugin.rs:159-167,322-326,411-415,
system/any.rs:10-27 and the per-system parameter structs), with `--flag value` or `--flag=value` and numeric values
given as expressions (`'10^(1/8)'`, `1e9`, `1/3`).  Behaviour follows `MonteCarlo::from_args` (mc/mod.rs:55-107):
`--save-as` on an existing file resumes it (report and save parameters refreshed from the flags,
energy.rs:899-902), `--resume-from FILE` continues a checkpoint as it is, the run ends when `Report` says so
(`--max-iter`, `--max-independent-samples`) after a final checkpoint (plugin.rs:104-124).

What is added for the GPU: `--num-walkers W` independent walkers (walker w is the reference run with `--seed
seed+w`; with W > 1 every walker gets its own file `name-w000017.yaml`), `--gpu-device`, `--bin-window-lo/-hi`
(the device keeps a fixed bin window per walker), `--lanes-per-walker`, `--lj-stream-z` / `--lj-smem-z` (force one LJ31 / LJ38 kernel layout; same results), `--fast-math` (LJ: tolerance tier, <= 1e-12
relative per move), `--checkpoint-walkers K` (write only the first K walkers: such a set is marked `name.partial` and
cannot be resumed), `--dry-run` (print the parsed configuration as JSON and stop: needs no GPU).  `--num-threads` is
accepted and ignored, as `EnergyMC` ignores rayon.  With W > 1 the per-run quantities of the reference are per-walker
quantities: `--max-independent-samples S` ends the run when EVERY walker has accepted S moves, and the progress line
reports the mean number of accepted moves per walker.  A walker that leaves the bin window ends the run with an error.
"""
import json
import math
import os
import sys

import numpy as np

from . import _abi


# ---- expression-valued numbers ("meval" in auto_args) --------------------------------------------------------------

class ExprError(ValueError):
    pass


_FUNCS = {"sqrt": math.sqrt, "abs": abs, "exp": math.exp, "ln": math.log, "log": math.log10, "sin": math.sin, "cos": math.cos,
          "tan": math.tan, "floor": math.floor, "ceil": math.ceil, "round": round}
_CONSTS = {"pi": math.pi, "e": math.e}


def evaluate(text):
    """Arithmetic on f64: + - * / % ^ (right-associative power), parentheses, unary minus, sqrt/exp/ln/..., pi, e."""
    s = text.strip()
    pos = 0

    def peek():
        nonlocal pos
        while pos < len(s) and s[pos].isspace():
            pos += 1
        return s[pos] if pos < len(s) else ""

    def number():
        nonlocal pos
        start = pos
        while pos < len(s) and (s[pos].isdigit() or s[pos] == "."):
            pos += 1
        if pos < len(s) and s[pos] in "eE" and pos > start:
            q = pos + 1
            if q < len(s) and s[q] in "+-":
                q += 1
            if q < len(s) and s[q].isdigit():
                pos = q
                while pos < len(s) and s[pos].isdigit():
                    pos += 1
        try:
            return float(s[start:pos])
        except ValueError:
            raise ExprError("bad number in %r" % text)

    def atom():
        nonlocal pos
        c = peek()
        if c == "(":
            pos += 1
            v = expr()
            if peek() != ")":
                raise ExprError("missing ) in %r" % text)
            pos += 1
            return v
        if c.isdigit() or c == ".":
            return number()
        if c.isalpha():
            start = pos
            while pos < len(s) and (s[pos].isalnum() or s[pos] == "_"):
                pos += 1
            name = s[start:pos]
            if peek() == "(":
                if name not in _FUNCS:
                    raise ExprError("unknown function %r in %r" % (name, text))
                pos += 1
                v = expr()
                if peek() != ")":
                    raise ExprError("missing ) in %r" % text)
                pos += 1
                return float(_FUNCS[name](v))
            if name in ("inf", "infinity"):
                return math.inf
            if name not in _CONSTS:
                raise ExprError("unknown name %r in %r" % (name, text))
            return _CONSTS[name]
        raise ExprError("cannot parse %r" % text)

    def power():
        nonlocal pos
        base = atom()
        if peek() == "^":
            pos += 1
            return base ** unary()
        return base

    def unary():
        nonlocal pos
        c = peek()
        if c == "-":
            pos += 1
            return -unary()
        if c == "+":
            pos += 1
            return unary()
        return power()

    def term():
        nonlocal pos
        v = unary()
        while peek() in ("*", "/", "%") and peek() != "":
            op = s[pos]
            pos += 1
            r = unary()
            v = v * r if op == "*" else (v / r if op == "/" else math.fmod(v, r))
        return v

    def expr():
        nonlocal pos
        v = term()
        while peek() in ("+", "-") and peek() != "":
            op = s[pos]
            pos += 1
            r = term()
            v = v + r if op == "+" else v - r
        return v

    if not s:
        raise ExprError("empty number")
    try:
        v = expr()
    except ZeroDivisionError:
        raise ExprError("division by zero in %r" % text)
    if peek() != "":
        raise ExprError("trailing characters in %r" % text)
    return float(v)


def _as_int(name, v):
    if v != v or v < 0 or v != math.floor(v) or v >= 2.0 ** 64:
        raise ExprError("--%s needs a non-negative integer, got %r" % (name, v))
    return int(v)


# ---- the flag table ---------------------------------------------------------------------------------------------

F64, INT, FLAG, PATH, VEC3 = "f64", "int", "flag", "path", "vec3"

SYSTEM_FLAGS = {
    # AnyParams variants (any.rs:10-27) -> prefix; fields of the variant's parameter struct
    "fake": {"fake-linear": FLAG, "fake-quadratic-dimensions": INT, "fake-pieces-a": F64, "fake-pieces-b": F64,
             "fake-pieces-e1": F64, "fake-pieces-e2": F64, "fake-gaussian-sigma": F64},        # fake.rs:12-36,66-74
    "fake-erfinv": {"fake-erfinv-mean-energy": F64, "fake-erfinv-N": INT},                       # erfinv.rs:11-26
    "wca": {"wca-cell-width": VEC3, "wca-cell-volume": F64, "wca-reduced-density": F64, "wca-N": INT, "wca-fcc": FLAG},  # wca.rs:361-380
    "lj": {"lj-N": INT, "lj-radius": F64},                                                        # lj.rs:14-24
    "ising": {"ising-N": INT},                                                                    # ising.rs:10-17
    "sw": {"sw-well-width": F64, "sw-cell-width": VEC3, "sw-cell-volume": F64, "sw-filling-fraction": F64, "sw-N": INT},  # optsquare.rs:326-345
    "two-wells": {"two-wells-N": INT, "two-wells-h2-to-h1": F64, "two-wells-barrier-over-h1": F64, "two-wells-r2": F64},  # two_wells.rs:13-22
    "water": {"water-N": INT},  # parsed so that the error can say why: no device kernel (out of scope)
}
METHOD_FLAGS = {"sad-min-T": F64, "samc-t0": F64, "wl": FLAG, "wl-min-gamma": F64, "Inv-t-WL": FLAG, "inv-t-wl": FLAG,
                "T": F64, "canonical-T": F64}                                                     # energy.rs:41-69
MC_FLAGS = {"seed": INT, "energy-bin": F64, "min-allowed-energy": F64, "max-allowed-energy": F64,  # energy.rs:81-97
            "translation-scale": F64, "acceptance-rate": F64,                                     # MoveParams 71-78
            "max-iter": INT, "max-independent-samples": INT, "quiet": FLAG,                       # plugin.rs:159-167
            "movie-time": F64, "save-time": F64,                                                  # plugin.rs:411-415, 322-326
            "save-as": PATH, "num-threads": INT, "resume-from": PATH}                             # mc/mod.rs:22-32
GPU_FLAGS = {"num-walkers": INT, "gpu-device": INT, "bin-window-lo": F64, "bin-window-hi": F64, "lanes-per-walker": INT,
             "fast-math": FLAG, "lj-stream-z": FLAG, "lj-smem-z": FLAG,  # SADMC_FLAG_LJ_STREAM_Z / _SMEM_Z: force one LJ31 / LJ38 layout
             "checkpoint-walkers": INT, "dry-run": FLAG, "max-launch": INT, "help": FLAG}

ALL_FLAGS = {}
for _t in list(SYSTEM_FLAGS.values()) + [METHOD_FLAGS, MC_FLAGS, GPU_FLAGS]:
    ALL_FLAGS.update(_t)


class UsageError(SystemExit):
    def __init__(self, msg):
        super().__init__("error: %s\n(see --help)" % msg)
        self.msg = msg


def parse_flags(argv):
    """`--flag value`, `--flag=value`, bare boolean flags; numbers are expressions.  Returns {flag: value}."""
    out = {}
    i = 0
    while i < len(argv):
        a = argv[i]
        if not a.startswith("--"):
            raise UsageError("unexpected argument %r" % a)
        name, eq, inline = a[2:].partition("=")
        if name not in ALL_FLAGS:
            raise UsageError("unknown flag --%s" % name)
        if name in out:
            raise UsageError("--%s given twice" % name)
        kind = ALL_FLAGS[name]
        if kind == FLAG:
            if eq:
                raise UsageError("--%s takes no value" % name)
            out[name] = True
            i += 1
            continue
        need = 3 if kind == VEC3 else 1
        if eq:
            vals = [inline] if need == 1 else inline.replace(",", " ").split()
            i += 1
        else:
            vals = argv[i + 1:i + 1 + need]
            if any(v.startswith("--") for v in vals):  # the next flag, not a value ("-1.5" is a value)
                vals = []
            i += 1 + need
        if len(vals) != need:
            raise UsageError("--%s needs %d value%s" % (name, need, "s" if need > 1 else ""))
        try:
            if kind == PATH:
                out[name] = vals[0]
            elif kind == F64:
                out[name] = evaluate(vals[0])
            elif kind == INT:
                out[name] = _as_int(name, evaluate(vals[0]))
            else:
                out[name] = tuple(evaluate(v) for v in vals)
        except ExprError as e:
            raise UsageError("--%s: %s" % (name, e))
    return out


def _one_of(flags, groups, what, required=True):
    present = [g for g, names in groups.items() if any(n in flags for n in names)]
    if len(present) > 1:
        raise UsageError("more than one %s given: %s" % (what, ", ".join(sorted(present))))
    if not present:
        if required:
            raise UsageError("no %s given (one of: %s)" % (what, ", ".join("--" + sorted(n)[0] for n in groups.values())))
        return None
    return present[0]


def config_from_flags(flags):
    """`AnyParams` + `EnergyMCParams` -> sadmc_config (+ the plugin parameters and the host-side options)."""
    system = _one_of(flags, SYSTEM_FLAGS, "system")
    kw = {}
    if system == "water":
        raise UsageError("--water-*: the water model has no device kernel (out of scope, SURVEY.md section 8)")
    if system == "lj":
        for req in ("lj-N", "lj-radius"):
            if req not in flags:
                raise UsageError("--%s is required" % req)
        kw = dict(N=flags["lj-N"], lj_radius=flags["lj-radius"])
    elif system == "ising":
        kw = dict(N=flags["ising-N"])
    elif system == "fake":
        fn = _one_of(flags, {"linear": ["fake-linear"], "quadratic": ["fake-quadratic-dimensions"],
                             "pieces": ["fake-pieces-a", "fake-pieces-b", "fake-pieces-e1", "fake-pieces-e2"],
                             "gaussian": ["fake-gaussian-sigma"]}, "fake function")
        if fn == "linear":
            kw = dict(fake_function=_abi.FAKE_LINEAR, N=1)
        elif fn == "quadratic":
            kw = dict(fake_function=_abi.FAKE_QUADRATIC, N=flags["fake-quadratic-dimensions"])
        elif fn == "pieces":
            for req in ("fake-pieces-a", "fake-pieces-b", "fake-pieces-e1", "fake-pieces-e2"):
                if req not in flags:
                    raise UsageError("--%s is required" % req)
            kw = dict(fake_function=_abi.FAKE_PIECES, N=3, fake_a=flags["fake-pieces-a"], fake_b=flags["fake-pieces-b"],
                      fake_e1=flags["fake-pieces-e1"], fake_e2=flags["fake-pieces-e2"])
        else:
            kw = dict(fake_function=_abi.FAKE_GAUSSIAN, N=3, fake_sigma=flags["fake-gaussian-sigma"])
    elif system == "fake-erfinv":
        for req in ("fake-erfinv-N", "fake-erfinv-mean-energy"):
            if req not in flags:
                raise UsageError("--%s is required" % req)
        kw = dict(N=flags["fake-erfinv-N"], erfinv_mean_energy=flags["fake-erfinv-mean-energy"])
    elif system in ("wca", "sw"):
        dims = {"cell-width": [system + "-cell-width"], "cell-volume": [system + "-cell-volume"]}
        dims["reduced-density" if system == "wca" else "filling-fraction"] = [system + ("-reduced-density" if system == "wca" else "-filling-fraction")]
        d = _one_of(flags, dims, "cell dimension")
        if system + "-N" not in flags:
            raise UsageError("--%s-N is required" % system)
        kw = dict(N=flags[system + "-N"])
        if d == "cell-width":
            kw["cell_width"] = flags[system + "-cell-width"]
        elif d == "cell-volume":
            w = float(np.cbrt(flags[system + "-cell-volume"]))
            kw["cell_width"] = (w, w, w)  # Cell::new: CellVolume(v) -> v.cbrt() on each side (optcell.rs:47-50)
        elif d == "reduced-density":
            kw["reduced_density"] = flags["wca-reduced-density"]
        else:
            kw["filling_fraction"] = flags["sw-filling-fraction"]
        if system == "sw":
            if "sw-well-width" not in flags:
                raise UsageError("--sw-well-width is required")
            kw["sw_well_width"] = flags["sw-well-width"]
        if flags.get("wca-fcc"):
            raise UsageError("--wca-fcc: the fcc start (rand's choose_multiple over the stretched grid, wca.rs:406-446) is not restated")
    elif system == "two-wells":
        for req in SYSTEM_FLAGS["two-wells"]:
            if req not in flags:
                raise UsageError("--%s is required" % req)
        kw = dict(N=flags["two-wells-N"], tw_h2_to_h1=flags["two-wells-h2-to-h1"],
                  tw_barrier_over_h1=flags["two-wells-barrier-over-h1"], tw_r2=flags["two-wells-r2"])

    method = _one_of(flags, {"sad": ["sad-min-T"], "samc": ["samc-t0"], "wl": ["wl", "wl-min-gamma"],
                             "inv-t-wl": ["Inv-t-WL", "inv-t-wl"], "canonical": ["T", "canonical-T"]}, "method")
    if method == "sad":
        kw["sad_min_T"] = flags["sad-min-T"]
    elif method == "samc":
        kw["samc_t0"] = flags["samc-t0"]
    elif method == "wl":
        if "wl-min-gamma" in flags:
            kw["wl_min_gamma"] = flags["wl-min-gamma"]
    elif method == "canonical":
        kw["canonical_T"] = flags.get("T", flags.get("canonical-T"))

    moves = _one_of(flags, {"translation-scale": ["translation-scale"], "acceptance-rate": ["acceptance-rate"]}, "move plan", required=False)
    if moves == "acceptance-rate":
        kw["move_plan"], kw["move_value"] = _abi.MOVE_ACCEPTANCE_RATE, flags["acceptance-rate"]
    elif moves == "translation-scale":
        kw["move_plan"], kw["move_value"] = _abi.MOVE_TRANSLATION_SCALE, flags["translation-scale"]
    for f, field in (("energy-bin", "energy_bin"), ("min-allowed-energy", "min_allowed_energy"), ("max-allowed-energy", "max_allowed_energy"),
                     ("bin-window-lo", "bin_window_lo"), ("bin-window-hi", "bin_window_hi"), ("lanes-per-walker", "lanes_per_walker"),
                     ("gpu-device", "device")):
        if f in flags:
            kw[field] = flags[f]
    kw["seed"] = flags.get("seed", 0)  # energy.rs:835: params.seed.unwrap_or(0)
    kw["n_walkers"] = flags.get("num-walkers", 1)
    if flags.get("fast-math"):
        kw["flags"] = _abi.FLAG_FAST_MATH
    if flags.get("lj-stream-z"):
        kw["flags"] = kw.get("flags", 0) | _abi.FLAG_LJ_STREAM_Z
    if flags.get("lj-smem-z"):
        kw["flags"] = kw.get("flags", 0) | _abi.FLAG_LJ_SMEM_Z
    return _abi.make_config(system, method, **kw)


def plugin_params(flags):
    """ReportParams / SaveParams / MovieParams with their defaults (plugin.rs:169-177, 328-334, 417-421)."""
    return dict(max_iter=flags.get("max-iter"), max_independent_samples=flags.get("max-independent-samples"),
                quiet=bool(flags.get("quiet", False)),  # a bool field is false unless its flag is given
                save_time=flags.get("save-time", 1.0), movie_time=flags.get("movie-time"))


def config_summary(cfg):
    d = {}
    for name, _ in cfg._fields_:
        if name.startswith("_"):
            continue
        v = getattr(cfg, name)
        if name == "cell_width":
            v = [v[0], v[1], v[2]]
        if isinstance(v, float) and v != v:
            v = None
        d[name] = v
    return d


HELP = __doc__ + "\nFlags:\n" + "\n".join(
    "  --%s%s" % (n, {F64: " <f64 expression>", INT: " <integer expression>", FLAG: "", PATH: " <path>", VEC3: " <x> <y> <z>"}[k])
    for n, k in ALL_FLAGS.items())


def shard_env():
    """(world, rank, local rank) of a one-process-per-GPU launch (torchrun / mpirun export these); (1, 0, 0) otherwise."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    return world, rank, int(os.environ.get("LOCAL_RANK", rank))


def rank_path(path, rank, world):
    """`name.ext` -> `name.rank3of8.ext` when the walkers are sharded over several processes."""
    if world == 1:
        return path
    stem, ext = os.path.splitext(path)
    return "%s.rank%dof%d%s" % (stem, rank, world, ext)


def main(argv=None, out=print):
    argv = list(sys.argv[1:] if argv is None else argv)
    flags = parse_flags(argv)
    if flags.get("help"):
        out(HELP)
        return 0
    pp = plugin_params(flags)
    from . import checkpoint, plugins
    n_walkers = flags.get("num-walkers", 1)
    # One process per GPU (torchrun --nproc-per-node G -m sad_monte_carlo_b200.histogram ...): --num-walkers is the total,
    # rank r runs global walkers [r W/G, (r+1) W/G) on device LOCAL_RANK -- walker w is still the reference run with
    # --seed seed+w, whatever G is -- and writes its own files `name.rankRofG[-wNNNNNN].ext`.  No traffic between ranks.
    world, rank, local_rank = shard_env()
    if n_walkers % world:
        raise UsageError("--num-walkers %d does not divide over %d processes" % (n_walkers, world))
    n_walkers //= world

    def place(cfg):
        if world > 1:
            cfg.n_walkers = n_walkers
            cfg.walker_offset = rank * n_walkers
            if "gpu-device" not in flags:
                cfg.device = local_rank
        return cfg

    if "resume-from" in flags:  # Params::ResumeFrom, mc/mod.rs:92-106: nothing else is read from the command line
        path = rank_path(flags["resume-from"], rank, world)
        if os.path.splitext(path)[1].lstrip(".") not in ("yaml", "json", "cbor"):
            raise UsageError("I don't know how to read file %r" % path)
        try:
            doc0 = checkpoint.check_resumable(None, path, n_walkers)
        except ValueError as ex:
            raise UsageError(str(ex))
        over = {k: flags[f] for f, k in (("bin-window-lo", "bin_window_lo"), ("bin-window-hi", "bin_window_hi"), ("gpu-device", "device")) if f in flags}
        cfg = place(checkpoint.config_from_document(doc0, n_walkers=n_walkers, **over))
        save_as = doc0.get("save_as", path) if n_walkers == 1 else path
        rep = doc0.get("report", {})
        mi = rep.get("max_iter", "Never")
        pp = dict(max_iter=None if mi == "Never" else mi["TotalMoves"], max_independent_samples=rep.get("max_independent_samples"),
                  quiet=rep.get("quiet", True), save_time=(doc0.get("save", {}).get("save_time_seconds") or 3600.0) / 3600.0,
                  movie_time=doc0.get("movies", {}).get("movie_time"))
        if flags.get("dry-run"):
            out(json.dumps({"resume_from": path, "config": config_summary(cfg), "plugins": pp}))
            return 0
        engine = checkpoint.resume(cfg, path)
        out("Resuming from file %r" % path)
        movie_state = doc0.get("movies")
        resumed = True
    else:
        cfg = place(config_from_flags(flags))
        save_as = rank_path(flags.get("save-as", "resume.yaml"), rank, world)  # mc/mod.rs:88
        if os.path.splitext(save_as)[1].lstrip(".") not in ("yaml", "json", "cbor"):
            raise UsageError("I don't know how to create file %r" % save_as)  # mc/mod.rs:118
        first = checkpoint.walker_path(save_as, 0, n_walkers)
        # Every rank takes the same decision, from the files of ALL ranks (one node, one file system): a run resumes
        # when some rank has a checkpoint, and then every rank's set must be complete and written for this command line.
        all_sets = [rank_path(flags.get("save-as", "resume.yaml"), r, world) for r in range(world)]
        resuming = "save-as" in flags and any(
            os.path.exists(checkpoint.walker_path(p, 0, n_walkers)) or os.path.exists(os.path.splitext(p)[0] + ".partial") for p in all_sets)
        if resuming:
            if "checkpoint-walkers" in flags and flags["checkpoint-walkers"] < n_walkers:
                raise UsageError("--checkpoint-walkers %d writes a partial set that cannot be resumed; %s exists: remove it or "
                                 "drop --checkpoint-walkers" % (flags["checkpoint-walkers"], first))
            try:
                for p in all_sets:
                    checkpoint.check_resumable(cfg, p, n_walkers)
            except ValueError as ex:
                raise UsageError(str(ex))
        if flags.get("dry-run"):
            out(json.dumps({"config": config_summary(cfg), "plugins": pp, "save_as": save_as, "resuming": resuming}))
            return 0
        movie_state = None
        resumed = resuming
        if resuming:  # mc/mod.rs:70-84, then update_from_params (energy.rs:899-902): report + save come from the flags
            engine = checkpoint.resume(cfg, save_as)
            out("Resuming from file %r" % save_as)
            movie_state = checkpoint.load(first).get("movies")
        else:
            from .engine import WalkerEngine
            engine = WalkerEngine(cfg)

    report = plugins.Report(max_iter=pp["max_iter"], max_independent_samples=pp["max_independent_samples"], quiet=pp["quiet"], out=out,
                            resumed=resumed)
    save = plugins.Save(save_time_hours=pp["save_time"], resumed=resumed)
    movies = plugins.Movie(movie_time=pp["movie_time"])
    if movie_state is not None:
        movies.restore(movie_state)
    ck = flags.get("checkpoint-walkers")
    launches = plugins.run_simulation(engine, report, save, movies, save_as=save_as,
                                      checkpoint_walkers=None if ck is None else range(min(ck, engine.n_walkers)),
                                      max_launch=flags.get("max-launch"))
    if not pp["quiet"]:
        out("%d moves per walker, %d walkers, %d launches" % (engine.num_moves(), engine.n_walkers, launches))
    engine.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
