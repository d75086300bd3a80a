"""The reference's `binning` command line on the GPU engine (src/bin/binning.rs: `EnergyMC<Any>` of
src/mc/energy_binning.rs over `binning::Bins`).

    python -m sad_monte_carlo_b200.binning --fake-linear --histogram-bin 0.01 --translation-scale 0.05 \\
        --sad-min-T 0.001 --max-iter 1e9 --save-as sad-linear-0.01.yaml            (fake/run-fake.py:25-36)

Flags are the `histogram` command line's (histogram.py) with `BinningParams` in place of `--energy-bin`
(binning.rs:50-69): `--histogram-bin <de>` (default 1.0) or `--linear-bin <de>` (binning::linear: a correctness path on the device; its
checkpoints are written, resuming them is not built), and `--high-resolution-de <de>` (energy_binning.rs:62-63).  Checkpoints are written in the reference's serde schema for this Monte Carlo -- one document per
walker, `bins: {Histogram: {min, min_e, max_e, width, lnw: BinCounts, extra: {name: BinCounts}}}` (histogram.rs:12-32,
99-111) -- so `plotting/parse-binning.py` reads them.  `--save-as` on an existing checkpoint set resumes it
(mc/mod.rs:66-84: state from the file, report / save parameters from the flags) and continues bit for bit.
"""
import json
import os
import sys

import numpy as np

from . import _abi
from . import histogram as H

BINNING_FLAGS = {"histogram-bin": H.F64, "linear-bin": H.F64, "high-resolution-de": H.F64}
H.ALL_FLAGS.update(BINNING_FLAGS)


def _bincounts(total, count, max_count=None, max_total=None, centres=None):
    """`BinCounts` (histogram.rs:12-32).  The aggregates the sampler never reads are reported as what a full rescan of
    the vectors gives (the reference maintains them lazily, histogram.rs:60-80); max_count / max_total, which the
    sampler does read, are the engine's exact running values where given."""
    total = np.asarray(total, float)
    count = np.asarray(count)
    n = len(total)
    i_t = int(np.argmax(total)) if n else 0
    i_c = int(np.argmax(count)) if n else 0
    return {"total": [float(x) for x in total], "min_total": float(total.min()) if n else 0.0,
            "max_total": float(max_total if max_total is not None else (total.max() if n else 0.0)),
            "e_max_total": float(centres[i_t]) if n and centres is not None else float("-inf"),
            "count": [int(x) for x in count], "min_count": int(count.min()) if n else 0,
            "max_count": int(max_count if max_count is not None else (count.max() if n else 0)),
            "e_max_count": float(centres[i_c]) if n and centres is not None else float("-inf"),
            "total_count": int(count.sum())}


def _high_resolution_document(engine, w, st):
    """`high_resolution: Option<histogram::Bins>` (energy_binning.rs:124-125): counts only, no extras."""
    de = engine.cfg.high_resolution_de
    if not de > 0:
        return None
    mn, cnt = engine.high_resolution(w)
    centres = mn + (np.arange(len(cnt)) + 0.5) * de
    return {"min": mn, "min_e": st.bins_min_e, "max_e": st.bins_max_e, "width": de, "lnw": _bincounts(np.zeros(len(cnt)), cnt, centres=centres),
            "extra": {}}


def _common_document(engine, w, st, save_as, report, movies, save):
    """Everything of energy_binning.rs's `EnergyMC` (92-126) except `bins`."""
    from .checkpoint import _opt, _system_document
    cfg = engine.cfg
    m = st.method
    if m == _abi.METHOD_SAD:
        method = {"Sad": {"num_states": int(st.num_states), "min_T": cfg.sad_min_T, "too_lo": st.too_lo, "too_hi": st.too_hi, "tL": int(st.tL),
                          "tF": st.tF, "latest_parameter": st.latest_parameter}}
    elif m == _abi.METHOD_SAMC:
        method = {"Samc": {"t0": st.samc_t0}}
    else:
        method = {"WL": {"gamma": st.wl_gamma, "inv_t": bool(st.wl_inv_t), "min_gamma": _opt(cfg.wl_min_gamma)}}
    move_plan = ({"TranslationScale": cfg.move_value} if cfg.move_plan == _abi.MOVE_TRANSLATION_SCALE else {"AcceptanceRate": cfg.move_value})
    return {
        "system": _system_document(cfg, engine.system(w), engine.cell_box() if cfg.system in (_abi.SYS_WCA, _abi.SYS_SW) else None),
        "method": method, "moves": int(st.moves), "time_L": 0, "accepted_moves": int(st.accepted_moves),
        "min_allowed_energy": _opt(cfg.min_allowed_energy), "max_allowed_energy": _opt(cfg.max_allowed_energy),
        "move_plan": move_plan, "translation_scale": st.translation_scale, "acceptance_rate": st.acceptance_rate,
        "rng": {"s0": int(st.rng_s0), "s1": int(st.rng_s1)}, "save_as": str(save_as),
        "report": report if report is not None else {"max_iter": "Never", "max_independent_samples": None, "quiet": True},
        "save": save if save is not None else {"save_time_seconds": 3600.0},
        "movies": movies if movies is not None else {"movie_time": None, "which_frame": 0, "period": "Never"},
        "manager": {},
        "bins": None,
        "high_resolution": _high_resolution_document(engine, w, st),
    }


def walker_document(engine, w, save_as="resume.yaml", report=None, movies=None, save=None):
    """The serde document of walker `w`: energy_binning.rs:92-126 (`EnergyMC`), 128-148 (`Method`), binning.rs:71-78 (`Bins`)."""
    from .checkpoint import EXTRA_LABEL
    cfg = engine.cfg
    st = engine.binning_walker(w)
    if st.status != 0:
        raise RuntimeError("walker %d is halted (status %d)" % (w, st.status))
    if cfg.flags & _abi.FLAG_BINNING_LINEAR:
        return _linear_walker_document(engine, w, st, save_as, report, movies, save)
    b = engine.binning_bins(w)
    n = st.bins_len
    centres = st.bins_min + (np.arange(n) + 0.5) * st.bins_width
    extra = {}
    if n:
        extra["energy"] = _bincounts(b["energy_total"], b["energy_count"], centres=centres)
        if b["t_found_count"].any():
            extra["t_found"] = _bincounts(b["t_found_total"], b["t_found_count"], max_total=st.t_found_max_total, centres=centres)
        if st.method in (_abi.METHOD_WL, _abi.METHOD_INV_T_WL) or b["hist_count"].any():
            extra["hist"] = _bincounts(np.zeros(n), b["hist_count"], centres=centres)
        if cfg.system in EXTRA_LABEL and b["extra_count"].any():
            extra[EXTRA_LABEL[cfg.system]] = _bincounts(b["extra_total"], b["extra_count"], centres=centres)
    lnw = _bincounts(b["lnw_total"], b["lnw_count"], max_count=st.lnw_max_count, centres=centres)
    lnw["total_count"] = int(st.lnw_total_count)
    doc = _common_document(engine, w, st, save_as, report, movies, save)
    doc["bins"] = {"Histogram": {"min": st.bins_min, "min_e": st.bins_min_e, "max_e": st.bins_max_e, "width": st.bins_width, "lnw": lnw,
                                 "extra": extra}}
    return doc


def restore_walker(engine, w, doc):
    """Feed a `binning` document back into walker `w` of an engine created with INIT_EXTERNAL + FLAG_BINNING (then
    engine.resume(moves)): what `MonteCarlo::from_args` does with an existing --save-as file (mc/mod.rs:70-84)."""
    from .checkpoint import _system_image
    cfg = engine.cfg
    engine.set_system(w, _system_image(cfg, doc["system"], engine.system_len))
    st = _abi.BinningState()
    st.moves, st.accepted_moves = doc["moves"], doc["accepted_moves"]
    st.acceptance_rate, st.translation_scale = doc["acceptance_rate"], doc["translation_scale"]
    st.rng_s0, st.rng_s1 = doc["rng"]["s0"], doc["rng"]["s1"]
    kind, h = next(iter(doc["bins"].items()))
    if kind != "Histogram":
        raise NotImplementedError("bins variant %r has no device kernel" % kind)
    n = len(h["lnw"]["total"])
    st.bins_min, st.bins_width, st.bins_len = h["min"], h["width"], n
    st.bins_min_e, st.bins_max_e = h["min_e"], h["max_e"]
    st.lnw_max_count, st.lnw_total_count = h["lnw"]["max_count"], h["lnw"]["total_count"]
    mtag, m = next(iter(doc["method"].items()))
    if mtag == "Sad":
        st.method = _abi.METHOD_SAD
        st.too_lo, st.too_hi, st.latest_parameter, st.tF = m["too_lo"], m["too_hi"], m["latest_parameter"], m["tF"]
        st.tL, st.num_states = m["tL"], m["num_states"]
    elif mtag == "Samc":
        st.method, st.samc_t0 = _abi.METHOD_SAMC, m["t0"]
    else:
        st.method = _abi.METHOD_INV_T_WL if m["inv_t"] else _abi.METHOD_WL
        st.wl_gamma, st.wl_inv_t = m["gamma"], int(m["inv_t"])
    ex = h.get("extra", {})
    arrays = {"lnw_total": h["lnw"]["total"], "lnw_count": h["lnw"]["count"]}
    zero_f, zero_u = np.zeros(n), np.zeros(n, np.uint64)
    arrays["energy_total"] = ex["energy"]["total"] if "energy" in ex else zero_f
    arrays["energy_count"] = ex["energy"]["count"] if "energy" in ex else zero_u
    if "t_found" in ex:
        arrays["t_found_total"], arrays["t_found_count"] = ex["t_found"]["total"], ex["t_found"]["count"]
        st.t_found_max_total = ex["t_found"]["max_total"]
    if "hist" in ex:
        arrays["hist_count"] = ex["hist"]["count"]
        st.hist_min_count, st.hist_total_count = ex["hist"]["min_count"], ex["hist"]["total_count"]
    for label, bc in ex.items():
        if label not in ("energy", "t_found", "hist"):
            arrays["extra_total"], arrays["extra_count"] = bc["total"], bc["count"]
    tag, body = next(iter(doc["system"].items()))
    st.energy = body["E"] if "E" in body else engine.compute_energy(w)  # the analytic systems keep no cached energy
    hr = doc.get("high_resolution")
    if cfg.high_resolution_de > 0:
        if hr is None or hr["width"] != cfg.high_resolution_de:
            raise ValueError("the checkpoint's high_resolution histogram does not match --high-resolution-de %r" % cfg.high_resolution_de)
        engine.set_high_resolution(w, hr["min"], hr["lnw"]["count"])
    elif hr is not None:
        raise ValueError("the checkpoint carries a high_resolution histogram: give --high-resolution-de %r" % hr["width"])
    engine.set_binning_walker(w, st, arrays)


def resume(cfg, save_as):
    """An engine continued from the per-walker documents written by save()."""
    from . import checkpoint as ck
    from .engine import WalkerEngine
    docs = [ck.load(ck.walker_path(save_as, w, cfg.n_walkers)) for w in range(cfg.n_walkers)]
    moves = {d["moves"] for d in docs}
    if len(moves) != 1:
        raise ValueError("walker checkpoints disagree on moves: %s" % sorted(moves)[:4])
    for d in docs:
        if "Linear" in d.get("bins", {}):
            raise ValueError("%s holds binning::linear bins: resuming them is not built" % save_as)
        if "Histogram" not in d.get("bins", {}):
            raise ValueError("%s is not a `binning` checkpoint (bins: {Histogram: ...})" % save_as)
        if d["bins"]["Histogram"]["width"] != cfg.energy_bin:
            raise ValueError("checkpoint bin width %r differs from --histogram-bin %r" % (d["bins"]["Histogram"]["width"], cfg.energy_bin))
    cfg.init_mode = _abi.INIT_EXTERNAL
    engine = WalkerEngine(cfg)
    for w, d in enumerate(docs):
        restore_walker(engine, w, d)
    engine.resume(moves.pop())
    return engine


def _linear_walker_document(engine, w, st, save_as, report, movies, save):
    """The same document with `bins: {Linear: ...}`: linear.rs:153-165, its BinCounts (12-32) keep f64 counts."""
    from .checkpoint import EXTRA_LABEL
    cfg = engine.cfg
    b = engine.binning_bins_f64(w)
    n = st.bins_len
    centres = st.bins_min + (np.arange(n) + 0.5) * st.bins_width

    def bc(total, count, max_count=None, max_total=None):
        d = _bincounts(total, np.zeros(len(total), np.uint64), max_total=max_total, centres=centres)
        count = np.asarray(count, float)
        d["count"] = [float(x) for x in count]
        d["min_count"] = float(count.min()) if n else 0.0
        d["max_count"] = float(max_count if max_count is not None else (count.max() if n else 0.0))
        d["e_max_count"] = float(centres[int(np.argmax(count))]) if n else float("-inf")
        return d
    extra = {}
    if n:
        extra["energy"] = bc(b["energy_total"], b["energy_count"])
        extra["energy"]["total_count"] = int(st.moves)
        if b["t_found_count"].any():
            extra["t_found"] = bc(b["t_found_total"], b["t_found_count"], max_total=st.t_found_max_total)
            extra["t_found"]["total_count"] = int(round(b["t_found_count"].sum()))
        if st.method in (_abi.METHOD_WL, _abi.METHOD_INV_T_WL) or b["hist_count"].any():
            extra["hist"] = bc(np.zeros(n), b["hist_count"])
            extra["hist"]["total_count"] = int(st.hist_total_count)
        if cfg.system in EXTRA_LABEL and b["extra_count"].any():
            extra[EXTRA_LABEL[cfg.system]] = bc(b["extra_total"], b["extra_count"])
            extra[EXTRA_LABEL[cfg.system]]["total_count"] = int(round(b["extra_count"].sum()))
    lnw = bc(b["lnw_total"], b["lnw_count"], max_count=st.lnw_max_count_f64)
    lnw["total_count"] = int(st.lnw_total_count)
    doc = _common_document(engine, w, st, save_as, report, movies, save)
    doc["bins"] = {"Linear": {"min": st.bins_min, "min_e": st.bins_min_e, "max_e": st.bins_max_e, "width": st.bins_width, "lnw": lnw, "extra": extra}}
    return doc


def save(engine, save_as, walkers=None, **plugin_docs):
    """MonteCarlo::checkpoint for a binning engine: one file per walker, all or nothing (as checkpoint.save)."""
    from . import checkpoint as ck
    ext = os.path.splitext(str(save_as))[1].lstrip(".")
    walkers = list(range(engine.n_walkers) if walkers is None else walkers)
    left, failed = engine.num_halted()
    if left or failed:
        raise RuntimeError("no checkpoint written: %d walker(s) left the bin window and %d failed verify_energy" % (left, failed))
    staged = []
    try:
        for w in walkers:
            p = ck.walker_path(save_as, w, engine.n_walkers)
            staged.append((ck.stage(p, ck.dumps(walker_document(engine, w, save_as=p, **plugin_docs), ext)), p))
        for tmp, p in staged:
            os.replace(tmp, p)
    except BaseException:
        for tmp, _ in staged:
            if os.path.exists(tmp):
                os.unlink(tmp)
        raise
    return [p for _, p in staged]


def config_from_flags(flags):
    """`AnyParams` + energy_binning.rs `EnergyMCParams` (52-69) -> sadmc_config with SADMC_FLAG_BINNING."""
    if "linear-bin" in flags and "histogram-bin" in flags:
        raise H.UsageError("more than one binning given: --histogram-bin, --linear-bin (BinningParams, binning.rs:50-61)")
    if "energy-bin" in flags:
        raise H.UsageError("--energy-bin belongs to `histogram`; `binning` takes --histogram-bin (binning.rs:50-69)")
    if "T" in flags or "canonical-T" in flags:
        raise H.UsageError("energy_binning.rs has no canonical method (MethodParams, energy_binning.rs:22-40)")
    f = dict(flags)
    linear = "linear-bin" in f
    f["energy-bin"] = f.pop("linear-bin") if linear else f.pop("histogram-bin", 1.0)  # BinningParams::default: Histogram { bin: 1.0 }
    hr = f.pop("high-resolution-de", None)
    cfg = H.config_from_flags(f)
    cfg.flags |= _abi.FLAG_BINNING | (_abi.FLAG_BINNING_LINEAR if linear else 0)
    if hr is not None:
        if not hr > 0:
            raise H.UsageError("--high-resolution-de must be positive (histogram.rs:148)")
        cfg.high_resolution_de = hr
    return cfg


def main(argv=None, out=print):
    argv = list(sys.argv[1:] if argv is None else argv)
    flags = H.parse_flags(argv)
    if flags.get("help"):
        out(__doc__)
        return 0
    if "resume-from" in flags:
        raise H.UsageError("--resume-from: give the system flags and --save-as <the checkpoint> instead (the configuration is not rebuilt from a `binning` document)")
    pp = H.plugin_params(flags)
    cfg = config_from_flags(flags)
    save_as = flags.get("save-as", "resume.yaml")
    if os.path.splitext(save_as)[1].lstrip(".") not in ("yaml", "json", "cbor"):
        raise H.UsageError("I don't know how to create file %r" % save_as)
    from . import checkpoint as ck
    resuming = "save-as" in flags and os.path.exists(ck.walker_path(save_as, 0, cfg.n_walkers))  # mc/mod.rs:66-84
    if flags.get("dry-run"):
        out(json.dumps({"config": H.config_summary(cfg), "binning": {"Histogram": {"bin": cfg.energy_bin}}, "plugins": pp, "save_as": save_as,
                        "resuming": resuming}))
        return 0
    from . import plugins
    from .engine import WalkerEngine
    movie_state = None
    if resuming:
        try:
            engine = resume(cfg, save_as)
        except (ValueError, OSError) as ex:
            raise H.UsageError(str(ex))
        out("Resuming from file %r" % save_as)
        movie_state = ck.load(ck.walker_path(save_as, 0, cfg.n_walkers)).get("movies")
    else:
        engine = WalkerEngine(cfg)

    class BinningMC(plugins.EngineMC):
        def checkpoint(self):
            return save(self.engine, self.save_as, walkers=self.checkpoint_walkers, **self._docs())

        def save_movie_frame(self, moves):
            d = os.path.splitext(self.save_as)[0]
            return save(self.engine, os.path.join(d, "%014d.cbor" % moves), walkers=self.checkpoint_walkers, **self._docs())

    report = plugins.Report(pp["max_iter"], pp["max_independent_samples"], pp["quiet"], out=out, resumed=resuming)
    saver = plugins.Save(pp["save_time"], resumed=resuming)
    movies = plugins.Movie(pp["movie_time"])
    if movie_state is not None:
        movies.restore(movie_state)
    kw = flags.get("checkpoint-walkers")
    mc = BinningMC(engine, save_as, list(range(kw)) if kw is not None else None, report, saver, movies)
    manager = plugins.PluginManager()
    launches = 0
    while True:
        n = manager.moves_until_next_action()
        if "max-launch" in flags:
            n = min(n, flags["max-launch"])
        engine.run(n)
        launches += 1
        if manager.run(mc, [report, saver, movies], moves_made=n) == plugins.Action.EXIT:
            break
    engine.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
