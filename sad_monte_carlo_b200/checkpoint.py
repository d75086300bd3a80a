"""Checkpoints in the reference's serde schema, one document per walker.

The reference serialises the whole `EnergyMC<Any>` with serde into yaml / json / cbor chosen by the file extension
(`MonteCarlo::checkpoint`, src/mc/mod.rs:110-120), through `AtomicFile` (src/atomicfile.rs: write to a temporary
name, rename on success), and reads it back on `--save-as` / `--resume-from` (mc/mod.rs:70-106).  The Python tools
(plotting/parse-binning.py:103-170) index `data['bins']`, `data['method']['Sad']`, `data['moves']`, ...

`walker_document` builds exactly that mapping for one GPU walker from the C ABI getters; `save` / `load` write and
read the three formats; `restore_walker` feeds a document back through the resume entry points
(sadmc_set_system / sadmc_set_walker_bins), so a GPU run can be checkpointed, inspected with the reference's
scripts and continued.  Externally tagged enums, `Option::None` = null, unit newtypes = bare f64, as serde emits.
Host-side only; nothing here is on the hot path.
"""
import json
import math
import os
import struct
import tempfile

import numpy as np

from . import _abi

SYSTEM_TAGS = {_abi.SYS_LJ: "Lj", _abi.SYS_ISING: "Ising", _abi.SYS_FAKE: "Fake", _abi.SYS_WCA: "Wca", _abi.SYS_SW: "Sw",
               _abi.SYS_TWO_WELLS: "TwoWells", _abi.SYS_FAKE_ERFINV: "FakeErfinv"}  # Any variants, src/system/any.rs:61-78
EXTRA_LABEL = {_abi.SYS_WCA: "pressure", _abi.SYS_TWO_WELLS: "which"}           # data_to_collect labels


# ---- one walker <-> the reference's document ------------------------------------------------------------------------

def _opt(x):
    return None if (isinstance(x, float) and math.isnan(x)) else x


def _system_document(cfg, image, cell=None):
    """`system` field: the `Any` variant of this walker (positions etc.) from the ABI's f64 image."""
    if cfg.system == _abi.SYS_LJ:  # lj.rs:32-45
        n = cfg.N
        pos = np.asarray(image[:3 * n]).reshape(n, 3)
        return {"Lj": {"E": float(image[3 * n]), "error": float(image[3 * n + 1]), "possible_change": "None",
                       "positions": [{"x": float(p[0]), "y": float(p[1]), "z": float(p[2])} for p in pos],
                       "max_radius_squared": cfg.lj_radius * cfg.lj_radius, "max_radius": cfg.lj_radius}}
    if cfg.system == _abi.SYS_ISING:  # ising.rs:20-29
        n = cfg.N
        return {"Ising": {"E": float(image[n * n]), "N": int(n), "S": [int(s) for s in image[:n * n]], "possible_change": None}}
    if cfg.system == _abi.SYS_FAKE:  # fake.rs:76-83, Function 12-36
        fn = cfg.fake_function
        if fn == _abi.FAKE_LINEAR:
            function, dim = "Linear", 1
        elif fn == _abi.FAKE_QUADRATIC:
            function, dim = {"Quadratic": {"dimensions": int(cfg.N)}}, int(cfg.N)
        elif fn == _abi.FAKE_PIECES:
            function, dim = {"Pieces": {"a": cfg.fake_a, "b": cfg.fake_b, "e1": cfg.fake_e1, "e2": cfg.fake_e2}}, 3
        else:
            function, dim = {"Gaussian": {"sigma": cfg.fake_sigma}}, 3
        return {"Fake": {"position": [float(x) for x in image[:dim]], "function": function, "possible_change": [0.0] * dim}}
    if cfg.system in (_abi.SYS_WCA, _abi.SYS_SW):  # wca.rs:23-33, optsquare.rs:24-31 around optcell.rs:27-40
        n = cfg.N
        pos = np.asarray(image[:3 * n]).reshape(n, 3)
        box, r_cutoff = cell  # from the engine (sadmc_cell_box): a cube root taken twice need not round the same way
        sw = cfg.system == _abi.SYS_SW
        cell = {"box_diagonal": {"x": box[0], "y": box[1], "z": box[2]}, "r_cutoff": r_cutoff,
                "positions": [{"x": float(p[0]), "y": float(p[1]), "z": float(p[2])} for p in pos]}  # subcells: #[serde(skip)]
        if sw:
            return {"Sw": {"E": float(image[3 * n]), "cell": cell, "possible_change": "None"}}
        return {"Wca": {"E": float(image[3 * n]), "error": float(image[3 * n + 1]), "cell": cell, "possible_change": "None"}}
    if cfg.system == _abi.SYS_TWO_WELLS:  # two_wells.rs:219-232
        n = cfg.N
        params = {"N": int(n), "h2_to_h1": cfg.tw_h2_to_h1, "barrier_over_h1": cfg.tw_barrier_over_h1, "r2": cfg.tw_r2}
        well = math.sqrt(cfg.tw_barrier_over_h1) * 1.0 + cfg.tw_r2 * math.sqrt(1.0 + cfg.tw_barrier_over_h1 - 1.0 / cfg.tw_h2_to_h1)
        return {"TwoWells": {"position": [float(x) for x in image[:n]], "d_squared": float(image[n]), "parameters": params,
                             "change": {"index": 0, "values": {"x": 0.0, "y": 0.0, "z": 0.0}},
                             "well_position": well, "invcdf": two_wells_invcdf(int(n), cfg.tw_r2)}}
    if cfg.system == _abi.SYS_FAKE_ERFINV:  # erfinv.rs:29-38
        return {"FakeErfinv": {"position": [float(x) for x in image[:cfg.N]], "parameters": {"mean_energy": cfg.erfinv_mean_energy},
                               "possible_change": []}}
    raise NotImplementedError("no checkpoint document for system kind %d" % cfg.system)


_INVCDF = {}


def two_wells_invcdf(dim, r2, num_points=10000, mult=100):
    """`SystemInvCdf::new` (two_wells.rs:46-137): cumulative distributions used only by TwoWells::randomize, which no
    `EnergyMC` run calls.  They are derived data, written so that the reference can deserialise the document; the
    midpoint sums are vectorised here, so the last digits may differ from the reference's sequential loop."""
    key = (dim, r2)
    if key in _INVCDF:
        return _INVCDF[key]
    r1 = 1.0

    def V(n):  # two_wells.rs:211-213
        return math.pi ** (0.5 * n) / math.gamma(n * 0.5 + 1.0)

    def cumulative(grid, pdf):
        a, b = grid[:-1, None], grid[1:, None]
        i = np.arange(mult)[None, :]
        us = ((mult - 1 - i) * a + i * b) / (mult - 1)  # linspace(x[w], x[w+1], mult), two_wells.rs:36-43
        du = us[:, 1] - us[:, 0]
        mid = 0.5 * (us[:, 1:] + us[:, :-1])
        return np.concatenate([[0.0], np.cumsum((du[:, None] * pdf(mid)).sum(axis=1))])

    def pdf_x1(x):
        x = np.asarray(x)
        first = np.sqrt(np.maximum(r1 * r1 - x * x, 0.0)) ** (dim - 1)
        hemi = np.sqrt(np.maximum(r2 * r2 - (x - r1 - r2) ** 2, 0.0)) ** (dim - 1)
        return np.where(x <= math.sqrt(r1 * r1 - r2 * r2), first, np.where(x < r1 + r2, r2 ** (dim - 1), hemi)) * V(dim - 1)

    i = np.arange(num_points)
    x1 = ((num_points - 1 - i) * (-r1) + i * (r1 + 2.0 * r2)) / (num_points - 1)
    stencils = np.zeros(num_points * dim)
    c = cumulative(x1, pdf_x1)
    stencils[:num_points] = c / c[-1]
    xn = ((num_points - 1 - i) * (-1.0) + i * 1.0) / (num_points - 1)
    for which in range(1, dim):
        d = dim - which
        c = cumulative(xn, lambda x: np.maximum(1.0 - x * x, 0.0) ** (0.5 * d) * V(d) / V(d + 1))
        stencils[which * num_points:(which + 1) * num_points] = c / c[-1]
    out = {"num_points": num_points, "dim": dim, "r1": r1, "r2": r2, "dx1_ball1": float(x1[1] - x1[0]),
           "stencils": [float(v) for v in stencils]}
    _INVCDF[key] = out
    return out


def _system_image(cfg, doc, length):
    """Inverse of _system_document: the ABI f64 image of the walker's system."""
    img = np.zeros(length)
    tag, body = next(iter(doc.items()))
    if tag == "Lj":
        n = len(body["positions"])
        for i, p in enumerate(body["positions"]):
            img[3 * i:3 * i + 3] = (p["x"], p["y"], p["z"])
        img[3 * n], img[3 * n + 1] = body["E"], body["error"]
    elif tag == "Ising":
        n = body["N"]
        img[:n * n] = body["S"]
        img[n * n] = body["E"]
    elif tag in ("Fake", "FakeErfinv"):
        img[:len(body["position"])] = body["position"]
    elif tag in ("Wca", "Sw"):
        pos = body["cell"]["positions"]
        n = len(pos)
        for i, p in enumerate(pos):
            img[3 * i:3 * i + 3] = (p["x"], p["y"], p["z"])
        img[3 * n] = body["E"]
        img[3 * n + 1] = body.get("error", 0.0)
    elif tag == "TwoWells":
        n = len(body["position"])
        img[:n] = body["position"]
        img[n] = body["d_squared"]
    else:
        raise NotImplementedError("cannot restore system variant %r" % tag)
    return img


def _method_document(cfg, st, bins):
    m = st.method
    if m == _abi.METHOD_SAD:  # energy.rs:215-226
        return {"Sad": {"min_T": cfg.sad_min_T, "too_lo": st.too_lo, "too_hi": st.too_hi, "tL": int(st.tL), "tF": int(st.tF),
                        "num_states": int(st.num_states), "highest_hist": int(st.highest_hist), "version": "Sad",
                        "latest_parameter": st.latest_parameter}}
    if m == _abi.METHOD_SAMC:  # 227-228 (also what 1/t-WL turns into, 754-756)
        return {"Samc": {"t0": st.samc_t0}}
    if m in (_abi.METHOD_WL, _abi.METHOD_INV_T_WL):  # 229-240
        return {"WL": {"gamma": st.wl_gamma, "lowest_hist": int(st.wl_lowest_hist), "highest_hist": int(st.wl_highest_hist),
                       "total_hist": int(st.wl_total_hist), "num_states": st.wl_num_states,
                       "hist": [int(x) for x in bins["wl_hist"][:st.wl_hist_len]], "min_energy": st.wl_min_energy,
                       "inv_t": bool(st.wl_inv_t), "min_gamma": _opt(cfg.wl_min_gamma)}}
    return {"Canonical": {"temperature": cfg.canonical_T}}


def walker_document(engine, w, save_as="resume.yaml", report=None, movies=None, save=None):
    """The serde document of walker `w` as the reference would write it (energy.rs:167-210)."""
    cfg = engine.cfg
    st = engine.walker(w)
    if st.status != 0:
        raise RuntimeError("walker %d is halted (status %d)" % (w, st.status))
    b = engine.bins(w)
    extra = {}
    if cfg.system in EXTRA_LABEL and b["extra_count"].any():  # Bins::accumulate_extra creates the entry on first use
        extra[EXTRA_LABEL[cfg.system]] = {"total": [float(x) for x in b["extra_total"]], "count": [int(x) for x in b["extra_count"]]}
    move_plan = ({"TranslationScale": cfg.move_value} if cfg.move_plan == _abi.MOVE_TRANSLATION_SCALE
                 else {"AcceptanceRate": cfg.move_value})
    return {
        "system": _system_document(cfg, engine.system(w), engine.cell_box() if cfg.system in (_abi.SYS_WCA, _abi.SYS_SW) else None),
        "method": _method_document(cfg, st, b),
        "moves": int(st.moves), "time_L": 0, "accepted_moves": int(st.accepted_moves),
        "min_allowed_energy": _opt(cfg.min_allowed_energy), "max_allowed_energy": _opt(cfg.max_allowed_energy),
        "move_plan": move_plan, "translation_scale": st.translation_scale, "acceptance_rate": st.acceptance_rate,
        "rng": {"s0": int(st.rng_s0), "s1": int(st.rng_s1)},  # rand_xoshiro "serde1"
        "save_as": str(save_as),
        "report": report if report is not None else {"max_iter": "Never", "max_independent_samples": None, "quiet": True},
        "movies": movies if movies is not None else {"movie_time": None, "which_frame": 0, "period": "Never"},
        "save": save if save is not None else {"save_time_seconds": 3600.0},
        "manager": {},
        "bins": {"min": st.bins_min, "width": st.bins_width, "histogram": [int(x) for x in b["histogram"]],
                 "t_found": [int(x) for x in b["t_found"]], "lnw": [float(x) for x in b["lnw"]],
                 "energy_total": [float(x) for x in b["energy_total"]],
                 "energy_squared_total": [float(x) for x in b["energy_squared_total"]], "extra": extra},
        "have_visited_since_maxentropy": [bool(x) for x in b["have_visited"]],
        "round_trips": [int(x) for x in b["round_trips"]],
        "max_S": st.max_S, "max_S_index": int(st.max_S_index),
    }


def restore_walker(engine, w, doc):
    """Feed a document back into walker `w` of an engine created with INIT_EXTERNAL (then call engine.resume(moves))."""
    cfg = engine.cfg
    engine.set_system(w, _system_image(cfg, doc["system"], engine.system_len))
    st = _abi.WalkerState()
    st.moves, st.accepted_moves = doc["moves"], doc["accepted_moves"]
    st.acceptance_rate, st.translation_scale = doc["acceptance_rate"], doc["translation_scale"]
    st.rng_s0, st.rng_s1 = doc["rng"]["s0"], doc["rng"]["s1"]
    tag, sys_body = next(iter(doc["system"].items()))
    bins = doc["bins"]
    n = len(bins["lnw"])
    st.bins_min, st.bins_width, st.bins_len = bins["min"], bins["width"], n
    st.max_S, st.max_S_index = doc["max_S"], doc["max_S_index"]
    mtag, m = next(iter(doc["method"].items()))
    wl_hist = np.zeros(n, np.uint64)
    if mtag == "Sad":
        st.method = _abi.METHOD_SAD
        st.too_lo, st.too_hi, st.latest_parameter = m["too_lo"], m["too_hi"], m["latest_parameter"]
        st.tL, st.tF, st.num_states, st.highest_hist = m["tL"], m["tF"], m["num_states"], m["highest_hist"]
    elif mtag == "Samc":
        st.method, st.samc_t0 = _abi.METHOD_SAMC, m["t0"]
    elif mtag == "WL":
        st.method = _abi.METHOD_INV_T_WL if m["inv_t"] else _abi.METHOD_WL
        st.wl_gamma, st.wl_num_states, st.wl_min_energy = m["gamma"], m["num_states"], m["min_energy"]
        st.wl_lowest_hist, st.wl_highest_hist, st.wl_total_hist = m["lowest_hist"], m["highest_hist"], m["total_hist"]
        st.wl_hist_len, st.wl_inv_t = len(m["hist"]), int(m["inv_t"])
        wl_hist[:len(m["hist"])] = m["hist"]
    else:
        st.method = _abi.METHOD_CANONICAL
    # the cached system energy (System::energy) travels inside the system variant
    st.energy = sys_body["E"] if "E" in sys_body else float("nan")
    arrays = {"histogram": bins["histogram"], "t_found": bins["t_found"], "lnw": bins["lnw"], "energy_total": bins["energy_total"],
              "energy_squared_total": bins["energy_squared_total"], "round_trips": doc["round_trips"],
              "have_visited": [1 if x else 0 for x in doc["have_visited_since_maxentropy"]], "wl_hist": wl_hist}
    for label, bc in bins.get("extra", {}).items():
        arrays["extra_total"], arrays["extra_count"] = bc["total"], bc["count"]
    if tag in ("Fake", "FakeErfinv", "TwoWells"):
        # these keep no cached energy: System::energy evaluates the function (fake.rs:96-99) -- on the device, so
        # that the restored value is the one the kernels would compute
        st.energy = engine.compute_energy(w)
    engine.set_walker_bins(w, st, arrays)


# ---- files: yaml / json / cbor by extension, written atomically (src/atomicfile.rs) -------------------------------------

def _cbor_encode(x, out):
    def head(major, n):
        if n < 24:
            out.append(bytes([major << 5 | n]))
        elif n < 1 << 8:
            out.append(bytes([major << 5 | 24, n]))
        elif n < 1 << 16:
            out.append(bytes([major << 5 | 25]) + struct.pack(">H", n))
        elif n < 1 << 32:
            out.append(bytes([major << 5 | 26]) + struct.pack(">I", n))
        else:
            out.append(bytes([major << 5 | 27]) + struct.pack(">Q", n))
    if x is None:
        out.append(b"\xf6")
    elif x is True:
        out.append(b"\xf5")
    elif x is False:
        out.append(b"\xf4")
    elif isinstance(x, (int, np.integer)):
        x = int(x)
        head(0, x) if x >= 0 else head(1, -1 - x)
    elif isinstance(x, (float, np.floating)):
        out.append(b"\xfb" + struct.pack(">d", float(x)))
    elif isinstance(x, str):
        b = x.encode()
        head(3, len(b))
        out.append(b)
    elif isinstance(x, (list, tuple)):
        head(4, len(x))
        for v in x:
            _cbor_encode(v, out)
    elif isinstance(x, dict):
        head(5, len(x))
        for k, v in x.items():
            _cbor_encode(k, out)
            _cbor_encode(v, out)
    else:
        raise TypeError("cannot encode %r" % type(x))


def _cbor_decode(b, i=0):
    ib = b[i]
    major, info = ib >> 5, ib & 31
    i += 1
    if major == 7:
        if info == 20:
            return False, i
        if info == 21:
            return True, i
        if info == 22:
            return None, i
        if info == 27:
            return struct.unpack(">d", b[i:i + 8])[0], i + 8
        if info == 26:
            return struct.unpack(">f", b[i:i + 4])[0], i + 4
        if info == 25:
            return float(np.frombuffer(b[i:i + 2], dtype=">f2")[0]), i + 2
        raise ValueError("unsupported cbor simple value %d" % info)
    if info < 24:
        n = info
    else:
        size = {24: 1, 25: 2, 26: 4, 27: 8}[info]
        n = int.from_bytes(b[i:i + size], "big")
        i += size
    if major == 0:
        return n, i
    if major == 1:
        return -1 - n, i
    if major in (2, 3):
        s = bytes(b[i:i + n])
        return (s.decode() if major == 3 else s), i + n
    if major == 4:
        out = []
        for _ in range(n):
            v, i = _cbor_decode(b, i)
            out.append(v)
        return out, i
    if major == 5:
        out = {}
        for _ in range(n):
            k, i = _cbor_decode(b, i)
            v, i = _cbor_decode(b, i)
            out[k] = v
        return out, i
    raise ValueError("unsupported cbor major type %d" % major)


def dumps(doc, ext):
    if ext == "yaml":
        import yaml
        dumper = getattr(yaml, "CSafeDumper", yaml.SafeDumper)  # libyaml when present: same text, several times faster
        return yaml.dump(doc, Dumper=dumper, default_flow_style=None, sort_keys=False).encode()
    if ext == "json":
        return json.dumps(doc).encode()
    if ext == "cbor":
        out = []
        _cbor_encode(doc, out)
        return b"".join(out)
    raise ValueError("I don't know how to create file with extension %r" % ext)  # mc/mod.rs:118


def loads(data, ext):
    if ext == "yaml":
        import yaml
        return yaml.load(data, Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))
    if ext == "json":
        return json.loads(data)
    if ext == "cbor":
        return _cbor_decode(memoryview(data))[0]
    raise ValueError("I don't know how to read file with extension %r" % ext)  # mc/mod.rs:79,104


def write_atomic(path, data):
    """AtomicFile (src/atomicfile.rs): the file appears under its name only when it is complete."""
    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    fd, tmp = tempfile.mkstemp(dir=d, prefix="." + os.path.basename(path) + ".")
    try:
        with os.fdopen(fd, "wb") as f:
            f.write(data)
        os.replace(tmp, path)
    except BaseException:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise


def stage(path, data):
    """First half of write_atomic: the complete temporary file beside `path`; os.replace(tmp, path) publishes it."""
    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    fd, tmp = tempfile.mkstemp(dir=d, prefix="." + os.path.basename(path) + ".")
    with os.fdopen(fd, "wb") as f:
        f.write(data)
    return tmp


def walker_path(save_as, w, n_walkers):
    """One file per walker: `name.ext` for a single walker (as the reference), `name-w000017.ext` otherwise."""
    if n_walkers == 1:
        return str(save_as)
    stem, ext = os.path.splitext(str(save_as))
    return "%s-w%06d%s" % (stem, w, ext)


def save(engine, save_as, walkers=None, **plugin_docs):
    """MonteCarlo::checkpoint (mc/mod.rs:110-120) for the chosen walkers (default: all).

    The set of per-walker files is written all or nothing: every walker is checked first (a halted walker stops the
    save before any file is touched), every document is written to a temporary file beside its target, and only
    then are the temporaries renamed into place; a partial set (`walkers` shorter than the engine) is marked by
    `name.partial` so that a later resume refuses it with a clear message instead of failing on a missing file."""
    ext = os.path.splitext(str(save_as))[1].lstrip(".")
    partial = walkers is not None and len(list(walkers)) < engine.n_walkers
    walkers = list(range(engine.n_walkers) if walkers is None else walkers)
    left, failed = engine.num_halted()
    if left or failed:
        bad = [w for w in range(engine.n_walkers) if engine.walker(w).status != 0][:8]
        raise RuntimeError("no checkpoint written: %d walker(s) left the bin window and %d failed verify_energy (first: %s)"
                           % (left, failed, bad))
    staged = []
    try:
        for w in walkers:
            p = walker_path(save_as, w, engine.n_walkers)
            staged.append((stage(p, dumps(walker_document(engine, w, save_as=p, **plugin_docs), ext)), p))
        for tmp, p in staged:
            os.replace(tmp, p)
    except BaseException:
        for tmp, _ in staged:
            if os.path.exists(tmp):
                os.unlink(tmp)
        raise
    marker = os.path.splitext(str(save_as))[0] + ".partial"
    if partial:
        write_atomic(marker, ("%d of %d walkers\n" % (len(walkers), engine.n_walkers)).encode())
    elif os.path.exists(marker):
        os.unlink(marker)
    return [p for _, p in staged]


def check_resumable(cfg, save_as, n_walkers):
    """Before any engine is created: the checkpoint set must be complete and must describe the configuration `cfg`.

    Returns the document of walker 0.  Raises ValueError with what is wrong otherwise (missing files, a set written
    with --checkpoint-walkers, a different system / size / method / bin width)."""
    marker = os.path.splitext(str(save_as))[0] + ".partial"
    if os.path.exists(marker):
        raise ValueError("%s holds only %s (written with --checkpoint-walkers): it cannot be resumed"
                         % (save_as, open(marker).read().strip()))
    missing = [w for w in range(n_walkers) if not os.path.exists(walker_path(save_as, w, n_walkers))]
    if missing:
        raise ValueError("checkpoint set %s is incomplete: %d of %d walker files are missing (first: %s)"
                         % (save_as, len(missing), n_walkers, walker_path(save_as, missing[0], n_walkers)))
    doc0 = load(walker_path(save_as, 0, n_walkers))
    if cfg is not None:
        want = config_from_document(doc0, n_walkers=n_walkers)
        for field, what in (("system", "system"), ("N", "system size"), ("energy_bin", "bin width")):
            a, b = getattr(cfg, field), getattr(want, field)
            if field == "energy_bin" and _abi.isnan(a):
                continue  # no --energy-bin given: the system's own (what the document holds)
            if a != b:
                raise ValueError("checkpoint %s was written for another %s (%r, the command line says %r)" % (save_as, what, b, a))
        m_doc, m_cfg = want.method, cfg.method
        same = m_doc == m_cfg or (m_cfg == _abi.METHOD_INV_T_WL and m_doc == _abi.METHOD_SAMC)  # 1/t-WL after its switch
        if not same:
            raise ValueError("checkpoint %s was written by another method (%d, the command line says %d)" % (save_as, m_doc, m_cfg))
    return doc0


def load(path):
    ext = os.path.splitext(str(path))[1].lstrip(".")
    with open(path, "rb") as f:
        return loads(f.read(), ext)


def config_from_document(doc, n_walkers=1, **overrides):
    """The sadmc_config a checkpoint document implies -- what `--resume-from` needs (mc/mod.rs:92-106 deserialises the
    whole EnergyMC; here the engine is re-created from the parameters the document carries and the walker restored)."""
    tag, body = next(iter(doc["system"].items()))
    kw = {}
    if tag == "Lj":
        system, kw = "lj", dict(N=len(body["positions"]), lj_radius=body["max_radius"])
    elif tag == "Ising":
        system, kw = "ising", dict(N=body["N"])
    elif tag == "Fake":
        system, fn = "fake", body["function"]
        if fn == "Linear":
            kw = dict(fake_function=_abi.FAKE_LINEAR, N=1)
        else:
            ftag, f = next(iter(fn.items()))
            if ftag == "Quadratic":
                kw = dict(fake_function=_abi.FAKE_QUADRATIC, N=f["dimensions"])
            elif ftag == "Pieces":
                kw = dict(fake_function=_abi.FAKE_PIECES, N=3, fake_a=f["a"], fake_b=f["b"], fake_e1=f["e1"], fake_e2=f["e2"])
            else:
                kw = dict(fake_function=_abi.FAKE_GAUSSIAN, N=3, fake_sigma=f["sigma"])
    elif tag in ("Wca", "Sw"):
        system = "wca" if tag == "Wca" else "sw"
        b = body["cell"]["box_diagonal"]
        kw = dict(N=len(body["cell"]["positions"]), cell_width=(b["x"], b["y"], b["z"]))
        if tag == "Sw":
            kw["sw_well_width"] = body["cell"]["r_cutoff"]
    elif tag == "TwoWells":
        p = body["parameters"]
        system, kw = "two-wells", dict(N=p["N"], tw_h2_to_h1=p["h2_to_h1"], tw_barrier_over_h1=p["barrier_over_h1"], tw_r2=p["r2"])
    elif tag == "FakeErfinv":
        system, kw = "fake-erfinv", dict(N=len(body["position"]), erfinv_mean_energy=body["parameters"]["mean_energy"])
    else:
        raise NotImplementedError("system variant %r has no device kernel" % tag)
    mtag, m = next(iter(doc["method"].items()))
    if mtag == "Sad":
        method, mk = "sad", dict(sad_min_T=m["min_T"])
    elif mtag == "Samc":
        method, mk = "samc", dict(samc_t0=m["t0"])
    elif mtag == "WL":
        method = "inv-t-wl" if m["inv_t"] else "wl"
        mk = {} if m.get("min_gamma") is None else dict(wl_min_gamma=m["min_gamma"])
    else:
        method, mk = "canonical", dict(canonical_T=m["temperature"])
    kw.update(mk)
    for k in ("min_allowed_energy", "max_allowed_energy"):
        if doc.get(k) is not None:
            kw[k] = doc[k]
    plan, value = next(iter(doc["move_plan"].items()))
    kw["move_plan"] = _abi.MOVE_TRANSLATION_SCALE if plan == "TranslationScale" else _abi.MOVE_ACCEPTANCE_RATE
    kw["move_value"] = value
    bins = doc["bins"]
    kw["energy_bin"] = bins["width"]
    # The device keeps a fixed bin window.  The engine derives it from min/max_allowed_energy or from the system's own
    # bounds (Ising, square well, fake, two wells); where a side has neither (LJ and WCA above, erfinv on both sides)
    # leave room for as many new bins as the document holds.
    n = len(bins["lnw"])
    span = max(n, 64) * bins["width"]
    if system in ("lj", "wca", "fake-erfinv") and kw.get("max_allowed_energy") is None:
        kw["bin_window_hi"] = bins["min"] + n * bins["width"] + span
    if system == "fake-erfinv" and kw.get("min_allowed_energy") is None:
        kw["bin_window_lo"] = bins["min"] - span
    kw.update(overrides)
    return _abi.make_config(system, method, n_walkers=n_walkers, init_mode=_abi.INIT_EXTERNAL, **kw)


def resume(cfg, save_as):
    """`--save-as` on an existing file (mc/mod.rs:70-84): a new engine whose walkers continue the checkpointed ones."""
    import ctypes as C
    from .engine import WalkerEngine
    local = _abi.Config()
    C.memmove(C.byref(local), C.byref(cfg), C.sizeof(cfg))
    local.init_mode = _abi.INIT_EXTERNAL
    check_resumable(None, save_as, local.n_walkers)
    eng = WalkerEngine(local)
    moves = None
    for w in range(eng.n_walkers):
        doc = load(walker_path(save_as, w, eng.n_walkers))
        restore_walker(eng, w, doc)
        if moves is not None and doc["moves"] != moves:
            raise ValueError("walker checkpoints disagree on `moves` (%d vs %d): the set %s was interrupted while it was "
                             "being replaced" % (doc["moves"], moves, save_as))
        moves = doc["moves"]
    eng.resume(moves)
    return eng
