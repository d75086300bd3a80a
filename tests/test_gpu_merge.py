"""GPU checks of the merge for reporting: the one-pass fold (walkers aligned by their running maximum ln w), the
selection of shards and of SAD-interior bins, the packed form for a single collective, and that a run sharded over
several engines (one per rank) merges to exactly what one engine holding all the walkers reports."""
import os
import socket

import numpy as np
import pytest

from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi
from sad_monte_carlo_b200.parallel import MERGED_KEYS, PACKED_FIELDS, shard_config, sum_shards, unpack_merged
from tests.gpu_common import clone_config

pytestmark = pytest.mark.gpu


def _host_fold(eng, walkers, mode=0):
    """The fold restated with numpy from per-walker bins (walker order = the kernel's order within one chunk)."""
    lo, width, n = eng.window()
    out = {"histogram": np.zeros(n, np.uint64), "lnw_count": np.zeros(n, np.uint64), "energy_total": np.zeros(n),
           "energy_squared_total": np.zeros(n), "lnw_sum": np.zeros(n), "lnw_sq_sum": np.zeros(n)}
    for w in walkers:
        s, b = eng.walker(w), eng.bins(w)
        sl = slice(s.window_first, s.window_first + s.bins_len)
        out["histogram"][sl] += b["histogram"]
        out["energy_total"][sl] += b["energy_total"]
        out["energy_squared_total"][sl] += b["energy_squared_total"]
        use = b["histogram"] != 0
        if mode and s.method == _abi.METHOD_SAD:
            E = s.bins_min + (np.arange(s.bins_len) + 0.5) * s.bins_width
            i_lo, i_hi = int(np.abs(E - s.too_lo).argmin()), int(np.abs(E - s.too_hi).argmin())
            inside = np.zeros(s.bins_len, bool)
            inside[i_lo + (mode == 2):i_hi + 1 - (mode == 2)] = True
            use &= inside
        if use.any():
            a = np.where(use, b["lnw"] - b["lnw"][use].max(), 0.0)
            out["lnw_count"][sl] += use.astype(np.uint64)
            out["lnw_sum"][sl] += a
            out["lnw_sq_sum"][sl] += a * a
    return out


@pytest.mark.parametrize("system,method,kw,flags", [
    ("ising", "sad", dict(N=8, sad_min_T=1.0), 0),
    ("ising", "sad", dict(N=8, sad_min_T=1.0), _abi.FLAG_NO_ROUND_TRIPS),  # max_S is still kept: the fold needs it
    ("ising", "wl", dict(N=8), 0),
    ("ising", "samc", dict(N=8, samc_t0=1e3), _abi.FLAG_NO_ROUND_TRIPS),
    ("fake", "sad", dict(fake_function=_abi.FAKE_QUADRATIC, N=3, sad_min_T=0.001, energy_bin=0.01), 0),
])
def test_one_pass_fold_equals_the_fold_restated_on_the_host(system, method, kw, flags):
    cfg = make_config(system, method, n_walkers=40, seed=4, flags=flags, **kw)
    eng = WalkerEngine(cfg)
    eng.run(30000)
    f = eng.fold()  # 40 walkers = one chunk: sums are taken in walker order, so even the f64 sums are bit-equal
    want = _host_fold(eng, range(40))
    for k in MERGED_KEYS:
        assert np.array_equal(f[k], want[k]), k
    # the alignment constant really is the walker's largest ln w
    for w in (0, 17, 39):
        b = eng.bins(w)
        assert eng.walker(w).max_S == b["lnw"][b["histogram"] != 0].max()


@pytest.mark.parametrize("mode", [1, 2])
def test_fold_sad_range_modes(mode):
    cfg = make_config("fake", "sad", fake_function=_abi.FAKE_LINEAR, sad_min_T=0.001, energy_bin=0.01, n_walkers=24, seed=1)
    eng = WalkerEngine(cfg)
    eng.run(200000)
    eng.fold_select(0, 1, mode)
    f = eng.fold()
    eng.fold_select(0, 1, 0)
    want = _host_fold(eng, range(24), mode)
    for k in MERGED_KEYS:
        assert np.array_equal(f[k], want[k]), k
    if mode == 2:  # the end bins of every walker's range are left out
        s = eng.walker(0)
        j_hi = s.window_first + int(round((s.too_hi - s.bins_min) / s.bins_width - 0.5))
        one = _host_fold(eng, [0], 2)
        assert one["lnw_count"][j_hi] == 0 and one["lnw_count"][j_hi - 1] == 1


def test_contiguous_shards_and_packed_fold():
    cfg = make_config("ising", "sad", N=8, sad_min_T=1.0, n_walkers=48, seed=2)
    eng = WalkerEngine(cfg)
    eng.run(20000)
    import torch
    _, _, nb = eng.window()
    parts = []
    for first in (0, 16, 32):
        eng.fold_select(first, 1, 0, walker_count=16)
        f = eng.fold()
        want = _host_fold(eng, range(first, first + 16))
        for k in MERGED_KEYS:
            assert np.array_equal(f[k], want[k]), (first, k)
        t = torch.zeros((PACKED_FIELDS, nb), dtype=torch.float64, device="cuda")
        eng.fold_packed_device(t.data_ptr())
        eng.sync()
        u = unpack_merged(t)
        for k in MERGED_KEYS:
            assert np.array_equal(u[k], f[k]), ("packed", first, k)
        parts.append(t)
    eng.fold_select(0, 1, 0)
    total = unpack_merged(sum_shards(torch.stack(parts)))
    full = eng.fold()
    assert np.array_equal(total["histogram"], full["histogram"]) and np.array_equal(total["lnw_count"], full["lnw_count"])
    assert np.allclose(total["lnw_sum"], full["lnw_sum"], rtol=1e-13, atol=1e-9)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_engines_merge_to_the_single_engine_report(world):
    """N ranks' worth of engines on this GPU (shard_config: the walker blocks and seeds a rank would hold), merged in
    rank order like parallel.merge_packed does after its all-gather, against ONE engine holding all the walkers and
    folding the same contiguous blocks: every array bit for bit, f64 sums included (walker w is `--seed seed + w`
    wherever it lives)."""
    import torch
    total = 256
    base = make_config("lj", "sad", N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01, n_walkers=1,
                       seed=11, init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=1, flags=_abi.FLAG_FAST_MATH,
                       bin_window_lo=-133.62, bin_window_hi=0.02)
    moves = 3000
    packed = []
    for r in range(world):
        eng = WalkerEngine(shard_config(base, total, r, world, device=0))
        eng.run(moves)
        _, _, nb = eng.window()
        t = torch.zeros((PACKED_FIELDS, nb), dtype=torch.float64, device="cuda")
        eng.fold_packed_device(t.data_ptr())
        eng.sync()
        packed.append(t)
        eng.close()
    merged = unpack_merged(sum_shards(torch.stack(packed)))
    one = WalkerEngine(clone_config(base, n_walkers=total))
    one.run(moves)
    per = total // world
    blocks = []
    for r in range(world):
        one.fold_select(r * per, 1, 0, walker_count=per)
        t = torch.zeros((PACKED_FIELDS, nb), dtype=torch.float64, device="cuda")
        one.fold_packed_device(t.data_ptr())
        one.sync()
        blocks.append(t)
    single = unpack_merged(sum_shards(torch.stack(blocks)))
    for k in MERGED_KEYS:
        assert np.array_equal(merged[k], single[k]), k
    assert int(merged["histogram"].sum()) == total * (moves + 1)
    one.fold_select(0, 1, 0)
    whole = one.fold()  # the plain fold of all walkers: integers identical, f64 sums equal up to the order of addition
    assert np.array_equal(whole["histogram"], merged["histogram"]) and np.array_equal(whole["lnw_count"], merged["lnw_count"])
    for k in ("energy_total", "energy_squared_total", "lnw_sum", "lnw_sq_sum"):
        assert np.allclose(whole[k], merged[k], rtol=1e-12, atol=1e-9), k


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_rank(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from sad_monte_carlo_b200.parallel import ShardedEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = make_config("ising", "sad", N=16, sad_min_T=1.0, n_walkers=1, seed=21)
    se = ShardedEngine(cfg, 512, rank=rank, world=world, device=rank)
    se.run(20000)
    m = se.merged()
    np.savez(os.path.join(out, "r%d.npz" % rank), **m)
    se.close()
    dist.destroy_process_group()


def test_nccl_ranks_merge_to_the_single_engine_report(tmp_path):
    """The same statement over real ranks (one process per GPU, NCCL all-gather); needs >= 2 GPUs (gpurun --gpus 2)."""
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_nccl_rank, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [dict(np.load(os.path.join(tmp_path, "r%d.npz" % r))) for r in range(world)]
    cfg = make_config("ising", "sad", N=16, sad_min_T=1.0, n_walkers=512, seed=21)
    one = WalkerEngine(cfg)
    one.run(20000)
    _, _, nb = one.window()
    per = 512 // world
    blocks = []
    for r in range(world):
        one.fold_select(r * per, 1, 0, walker_count=per)
        t = torch.zeros((PACKED_FIELDS, nb), dtype=torch.float64, device="cuda:0")
        one.fold_packed_device(t.data_ptr())
        one.sync()
        blocks.append(t)
    single = unpack_merged(sum_shards(torch.stack(blocks)))
    for r in range(world):
        for k in MERGED_KEYS:
            assert np.array_equal(got[r][k], single[k]), (r, k)


def test_randomize_shim_matches_oracle():
    from tests.oracle_lib import OracleMC
    for system, kw in (("lj", dict(N=13, lj_radius=3.0)), ("fake", dict(fake_function=_abi.FAKE_QUADRATIC, N=3)),
                       ("ising", dict(N=8))):
        if system == "lj":
            kw = dict(kw, lanes_per_walker=1, bin_window_lo=-60.0, bin_window_hi=1e4, energy_bin=1.0)
        cfg = make_config(system, "sad", n_walkers=3, seed=5, **kw)
        eng = WalkerEngine(cfg)
        o = OracleMC(cfg, walker=2)
        for _ in range(3):
            eg, eo = eng.randomize(2), o.randomize()
            assert eg == eo == eng.energy(2)
            assert np.array_equal(eng.system(2), o.system())
            st = o.walker()
            assert tuple(eng.rngs()[2]) == (st.rng_s0, st.rng_s1)


def test_in_loop_verify_energy_runs_at_the_reference_cadence_and_halts_a_corrupted_walker():
    """energy.rs:907-911: verify_energy every len^2 * 1000 moves.  A walker whose cached energy was tampered with
    (set_system with a wrong E) is caught at the first such move: the reference panics, the engine halts the walker with
    SADMC_ERR_VERIFY and says so; the intact walkers carry on and stay equal to the oracle."""
    from sad_monte_carlo_b200.engine import SadmcError
    from tests.oracle_lib import OracleMC
    # square well: integer energies, no periodic recomputation that would heal a wrong cache (optsquare.rs:213-221), and
    # verify_energy is an exact comparison with the slow recount (199-201).  energy_bin 20 keeps len, and with it the
    # period len^2 * 1000, small.
    cfg = make_config("sw", "sad", N=50, filling_fraction=0.3, sw_well_width=1.3, sad_min_T=0.5, energy_bin=20.0, n_walkers=4, seed=3)
    eng = WalkerEngine(cfg)
    img = eng.system(1)
    img[3 * 50] += 1.0  # wrong cached energy
    eng.set_system(1, img)
    with pytest.raises(SadmcError) as ei:
        eng.run(600_000)
    assert ei.value.code == _abi.ERR_VERIFY
    assert eng.walker(1).status == _abi.ERR_VERIFY and eng.num_halted() == (0, 1)
    o = OracleMC(cfg, walker=2)
    o.run(600_000)
    g, s = eng.walker(2), o.walker()
    assert s.bins_len ** 2 * 1000 <= 600_000  # the cadence was reached
    assert g.status == 0 and (g.rng_s0, g.rng_s1, g.energy) == (s.rng_s0, s.rng_s1, s.energy)


def test_fixed_weights_production_run_reweights_to_the_exact_dos():
    """sadmc_set_lnw + Method::Samc with t0 = 0: the weights never change (gamma = 0), every walker samples the same
    multicanonical ensemble, and S = ln w + ln H recovers the exact density of states whatever the weights are --
    here deliberately wrong ones (a tilt of +-1.5 across the range)."""
    from sad_monte_carlo_b200 import analysis
    W = 16384
    cfg = make_config("fake", "samc", fake_function=_abi.FAKE_QUADRATIC, N=3, samc_t0=0.0, energy_bin=0.02, move_value=0.1, n_walkers=W,
                      seed=9, bin_window_lo=-0.04, bin_window_hi=1.04)
    eng = WalkerEngine(cfg)
    lo, width, nb = eng.window()
    E = lo + (np.arange(nb) + 0.5) * width
    exact = analysis.fake_bin_weights("quadratic", lo, width, nb, 3)
    w = np.where(exact > 0, np.log(np.where(exact > 0, exact, 1.0)), 0.0) + 3.0 * (E - 0.5)  # exact ln D, tilted
    eng.set_lnw(w)
    eng.run(20000)   # equilibration from the common start at the origin
    h0 = eng.fold()["histogram"].astype(np.float64)
    eng.run(200000)
    H = eng.fold()["histogram"].astype(np.float64) - h0
    for wk in (0, W - 1):  # the weights are still exactly what was set
        s, b = eng.walker(wk), eng.bins(wk)
        sl = slice(s.window_first, s.window_first + s.bins_len)
        assert np.array_equal(b["lnw"], w[sl])
    ok = (exact > 0) & (H > 0)
    assert ok.sum() >= 50
    d = w[ok] + np.log(H[ok]) - np.log(exact[ok])
    d -= d.mean()
    assert np.sqrt(np.mean(d * d)) < 5e-3 and np.abs(d).max() < 2e-2, (np.sqrt(np.mean(d * d)), np.abs(d).max())
