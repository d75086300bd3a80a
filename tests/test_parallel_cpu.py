"""Host-side multi-GPU logic on CPU: walker sharding and the merged-report all-reduce (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sad_monte_carlo_b200 import make_config
from sad_monte_carlo_b200.parallel import MERGED_KEYS, PACKED_FIELDS, merge_packed, shard, shard_config, sum_shards, unpack_merged


def test_shard_partitions_walkers_exactly():
    for total in (1, 7, 64, 65536, 75776):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard(total, r, world) for r in range(world)]
            assert sum(n for n, _ in blocks) == total
            cursor = 0
            for n, off in blocks:
                assert off == cursor
                cursor += n
            assert max(n for n, _ in blocks) - min(n for n, _ in blocks) <= 1
    with pytest.raises(ValueError):
        shard(10, 4, 4)


def test_shard_config_keeps_global_walker_identity():
    cfg = make_config("ising", N=32, n_walkers=1, seed=100, walker_offset=5)
    seen = []
    for r in range(4):
        c = shard_config(cfg, 10, r, 4, device=r)
        assert c.seed == 100 and c.device == r
        seen += [c.walker_offset + k for k in range(c.n_walkers)]
    # the union over ranks is the same set of global walkers (seed + w) whatever the world size
    assert seen == list(range(5, 15))
    one = shard_config(cfg, 10, 0, 1)
    assert [one.walker_offset + k for k in range(one.n_walkers)] == seen


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _packed_for_rank(rank, nb=50):
    """A synthetic packed fold [7, nb]: histogram halves that need more than 53 bits when recombined, and
    floating-point sums whose result depends on the order of addition."""
    rng = np.random.default_rng(rank)
    hist = rng.integers(0, 2 ** 62, nb, dtype=np.uint64)
    p = np.zeros((PACKED_FIELDS, nb))
    p[0] = (hist >> np.uint64(32)).astype(np.float64)
    p[1] = (hist & np.uint64(0xffffffff)).astype(np.float64)
    p[2] = rng.integers(0, 75776, nb)
    p[3:] = rng.standard_normal((4, nb)) * 10.0 ** rng.integers(-8, 8, (4, nb))
    return p, hist


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p, _ = _packed_for_rank(rank)
    merged = merge_packed(torch.from_numpy(p))
    torch.save({"merged": merged}, os.path.join(out, "r%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_merged_report_one_collective_rank_ordered_gloo(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, "r%d.pt" % r))["merged"] for r in range(world)]
    shards = [_packed_for_rank(r) for r in range(world)]
    # what a single process computes when it adds the same shards one after the other
    want = sum_shards(torch.from_numpy(np.stack([p for p, _ in shards])))
    for r in range(world):
        assert torch.equal(res[r], want)  # bit-identical on every rank, floating-point sums included
    m = unpack_merged(res[0])
    assert set(m) == set(MERGED_KEYS)
    hist = np.zeros(50, dtype=object)
    for _, h in shards:
        hist = hist + h.astype(object)
    assert [int(x) for x in m["histogram"]] == [int(x) % 2 ** 64 for x in hist]  # exact beyond 2^53
    assert np.array_equal(m["lnw_count"], sum(p[2] for p, _ in shards).astype(np.uint64))


def test_merge_without_process_group_is_the_identity():
    p, hist = _packed_for_rank(7)
    t = torch.from_numpy(p)
    assert merge_packed(t) is t
    assert np.array_equal(unpack_merged(t)["histogram"], hist)
