"""Host-side multi-GPU logic on CPU: walker sharding and the merged-report all-reduce (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sad_monte_carlo_b200 import make_config
from sad_monte_carlo_b200.parallel import MERGED_KEYS, all_reduce_merged, shard, shard_config


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


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nb = 50
    rng = np.random.default_rng(rank)
    t = {k: torch.from_numpy(rng.integers(0, 100, nb)).to(torch.int64 if k in ("histogram", "lnw_count") else torch.float64)
         for k in MERGED_KEYS}
    local = {k: v.clone() for k, v in t.items()}
    all_reduce_merged(t)
    torch.save({"local": local, "merged": t}, os.path.join(out, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_merged_report_all_reduce_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, "r%d.pt" % r)) for r in range(world)]
    for k in MERGED_KEYS:
        want = res[0]["local"][k] + res[1]["local"][k]
        for r in range(world):
            assert torch.equal(res[r]["merged"][k], want), k
