"""Walkers sharded over GPUs: one process per GPU, no traffic while moving, one all-reduce when reporting.

Every walker is a complete `EnergyMC` (reference src/mc/energy.rs:167-210: own system, RNG, bins, method state), so
the hot path shards with no data-path collective: rank r of W holds the contiguous block of global walkers
[offset, offset + n_local) and walker w is seeded `seed + w` -- results do not depend on the GPU count.
Only the merged report (histogram, energy moments, aligned ln w sums) crosses NVLink: a device fold
(sadmc_fold_device) straight into torch tensors followed by `torch.distributed.all_reduce` (NCCL on GPUs, gloo in
the CPU tests), at reporting intervals.
"""
import ctypes as C

import numpy as np

MERGED_KEYS = ("histogram", "energy_total", "energy_squared_total", "lnw_sum", "lnw_sq_sum", "lnw_count")


def shard(n_walkers_total, rank, world):
    """(n_local, walker_offset) of `rank`: contiguous blocks, the first `total % world` ranks hold one extra."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(int(n_walkers_total), int(world))
    n_local = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return n_local, offset


def shard_config(cfg, n_walkers_total, rank, world, device=None):
    """This rank's copy of `cfg`: its walker block and its device."""
    from ._abi import Config
    n_local, offset = shard(n_walkers_total, rank, world)
    local = Config()
    C.memmove(C.byref(local), C.byref(cfg), C.sizeof(cfg))
    local.n_walkers = n_local
    local.walker_offset = cfg.walker_offset + offset
    if device is not None:
        local.device = device
    return local


def all_reduce_merged(tensors, group=None):
    """Sum the per-rank fold tensors in place over the process group (NCCL all-reduce over NVLink on GPUs)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for k in MERGED_KEYS:
            dist.all_reduce(tensors[k], op=dist.ReduceOp.SUM, group=group)
    return tensors


class ShardedEngine:
    """A WalkerEngine for this rank's shard + the merged report across ranks."""

    def __init__(self, cfg, n_walkers_total, rank=None, world=None, device=None, group=None):
        import torch
        import torch.distributed as dist
        from . import WalkerEngine
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        local = shard_config(cfg, n_walkers_total, rank, world, device)
        self.rank, self.world, self.group = rank, world, group
        self.n_local, self.offset = local.n_walkers, local.walker_offset - cfg.walker_offset
        self.engine = WalkerEngine(local)
        self.device = torch.device("cuda", local.device)
        _, _, nb = self.engine.window()
        self._t = {k: torch.zeros(nb, dtype=torch.int64 if k in ("histogram", "lnw_count") else torch.float64,
                                  device=self.device) for k in MERGED_KEYS}

    def run(self, n_moves):
        self.engine.run(n_moves)

    def merged(self):
        """Fold the local walkers on the device, all-reduce across ranks; host arrays, identical on all ranks."""
        import torch
        t = self._t
        self.engine.fold_device(*[t[k].data_ptr() for k in MERGED_KEYS])
        self.engine.sync()
        torch.cuda.synchronize(self.device)
        all_reduce_merged(t, self.group)
        out = {k: t[k].cpu().numpy() for k in MERGED_KEYS}
        out["histogram"] = out["histogram"].astype(np.uint64)
        out["lnw_count"] = out["lnw_count"].astype(np.uint64)
        return out

    def close(self):
        self.engine.close()
