"""Walkers sharded over GPUs: one process per GPU, no traffic while moving, one all-reduce when reporting.

Every walker is a complete `EnergyMC` (reference src/mc/energy.rs:167-210: own system, RNG, bins, method state), so
the hot path shards with no data-path collective: rank r of W holds the contiguous block of global walkers
[offset, offset + n_local) and walker w is seeded `seed + w` -- results do not depend on the GPU count.
Only the merged report (histogram, energy moments, aligned ln w sums) crosses NVLink: a device fold
(sadmc_fold_packed_device) straight into one torch tensor followed by ONE `all_gather_into_tensor` (NCCL on GPUs, gloo
in the CPU tests) and a rank-ordered sum, at reporting intervals.
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


PACKED_FIELDS = 7  # sadmc_fold_packed_device: histogram >> 32, histogram & 0xffffffff, lnw_count, then the four f64 sums


def merge_packed(packed, group=None):
    """ONE collective for the whole report: every rank's packed fold [7, nbins] (f64) is all-gathered and the shards
    are added in RANK ORDER on every rank.

    The integer fields travel as exact doubles (32-bit halves of the histogram, walker counts), so their sums are
    exact; the floating-point sums are added shard by shard in a fixed order, so the merged report does not depend on
    NCCL's reduction tree and is bit-identical to folding the same shards one after the other on a single GPU
    (tests/test_gpu_merge.py).  The payload is 56 bytes per bin and rank (LJ31: 0.75 MB per rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return packed
    world = dist.get_world_size(group)
    # concatenated along dim 0 (the form every backend accepts), then viewed as [world, 7, nbins]
    gathered = torch.empty((world * packed.shape[0],) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(gathered, packed.contiguous(), group=group)
    return sum_shards(gathered.view((world,) + tuple(packed.shape)))


def sum_shards(gathered):
    """Shard-ordered sum of [n_shards, 7, nbins]."""
    acc = gathered[0].clone()
    for r in range(1, gathered.shape[0]):
        acc += gathered[r]
    return acc


def unpack_merged(packed):
    """Packed [7, nbins] f64 tensor (or array) -> the host arrays of MERGED_KEYS."""
    p = packed.detach().cpu().numpy() if hasattr(packed, "detach") else np.asarray(packed)
    hi = p[0].astype(np.uint64)
    lo = p[1].astype(np.uint64)
    return {"histogram": (hi << np.uint64(32)) + lo, "lnw_count": p[2].astype(np.uint64), "energy_total": p[3].copy(),
            "energy_squared_total": p[4].copy(), "lnw_sum": p[5].copy(), "lnw_sq_sum": p[6].copy()}


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
        self._packed = torch.zeros((PACKED_FIELDS, nb), dtype=torch.float64, device=self.device)

    def run(self, n_moves):
        self.engine.run(n_moves)

    def merged_device(self):
        """Fold the local walkers on the device (one pass over the bin records) and merge the ranks with one
        all-gather; returns the packed [7, nbins] tensor, identical on all ranks."""
        import torch
        self.engine.fold_packed_device(self._packed.data_ptr())
        self.engine.sync()
        torch.cuda.synchronize(self.device)
        return merge_packed(self._packed, self.group)

    def merged(self):
        """The merged report as host arrays (MERGED_KEYS), identical on all ranks."""
        return unpack_merged(self.merged_device())

    def close(self):
        self.engine.close()
