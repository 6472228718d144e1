"""Data-parallel sharding of the sampling batch: samples are independent and
the model is replicated, so the batch is split by rank, NO collective runs inside a step, and the
generated latents are all-gathered once at the end -- what the reference does through
``accelerator.gather`` (sample_acc.py:435-436, train_acc.py:573)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n samples owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def global_latents(n, shape, seed=0, dtype=torch.float32):
    """Initial noise indexed by the GLOBAL sample id, so results do not depend on the world size."""
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.randn(n, *shape).astype(np.float32)).to(dtype)


def gather_latents(local, n, world):
    """One all-gather of the final latents; handles uneven shards by padding to the largest."""
    if world == 1 or not dist.is_initialized():
        return local
    sizes = [shard_range(n, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)
