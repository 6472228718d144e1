"""world_size-2 gloo worker for tests/test_host_logic.py::test_shard_sampling_world2_gloo."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zigma_b200.sharding import shard_range, gather_latents, global_latents  # noqa: E402
from zigma_b200 import create_transport, Sampler  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
B = 6
W = torch.linspace(-1, 1, 16).reshape(4, 4)
model = lambda x, t, **kw: torch.tanh(x.flatten(1)[:, :4] @ W).repeat(1, x[0].numel() // 4).reshape(x.shape) * (1 + t.view(-1, 1, 1, 1))
fn = Sampler(create_transport()).sample_ode(sampling_method="euler", num_steps=8)
z_all = global_latents(B, (4, 2, 2), seed=0)          # depends only on the GLOBAL sample index
lo, hi = shard_range(B, rank, world)
mine = fn(z_all[lo:hi], model)[-1]
full = gather_latents(mine, B, world)
if rank == 0:
    ref = fn(z_all, model)[-1]
    assert full.shape == ref.shape and torch.allclose(full, ref, atol=1e-6), (full - ref).abs().max()
    print("DIST_OK")
dist.destroy_process_group()
