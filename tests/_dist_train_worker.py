"""world_size-2 gloo worker for tests/test_train.py: the bucketed, hook-driven all-reduce of GradSync turns per-rank
gradients into the full-batch gradient (no optimiser here: that kernel is CUDA only)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zigma_b200.train import FlatParams, GradSync  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(100 + rank)                          # DIFFERENT initial weights per rank, like set_seed(device_specific=True) (train_acc.py:125)
net = torch.nn.Sequential(torch.nn.Linear(7, 13), torch.nn.Tanh(), torch.nn.Linear(13, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
gen = torch.Generator().manual_seed(0)                 # the same data on every rank
X, Y = torch.randn(8, 7, generator=gen), torch.randn(8, 3, generator=gen)
ref = [p.detach().clone() for p in net.parameters()]
flat = FlatParams(net)
assert all(torch.equal(a, b) for a, b in zip(ref, net.parameters()))
sync = GradSync(flat, bucket_mb=1e-4)                  # tiny buckets: several all-reduces, launched from the hooks
# GradSync starts every replica from rank 0's weights (what DDP / accelerate do at wrap time)
w0 = [torch.empty_like(flat.flat) for _ in range(world)]
dist.all_gather(w0, flat.flat)
assert all(torch.equal(w0[0], w) for w in w0), "replicas must start from rank 0's weights"
assert (rank == 0) == all(torch.equal(a, b) for a, b in zip(ref, net.parameters()))
assert len(sync.buckets) >= 3, sync.buckets
for it in range(2):                                    # two steps: counters re-arm, zero_grad keeps the views
    flat.zero_grad()
    sync.begin()
    lo, hi = rank * 4, rank * 4 + 4
    loss = ((net(X[lo:hi]) - Y[lo:hi]) ** 2).mean()
    loss.backward()
    launched_in_backward = sum(sync._launched)
    scale = sync.finish()
    assert launched_in_backward == len(sync.buckets), (launched_in_backward, len(sync.buckets))
    g = flat.grad * scale
    # single-process full-batch gradient
    net2 = torch.nn.Sequential(torch.nn.Linear(7, 13), torch.nn.Tanh(), torch.nn.Linear(13, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    net2.load_state_dict({k: v.clone() for k, v in net.state_dict().items()})
    ((net2(X) - Y) ** 2).mean().backward()
    for i, ((n, p), p2) in enumerate(zip(flat.named, net2.parameters())):
        assert p.grad.data_ptr() == flat.grad.data_ptr() + 4 * flat.offsets[i]
        assert torch.allclose(flat.view_of(g, i), p2.grad, rtol=1e-5, atol=1e-7), n
if rank == 0:
    print("TRAIN_DIST_OK")
dist.destroy_process_group()
