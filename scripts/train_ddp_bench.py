"""Data-parallel training iteration (zigma_b200.train.train_step: flow-matching loss -> forward/backward kernels ->
overlapped bucketed gradient all-reduce -> fused AdamW+EMA) on N GPUs of one node, bs per GPU fixed (weak scaling).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/train_ddp_bench.py
Rank 0 prints one JSON line: ms per iteration (max over ranks, CUDA events), samples/s, and the same with the
exchange not overlapped (all-reduce after the backward) for comparison."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from zigma_b200 import ZigMa, synth, create_transport
from zigma_b200.train import FlatParams, GradSync, FusedAdamWEMA, train_step

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
bs = int(os.environ.get("BS", 16))
steps, warm = int(os.environ.get("STEPS", 10)), 3
CFG = dict(img_dim=32, patch_size=1, in_channels=4, embed_dim=640, depth=18, scan_type="zigzagN8", num_classes=-1,
           has_text=False, use_pe=0, rms_norm=True, fused_add_norm=True, residual_in_fp32=True)
m = ZigMa(device=dev, **CFG)
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict(synth.synth_state_dict(shapes, seed=0))
m.eval()
tr = create_transport()
flat = FlatParams(m)
opt = FusedAdamWEMA(flat, lr=1e-4, weight_decay=0.0)
x1 = torch.stack([synth.synth_latents((4, 32, 32), seed=5000 + rank * bs + i) for i in range(bs)]).to(dev)


def run(sync):
    for _ in range(warm):
        train_step(m, tr, opt, sync, x1, {"y": None}, autocast_dtype=torch.bfloat16)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = train_step(m, tr, opt, sync, x1, {"y": None}, autocast_dtype=torch.bfloat16)
    b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item(), float(loss)


res = {"n_gpus": world, "bs_per_gpu": bs, "steps": steps, "mode": "fp32 master weights, bf16 autocast", "n_params": flat.numel}
sync = GradSync(flat, bucket_mb=float(os.environ.get("BUCKET_MB", 32)), overlap=True)
res["buckets"] = len(sync.buckets)
ms, loss = run(sync)
res.update({"ms_per_iter": ms, "samples_per_s": bs * world / (ms * 1e-3), "loss": loss})
if world > 1:
    sync.remove_hooks()
    ms2, _ = run(GradSync(flat, overlap=False))
    res["ms_per_iter_no_overlap"] = ms2
    # every rank holds the same weights after the same number of averaged steps
    chk = flat.flat.double().sum().reshape(1)
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    res["weights_identical_across_ranks"] = bool(all(torch.equal(lst[0], v) for v in lst))
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
