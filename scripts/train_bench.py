"""One flow-matching TRAINING step (forward + backward, MSE to a target velocity) of the BASELINE
config-2 denoiser: ours (ZigMa.forward_autograd -> our forward/backward kernels) vs the reference's
CUDA forward/backward kernels (oracle/_ref) inside the restated reference glue under autograd.
    BS=16 DTYPE=bf16|fp32 PROFILE=1 python scripts/train_bench.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zigma_b200 import ZigMa, synth, rms_norm_fn
from oracle import ref_cuda, zigma_oracle as zo

dev = "cuda"
bs = int(os.environ.get("BS", 16))
MODE = os.environ.get("DTYPE", "bf16")          # bf16 (weights + activations), fp32 (the reference's default), amp (fp32 weights, bf16 autocast)
dtype = torch.bfloat16 if MODE == "bf16" else torch.float32
amp = lambda: torch.autocast("cuda", dtype=torch.bfloat16, enabled=(MODE == "amp"))
CFG = dict(img_dim=32, patch_size=1, in_channels=4, embed_dim=640, depth=18, scan_type="zigzagN8", num_classes=-1,
           has_text=False, use_pe=0, rms_norm=True, fused_add_norm=True, residual_in_fp32=True)
if os.environ.get("DEPTH"):
    CFG["depth"] = int(os.environ["DEPTH"])
m = ZigMa(device=dev, dtype=dtype, **CFG)
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
sd = synth.synth_state_dict(shapes, seed=0, dtype=dtype)
m.load_state_dict(sd)
m.eval()                                     # drop_path off; gradients still flow
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(bs, 4, 32, 32, device=dev, generator=g).to(dtype)
t = torch.rand(bs, device=dev, generator=g).to(dtype)
target = torch.randn(bs, 4, 32, 32, device=dev, generator=g).to(dtype)


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def ours_step():
    for p_ in m.parameters():
        p_.grad = None
    with amp():
        out = m.forward_autograd(x, t, None)
    loss = ((out.float() - target.float()) ** 2).mean()
    loss.backward()
    return loss


def ours_fwd():
    with torch.no_grad(), amp():
        return m.forward_autograd(x, t, None)


res = {"bs": bs, "mode": MODE, "depth": CFG["depth"], "token_major": os.environ.get("ZIGMA_TOKEN_MAJOR_TRAIN", "1")}
res["ours_train_step_ms"] = timeit(ours_step)
res["ours_fwd_only_autograd_path_ms"] = timeit(ours_fwd)
torch.cuda.reset_peak_memory_stats()
ours_step(); torch.cuda.synchronize()
res["ours_peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30

# ---- optimiser + EMA part of the iteration (fp32 master weights only): fused kernel vs torch AdamW + the reference's EMA loop
if MODE != "bf16":
    from zigma_b200.train import FlatParams, FusedAdamWEMA, reference_update_ema_
    import copy
    twin = copy.deepcopy(m)
    flat = FlatParams(m)
    fopt = FusedAdamWEMA(flat, lr=1e-4, weight_decay=0.0)
    ours_step()
    res["ours_adamw_ema_ms"] = timeit(lambda: fopt.step(), n=10)
    tp = [p_ for p_ in twin.parameters() if p_.requires_grad]
    for p_ in tp:
        p_.grad = torch.randn_like(p_)
    topt = torch.optim.AdamW(tp, lr=1e-4, weight_decay=0.0)
    ema_ref = [p_.detach().clone() for p_ in tp]
    def torch_opt():
        topt.step(); reference_update_ema_(ema_ref, tp)
    res["torch_adamw_plus_ema_loop_ms"] = timeit(torch_opt, n=10)
    res["n_params"] = flat.numel
    del twin, topt, ema_ref

if ref_cuda.available():
    sdr = {k: v.to(dev).clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    norm = lambda x_, w, b, residual=None, prenorm=False, residual_in_fp32=False, eps=1e-6: rms_norm_fn(
        x_, w, b, residual=residual, prenorm=prenorm, residual_in_fp32=residual_in_fp32, eps=eps)
    cfg = dict(CFG, norm_epsilon=1e-5)

    def ref_step():
        for v in sdr.values():
            v.grad = None
        zo.BACKEND = ref_cuda.train_backend(norm)
        try:
            with amp():
                out = zo.zigma_forward(sdr, cfg, x, t)
            loss = ((out.float() - target.float()) ** 2).mean()
            loss.backward()
        finally:
            zo.BACKEND = {}
        return loss
    res["reference_kernels_train_step_ms"] = timeit(ref_step)
    torch.cuda.reset_peak_memory_stats()
    ref_step(); torch.cuda.synchronize()
    res["reference_peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    res["speedup"] = res["reference_kernels_train_step_ms"] / res["ours_train_step_ms"]
    # gradient agreement on a few parameters
    ours_step(); ref_step()
    named = dict(m.named_parameters())
    for k in ("blocks.0.mixer.in_proj.weight", "blocks.17.mixer.A_log", "blocks.5.mixer.conv1d.weight", "x_embedder.proj.weight"):
        if k in named and named[k].grad is not None and sdr[k].grad is not None:
            a_, b_ = named[k].grad.float(), sdr[k].grad.float().reshape(named[k].grad.shape)
            res["grad_rel_" + k] = float((a_ - b_).norm() / (b_.norm() + 1e-30))
print(json.dumps(res))

if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        ours_step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
