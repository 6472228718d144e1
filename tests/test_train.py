"""Training step around the path: flat parameter / gradient buffers, overlapped
data-parallel gradient mean (CPU, gloo), fused AdamW + EMA kernel and the whole step (GPU)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _net():
    torch.manual_seed(1)
    return torch.nn.Sequential(torch.nn.Linear(6, 10), torch.nn.SiLU(), torch.nn.Linear(10, 3, bias=False))


def test_flat_params_are_views_and_grads_accumulate_in_place():
    from zigma_b200.train import FlatParams
    net = _net()
    net[0].bias.requires_grad_(False)                       # non-trainable parameters stay where they are
    before = {k: v.clone() for k, v in net.state_dict().items()}
    flat = FlatParams(net)
    assert [n for n, _ in flat.named] == ["0.weight", "2.weight"]
    assert flat.numel % 4 == 0 and all(o % 4 == 0 for o in flat.offsets)
    assert all(torch.equal(before[k], v) for k, v in net.state_dict().items())
    x = torch.randn(5, 6)
    net(x).sum().backward()
    g1 = flat.grad.clone()
    assert g1.abs().sum() > 0 and net[0].weight.grad.data_ptr() == flat.grad.data_ptr()
    net(x).sum().backward()                                 # accumulates INTO the flat buffer
    assert torch.allclose(flat.grad, 2 * g1)
    flat.zero_grad()
    assert flat.grad.abs().sum() == 0 and net[2].weight.grad.data_ptr() == flat.grad.data_ptr() + 4 * flat.offsets[1]
    net[2].weight.grad = None                               # e.g. a stray optimizer.zero_grad(set_to_none=True)
    flat.zero_grad()
    assert net[2].weight.grad is not None
    with torch.no_grad():                                   # writing the flat buffer IS writing the parameters
        flat.flat.mul_(0)
    assert all(v.abs().sum() == 0 for k, v in net.state_dict().items() if k != "0.bias")
    with pytest.raises(ValueError):
        FlatParams(torch.nn.Linear(2, 2).half())


def test_grad_sync_world2_gloo():
    script = os.path.join(ROOT, "tests", "_dist_train_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29547", script],
                       env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "TRAIN_DIST_OK" in r.stdout


def test_fused_optimizer_needs_cuda():
    from zigma_b200.train import FlatParams, FusedAdamWEMA
    with pytest.raises(RuntimeError):
        FusedAdamWEMA(FlatParams(_net()))


@pytest.mark.gpu
def test_fused_adamw_ema_matches_torch():
    """zg_adamw_ema_step vs torch.optim.AdamW + the reference's update_ema loop, 5 steps, odd sizes, with and without
    weight decay / gradient scale."""
    from zigma_b200.train import FlatParams, FusedAdamWEMA, reference_update_ema_
    for wd, scale in ((0.0, 1.0), (0.05, 0.5)):
        net = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.SiLU(), torch.nn.Linear(53, 11)).cuda()
        ref = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.SiLU(), torch.nn.Linear(53, 11)).cuda()
        ref.load_state_dict(net.state_dict())
        ema_ref = [p.detach().clone() for p in ref.parameters()]
        flat = FlatParams(net)
        opt = FusedAdamWEMA(flat, lr=3e-3, weight_decay=wd, ema_decay=0.9)
        topt = torch.optim.AdamW(ref.parameters(), lr=3e-3, weight_decay=wd)
        gen = torch.Generator(device="cuda").manual_seed(0)
        for it in range(5):
            x = torch.randn(16, 37, device="cuda", generator=gen)
            opt.zero_grad(); topt.zero_grad()
            (net(x) ** 2).mean().backward()
            ((ref(x) ** 2).mean() * scale).backward()            # the fused step scales the gradient itself
            opt.step(grad_scale=scale)
            topt.step()
            reference_update_ema_(ema_ref, list(ref.parameters()), 0.9)
        for (n, p), q in zip(net.named_parameters(), ref.parameters()):
            assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), (n, (p - q).abs().max())
        for (n, e), q in zip(opt.ema_state_dict().items(), ema_ref):
            assert torch.allclose(e, q, rtol=2e-5, atol=2e-6), n
        em = opt.ema_module()
        assert all(not p.requires_grad for p in em.parameters())
        assert torch.allclose(em[0].weight, ema_ref[0], rtol=2e-5, atol=2e-6)
        # device-side clip coefficient == torch's clip_grad_norm_
        coef, norm = opt.clip_coefficient(0.01)
        want = torch.nn.utils.clip_grad_norm_(net.parameters(), 1e9)
        assert torch.allclose(norm, want) and torch.allclose(coef, torch.clamp(0.01 / (want + 1e-6), max=1.0))
        v0, w0 = em[0].weight._version, net[0].weight._version
        opt.step()                                           # the raw-pointer update must bump autograd's version counters
        assert em[0].weight._version > v0 and net[0].weight._version > w0


@pytest.mark.gpu
def test_train_step_on_tiny_zigma_matches_unfused_reference_loop():
    """zigma_b200.train.train_step (flow-matching loss -> our forward/backward kernels -> fused AdamW+EMA) against the
    reference's loop written out with torch.optim.AdamW and update_ema on a twin model: same losses and weights after
    3 steps (fp32), and the loss goes down."""
    from zigma_b200 import ZigMa, create_transport
    from zigma_b200.train import FlatParams, FusedAdamWEMA, train_step, reference_update_ema_
    cfg = dict(img_dim=8, patch_size=1, in_channels=4, embed_dim=64, depth=2, scan_type="zigzagN8", num_classes=-1, has_text=False,
               use_pe=0, rms_norm=True, fused_add_norm=True, residual_in_fp32=True)
    torch.manual_seed(0)
    m = ZigMa(device="cuda", **cfg)
    twin = ZigMa(device="cuda", **cfg)
    with torch.no_grad():                                    # adaLN-zero init would make every gradient but the head's vanish
        for p in m.parameters():
            if p.requires_grad and p.abs().sum() == 0:
                p.normal_(0, 0.02)
    twin.load_state_dict(m.state_dict())
    m.eval(); twin.eval()                                    # drop_path off; gradients flow
    tr = create_transport()
    flat = FlatParams(m)
    opt = FusedAdamWEMA(flat, lr=2e-3, weight_decay=0.0, ema_decay=0.99)
    topt = torch.optim.AdamW([p for p in twin.parameters() if p.requires_grad], lr=2e-3, weight_decay=0.0)
    ema_ref = [p.detach().clone() for p in twin.parameters() if p.requires_grad]
    x1 = torch.randn(6, 4, 8, 8, device="cuda")
    losses = []
    for it in range(3):
        torch.manual_seed(10 + it)
        loss = train_step(m, tr, opt, None, x1, {"y": None})
        torch.manual_seed(10 + it)                            # same (t, x0) draw
        lr_ = tr.training_losses(twin, x1, {"y": None})["loss"].mean()
        topt.zero_grad(); lr_.backward(); topt.step()
        reference_update_ema_(ema_ref, [p for p in twin.parameters() if p.requires_grad], 0.99)
        assert torch.allclose(loss, lr_.detach(), rtol=1e-4, atol=1e-6), (it, loss.item(), lr_.item())
        losses.append(loss.item())
    for (n, p), q in zip(m.named_parameters(), twin.parameters()):
        if p.requires_grad:
            assert torch.allclose(p, q, rtol=1e-3, atol=2e-5), (n, (p - q).abs().max())
    for e, q in zip(opt.ema_state_dict().values(), ema_ref):
        assert torch.allclose(e, q, rtol=1e-3, atol=2e-5)
    fixed = []
    for it in range(40):
        torch.manual_seed(3)                                  # a fixed (t, x0): the loss must go down
        fixed.append(train_step(m, tr, opt, None, x1, {"y": None}).item())
    assert fixed[-1] < 0.7 * fixed[0], (fixed[0], fixed[-1])


def test_flat_params_and_ema_module_drop_a_stale_engine():
    """Host logic: a sampling engine cached on the module holds views of the old parameter storage (and CUDA graphs that
    cannot be deep-copied) -- FlatParams drops it, ema_module() copies around it."""
    from zigma_b200.train import FlatParams, FusedAdamWEMA
    net = _net()
    net._engine = object()
    flat = FlatParams(net)
    assert net._engine is None

    class NoCopy:
        def __deepcopy__(self, memo):
            raise RuntimeError("engines must not be deep-copied")
    net._engine = NoCopy()
    opt = FusedAdamWEMA.__new__(FusedAdamWEMA)            # host-side part only (the step kernel is CUDA)
    opt.flat, opt.ema, opt._versioned = flat, flat.flat.clone() + 1.0, []
    em = opt.ema_module()
    assert isinstance(net._engine, NoCopy) and getattr(em, "_engine", None) is None
    assert torch.equal(em[0].weight, net[0].weight + 1.0) and not em[0].weight.requires_grad


def test_optimizer_state_dict_round_trip_host_logic():
    """FusedAdamWEMA.state_dict() has torch.optim.AdamW's layout (loads into a real AdamW) and load_state_dict restores the
    moments, the step count and the EMA (host-side logic only: no kernel launch)."""
    from zigma_b200.train import FlatParams, FusedAdamWEMA
    net = _net()
    flat = FlatParams(net)
    opt = FusedAdamWEMA.__new__(FusedAdamWEMA)
    opt.flat, opt.lr, opt.weight_decay, opt.betas, opt.eps, opt.ema_decay = flat, 3e-4, 0.01, (0.9, 0.999), 1e-8, 0.999
    opt.exp_avg, opt.exp_avg_sq, opt.ema, opt.steps = torch.randn_like(flat.flat), torch.rand_like(flat.flat), flat.flat.clone() + 0.5, 17
    sd = opt.state_dict()
    ref_opt = torch.optim.AdamW([p for _, p in flat.named], lr=1.0)
    ref_opt.load_state_dict({k: v for k, v in sd.items() if k != "ema_flat"})       # AdamW accepts the layout
    st0 = ref_opt.state[flat.named[0][1]]
    assert float(st0["step"]) == 17 and torch.equal(st0["exp_avg"], flat.view_of(opt.exp_avg, 0))
    assert ref_opt.param_groups[0]["lr"] == 3e-4 and ref_opt.param_groups[0]["weight_decay"] == 0.01
    opt2 = FusedAdamWEMA.__new__(FusedAdamWEMA)
    opt2.flat, opt2.lr, opt2.weight_decay, opt2.betas, opt2.eps, opt2.ema_decay = flat, 0.0, 0.0, (0.5, 0.5), 1.0, 0.999
    opt2.exp_avg, opt2.exp_avg_sq, opt2.ema, opt2.steps = torch.zeros_like(flat.flat), torch.zeros_like(flat.flat), torch.zeros_like(flat.flat), 0
    opt2.load_state_dict(sd)
    assert opt2.steps == 17 and opt2.lr == 3e-4 and opt2.betas == (0.9, 0.999)
    for i in range(len(flat.named)):      # (the flat buffers have alignment padding between parameters: compare the views)
        assert torch.equal(flat.view_of(opt2.exp_avg, i), flat.view_of(opt.exp_avg, i))
        assert torch.equal(flat.view_of(opt2.exp_avg_sq, i), flat.view_of(opt.exp_avg_sq, i))
    assert torch.equal(opt2.ema, opt.ema)
    # a real AdamW state loads too (resuming a reference 'opt' entry)
    real = torch.optim.AdamW([p for _, p in flat.named], lr=2e-4)
    for _, p in flat.named:
        p.grad = torch.ones_like(p)
    real.step()
    opt2.load_state_dict(real.state_dict())
    assert opt2.steps == 1 and opt2.lr == 2e-4


@pytest.mark.gpu
def test_autocast_with_fp32_pos_embed_train_and_sample():
    """bf16 autocast over fp32 master weights with use_pe=2 -- embed() returns fp32 tokens (bf16 linear +
    fp32 pos_embed) while the adaLN chunks and the mixer output are bf16.  The block-tail kernels must not reinterpret the
    mixed buffers: train step (fused tail, drop_path 0) and eval sampling both agree with the fp32 run to bf16 accuracy."""
    from zigma_b200 import ZigMa
    cfg = dict(img_dim=8, patch_size=1, in_channels=4, embed_dim=64, depth=3, scan_type="zigzagN8", use_pe=2, drop_path_rate=0.0)
    torch.manual_seed(0)
    m = ZigMa(device="cuda", **cfg)
    with torch.no_grad():
        for p in m.parameters():
            if p.requires_grad and p.abs().sum() == 0:
                p.normal_(0, 0.02)
    x = torch.randn(3, 4, 8, 8, device="cuda")
    t = torch.rand(3, device="cuda")
    m.train()
    ref = m(x, t)
    ref.square().mean().backward()
    g_ref = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(x, t)
        loss = out.float().square().mean()
    loss.backward()
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max() <= 5e-2 * max(1.0, ref.abs().max().item())
    rels = []
    for n, p in m.named_parameters():
        if p.grad is not None and g_ref[n].abs().max() > 0:
            rel = ((p.grad - g_ref[n]).norm() / g_ref[n].norm()).item()
            assert torch.isfinite(p.grad).all() and rel < 0.4, (n, rel)     # (bf16 noise; garbage reads give rel >> 1 or NaN)
            rels.append(rel)
    assert sorted(rels)[len(rels) // 2] < 0.05, sorted(rels)
    m.eval()
    with torch.no_grad():
        want = m(x, t)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            got = m(x, t)
    assert torch.isfinite(got.float()).all()
    assert (got.float() - want).abs().max() <= 5e-2 * max(1.0, want.abs().max().item())


@pytest.mark.gpu
def test_engine_not_used_for_layernorm_models():
    """rms_norm=False builds nn.LayerNorm blocks; the RMSNorm-only sampling engine must not run them."""
    from zigma_b200 import ZigMa
    cfg = dict(img_dim=8, patch_size=1, in_channels=4, embed_dim=64, depth=2, scan_type="zigzagN8", use_pe=0, rms_norm=False)
    torch.manual_seed(0)
    m = ZigMa(device="cuda", **cfg).eval()
    with torch.no_grad():
        for p in m.parameters():
            if p.requires_grad and p.abs().sum() == 0:
                p.normal_(0, 0.02)
        x, t = torch.randn(2, 4, 8, 8, device="cuda"), torch.rand(2, device="cuda")
        got = m(x, t)
        want = m.forward_autograd(x, t)
    assert m._engine is None
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
