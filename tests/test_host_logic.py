"""CPU: host-side logic of the product package that needs no kernel -- module construction and
state-dict parity with the reference, scan tables per layer, the sampler driver, the sharding used
by bench.py (world_size-2 gloo)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import synth, zigma_oracle as zo
from util import ROOT, model_case

CASES = ["tiny_zigzag8", "tiny_sweep2", "tiny_hilbert2", "tiny_patch2_cls", "tiny_video_sst", "full_zigzag8_b1"]


def build(cfg, device="cpu", dtype=torch.float32):
    from zigma_b200 import ZigMa
    return ZigMa(device=device, dtype=dtype, **cfg).eval()


@pytest.mark.parametrize("name", CASES)
def test_state_dict_layout_matches_reference(name):
    g, cfg, shapes = model_case(name)
    m = build(cfg)
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert mine == shapes
    m.load_state_dict(synth.synth_state_dict(shapes), strict=True)


def test_layer_tables_follow_reference_rules():
    g, cfg, _ = model_case("tiny_zigzag8")
    m = build(cfg)
    fwd, rev, _ = zo.build_scan_tables(cfg["scan_type"], cfg["depth"], (cfg["img_dim"] // cfg["patch_size"]) ** 2)
    for i, blk in enumerate(m.blocks):
        assert np.array_equal(blk.mixer.zigzag_paths[i].numpy(), fwd[i])
        assert np.array_equal(blk.mixer.zigzag_paths_reverse[i].numpy(), rev[i])
    g, cfg, _ = model_case("tiny_video_sst")
    m = build(cfg)
    fwd, rev, st = zo.build_scan_tables(cfg["scan_type"], cfg["depth"], (cfg["img_dim"] // cfg["patch_size"]) ** 2, cfg["video_frames"])
    for i, blk in enumerate(m.blocks):
        assert blk.mixer.st_order[i] == st[i]
        assert np.array_equal(blk.mixer.zigzag_paths[i].numpy(), fwd[i])
        assert np.array_equal(blk.mixer.zigzag_paths_reverse[i].numpy(), rev[i])


def test_reference_constructor_init_properties():
    """Init rules of the reference constructor (model_zigma.py:840-872, mamba_simple.py:128-162)."""
    from zigma_b200 import ZigMa
    torch.manual_seed(0)
    m = ZigMa(in_channels=4, embed_dim=32, depth=2, img_dim=8, scan_type="zigzagN8", use_pe=1, device="cpu")
    for blk in m.blocks:
        assert torch.count_nonzero(blk.adaLN_modulation[1].weight) == 0 and torch.count_nonzero(blk.adaLN_modulation[1].bias) == 0
        A = torch.exp(blk.mixer.A_log)
        assert torch.allclose(A, torch.arange(1, 17, dtype=torch.float32).repeat(64, 1), rtol=1e-5)
        dt = torch.nn.functional.softplus(blk.mixer.dt_proj.bias)
        assert dt.min() >= 1e-4 * 0.99 and dt.max() <= 0.1 * 1.01
    assert m.pos_embed.abs().sum() > 0 and not m.pos_embed.requires_grad


def test_embed_without_pos_embed_is_the_hand_off_of_the_folded_first_tail():
    """ZigMa.embed(add_pos=False) returns the tokens BEFORE `x = x + self.pos_embed` (model_zigma.py:941): the sampling engine adds the
    table inside its first fused tail (zg_block_tail_fwd_pe).  Everything else -- conditioning, the temporal embedding of video
    models -- is unchanged, and the default still includes the table."""
    from zigma_b200 import ZigMa
    torch.manual_seed(0)
    for cfg in (dict(in_channels=4, embed_dim=32, depth=2, img_dim=8, scan_type="zigzagN8", use_pe=2),
                dict(in_channels=4, embed_dim=32, depth=3, img_dim=8, patch_size=2, scan_type="zzvideo_sst", use_pe=2, video_frames=4, num_classes=5)):
        m = ZigMa(device="cpu", **cfg).eval()
        with torch.no_grad():
            m.pos_embed.normal_(0, 0.5)
        video = cfg.get("video_frames", 0) > 0
        x = torch.randn(2, 4, 4, 8, 8) if video else torch.randn(2, 4, 8, 8)
        t = torch.tensor([0.1, 0.7])
        y = torch.tensor([1, 3]) if cfg.get("num_classes", -1) > 0 else None
        with torch.no_grad():
            full, c1, _ = m.embed(x, t, y)
            bare, c2, _ = m.embed(x, t, y, add_pos=False)
        assert torch.equal(c1, c2)
        assert m.pos_embed.shape[1] == full.shape[1]            # one row per token: the (seqlen, dim) table of the fold
        if not (video and m.tpe):
            assert torch.allclose(bare + m.pos_embed, full, atol=1e-6)
        assert not torch.allclose(bare, full)


def test_temporal_composite_row_tables_are_the_reference_rearrange_plus_permutation():
    """engine._temporal_tables: the copy-free temporal layer addresses the (b, t k) token-major rows through ONE int32 table per
    direction.  In: position k T + t of the (b k) t working order reads model row perm[t] K + k -- the reference's
    `rearrange(x, "b (t k) d -> (b k) t d")` followed by the gather along t (mamba_simple.py:416-425); out: model row t K + k
    takes working row k T + perm_rev[t] (:436-442).  Pure index arithmetic, checked against the rearranged tensors."""
    import types
    from zigma_b200.engine import ZigMaEngine
    T, K, D = 8, 6, 3
    rng = np.random.RandomState(1)
    perm = torch.from_numpy(rng.permutation(T))
    rev = torch.empty_like(perm)
    rev[perm] = torch.arange(T)
    lay = {"perm64": perm, "perm_rev64": rev}
    stub = types.SimpleNamespace(_rev_cache={})
    tb = ZigMaEngine._temporal_tables(stub, lay, T, K, "cpu")
    assert tb["in"].dtype == torch.int32 and tb["out"].dtype == torch.int32 and tb["in"].numel() == T * K
    x = torch.randn(T * K, D)                                        # rows in (t, k) order
    work = x.view(T, K, D).permute(1, 0, 2)[:, perm].reshape(T * K, D)   # (k, t) order, every sequence permuted along t
    assert torch.equal(x[tb["in"].long()], work)
    y = torch.randn(K * T, D)                                        # a layer output in the working order
    back = y.view(K, T, D)[:, rev].permute(1, 0, 2).reshape(T * K, D)
    assert torch.equal(y[tb["out"].long()], back)
    assert torch.equal(x[tb["in"].long()][tb["out"].long()], x)      # the two tables are inverse of each other
    assert ZigMaEngine._temporal_tables(stub, lay, T, K, "cpu") is tb  # cached per (layer, T, K)


def test_sampler_euler_matches_oracle_and_reference_grid():
    from zigma_b200 import create_transport, Sampler
    tr = create_transport("Linear", "velocity", None, None, None)
    fn = Sampler(tr).sample_ode(sampling_method="euler", num_steps=50)
    W = torch.randn(6, 6) * 0.3
    model = lambda x, t, **kw: torch.tanh(x @ W) * (1 + t.view(-1, 1))
    x0 = torch.randn(4, 6)
    out = fn(x0, model)
    assert len(out) == 50                     # one state per grid point, like odeint
    ref = zo.sample_ode_fixed(model, x0, num_steps=50)
    assert torch.allclose(out[-1], ref, rtol=1e-5, atol=1e-6)
    calls = []
    fn(x0, lambda x, t, **kw: (calls.append(float(t[0])), x)[1])
    assert len(calls) == 49 and abs(calls[1] - 1 / 49) < 1e-6   # 49 evaluations on linspace(0, 1, 50)


def test_training_loss_shapes():
    from zigma_b200 import create_transport
    tr = create_transport()
    out = tr.training_losses(lambda x, t: x * 0, torch.randn(5, 4, 8, 8))
    assert out["loss"].shape == (5,)


def test_shard_sampling_world2_gloo():
    """bench.py's multi-GPU contract on CPU: 2 gloo ranks shard the batch, no collective inside the
    steps, one all_gather of the latents at the end, identical to the unsharded run."""
    script = os.path.join(ROOT, "tests", "_dist_worker.py")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", script],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "DIST_OK" in r.stdout


def test_permute_along_matches_advanced_indexing_fwd_and_bwd():
    """The permutation gather with its inverse-gather backward (mamba_simple._PermuteFn) vs the
    reference's ``x[:, :, perm]`` under autograd; a non-bijective index falls back to plain indexing."""
    from zigma_b200.mamba_simple import permute_along, _inverse_of
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(37, generator=g)
    for dim, shape in ((2, (3, 5, 37)), (1, (2, 37, 4))):
        x = torch.randn(*shape, generator=g, requires_grad=True)
        xr = x.detach().clone().requires_grad_()
        w = torch.randn(*shape, generator=g)
        y = permute_along(x, perm, dim)
        yr = xr[:, :, perm] if dim == 2 else xr[:, perm, :]
        assert torch.equal(y, yr) and y.is_contiguous()
        (y * w).sum().backward(); (yr * w).sum().backward()
        assert torch.equal(x.grad, xr.grad)
    assert torch.equal(_inverse_of(perm)[perm], torch.arange(37))
    dup = torch.tensor([0, 0, 2, 1])
    assert _inverse_of(dup) is None
    x = torch.randn(2, 3, 4, requires_grad=True)
    xr = x.detach().clone().requires_grad_()
    permute_along(x, dup, 2).sum().backward(); xr[:, :, dup].sum().backward()
    assert torch.equal(x.grad, xr.grad)


def test_reference_checkpoint_roundtrip(tmp_path):
    """train_acc.py:492-503 format ({"model", "ema", "opt"}, optionally "module."-prefixed) -> sample_acc.py:70-78 load."""
    from zigma_b200 import ZigMa
    from zigma_b200.checkpoint import load_reference_checkpoint, save_reference_checkpoint
    cfg = dict(img_dim=8, patch_size=1, in_channels=4, embed_dim=32, depth=2, scan_type="zigzagN8", num_classes=-1, has_text=False,
               use_pe=0, rms_norm=True, fused_add_norm=True, residual_in_fp32=True)
    torch.manual_seed(0)
    a, ema, b = ZigMa(device="cpu", **cfg), ZigMa(device="cpu", **cfg), ZigMa(device="cpu", **cfg)
    with torch.no_grad():
        for p in ema.parameters():
            p.add_(0.5)
    path = str(tmp_path / "0001000.pt")
    save_reference_checkpoint(path, a, ema_model=ema, opt_state={"state": {}}, args={"note": "x"}, ddp_prefix=True)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"model", "ema", "opt", "args", "train_steps", "best_fid"} and all(k.startswith("module.") for k in ck["ema"])
    missing, unexpected = load_reference_checkpoint(b, path)                # "ema", like sample_acc.py
    assert not missing and not unexpected
    assert all(torch.equal(v, ema.state_dict()[k]) for k, v in b.state_dict().items())
    load_reference_checkpoint(b, path, which="model")
    assert all(torch.equal(v, a.state_dict()[k]) for k, v in b.state_dict().items())
    load_reference_checkpoint(b, ema.state_dict())                           # bare state dict
    assert torch.equal(b.state_dict()["blocks.0.mixer.A_log"], ema.state_dict()["blocks.0.mixer.A_log"])
    with pytest.raises(RuntimeError):
        load_reference_checkpoint(b, {"ema": {"module.nope": torch.zeros(1)}})


def test_vae_decode_handoff_matches_the_reference_driver_lines():
    """zigma_b200.handoff against sample_acc.py:318-320,362-386 written out longhand with a stand-in VAE (the real one is a
    third-party diffusers model): latents / 0.18215 -> decode(...).sample -> clamp(127.5 x + 128, 0, 255) -> uint8; videos decode
    latents[i] for every index of the first axis and stack along dim 1; ground-truth latents are not rescaled."""
    import types
    import torch
    from zigma_b200 import decode_latents, to_uint8_pixels
    torch.manual_seed(0)
    up = torch.nn.ConvTranspose2d(4, 3, kernel_size=8, stride=8)

    class FakeVAE:
        def decode(self, z):
            return types.SimpleNamespace(sample=up(z))
    vae = FakeVAE()
    z = torch.randn(3, 4, 4, 4)
    with torch.no_grad():
        want = vae.decode(z / 0.18215).sample
        got = decode_latents(z, vae)
        assert torch.equal(got, want) and got.shape == (3, 3, 32, 32)
        assert torch.equal(decode_latents(z, vae, from_flow=False), vae.decode(z).sample)
        px = to_uint8_pixels(got)
        assert px.dtype == torch.uint8 and torch.equal(px, torch.clamp(127.5 * want + 128.0, 0, 255).to(torch.uint8))
        zv = torch.randn(2, 5, 4, 4, 4)                     # (batch, frames, C, h, w)
        gv = decode_latents(zv, vae, video=True)
        wv = torch.stack([vae.decode(zv[i] / 0.18215).sample for i in range(len(zv))], dim=1)
        assert torch.equal(gv, wv) and gv.shape == (5, 2, 3, 32, 32)
