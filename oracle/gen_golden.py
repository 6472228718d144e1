"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(imported from /root/reference through oracle/ref_loader.py) on deterministic synthetic inputs
(oracle/synth.py), and at the same time pins oracle/zigma_oracle.py + oracle/scan_oracle.c against
it (hard asserts below).  Runs only in the build container; the vectors it writes are committed.

    python oracle/gen_golden.py            # all
    python oracle/gen_golden.py tables scan # subset
"""
import contextlib
import io
import json
import os
import sys
import hashlib

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader, synth, zigma_oracle as zo, c_oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
torch.set_grad_enabled(True)


def close(a, b, rtol=1e-4, atol=2e-5, what=""):
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    err = (a - b).abs().max().item()
    ok = torch.allclose(a, b, rtol=rtol, atol=atol)
    print(f"    pin {what:<40s} max|diff|={err:.3e} {'OK' if ok else 'FAIL'}")
    assert ok, what


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().float().numpy() if v.dtype in (torch.bfloat16, torch.float16) else v.detach().numpy()
        out[k] = v
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {os.path.relpath(path, ROOT)} ({os.path.getsize(path) / 1024:.1f} KiB)")


# ------------------------------------------------------------------------------------------------
def gen_tables(ref):
    arrs = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for N in (1, 2, 3, 4, 5, 8, 16, 32, 64):
            r = np.stack(ref.zigzag.zigzag_path(N)).astype(np.int64)
            assert np.array_equal(r, np.stack(zo.zigzag_path(N))), N
            for p in r:
                assert np.array_equal(ref.zigzag.reverse_permut_np(p), zo.reverse_permut_np(p))
            arrs[f"zigzag_{N}"] = r.astype(np.int32)
        for N in (2, 3, 4, 6, 8, 16, 32):
            r = np.stack(ref.zigzag.hilbert_path(N)).astype(np.int64)
            assert np.array_equal(r, np.stack(zo.hilbert_path(N))), N
            arrs[f"hilbert_{N}"] = r.astype(np.int32)
    z32 = np.stack(zo.zigzag_path(32)).astype(np.int64)
    arrs["sha256_zigzag_32_int64"] = np.frombuffer(hashlib.sha256(z32.tobytes()).digest(), dtype=np.uint8)
    print("    sha256(zigzag_path(32) int64) =", hashlib.sha256(z32.tobytes()).hexdigest()[:16], "(SURVEY: 01b6ef874ac9cd89)")
    save("tables", **arrs)


# ------------------------------------------------------------------------------------------------
SCAN_CASES = [
    # name, Bt, E, L, N, G, has_D, has_z, has_bias, softplus
    ("t128_g1", 2, 4, 128, 8, 1, True, True, True, True),      # test_selective_scan.py:54-56
    ("t131_g2", 2, 4, 131, 8, 2, True, True, True, True),      # ragged L, 2 groups
    ("e64_n16", 1, 64, 256, 16, 1, True, True, True, True),
    ("plain", 2, 8, 37, 16, 1, False, False, False, False),
    ("noz", 2, 8, 64, 16, 1, True, False, True, True),
    ("l1", 1, 8, 1, 16, 1, True, True, True, True),
    ("l16_many", 12, 16, 16, 16, 1, True, True, True, True),   # video temporal shape
    ("n4", 1, 8, 40, 4, 1, True, True, True, True),
    ("e128_g2_n16", 2, 128, 72, 16, 2, True, True, True, True),   # whole 64-channel tiles, 2 groups: the staged dstate-16 backward
    ("e96_l45_n16", 2, 96, 45, 16, 1, True, True, True, True),    # partial tile + ragged L: the unstaged dstate-16 backward
]


def gen_scan(ref, only=None):
    for (name, Bt, E, L, N, G, hasD, hasz, hasb, sp) in SCAN_CASES:
        if only is not None and name not in only:
            continue
        print(f"  scan case {name}")
        inp = synth.synth_scan_inputs(Bt, E, L, N, G, seed=1)
        req = {k: v.clone().requires_grad_() for k, v in inp.items()}
        B_in = req["B"] if G > 1 else req["B"][:, 0]
        C_in = req["C"] if G > 1 else req["C"][:, 0]
        out, last = ref.ssi.selective_scan_ref(
            req["u"], req["delta"], req["A"], B_in, C_in, req["D"] if hasD else None,
            z=req["z"] if hasz else None, delta_bias=req["delta_bias"] if hasb else None,
            delta_softplus=sp, return_last_state=True)
        g = torch.from_numpy(np.random.RandomState(5).randn(*out.shape).astype(np.float32))
        out.backward(g)
        grads = {"d" + k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in req.items()}
        # pin the restatements
        o2, l2 = zo.selective_scan(inp["u"], inp["delta"], inp["A"],
                                   inp["B"] if G > 1 else inp["B"][:, 0], inp["C"] if G > 1 else inp["C"][:, 0],
                                   inp["D"] if hasD else None, inp["z"] if hasz else None,
                                   inp["delta_bias"] if hasb else None, sp, True)
        close(o2, out, what="torch restatement out"); close(l2, last, what="torch restatement state")
        o3, l3 = c_oracle.scan_fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"],
                                   inp["D"] if hasD else None, inp["z"] if hasz else None,
                                   inp["delta_bias"] if hasb else None, sp)
        close(o3, out, rtol=1e-4, atol=1e-5, what="C restatement out"); close(l3, last, rtol=1e-4, atol=1e-5, what="C state")
        save("scan_" + name, out=out, last_state=last, g=g,
             flags=np.array([Bt, E, L, N, G, hasD, hasz, hasb, sp], dtype=np.int32), **inp, **grads)

    if only is not None:
        return
    # config 1 of BASELINE.json: B=2 L=1024 D=640 N=16 -- digest only (full output is 5 MB)
    print("  scan case config1 (B=2 L=1024 D=640 N=16)")
    inp = synth.synth_scan_inputs(2, 640, 1024, 16, 1, seed=2)
    out, last = ref.ssi.selective_scan_ref(inp["u"], inp["delta"], inp["A"], inp["B"][:, 0], inp["C"][:, 0],
                                           inp["D"], z=inp["z"], delta_bias=inp["delta_bias"],
                                           delta_softplus=True, return_last_state=True)
    o3, l3 = c_oracle.scan_fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"],
                               inp["delta_bias"], True)
    close(o3, out, rtol=1e-3, atol=2e-4, what="C restatement config1 (fp32 noise floor, see DESIGN.md)")
    flat = out.reshape(-1)
    idx = np.arange(0, flat.numel(), 97)
    save("scan_config1_digest", idx=idx.astype(np.int64), out_sub=flat[idx], last_sub=last.reshape(-1)[::13],
         out_sum=np.float64(flat.double().sum().item()), out_abs_sum=np.float64(flat.double().abs().sum().item()))


# ------------------------------------------------------------------------------------------------
def gen_conv(ref):
    arrs = {}
    rs = np.random.RandomState(11)
    Bt, E, L = 2, 24, 151
    x = torch.from_numpy(rs.randn(Bt, E, L).astype(np.float32))
    g = torch.from_numpy(rs.randn(Bt, E, L).astype(np.float32))
    arrs["x"], arrs["g"] = x, g
    for W in (2, 3, 4):
        w = torch.from_numpy(rs.randn(E, W).astype(np.float32))
        b = torch.from_numpy(rs.randn(E).astype(np.float32))
        arrs[f"w{W}"], arrs[f"b{W}"] = w, b
        for silu in (0, 1):
            for hb in (0, 1):
                xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
                out = ref.conv.causal_conv1d_ref(xr, wr, br if hb else None, "silu" if silu else None)
                out.backward(g)
                tag = f"W{W}_s{silu}_b{hb}"
                close(zo.causal_conv1d(x, w, b if hb else None, "silu" if silu else None), out, what="conv " + tag)
                close(c_oracle.conv1d_fwd(x, w, b if hb else None, bool(silu)), out, rtol=1e-5, atol=1e-5, what="conv C " + tag)
                arrs["out_" + tag] = out
                if tag in ("W4_s1_b1", "W3_s0_b0", "W2_s1_b0"):   # gradient pins (kept few: file size)
                    arrs["dx_" + tag], arrs["dw_" + tag] = xr.grad, wr.grad
                    if hb:
                        arrs["db_" + tag] = br.grad
    save("conv", **arrs)


def gen_norm(ref):
    rs = np.random.RandomState(12)
    M, N = 6, 48
    x = torch.from_numpy(rs.randn(2, 3, N).astype(np.float32))
    res = torch.from_numpy(rs.randn(2, 3, N).astype(np.float32))
    w = torch.from_numpy((1 + 0.1 * rs.randn(N)).astype(np.float32))
    b = torch.from_numpy((0.1 * rs.randn(N)).astype(np.float32))
    arrs = dict(x=x, res=res, w=w, b=b)
    for rms in (1, 0):
        for hr in (1, 0):
            fn = ref.ln.rms_norm_fn if rms else ref.ln.layer_norm_fn
            y, r = fn(x, w, None if rms else b, residual=res if hr else None, prenorm=True,
                      residual_in_fp32=True, eps=1e-5)
            y2, r2 = zo.add_norm(x, w, None if rms else b, res if hr else None, True, True, 1e-5, bool(rms))
            close(y2, y, what=f"norm rms={rms} res={hr}"); close(r2, r, what="norm residual")
            arrs[f"y_rms{rms}_res{hr}"], arrs[f"r_rms{rms}_res{hr}"] = y, r
    # bf16 input, fp32 residual stream (the model's use)
    xb = x.bfloat16()
    y, r = ref.ln.rms_norm_fn(xb, w.bfloat16(), None, residual=res, prenorm=True, residual_in_fp32=True, eps=1e-5)
    y2, r2 = zo.add_norm(xb, w.bfloat16(), None, res, True, True, 1e-5, True)
    close(y2, y, what="norm bf16"); close(r2, r, what="norm bf16 residual")
    arrs["y_bf16"], arrs["r_bf16"] = y, r
    save("norm", **arrs)


def gen_inner(ref):
    rs = np.random.RandomState(13)
    Bt, E, L, N, R, W, Dm = 2, 32, 48, 8, 3, 4, 16
    f = lambda *s, sc=1.0: torch.from_numpy((sc * rs.randn(*s)).astype(np.float32))
    a = dict(xz=f(Bt, 2 * E, L), conv_w=f(E, 1, W, sc=0.5), conv_b=f(E, sc=0.1), x_proj_w=f(R + 2 * N, E, sc=E ** -0.5),
             dt_proj_w=f(E, R, sc=R ** -0.5), out_proj_w=f(Dm, E, sc=E ** -0.5), out_proj_b=f(Dm, sc=0.1),
             A=-torch.from_numpy(rs.rand(E, N).astype(np.float32)) - 0.1, D=f(E), delta_bias=f(E, sc=0.3))
    out = ref.ssi.mamba_inner_ref(a["xz"], a["conv_w"], a["conv_b"], a["x_proj_w"], a["dt_proj_w"], a["out_proj_w"],
                                  a["out_proj_b"], a["A"], None, None, a["D"], a["delta_bias"], delta_softplus=True)
    out2 = zo.mamba_inner(a["xz"], a["conv_w"], a["conv_b"], a["x_proj_w"], a["dt_proj_w"], a["out_proj_w"],
                          a["out_proj_b"], a["A"], a["D"], a["delta_bias"])
    close(out2, out, rtol=1e-4, atol=1e-5, what="mamba_inner")
    # also the autograd.Function path of the reference (MambaInnerFn.forward through the stubs)
    out3 = ref.ssi.mamba_inner_fn(a["xz"], a["conv_w"], a["conv_b"], a["x_proj_w"], a["dt_proj_w"], a["out_proj_w"],
                                  a["out_proj_b"], a["A"], None, None, a["D"], a["delta_bias"], None, None, True)
    close(out3, out, rtol=1e-4, atol=1e-5, what="mamba_inner_fn(stubbed) vs ref")
    # bidirectional variant (unused by ZigMa, part of the op surface): bimamba_inner_ref with a second A
    a["A_b"] = -torch.from_numpy(rs.rand(E, N).astype(np.float32)) - 0.1
    out_bi = ref.ssi.bimamba_inner_ref(a["xz"], a["conv_w"], a["conv_b"], a["x_proj_w"], a["dt_proj_w"], a["out_proj_w"],
                                       a["out_proj_b"], a["A"], a["A_b"], None, None, a["D"], a["delta_bias"], delta_softplus=True)
    save("mamba_inner", out=out, out_bi=out_bi, **a)


# ------------------------------------------------------------------------------------------------
MODEL_CASES = {
    # name: (ctor kwargs, batch, dtype, has_y)
    "tiny_zigzag8": (dict(in_channels=4, embed_dim=32, depth=8, img_dim=8, patch_size=1, scan_type="zigzagN8", use_pe=2), 2, "fp32"),
    "tiny_zigzag8_bf16": (dict(in_channels=4, embed_dim=32, depth=8, img_dim=8, patch_size=1, scan_type="zigzagN8", use_pe=2), 2, "bf16"),
    "tiny_sweep2": (dict(in_channels=4, embed_dim=32, depth=2, img_dim=8, patch_size=1, scan_type="v2", use_pe=2), 2, "fp32"),
    "tiny_hilbert2": (dict(in_channels=4, embed_dim=32, depth=3, img_dim=8, patch_size=1, scan_type="hilbertN2", use_pe=0), 2, "fp32"),
    "tiny_patch2_cls": (dict(in_channels=4, embed_dim=32, depth=3, img_dim=16, patch_size=2, scan_type="zigzagN8", use_pe=1, num_classes=10), 3, "fp32"),
    "tiny_video_sst": (dict(in_channels=4, embed_dim=32, depth=6, img_dim=8, patch_size=2, scan_type="zzvideo_sst", use_pe=2,
                            video_frames=4, tpe=True, num_classes=5), 2, "fp32"),
    # has_text: the gated cross-attention branch of every block + the text conditioning path (model_zigma.py:95-135, 446-458, 930-933)
    "tiny_text": (dict(in_channels=4, embed_dim=64, depth=3, img_dim=8, patch_size=1, scan_type="zigzagN8", use_pe=2, has_text=True,
                       d_context=24, n_context_token=7), 2, "fp32"),
    # has_text on a factorised video scan (spatial, spatial, temporal layers): the cross-attention branch on (b, t k) tokens
    "tiny_video_text": (dict(in_channels=4, embed_dim=64, depth=3, img_dim=8, patch_size=2, scan_type="zzvideo_sst", use_pe=2,
                             video_frames=8, tpe=True, has_text=True, d_context=24, n_context_token=7), 2, "fp32"),
    "full_zigzag8_b1": (dict(in_channels=4, embed_dim=640, depth=18, img_dim=32, patch_size=1, scan_type="zigzagN8", use_pe=2), 1, "fp32"),
}


def build_reference_model(ref, kw, dtype):
    mz = ref.model_zigma
    orig = mz.Mamba

    def mamba_video_shim(*a, scan_type="v2", **k):
        # The shipped reference cannot construct its own 3d configs: ZigMa forwards
        # scan_type="zzvideo_*" (model_zigma.py:746,807) to Mamba whose assert only knows the
        # "video_" prefix (mamba_simple.py:164-171,396).  The harness (not the reference source)
        # renames the prefix on the way in so that branch can be exercised.
        if scan_type.startswith("zzvideo_"):
            scan_type = scan_type.replace("zzvideo_", "video_")
        return orig(*a, scan_type=scan_type, **k)
    mz.Mamba = mamba_video_shim
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            m = mz.ZigMa(device="cpu", dtype=dtype, drop_path_rate=0.1, **kw).eval()
    finally:
        mz.Mamba = orig
    return m


def model_io(kw, bs, seed=3):
    if kw.get("video_frames", 0) > 0:
        x = synth.synth_latents((bs, kw["video_frames"], kw["in_channels"], kw["img_dim"], kw["img_dim"]), seed)
    else:
        x = synth.synth_latents((bs, kw["in_channels"], kw["img_dim"], kw["img_dim"]), seed)
    t = torch.linspace(0.1, 0.9, bs)
    y = None
    if kw.get("has_text", False):
        y = synth.synth_latents((bs, kw["n_context_token"], kw["d_context"]), seed + 100)        # the text encoder's output
    elif kw.get("num_classes", -1) > 0:
        y = torch.arange(bs) % kw["num_classes"]
    return x, t, y


def gen_models(ref, only=None):
    for name, (kw, bs, dts) in MODEL_CASES.items():
        if only and name not in only:
            continue
        print(f"  model case {name}")
        dtype = torch.float32 if dts == "fp32" else torch.bfloat16
        m = build_reference_model(ref, kw, dtype)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        sd = synth.synth_state_dict(shapes, seed=0, dtype=dtype)
        m.load_state_dict(sd, strict=True)
        x, t, y = model_io(kw, bs)
        with torch.no_grad():
            if y is not None and y.is_floating_point():
                y = y.to(dtype)
            out = m(x.to(dtype), t.to(dtype), y)
            cfg = dict(kw); cfg.setdefault("norm_epsilon", 1e-5)
            out2 = zo.zigma_forward(sd, cfg, x.to(dtype), t.to(dtype), y)
        if dts == "fp32":
            close(out2, out, rtol=1e-3, atol=2e-5, what="zigma_forward restatement")
        else:
            close(out2, out, rtol=5e-2, atol=5e-2, what="zigma_forward restatement (bf16)")
        save("model_" + name, out=out, t=t, y=(y.float().numpy() if (y is not None and y.is_floating_point()) else y.numpy() if y is not None else np.zeros(0, np.int64)),
             shapes_json=np.frombuffer(json.dumps(shapes).encode(), dtype=np.uint8),
             cfg_json=np.frombuffer(json.dumps(kw).encode(), dtype=np.uint8))


def main():
    what = sys.argv[1:] or ["tables", "scan", "conv", "norm", "inner", "models"]
    ref = ref_loader.load_reference()
    for w in what:
        print(f"== {w}")
        if w.startswith("model:"):
            gen_models(ref, only=w.split(":", 1)[1].split(","))
        elif w.startswith("scan:"):
            gen_scan(ref, only=w.split(":", 1)[1].split(","))
        else:
            {"tables": gen_tables, "scan": gen_scan, "conv": gen_conv, "norm": gen_norm,
             "inner": gen_inner, "models": gen_models}[w](ref)


if __name__ == "__main__":
    main()
