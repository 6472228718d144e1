"""CPU: the oracle (torch restatement + plain-C port) against the golden vectors produced by the
UNMODIFIED reference (oracle/gen_golden.py).  This is the pin that lets the GPU tests trust it."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, synth, zigma_oracle as zo
from util import check_close, gold, model_case, t

SCAN = ["t128_g1", "t131_g2", "e64_n16", "plain", "noz", "l1", "l16_many", "n4", "e128_g2_n16", "e96_l45_n16"]


@pytest.mark.parametrize("name", SCAN)
def test_scan_oracles_match_reference(name):
    g = gold("scan_" + name)
    Bt, E, L, N, G, hasD, hasz, hasb, sp = [int(v) for v in g["flags"]]
    B, C = t(g["B"]), t(g["C"])
    args = (t(g["u"]), t(g["delta"]), t(g["A"]), B if G > 1 else B[:, 0], C if G > 1 else C[:, 0],
            t(g["D"]) if hasD else None, t(g["z"]) if hasz else None, t(g["delta_bias"]) if hasb else None, bool(sp))
    out, last = zo.selective_scan(*args, return_last_state=True)
    check_close(out, g["out"], f"torch oracle {name} out")
    check_close(last, g["last_state"], f"torch oracle {name} state")
    out_c, last_c = c_oracle.scan_fwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g["D"] if hasD else None,
                                      g["z"] if hasz else None, g["delta_bias"] if hasb else None, bool(sp))
    check_close(out_c, g["out"], f"C oracle {name} out")
    check_close(last_c, g["last_state"], f"C oracle {name} state")


def test_scan_config1_digest():
    """BASELINE config 1 (B=2 L=1024 D=640 N=16, fp32) through the C oracle vs the reference digest."""
    g = gold("scan_config1_digest")
    inp = synth.synth_scan_inputs(2, 640, 1024, 16, 1, seed=2)
    out, last = c_oracle.scan_fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"], inp["delta_bias"], True)
    check_close(out.reshape(-1)[g["idx"]], g["out_sub"], "config1 out (subsample)", max_strict_viol=1e-3)
    check_close(last.reshape(-1)[::13], g["last_sub"], "config1 last_state (subsample)", max_strict_viol=1e-3)


def test_conv_oracles_match_reference():
    g = gold("conv")
    x = t(g["x"])
    for W in (2, 3, 4):
        for silu in (0, 1):
            for hb in (0, 1):
                tag = f"W{W}_s{silu}_b{hb}"
                w, b = t(g[f"w{W}"]), t(g[f"b{W}"]) if hb else None
                check_close(zo.causal_conv1d(x, w, b, "silu" if silu else None), g["out_" + tag], "conv torch " + tag)
                check_close(c_oracle.conv1d_fwd(g["x"], g[f"w{W}"], g[f"b{W}"] if hb else None, bool(silu)), g["out_" + tag], "conv C " + tag)


def test_norm_oracle_matches_reference():
    g = gold("norm")
    x, res, w, b = t(g["x"]), t(g["res"]), t(g["w"]), t(g["b"])
    for rms in (1, 0):
        for hr in (1, 0):
            y, r = zo.add_norm(x, w, None if rms else b, res if hr else None, True, True, 1e-5, bool(rms))
            check_close(y, g[f"y_rms{rms}_res{hr}"], f"norm rms={rms} res={hr}")
            check_close(r, g[f"r_rms{rms}_res{hr}"], "norm residual")


def test_mamba_inner_oracle_matches_reference():
    g = gold("mamba_inner")
    a = {k: t(g[k]) for k in g.files}
    out = zo.mamba_inner(a["xz"], a["conv_w"], a["conv_b"], a["x_proj_w"], a["dt_proj_w"], a["out_proj_w"], a["out_proj_b"],
                         a["A"], a["D"], a["delta_bias"])
    check_close(out, g["out"], "mamba_inner oracle")


@pytest.mark.parametrize("name", ["tiny_zigzag8", "tiny_sweep2", "tiny_hilbert2", "tiny_patch2_cls", "tiny_video_sst", "tiny_text", "tiny_video_text"])
def test_zigma_forward_oracle_matches_reference(name):
    from oracle.gen_golden import model_io
    g, cfg, shapes = model_case(name)
    sd = synth.synth_state_dict(shapes, seed=0)
    x, tt, y = model_io(cfg, g["out"].shape[0])
    cfg = dict(cfg, norm_epsilon=1e-5)
    out = zo.zigma_forward(sd, cfg, x, tt, y)
    check_close(out, g["out"], f"zigma_forward oracle {name}")


def test_euler_sampler_oracle():
    """Fixed-grid Euler on dx/dt = -x has the closed form x0 * (1 - dt)^(n-1)."""
    x0 = torch.ones(2, 3)
    x = zo.sample_ode_fixed(lambda x, t: -x, x0, num_steps=11)
    assert torch.allclose(x, x0 * 0.9 ** 10, atol=1e-6)
