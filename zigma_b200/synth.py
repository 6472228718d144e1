"""Deterministic synthetic weights / inputs shared by the golden
generator (which loads them into the UNMODIFIED reference model) and by the tests / bench (which
load them into the product model).  Uses numpy's legacy RandomState (bit-stable across numpy
versions) keyed by the parameter NAME, so the same state dict is reproduced on any machine without
shipping megabytes of weights.

Recipe (SURVEY.md section 8d "Synthetic inputs"): realistic magnitudes, adaLN and pos_embed
re-randomised (they are zero at init in the reference, model_zigma.py:863-865,641-643, which would
null the whole mixer path), A_log jittered so the S4D-real structure A[d,n] = -(n+1) is not there
to exploit, dt bias = softplus^-1 of dt in [1e-3, 1e-1] as mamba_simple.py:137-146.
"""
import zlib

import numpy as np
import torch


def _rs(name, seed):
    return np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))


def synth_param(name, shape, seed=0):
    rs = _rs(name, seed)
    shape = tuple(shape)
    leaf = name.split(".")[-1]
    if leaf in ("A_log", "A_b_log"):
        E, N = shape
        base = np.log(np.tile(np.arange(1, N + 1, dtype=np.float64), (E, 1)))
        v = base + 0.1 * rs.randn(*shape)
    elif leaf in ("D", "D_b"):
        v = 1.0 + 0.1 * rs.randn(*shape)
    elif name.endswith("dt_proj.bias") or name.endswith("dt_proj_b.bias"):
        dt = np.exp(rs.rand(*shape) * (np.log(0.1) - np.log(0.001)) + np.log(0.001)).clip(min=1e-4)
        v = dt + np.log(-np.expm1(-dt))
    elif name.endswith("dt_proj.weight") or name.endswith("dt_proj_b.weight"):
        R = shape[1]
        v = rs.uniform(-R ** -0.5, R ** -0.5, size=shape)
    elif "norm" in name and leaf == "weight":
        v = 1.0 + 0.1 * rs.randn(*shape)
    elif leaf == "bias":
        v = 0.02 * rs.randn(*shape)
    elif "pos_embed" in name or "temporal_pos_embedding" in name:
        v = 0.02 * rs.randn(*shape)
    elif "adaLN_modulation" in name:
        v = 0.02 * rs.randn(*shape)
    elif "embedding_table" in name:
        v = 0.02 * rs.randn(*shape)
    elif "conv1d" in name and leaf == "weight":          # (E, 1, W) depthwise
        v = rs.uniform(-0.5, 0.5, size=shape)
    elif leaf == "weight" and len(shape) >= 2:            # Linear / Conv2d: U(-1/sqrt(fan_in), ..)
        fan_in = int(np.prod(shape[1:]))
        v = rs.uniform(-fan_in ** -0.5, fan_in ** -0.5, size=shape)
    else:
        v = 0.02 * rs.randn(*shape)
    return torch.from_numpy(np.asarray(v, dtype=np.float32))


def synth_state_dict(shapes, seed=0, dtype=torch.float32):
    """shapes: mapping name -> shape (e.g. {k: v.shape for k, v in model.state_dict().items()})."""
    return {k: synth_param(k, s, seed).to(dtype) for k, s in shapes.items()}


def synth_latents(shape, seed=0):
    return torch.from_numpy(_rs("latents", seed).randn(*shape).astype(np.float32))


def synth_scan_inputs(Bt, E, L, N, G=1, seed=0):
    """Distributions of dis_mamba/tests/ops/test_selective_scan.py:58-88."""
    rs = _rs("scan", seed)
    f = lambda a: torch.from_numpy(a.astype(np.float32))
    return dict(
        A=f(-0.5 * rs.rand(E, N)), B=f(rs.randn(Bt, G, N, L)), C=f(rs.randn(Bt, G, N, L)),
        D=f(rs.randn(E)), z=f(rs.randn(Bt, E, L)), delta_bias=f(0.5 * rs.rand(E)),
        u=f(rs.randn(Bt, E, L)), delta=f(0.5 * rs.rand(Bt, E, L)))
