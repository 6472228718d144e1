"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference (CompVis/zigma) from /root/reference.

Only usable in the build container (``/root/reference`` does not exist on the GPU box).  Used by
``oracle/gen_golden.py`` to (a) pin the numpy/torch restatement in ``oracle/zigma_oracle.py`` against
the real reference code and (b) generate the golden vectors committed under ``tests/golden``.

The reference cannot be imported as a package here:
  * ``dis_mamba/mamba_ssm/__init__.py:3-5`` imports the LM head -> transformers 4.36 API (absent);
  * ``selective_scan_interface.py:9-11`` imports the compiled ``selective_scan_cuda`` /
    ``causal_conv1d_cuda`` extensions (not built, no GPU);
  * ``utils/utils_zigzag.py:4-5`` imports matplotlib, ``model_zigma.py:17`` imports timm (absent);
  * ``ops/triton/layernorm.py`` launches Triton kernels (no GPU).
So each reference file is loaded BY PATH into a synthetic package tree and the missing native
modules are replaced by stubs that forward to the reference's OWN pure-PyTorch oracles
(``selective_scan_ref``, ``causal_conv1d_ref``, ``rms_norm_ref``/``layer_norm_ref``).  No reference
source is copied; nothing here is imported by the product package.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("ZIGMA_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "dis_mamba"))


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _pkg(name):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = []  # mark as package
        sys.modules[name] = m
    return sys.modules[name]


_cache = {}


def load_reference():
    """Returns a namespace with the reference modules: .ssi (selective_scan_interface), .conv
    (causal_conv1d_interface), .ln (layernorm refs), .mamba_simple, .model_zigma, .zigzag."""
    if "ns" in _cache:
        return _cache["ns"]
    assert available(), f"reference tree not found at {REF_ROOT}"

    # ---- stubs for absent third-party / native modules ------------------------------------
    mpl = types.ModuleType("matplotlib"); mpl.__path__ = []
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))

    class PatchEmbed(nn.Module):
        """Shim with timm's PatchEmbed contract used by model_zigma.py:608-622,845,880:
        .proj Conv2d(k=s=patch), .patch_size tuple, .num_patches, forward = conv -> flatten(2).T"""
        def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, bias=True):
            super().__init__()
            self.img_size = (img_size, img_size)
            self.patch_size = (patch_size, patch_size)
            self.grid_size = (img_size // patch_size, img_size // patch_size)
            self.num_patches = self.grid_size[0] * self.grid_size[1]
            self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)
        def forward(self, x):
            return self.proj(x).flatten(2).transpose(1, 2)

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
            super().__init__()
            self.fc1 = nn.Linear(in_features, hidden_features or in_features)
            self.act = act_layer() if isinstance(act_layer, type) else act_layer
            self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)
        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    timm = _pkg("timm"); _pkg("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer")
    vt.PatchEmbed, vt.Mlp = PatchEmbed, Mlp
    sys.modules["timm.models.vision_transformer"] = vt

    # causal_conv1d python interface of the reference, with the native module stubbed
    ccuda = types.ModuleType("causal_conv1d_cuda")
    sys.modules["causal_conv1d_cuda"] = ccuda
    conv = _load("causal_conv1d.causal_conv1d_interface",
                 "dis_causal_conv1d/causal_conv1d/causal_conv1d_interface.py")
    ccuda.causal_conv1d_fwd = lambda x, w, b, silu: conv.causal_conv1d_ref(x, w, b, "silu" if silu else None)
    cpkg = _pkg("causal_conv1d")
    cpkg.causal_conv1d_fn = lambda x, w, b=None, activation=None: conv.causal_conv1d_ref(x, w, b, activation)
    cpkg.causal_conv1d_update = conv.causal_conv1d_update_ref
    cpkg.causal_conv1d_interface = conv

    scuda = types.ModuleType("selective_scan_cuda")
    sys.modules["selective_scan_cuda"] = scuda

    for p in ["dis_mamba", "dis_mamba.mamba_ssm", "dis_mamba.mamba_ssm.ops",
              "dis_mamba.mamba_ssm.ops.triton", "dis_mamba.mamba_ssm.modules", "utils"]:
        _pkg(p)
    ssi = _load("dis_mamba.mamba_ssm.ops.selective_scan_interface",
                "dis_mamba/mamba_ssm/ops/selective_scan_interface.py")

    def _fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus):
        # selective_scan_cuda.fwd contract (selective_scan.cpp:226-336): returns [out, x, (out_z)],
        # x[:, :, -1, 1::2] is the last state (selective_scan_interface.py:40).
        out, last = ssi.selective_scan_ref(u, delta, A, B, C, D, None, delta_bias, delta_softplus,
                                           return_last_state=True)
        x = torch.zeros(u.shape[0], u.shape[1], 1, 2 * A.shape[1], dtype=torch.float32)
        x[:, :, 0, 1::2] = last
        if z is None:
            return [out, x]
        out_z = (out.float() * F.silu(z.float())).to(out.dtype)
        return [out, x, out_z]
    scuda.fwd = _fwd
    # the stub selective_scan_fn used by the *_ref inner functions must not hit autograd.Function
    ssi.selective_scan_fn = ssi.selective_scan_ref

    ln = _load("dis_mamba.mamba_ssm.ops.triton.layernorm", "dis_mamba/mamba_ssm/ops/triton/layernorm.py")

    def _norm_fn(is_rms):
        ref = ln.rms_norm_ref if is_rms else ln.layer_norm_ref
        def fn(x, weight, bias, residual=None, prenorm=False, residual_in_fp32=False, eps=1e-6, **kw):
            # semantics of _layer_norm_fwd_1pass_kernel (layernorm.py:64-120): fp32 add, fp32
            # residual_out, y stored in x.dtype.
            dt = x.dtype
            y, res = ref(x, weight, bias, residual=residual, eps=eps, prenorm=True, upcast=True)
            y = y.to(dt)
            if not (residual_in_fp32 or (residual is not None and residual.dtype == torch.float32)):
                res = res.to(residual.dtype if residual is not None else dt)
            return (y, res) if prenorm else y
        return fn
    ln.rms_norm_fn = _norm_fn(True)
    ln.layer_norm_fn = _norm_fn(False)

    ms = _load("dis_mamba.mamba_ssm.modules.mamba_simple", "dis_mamba/mamba_ssm/modules/mamba_simple.py")
    zz = _load("utils.utils_zigzag", "utils/utils_zigzag.py")
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        mz = _load("model_zigma", "model_zigma.py")
    mz.rms_norm_fn, mz.layer_norm_fn = ln.rms_norm_fn, ln.layer_norm_fn

    ns = types.SimpleNamespace(ssi=ssi, conv=conv, ln=ln, mamba_simple=ms, model_zigma=mz, zigzag=zz)
    _cache["ns"] = ns
    return ns


if __name__ == "__main__":
    import contextlib, io, time
    ns = load_reference()
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ns.model_zigma.ZigMa(in_channels=4, embed_dim=64, depth=4, img_dim=8, patch_size=1,
                                 scan_type="zigzagN8", device="cpu", use_pe=2).eval()
    x = torch.randn(2, 4, 8, 8); t = torch.rand(2)
    t0 = time.time(); y = m(x, t); print("ref tiny forward", y.shape, float(y.abs().mean()), time.time() - t0)
