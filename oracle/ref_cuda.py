"""BENCH/TEST INFRASTRUCTURE -- the reference's OWN compiled CUDA kernels (oracle/_ref/*.so, built
from the unmodified sources under /root/reference by oracle/build_ref.sh, gencode swapped to
sm_100a) behind the restated glue of oracle/zigma_oracle.py: the "reference vendored CUDA path"
baseline that BASELINE.md section 5 asks to time on the same box.  Never imported by the product
package; bench.py uses it only to report the baseline next to our numbers.

Glue = what the reference's Python does around its kernels (channel-first layout, index_select
permutation + .contiguous(), separate x_proj / dt_proj GEMMs, rearranged contiguous B/C, unfused
modulate / gate), restated in zigma_oracle.py and pinned against the reference there.  The one
substitution: the reference's Triton add+RMSNorm kernel cannot travel (Python source), so the
baseline borrows OUR fused add+norm kernel for that op -- this can only flatter the baseline.
"""
import importlib.util
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_mods = {}


def available():
    return all(os.path.exists(os.path.join(_REF, n + ".so")) for n in ("selective_scan_cuda", "causal_conv1d_cuda"))


def load():
    if not _mods:
        for name in ("selective_scan_cuda", "causal_conv1d_cuda"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(_REF, name + ".so"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _mods[name] = mod
    return _mods["selective_scan_cuda"], _mods["causal_conv1d_cuda"]


def scan_fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus=True):
    """selective_scan_cuda.fwd (selective_scan.cpp:226-336): returns out_z (or out when z is None)."""
    ss, _ = load()
    if B.dim() == 3:
        B, C = B.unsqueeze(1), C.unsqueeze(1)
    res = ss.fwd(u, delta, A, B.contiguous(), C.contiguous(), D, z, delta_bias, delta_softplus)
    return res[-1] if z is not None else res[0]


def conv_fwd(x, w, b, silu=True):
    _, cc = load()
    return cc.causal_conv1d_fwd(x, w, b, silu)


def backend(norm_fn=None):
    """BACKEND dict for zigma_oracle: reference CUDA kernels (+ our norm kernel, see module doc)."""
    be = {"conv": lambda x, w, b: conv_fwd(x, w.contiguous(), b, True),
          "scan": lambda u, d, A, B, C, D, z, bias: scan_fwd(u, d, A, B, C, D, z, bias, True)}
    if norm_fn is not None:
        be["norm"] = norm_fn
    return be
