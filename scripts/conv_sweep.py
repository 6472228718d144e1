"""Times the token-major conv kernel at the BASELINE config-2 layer shape (ZG_CONV_VEC read once per process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zigma_b200.causal_conv1d_interface import _conv_fwd
from zigma_b200.engine import block_tail
from zigma_b200 import zigzag_path, reverse_permut_np
dev = "cuda"
bs, L, E, D = 64, 1024, 1280, 640
xz = torch.randn(bs, L, 2 * E, device=dev).bfloat16()
w, b = torch.randn(E, 4, device=dev).bfloat16(), torch.randn(E, device=dev).bfloat16()
perm = torch.from_numpy(zigzag_path(32)[1]).to(dev).to(torch.int32)
out = torch.empty(bs, L, E, device=dev, dtype=torch.bfloat16).transpose(1, 2)
fn = lambda: _conv_fwd(xz[:, :, :E].transpose(1, 2), w, b, True, x_rowmap=perm, out=out)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c) / n
t = timeit(fn)
x = torch.randn(bs, L, D, device=dev).bfloat16(); mix = torch.randn(bs, L, D, device=dev).bfloat16()
mods = torch.randn(bs, 3 * D, device=dev).bfloat16(); res = torch.randn(bs, L, D, device=dev); nw = torch.ones(D, device=dev).bfloat16()
rev = torch.from_numpy(reverse_permut_np(zigzag_path(32)[1])).to(dev).to(torch.int32)
t2 = timeit(lambda: block_tail(x, mix, mods[:, :D], mods[:, D:2*D], mods[:, 2*D:], nw, res, rev, 1e-5))
if os.environ.get("CONV_SWEEP_JSON"):      # machine-readable line for scripts
    import json
    print(json.dumps({"conv_us": t * 1e3, "tail_us": t2 * 1e3, "smem": os.environ.get("ZG_CONV_SMEM", ""), "lch": os.environ.get("ZG_CONV_SMEM_LCH", "")}))
print(f"ZG_CONV_SMEM={os.environ.get('ZG_CONV_SMEM','default')} LCH={os.environ.get('ZG_CONV_SMEM_LCH','default(32)')} ZG_CONV_VEC={os.environ.get('ZG_CONV_VEC','default(4)')}: conv {t*1e3:.1f} us ({2*2*bs*L*E/t/1e6:.0f} GB/s of 335 MB)   block_tail {t2*1e3:.1f} us ({bs*L*D*(2+2+4+4+2+2)/t2/1e6:.0f} GB/s of 671 MB)")
