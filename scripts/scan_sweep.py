"""Times the selective-scan forward kernel at the BASELINE config-2 layer shape (bs=64, E=1280, L=1024,
N=16, bf16) in both layouts.  ZG_SCAN_NPOLY is read once per process: run once per setting."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zigma_b200.selective_scan_interface import _scan_fwd
from zigma_b200 import zigzag_path
dev = "cuda"
bs, L, E, N, R = 64, 1024, 1280, 16, 40
gen = torch.Generator(device=dev).manual_seed(0)
dt = torch.bfloat16
xz = torch.randn(bs, L, 2 * E, device=dev, generator=gen).to(dt)
xc = torch.randn(bs, L, E, device=dev, generator=gen).to(dt)
dl = (0.5 * torch.rand(bs, L, E, device=dev, generator=gen)).to(dt)
xdbl = torch.randn(bs, L, R + 2 * N, device=dev, generator=gen).to(dt)
A = -0.5 * torch.rand(E, N, device=dev, generator=gen)
Dp, bias = torch.randn(E, device=dev, generator=gen), 0.5 * torch.rand(E, device=dev, generator=gen)
perm = torch.from_numpy(zigzag_path(32)[1]).to(dev).to(torch.int32)
Bv = xdbl[:, :, R:R + N].permute(0, 2, 1).unsqueeze(1)
Cv = xdbl[:, :, R + N:].permute(0, 2, 1).unsqueeze(1)
outb = torch.empty(bs, L, E, device=dev, dtype=dt).transpose(1, 2)
tok = lambda: _scan_fwd(xc.transpose(1, 2), dl.transpose(1, 2), A, Bv, Cv, Dp, xz[:, :, E:].transpose(1, 2), bias, True, z_rowmap=perm, want_last_state=False, out=outb)
u_s, d_s, z_s = xc.transpose(1, 2).contiguous(), dl.transpose(1, 2).contiguous(), xz[:, :, E:].transpose(1, 2).contiguous()
B_s, C_s = Bv.contiguous(), Cv.contiguous()
outs = torch.empty(bs, E, L, device=dev, dtype=dt)
seq = lambda: _scan_fwd(u_s, d_s, A, B_s, C_s, Dp, z_s, bias, True, want_last_state=False, out=outs)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
t_tok, t_seq = timeit(tok), timeit(seq)
tok(); seq()
ref = outs.float()
err = (outb.float() - torch.gather(ref, 2, torch.zeros(1, dtype=torch.long, device=dev).expand(1, 1, 1).expand(bs, E, 1)) * 0).abs().max().item()  # (layouts use different z order: no cross-check)
abytes = 4 * 2 * bs * E * L + 2 * 2 * bs * N * L + 4 * (E * N + 2 * E)
print(f"NPOLY={os.environ.get('ZG_SCAN_NPOLY', 'default')} TPC2={os.environ.get('ZG_SCAN_TPC2', '1')} TPC2_NPOLY={os.environ.get('ZG_SCAN_TPC2_NPOLY', '0')} token-major {t_tok:.4f} ms ({abytes / t_tok / 1e6:.0f} GB/s)  seq {t_seq:.4f} ms ({abytes / t_seq / 1e6:.0f} GB/s)")
