"""Backward kernels at the BASELINE config-2 layer shape (channel-first, the autograd path): ours vs the
reference's selective_scan_cuda.bwd / causal_conv1d_bwd (oracle/_ref, when present)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zigma_b200.selective_scan_interface import _scan_fwd, _scan_bwd
from zigma_b200.causal_conv1d_interface import _conv_bwd
from oracle import ref_cuda
dev = "cuda"
bs, L, E, N = int(os.environ.get("BS", 16)), 1024, 1280, 16
dt = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
u = torch.randn(bs, E, L, device=dev, generator=g).to(dt); z = torch.randn(bs, E, L, device=dev, generator=g).to(dt)
delta = (0.5 * torch.rand(bs, E, L, device=dev, generator=g)).to(dt)
B = torch.randn(bs, 1, N, L, device=dev, generator=g).to(dt); C = torch.randn(bs, 1, N, L, device=dev, generator=g).to(dt)
A = -0.5 * torch.rand(E, N, device=dev, generator=g); D = torch.randn(E, device=dev, generator=g); bias = 0.5 * torch.rand(E, device=dev, generator=g)
dout = torch.randn(bs, E, L, device=dev, generator=g).to(dt)
def timeit(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
out, _, ckpt, saved = _scan_fwd(u, delta, A, B, C, D, z, bias, True, want_last_state=False, want_ckpt=True)
t_fwd_ck = timeit(lambda: _scan_fwd(u, delta, A, B, C, D, z, bias, True, want_last_state=False, want_ckpt=True))
t_bwd = timeit(lambda: _scan_bwd(saved, ckpt, dout, True))
print(f"bs={bs}: ours scan fwd(+ckpt) {t_fwd_ck:.3f} ms, scan bwd {t_bwd:.3f} ms")
# token-major storage of the same problem (the engine's layout)
tm = lambda a: a.transpose(-1, -2).contiguous().transpose(-1, -2)
ut, deltat, zt, Bt_, Ct_, doutt = tm(u), tm(delta), tm(z), tm(B), tm(C), tm(dout)
out_t, _, ckpt_t, saved_t = _scan_fwd(ut, deltat, A, Bt_, Ct_, D, zt, bias, True, want_last_state=False, want_ckpt=True)
t_fwd_t = timeit(lambda: _scan_fwd(ut, deltat, A, Bt_, Ct_, D, zt, bias, True, want_last_state=False, want_ckpt=True))
t_bwd_t = timeit(lambda: _scan_bwd(saved_t, ckpt_t, doutt, True))
print(f"bs={bs}: ours token-major scan fwd(+ckpt) {t_fwd_t:.3f} ms, scan bwd {t_bwd_t:.3f} ms")
m1, m2 = _scan_bwd(saved, ckpt, dout, True), _scan_bwd(saved_t, ckpt_t, doutt, True)
print("   channel-first vs token-major grads: " + ", ".join(f"{n} {(a_.float() - b_.float()).abs().max().item():.2e}" for n, a_, b_ in
      (("du", m1[0], m2[0]), ("ddelta", m1[1], m2[1]), ("dA", m1[2], m2[2]), ("dB", m1[3], m2[3]), ("dC", m1[4], m2[4]), ("dz", m1[7], m2[7]))))
w = torch.randn(E, 4, device=dev, generator=g).to(dt); cb = torch.randn(E, device=dev, generator=g).to(dt)
t_cb = timeit(lambda: _conv_bwd(u, w, cb, dout, True))
print(f"ours conv bwd {t_cb:.3f} ms")
from zigma_b200 import zigzag_path
perm = torch.from_numpy(zigzag_path(32)[1]).to(dev).to(torch.int32)
t_cbt = timeit(lambda: _conv_bwd(ut, w, cb, doutt, True, x_rowmap=perm))
print(f"ours token-major conv bwd (+rowmap) {t_cbt:.3f} ms")
if ref_cuda.available():
    ss, cc = ref_cuda.load()
    o = ss.fwd(u, delta, A, B, C, D, z, bias, True)
    outr, x = o[0], o[1]
    t_rf = timeit(lambda: ss.fwd(u, delta, A, B, C, D, z, bias, True))
    t_rb = timeit(lambda: ss.bwd(u, delta, A, B, C, D, z, bias, dout, x, outr, None, True, False))
    t_rcb = timeit(lambda: cc.causal_conv1d_bwd(u, w, cb, dout, None, True))
    print(f"reference scan fwd {t_rf:.3f} ms, scan bwd {t_rb:.3f} ms, conv bwd {t_rcb:.3f} ms")
    r = ss.bwd(u, delta, A, B, C, D, z, bias, dout, x, outr, None, True, False)
    mine = _scan_bwd(saved, ckpt, dout, True)
    for name, a_, b_ in (("du", mine[0], r[0]), ("ddelta", mine[1], r[1]), ("dA", mine[2], r[2]), ("dB", mine[3], r[3]), ("dC", mine[4], r[4]), ("dz", mine[7], r[7])):
        d = (a_.float() - b_.float()).abs().max().item(); m = b_.float().abs().max().item()
        print(f"   ours vs reference CUDA {name}: max|diff| {d:.3e} (max|ref| {m:.3e})")
