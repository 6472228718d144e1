import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zigma_b200 import rms_norm_fn
from oracle import zigma_oracle as zo
torch.manual_seed(1)
M, N = 5, 64
x = torch.randn(M, N, device="cuda", requires_grad=True); res = torch.randn(M, N, device="cuda", requires_grad=True)
w = (1 + 0.1 * torch.randn(N, device="cuda")).requires_grad_()
gy, gr = torch.randn(M, N, device="cuda"), torch.randn(M, N, device="cuda")
y, r = rms_norm_fn(x, w, None, residual=res, prenorm=True, residual_in_fp32=True, eps=1e-5)
for name, (a, b) in {"dy only": (gy, torch.zeros_like(gr)), "dres only": (torch.zeros_like(gy), gr), "both": (gy, gr)}.items():
    gx, gres, gw = torch.autograd.grad([y, r], [x, res, w], [a, b], retain_graph=True)
    xr, rr, wr = x.detach().cpu().requires_grad_(), res.detach().cpu().requires_grad_(), w.detach().cpu().requires_grad_()
    y2, r2 = zo.add_norm(xr, wr, None, rr, True, True, 1e-5, True)
    ex, er, ew = torch.autograd.grad([y2, r2], [xr, rr, wr], [a.cpu(), b.cpu()])
    print(name, "dx err", (gx.cpu() - ex).abs().max().item(), "dres err", (gres.cpu() - er).abs().max().item(), "dw err", (gw.cpu() - ew).abs().max().item(),
          "| ref dx max", ex.abs().max().item())
    print("   gx[0,:4]", gx[0, :4].tolist(), "ex[0,:4]", ex[0, :4].tolist())
