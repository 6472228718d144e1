"""Pure-write / pure-read / copy bandwidth of the box at the size of one (batch x seqlen x d_inner) bf16 tensor of config 2 (168 MB):
what a write-bound GEMM (dt_proj: K = 40, 168 MB written) or a read-bound one (x_proj: 168 MB read) can reach at best."""
import torch
dev = "cuda"
n = 64 * 1024 * 1280
bufs = [torch.empty(n, device=dev, dtype=torch.bfloat16) for _ in range(4)]      # 4 x 168 MB: rotating set larger than L2
def timeit(fn, iters=20):
    for i in range(4): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
mb = n * 2 / 1e6
t = timeit(lambda i: bufs[i % 4].zero_())
print(f"write  (fill kernel)   {t*1e3:6.1f} us  {mb/t/1e3:6.2f} TB/s written")
t = timeit(lambda i: torch.cuda.current_stream().synchronize() if False else bufs[i % 4].view(torch.int16).sum())
print(f"read   (sum reduction) {t*1e3:6.1f} us  {mb/t/1e3:6.2f} TB/s read")
t = timeit(lambda i: bufs[i % 4].copy_(bufs[(i + 1) % 4]))
print(f"copy   (r + w)         {t*1e3:6.1f} us  {2*mb/t/1e3:6.2f} TB/s total")
