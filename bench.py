#!/usr/bin/env python
"""bench.py -- ZigMa denoiser hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one denoiser evaluation (ZigMa.forward at bs=64 per GPU, zigzag8_b1: D=640, depth=18,
32x32 latents, patch 1, bf16, synthetic weights / latents) followed by the Euler update of the
flow-matching sampler (transport/integrators.py:105-123) -- BASELINE.json configs[1].
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(in_channels=4, embed_dim=640, depth=18, img_dim=32, patch_size=1, scan_type="zigzagN8", use_pe=2)
BS_PER_GPU = 64
L_TOKENS = 1024
NUM_GRID = 50           # linspace(0, 1, 50): the sampler's time grid


def scan_algorithmic_bytes(Bt, E, L, N, s):
    """4 s B E L (u, delta, z read; out written) + 2 s B N L (B, C) + 4 (E N + 2E)."""
    return 4 * s * Bt * E * L + 2 * s * Bt * N * L + 4 * (E * N + 2 * E)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy kernel, burst)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def cpu_reference_eval(nthreads=None, bs=2):
    """The reference's algorithm on the host cores: the CPU port in oracle/ (torch GEMMs + the
    plain-C OpenMP scan), fp32, one denoiser evaluation at batch `bs`.  Returns seconds."""
    import torch
    from oracle import zigma_oracle as zo
    from zigma_b200 import synth, ZigMa
    if nthreads:
        torch.set_num_threads(min(nthreads, 32))   # small-batch GEMMs stop scaling (and regress) beyond ~32 threads
    zo.USE_C_SCAN = True
    shapes = {k: tuple(v.shape) for k, v in ZigMa(device="cpu", **CFG).state_dict().items()}
    sd = synth.synth_state_dict(shapes, seed=0)
    x = synth.synth_latents((bs, 4, 32, 32), seed=0)
    t = torch.full((bs,), 0.5)
    cfg = dict(CFG, norm_epsilon=1e-5)
    with torch.no_grad():
        zo.zigma_forward(sd, cfg, x[:1], t[:1])     # warm-up (page in, build the C oracle)
        t0 = time.perf_counter()
        zo.zigma_forward(sd, cfg, x, t)
        return time.perf_counter() - t0


def reference_cuda_numbers(bs, dtype_name="bf16", n_iter=5):
    """The reference's vendored CUDA kernels (oracle/_ref, built from the unmodified sources for
    sm_100a) on this GPU: (a) selective_scan_cuda.fwd and causal_conv1d_fwd at the layer shape of the
    workload in the reference's own layout, (b) one denoiser evaluation through the reference's
    (restated) PyTorch glue around those kernels.  Returns None when oracle/_ref is absent."""
    import torch
    from oracle import ref_cuda, zigma_oracle as zo
    if not (ref_cuda.available() and torch.cuda.is_available()):
        return None
    from zigma_b200 import synth, ZigMa, rms_norm_fn
    dev = torch.device("cuda", torch.cuda.current_device())
    dtype = torch.bfloat16 if dtype_name == "bf16" else torch.float32
    E, N, L = 2 * CFG["embed_dim"], 16, L_TOKENS

    def timeit(fn, n=n_iter, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    gen = torch.Generator(device=dev).manual_seed(0)
    xz = torch.randn(bs, 2 * E, L, device=dev, generator=gen).to(dtype)
    u, z = xz[:, :E], xz[:, E:]
    delta = (0.5 * torch.rand(E, bs * L, device=dev, generator=gen)).to(dtype).reshape(E, bs, L).transpose(0, 1)   # view like :323
    Bm = torch.randn(bs, 1, N, L, device=dev, generator=gen).to(dtype)
    Cm = torch.randn(bs, 1, N, L, device=dev, generator=gen).to(dtype)
    A = -0.5 * torch.rand(E, N, device=dev, generator=gen)
    Dp, bias = torch.randn(E, device=dev, generator=gen), 0.5 * torch.rand(E, device=dev, generator=gen)
    cw, cb = torch.randn(E, 4, device=dev, generator=gen).to(dtype), torch.randn(E, device=dev, generator=gen).to(dtype)
    out = {"scan_fwd_ms": timeit(lambda: ref_cuda.scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True), 10),
           "conv_fwd_ms": timeit(lambda: ref_cuda.conv_fwd(u, cw, cb, True), 10)}
    from zigma_b200 import selective_scan_fn, causal_conv1d_fn
    out["ours_same_layout_scan_fwd_ms"] = timeit(lambda: selective_scan_fn(u, delta, A, Bm, Cm, Dp, z=z, delta_bias=bias, delta_softplus=True), 10)
    out["ours_same_layout_conv_fwd_ms"] = timeit(lambda: causal_conv1d_fn(u, cw, cb, "silu"), 10)
    # whole denoiser evaluation: restated reference glue + reference kernels
    shapes = {k: tuple(v.shape) for k, v in ZigMa(device="cpu", **CFG).state_dict().items()}
    sd = {k: v.to(dev) for k, v in synth.synth_state_dict(shapes, seed=0, dtype=dtype).items()}
    norm = lambda x, w, b, residual=None, prenorm=False, residual_in_fp32=False, eps=1e-6: rms_norm_fn(
        x, w, b, residual=residual, prenorm=prenorm, residual_in_fp32=residual_in_fp32, eps=eps)
    zo.BACKEND = ref_cuda.backend(norm)
    x = torch.randn(bs, 4, 32, 32, device=dev, generator=gen).to(dtype)
    t = torch.full((bs,), 0.5, device=dev, dtype=dtype)
    cfg = dict(CFG, norm_epsilon=1e-5)
    try:
        with torch.no_grad():
            ms = timeit(lambda: zo.zigma_forward(sd, cfg, x, t), n_iter)
    finally:
        zo.BACKEND = {}
    out.update({"denoiser_eval_ms": ms, "tokens_per_s": bs * L / (ms * 1e-3), "bs": bs, "dtype": dtype_name,
                "what": "reference dis_mamba/dis_causal_conv1d CUDA kernels (sm_100a build of the unmodified sources) + the reference's PyTorch glue (restated), eager"})
    return out


def run_reference(args, rank):
    """--impl reference: the reference's own (CPU) implementation of the path, timed on the host
    cores.  The reference is Python; it cannot travel to the GPU box, so its CPU restatement
    (oracle/, pinned against the unmodified reference by tests/golden) is what runs -- kind 'port'."""
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    bs = 2
    for _ in range(min(args.warmup, 1)):
        cpu_reference_eval(cores, bs)
    n = max(1, min(args.steps, 5))
    dts = [cpu_reference_eval(cores, bs) for _ in range(n)]
    dt = sum(dts) / len(dts)
    val = bs * L_TOKENS / dt
    line = {
        "impl": "reference", "metric": "denoiser tokens/s (bs*L*evals/s), zigzag8_b1 32x32, Euler sampling loop",
        "value": val, "unit": "tokens/s", "n_gpus": args.gpus, "steps": n, "warmup": min(args.warmup, 1),
        "ms_per_step": dt * 1e3 * (BS_PER_GPU / bs), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "zigzag8_b1 D=640 depth=18 32x32 patch1; bounded sample: bs=2 per eval (ms_per_step scaled to bs=64)",
                   "denoiser_steps_per_s_at_bs64": val / (BS_PER_GPU * L_TOKENS)},
        "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": cores, "kind": "port",
                         "sample": f"{n} denoiser evals at bs={bs} (fp32, torch {torch.get_num_threads()} threads + OpenMP C scan)"},
        "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        line["reference_cuda"] = reference_cuda_numbers(BS_PER_GPU)
    except Exception as ex:   # the CUDA baseline is extra information; the arm's contract is the CPU number
        line["reference_cuda"] = {"error": repr(ex)[:300]}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bs", type=int, default=BS_PER_GPU, help="batch per GPU (BASELINE config: 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly (for ncu launch lists)")
    ap.add_argument("--no-train", action="store_true", help="skip the forward+backward (training step) side measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    from zigma_b200 import ZigMa, _lib, synth, create_transport, Sampler
    from zigma_b200.sharding import gather_latents
    from zigma_b200.selective_scan_interface import _scan_fwd

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    bs = args.bs
    dtype = torch.bfloat16

    model = ZigMa(device=dev, dtype=dtype, **CFG).eval()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synth.synth_state_dict(shapes, seed=0, dtype=dtype))
    # initial noise indexed by the global sample id (weak scaling: bs per GPU fixed)
    z0 = torch.stack([synth.synth_latents((4, 32, 32), seed=1000 + rank * bs + i) for i in range(bs)]).to(dev).to(dtype)
    ts = torch.linspace(0, 1, NUM_GRID).tolist()
    dt_step = ts[1] - ts[0]
    tvec = torch.empty(bs, device=dev, dtype=dtype)

    def euler_step(x, i):
        tvec.fill_(ts[i % (NUM_GRID - 1)])
        return x + dt_step * model(x, tvec)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        # launches of OUR kernels per evaluation, counted on one eager (non-graph) pass
        os.environ["ZIGMA_CUDA_GRAPH"] = "0"
        from zigma_b200.engine import ZigMaEngine
        eager = ZigMaEngine(model)
        c0 = _lib.launch_count()
        eager._forward_impl(z0, tvec.fill_(0.5), None)
        per_eval = _lib.launch_count() - c0
        os.environ["ZIGMA_CUDA_GRAPH"] = "0" if args.no_graph else "1"
        model._engine = None

        x = z0
        for i in range(W):
            x = euler_step(x, i)
        barrier()
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.profiler.start()     # no-op unless run under `ncu --profile-from-start off`
        e0.record()
        for i in range(K):
            x = euler_step(x, W + i)
        full = gather_latents(x, bs * world, world)      # the single collective of the sampling job
        e1.record()
        barrier()
        torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        clk = clocks.stop() if rank == 0 else None
        if world > 1:
            tm = torch.tensor([ms], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms = tm.item()
        assert torch.isfinite(full.float()).all()

        # ---- e2e: public API with HOST buffers: pinned latents/t H2D, forward, result D2H, every step
        hx = z0.cpu().pin_memory()
        ht = torch.empty(bs, dtype=dtype).pin_memory()
        hout = torch.empty_like(hx).pin_memory()
        dx = torch.empty_like(z0)

        def e2e_step(i):
            ht.fill_(ts[i % (NUM_GRID - 1)])
            dx.copy_(hx, non_blocking=True)
            tvec.copy_(ht, non_blocking=True)
            out = model(dx, tvec)
            hout.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()     # the caller reads the result on the host
        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            e2e_step(i)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            tm = torch.tensor([e2e_ms], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            e2e_ms = tm.item()

        # ---- roofline of the dominant kernel (selective scan, layer shape of this config) -----------------
        roof = None
        if rank == 0:
            E, N, R = 2 * CFG["embed_dim"], 16, 40
            gen = torch.Generator(device=dev).manual_seed(0)
            xz = torch.randn(bs, L_TOKENS, 2 * E, device=dev, generator=gen).to(dtype)
            xc = torch.randn(bs, L_TOKENS, E, device=dev, generator=gen).to(dtype)
            dl = (0.5 * torch.rand(bs, L_TOKENS, E, device=dev, generator=gen)).to(dtype)
            xdbl = torch.randn(bs, L_TOKENS, R + 2 * N, device=dev, generator=gen).to(dtype)
            A = -0.5 * torch.rand(E, N, device=dev, generator=gen)
            Dp, bias = torch.randn(E, device=dev, generator=gen), 0.5 * torch.rand(E, device=dev, generator=gen)
            perm = eager.layers[1]["perm"]
            Bv = xdbl[:, :, R:R + N].permute(0, 2, 1).unsqueeze(1)
            Cv = xdbl[:, :, R + N:].permute(0, 2, 1).unsqueeze(1)
            outb = torch.empty(bs, L_TOKENS, E, device=dev, dtype=dtype).transpose(1, 2)
            call = lambda: _scan_fwd(xc.transpose(1, 2), dl.transpose(1, 2), A, Bv, Cv, Dp, xz[:, :, E:].transpose(1, 2), bias, True,
                                     z_rowmap=perm, want_last_state=False, out=outb)
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            n_it = 20
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(n_it):
                call()      # working set 4 x 168 MB > 126 MB L2: every launch streams from HBM
            s1.record()
            torch.cuda.synchronize()
            kms = s0.elapsed_time(s1) / n_it
            peak, how = measured_peaks()
            abytes = scan_algorithmic_bytes(bs, E, L_TOKENS, N, 2)
            ach = abytes / (kms * 1e-3) / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", "scan_fwd_traffic.json")
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            roof = {"bound": "hbm", "kernel": "zg::scan_fwd_tpc2_kernel<bf16> (dstate 16, token-major, z gathered through the zigzag table)", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "peak_source": how, "ms_per_launch": kms, "algorithmic_bytes": abytes,
                    "launches_per_step": CFG["depth"], "share_of_step": CFG["depth"] * kms / (ms / K)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tokens = bs * world * L_TOKENS * K
    value = tokens / (ms * 1e-3)
    line = {
        "metric": "denoiser tokens/s (bs*L*evals/s), zigzag8_b1 32x32, Euler sampling loop",
        "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"zigzag8_b1: ZigMa D=640 depth=18 img 32 patch 1 zigzagN8, bs={bs}/GPU, fixed-grid Euler (linspace(0,1,50))",
                   "global_batch": bs * world, "seq_len": L_TOKENS, "parallelism": f"dp{world}",
                   "l2_policy": "no flush: per-step working set (xz alone 335 MB per layer) >> 126 MB L2",
                   "denoiser_steps_per_s": K / (ms * 1e-3), "cuda_graph": not args.no_graph,
                   "gemm": "hand-written tcgen05 (zg_gemm_bf16_tn)" if os.environ.get("ZIGMA_TCGEN05", "1") == "1" else "library (cuBLAS)", "collective": "one all_gather of final latents, inside the timed region"},
        "clocks": clk,
        "e2e": {"value": bs * world * L_TOKENS * K / (e2e_ms * 1e-3), "unit": "tokens/s",
                "h2d_bytes_per_step": z0.numel() * z0.element_size() + bs * 2, "d2h_bytes_per_step": z0.numel() * z0.element_size(),
                "ms_per_step": e2e_ms / K, "api": "ZigMa.forward(x, t) with pinned host latents in, host velocity out"},
        "gpu_launches": per_eval * K,
        "gpu_launches_per_eval": per_eval,
        "roofline": roof,
    }
    if not args.no_ref_cuda:
        try:
            rc = reference_cuda_numbers(bs)
        except Exception as ex:
            rc = {"error": repr(ex)[:300]}
        line["reference_cuda"] = rc
        if rc and "denoiser_eval_ms" in rc:
            line["speedup_vs_reference_cuda"] = rc["denoiser_eval_ms"] / (ms / K)
    if not args.no_train and world == 1:
        # side measurement (not the headline): one flow-matching training step, forward + backward, of the same
        # denoiser at bs 16 -- ours vs the reference's forward/backward CUDA kernels in the reference's glue
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "train_bench.py")], capture_output=True, text=True, timeout=600,
                               env=dict(os.environ, BS="16", DTYPE="bf16", CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local))))
            js = [l for l in r.stdout.splitlines() if l.startswith("{")]
            line["train_step"] = json.loads(js[-1]) if js else {"error": (r.stderr or r.stdout)[-300:]}
        except Exception as ex:
            line["train_step"] = {"error": repr(ex)[:300]}
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        dt = cpu_reference_eval(cores, 2)
        line["cpu_baseline"] = {"value": 2 * L_TOKENS / dt, "unit": "tokens/s", "cores": cores, "kind": "port",
                                "sample": "1 denoiser eval at bs=2 (fp32 CPU port of the reference path: torch GEMMs + OpenMP C scan)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
