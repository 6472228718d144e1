"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the ZigMa denoiser hot path.

A from-scratch restatement (numpy for the integer tables, torch-CPU fp32 for the arithmetic) of
the algorithm the reference implements on the path named by BASELINE.json ``north_star``.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this file; the product package ``zigma_b200`` never does.

PARITY PIN: every function below is checked against the UNMODIFIED reference, imported from
``/root/reference`` by ``oracle/ref_loader.py`` (stubs forward to the reference's own
``selective_scan_ref`` / ``causal_conv1d_ref`` / ``rms_norm_ref``), by ``oracle/gen_golden.py``;
the vectors it produced are committed under ``tests/golden`` and re-checked by
``tests/test_oracle_golden.py`` on every run.  The reference itself stores no golden vectors for
this path, so the pin is "outputs of the reference itself run here".

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

USE_C_SCAN = False   # route mamba_inner's scan through the plain-C port (fp32 only)
# Optional op backends for mamba_inner / zigma_forward: {"conv": f(x, w, b) -> silu(conv),
# "scan": f(u, delta, A, B, C, D, z, delta_bias) -> y * silu(z), "norm": add_norm-like}.  Used by
# oracle/ref_cuda.py to drive the SAME restated glue with the reference's compiled CUDA kernels
# (the "reference vendored CUDA path" baseline of bench.py).  None = the CPU restatements below.
BACKEND = {}

# ------------------------------------------------------------------------------------------------
# a1. scan-path tables (integer, bit exact)
# ------------------------------------------------------------------------------------------------

def zigzag_path(N):
    """8 boustrophedon paths over an N x N row-major grid.  utils/utils_zigzag.py:144-175.

    For each start corner (top-left, top-right, bottom-left, bottom-right) first the row snake
    ("lr") then the column snake ("tb").  Entry k of a path is the flat grid index visited k-th."""
    paths = []
    i, j = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")  # i: slow loop, j: fast loop
    snake = np.where(i % 2 == 0, j, N - 1 - j)                      # position inside line i
    for r0, c0, dr, dc in [(0, 0, 1, 1), (0, N - 1, 1, -1), (N - 1, 0, -1, 1), (N - 1, N - 1, -1, -1)]:
        # lr: line = row i, inside the row walk columns `snake`
        paths.append(((r0 + dr * i) * N + c0 + dc * snake).reshape(-1).astype(np.int64))
        # tb: line = column i (reference's j), inside the column walk rows `snake`
        paths.append(((r0 + dr * snake) * N + c0 + dc * i).reshape(-1).astype(np.int64))
    return paths


def reverse_permut_np(perm):
    """Inverse permutation: rev[perm[i]] = i.  utils/utils_zigzag.py:136-141."""
    perm = np.asarray(perm)
    rev = np.zeros(len(perm), dtype=np.int64)
    rev[perm] = np.arange(len(perm), dtype=np.int64)
    return rev


def _sgn(v):
    return (v > 0) - (v < 0)


def _gilbert_d(x, y, w, h):
    """Index along the generalised Hilbert curve of cell (x, y) of a w x h grid.
    utils/utils_zigzag.py:16-120 (itself a port of jakubcerveny/gilbert).  Iterative restatement
    of the tail-recursive descent."""
    idx, px, py = 0, 0, 0
    ax, ay, bx, by = (w, 0, 0, h) if w >= h else (0, h, w, 0)

    def inside(qx, qy, sx, sy, ax, ay, bx, by):
        dx, dy = ax + bx, ay + by
        okx = (sx + dx < qx <= sx) if dx < 0 else (sx <= qx < sx + dx)
        oky = (sy + dy < qy <= sy) if dy < 0 else (sy <= qy < sy + dy)
        return okx and oky

    while True:
        ww, hh = abs(ax + ay), abs(bx + by)
        dax, day, dbx, dby = _sgn(ax), _sgn(ay), _sgn(bx), _sgn(by)
        dx, dy = dax + dbx, day + dby
        if hh == 1:
            return idx + (dy * (y - py) if dax == 0 else dx * (x - px))
        if ww == 1:
            return idx + (dy * (y - py) if dbx == 0 else dx * (x - px))
        ax2, ay2, bx2, by2 = ax // 2, ay // 2, bx // 2, by // 2
        w2, h2 = abs(ax2 + ay2), abs(bx2 + by2)
        if 2 * ww > 3 * hh:
            if (w2 % 2) and ww > 2:
                ax2, ay2 = ax2 + dax, ay2 + day
            if inside(x, y, px, py, ax2, ay2, bx, by):
                ax, ay = ax2, ay2
                continue
            idx += abs((ax2 + ay2) * (bx + by))
            px, py, ax, ay = px + ax2, py + ay2, ax - ax2, ay - ay2
            continue
        if (h2 % 2) and hh > 2:
            bx2, by2 = bx2 + dbx, by2 + dby
        if inside(x, y, px, py, bx2, by2, ax2, ay2):
            ax, ay, bx, by = bx2, by2, ax2, ay2
            continue
        idx += abs((bx2 + by2) * (ax2 + ay2))
        if inside(x, y, px + bx2, py + by2, ax, ay, bx - bx2, by - by2):
            px, py, bx, by = px + bx2, py + by2, bx - bx2, by - by2
            continue
        idx += abs((ax + ay) * ((bx - bx2) + (by - by2)))
        px, py = px + (ax - dax) + (bx2 - dbx), py + (ay - day) + (by2 - dby)
        ax, ay, bx, by = -bx2, -by2, -(ax - ax2), -(ay - ay2)


def hilbert_path(N):
    """8 variants (rot90 x transpose) of the N x N gilbert ORDER-INDEX map, flattened.
    utils/utils_zigzag.py:123-130,285-302.  NB the reference uses the order-index array itself
    as the gather permutation (not its inverse); we reproduce that."""
    order = np.zeros((N, N), dtype=np.int64)
    for x in range(N):
        for y in range(N):
            order[x, y] = _gilbert_d(x, y, N, N)
    out = []
    for k in range(4):
        r = np.rot90(order, k)
        out += [r, r.T]
    return [np.ascontiguousarray(o).reshape(-1) for o in out]


_BUILD_CACHE = {}


def build_scan_tables(scan_type, depth, num_patches, video_frames=0):
    """Per-layer (perm, perm_rev, st_order) lists as ZigMa.__init__ builds them.
    model_zigma.py:689-794.  Memoised (the reference builds them once in the constructor)."""
    key = (scan_type, depth, num_patches, video_frames)
    if key not in _BUILD_CACHE:
        _BUILD_CACHE[key] = _build_scan_tables(scan_type, depth, num_patches, video_frames)
    return _BUILD_CACHE[key]


def _build_scan_tables(scan_type, depth, num_patches, video_frames=0):
    side = int(math.sqrt(num_patches))
    if scan_type.startswith("zigzagN") or scan_type.startswith("hilbertN"):
        n = int(scan_type.replace("zigzagN", "").replace("hilbertN", ""))
        base = (zigzag_path(side) if scan_type.startswith("zigzagN") else hilbert_path(side))[:n]
        assert len(base) == n
        fwd = base * depth
        rev = [reverse_permut_np(p) for p in base] * depth
        return fwd, rev, None
    if scan_type.startswith("zzvideo_"):
        st = list(scan_type.replace("zzvideo_", "")) * depth
        zz = zigzag_path(side) * depth
        zr = [reverse_permut_np(p) for p in zigzag_path(side)] * depth
        tp = np.arange(video_frames, dtype=np.int64)
        tn = tp[::-1].copy()
        tz, tr = [tp, tn] * depth, [tn, tp] * depth
        fwd, rev = [], []
        for d in range(depth):
            if st[d] == "s":
                fwd.append(zz.pop(0)); rev.append(zr.pop(0))
            else:
                fwd.append(tz.pop(0)); rev.append(tr.pop(0))
        return fwd, rev, st
    if scan_type in ("v1", "v2"):
        return None, None, None
    raise ValueError(scan_type)


# ------------------------------------------------------------------------------------------------
# a3. causal depthwise conv1d (+SiLU)
# ------------------------------------------------------------------------------------------------

def causal_conv1d(x, weight, bias=None, activation=None):
    """x (B, C, L), weight (C, W), bias (C,).  out[l] = bias + sum_w weight[w] * x[l-(W-1-w)],
    zero history, optional SiLU; computed in weight dtype, returned in x dtype.
    dis_causal_conv1d/causal_conv1d/causal_conv1d_interface.py:49-65, causal_conv1d_fwd.cu:103-118."""
    dt = x.dtype
    xw = x.to(weight.dtype)
    C, W = weight.shape
    L = x.shape[-1]
    xp = F.pad(xw, (W - 1, 0))
    acc = torch.zeros_like(xw)
    for w in range(W):
        acc = acc + weight[:, w].view(1, C, 1) * xp[..., w:w + L]
    if bias is not None:
        acc = acc + bias.view(1, C, 1)
    if activation in ("silu", "swish"):
        acc = acc * torch.sigmoid(acc)
    return acc.to(dt)


# ------------------------------------------------------------------------------------------------
# a4. selective scan (S6 recurrence)
# ------------------------------------------------------------------------------------------------

def selective_scan(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                   return_last_state=False):
    """u, delta, z: (Bt, E, L); A: (E, N) real; B, C: (Bt, N, L) | (Bt, G, N, L) | (E, N);
    D, delta_bias: (E,).  All math fp32, result cast to u.dtype.
    dis_mamba/mamba_ssm/ops/selective_scan_interface.py:86-152 and the fp32 semantics of
    selective_scan_fwd_kernel.cuh:153-171,216-261,280-298 (softplus threshold 20, h0 = 0)."""
    dt_in = u.dtype
    u32, d32 = u.float(), delta.float()
    if delta_bias is not None:
        d32 = d32 + delta_bias.float().view(1, -1, 1)
    if delta_softplus:
        d32 = torch.where(d32 <= 20.0, torch.log1p(torch.exp(torch.clamp(d32, max=20.0))), d32)
    Bt, E, L = u32.shape
    N = A.shape[1]
    A32 = A.float()

    def per_channel(M):  # -> callable l -> (Bt, E, N)
        M = M.float()
        if M.dim() == 2:
            return lambda l: M.unsqueeze(0)
        if M.dim() == 3:
            return lambda l: M[:, None, :, l]
        G = M.shape[1]
        Mx = M.repeat_interleave(E // G, dim=1)
        return lambda l: Mx[:, :, :, l]

    Bf, Cf = per_channel(B), per_channel(C)
    h = torch.zeros(Bt, E, N, dtype=torch.float32)
    ys = []
    for l in range(L):
        dl = d32[:, :, l].unsqueeze(-1)
        h = torch.exp(dl * A32.unsqueeze(0)) * h + (dl * u32[:, :, l].unsqueeze(-1)) * Bf(l)
        ys.append((h * Cf(l)).sum(-1))
    y = torch.stack(ys, dim=2)
    if D is not None:
        y = y + u32 * D.float().view(1, -1, 1)
    if z is not None:
        zf = z.float()
        y = y * (zf * torch.sigmoid(zf))
    y = y.to(dt_in)
    return (y, h) if return_last_state else y


# ------------------------------------------------------------------------------------------------
# a7. fused add + RMSNorm / LayerNorm
# ------------------------------------------------------------------------------------------------

def add_norm(x, weight, bias=None, residual=None, prenorm=False, residual_in_fp32=False, eps=1e-6,
             is_rms_norm=True):
    """r = x (+ residual) in fp32; y = norm(r) * w (+ b) stored in x.dtype; residual_out = r stored
    in residual dtype, or fp32 when residual_in_fp32, else x dtype.
    dis_mamba/mamba_ssm/ops/triton/layernorm.py:64-120,123-177,380-422 (+ oracles :19-48)."""
    r = x.float()
    if residual is not None:
        r = r + residual.float()
    if is_rms_norm:
        y = r * torch.rsqrt(r.square().mean(-1, keepdim=True) + eps)
    else:
        mu = r.mean(-1, keepdim=True)
        y = (r - mu) * torch.rsqrt((r - mu).square().mean(-1, keepdim=True) + eps)
    if weight is not None:
        y = y * weight.float()
    if bias is not None:
        y = y + bias.float()
    y = y.to(x.dtype)
    if not prenorm:
        return y
    if residual is not None:
        rdt = residual.dtype
    else:
        rdt = torch.float32 if residual_in_fp32 else x.dtype
    return y, r.to(rdt)


# ------------------------------------------------------------------------------------------------
# a5. mamba inner (conv -> x_proj -> dt_proj -> scan -> [out_proj])
# ------------------------------------------------------------------------------------------------

def mamba_inner(xz, conv_w, conv_b, x_proj_w, dt_proj_w, out_proj_w, out_proj_b, A, D, delta_bias,
                no_out_proj=False):
    """xz (Bt, 2E, L) -> (Bt, L, Dm) (or (Bt, E, L) when no_out_proj).  Intermediate tensors carry
    xz.dtype exactly as in MambaInnerFn.forward (selective_scan_interface.py:296-365): conv output,
    x_dbl, delta, B, C are all materialised in the activation dtype; scan math is fp32."""
    Bt, E2, L = xz.shape
    E = E2 // 2
    R = dt_proj_w.shape[1]
    N = A.shape[1]
    x, z = xz[:, :E], xz[:, E:]
    if "conv" in BACKEND:
        xc = BACKEND["conv"](x, conv_w.reshape(E, -1), conv_b)
    else:
        xc = causal_conv1d(x, conv_w.reshape(E, -1), conv_b, "silu")
    x_dbl = F.linear(xc.transpose(1, 2).reshape(Bt * L, E), x_proj_w)           # (Bt*L, R+2N)
    delta = (dt_proj_w @ x_dbl[:, :R].t()).reshape(E, Bt, L).transpose(0, 1)    # (Bt, E, L)
    Bm = x_dbl[:, R:R + N].reshape(Bt, L, N).transpose(1, 2)                    # (Bt, N, L)
    Cm = x_dbl[:, R + N:].reshape(Bt, L, N).transpose(1, 2)
    if "scan" in BACKEND:
        y = BACKEND["scan"](xc, delta, A, Bm, Cm, D, z, delta_bias)
    elif USE_C_SCAN and xz.dtype == torch.float32:
        # same recurrence through oracle/scan_oracle.c (OpenMP over (batch, channel)): used where the
        # python time loop is too slow (bench cpu_baseline, full-size model checks)
        from . import c_oracle
        yc, _ = c_oracle.scan_fwd(xc.contiguous().numpy(), delta.contiguous().numpy(), A.numpy(), Bm.contiguous().numpy()[:, None],
                                  Cm.contiguous().numpy()[:, None], D.numpy(), z.contiguous().numpy(), delta_bias.numpy(), True)
        y = torch.from_numpy(yc)
    else:
        y = selective_scan(xc, delta, A, Bm, Cm, D, z=z, delta_bias=delta_bias, delta_softplus=True)
    if no_out_proj:
        return y
    return F.linear(y.transpose(1, 2), out_proj_w, out_proj_b)


# ------------------------------------------------------------------------------------------------
# a6. Mamba mixer with the ZigMa scan-type dispatch
# ------------------------------------------------------------------------------------------------

_TABLE_CACHE = {}


def _dev_table(a, device):
    """int64 index tensor of a path table on `device`, cached (the reference keeps them resident too)."""
    key = (id(a), str(device))
    if key not in _TABLE_CACHE:
        _TABLE_CACHE[key] = (a, torch.as_tensor(a, dtype=torch.long).to(device))   # keep `a` alive: id() stays unique
    return _TABLE_CACHE[key][1]


def mamba_mixer(h, p, scan_type, perm=None, perm_rev=None, st=None, video_frames=0):
    """h (Bt, L, Dm); p: dict of the mixer's parameters (reference state-dict names without the
    ``blocks.i.mixer.`` prefix).  dis_mamba/mamba_ssm/modules/mamba_simple.py:274-444."""
    Bt, L, Dm = h.shape
    xz = (p["in_proj.weight"] @ h.reshape(Bt * L, Dm).t()).reshape(-1, Bt, L).transpose(0, 1)  # (Bt,2E,L)
    A = -torch.exp(p["A_log"].float())
    args = dict(conv_w=p["conv1d.weight"], conv_b=p["conv1d.bias"], x_proj_w=p["x_proj.weight"],
                dt_proj_w=p["dt_proj.weight"], out_proj_w=p["out_proj.weight"], out_proj_b=None,
                A=A, D=p["D"].float(), delta_bias=p["dt_proj.bias"].float())
    if scan_type == "v1":
        return mamba_inner(xz, **args)
    if scan_type == "v2":  # bidirectional sweep with a second parameter set, one out_proj (:304-339)
        yf = mamba_inner(xz, no_out_proj=True, **args)
        argb = dict(conv_w=p["conv1d_b.weight"], conv_b=p["conv1d_b.bias"], x_proj_w=p["x_proj_b.weight"],
                    dt_proj_w=p["dt_proj_b.weight"], out_proj_w=None, out_proj_b=None,
                    A=-torch.exp(p["A_b_log"].float()), D=p["D_b"].float(),
                    delta_bias=p["dt_proj_b.bias"].float())
        yb = mamba_inner(xz.flip(-1), no_out_proj=True, **argb)
        return F.linear((yf + yb.flip(-1)).transpose(1, 2), p["out_proj.weight"], None)
    perm_t, rev_t = _dev_table(perm, h.device), _dev_table(perm_rev, h.device)
    if st is None:  # zigzagN / hilbertN / randomN (:356-395)
        out = mamba_inner(xz[:, :, perm_t].contiguous(), **args)
        return out[:, rev_t, :].contiguous()
    # video: factorised spatial / temporal scans (:396-442)
    T = video_frames
    K = L // T
    if st == "s":
        xr = xz.reshape(Bt, -1, T, K).permute(0, 2, 1, 3).reshape(Bt * T, -1, K)
    else:
        xr = xz.reshape(Bt, -1, T, K).permute(0, 3, 1, 2).reshape(Bt * K, -1, T)
    out = mamba_inner(xr[:, :, perm_t].contiguous(), **args)[:, rev_t, :]
    if st == "s":
        return out.reshape(Bt, T, K, Dm).reshape(Bt, L, Dm)
    return out.reshape(Bt, K, T, Dm).permute(0, 2, 1, 3).reshape(Bt, L, Dm)


# ------------------------------------------------------------------------------------------------
# a8/a9. Block and ZigMa.forward, functional over a reference-layout state dict
# ------------------------------------------------------------------------------------------------

def timestep_embedding(t, dim, dtype, max_period=10000):
    """model_zigma.py:247-268 -- NB the frequency table is computed in the MODEL dtype."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=dtype) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def cross_attention(x, text, p, heads=8):
    """CrossAttention.forward (model_zigma.py:95-135): q from the tokens, k / v from the (embedded) text tokens, `heads` heads,
    softmax(q k^T / sqrt(dim_head)) v, output projection with bias.  p: to_q / to_k / to_v .weight, to_out.0.weight / .bias.
    (The reference dispatches to torch SDPA, xformers or explicit math by availability; all three are this formula.)"""
    Bt, L, _ = x.shape
    q, k, v = F.linear(x, p["to_q.weight"]), F.linear(text, p["to_k.weight"]), F.linear(text, p["to_v.weight"])
    split = lambda a: a.reshape(Bt, a.shape[1], heads, -1).transpose(1, 2)          # "B L (H D) -> B H L D"
    q, k, v = split(q), split(k), split(v)
    attn = ((q.float() @ k.float().transpose(-2, -1)) * (q.shape[-1] ** -0.5)).softmax(dim=-1)
    o = (attn @ v.float()).to(x.dtype).transpose(1, 2).reshape(Bt, L, -1)
    return F.linear(o, p["to_out.0.weight"], p["to_out.0.bias"])


def zigma_forward(sd, cfg, x, t, y=None):
    """Functional ZigMa.forward (model_zigma.py:911-990) in eval mode (drop_path = identity).

    sd: state dict with the reference's key layout; cfg: dict with
    in_channels, embed_dim, depth, img_dim, patch_size, scan_type, video_frames, use_pe, tpe,
    num_classes, has_text, norm_epsilon.  has_text (model_zigma.py:446-458, 930-933): y is the (B, n_tokens, d_context) text
    embedding; c = t_emb + mean(y_embedder(y)), and every block adds a gated cross-attention over the embedded text."""
    D, depth, p = cfg["embed_dim"], cfg["depth"], cfg.get("patch_size", 1)
    vf = cfg.get("video_frames", 0)
    eps = cfg.get("norm_epsilon", 1e-5)
    dt = sd["x_embedder.proj.weight"].dtype
    if vf > 0:
        Bt, T = x.shape[:2]
        hs = F.conv2d(x.reshape(Bt * T, *x.shape[2:]), sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=p)
        hs = hs.flatten(2).transpose(1, 2).reshape(Bt, -1, D)
    else:
        Bt = x.shape[0]
        hs = F.conv2d(x, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=p).flatten(2).transpose(1, 2)
    L = hs.shape[1]
    tt = (t * 1000.0).to(hs)
    te = timestep_embedding(tt, 256, dt).to(dt)
    te = F.linear(F.silu(F.linear(te, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])),
                  sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    c = te
    text = None
    if cfg.get("has_text", False):
        text = F.linear(y.to(dt), sd["y_embedder.weight"], sd["y_embedder.bias"])           # (B, n_tokens, D)
        c = te + text.mean(dim=1)
    elif cfg.get("num_classes", -1) > 0:
        c = te + sd["y_embedder.embedding_table.weight"][y]
    if cfg.get("use_pe", 0) in (1, 2):
        hs = hs + sd["pos_embed"]
    if vf > 0 and cfg.get("tpe", False):
        K = L // vf
        hs = (hs.reshape(Bt, vf, K, D) + sd["temporal_pos_embedding"].reshape(1, vf, 1, D)).reshape(Bt, L, D)

    num_patches = (cfg["img_dim"] // p) ** 2
    fwd, rev, st = build_scan_tables(cfg["scan_type"], depth, num_patches, vf)
    stype = cfg["scan_type"]
    residual = None
    for i in range(depth):
        pre = f"blocks.{i}."
        hs, residual = BACKEND.get("norm", add_norm)(hs, sd[pre + "norm.weight"], None, residual, prenorm=True,
                                                     residual_in_fp32=True, eps=eps)
        mod = F.linear(F.silu(c), sd[pre + "adaLN_modulation.1.weight"], sd[pre + "adaLN_modulation.1.bias"])
        if text is None:
            shift, scale, gate = mod.chunk(3, dim=1)
        else:
            shift, scale, gate, shift_msa, scale_msa, gate_msa = mod.chunk(6, dim=1)
        mp = {k[len(pre + "mixer."):]: v for k, v in sd.items() if k.startswith(pre + "mixer.")}
        mixed = mamba_mixer(hs * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1), mp, stype,
                            None if fwd is None else fwd[i], None if rev is None else rev[i],
                            None if st is None else st[i], vf)
        hs = hs + gate.unsqueeze(1) * mixed
        if text is not None:      # gated cross-attention branch (model_zigma.py:446-458); norm_msa has no affine parameters
            ap = {k[len(pre + "msa."):]: v for k, v in sd.items() if k.startswith(pre + "msa.")}
            qin = F.layer_norm(hs, (D,), None, None, 1e-6) * (1 + scale_msa.unsqueeze(1)) + shift_msa.unsqueeze(1)
            hs = hs + gate_msa.unsqueeze(1) * cross_attention(qin, text, ap)
    hs = BACKEND.get("norm", add_norm)(hs, sd["norm_f.weight"], None, residual, prenorm=False, residual_in_fp32=True, eps=eps)
    hs = F.layer_norm(hs, (D,), None, None, 1e-6)
    hs = F.linear(hs, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    C = cfg["in_channels"]
    if vf > 0:
        g = int((L // vf) ** 0.5)
        return hs.reshape(Bt, vf, g, g, p, p, C).permute(0, 1, 6, 2, 4, 3, 5).reshape(Bt, vf, C, g * p, g * p)
    g = int(L ** 0.5)
    return hs.reshape(Bt, g, g, p, p, C).permute(0, 5, 1, 3, 2, 4).reshape(Bt, C, g * p, g * p)


# ------------------------------------------------------------------------------------------------
# a11. fixed-grid ODE sampling (velocity model, linear path)
# ------------------------------------------------------------------------------------------------

def sample_ode_fixed(model_fn, x0, num_steps=50, method="euler", t0=0.0, t1=1.0):
    """Fixed-grid integration of dx/dt = model(x, t) on linspace(t0, t1, num_steps) (num_steps-1
    model evaluations for euler).  transport/integrators.py:83-123 hands this grid to
    torchdiffeq.odeint (third-party, unpinned: README.md:181, absent here); fixed-grid 'euler'
    there is x_{i+1} = x_i + (t_{i+1}-t_i) f(t_i, x_i), 'heun' (heun2... torchdiffeq name "heun3"
    differs) is NOT restated.  Returns the final state (the reference keeps [-1]: sample_acc.py:362)."""
    assert method == "euler"
    ts = torch.linspace(t0, t1, num_steps)
    x = x0
    for i in range(num_steps - 1):
        tv = torch.ones(x.shape[0]) * ts[i]
        x = x + (ts[i + 1] - ts[i]) * model_fn(x, tv)
    return x
