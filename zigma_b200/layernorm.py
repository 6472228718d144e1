"""Host-side mirror of ``dis_mamba/mamba_ssm/ops/triton/layernorm.py`` (reference) over hand-written
sm_100a kernels -- the reference launches Triton here; north_star forbids Triton on the hot path.
``rms_norm_fn`` (:477-478), ``layer_norm_fn`` (:464-474), ``RMSNorm`` (:481-503), ``LayerNormFn``
(:380-461)."""
import torch

from . import _lib


def _norm_fwd(x, weight, bias, eps, residual, residual_dtype, is_rms):
    """_layer_norm_fwd (layernorm.py:123-177) on a 2-D view.  Returns y, mean, rstd, residual_out."""
    M, N = x.shape
    y = torch.empty_like(x, memory_format=torch.contiguous_format)
    if residual is not None:
        residual_dtype = residual.dtype
    store_res = residual is not None or (residual_dtype is not None and residual_dtype != x.dtype)
    res_out = torch.empty((M, N), dtype=residual_dtype, device=x.device) if store_res else None
    mean = torch.empty((M,), dtype=torch.float32, device=x.device) if not is_rms else None
    rstd = torch.empty((M,), dtype=torch.float32, device=x.device)
    p = _lib.NormParams()
    p.x, p.residual, p.weight, p.bias = _lib.ptr(x), _lib.ptr(residual), _lib.ptr(weight), _lib.ptr(bias)
    p.y, p.residual_out, p.mean, p.rstd = _lib.ptr(y), _lib.ptr(res_out), _lib.ptr(mean), _lib.ptr(rstd)
    p.x_rs, p.y_rs = x.stride(0), y.stride(0)
    p.res_rs = residual.stride(0) if residual is not None else 0
    p.resout_rs = res_out.stride(0) if res_out is not None else 0
    p.nrows, p.ncols = M, N
    p.dtype = _lib.dt(x)
    p.res_dtype = _lib.dt(residual_dtype) if store_res else p.dtype
    p.wdtype = _lib.dt(weight) if weight is not None else p.dtype
    p.is_rms, p.eps = int(is_rms), float(eps)
    _lib.call("zg_add_norm_fwd", p)
    return y, mean, rstd, (res_out if res_out is not None else x)


def _norm_bwd(dy, x, weight, bias, mean, rstd, dresidual, has_residual, is_rms, x_dtype):
    """_layer_norm_bwd (layernorm.py:293-377)."""
    M, N = x.shape
    dx = torch.empty((M, N), dtype=x_dtype, device=x.device)
    dres_in = torch.empty_like(x) if (has_residual and dx.dtype != x.dtype) else None
    dw = torch.zeros((N,), dtype=torch.float32, device=x.device) if weight is not None else None
    db = torch.zeros((N,), dtype=torch.float32, device=x.device) if bias is not None else None
    p = _lib.NormBwdParams()
    p.dy, p.dresidual, p.x, p.weight = _lib.ptr(dy), _lib.ptr(dresidual), _lib.ptr(x), _lib.ptr(weight)
    p.mean, p.rstd, p.dx, p.dresidual_in = _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(dx), _lib.ptr(dres_in)
    p.dweight, p.dbias = _lib.ptr(dw), _lib.ptr(db)
    p.dy_rs, p.x_rs, p.dx_rs = dy.stride(0), x.stride(0), dx.stride(0)
    p.dres_rs = dresidual.stride(0) if dresidual is not None else 0
    p.dresin_rs = dres_in.stride(0) if dres_in is not None else 0
    p.nrows, p.ncols = M, N
    p.dtype, p.res_dtype = _lib.dt(x_dtype), _lib.dt(x)
    p.wdtype = _lib.dt(weight) if weight is not None else p.dtype
    p.is_rms = int(is_rms)
    if dy.dtype != x_dtype:
        dy = dy.to(x_dtype)
        p.dy, p.dy_rs = _lib.ptr(dy), dy.stride(0)
    _lib.call("zg_add_norm_bwd", p)
    if has_residual and dx.dtype == x.dtype:
        dres_in = dx
    return dx, (dw.to(weight.dtype) if dw is not None else None), (db.to(bias.dtype) if db is not None else None), dres_in


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual=None, eps=1e-6, prenorm=False, residual_in_fp32=False,
                is_rms_norm=False):
        _lib.require_cuda(x, weight, bias, residual)
        x_shape_og = x.shape
        x = x.reshape(-1, x.shape[-1])
        if x.stride(-1) != 1:
            x = x.contiguous()
        if residual is not None:
            assert residual.shape == x_shape_og
            residual = residual.reshape(-1, residual.shape[-1])
            if residual.stride(-1) != 1:
                residual = residual.contiguous()
        weight = weight.contiguous()
        if bias is not None:
            bias = bias.contiguous()
        residual_dtype = residual.dtype if residual is not None else (torch.float32 if residual_in_fp32 else None)
        y, mean, rstd, residual_out = _norm_fwd(x, weight, bias, eps, residual, residual_dtype, is_rms_norm)
        ctx.save_for_backward(residual_out, weight, bias, mean, rstd)
        ctx.x_shape_og = x_shape_og
        ctx.eps = eps
        ctx.is_rms_norm = is_rms_norm
        ctx.has_residual = residual is not None
        ctx.prenorm = prenorm
        ctx.x_dtype = x.dtype
        y = y.reshape(x_shape_og)
        return y if not prenorm else (y, residual_out.reshape(x_shape_og))

    @staticmethod
    def backward(ctx, dy, *args):
        x, weight, bias, mean, rstd = ctx.saved_tensors
        dy = dy.reshape(-1, dy.shape[-1])
        if dy.stride(-1) != 1:
            dy = dy.contiguous()
        assert dy.shape == x.shape
        dresidual = None
        if ctx.prenorm:
            dresidual = args[0].reshape(-1, args[0].shape[-1])
            if dresidual.stride(-1) != 1:
                dresidual = dresidual.contiguous()
            if dresidual.dtype != x.dtype:
                dresidual = dresidual.to(x.dtype)
        dx, dw, db, dresidual_in = _norm_bwd(dy, x, weight, bias, mean, rstd, dresidual, ctx.has_residual,
                                             ctx.is_rms_norm, ctx.x_dtype)
        return (dx.reshape(ctx.x_shape_og), dw, db,
                dresidual_in.reshape(ctx.x_shape_og) if ctx.has_residual else None, None, None, None, None)


def layer_norm_fn(x, weight, bias, residual=None, eps=1e-6, prenorm=False, residual_in_fp32=False,
                  is_rms_norm=False):
    return LayerNormFn.apply(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms_norm)


def rms_norm_fn(x, weight, bias, residual=None, prenorm=False, residual_in_fp32=False, eps=1e-6):
    return LayerNormFn.apply(x, weight, bias, residual, eps, prenorm, residual_in_fp32, True)


class RMSNorm(torch.nn.Module):
    def __init__(self, hidden_size, eps=1e-5, device=None, dtype=None):
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.empty(hidden_size, **factory_kwargs))
        self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.ones_(self.weight)

    def forward(self, x, residual=None, prenorm=False, residual_in_fp32=False):
        return rms_norm_fn(x, self.weight, self.bias, residual=residual, eps=self.eps, prenorm=prenorm,
                           residual_in_fp32=residual_in_fp32)
