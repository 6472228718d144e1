"""ODE integration for the sampler: the job the reference hands to ``torchdiffeq.odeint``
(transport/integrators.py:4,121; torchdiffeq is a third-party dependency, unpinned in the
reference -- README.md:181 "pip install torchdiffeq" -- and absent from this image, so its
PUBLISHED algorithms are restated here, for the 0.2.x line).

``odeint(func, y0, t, method=..., rtol=..., atol=...)`` keeps torchdiffeq's calling convention:
``func(t, y)`` with ``t`` a 0-dim tensor, ``y0`` a tensor or a tuple of tensors, ``t`` a 1-D
monotone grid (either direction); the result stacks the solution at every ``t`` along dim 0 (a tuple
of such stacks for a tuple state).

Methods
  fixed grid  : euler, midpoint, rk4 (the 3/8-rule torchdiffeq uses), heun2 (alias heun), heun3
  adaptive    : dopri5 (Dormand-Prince 5(4), FSAL, 4th-order dense output), bosh3, adaptive_heun
The adaptive controller is torchdiffeq's: mixed error norm (max over state tensors of the RMS of
err / (atol + rtol max(|y0|,|y1|))), accept when <= 1, step factor min(10, max(0.9 err^(-1/order),
0.2)) with the lower clamp lifted to 1 after an accepted step, Hairer's initial-step heuristic.
Parity: no torchdiffeq here to pin against -- the tests pin the tableaux through convergence order
and closed-form solutions, and dopri5 against scipy.integrate.solve_ivp(RK45) (same tableau).
"""
import torch

__all__ = ["odeint", "FIXED_METHODS", "ADAPTIVE_METHODS"]


# ---- Butcher tableaux: (c, a-rows, b, b_err (= b - b_hat), c_mid, order) ---------------------------
def _dopri5():
    c = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
    a = [[1 / 5],
         [3 / 40, 9 / 40],
         [44 / 45, -56 / 15, 32 / 9],
         [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
         [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
         [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]]
    b = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0]
    bhat = [1951 / 21600, 0.0, 22642 / 50085, 451 / 720, -12231 / 42400, 649 / 6300, 1 / 60]
    mid = [6025192743 / 30085553152 / 2, 0.0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
           187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]
    return c, a, b, [x - y for x, y in zip(b, bhat)], mid, 5


def _bosh3():
    c = [1 / 2, 3 / 4, 1.0]
    a = [[1 / 2], [0.0, 3 / 4], [2 / 9, 1 / 3, 4 / 9]]
    b = [2 / 9, 1 / 3, 4 / 9, 0.0]
    bhat = [7 / 24, 1 / 4, 1 / 3, 1 / 8]
    return c, a, b, [x - y for x, y in zip(b, bhat)], [0.0, 0.5, 0.0, 0.0], 3


def _adaptive_heun():
    c = [1.0]
    a = [[1.0]]
    b = [0.5, 0.5]
    return c, a, b, [0.5 - 1.0, 0.5 - 0.0], [0.5, 0.0], 2


ADAPTIVE_METHODS = {"dopri5": _dopri5, "bosh3": _bosh3, "adaptive_heun": _adaptive_heun}
FIXED_METHODS = ("euler", "midpoint", "rk4", "heun2", "heun", "heun3")


# ---- tuple-state helpers -------------------------------------------------------------------------------
class _State:
    """A tuple of tensors with the few vector-space ops the solvers need."""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = tuple(v)

    @staticmethod
    def lincomb(coefs, ks, h):
        """h * sum_i coefs[i] * ks[i] (zero coefficients skipped); None entries where the sum is empty."""
        out = []
        for j in range(len(ks[0].v)):
            acc = None
            for cf, k in zip(coefs, ks):
                if cf == 0.0:
                    continue
                term = k.v[j] * cf
                acc = term if acc is None else acc + term
            out.append(None if acc is None else acc * h)
        return out

    def axpy(self, coefs, ks, h):
        """self + h * sum_i coefs[i] * ks[i]."""
        return _State([y if d is None else y + d for y, d in zip(self.v, _State.lincomb(coefs, ks, h))])


def _rms(x):
    return x.float().pow(2).mean().sqrt() if x.numel() else x.new_zeros((), dtype=torch.float32)


def _mixed_norm(vals):
    return max(float(_rms(v)) for v in vals)


def _call(func, t, y, sign, is_tuple, tdtype):
    """func in solver time s = sign * t: dy/ds = sign * func(sign * s, y)."""
    tt = torch.as_tensor(sign * t, dtype=tdtype)
    out = func(tt, y.v if is_tuple else y.v[0])
    out = tuple(out) if is_tuple else (out,)
    if sign < 0:
        out = tuple(-o for o in out)
    return _State(out)


def _per_state(tol, n):
    if isinstance(tol, (list, tuple)):
        if len(tol) == 1 and n > 1:
            tol = list(tol) * n
        assert len(tol) == n, "rtol/atol lists must match the number of state tensors"
        return [float(x) for x in tol]
    return [float(tol)] * n


# ---- fixed-grid solvers -----------------------------------------------------------------------------------
def _fixed_step(method, f, t0, h, y):
    k1 = f(t0, y)
    if method == "euler":
        return y.axpy([1.0], [k1], h)
    if method == "midpoint":
        k2 = f(t0 + 0.5 * h, y.axpy([0.5], [k1], h))
        return y.axpy([1.0], [k2], h)
    if method in ("heun2", "heun"):
        k2 = f(t0 + h, y.axpy([1.0], [k1], h))
        return y.axpy([0.5, 0.5], [k1, k2], h)
    if method == "heun3":
        k2 = f(t0 + h / 3, y.axpy([1 / 3], [k1], h))
        k3 = f(t0 + 2 * h / 3, y.axpy([0.0, 2 / 3], [k1, k2], h))
        return y.axpy([0.25, 0.0, 0.75], [k1, k2, k3], h)
    if method == "rk4":          # Kutta's 3/8 rule
        k2 = f(t0 + h / 3, y.axpy([1 / 3], [k1], h))
        k3 = f(t0 + 2 * h / 3, y.axpy([-1 / 3, 1.0], [k1, k2], h))
        k4 = f(t0 + h, y.axpy([1.0, -1.0, 1.0], [k1, k2, k3], h))
        return y.axpy([1 / 8, 3 / 8, 3 / 8, 1 / 8], [k1, k2, k3, k4], h)
    raise ValueError(method)


# ---- adaptive embedded Runge-Kutta -----------------------------------------------------------------------
def _initial_step(f, t0, y0, f0, order, rtol, atol):
    scale = [a + y.abs() * r for y, a, r in zip(y0.v, atol, rtol)]
    d0 = _mixed_norm([y / s for y, s in zip(y0.v, scale)])
    d1 = _mixed_norm([k / s for k, s in zip(f0.v, scale)])
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = f(t0 + h0, y0.axpy([1.0], [f0], h0))
    d2 = _mixed_norm([(b - a) / s for a, b, s in zip(f0.v, f1.v, scale)]) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / (order + 1))
    return min(100 * h0, h1)


def _quartic(y0, y1, ymid, f0, f1, h):
    """Coefficients of the 4th-degree interpolant matching y0, y1, y(mid), f0, f1."""
    co = []
    for a0, a1, am, g0, g1 in zip(y0.v, y1.v, ymid.v, f0.v, f1.v):
        p4 = 2 * h * (g1 - g0) - 8 * (a1 + a0) + 16 * am
        p3 = h * (5 * g0 - 3 * g1) + 18 * a0 + 14 * a1 - 32 * am
        p2 = h * (g1 - 4 * g0) - 11 * a0 - 5 * a1 + 16 * am
        co.append((p4, p3, p2, h * g0, a0))
    return co


def _eval_quartic(co, x):
    return _State([(((p4 * x + p3) * x + p2) * x + p1) * x + p0 for p4, p3, p2, p1, p0 in co])


def _adaptive(f, y0, ts, tableau, rtol, atol, first_step=None, max_steps=2 ** 31 - 1,
              safety=0.9, ifactor=10.0, dfactor=0.2):
    c, a, b, berr, mid, order = tableau
    t = ts[0]
    y = y0
    fy = f(t, y)
    h = first_step if first_step is not None else _initial_step(f, t, y, fy, order - 1, rtol, atol)
    out = [y0]
    nxt = 1
    interp = None            # (t_lo, t_hi, coeffs) of the last accepted step
    nsteps = 0
    while nxt < len(ts):
        # emit every requested time inside the last accepted step
        if interp is not None and ts[nxt] <= interp[1]:
            lo, hi, co = interp
            out.append(_eval_quartic(co, (ts[nxt] - lo) / (hi - lo)))
            nxt += 1
            continue
        assert nsteps < max_steps, "max_num_steps exceeded"
        nsteps += 1
        assert t + h > t, "underflow in dt"
        ks = [fy]
        for ci, ai in zip(c, a):
            ks.append(f(t + ci * h, y.axpy(ai, ks, h)))
        y1 = y.axpy(b, ks, h)
        err = _State.lincomb(berr, ks, h)
        ratio = _mixed_norm([e / (at + rt * torch.maximum(y_.abs(), n_.abs()))
                             for e, y_, n_, at, rt in zip(err, y.v, y1.v, atol, rtol)])
        accept = ratio <= 1.0
        if accept:
            ymid = y.axpy(mid, ks, h)
            f1 = ks[-1] if b[-1] == 0.0 and a[-1] == b[:-1] else f(t + h, y1)   # FSAL when the tableau allows
            interp = (t, t + h, _quartic(y, y1, ymid, fy, f1, h))
            t, y, fy = t + h, y1, f1
        if ratio == 0.0:
            fac = ifactor
        else:
            fac = min(ifactor, max(safety / ratio ** (1.0 / order), 1.0 if ratio < 1.0 else dfactor))
        h = h * fac
    return out


# ---- public entry point ------------------------------------------------------------------------------------
def odeint(func, y0, t, *, method="dopri5", rtol=1e-7, atol=1e-9, options=None):
    options = options or {}
    is_tuple = isinstance(y0, (tuple, list))
    y = _State(y0 if is_tuple else (y0,))
    n = len(y.v)
    tdtype = t.dtype if torch.is_tensor(t) and t.is_floating_point() else torch.float32
    ts = [float(v) for v in (t.tolist() if torch.is_tensor(t) else t)]
    assert len(ts) >= 1
    inc = all(b > a for a, b in zip(ts, ts[1:]))
    dec = all(b < a for a, b in zip(ts, ts[1:]))
    assert inc or dec or len(ts) == 1, "t must be strictly increasing or decreasing"
    sign = -1.0 if dec and len(ts) > 1 else 1.0
    ss = [sign * v for v in ts]
    f = lambda s, st: _call(func, s, st, sign, is_tuple, tdtype)
    m = method.lower() if isinstance(method, str) else method
    if m in FIXED_METHODS:
        sol = [y]
        for s0, s1 in zip(ss, ss[1:]):
            y = _fixed_step(m, f, s0, s1 - s0, y)
            sol.append(y)
    elif m in ADAPTIVE_METHODS:
        sol = _adaptive(f, y, ss, ADAPTIVE_METHODS[m](), _per_state(rtol, n), _per_state(atol, n),
                        first_step=options.get("first_step"), max_steps=options.get("max_num_steps", 2 ** 31 - 1),
                        safety=options.get("safety", 0.9), ifactor=options.get("ifactor", 10.0),
                        dfactor=options.get("dfactor", 0.2))
    else:
        raise ValueError(f"Invalid method \"{method}\". Must be one of {sorted(list(ADAPTIVE_METHODS) + list(FIXED_METHODS))}")
    stacked = tuple(torch.stack([s.v[j] for s in sol], 0) for j in range(n))
    return stacked if is_tuple else stacked[0]
