"""Flow-matching sampler surface of the reference (``transport/`` of CompVis/zigma, after SiT) --
the part that DRIVES the denoiser hot path: ``create_transport`` (transport/__init__.py:4-75),
``Sampler.sample_ode`` (transport/transport.py:372-417), the ``ode`` driver
(transport/integrators.py:83-123) and the velocity drift (transport/transport.py:175-210).

torchdiffeq (third party, unpinned: README.md:181) is what the reference hands the time grid to;
it is absent here, so the fixed-grid solvers it would run are restated: 'euler'
(x_{i+1} = x_i + dt f(t_i, x_i)) and 'heun' (trapezoidal predictor-corrector, torchdiffeq's
"heun2"... the reference's README only ever uses euler/dopri5).  Adaptive dopri5, the SDE samplers
and the training losses are SURVEY.md section 8f rows ("next").
"""
import enum

import torch


class ModelType(enum.Enum):
    NOISE = enum.auto()
    SCORE = enum.auto()
    VELOCITY = enum.auto()


class PathType(enum.Enum):
    LINEAR = enum.auto()
    GVP = enum.auto()
    VP = enum.auto()


class WeightType(enum.Enum):
    NONE = enum.auto()
    VELOCITY = enum.auto()
    LIKELIHOOD = enum.auto()


class Transport:
    """Velocity-prediction transport on the linear path (the configuration every ZigMa config uses:
    config/*.yaml path_type Linear, prediction velocity)."""

    def __init__(self, *, model_type, path_type, loss_type, train_eps, sample_eps):
        if model_type != ModelType.VELOCITY or path_type != PathType.LINEAR:
            raise NotImplementedError("zigma_b200.transport: only velocity prediction on the Linear path is implemented")
        self.model_type, self.path_type, self.loss_type = model_type, path_type, loss_type
        self.train_eps, self.sample_eps = train_eps, sample_eps

    def check_interval(self, train_eps, sample_eps, *, diffusion_form="SBDM", sde=False, reverse=False,
                       eval=False, last_step_size=0.0):
        """transport/transport.py:79-112: for velocity + Linear the interval is [0, 1]."""
        t0, t1 = 0.0, 1.0
        if reverse:
            t0, t1 = 1 - t0, 1 - t1
        return t0, t1

    def get_drift(self):
        """transport/transport.py:175-210 (velocity_ode + the output-shape assert of body_fn)."""
        def body_fn(x, t, model, **model_kwargs):
            out = model(x, t, **model_kwargs)
            assert out.shape == x.shape, "Output shape from ODE solver must match input shape"
            return out
        return body_fn

    def training_losses(self, model, x1, model_kwargs=None):
        """transport/transport.py:126-173 for velocity / Linear / no weighting: t ~ U(0,1), x0 ~ N,
        xt = t x1 + (1 - t) x0, target ut = x1 - x0, per-sample mean squared error."""
        model_kwargs = model_kwargs or {}
        x0 = torch.randn_like(x1)
        t = torch.rand((x1.shape[0],), device=x1.device, dtype=x1.dtype)
        t = t * (1 - self.train_eps - self.sample_eps) + self.train_eps if (self.train_eps or self.sample_eps) else t
        tb = t.view(-1, *([1] * (x1.dim() - 1)))
        xt = tb * x1 + (1 - tb) * x0
        ut = x1 - x0
        pred = model(xt, t, **model_kwargs)
        return {"pred": pred, "loss": ((pred - ut) ** 2).flatten(1).mean(1)}


class ode:
    """Fixed-grid ODE driver (transport/integrators.py:83-123)."""

    def __init__(self, drift, *, t0, t1, sampler_type, num_steps, atol, rtol):
        self.drift = drift
        self.t = torch.linspace(t0, t1, num_steps)
        self.atol, self.rtol, self.sampler_type = atol, rtol, sampler_type

    def sample(self, x, model, **model_kwargs):
        dev = x.device
        method = self.sampler_type.lower()
        if method not in ("euler", "heun"):
            raise NotImplementedError(f"sampler_type {self.sampler_type}: only the fixed-grid 'euler' and 'heun' are implemented (no torchdiffeq)")
        ts = self.t.tolist()
        fn = lambda tt, xx: self.drift(xx, torch.ones(xx.size(0), device=dev) * tt, model, **model_kwargs)
        samples = [x]
        for i in range(len(ts) - 1):
            dt = ts[i + 1] - ts[i]
            k1 = fn(ts[i], x)
            if method == "euler":
                x = x + dt * k1
            else:
                k2 = fn(ts[i + 1], x + dt * k1)
                x = x + (0.5 * dt) * (k1 + k2)
            samples.append(x)
        return samples          # indexable like odeint's (T, ...) result; callers take [-1]


class Sampler:
    """transport/transport.py:236-250,372-417."""

    def __init__(self, transport):
        self.transport = transport
        self.drift = transport.get_drift()

    def sample_ode(self, *, sampling_method="dopri5", num_steps=50, atol=1e-6, rtol=1e-3, reverse=False):
        if reverse:
            drift = lambda x, t, model, **kw: self.drift(x, torch.ones_like(t) * (1 - t), model, **kw)
        else:
            drift = self.drift
        t0, t1 = self.transport.check_interval(self.transport.train_eps, self.transport.sample_eps, sde=False,
                                               eval=True, reverse=reverse, last_step_size=0.0)
        return ode(drift=drift, t0=t0, t1=t1, sampler_type=sampling_method, num_steps=num_steps, atol=atol, rtol=rtol).sample


def create_transport(path_type="Linear", prediction="velocity", loss_weight=None, train_eps=None, sample_eps=None):
    """transport/__init__.py:4-75."""
    model_type = {"noise": ModelType.NOISE, "score": ModelType.SCORE, "velocity": ModelType.VELOCITY}.get(prediction)
    if model_type is None:
        raise ValueError(f"Model type {prediction} not implemented")
    loss_type = {"velocity": WeightType.VELOCITY, "likelihood": WeightType.LIKELIHOOD, None: WeightType.NONE}.get(loss_weight, "bad")
    if loss_type == "bad":
        raise ValueError(f"Loss type {loss_weight} not implemented")
    ptype = {"Linear": PathType.LINEAR, "GVP": PathType.GVP, "VP": PathType.VP}[path_type]
    if ptype == PathType.VP:
        train_eps = 1e-5 if train_eps is None else train_eps
        sample_eps = 1e-3 if train_eps is None else sample_eps
    elif model_type != ModelType.VELOCITY:
        train_eps = 1e-3 if train_eps is None else train_eps
        sample_eps = 1e-3 if train_eps is None else sample_eps
    else:   # velocity & [GVP, LINEAR] is stable everywhere
        train_eps, sample_eps = 0, 0
    return Transport(model_type=model_type, path_type=ptype, loss_type=loss_type, train_eps=train_eps, sample_eps=sample_eps)
