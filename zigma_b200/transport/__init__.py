"""Flow-matching transport surface of the reference (``transport/`` of CompVis/zigma, after SiT):
``create_transport`` (transport/__init__.py:4-75), ``Transport`` / ``Sampler``
(transport/transport.py), the ``sde`` / ``ode`` drivers (transport/integrators.py) and the
coupling plans (transport/path.py, here ``zigma_b200.transport.path``).

This is host logic around the denoiser hot path: every ``model(x, t, **kw)`` it issues is one
ZigMaEngine step on the GPU.  torchdiffeq (third party, unpinned: README.md:181) is what the
reference hands the ODE to; it is absent here, so its solvers are restated in ``_odeint``.
"""
from . import _plans as path
from . import _odeint
from ._odeint import odeint
from ._sampler import (EasyDict, ModelType, PathType, Sampler, Transport, WeightType, mean_flat, ode, sde)

__all__ = ["create_transport", "Transport", "Sampler", "ModelType", "PathType", "WeightType", "ode", "sde",
           "odeint", "path", "mean_flat", "EasyDict"]


def create_transport(path_type="Linear", prediction="velocity", loss_weight=None, train_eps=None, sample_eps=None):
    """transport/__init__.py:4-75, including its epsilon defaults: the reference tests ``train_eps is
    None`` AFTER assigning train_eps, so a ``sample_eps`` left at None stays None (and the velocity /
    Linear-or-GVP case forces both to 0)."""
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
