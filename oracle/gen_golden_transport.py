"""TEST INFRASTRUCTURE ONLY -- golden vectors for the sampler surface.

Imports the UNMODIFIED reference ``transport/`` package from /root/reference (build container only;
``torchdiffeq`` -- absent here -- is stubbed with a module whose ``odeint`` raises, so only the paths
that do not need it run: check_interval, the plans' algebra, training_losses, get_drift/get_score,
and the SDE samplers).  Writes tests/golden/transport.npz; tests/test_transport.py replays the same
seeded calls through zigma_b200.transport and compares.

    python oracle/gen_golden_transport.py
"""
import importlib
import itertools
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_ROOT = os.environ.get("ZIGMA_REFERENCE_ROOT", "/root/reference")


def load_reference_transport():
    if "torchdiffeq" not in sys.modules:
        stub = types.ModuleType("torchdiffeq")
        def odeint(*a, **k):
            raise RuntimeError("torchdiffeq is not installed in this image")
        stub.odeint = odeint
        sys.modules["torchdiffeq"] = stub
    sys.path.insert(0, REF_ROOT)
    try:
        for k in [k for k in sys.modules if k == "transport" or k.startswith("transport.")]:
            del sys.modules[k]
        mod = importlib.import_module("transport")
        assert os.path.realpath(mod.__file__).startswith(os.path.realpath(REF_ROOT)), mod.__file__
    finally:
        sys.path.remove(REF_ROOT)
    return mod


# ---- the shared, deterministic toy problem (also imported by tests/test_transport.py) ------------------
def toy_model(dim=6, seed=11):
    W = torch.from_numpy(np.random.RandomState(seed).randn(dim, dim).astype(np.float32)) * 0.4
    def model(x, t, **kw):
        scale = kw.get("scale", 1.0)
        return torch.tanh(x @ W) * (0.5 + t.view(-1, *([1] * (x.dim() - 1)))) * scale - 0.1 * x
    return model


def toy_x(bs=5, dim=6, seed=5):
    return torch.from_numpy(np.random.RandomState(seed).randn(bs, dim).astype(np.float32))


INTERVAL_CASES = [dict(diffusion_form=f, sde=s, reverse=r, eval=e, last_step_size=l)
                  for f, s, r, e, l in itertools.product(["SBDM", "sigma"], [False, True], [False, True],
                                                         [False, True], [0.0, 0.04])]
TRANSPORTS = [("Linear", "velocity", None), ("Linear", "noise", None), ("Linear", "score", "velocity"),
              ("GVP", "velocity", None), ("GVP", "score", "likelihood"), ("GVP", "noise", "velocity"),
              ("VP", "velocity", None), ("VP", "noise", "likelihood"), ("VP", "score", None)]
SDE_CASES = [("Euler", "SBDM", 1.0, "Mean"), ("Euler", "sigma", 0.7, "Tweedie"), ("Euler", "linear", 1.0, "Euler"),
             ("Heun", "SBDM", 1.0, None), ("Heun", "decreasing", 1.0, "Mean"),
             ("Heun", "inccreasing-decreasing", 0.5, "Euler")]
EPS = dict(train_eps=2e-3, sample_eps=3e-3)      # explicit: the reference's defaults leave sample_eps None


def run_all(tp):
    """Every pinned quantity, from a module with the reference's transport API."""
    out = {}
    model, x = toy_model(), toy_x()
    for ti, (pt, pred, lw) in enumerate(TRANSPORTS):
        kw = {} if (pred == "velocity" and pt != "VP") else EPS
        tr = tp.create_transport(pt, pred, lw, **kw)
        tag = f"{pt}_{pred}_{lw}"
        out[f"eps_{tag}"] = np.array([tr.train_eps, tr.sample_eps], dtype=np.float64)
        out[f"interval_{tag}"] = np.array([tr.check_interval(tr.train_eps, tr.sample_eps, **c) for c in INTERVAL_CASES],
                                          dtype=np.float64)
        t = torch.linspace(0.1, 0.9, x.size(0))
        ps = tr.path_sampler
        out[f"drift_{tag}"] = tr.get_drift()(x, t, model, scale=1.3).numpy()
        out[f"score_{tag}"] = tr.get_score()(x, t, model).numpy()
        dm, dv = ps.compute_drift(x, t)
        out[f"plan_drift_{tag}"] = np.stack([dm.numpy(), (dv + 0 * x).numpy()])
        v = model(x, t)
        out[f"plan_conv_{tag}"] = np.stack([ps.get_score_from_velocity(v, x, t).numpy(),
                                            ps.get_noise_from_velocity(v, x, t).numpy(),
                                            ps.get_velocity_from_score(v, x, t).numpy()])
        x0 = toy_x(seed=6)
        _, xt, ut = ps.plan(t, x0, x)
        out[f"plan_xt_ut_{tag}"] = np.stack([xt.numpy(), (ut + 0 * x).numpy()])
        torch.manual_seed(100 + ti)
        terms = tr.training_losses(model, x, dict(scale=0.9))
        out[f"loss_{tag}"] = terms["loss"].detach().numpy()
        out[f"pred_{tag}"] = terms["pred"].detach().numpy()
        if ti in (0, 2, 3, 7):
            for ci, (meth, form, norm, last) in enumerate(SDE_CASES):
                fn = tp.Sampler(tr).sample_sde(sampling_method=meth, diffusion_form=form, diffusion_norm=norm,
                                               last_step=last, last_step_size=0.04, num_steps=12)
                torch.manual_seed(7 + ci)
                xs = fn(x, model, scale=1.1)
                out[f"sde_{tag}_{ci}"] = np.stack([xs[0].numpy(), xs[5].numpy(), xs[-2].numpy(), xs[-1].numpy()])
    return out


def main():
    ref = run_all(load_reference_transport())
    path = os.path.join(ROOT, "tests", "golden", "transport.npz")
    np.savez_compressed(path, **ref)
    print(f"wrote {path}: {len(ref)} arrays, {os.path.getsize(path)} bytes")
    # pin the restatement right here too
    sys.path.insert(0, ROOT)
    from zigma_b200 import transport as mine
    got = run_all(mine)
    worst = 0.0
    for k, v in ref.items():
        assert got[k].shape == v.shape, k
        assert np.array_equal(np.isfinite(got[k]), np.isfinite(v)), k     # (SBDM from t0=0 is inf/nan in the reference too)
        fin = np.isfinite(v)
        d = np.abs(got[k][fin] - v[fin]).max() if fin.any() else 0.0
        worst = max(worst, float(d / (1e-30 + np.abs(v[fin]).max())) if fin.any() else 0.0)
        assert np.allclose(got[k], v, rtol=1e-6, atol=1e-7, equal_nan=True), (k, d)
    print(f"zigma_b200.transport matches the reference on {len(ref)} arrays (worst scaled diff {worst:.2e})")


if __name__ == "__main__":
    main()
