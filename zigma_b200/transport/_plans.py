"""Interpolant ("coupling plan") algebra x_t = alpha_t x1 + sigma_t x0 of the reference's
``transport/path.py`` (after SiT): the Linear plan every ZigMa config uses (path.py:18-136), plus the
GVP (path.py:173-192) and VP (path.py:139-170) variants ``create_transport`` can select.

Same class and method names as the reference so ``transport.path_sampler.<method>`` call sites keep
working; each schedule is given once as a ``_coef`` pair and the derived quantities are shared.
"""
import math

import torch


def expand_t_like_x(t, x):
    """(B,) -> (B, 1, ..., 1) broadcastable against x (path.py:5-13)."""
    return t.view(t.size(0), *([1] * (x.dim() - 1)))


class ICPlan:
    """Linear plan: alpha_t = t, sigma_t = 1 - t."""

    def __init__(self, sigma=0.0):
        self.sigma = sigma

    # -- schedule (value, time-derivative) ---------------------------------------------------------
    def compute_alpha_t(self, t):
        return t, 1

    def compute_sigma_t(self, t):
        return 1 - t, -1

    def compute_d_alpha_alpha_ratio_t(self, t):
        return 1 / t

    # -- SDE pieces -----------------------------------------------------------------------------------
    def compute_drift(self, x, t):
        """(-f_t x, beta_t) of the score-parametrised SDE (path.py:35-43)."""
        t = expand_t_like_x(t, x)
        ratio = self.compute_d_alpha_alpha_ratio_t(t)
        sig, dsig = self.compute_sigma_t(t)
        return -(ratio * x), ratio * (sig ** 2) - sig * dsig

    def compute_diffusion(self, x, t, form="constant", norm=1.0):
        """w_t of SiT eq. 4 in the named functional form (path.py:45-68)."""
        t = expand_t_like_x(t, x)
        if form == "constant":
            return norm
        if form == "SBDM":
            return norm * self.compute_drift(x, t)[1]
        if form == "sigma":
            return norm * self.compute_sigma_t(t)[0]
        if form == "linear":
            return norm * (1 - t)
        if form == "decreasing":
            return 0.25 * (norm * torch.cos(math.pi * t) + 1) ** 2
        if form == "inccreasing-decreasing":       # (sic) the reference's key, path.py:61
            return norm * torch.sin(math.pi * t) ** 2
        raise NotImplementedError(f"Diffusion form {form} not implemented")

    # -- model-output conversions -----------------------------------------------------------------------
    def _ratio_terms(self, x, t):
        t = expand_t_like_x(t, x)
        al, dal = self.compute_alpha_t(t)
        sig, dsig = self.compute_sigma_t(t)
        return al / dal, sig, dsig

    def get_score_from_velocity(self, velocity, x, t):
        """SiT eq. 9 (path.py:70-84)."""
        rar, sig, dsig = self._ratio_terms(x, t)
        var = sig ** 2 - rar * dsig * sig
        return (rar * velocity - x) / var

    def get_noise_from_velocity(self, velocity, x, t):
        """path.py:86-100."""
        rar, sig, dsig = self._ratio_terms(x, t)
        var = rar * dsig - sig
        return (rar * velocity - x) / var

    def get_velocity_from_score(self, score, x, t):
        """path.py:102-112."""
        t = expand_t_like_x(t, x)
        drift, var = self.compute_drift(x, t)
        return var * score - drift

    # -- sampling the path (training) ---------------------------------------------------------------------
    def compute_mu_t(self, t, x0, x1):
        t = expand_t_like_x(t, x1)
        return self.compute_alpha_t(t)[0] * x1 + self.compute_sigma_t(t)[0] * x0

    def compute_xt(self, t, x0, x1):
        return self.compute_mu_t(t, x0, x1)

    def compute_ut(self, t, x0, x1, xt):
        t = expand_t_like_x(t, x1)
        return self.compute_alpha_t(t)[1] * x1 + self.compute_sigma_t(t)[1] * x0

    def plan(self, t, x0, x1):
        xt = self.compute_xt(t, x0, x1)
        return t, xt, self.compute_ut(t, x0, x1, xt)


class VPCPlan(ICPlan):
    """Variance-preserving plan with a linear beta schedule (path.py:139-170)."""

    def __init__(self, sigma_min=0.1, sigma_max=20.0):
        self.sigma_min, self.sigma_max = sigma_min, sigma_max

    def log_mean_coeff(self, t):
        return -0.25 * ((1 - t) ** 2) * (self.sigma_max - self.sigma_min) - 0.5 * (1 - t) * self.sigma_min

    def d_log_mean_coeff(self, t):
        return 0.5 * (1 - t) * (self.sigma_max - self.sigma_min) + 0.5 * self.sigma_min

    def compute_alpha_t(self, t):
        al = torch.exp(self.log_mean_coeff(t))
        return al, al * self.d_log_mean_coeff(t)

    def compute_sigma_t(self, t):
        p = 2 * self.log_mean_coeff(t)
        sig = torch.sqrt(1 - torch.exp(p))
        return sig, torch.exp(p) * (2 * self.d_log_mean_coeff(t)) / (-2 * sig)

    def compute_d_alpha_alpha_ratio_t(self, t):
        return self.d_log_mean_coeff(t)

    def compute_drift(self, x, t):
        t = expand_t_like_x(t, x)
        beta = self.sigma_min + (1 - t) * (self.sigma_max - self.sigma_min)
        return -0.5 * beta * x, beta / 2


class GVPCPlan(ICPlan):
    """Trigonometric (generalised VP) plan (path.py:173-192)."""

    def compute_alpha_t(self, t):
        return torch.sin(t * math.pi / 2), math.pi / 2 * torch.cos(t * math.pi / 2)

    def compute_sigma_t(self, t):
        return torch.cos(t * math.pi / 2), -math.pi / 2 * torch.sin(t * math.pi / 2)

    def compute_d_alpha_alpha_ratio_t(self, t):
        return math.pi / (2 * torch.tan(t * math.pi / 2))
