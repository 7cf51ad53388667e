"""Synthetic observations from a parameterised model (reference: SimulateData.simMarkov /
simStep, model/Data.scala:81-100,186-193).  Host side, one latent path, numpy RNG: this is input
generation for examples and bench.py, not part of the filter path.
"""
import math

import numpy as np

from . import _abi


def _step_exact(sde, x, dt, z):
    if sde.kind == _abi.SDE_BROWNIAN:
        return np.sqrt(sde.sigma * dt) * z + x
    if sde.kind == _abi.SDE_GEN_BROWNIAN:
        return np.sqrt(sde.sigma * dt) * z + (x + sde.mu * dt)
    var = (sde.sigma * sde.sigma / (sde.phi * 2.0)) * (1.0 - np.exp(sde.phi * -2.0 * dt))
    return np.sqrt(var) * z + (sde.mu + (x - sde.mu) * np.exp(-sde.phi * dt))


def _observe(mod, gamma, rng):
    k = mod.obs_kind
    if k == _abi.OBS_POISSON:
        return float(rng.poisson(math.exp(gamma)))
    if k == _abi.OBS_NEGBIN:
        size, mu = math.exp(mod.scale), math.exp(gamma)
        prob = mu / (size + mu)
        return float(rng.poisson(rng.gamma(size, prob / (1 - prob))))
    if k == _abi.OBS_NORMAL:
        return float(gamma + math.exp(mod.scale) * rng.standard_normal())
    if k == _abi.OBS_BERNOULLI:
        return 1.0 if rng.random() < mod.link(gamma) else 0.0
    if k == _abi.OBS_STUDENT_T:  # StudentsT(df) * v + x, model/Model.scala:145-150
        return float(rng.standard_t(mod.df) * math.exp(mod.scale) + gamma)
    if k == _abi.OBS_ZIP:  # model/Model.scala:282-292
        p = math.exp(mod.scale) / (1 + math.exp(mod.scale))
        return 0.0 if rng.random() < p else float(rng.poisson(math.exp(gamma)))
    if k == _abi.OBS_BETA:  # new Beta(link(gamma), beta), model/Model.scala:340-343 (the scale is the second shape)
        return float(min(max(rng.beta(mod.link(gamma), mod.scale), 1e-12), 1 - 1e-12))
    return 1.0


def simRegular(mod, dt, T, seed=1):
    """T observations on the grid t = 0, dt, 2 dt, ... (SimulateData.observations uses dt = 0.1).
    Returns (t[T], y[T], x[T, d])."""
    rng = np.random.default_rng(seed)
    xs = [l.sde.m0 + np.sqrt(l.sde.c0) * rng.standard_normal(l.sde.dimension) for l in mod.leaves]
    t = 0.0
    ts, ys, states = [], [], []
    for s in range(T):
        if s > 0:
            xs = [_step_exact(l.sde, x, dt, rng.standard_normal(l.sde.dimension)) for l, x in zip(mod.leaves, xs)]
            t = t + dt
        flat = np.concatenate(xs)
        ts.append(t)
        ys.append(_observe(mod, mod.f(flat, t), rng))
        states.append(flat)
    return np.array(ts), np.array(ys), np.array(states)


def simLgcpEvents(T, mean_gap=0.1, seed=1):
    """Event times with exponential gaps (synthetic LGCP input: every datum is an event)."""
    rng = np.random.default_rng(seed)
    return np.cumsum(rng.exponential(mean_gap, T)), np.ones(T)
