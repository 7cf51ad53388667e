"""Particle marginal Metropolis-Hastings (reference: model/PMMH.scala).

The chain logic (propose, accept/reject) is a few scalar operations per iteration and stays on
the host exactly as in the reference; the particle filter it calls -- the `pf: BootstrapFilter`
argument, model/PMMH.scala:58 -- is the GPU one.
"""
import math

import numpy as np

from . import _abi
from .filter import GpuFilterHandle, StateSpace


class MetropState:
    """model/PMMH.scala:26"""

    def __init__(self, ll, params, state, accepted):
        self.ll, self.params, self.state, self.accepted = ll, params, state, accepted


class GpuBootstrapFilter:
    """BootstrapFilter[Parameters, StateSpace[State]] (model/package.scala:24) backed by ONE
    GPU filter handle that is re-parameterised per call (cssm_filter_set_params) instead of being
    rebuilt, with the observations resident on the device.  Equivalent of
        Reader { p => ParticleFilter.filterLlState(data, resample, n)(model.run(p).get) }
    (examples/DetermineParameters.scala:67-72).  Returns (ll, [last sampled state])."""

    def __init__(self, unparamModel, initParams, data, resample, n, dtype=_abi.F32, device=0, seed=0, stream_id=0,
                 precision=None):
        from .resampling import Resampling
        self.unparamModel = unparamModel
        mod = unparamModel(initParams)
        if precision is not None:
            from .model import Model
            mod = Model(mod.leaves, mod.step_mode, precision)
        self.precision = precision
        self.handle = GpuFilterHandle(mod, Resampling.kind_of(resample), n, dtype, device, seed, stream_id)
        t, y, h = GpuFilterHandle._series(data)
        self.t_last = float(t[-1])
        self.handle.load_series(t, y, h)

    def __call__(self, p):
        mod = self.unparamModel(p)
        if self.precision is not None:
            from .model import Model
            mod = Model(mod.leaves, mod.step_mode, self.precision)
        self.handle.set_params(mod)
        ll = self.handle.ll_resident()
        return ll, [StateSpace(self.t_last, self.handle.sample_one())]

    def close(self):
        self.handle.close()


class MetropolisHastings:
    """trait MetropolisHastings, model/PMMH.scala:28-99"""

    def __init__(self, initialParams, proposal, logTransition, prior, pf, rng=None):
        self.initialParams, self.proposal, self.logTransition, self.prior, self.pf = initialParams, proposal, logTransition, prior, pf
        self.rng = rng if rng is not None else np.random.default_rng()
        # model/PMMH.scala:121: ll = -1e99 so that the first proposal is always accepted
        self.init = MetropState(-1e99, initialParams, None, 0)

    def mhStep(self, s):
        """model/PMMH.scala:68-81"""
        propParams = self.proposal(s.params)
        state = self.pf(propParams)
        a = (state[0] + self.logTransition(propParams, s.params) + self.prior(propParams)
             - self.logTransition(s.params, propParams) - s.ll - self.prior(s.params))
        u = self.rng.random()
        if math.log(u) < a:
            return MetropState(state[0], propParams, state[1][-1], s.accepted + 1)
        return s

    def iters(self):
        """model/PMMH.scala:95-98: the chain without its initial state (drop(1))."""
        s = self.init
        while True:
            s = self.mhStep(s)
            yield s

    def params(self):
        for s in self.iters():
            yield (s.ll, s.params, s.accepted)

    @staticmethod
    def pmmhState(initP, proposal, logTransition, prior, rng=None):
        """model/PMMH.scala:161-167: Reader from the bootstrap filter to the stream of states."""
        return lambda pf: ParticleMetropolisHastings(initP, proposal, logTransition, prior, pf, rng).iters()


class ParticleMetropolisHastings(MetropolisHastings):
    """model/PMMH.scala:114-123"""


class ApproxPMMH(MetropolisHastings):
    """model/PMMH.scala:128-153: the likelihood of the CURRENT parameters is re-estimated in every
    iteration (two filter runs per step); on rejection the chain keeps the re-estimated value."""

    def mhStep(self, s):
        propParams = self.proposal(s.params)
        state = self.pf(propParams)
        oldState = self.pf(s.params)
        a = (state[0] + self.logTransition(propParams, s.params) + self.prior(propParams)
             - self.logTransition(s.params, propParams) - oldState[0] - self.prior(s.params))
        u = self.rng.random()
        if math.log(u) < a:
            return MetropState(state[0], propParams, state[1][-1], s.accepted + 1)
        return MetropState(oldState[0], s.params, oldState[1][-1], s.accepted)


def approxPmmh(initP, proposal, logTransition, prior, rng=None):
    """model/PMMH.scala:169-175."""
    return lambda pf: ApproxPMMH(initP, proposal, logTransition, prior, pf, rng).iters()


def pmmhStep(pos, proposal, rng=None):
    """model/PMMH.scala:177-191: one step of a Metropolis chain on (ll, parameters) with a symmetric proposal."""
    rng = rng if rng is not None else np.random.default_rng()

    def step(s):
        prop = proposal(s[1])
        ll = pos(prop)
        return (ll, prop) if math.log(rng.random()) < ll - s[0] else s
    return step


MetropolisHastings.approxPmmh = staticmethod(approxPmmh)
MetropolisHastings.pmmhStep = staticmethod(pmmhStep)
