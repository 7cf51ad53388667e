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
                 precision=None, replicas=1):
        """`replicas` > 1: that many handles (each with its own CUDA stream and Philox stream id) so that `many`
        can evaluate several parameter values at once -- the two filter runs of an ApproxPMMH step, or the proposals of
        several chains.  A small cloud leaves most of the GPU idle (a 2^16-particle filter is latency bound), so
        concurrent evaluations cost little more than one."""
        from .resampling import Resampling
        self.unparamModel = unparamModel
        mod = unparamModel(initParams)
        if precision is not None:
            from .model import Model
            mod = Model(mod.leaves, mod.step_mode, precision)
        self.precision = precision
        t, y, h = GpuFilterHandle._series(data)
        self.t_last = float(t[-1])
        self.handles = []
        for r in range(max(1, int(replicas))):
            hd = GpuFilterHandle(mod, Resampling.kind_of(resample), n, dtype, device, seed, stream_id + r)
            hd.load_series(t, y, h)
            self.handles.append(hd)
        self.handle = self.handles[0]
        self._pool = None

    def _model(self, p):
        mod = self.unparamModel(p)
        if self.precision is not None:
            from .model import Model
            mod = Model(mod.leaves, mod.step_mode, self.precision)
        return mod

    def _eval(self, handle, p):
        handle.set_params(self._model(p))
        ll = handle.ll_resident()
        return ll, [StateSpace(self.t_last, handle.sample_one())]

    def __call__(self, p):
        return self._eval(self.handle, p)

    def many(self, ps):
        """[pf(p) for p in ps], evaluated concurrently on the replica handles (one host thread each; the C calls
        release the GIL and every handle has its own stream)."""
        ps = list(ps)
        if len(ps) > len(self.handles):
            raise ValueError(f"{len(ps)} parameter values for {len(self.handles)} replica handle(s)")
        if len(ps) == 1:
            return [self._eval(self.handles[0], ps[0])]
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=len(self.handles))
        futs = [self._pool.submit(self._eval, h, p) for h, p in zip(self.handles, ps)]
        return [f.result() for f in futs]

    def close(self):
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None
        for h in self.handles:
            h.close()


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

    def markovIters(self):
        """model/PMMH.scala:85-87 with model/MarkovChain.scala:7-17 and Breeze's Process.steps: the first element is
        already ONE mhStep past `init` (MarkovChain.draw = resample(init).draw), every further one a step later."""
        s = self.init
        while True:
            s = self.mhStep(s)
            yield s

    def iters(self):
        """model/PMMH.scala:95-98: `markovIters.steps.drop(1)` -- the first state emitted is the result of the SECOND
        mhStep (the first, always-accepted move away from the artificial ll = -1e99 start is dropped)."""
        it = self.markovIters()
        next(it)
        return it

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
        if hasattr(self.pf, "many") and len(getattr(self.pf, "handles", ())) >= 2:
            state, oldState = self.pf.many([propParams, s.params])  # the two filter runs of a step, side by side
        else:
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


def runChains(chains, n_iters, parallelism=2):
    """examples/DetermineParameters.scala:68-80: `Source(1 to k).mapAsync(2) { chain => iters.take(n).runWith(sink) }` --
    k independent chains, `parallelism` of them in flight at once (the reference's 2).  `chains`: iterators of MetropState
    (MetropolisHastings.iters() over distinct GpuBootstrapFilters).  One host thread per running chain; the library calls
    release the GIL, every handle owns its stream, and a 2^16-particle likelihood evaluation occupies a fraction of the
    GPU, so chains on one device overlap.  Returns one list of n_iters states per chain, in input order."""
    from concurrent.futures import ThreadPoolExecutor
    chains = list(chains)

    def run(it):
        return [next(it) for _ in range(int(n_iters))]

    with ThreadPoolExecutor(max_workers=max(1, min(int(parallelism), len(chains)))) as pool:
        return list(pool.map(run, chains))


MetropolisHastings.approxPmmh = staticmethod(approxPmmh)
MetropolisHastings.pmmhStep = staticmethod(pmmhStep)
