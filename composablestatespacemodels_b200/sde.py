"""SDE parameters and latent-state constructors (reference: model/SdeParameters.scala,
model/Sde.scala).  Host side only: these objects carry the EFFECTIVE parameter vectors the device
kernels consume; the arithmetic of the transitions lives in csrc/ (and, as the checker, oracle/).
"""
import math

import numpy as np

from . import _abi


def _vec(x):
    return np.atleast_1d(np.asarray(x, dtype=np.float64)).copy()


class SdeParameter:
    """model/SdeParameters.scala:14-48.  Values are the UNCONSTRAINED ones the reference stores."""

    fields = ()

    def flatten(self):
        return np.concatenate([getattr(self, f) for f in self.fields])

    @property
    def length(self):
        return self.flatten().size

    def add(self, delta):
        """model/SdeParameters.scala:62-69,104-110,141-149: add a flat vector field by field."""
        delta = _vec(delta)
        out, o = [], 0
        for f in self.fields:
            v = getattr(self, f)
            out.append(v + delta[o:o + v.size])
            o += v.size
        return type(self)(*out)

    def plus(self, that):
        if type(that) is not type(self):
            raise Exception(f"Can't add {type(self).__name__} to {that}")
        return type(self)(*[getattr(self, f) + getattr(that, f) for f in self.fields])

    def map(self, f):
        return type(self)(*[f(getattr(self, n)) for n in self.fields])

    def perturb(self, delta, rng):
        """model/SdeParameters.scala:21-26 (innovation sd = 1/sqrt(delta), as written there)."""
        n = self.length
        return self.add((1.0 / math.sqrt(delta)) * rng.standard_normal(n))

    def __repr__(self):
        return type(self).__name__ + "(" + ", ".join(f"{n}={getattr(self, n).tolist()}" for n in self.fields) + ")"

    # smart constructors (model/SdeParameters.scala:176-205)
    @staticmethod
    def genBrownianParameterUnconstrained(m0, c0, mu, sigma):
        return GenBrownianParameter(m0, c0, mu, sigma)

    @staticmethod
    def brownianParameterUnconstrained(m0, c0, sigma):
        return BrownianParameter(m0, c0, sigma)

    @staticmethod
    def ouParameterUnconstrained(m0, c0, phi, mu, sigma):
        return OuParameter(m0, c0, phi, mu, sigma)

    @staticmethod
    def genBrownianParameter(m0, c0, mu, sigma):
        return GenBrownianParameter(m0, np.log(_vec(c0)), mu, np.log(_vec(sigma)))

    @staticmethod
    def brownianParameter(m0, c0, sigma):
        return BrownianParameter(m0, np.log(_vec(c0)), np.log(_vec(sigma)))

    @staticmethod
    def ouParameter(m0, c0, phi, mu, sigma):
        # the reference applies `logistic` (not logit) here, model/SdeParameters.scala:202-205,
        # and OuProcess applies logistic again (model/Sde.scala:136)
        return OuParameter(m0, np.log(_vec(c0)), SdeParameter.logistic(_vec(phi)), mu, np.log(_vec(sigma)))

    @staticmethod
    def logit(p):
        return np.log(p) - np.log(1 - p)

    @staticmethod
    def logistic(x):
        return 1.0 / (1 + np.exp(-x))


class GenBrownianParameter(SdeParameter):
    fields = ("m0", "c0", "mu", "sigma")

    def __init__(self, m0, c0, mu, sigma):
        self.m0, self.c0, self.mu, self.sigma = _vec(m0), _vec(c0), _vec(mu), _vec(sigma)


class BrownianParameter(SdeParameter):
    fields = ("m0", "c0", "sigma")

    def __init__(self, m0, c0, sigma):
        self.m0, self.c0, self.sigma = _vec(m0), _vec(c0), _vec(sigma)


class OuParameter(SdeParameter):
    fields = ("m0", "c0", "phi", "mu", "sigma")

    def __init__(self, m0, c0, phi, mu, sigma):
        self.m0, self.c0, self.phi, self.mu, self.sigma = _vec(m0), _vec(c0), _vec(phi), _vec(mu), _vec(sigma)


def buildParamRepeat(dim, m):
    """model/Sde.scala:177-179: cyclically repeat `m` to length `dim`."""
    m = _vec(m)
    return np.array([m[i % m.size] for i in range(dim)], dtype=np.float64)


class SdeInstance:
    """A parameterised SDE: kind, dimension and effective parameter vectors of length `dimension`.

    Effective = after buildParamRepeat and the constructor transforms of model/Sde.scala:70-73
    (GenBM: c0, sigma -> exp), :99-102 (BM: c0, sigma -> exp), :133-137 (OU: c0, sigma -> exp,
    phi -> logistic)."""

    def __init__(self, kind, dimension, m0, c0, sigma, mu=None, phi=None):
        self.kind, self.dimension = kind, dimension
        self.m0, self.c0, self.sigma, self.mu, self.phi = m0, c0, sigma, mu, phi


class Sde:
    """Constructors return an `UnparamSde`: a function SdeParameter -> SdeInstance that raises on
    the wrong parameter type exactly where the reference fails (model/Sde.scala:181-202)."""

    @staticmethod
    def brownianMotion(dimension):
        def run(p):
            if not isinstance(p, BrownianParameter):
                raise Exception(f"Incorrect parameters supplied to Brownianmotion, expected BrownianParameter, received {p}")
            r = lambda v: buildParamRepeat(dimension, v)
            return SdeInstance(_abi.SDE_BROWNIAN, dimension, r(p.m0), np.exp(r(p.c0)), np.exp(r(p.sigma)))
        run.dimension = dimension
        return run

    @staticmethod
    def genBrownianMotion(dimension):
        def run(p):
            if not isinstance(p, GenBrownianParameter):
                raise Exception(f"Incorrect parameters supplied to GenBrownianmotion, expected GenBrownianParameter, received {p}")
            r = lambda v: buildParamRepeat(dimension, v)
            return SdeInstance(_abi.SDE_GEN_BROWNIAN, dimension, r(p.m0), np.exp(r(p.c0)), np.exp(r(p.sigma)), mu=r(p.mu))
        run.dimension = dimension
        return run

    @staticmethod
    def ouProcess(dimension):
        def run(p):
            if not isinstance(p, OuParameter):
                raise Exception(f"Incorrect parameters supplied to OuProcess, expected OuParameter, received {p}")
            r = lambda v: buildParamRepeat(dimension, v)
            return SdeInstance(_abi.SDE_OU, dimension, r(p.m0), np.exp(r(p.c0)), np.exp(r(p.sigma)), mu=r(p.mu),
                               phi=SdeParameter.logistic(r(p.phi)))
        run.dimension = dimension
        return run
