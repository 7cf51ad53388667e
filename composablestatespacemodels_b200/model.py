"""Observation models and model composition (reference: model/Model.scala).

A parameterised `Model` here is the flat leaf list the device consumes (`desc()` builds the
cssm_model_desc_t of include/cssm.h).  The per-particle arithmetic of `f` and `dataLikelihood`
runs in csrc/; the small host versions below exist for single states (forecast mean, tests).
"""
import ctypes as C
import math

import numpy as np

from . import _abi
from .tree import Leaf, Branch


class ModelLeaf:
    def __init__(self, obs_kind, f_kind, sde, scale, period=0, harmonics=0, df=0):
        self.obs_kind, self.f_kind, self.sde, self.scale = obs_kind, f_kind, sde, scale
        self.period, self.harmonics, self.df = period, harmonics, df


class Model:
    """A parameterised (possibly composed) model.  Observation model and link come from the
    left-most leaf (model/Model.scala:118-132); the latent state is the concatenation of the
    leaves' SDE states in Tree.flatten order (model/Sde.scala:204-231)."""

    def __init__(self, leaves, step_mode=_abi.STEP_EXACT, lgcp_precision=0):
        self.leaves = list(leaves)
        self.step_mode = step_mode
        self.lgcp_precision = lgcp_precision

    # -- structure ---------------------------------------------------------------------------
    @property
    def obs_kind(self):
        return self.leaves[0].obs_kind

    @property
    def scale(self):
        return self.leaves[0].scale

    @property
    def df(self):
        return self.leaves[0].df

    @property
    def dimension(self):
        return sum(l.sde.dimension for l in self.leaves)

    def withStepMode(self, step_mode):
        return Model(self.leaves, step_mode, self.lgcp_precision)

    # -- host-side single-state helpers ------------------------------------------------------
    def f(self, s, t):
        """Model.f for ONE flat state vector s[d] at time t (model/Model.scala:122-128,217-225)."""
        s = np.asarray(s, dtype=np.float64)
        g, k = None, 0
        for l in self.leaves:
            x = s[k:k + l.sde.dimension]
            if l.f_kind == _abi.F_SEASONAL:
                frequency = 2 * math.pi / l.period
                fl = 0.0
                for a in range(1, l.harmonics + 1):
                    fl += math.cos(frequency * a * t) * x[2 * (a - 1)]
                    fl += math.sin(frequency * a * t) * x[2 * (a - 1) + 1]
            else:
                fl = float(x[0])
            g = fl if g is None else g + fl
            k += l.sde.dimension
        return g

    def link(self, x):
        k = self.obs_kind
        if k in (_abi.OBS_POISSON, _abi.OBS_NEGBIN, _abi.OBS_ZIP):
            return math.exp(x)
        if k == _abi.OBS_BETA:
            return math.exp(-x)  # model/Model.scala:345
        if k == _abi.OBS_BERNOULLI:
            return 1.0 if x > 6 else 0.0 if x < -6 else 1.0 / (1 + math.exp(-x))
        return x

    # -- the boundary ------------------------------------------------------------------------
    def desc(self):
        """Build the cssm_model_desc_t.  Returns (desc, keepalive)."""
        n = len(self.leaves)
        arr = (_abi.Leaf * n)()
        keep = [arr]
        for i, l in enumerate(self.leaves):
            s = l.sde
            if l.f_kind == _abi.F_SEASONAL and s.dimension != 2 * l.harmonics:
                raise Exception(f"seasonal model with {l.harmonics} harmonics needs an SDE of dimension "
                                f"{2 * l.harmonics}, got {s.dimension}")
            arr[i].sde_kind, arr[i].dim, arr[i].f_kind = s.kind, s.dimension, l.f_kind
            arr[i].period, arr[i].harmonics = l.period, l.harmonics
            for name in ("m0", "c0", "phi", "mu", "sigma"):
                v = getattr(s, name)
                if v is None:
                    setattr(arr[i], name, None)
                else:
                    v = np.ascontiguousarray(v, dtype=np.float64)
                    keep.append(v)
                    setattr(arr[i], name, v.ctypes.data_as(_abi.c_double_p))
        d = _abi.ModelDesc()
        d.n_leaves, d.leaves = n, arr
        d.obs_kind = self.obs_kind
        d.has_scale = 0 if self.scale is None else 1
        d.scale = 0.0 if self.scale is None else self.scale
        d.step_mode, d.lgcp_precision = self.step_mode, self.lgcp_precision
        d.obs_df = int(self.df)
        keep.append(d)
        return d, keep


class UnparamModel:
    """ReaderT[Try, Parameters, Model] of the reference (model/package.scala:18): call it with a
    parameter tree to get a Model; `a | b` is the reference's `a |+| b` (model/Model.scala:96-137)."""

    def __init__(self, run):
        self.run = run

    def __call__(self, p):
        return self.run(p)

    def __or__(self, other):
        return compose(self, other)


def _leaf_model(obs_kind, f_kind, sde, needs_scale=None, **kw):
    def run(p):
        if not isinstance(p, Leaf):
            raise Exception("Can't build model from branch parameter")
        node = p.value
        s = sde(node.sdeParam)
        return Model([ModelLeaf(obs_kind, f_kind, s, node.scale, **kw)])
    return UnparamModel(run)


def poisson(sde):
    return _leaf_model(_abi.OBS_POISSON, _abi.F_FIRST, sde)


def negativeBinomial(sde):
    return _leaf_model(_abi.OBS_NEGBIN, _abi.F_FIRST, sde)


def linear(sde):
    return _leaf_model(_abi.OBS_NORMAL, _abi.F_FIRST, sde)


def seasonal(period, harmonics, sde):
    return _leaf_model(_abi.OBS_NORMAL, _abi.F_SEASONAL, sde, period=period, harmonics=harmonics)


def bernoulli(sde):
    return _leaf_model(_abi.OBS_BERNOULLI, _abi.F_FIRST, sde)


def lgcp(sde):
    return _leaf_model(_abi.OBS_LGCP, _abi.F_FIRST, sde)


def studentsT(sde, df):
    """model/Model.scala:76-79,144-162; the scale parameter is the log of the t scale."""
    return _leaf_model(_abi.OBS_STUDENT_T, _abi.F_FIRST, sde, df=int(df))


def zeroInflatedPoisson(sde):
    """model/Model.scala:88-91,281-309; the scale parameter is the logit of the extra-zero probability."""
    return _leaf_model(_abi.OBS_ZIP, _abi.F_FIRST, sde)


def beta(sde):
    """model/Model.scala:49-54,339-353."""
    return _leaf_model(_abi.OBS_BETA, _abi.F_FIRST, sde)


def compose(mod1, mod2):
    """model/Model.scala:110-136: needs a Branch parameter; the observation model is mod1's."""
    def run(p):
        if not isinstance(p, Branch):
            raise Exception("Can't Build composed model from Leaf Parameter")
        m1, m2 = mod1(p.left), mod2(p.right)
        return Model(m1.leaves + m2.leaves, m1.step_mode, m1.lgcp_precision)
    return UnparamModel(run)
