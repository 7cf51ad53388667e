"""Parameter trees (reference: model/Parameters.scala).  Host side only."""
import math

import numpy as np

from .tree import Tree, Leaf, Branch


class ParamNode:
    """model/Parameters.scala:14: optional observation scale + the SDE parameters of one model."""

    def __init__(self, scale, sdeParam):
        self.scale = None if scale is None else float(scale)
        self.sdeParam = sdeParam

    def __repr__(self):
        return f"ParamNode({self.scale}, {self.sdeParam})"


def Parameters(scale, sdeParam):
    """Constructor for a leaf parameter value (model/Parameters.scala:20-22)."""
    return Leaf(ParamNode(scale, sdeParam))


def flattenParams(fa):
    """model/Parameters.scala:88-95: scale first (when present), then the SDE parameters."""
    out = []
    for node in fa.flatten():
        if node.scale is not None:
            out.append(node.scale)
        out.extend(node.sdeParam.flatten().tolist())
    return out


def paramSize(fa):
    return len(flattenParams(fa))


def add(fa, that):
    """Addable[Parameters] (model/Parameters.scala:72-103): add a flat vector, leaf by leaf."""
    that = np.asarray(that, dtype=np.float64)
    if isinstance(fa, Leaf):
        v = fa.value
        if v.scale is not None:
            return Leaf(ParamNode(v.scale + that[0], v.sdeParam.add(that[1:])))
        return Leaf(ParamNode(None, v.sdeParam.add(that)))
    n = paramSize(fa.left)
    return Branch(add(fa.left, that[:n]), add(fa.right, that[n:]))


def perturb(delta, rng=None):
    """Parameters.perturb (model/Parameters.scala:65-67): every scalar of the tree, in
    flattenParams order, gets independent N(theta, sqrt(delta)) noise."""
    rng = rng if rng is not None else np.random.default_rng()

    def run(p):
        n = paramSize(p)
        return add(p, math.sqrt(delta) * rng.standard_normal(n))
    return run


def perturbMvn(chol, rng=None):
    """model/Parameters.scala:111-114: add chol * z."""
    rng = rng if rng is not None else np.random.default_rng()
    chol = np.asarray(chol, dtype=np.float64)

    def run(p):
        return add(p, chol @ rng.standard_normal(chol.shape[1]))
    return run


def proposeIdent(p):
    return p
