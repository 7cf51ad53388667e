"""The reference's own tests for this path, restated against the oracle (CPU) and the library (GPU):

  src/test/scala/SamplingTest.scala:12-22  ScalaCheck properties: for a non-empty vector of weights in [0, 1] each of
                                           the three resamplers returns a vector of the same length
  src/test/scala/ModelTest.scala:66-89     "Brownian Motion step function should change the value of the state",
                                           "Compose two models should work" (state of the composed model has the
                                           composed dimension, the leaves advance separately, f sums the leaves)

The ScalaCheck generators become hypothesis strategies; on top of the reference's only property (the length) the
oracle's two summation orders and the literal TreeMap are cross-checked on every generated vector."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle
from composablestatespacemodels_b200 import Model, Sde, SdeParameter, Parameters, _abi
from configs import SYS, STRAT, MULTI

weights = st.lists(st.floats(min_value=0.0, max_value=1.0, allow_nan=False), min_size=1, max_size=300)


def _usable(w):
    """Not pinned: an all-zero vector (the reference divides by a zero total), and vectors whose LARGEST weight is below
    1e-20 -- the device definition sums weights as multiples of 2^-96 (weights above 1 are scaled down by a power of two,
    tiny ones are not scaled up), so a vector of subnormals is a zero total there.  Inside a filter the largest weight is
    exp(0) = 1 by construction."""
    return np.max(w) >= 1e-20


@settings(max_examples=200, deadline=None)
@given(w=weights, seed=st.integers(0, 2 ** 32 - 1))
def test_resampling_returns_a_vector_of_the_same_length(w, seed):
    w = np.array(w)
    if not _usable(w):
        return  # the reference divides by a zero total here (NaN keys): nothing to pin
    rng = np.random.default_rng(seed)
    n = w.size
    u1, un = rng.random(1), rng.random(n)
    for kind, u in ((SYS, u1), (STRAT, un), (MULTI, un)):
        for order in (oracle.ORDER_REFERENCE, oracle.ORDER_DEVICE):
            anc = oracle.resample(kind, w, u, order)
            assert anc.size == n                                   # the reference's property
            assert anc.min() >= 0 and anc.max() < n
            if kind != MULTI:
                assert np.all(np.diff(anc) >= 0)
                assert np.all(w[anc] > 0) or np.any(w == 0)        # a zero weight is selected only through the TreeMap quirk
    # the literal TreeMap (std::map) and the merge restatement of the reference order agree on every vector
    np.testing.assert_array_equal(oracle.resample_treemap(SYS, w, u1), oracle.resample(SYS, w, u1, oracle.ORDER_REFERENCE))
    np.testing.assert_array_equal(oracle.resample_treemap(STRAT, w, un), oracle.resample(STRAT, w, un, oracle.ORDER_REFERENCE))
    # the textbook rule never selects a particle of zero weight
    for kind, u in ((SYS, u1), (STRAT, un)):
        a = oracle.resample(kind, w, u, oracle.ORDER_DEVICE | oracle.TIE_FIRST)
        keep = a < n - 1                                           # the clamp at the top end may land on a zero weight
        assert np.all(w[a[keep]] > 0)


@settings(max_examples=100, deadline=None)
@given(w=weights, seed=st.integers(0, 2 ** 32 - 1))
def test_device_and_reference_summation_orders_agree(w, seed):
    """Exact integer sums rounded once (what the GPU computes) against the reference's sequential fp64 sums: the same
    ancestors except where a uniform falls within a few ulps of a cumulative weight."""
    w = np.array(w)
    if not _usable(w):
        return
    rng = np.random.default_rng(seed)
    u = rng.random(1)
    a = oracle.resample(SYS, w, u, oracle.ORDER_DEVICE)
    b = oracle.resample(SYS, w, u, oracle.ORDER_REFERENCE)
    assert np.mean(a != b) <= 0.02 + 1.0 / w.size


def _bm(dim=1):
    return Sde.brownianMotion(dim), SdeParameter.brownianParameter([1.0] * dim, [1.0] * dim, [1.0] * dim)


def test_brownian_motion_step_changes_the_state():
    """ModelTest.scala:66-72"""
    sde, p = _bm()
    mod = Model.linear(sde)(Parameters(1.0, p))
    orc = oracle.Oracle(mod)
    x0 = np.array([[1.0]])
    x1 = orc.propagate(x0, np.array([[0.37]]), 2.0)
    assert x1[0, 0] != x0[0, 0]
    # x + sqrt(sigma dt) z (model/Sde.scala:114-123); brownianParameter stores log(sigma) and the SDE applies exp: sigma = 1
    assert x1[0, 0] == 1.0 + np.sqrt(1.0 * 2.0) * 0.37


def test_compose_two_models_works():
    """ModelTest.scala:74-89: the composed state has the composed dimension, the two leaves advance separately and f
    sums their first components (the reference's no-noise model makes y == eta; here the link of the Normal model is
    the identity, so eta == gamma == x1_left + x1_right)."""
    sde, p = _bm()
    single = Parameters(1.0, p)
    mod = (Model.linear(sde) | Model.linear(sde))(single | single)
    assert mod.dimension == 2
    orc = oracle.Oracle(mod)
    z0 = np.array([[0.3], [-1.1]])
    x0 = orc.init_state(z0)
    assert x0.shape == (2, 1)
    x1 = orc.propagate(x0, np.array([[0.5], [0.25]]), 1.0)
    assert x1[0, 0] != x0[1, 0] and x1[0, 0] != x0[0, 0] and x1[1, 0] != x0[1, 0]
    # the leaves do not mix: the left leaf only sees its own noise
    x1b = orc.propagate(x0, np.array([[0.5], [-2.0]]), 1.0)
    assert x1b[0, 0] == x1[0, 0] and x1b[1, 0] != x1[1, 0]
    gamma = orc.f(x1, 1.0)[0]
    assert gamma == x1[0, 0] + x1[1, 0]
    assert orc.link(gamma) == gamma == mod.link(gamma)
    assert mod.f(x1[:, 0], 1.0) == gamma


@pytest.mark.gpu
@settings(max_examples=40, deadline=None)
@given(w=weights, seed=st.integers(0, 2 ** 32 - 1))
def test_gpu_resampling_properties(w, seed):
    """SamplingTest.scala against the library: same length, and the oracle's ancestors bit for bit, on every
    generated vector (ragged sizes 1 .. 300, zeros, ties)."""
    from composablestatespacemodels_b200.resampling import ancestors
    w = np.array(w)
    if not _usable(w):
        return
    rng = np.random.default_rng(seed)
    n = w.size
    for kind, u in ((SYS, rng.random(1)), (STRAT, rng.random(n)), (MULTI, rng.random(n))):
        got = ancestors(kind, w, u)
        assert got.size == n
        np.testing.assert_array_equal(got, oracle.resample(kind, w, u, oracle.ORDER_DEVICE))
