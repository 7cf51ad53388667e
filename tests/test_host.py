"""CPU tests of the host-side mirror of the reference API: parameters, model composition,
descriptor building, PMMH chain logic (with a fake bootstrap filter) -- no GPU needed."""
import math

import numpy as np
import pytest

import composablestatespacemodels_b200 as cs
from composablestatespacemodels_b200 import Model, Sde, SdeParameter, Parameters, Leaf, Branch, _abi, flattenParams
from composablestatespacemodels_b200 import parameters as P
from configs import c1, c2, c4, c4_params, c4_unparam


def test_parameter_transforms_follow_the_reference():
    # user-facing ouParameter applies log / logistic; OuProcess applies exp / logistic again
    m = c1()
    s = m.leaves[0].sde
    assert s.kind == _abi.SDE_OU and s.dimension == 1
    assert abs(s.c0[0] - 0.5) < 1e-15 and abs(s.sigma[0] - 0.05) < 1e-15 and s.mu[0] == 1.5 and s.m0[0] == 1.0
    logistic = lambda x: 1 / (1 + math.exp(-x))
    assert abs(s.phi[0] - logistic(logistic(0.2))) < 1e-16
    assert abs(s.phi[0] - 0.6340970762609854) < 1e-15


def test_build_param_repeat_and_dimension():
    m = c2()
    assert m.dimension == 7 and [l.sde.dimension for l in m.leaves] == [1, 6]
    np.testing.assert_allclose(m.leaves[1].sde.sigma, np.full(6, 0.5))
    s8 = Sde.ouProcess(8)(SdeParameter.ouParameter([1.0], [2.0], [0.2], [-4, -4, 0, 0, 0, 0, -0.5, -0.5], [0.3]))
    np.testing.assert_array_equal(s8.mu, [-4, -4, 0, 0, 0, 0, -0.5, -0.5])   # examples/Simulation.scala:19


def test_composition_rules():
    with pytest.raises(Exception, match="Can't Build composed model from Leaf Parameter"):
        c4_unparam()(Parameters(1.0, SdeParameter.brownianParameter([0.0], [1.0], [0.01])))
    with pytest.raises(Exception, match="Can't build model from branch parameter"):
        Model.poisson(Sde.ouProcess(1))(c4_params())
    with pytest.raises(Exception, match="Incorrect parameters supplied to OuProcess"):
        Model.poisson(Sde.ouProcess(1))(Parameters(None, SdeParameter.brownianParameter([0.0], [1.0], [0.01])))
    m = c4()
    assert m.obs_kind == _abi.OBS_NEGBIN and m.scale == 2.0           # observation model of the left-most leaf
    three = (Model.poisson(Sde.brownianMotion(1)) | Model.linear(Sde.brownianMotion(2)) | Model.seasonal(12, 1, Sde.ouProcess(2)))
    bp = SdeParameter.brownianParameter([0.0], [1.0], [0.1])
    mm = three((Parameters(None, bp) | Parameters(0.0, bp)) | Parameters(None, SdeParameter.ouParameter([0.0], [1.0], [0.1], [0.0], [0.2])))
    assert mm.dimension == 5 and len(mm.leaves) == 3


def test_descriptor_layout():
    d, keep = c2().desc()
    assert d.n_leaves == 2 and d.obs_kind == _abi.OBS_POISSON and d.has_scale == 0
    assert d.leaves[1].f_kind == _abi.F_SEASONAL and d.leaves[1].period == 24 and d.leaves[1].harmonics == 3 and d.leaves[1].dim == 6
    assert d.leaves[0].phi[0] == c2().leaves[0].sde.phi[0]
    bad = Model.seasonal(24, 3, Sde.ouProcess(4))(Parameters(0.0, SdeParameter.ouParameter([0.0], [1.0], [0.1], [0.0], [0.2])))
    with pytest.raises(Exception, match="seasonal"):
        bad.desc()


def test_flatten_add_perturb():
    p = c4_params()
    flat = flattenParams(p)
    assert flat[0] == 2.0 and len(flat) == 1 + 3 + 4                      # scale first, model/Parameters.scala:88-95
    q = P.add(p, np.arange(8.0))
    assert flattenParams(q) == pytest.approx(np.array(flat) + np.arange(8.0))
    rng = np.random.default_rng(0)
    prop = P.perturb(0.05, rng)
    draws = np.array([flattenParams(prop(p)) for _ in range(4000)])
    assert np.allclose(draws.mean(0), flat, atol=0.02)
    assert np.allclose(draws.std(0), math.sqrt(0.05), atol=0.02)           # sd sqrt(delta), :65-67


def test_pmmh_chain_logic_with_a_fake_filter():
    """MetropolisHastings.mhStep (model/PMMH.scala:68-81) against a closed-form target: with an exact
    'filter' returning log N(theta; 1, 0.5^2) the chain must sample that normal."""
    class FakeP:
        def __init__(self, v): self.v = v
    rng = np.random.default_rng(1)
    pf = lambda p: (-0.5 * ((p.v - 1.0) / 0.5) ** 2, [("state", p.v)])
    prop = lambda p: FakeP(p.v + 0.8 * rng.standard_normal())
    mh = cs.ParticleMetropolisHastings(FakeP(5.0), prop, lambda a, b: 0.0, lambda p: 0.0, pf, rng)
    first = next(mh.markovIters())
    assert first.accepted == 1                                            # init ll = -1e99: first proposal always accepted (:121)
    # iters = markovIters.steps.drop(1) (model/PMMH.scala:95-98): the first emitted state is the SECOND mhStep
    rng2 = np.random.default_rng(7)
    a = cs.ParticleMetropolisHastings(FakeP(5.0), lambda p: FakeP(p.v + 0.8 * rng2.standard_normal()), lambda a, b: 0.0,
                                      lambda p: 0.0, pf, rng2)
    full = a.markovIters()
    two = [next(full), next(full)]
    rng2 = np.random.default_rng(7)
    b = cs.ParticleMetropolisHastings(FakeP(5.0), lambda p: FakeP(p.v + 0.8 * rng2.standard_normal()), lambda a, b: 0.0,
                                      lambda p: 0.0, pf, rng2)
    emitted = next(b.iters())
    assert emitted.params.v == two[1].params.v and emitted.accepted == two[1].accepted
    it = mh.iters()
    xs = np.array([next(it).params.v for _ in range(20000)])[2000:]
    assert abs(xs.mean() - 1.0) < 0.05 and abs(xs.std() - 0.5) < 0.05


def test_resample_kind_mapping():
    R = cs.Resampling
    assert R.kind_of(R.systematicResampling) == 0 and R.kind_of(R.stratifiedResampling) == 1 and R.kind_of(R.multinomialResampling) == 2
    with pytest.raises(Exception):
        R.kind_of(lambda p, w: p)
    np.testing.assert_allclose(R.normalise([1, 1, 2]), [0.25, 0.25, 0.5])


def test_simulator_is_deterministic_and_shaped():
    from composablestatespacemodels_b200 import simulate
    t, y, x = simulate.simRegular(c2(), 0.1, 50, seed=3)
    t2, y2, _ = simulate.simRegular(c2(), 0.1, 50, seed=3)
    assert np.array_equal(y, y2) and x.shape == (50, 7) and np.allclose(np.diff(t), 0.1)
    assert np.all(y >= 0) and np.all(y == np.floor(y))


def test_approx_pmmh_reestimates_the_current_likelihood():
    """model/PMMH.scala:128-153 with a fake filter: two filter calls per step, and a rejected
    proposal still replaces the stored likelihood by the re-estimate."""
    from composablestatespacemodels_b200.pmmh import ApproxPMMH, pmmhStep
    calls = []

    def pf(p):
        calls.append(p)
        return (-10.0 - 100.0 * abs(p) - 0.001 * len(calls), [("state", len(calls))])

    rng = np.random.default_rng(3)
    mh = ApproxPMMH(0.0, lambda p: p + 5.0, lambda a, b: 0.0, lambda p: 0.0, pf, rng)
    it = mh.markovIters()
    s1 = next(it)
    assert len(calls) == 2 and calls == [5.0, 0.0]
    assert s1.accepted == 0 and s1.params == 0.0 and s1.ll == -10.0 - 0.002 and s1.state == ("state", 2)
    s2 = next(it)
    assert len(calls) == 4 and s2.ll == -10.0 - 0.004
    step = pmmhStep(lambda p: -abs(p), lambda p: p * 0.5, rng)
    assert step((-4.0, 4.0)) == (-2.0, 2.0)


def test_resampling_host_helpers():
    """model/Resampling.scala:29,102-122,151-162"""
    from composablestatespacemodels_b200 import Resampling
    import numpy as np
    lw = np.array([-1000.0, -1001.0, -1002.0])
    en = Resampling.expNormalise(lw)
    assert abs(en.sum() - 1) < 1e-15 and en[0] > en[1] > en[2] > 0
    np.testing.assert_allclose(Resampling.cumSum([1, 2, 3]), [0, 1, 3, 6])
    np.testing.assert_allclose(Resampling.empDist([1, 1, 2]), [0, 0.25, 0.5, 1.0])
    assert Resampling.indentity([1, 2], [0.5, 0.5]) == [1, 2]
    s = list(range(10))
    assert Resampling.sampleOne(s) in s
    m = Resampling.sampleMany(4, s)
    assert len(m) == 4 and len(set(m)) == 4


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU restatement on the host cores, no GPU involved) prints one JSON line with
    the keys of the bench contract; rank != 0 of a multi-process launch prints nothing and exits 0."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CSSM_BENCH_BUDGET_S="0.3")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "dtype", "data", "config",
              "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in j, k
    assert j["impl"] == "reference" and j["value"] > 0 and j["gpu_launches"] == 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
