"""GPU tests of the PMMH path and of the remaining drivers (a11):

  * cssm_filter_set_params + a resident series is the same filter as a handle created with those
    parameters (model/PMMH.scala:71 re-runs `pf(propParams)`; examples/DetermineParameters.scala:67-72),
  * a pseudo-marginal chain over the GPU filter samples the same posterior as a chain over the exact
    (Kalman) likelihood -- the defining property of PMMH (model/PMMH.scala:68-81),
  * Resampling.sampleOne (model/Resampling.scala:151-154) returns a member of the cloud,
  * `filter` (model/ParticleFilter.scala:152-158), `FilterInit` (:252-271), `filterStream` (:163-166),
  * chains / filter evaluations running side by side on one GPU give the bits they give alone.
"""
import math

import numpy as np
import pytest

import composablestatespacemodels_b200 as cs
from composablestatespacemodels_b200 import (_abi, Model, Sde, SdeParameter, Parameters, Data, Filter, FilterInit, ParticleFilter,
                                             Resampling, GpuBootstrapFilter, ParticleMetropolisHastings, ApproxPMMH, runChains)
from composablestatespacemodels_b200.parameters import add, flattenParams
import oracle
from configs import SYS, STRAT, c1, c2, c4, c4_unparam, c4_params, ou1, ou6
from test_oracle import kalman_loglik

pytestmark = pytest.mark.gpu


def _c2_unparam():
    return Model.poisson(Sde.ouProcess(1)) | Model.seasonal(24, 3, Sde.ouProcess(6))


def _c2_params(shift=0.0):
    return (Parameters(None, SdeParameter.ouParameter([1.0 + shift], [0.5], [0.2 + shift], [1.5], [0.05 + shift])) |
            Parameters(None, SdeParameter.ouParameter([0.1], [1.0 + shift], [0.4], [0.1 - shift], [0.5])))


def _c4_params(shift):
    return (Parameters(2.0 + shift, SdeParameter.brownianParameter([0.0 + shift], [1.0], [0.01 + shift])) |
            Parameters(None, SdeParameter.genBrownianParameter([0.0], [1.0 - shift], [0.01], [0.01 + 2 * shift])))


@pytest.mark.parametrize("name", ["c4", "c2"])
@pytest.mark.parametrize("dtype", [_abi.F32, _abi.F64])
@pytest.mark.parametrize("mode", [_abi.SERIES_AUTO, _abi.SERIES_THREE_LAUNCH])
def test_set_params_on_a_resident_series_equals_a_fresh_handle(name, dtype, mode):
    """theta_1 -> set_params(theta_2) -> ll_resident must give the bits of a handle created with theta_2: log-likelihood,
    every per-step value, ESS, and the final cloud."""
    um, p1, p2 = (c4_unparam(), _c4_params(0.0), _c4_params(0.07)) if name == "c4" else (_c2_unparam(), _c2_params(0.0), _c2_params(0.05))
    m1, m2 = um(p1), um(p2)
    N, T = 3000, 40
    t, y, _ = oracle.Oracle(m1).simulate(T, 0.1, 7)
    ho = np.ones(T, dtype=np.uint8)
    ho[[3, 4, 17]] = 0
    fresh = cs.GpuFilterHandle(m2, SYS, N, dtype=dtype, seed=5, stream_id=2)
    fresh.series_mode(mode)
    fresh.load_series(t, y, ho)
    want = fresh.ll_resident(steps=True)
    x_want = fresh.get_particles()
    reused = cs.GpuFilterHandle(m1, SYS, N, dtype=dtype, seed=99, stream_id=0)
    reused.series_mode(mode)
    reused.load_series(t, y, ho)
    first = reused.ll_resident()
    assert first != want[0]                       # other parameters, other seed
    reused.set_params(m2)                         # rebuilds the resident series for theta_2 (no reallocation)
    assert reused.series_len() == T
    reused.reseed(5, 2)
    got = reused.ll_resident(steps=True)
    assert got[0] == want[0]
    np.testing.assert_array_equal(got[1], want[1])
    np.testing.assert_array_equal(got[2], want[2])
    np.testing.assert_array_equal(reused.get_particles(), x_want)
    # unobserved data leave ll and ess unchanged (model/ParticleFilter.scala:121)
    assert got[1][3] == got[1][2] and got[1][4] == got[1][2] and got[2][3] == got[2][2]
    # and back again: theta_1 with its original seed reproduces the first evaluation
    reused.set_params(m1)
    reused.reseed(99, 0)
    assert reused.ll_resident() == first
    fresh.close()
    reused.close()


def test_ll_resident_buffers_follow_the_series_the_handle_holds():
    """ll_arrays / set_params (re)load a series on the C side; the per-step outputs are sized from
    cssm_filter_series_len, not from what load_series last saw."""
    mod = c1()
    t, y, _ = oracle.Oracle(mod).simulate(60, 0.1, 3)
    h = cs.GpuFilterHandle(mod, SYS, 2000, dtype=_abi.F64, seed=1)
    with pytest.raises(_abi.CssmError):
        h.ll_resident(steps=True)                 # nothing loaded yet
    h.load_series(t[:10], y[:10])
    h.ll_arrays(t, y)                             # loads all 60
    assert h.series_len() == 60
    ll, lls, ess = h.ll_resident(steps=True)
    assert lls.shape == (60,) and ess.shape == (60,) and lls[-1] == ll
    h.close()


@pytest.mark.parametrize("dtype", [_abi.F32, _abi.F64])
def test_sample_one_is_a_member_of_the_cloud(dtype):
    mod = c4()
    t, y, _ = oracle.Oracle(mod).simulate(12, 0.1, 2)
    h = cs.GpuFilterHandle(mod, SYS, 5000, dtype=dtype, seed=8)
    h.ll_arrays(t, y)
    cloud = h.get_particles()                     # [d, N]
    seen = set()
    for _ in range(6):
        x = h.sample_one()
        hit = np.where(np.all(cloud == x[:, None], axis=0))[0]
        assert hit.size >= 1, "sampleOne returned a state that is not in the cloud"
        seen.add(int(hit[0]))
    assert len(seen) > 1                          # fresh uniform index per call
    h.close()


def _chain_summary(xs, burn):
    xs = np.asarray(xs)[burn:]
    nb = 20
    m = xs[: len(xs) // nb * nb].reshape(nb, -1).mean(1)
    return xs.mean(), xs.var(), m.std(ddof=1) / math.sqrt(nb)


def test_pmmh_chain_samples_the_posterior_of_the_exact_likelihood():
    """Normal observations of a Brownian motion: the Kalman filter gives the likelihood exactly, so a Metropolis chain on
    it is the reference answer for the posterior of the (log) observation scale.  The pseudo-marginal chain over the GPU
    filter (llFilter estimate in place of the likelihood, model/PMMH.scala:68-81) has the same invariant distribution:
    posterior mean and variance agree within Monte-Carlo error."""
    um = Model.linear(Sde.brownianMotion(1))
    p_true = Parameters(math.log(0.5), SdeParameter.brownianParameter([0.0], [1.0], [0.3]))
    mod = um(p_true)
    T = 40
    t, y, _ = oracle.Oracle(mod).simulate(T, 0.1, 17)
    data = [Data(a, b) for a, b in zip(t, y)]

    def prior(p):
        s = flattenParams(p)[0]
        return 0.0 if -3.0 < s < 1.5 else -1e300

    def make_prop(rng):
        return lambda p: add(p, np.array([0.35 * rng.standard_normal(), 0.0, 0.0, 0.0]))

    # reference chain on the exact likelihood
    rng = np.random.default_rng(5)
    exact_pf = lambda p: (kalman_loglik(um(p), t, y), [None])
    ref = ParticleMetropolisHastings(p_true, make_prop(rng), lambda a, b: 0.0, prior, exact_pf, rng)
    it = ref.iters()
    xs_ref = [flattenParams(next(it).params)[0] for _ in range(12000)]
    # pseudo-marginal chain on the GPU
    rng = np.random.default_rng(6)
    pf = GpuBootstrapFilter(um, p_true, data, Resampling.systematicResampling, 4096, dtype=_abi.F32, seed=3)
    mh = ParticleMetropolisHastings(p_true, make_prop(rng), lambda a, b: 0.0, prior, pf, rng)
    it = mh.iters()
    states = [next(it) for _ in range(2500)]
    pf.close()
    xs = [flattenParams(s.params)[0] for s in states]
    assert 0.15 < states[-1].accepted / len(states) < 0.9
    assert states[-1].state.state.shape == (1,) and states[-1].state.time == t[-1]   # state._2.last: one particle
    m_ref, v_ref, se_ref = _chain_summary(xs_ref, 1000)
    m, v, se = _chain_summary(xs, 300)
    assert abs(m - m_ref) < 5 * math.hypot(se, se_ref) + 0.01, (m, m_ref, se, se_ref)
    assert 0.6 < v / v_ref < 1.6, (v, v_ref)
    # and the estimator behind it: the GPU log-likelihood at the true parameters is unbiased on the likelihood scale
    h = cs.GpuFilterHandle(mod, SYS, 1 << 14, dtype=_abi.F32, seed=11)
    est = np.array([h.ll_arrays(t, y) for _ in range(12)])
    h.close()
    exact = kalman_loglik(mod, t, y)
    assert abs(math.log(np.mean(np.exp(est - exact)))) < 0.05, (est, exact)


def test_approx_pmmh_and_concurrent_filter_evaluations():
    """ApproxPMMH (model/PMMH.scala:128-153) evaluates pf(proposed) and pf(current) in every step; with two replica
    handles the two runs go side by side (`many`).  Each replica returns exactly what it returns alone."""
    um, p0 = c4_unparam(), c4_params()
    t, y, _ = oracle.Oracle(um(p0)).simulate(60, 0.1, 4)
    data = [Data(a, b) for a, b in zip(t, y)]
    p1 = _c4_params(0.05)
    pf = GpuBootstrapFilter(um, p0, data, Resampling.systematicResampling, 1 << 14, seed=21, replicas=2)
    both = pf.many([p1, p0])
    pf.close()
    pf = GpuBootstrapFilter(um, p0, data, Resampling.systematicResampling, 1 << 14, seed=21, replicas=2)
    alone = [pf._eval(pf.handles[0], p1), pf._eval(pf.handles[1], p0)]
    assert both[0][0] == alone[0][0] and both[1][0] == alone[1][0]
    np.testing.assert_array_equal(both[0][1][0].state, alone[0][1][0].state)
    rng = np.random.default_rng(2)
    mh = ApproxPMMH(p0, cs.perturb(0.001, rng), lambda a, b: 0.0, lambda p: 0.0, pf, rng)
    it = mh.iters()
    s = [next(it) for _ in range(5)]
    pf.close()
    assert all(np.isfinite(v.ll) for v in s) and s[-1].state.state.shape == (2,)


def test_chains_side_by_side_on_one_gpu_give_the_bits_they_give_alone():
    """examples/DetermineParameters.scala:68-80 runs its chains with mapAsync(2).  Different handles are used from
    different host threads at the same time (the library is re-entrant across handles, include/cssm.h); a chain's
    states do not depend on what else the GPU is doing."""
    um, p0 = c4_unparam(), c4_params()
    t, y, _ = oracle.Oracle(um(p0)).simulate(80, 0.1, 9)
    data = [Data(a, b) for a, b in zip(t, y)]

    def chains():
        pfs, its = [], []
        for c in range(3):
            rng = np.random.default_rng(100 + c)
            pf = GpuBootstrapFilter(um, p0, data, Resampling.systematicResampling, 1 << 14, seed=7, stream_id=c)
            pfs.append(pf)
            its.append(ParticleMetropolisHastings(p0, cs.perturb(0.002, rng), lambda a, b: 0.0, lambda p: 0.0, pf, rng).iters())
        return pfs, its

    pfs, its = chains()
    together = runChains(its, 12, parallelism=3)
    for pf in pfs:
        pf.close()
    pfs, its = chains()
    alone = runChains(its, 12, parallelism=1)
    for pf in pfs:
        pf.close()
    for a, b in zip(together, alone):
        assert [s.ll for s in a] == [s.ll for s in b]
        assert [s.accepted for s in a] == [s.accepted for s in b]
    assert together[0][-1].ll != together[1][-1].ll   # independent chains


@pytest.mark.parametrize("dtype", [_abi.F32, _abi.F64])
@pytest.mark.parametrize("kind", [SYS, STRAT])
def test_filter_returns_one_member_of_the_cloud_per_time(dtype, kind):
    """`filter` (model/ParticleFilter.scala:152-158): log-likelihood + T+1 states, state s drawn uniformly from the cloud
    at time s (Resampling.sampleOne per scanLeft element).  A second handle with the same seed steps through the same
    data and shows the cloud at every time: each returned state must be one of its columns."""
    mod = c2()
    T, N = 9, 3000
    t, y, _ = oracle.Oracle(mod).simulate(T, 0.1, 12)
    ho = np.ones(T, dtype=np.uint8)
    ho[4] = 0
    a = cs.GpuFilterHandle(mod, kind, N, dtype=dtype, seed=4)
    ll, states = a.run_arrays(t, y, ho)
    assert states.shape == (T + 1, mod.dimension)
    b = cs.GpuFilterHandle(mod, kind, N, dtype=dtype, seed=4)
    b.init(float(t.min()))
    clouds = [b.get_particles()]
    ll_b = 0.0
    for s in range(T):
        ll_b, _ = b.step(t[s], float(y[s]) if ho[s] else None)
        clouds.append(b.get_particles())
    assert ll == ll_b                                  # the driver is the fold of stepFilter
    idx = []
    for s in range(T + 1):
        hit = np.where(np.all(clouds[s] == states[s][:, None], axis=0))[0]
        assert hit.size >= 1, f"state {s} is not a particle of the cloud at that time"
        idx.append(int(hit[0]))
    assert len(set(idx)) > 3                           # not the same slot every time
    # the host mirror: (ll, Vector[StateSpace]) with the times of the data behind t0
    f = Filter(mod, Resampling.systematicResampling if kind == SYS else Resampling.stratifiedResampling, dtype=dtype, seed=4)
    ll_f, sts = f.filter([Data(tt, yy if o else None) for tt, yy, o in zip(t, y, ho)], N)
    assert ll_f == ll and len(sts) == T + 1 and sts[0].time == t.min() and sts[-1].time == t[-1]
    np.testing.assert_array_equal(np.array([s.state for s in sts]), states)
    a.close()
    b.close()
    f.close()


@pytest.mark.parametrize("dtype", [_abi.F32, _abi.F64])
def test_filter_init_starts_every_particle_at_the_given_state(dtype):
    """FilterInit (model/ParticleFilter.scala:252-271): x0 = Vector.fill(particles)(initState); the first step then
    propagates from that state (checked against the oracle with injected noise)."""
    mod = c2()
    d, N = mod.dimension, 2500
    x0 = np.linspace(-0.5, 0.8, d)
    fi = FilterInit(mod, Resampling.systematicResampling, x0, dtype=dtype, seed=2)
    s0 = fi.initialiseState(N, 0.25)
    assert s0.ll == 0.0 and s0.ess == N and s0.t == 0.25
    want = np.tile((x0.astype(np.float32) if dtype == _abi.F32 else x0).astype(np.float64), (N, 1))
    np.testing.assert_array_equal(s0.particles, want)
    h = s0._handle
    rng = np.random.default_rng(1)
    z, u = rng.standard_normal((d, N)), rng.random(1)
    g = h.step_injected(0.35, 3.0, z, u)
    r = oracle.Oracle(mod)
    r.reset(N)
    ref = r.step(want.T.copy(), 0.25, 0.35, 3.0, z, u, SYS, oracle.device_order(dtype))
    tol = 1e-5 if dtype == _abi.F32 else 1e-12
    assert np.max(np.abs(g["x_prop"] - ref["x_prop"]) / np.maximum(1, np.abs(ref["x_prop"]))) <= tol
    assert np.max(np.abs(g["logw"] - ref["logw"]) / np.maximum(1, np.abs(ref["logw"]))) <= tol
    # the Reader form: ParticleFilter.filterInit(resample, t0, n, initState).run(model) is a Flow
    flow = ParticleFilter.filterInit(Resampling.systematicResampling, 0.25, N, x0, dtype=dtype, seed=2)(mod)
    out = list(flow([Data(0.35, 3.0), Data(0.45, None)]))
    assert len(out) == 3 and out[1].ess <= N and out[2].ll == out[1].ll and out[2].ess == out[1].ess
    fi.close()


def test_filter_stream_is_the_scan_of_step_filter():
    """filterStream (model/ParticleFilter.scala:163-166) = Flow[Data].scan(init)(stepFilter): T+1 states, the last ll is
    llFilter's (the single-launch series kernel and the stepping API return the same bits), and a state that the
    handle has moved past refuses to show a cloud that is no longer its own."""
    mod = c1()
    T, N = 25, 2048
    t, y, _ = oracle.Oracle(mod).simulate(T, 0.1, 6)
    data = [Data(a, b) for a, b in zip(t, y)]
    data[7] = Data(t[7], None)
    f = Filter(mod, Resampling.systematicResampling, dtype=_abi.F64, seed=13)
    states = []
    for s in f.filterStream(float(t[0]), N)(data):
        if len(states) == 3:
            s.materialise()                            # a value that outlives the next step
        states.append(s)
    assert len(states) == T + 1 and states[0].ll == 0.0 and states[0].ess == N
    assert states[8].ll == states[7].ll and states[8].ess == states[7].ess           # None observation (:121)
    g = Filter(mod, Resampling.systematicResampling, dtype=_abi.F64, seed=13)
    assert g.llFilter(data, N) == states[-1].ll
    g._handle.reseed(13, 0)                            # the same Philox streams once more, now with the per-step values
    ll, lls, ess = g._handle.ll_resident(steps=True)
    assert ll == states[-1].ll
    np.testing.assert_array_equal(lls, [s.ll for s in states[1:]])
    np.testing.assert_array_equal(ess, [s.ess for s in states[1:]])
    assert states[-1].particles.shape == (N, 1)        # the newest state owns the handle's cloud
    assert states[3].particles.shape == (N, 1)         # materialised in time
    with pytest.raises(cs.StaleStateError):
        states[5].particles
    with pytest.raises(cs.StaleStateError):
        ParticleFilter.getIntervals(mod, states[5])
    with pytest.raises(cs.StaleStateError):
        f.stepFilter(states[5], data[6])
    out = ParticleFilter.getIntervals(mod, states[-1])
    assert out.stateIntervals[0].lower <= out.state[0] <= out.stateIntervals[0].upper
    f.close()
    g.close()


def test_lgcp_call_counter_bound_is_enforced():
    """The Philox call counter of a step shares a word with the purpose tag: sub-steps x dimension beyond 2^24 calls
    would alias the resampling stream, so the library refuses it."""
    m = Model.lgcp(Sde.brownianMotion(32))(Parameters(None, SdeParameter.brownianParameter([0.0], [1.0], [0.01])))
    m = cs.model.Model(m.leaves, m.step_mode, 6)       # sub-step 1e-6: dt = 1.1 is 1.1e6 sub-steps x 32 coordinates / 2 per call
    h = cs.GpuFilterHandle(m, STRAT, 64, dtype=_abi.F64, seed=1)
    h.init(0.0)
    with pytest.raises(_abi.CssmError) as e:
        h.step(1.1, 1.0)
    assert e.value.status == -4 and "Philox" in str(e.value)
    h.step(0.001, 1.0)                                 # a short increment is fine
    h.close()


def test_residual_resampling_corrected():
    """Residual resampling (model/Resampling.scala:130-146; the reference's version cannot run, see the docstring): every
    particle keeps its floor(n w_i) deterministic copies, the remaining places are the device's multinomial draws over
    the residual weights -- bit-exact against the oracle's multinomial on the same uniforms -- and the output has n items."""
    from composablestatespacemodels_b200 import resampling
    rng = np.random.default_rng(4)
    n = 5000
    lw = rng.normal(0.0, 1.5, n)
    w = Resampling.expNormalise(lw)
    ki = np.floor(w * n).astype(int)
    resampling.seed(77)
    anc = Resampling.residualResampling(np.arange(n), lw, return_ancestors=True)
    assert anc.size == n
    m = n - ki.sum()
    np.testing.assert_array_equal(anc[: n - m], np.repeat(np.arange(n), ki))
    us = np.random.default_rng(77).random(n)           # the same host stream the mirror consumed
    want = oracle.resample(2, n * w - ki, us)[:m]
    np.testing.assert_array_equal(anc[n - m:], want)
    counts = np.bincount(anc, minlength=n)
    assert np.all(counts >= ki)
    # expected offspring n w_i: the residual part makes the scheme unbiased
    resampling.seed(5)
    tot = np.zeros(n)
    for _ in range(60):
        tot += np.bincount(Resampling.residualResampling(np.arange(n), lw, return_ancestors=True), minlength=n)
    heavy = np.argsort(w)[-50:]
    assert np.all(np.abs(tot[heavy] / 60 - n * w[heavy]) < 0.5)
    out = Resampling.residualResampling(list(range(n)), lw)
    assert len(out) == n
