"""CPU tests of the oracle itself.  The reference's own tests pin no value of this path
(PARITY UNPINNED, see oracle/cssm_oracle.cpp), so the restatement is pinned against
  - independent implementations of the third-party (Breeze) densities: scipy.stats,
  - the literal TreeMap formulation of the reference's resampling (std::map),
  - an exact Kalman-filter likelihood for a linear-Gaussian composition,
  - properties the reference tests do state (SamplingTest.scala: output length).
"""
import math

import numpy as np
import pytest
from scipy import stats, special

import oracle
from composablestatespacemodels_b200 import Model, Sde, SdeParameter, Parameters, _abi
from configs import ALL, SYS, STRAT, MULTI, c1, c2, c3, c4, c5


def test_exp_det_is_accurate_and_monotone():
    xs = -np.abs(np.random.default_rng(0).normal(0, 60, 20000))
    e = np.array([oracle.exp_det(x) for x in xs])
    r = np.exp(xs)
    ok = r > 1e-300
    assert np.max(np.abs(e[ok] - r[ok]) / r[ok]) < 2.5e-16
    assert oracle.exp_det(0.0) == 1.0 and oracle.exp_det(-1e4) == 0.0
    s = np.sort(xs)
    es = np.array([oracle.exp_det(x) for x in s])
    assert np.all(np.diff(es) >= 0)


def test_expf_det_is_accurate_and_monotone():
    """The fp32 weight evaluation of F32 filters (ORDER_DEVICE_F32)."""
    xs = np.sort(-np.abs(np.random.default_rng(3).normal(0, 25, 20000)).astype(np.float32))
    e = np.array([oracle.expf_det(x) for x in xs], dtype=np.float64)
    r = np.exp(xs.astype(np.float64))
    ok = xs >= -86.0
    assert np.max(np.abs(e[ok] - r[ok]) / r[ok]) < 1.2e-7            # < 1 ulp(fp32) = 2^-23
    assert np.all(e[~ok] == 0.0)                                     # would be subnormal: flushed to exactly 0
    assert np.all(np.diff(e) >= 0) and oracle.expf_det(0.0) == 1.0
    lw = np.random.default_rng(4).normal(-5, 3, 1000).astype(np.float32).astype(np.float64)
    w = oracle.w1(lw, lw.max(), oracle.ORDER_DEVICE_F32)
    assert np.all(w == w.astype(np.float32)) and w.max() == 1.0      # fp32 values, the max weight is exactly 1
    np.testing.assert_allclose(w, np.exp(lw - lw.max()), rtol=4e-6)   # + the rounding of the fp32 subtraction


def test_fixed_point_round_trip():
    L = oracle.lib()
    import ctypes as C
    rng = np.random.default_rng(1)
    for x in list(rng.random(200)) + [1.0, 0.5, 2.0 ** -40, 2.0 ** -95, 3e-29, 0.0]:
        lo, hi = C.c_uint64(), C.c_uint64()
        L.orc_fix96(x, C.byref(lo), C.byref(hi))
        v = (hi.value << 64) | lo.value
        assert v == int(x * 2 ** 96) if x >= 2.0 ** -43 else v <= x * 2 ** 96
        if x >= 2.0 ** -43:
            assert L.orc_unfix96(lo.value, hi.value) == x   # 53 significant bits fit above 2^-96


def test_densities_against_scipy():
    g = np.linspace(-4, 4, 33)
    m = c1()
    o = oracle.Oracle(m)
    for y in (0.0, 1.0, 7.0, 23.9):
        np.testing.assert_allclose(o.loglik(g, y), stats.poisson.logpmf(int(y), np.exp(g)), rtol=1e-12, atol=1e-12)
    o = oracle.Oracle(c4())
    size = math.exp(2.0)
    for y in (0.0, 3.0, 11.0):
        mu = np.exp(g)
        np.testing.assert_allclose(o.loglik(g, y), stats.nbinom.logpmf(int(y), size, size / (size + mu)), rtol=1e-11, atol=1e-11)
    o = oracle.Oracle(c5())
    for y in (-1.3, 0.0, 2.5):
        np.testing.assert_allclose(o.loglik(g, y), stats.norm.logpdf(y, g, 1.0), rtol=1e-13, atol=1e-13)
    o = oracle.Oracle(ALL["bernoulli"]())
    p = special.expit(g)
    np.testing.assert_allclose(o.loglik(g, 1.0), np.log(p), rtol=1e-12)
    np.testing.assert_allclose(o.loglik(g, 0.0), np.log1p(-p), rtol=1e-9, atol=1e-12)
    assert o.loglik(np.array([-7.0]), 1.0)[0] == -1e99 and o.loglik(np.array([7.0]), 0.0)[0] == -1e99  # model/Model.scala:332-334


def test_further_densities_against_scipy():
    """Student-t (with the reference's 1/v factor on the log-density), zero-inflated Poisson, Beta
    (model/Model.scala:154-160, :298-306, :349-352)."""
    g = np.linspace(-2, 2, 17)
    o = oracle.Oracle(ALL["student_t"]())
    v = math.exp(-0.7)
    for y in (-1.0, 0.4, 3.0):
        np.testing.assert_allclose(o.loglik(g, y), stats.t.logpdf((y - g) / v, 5) / v, rtol=1e-12, atol=1e-12)
    o = oracle.Oracle(ALL["zip"]())
    p = special.expit(-1.2)
    np.testing.assert_allclose(o.loglik(g, 0.0), np.log(p + (1 - p) * stats.poisson.pmf(0, np.exp(g))), rtol=1e-12)
    for y in (1.0, 6.0):
        np.testing.assert_allclose(o.loglik(g, y), np.log1p(-p) + stats.poisson.logpmf(int(y), np.exp(g)), rtol=1e-12, atol=1e-12)
    o = oracle.Oracle(ALL["beta"]())
    for y in (0.1, 0.5, 0.93):
        np.testing.assert_allclose(o.loglik(g, y), stats.beta.logpdf(y, np.exp(-g), 1.0), rtol=1e-11, atol=1e-12)


def test_intervals_are_the_literal_order_statistics():
    """getIntervals / getCredibleInterval / getOrderStatistic (model/ParticleFilter.scala:415-424,455-460,490-505)."""
    for name in ("c2", "beta", "bernoulli", "c4"):
        mod = ALL[name]()
        o = oracle.Oracle(mod)
        rng = np.random.default_rng(1)
        N, t = 1000, 0.7
        x = 0.5 * rng.standard_normal((mod.dimension, N))
        r = o.intervals(x, t)
        idx = math.floor(0.975 * N)
        xs = np.sort(x, axis=1)
        np.testing.assert_array_equal(r["lower"], xs[:, N - idx - 1])
        np.testing.assert_array_equal(r["upper"], xs[:, idx - 1])
        np.testing.assert_allclose(r["mean"], x.mean(axis=1), rtol=1e-12, atol=1e-14)
        eta = np.sort([mod.link(mod.f(x[:, i], t)) for i in range(N)])
        np.testing.assert_allclose(r["eta"], [mod.link(mod.f(r["mean"], t)), eta[N - idx], eta[idx]], rtol=1e-12)
    with pytest.raises(IndexError):
        o.intervals(x, t, 1.0)


def test_transitions_have_the_reference_moments():
    # OU exact transition: mean mu + (x - mu) e^{-phi dt}, variance sigma^2/(2 phi) (1 - e^{-2 phi dt})  (model/Sde.scala:139-150)
    m = c1()
    o = oracle.Oracle(m)
    sde = m.leaves[0].sde
    N, dt = 400000, 0.7
    rng = np.random.default_rng(3)
    x0 = np.full((1, N), 0.3)
    x1 = o.propagate(x0, rng.standard_normal((1, N)), dt)
    phi, mu, sig = sde.phi[0], sde.mu[0], sde.sigma[0]
    assert abs(x1.mean() - (mu + (0.3 - mu) * math.exp(-phi * dt))) < 5e-4
    assert abs(x1.var() - sig * sig / (2 * phi) * (1 - math.exp(-2 * phi * dt))) < 1e-4
    # dt = 0: zero variance, mean mu + (x - mu) * 1 -- the identity up to one rounding, noise still consumed
    np.testing.assert_allclose(o.propagate(x0, rng.standard_normal((1, N)), 0.0), x0, rtol=3e-16)
    # Brownian motion: sigma is a variance rate in the exact step (model/Sde.scala:117)
    mb = ALL["bernoulli"]()
    ob = oracle.Oracle(mb)
    xb = ob.propagate(np.zeros((2, N)), rng.standard_normal((2, N)), 2.0)
    assert abs(xb[0].var() - 0.3 * 2.0) < 5e-3


def test_seasonal_f_matches_buildF():
    m = c2()
    o = oracle.Oracle(m)
    x = np.random.default_rng(4).standard_normal((7, 5))
    t = 3.7
    w = 2 * math.pi / 24
    F = np.array([f(w * a * t) for a in (1, 2, 3) for f in (math.cos, math.sin)])
    np.testing.assert_allclose(o.f(x, t), x[0] + F @ x[1:], rtol=1e-14)
    assert abs(m.f(x[:, 0], t) - o.f(x, t)[0]) < 1e-14   # host-side Model.f agrees


@pytest.mark.parametrize("kind", [SYS, STRAT])
def test_merge_equals_literal_treemap(kind):
    rng = np.random.default_rng(5)
    for n in (1, 2, 17, 1000, 5000):
        for w in (rng.random(n), np.exp(rng.normal(0, 5, n)), np.where(rng.random(n) < 0.5, 0.0, rng.random(n)) + (np.arange(n) == 0)):
            u = rng.random(1 if kind == SYS else n)
            a = oracle.resample(kind, w, u, oracle.ORDER_REFERENCE)
            np.testing.assert_array_equal(a, oracle.resample_treemap(kind, w, u))
            assert a.size == n                       # SamplingTest.scala:12-22
            assert np.all(np.diff(a) >= 0)


@pytest.mark.parametrize("kind", [SYS, STRAT, MULTI])
def test_device_order_agrees_with_reference_order(kind):
    """The order-invariant (exact fixed-point) definition the GPU uses picks the same ancestors as
    the reference's sequential fp64 arithmetic, up to last-ulp events."""
    rng = np.random.default_rng(6)
    n_diff = n_tot = 0
    for n in (10, 1000, 20000):
        for w in (rng.random(n), np.exp(rng.normal(0, 3, n)), np.exp(rng.normal(0, 12, n))):
            u = rng.random(1 if kind == SYS else n)
            a = oracle.resample(kind, w, u, oracle.ORDER_REFERENCE)
            b = oracle.resample(kind, w, u, oracle.ORDER_DEVICE)
            n_diff += int(np.sum(a != b))
            n_tot += n
    assert n_diff <= 2, (n_diff, n_tot)


def test_systematic_offspring_counts():
    rng = np.random.default_rng(7)
    n = 5000
    w = rng.random(n)
    a = oracle.resample(SYS, w, rng.random(1), oracle.ORDER_DEVICE)
    counts = np.bincount(a, minlength=n)
    expect = n * w / w.sum()
    assert np.all(np.abs(counts - expect) < 1.0 + 1e-9)


def test_multinomial_reference_walk():
    # Breeze Multinomial.draw (first draw): prob = u * sum; subtract weights in order until prob <= 0
    w = np.array([0.2, 0.0, 0.5, 0.3, 0.0, 0.0])
    u = np.array([0.0, 0.1999, 0.2, 0.21, 0.75, 0.999])
    a = oracle.resample(MULTI, w, u, oracle.ORDER_REFERENCE)
    np.testing.assert_array_equal(a, [0, 0, 0, 2, 3, 3])
    np.testing.assert_array_equal(oracle.resample(MULTI, w, u, oracle.ORDER_DEVICE), [0, 0, 0, 2, 3, 3])


def test_ll_ess_definitions():
    rng = np.random.default_rng(8)
    lw = rng.normal(-3, 2, 1000)
    mx = lw.max()
    for order in (oracle.ORDER_REFERENCE, oracle.ORDER_DEVICE):
        w1 = oracle.w1(lw, mx, order)
        incr, ess = oracle.ll_ess(w1, mx, order)
        assert abs(incr - special.logsumexp(lw) + math.log(1000)) < 1e-12   # max + log(mean(exp(w - max)))
        wn = w1 / w1.sum()
        assert ess == int(math.floor(1 / np.sum(wn * wn)))


def kalman_loglik(mod, t, y):
    """Exact log-likelihood of a Normal-observation composition of OU / Brownian leaves."""
    d = mod.dimension
    m0 = np.concatenate([l.sde.m0 for l in mod.leaves])
    P = np.diag(np.concatenate([l.sde.c0 for l in mod.leaves]))
    m = m0.copy()
    r = math.exp(mod.scale) ** 2
    ll, tp = 0.0, t[0]
    for s in range(len(t)):
        dt = t[s] - tp
        A, c, Q = np.zeros(d), np.zeros(d), np.zeros(d)
        k = 0
        for l in mod.leaves:
            sde = l.sde
            for j in range(sde.dimension):
                if sde.kind == _abi.SDE_OU:
                    a = math.exp(-sde.phi[j] * dt)
                    A[k], c[k] = a, sde.mu[j] * (1 - a)
                    Q[k] = sde.sigma[j] ** 2 / (2 * sde.phi[j]) * (1 - math.exp(-2 * sde.phi[j] * dt))
                else:
                    A[k], c[k], Q[k] = 1.0, (sde.mu[j] * dt if sde.mu is not None else 0.0), sde.sigma[j] * dt
                k += 1
        m = A * m + c
        P = (A[:, None] * P) * A[None, :] + np.diag(Q)
        H = np.zeros(d)
        k = 0
        for l in mod.leaves:
            if l.f_kind == _abi.F_SEASONAL:
                w = 2 * math.pi / l.period
                for a_ in range(1, l.harmonics + 1):
                    H[k + 2 * (a_ - 1)] = math.cos(w * a_ * t[s])
                    H[k + 2 * (a_ - 1) + 1] = math.sin(w * a_ * t[s])
            else:
                H[k] = 1.0
            k += l.sde.dimension
        S = H @ P @ H + r
        e = y[s] - H @ m
        ll += -0.5 * (math.log(2 * math.pi * S) + e * e / S)
        K = P @ H / S
        m = m + K * e
        P = P - np.outer(K, H @ P)
        tp = t[s]
    return ll


def test_oracle_filter_matches_kalman():
    mod = c5()
    o = oracle.Oracle(mod)
    t, y, _ = o.simulate(30, 0.1, 5)
    exact = kalman_loglik(mod, t, y)
    est = o.filter_ll_many(20000, SYS, t, y, seed=1, variant=1, R=8, threads=8)
    # E[exp(ll_hat)] = exp(ll): compare on the likelihood scale with the MC standard error
    lm = special.logsumexp(est) - math.log(len(est))
    se = np.std(np.exp(est - exact)) / math.sqrt(len(est))
    assert abs(math.exp(lm - exact) - 1) < 5 * se + 0.02, (lm, exact, se)
    # both cost-model variants estimate the same quantity
    est0 = o.filter_ll_many(5000, SYS, t, y, seed=2, variant=0, R=4, threads=4)
    assert abs(est0.mean() - exact) < 0.3


def test_lgcp_step_definition():
    mod = c3(precision=2)
    o = oracle.Oracle(mod)
    N = 4
    x = np.array([[0.1, -0.2, 0.3, 0.0]])
    assert oracle.lgcp_nsub(0.0, 2) == 0 and oracle.lgcp_nsub(0.031, 2) == 4 and oracle.lgcp_nsub(0.03, 2) in (3, 4)
    n = oracle.lgcp_nsub(0.05, 2)
    z = np.random.default_rng(9).standard_normal((n, 1, N))
    o.reset(N)
    r = o.step(x, 0.0, 0.05, 1.0, z, np.random.default_rng(1).random(N), STRAT, oracle.ORDER_REFERENCE)
    # by hand: Brownian exact sub-steps of length 0.01 (variance rate sigma), hazard over post-step states
    sd = math.sqrt(0.01 * 0.01)
    xs, hz = x[0].copy(), np.zeros(N)
    for s in range(n):
        xs = sd * z[s, 0] + xs
        hz = hz + np.exp(xs) * 0.01
    np.testing.assert_allclose(r["x_prop"][0], xs, rtol=1e-15)
    np.testing.assert_allclose(r["logw"], xs - hz, rtol=1e-14)
    # dt == 0: weights f - f = 0, state untouched
    o.reset(N)
    r0 = o.step(x, 0.05, 0.05, 1.0, z, np.random.default_rng(1).random(N), STRAT, oracle.ORDER_REFERENCE)
    np.testing.assert_array_equal(r0["logw"], np.zeros(N))
    np.testing.assert_array_equal(r0["x_prop"], x)


def test_euler_maruyama_definition():
    mod = c1().withStepMode(_abi.STEP_EULER)
    o = oracle.Oracle(mod)
    sde = mod.leaves[0].sde
    x = np.array([[0.4, 1.9]])
    z = np.array([[0.5, -1.0]])
    dt = 0.1
    want = x + sde.phi[0] * (sde.mu[0] - x) * dt + sde.sigma[0] * (math.sqrt(dt) * z)
    np.testing.assert_allclose(o.propagate(x, z, dt), want, rtol=1e-15)
    mb = ALL["bernoulli"]().withStepMode(_abi.STEP_EULER)   # Brownian drift is the constant 1.0 (model/Sde.scala:110)
    ob = oracle.Oracle(mb)
    xb = np.zeros((2, 1))
    np.testing.assert_allclose(ob.propagate(xb, np.zeros((2, 1)), 0.25), np.full((2, 1), 0.25))


def test_oracle_forecast_moments():
    """Oracle.forecast (getForecast / getMeanForecast restated with NumPy samplers): law of total variance of the
    drawn observations, and the interval ranks of getCredibleInterval / getOrderStatistic."""
    import oracle
    from configs import c5, c1
    rng = np.random.default_rng(3)
    N = 100000
    for make, kind in ((c5, "normal"), (c1, "poisson")):
        mod = make()
        orc = oracle.Oracle(mod)
        x = orc.init_state(rng.standard_normal((mod.dimension, N)))
        r = orc.forecast(x, 0.0, 0.4, rng, 0.975)
        g, eta, obs = r["gamma"], r["eta"], r["obs"]
        if kind == "normal":
            np.testing.assert_allclose(eta, g)
            assert abs(obs.var() - (g.var() + np.exp(mod.scale) ** 2)) < 0.05 * obs.var()
        else:
            np.testing.assert_allclose(eta, np.exp(g))
            assert abs(obs.var() - (eta.mean() + eta.var())) < 0.05 * obs.var()
        assert abs(obs.mean() - eta.mean()) < 6 * obs.std() / np.sqrt(N)
        idx = int(np.floor(0.975 * N))
        xs = np.sort(r["x"][0])
        assert r["lower"][0] == xs[N - idx - 1] and r["upper"][0] == xs[idx - 1]
        es = np.sort(eta)
        assert r["eta_summary"][1] == es[N - idx] and r["eta_summary"][2] == es[idx]
