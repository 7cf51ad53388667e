"""GPU parity: the CUDA path through the C ABI against the CPU oracle on identical injected
noise and uniforms.  Tolerances (BASELINE.json north_star): states, log-weights and the
log-likelihood within 1e-5 relative in fp32 and 1e-12 in fp64; ancestor indices bit-exact."""
import numpy as np
import pytest

import composablestatespacemodels_b200 as cs
from composablestatespacemodels_b200 import _abi
import oracle
from configs import ALL, EXTRA, SYS, STRAT, MULTI, c1, c2, c3, c5

pytestmark = pytest.mark.gpu

TOL = {_abi.F64: 1e-12, _abi.F32: 1e-5}


def rel_close(a, b, tol, what):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), 1.0)
    err = np.max(np.abs(a - b) / scale)
    assert err <= tol, f"{what}: max relative error {err:.3e} > {tol:.1e}"


def run_steps(mod, N, T, kind, dtype, seed=0, missing=(), tie_first=False):
    """T stepFilters with injected noise on the GPU; each stage checked against the oracle fed
    with the device's own output of the previous stage (so every comparison is like for like)."""
    rng = np.random.default_rng(seed)
    orc = oracle.Oracle(mod)
    d = mod.dimension
    t, y, _ = orc.simulate(T, 0.1, seed + 11)
    h = cs.GpuFilterHandle(mod, kind, N, dtype=dtype, seed=3)
    if tie_first:
        h.set_tie_rule(_abi.TIE_FIRST)
    tol = TOL[dtype]
    DEV = oracle.device_order(dtype)   # F32 filters evaluate w1 = exp(logw - max) in fp32
    TIE = oracle.TIE_FIRST if tie_first else 0
    z0 = rng.standard_normal((d, N))
    h.init_injected(t[0], z0)
    x = h.get_particles()
    rel_close(x, orc.init_state(z0), tol, "initial state")
    tp, ll_ref = t[0], 0.0
    for s in range(T):
        obs = None if s in missing else float(y[s])
        z = rng.standard_normal((d, N))
        u = rng.random(1 if kind == SYS else N)
        g = h.step_injected(t[s], obs, z, u)
        # stage 1: propagate + weight, from the device's own previous cloud
        o = oracle.Oracle(mod)
        o.reset(N)
        r = o.step(x, tp, t[s], obs, z, u, kind, DEV)
        rel_close(g["x_prop"], r["x_prop"], tol, f"step {s} propagated state")
        if obs is None:
            x = h.get_particles()
            np.testing.assert_array_equal(x, g["x_prop"])
            tp = t[s]
            continue
        rel_close(g["logw"], r["logw"], tol, f"step {s} log-weights")
        # stage 2: from the device's log-weights -> w1, ll increment, ESS, ancestors
        mx = float(np.max(g["logw"]))
        w1 = oracle.w1(g["logw"], mx, DEV)
        np.testing.assert_array_equal(g["w1"], w1)  # deterministic exp: identical bits
        incr, ess = oracle.ll_ess(w1, mx, DEV)
        ll_ref += incr
        assert g["ess"] == ess
        rel_close(g["ll"], ll_ref, 1e-12, f"step {s} log-likelihood")
        anc = oracle.resample(kind, w1, u, DEV | TIE)
        np.testing.assert_array_equal(g["anc"], anc)
        # the reference-order (sequential fp64) restatement agrees except for last-ulp ties
        anc_ref = oracle.resample(kind, oracle.w1(g["logw"], mx, oracle.ORDER_REFERENCE), u, oracle.ORDER_REFERENCE | TIE)
        assert np.mean(anc_ref != anc) <= 2e-3
        # stage 3: gather
        x = h.get_particles()
        np.testing.assert_array_equal(x, g["x_prop"][:, anc])
        tp = t[s]
    h.close()


@pytest.mark.parametrize("name", sorted(ALL))
@pytest.mark.parametrize("dtype", [_abi.F64, _abi.F32])
def test_step_parity_systematic(name, dtype):
    run_steps(ALL[name](), 1000, 6, SYS, dtype, seed=1)


@pytest.mark.parametrize("kind", [STRAT, MULTI])
@pytest.mark.parametrize("dtype", [_abi.F64, _abi.F32])
def test_step_parity_other_resamplers(kind, dtype):
    run_steps(c2(), 3000, 4, kind, dtype, seed=2)


def test_step_parity_ragged_and_missing():
    # N not a multiple of anything, observations missing at steps 1 and 2
    run_steps(c2(), 4099, 5, SYS, _abi.F64, seed=3, missing=(1, 2))
    run_steps(c1(), 1, 3, SYS, _abi.F64, seed=4)
    run_steps(c1(), 2049, 3, STRAT, _abi.F32, seed=5)


def test_step_parity_large_tiles(monkeypatch):
    # clouds above 2^18 particles use 2048-particle tiles; force them on a small cloud
    monkeypatch.setenv("CSSM_TILE_ITEMS", "8")
    run_steps(c2(), 5000, 3, SYS, _abi.F32, seed=8)
    run_steps(c2(), 4100, 3, STRAT, _abi.F64, seed=9)
    monkeypatch.setenv("CSSM_PDL", "0")  # and without programmatic dependent launch
    run_steps(c2(), 3000, 3, SYS, _abi.F32, seed=10)


def test_step_parity_tie_first_option():
    """CSSM_TIE_FIRST (textbook inverse CDF, NOT the reference's TreeMap rule) against the oracle's restatement of it;
    the degenerate Poisson clouds are where the two rules differ."""
    run_steps(c2(), 6000, 4, SYS, _abi.F32, seed=12, tie_first=True)
    run_steps(c2(), 3000, 4, STRAT, _abi.F64, seed=13, tie_first=True)
    run_steps(ALL["bernoulli"](), 2500, 3, SYS, _abi.F64, seed=14, tie_first=True)


def test_step_parity_two_level_sums_on_a_small_cloud(monkeypatch):
    """Clouds of at most 1024 tiles run K2 / K3 "flat" (no super tiles, no atomics); CSSM_FLAT_MAX_NT=0 forces the
    two-level tables of the large clouds onto a small one: same bits either way."""
    monkeypatch.setenv("CSSM_FLAT_MAX_NT", "0")
    run_steps(c2(), 5000, 3, SYS, _abi.F32, seed=15)
    run_steps(c2(), 2300, 3, STRAT, _abi.F64, seed=16)
    run_steps(c2(), 2300, 3, MULTI, _abi.F64, seed=17)


def test_step_parity_euler():
    run_steps(c2().withStepMode(_abi.STEP_EULER), 1500, 4, SYS, _abi.F64, seed=6)
    run_steps(ALL["c4"]().withStepMode(_abi.STEP_EULER), 1500, 4, SYS, _abi.F32, seed=7)


@pytest.mark.parametrize("kind", [SYS, STRAT, MULTI])
def test_resample_bit_exact(kind):
    rng = np.random.default_rng(5)
    cases = []
    for n in (1, 2, 7, 2048, 2049, 5000, 70001):
        cases.append(rng.random(n))                                   # SamplingTest.scala: weights in [0, 1]
        cases.append(np.exp(rng.normal(0, 2, n)))                     # exp(N(0, 2^2)), may exceed 1
        w = np.full(n, 1e-30); w[n // 3] = 1.0
        cases.append(w)                                               # degenerate: one particle carries everything
        w = rng.random(n); w[rng.random(n) < 0.7] = 0.0
        if w.sum() == 0: w[0] = 1.0
        cases.append(w)                                               # runs of exact zeros -> duplicate TreeMap keys
    cases.append(np.ones(6400))                                       # src/bench/scala/Resampling.scala:17
    for w in cases:
        n = w.size
        u = rng.random(1 if kind == SYS else n)
        got = cs.resampling.ancestors(kind, w, u)
        want = oracle.resample(kind, w, u, oracle.ORDER_DEVICE)
        np.testing.assert_array_equal(got, want)
        assert got.size == n                                          # the reference's only pinned property
        if kind != MULTI:
            assert np.all(np.diff(got) >= 0)


def test_resample_large_bit_exact():
    rng = np.random.default_rng(9)
    n = 1 << 20
    w = np.exp(rng.normal(0, 3, n))
    w[100000:400000] = 0.0   # a run of zeros spanning many scan tiles
    for kind in (SYS, STRAT):
        u = rng.random(1 if kind == SYS else n)
        got = cs.resampling.ancestors(kind, w, u)
        np.testing.assert_array_equal(got, oracle.resample(kind, w, u, oracle.ORDER_DEVICE))
        ref = oracle.resample(kind, w, u, oracle.ORDER_REFERENCE)
        assert np.mean(ref != got) < 1e-3


def test_resample_one_heavy_particle_and_a_long_run_of_repeated_keys():
    """SURVEY 8(d)'s degenerate case: one weight 1, the rest 1e-30.  Under the TreeMap rule every output goes to the last
    particle of the run of repeated keys behind the heavy one -- a run that crosses many tiles, owned by a particle whose
    offspring fill many passes of the exact scan path (the per-thread memo of the last walk; before it 2^24 outputs took
    seconds).  Ancestors equal the oracle's; a second heavy particle splits the run."""
    rng = np.random.default_rng(4)
    for n, heavy in ((1 << 14, (5000,)), ((1 << 14) + 333, (17, 9000)), (40000, (39999,))):
        w = np.full(n, 1e-30)
        for h in heavy:
            w[h] = 1.0
        for kind in (SYS, STRAT):
            u = rng.random(1 if kind == SYS else n)
            got = cs.resampling.ancestors(kind, w, u)
            np.testing.assert_array_equal(got, oracle.resample(kind, w, u, oracle.ORDER_DEVICE))
            if len(heavy) == 1 and heavy[0] < n - 1:
                assert np.mean(got == n - 1) > 0.99  # the last particle of the cloud takes (nearly) everything


def test_lgcp_step_parity():
    mod = c3(precision=2)
    for dtype in (_abi.F64, _abi.F32):
        rng = np.random.default_rng(8)
        N, d = 1200, 1
        h = cs.GpuFilterHandle(mod, STRAT, N, dtype=dtype, seed=1)
        z0 = rng.standard_normal((d, N))
        h.init_injected(0.0, z0)
        x = h.get_particles()
        tp = 0.0
        for t in (0.0, 0.13, 0.2):
            n = oracle.lgcp_nsub(t - tp, 2)
            assert n == h.n_substeps(t - tp)
            z = rng.standard_normal((max(n, 1), d, N))
            u = rng.random(N)
            g = h.step_injected(t, 1.0, z if n > 0 else None, u)
            o = oracle.Oracle(mod); o.reset(N)
            r = o.step(x, tp, t, 1.0, z, u, STRAT, oracle.device_order(dtype))
            rel_close(g["x_prop"], r["x_prop"], TOL[dtype], "lgcp state")
            rel_close(g["logw"], r["logw"], TOL[dtype], "lgcp log-weights")  # north_star: 1e-5 fp32 / 1e-12 fp64, as everywhere
            mx = float(np.max(g["logw"]))
            w1 = oracle.w1(g["logw"], mx, oracle.device_order(dtype))
            np.testing.assert_array_equal(g["anc"], oracle.resample(STRAT, w1, u))
            x = h.get_particles()
            tp = t
        h.close()


def test_full_filter_fp64_matches_oracle_end_to_end():
    """C1 (1000 particles x 500 observations), every step with injected noise: the accumulated
    log-likelihood of the GPU equals the oracle's own end-to-end run (no hand-over of device
    intermediates), reference summation order included."""
    mod = c1()
    N, T, d = 1000, 500, 1
    orc = oracle.Oracle(mod)
    t, y, _ = orc.simulate(T, 0.1, 1)
    rng = np.random.default_rng(0)
    h = cs.GpuFilterHandle(mod, SYS, N, dtype=_abi.F64, seed=1)
    z0 = rng.standard_normal((d, N))
    h.init_injected(t[0], z0)
    xo = orc.init_state(z0)
    o_dev, o_ref = oracle.Oracle(mod), oracle.Oracle(mod)
    o_dev.reset(N); o_ref.reset(N)
    xr = xo.copy()
    tp = t[0]
    n_diff = 0
    for s in range(T):
        z, u = rng.standard_normal((d, N)), rng.random(1)
        g = h.step_injected(t[s], float(y[s]), z, u, want=("anc",))
        r = o_dev.step(xo, tp, t[s], float(y[s]), z, u, SYS, oracle.ORDER_DEVICE)
        rr = o_ref.step(xr, tp, t[s], float(y[s]), z, u, SYS, oracle.ORDER_REFERENCE)
        n_diff += int(np.sum(g["anc"] != r["anc"]))
        xo, xr, tp = r["x_out"], rr["x_out"], t[s]
    ll_gpu, _ = h.get_ll()
    assert n_diff == 0
    assert abs(ll_gpu - o_dev._ll) <= 1e-12 * abs(o_dev._ll)
    assert abs(ll_gpu - o_ref._ll) <= 1e-9 * abs(o_ref._ll)
    h.close()


@pytest.mark.parametrize("name", sorted(ALL))
@pytest.mark.parametrize("dtype", [_abi.F64, _abi.F32])
@pytest.mark.parametrize("kind", [SYS, STRAT])
def test_series_kernel_equals_three_launch_path(name, dtype, kind):
    """The single-launch series kernel (small clouds, PMMH) and the three-launch step are two
    schedules of the same arithmetic: per-step log-likelihood, ESS, the final cloud and a further
    stepFilter on it must agree bit for bit.  The three-launch path is the one pinned stage by
    stage against the oracle above, so this extends that parity to the series kernel.  Ragged
    cloud (not a multiple of the tile), several tiles, missing observations."""
    mod = ALL[name]()
    orc = oracle.Oracle(mod)
    N, T = 5 * 512 + 77, 30
    t, y, _ = orc.simulate(T, 0.1, 21)
    has = np.ones(T, dtype=np.uint8)
    has[[4, 5, 17]] = 0
    out = []
    for mode in (_abi.SERIES_THREE_LAUNCH, _abi.SERIES_SINGLE_LAUNCH):
        h = cs.GpuFilterHandle(mod, kind, N, dtype=dtype, seed=7)
        h.series_mode(mode)
        h.load_series(t, y, has)
        ll, lls, ess = h.ll_resident(steps=True)
        n_launch = h.last_launches()
        x = h.get_particles()
        ll2, ess2 = h.step(t[-1] + 0.1, float(y[-1]))   # the handle carries on from the same state
        x2 = h.get_particles()
        ll_again = h.ll_resident()                        # a second evaluation draws fresh noise (new epoch)
        out.append((ll, lls, ess, x, ll2, ess2, x2, ll_again, n_launch))
        h.close()
    a, b = out
    assert b[8] == 2 and a[8] > T            # init + ONE launch against init + 3 per observed step
    assert a[0] == b[0] and a[4] == b[4] and a[5] == b[5] and a[7] == b[7]
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(a[2], b[2])
    np.testing.assert_array_equal(a[3], b[3])
    np.testing.assert_array_equal(a[6], b[6])
    assert np.isfinite(a[0]) and a[7] != a[0]


@pytest.mark.parametrize("name", sorted(ALL))
@pytest.mark.parametrize("N", [100, 777, 1000, 1024])
@pytest.mark.parametrize("dtype,kind", [(_abi.F32, SYS), (_abi.F64, SYS), (_abi.F32, STRAT), (_abi.F64, STRAT)])
def test_one_block_series_kernel_equals_three_launch_path(name, N, dtype, kind):
    """Clouds of the reference's own examples (100 .. 1000 particles, examples/DetermineParameters.scala:70,
    examples/Filtering.scala:24) run the whole llFilter in ONE block (k_series_one): no grid-wide exchange.  Same bits as
    the three-launch step -- per-step log-likelihood and ESS, the final cloud, a further stepFilter on it -- on ragged
    clouds, with missing observations, and with extreme observations that make almost every weight vanish."""
    mod = ALL[name]()
    orc = oracle.Oracle(mod)
    T = 24
    t, y, _ = orc.simulate(T, 0.1, 5)
    has = np.ones(T, dtype=np.uint8)
    has[[0, 7, 8]] = 0
    out = []
    for mode in (_abi.SERIES_THREE_LAUNCH, _abi.SERIES_SINGLE_LAUNCH):
        h = cs.GpuFilterHandle(mod, kind, N, dtype=dtype, seed=3)
        h.series_mode(mode)
        h.load_series(t, y, has)
        ll, lls, ess = h.ll_resident(steps=True)
        n_launch = h.last_launches()
        x = h.get_particles()
        ll2, ess2 = h.step(t[-1] + 0.1, float(y[-1]))
        x2 = h.get_particles()
        out.append((ll, lls, ess, x, ll2, ess2, x2, n_launch))
        h.close()
    a, b = out
    assert b[7] == 2 and a[7] > T
    assert a[0] == b[0] and a[4] == b[4] and a[5] == b[5]
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(a[2], b[2])
    np.testing.assert_array_equal(a[3], b[3])
    np.testing.assert_array_equal(a[6], b[6])


def test_one_block_series_kernel_degenerate_weights():
    """Runs of repeated keys inside the single tile of the one-block kernel (extreme observations)."""
    mod = c5()
    N, T = 1000, 6
    t = 0.1 * np.arange(T)
    y = np.array([0.3, 75.0, -60.0, 0.0, 40.0, 0.1])
    out = []
    for mode in (_abi.SERIES_THREE_LAUNCH, _abi.SERIES_SINGLE_LAUNCH):
        for dtype in (_abi.F32, _abi.F64):
            h = cs.GpuFilterHandle(mod, SYS, N, dtype=dtype, seed=11)
            h.series_mode(mode)
            h.load_series(t, y)
            ll, lls, ess = h.ll_resident(steps=True)
            out.append((ll, lls, ess, h.get_particles()))
            h.close()
    for i in (0, 1):
        assert out[i][0] == out[i + 2][0]
        np.testing.assert_array_equal(out[i][1], out[i + 2][1])
        np.testing.assert_array_equal(out[i][2], out[i + 2][2])
        np.testing.assert_array_equal(out[i][3], out[i + 2][3])


def test_series_kernel_degenerate_weights_and_large_grid():
    """Duplicate-key runs that cross tiles inside the series kernel (extreme observations make
    almost every weight vanish), on a cloud that needs more than one block per SM."""
    mod = c5()
    N, T = 200 * 512, 6
    t = 0.1 * np.arange(T)
    y = np.array([0.3, 75.0, -60.0, 0.0, 40.0, 0.1])
    out = []
    for mode in (_abi.SERIES_THREE_LAUNCH, _abi.SERIES_SINGLE_LAUNCH):
        h = cs.GpuFilterHandle(mod, SYS, N, dtype=_abi.F32, seed=11)
        h.series_mode(mode)
        h.load_series(t, y)
        ll, lls, ess = h.ll_resident(steps=True)
        out.append((ll, lls, ess, h.get_particles()))
        h.close()
    assert out[0][0] == out[1][0]
    np.testing.assert_array_equal(out[0][1], out[1][1])
    np.testing.assert_array_equal(out[0][2], out[1][2])
    np.testing.assert_array_equal(out[0][3], out[1][3])


@pytest.mark.parametrize("name,N", [("c2", 300 * 512 + 77), ("c4", 300 * 512 + 77), ("c2", (1 << 18) + 6149), ("c5", (1 << 19) + 3)])
@pytest.mark.parametrize("dtype", [_abi.F64, _abi.F32])
@pytest.mark.parametrize("kind", [SYS, STRAT])
def test_series_multi_tile_kernel_equals_three_launch_path(name, N, dtype, kind):
    """Mid-size clouds: more tiles than resident blocks, every block loops over several tiles in
    each stage of the single-launch kernel (512- and 2048-particle tiles, compile-time and generic
    latent dimension).  Same bits as the three-launch step."""
    mod = ALL[name]()
    orc = oracle.Oracle(mod)
    T = 10
    t, y, _ = orc.simulate(T, 0.1, 33)
    has = np.ones(T, dtype=np.uint8)
    has[[3, 4]] = 0
    out = []
    for mode in (_abi.SERIES_THREE_LAUNCH, _abi.SERIES_SINGLE_LAUNCH):
        h = cs.GpuFilterHandle(mod, kind, N, dtype=dtype, seed=13)
        h.series_mode(mode)
        h.load_series(t, y, has)
        ll, lls, ess = h.ll_resident(steps=True)
        n_launch = h.last_launches()
        x = h.get_particles()
        ll2, ess2 = h.step(t[-1] + 0.1, float(y[-1]))
        out.append((ll, lls, ess, x, ll2, ess2, n_launch))
        h.close()
    a, b = out
    assert b[6] == 2 and a[6] > T
    assert a[0] == b[0] and a[4] == b[4] and a[5] == b[5]
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(a[2], b[2])
    np.testing.assert_array_equal(a[3], b[3])


@pytest.mark.parametrize("name", ["c2", "c4", "bernoulli", "beta", "student_t"])
@pytest.mark.parametrize("dtype", [_abi.F64, _abi.F32])
def test_intervals_on_device_match_the_sorted_cloud(name, dtype):
    """ParticleFilter.getIntervals (model/ParticleFilter.scala:415-424) by radix select on the
    device against the oracle's literal sort of the SAME cloud (read back from the device): state
    intervals are elements of the cloud, bit for bit; mean and eta within the dtype tolerance."""
    from composablestatespacemodels_b200 import Filter, Resampling, Data, ParticleFilter
    mod = ALL[name]()
    orc = oracle.Oracle(mod)
    N, T = 3000, 6
    t, y, _ = orc.simulate(T, 0.1, 17)
    flt = Filter(mod, Resampling.systematicResampling, dtype=dtype, seed=9)
    s = flt.initialiseState(N, t[0])
    for k in range(T):
        s = flt.stepFilter(s, Data(t[k], y[k]))
    x = s.particles.T                                  # [d][N], what the device holds (as doubles)
    ref = orc.intervals(x, s.t, 0.975)
    out = ParticleFilter.getIntervals(mod, s)
    np.testing.assert_array_equal([ci.lower for ci in out.stateIntervals], ref["lower"])
    np.testing.assert_array_equal([ci.upper for ci in out.stateIntervals], ref["upper"])
    tol = TOL[dtype]
    rel_close(out.state, ref["mean"], 1e-12, "mean state")
    rel_close([out.eta, out.etaIntervals.lower, out.etaIntervals.upper], ref["eta"], tol, "eta and its interval")
    with pytest.raises(cs._abi.CssmError):
        s._handle.intervals(s.t, 1.0)                  # index = n: the reference throws IndexOutOfBounds
    flt.close()


FORECAST_MODELS = ["c1", "c2", "c4", "c5", "bernoulli", "student_t", "zip", "beta"]


def _filtered(mod, N, dtype, T=4, seed=9):
    from composablestatespacemodels_b200 import Filter, Resampling, Data
    orc = oracle.Oracle(mod)
    t, y, _ = orc.simulate(T, 0.1, 17)
    flt = Filter(mod, Resampling.systematicResampling, dtype=dtype, seed=seed)
    s = flt.initialiseState(N, t[0])
    for k in range(T):
        s = flt.stepFilter(s, Data(t[k], y[k]))
    return flt, s, orc


@pytest.mark.parametrize("name", FORECAST_MODELS)
@pytest.mark.parametrize("dtype", [_abi.F64, _abi.F32])
def test_forecast_summaries_match_the_forecast_cloud(name, dtype):
    """cssm_filter_forecast (getMeanForecast, model/ParticleFilter.scala:394-412): the summaries computed on
    the device against a literal sort of the forecast cloud read back from it -- intervals are elements of
    the cloud, bit for bit; and the cloud is self-consistent: gamma = f(x1, t), eta = link(gamma), observations
    in the support of the model's distribution.  The filter's own cloud is not touched."""
    mod = ALL[name]()
    N = 4000
    flt, s, orc = _filtered(mod, N, dtype)
    tol = TOL[dtype]
    before = s.particles
    tf = s.t + 0.35
    r = s._handle.forecast(tf, 0.975)
    c = s._handle.forecast_cloud()
    np.testing.assert_array_equal(s.particles, before)
    idx = int(np.floor(0.975 * N))
    xs = np.sort(c["x"], axis=1)
    np.testing.assert_array_equal(r["lower"], xs[:, N - idx - 1])
    np.testing.assert_array_equal(r["upper"], xs[:, idx - 1])
    rel_close(r["mean"], c["x"].mean(axis=1), 1e-9, "forecast mean state")
    for key, col in (("eta", "eta"), ("obs", "obs2")):
        o = np.sort(c[col])
        assert r[key][1] == o[N - idx] and r[key][2] == o[idx], key
        rel_close(r[key][0], c[col].mean(), 1e-9, f"mean {key}")
    rel_close(c["gamma"], orc.f(c["x"], tf), tol, "gamma = f(x1, t)")
    rel_close(c["eta"], [orc.link(g) for g in c["gamma"]], tol, "eta = link(gamma)")
    for col in ("obs", "obs2"):
        o = c[col]
        assert np.all(np.isfinite(o))
        if name in ("c1", "c2", "c4", "zip"):
            assert np.all(o >= 0) and np.all(o == np.floor(o))
        if name == "bernoulli":
            assert set(np.unique(o)) <= {0.0, 1.0}
        if name == "beta":
            assert np.all((o >= 0) & (o <= 1))
    assert not np.array_equal(c["obs"], c["obs2"])
    with pytest.raises(cs._abi.CssmError):
        s._handle.forecast(s.t - 1.0)
    flt.close()


@pytest.mark.parametrize("name", FORECAST_MODELS + ["poisson_big"])
def test_forecast_agrees_with_the_oracle_in_distribution(name):
    """getForecast / getMeanForecast against the oracle's NumPy restatement started from the SAME filtering cloud:
    independent RNG streams, so means agree within 6 standard errors and every interval end point lies between the
    oracle's empirical quantiles at p -+ 6 sd of an empirical CDF value."""
    from composablestatespacemodels_b200 import ParticleFilter
    mod = {**ALL, **EXTRA}[name]()
    N = 200000
    flt, s, orc = _filtered(mod, N, _abi.F32)
    x = s.particles.T
    tf = s.t + 0.5
    out = ParticleFilter.getMeanForecast(s, mod, tf, 0.975)
    cloud = ParticleFilter.getForecast(s, mod, tf)
    assert cloud.sdeState.shape == (N, mod.dimension)
    rng = np.random.default_rng(5)
    ref = orc.forecast(x, s.t, tf, rng, 0.975)
    ref2 = orc.forecast(x, s.t, tf, rng, 0.975)   # a second oracle run gauges the Monte-Carlo error itself

    def close_mean(a, sample, what):
        se = np.std(sample) * np.sqrt(2.0 / N) + 1e-12
        assert abs(a - np.mean(sample)) <= 6 * se, f"{what}: {a} vs {np.mean(sample)} (se {se})"

    def within_quantiles(q, sample, p, what):
        dlt = 6 * np.sqrt(2 * p * (1 - p) / N)
        lo, hi = np.quantile(sample, max(p - dlt, 0.0), method="lower"), np.quantile(sample, min(p + dlt, 1.0), method="higher")
        assert lo <= q <= hi, f"{what}: {q} outside [{lo}, {hi}]"

    for k in range(mod.dimension):
        close_mean(out.state[k], ref["x"][k], f"state mean {k}")
        within_quantiles(out.stateIntervals[k].lower, ref["x"][k], 0.025, f"state lower {k}")
        within_quantiles(out.stateIntervals[k].upper, ref["x"][k], 0.975, f"state upper {k}")
    close_mean(out.eta, ref["eta"], "mean eta")
    within_quantiles(out.etaIntervals.lower, ref["eta"], 0.025, "eta lower")
    within_quantiles(out.etaIntervals.upper, ref["eta"], 0.975, "eta upper")
    close_mean(out.obs, ref["obs"], "mean observation")
    within_quantiles(out.obsIntervals.lower, ref["obs"], 0.025, "observation lower")
    within_quantiles(out.obsIntervals.upper, ref["obs"], 0.975, "observation upper")
    # the observation distribution beyond its mean: variance, and P(obs == 0) for the count models
    vg, vr, vr2 = np.var(cloud.observation), np.var(ref["obs"]), np.var(ref2["obs"])
    assert abs(vg - vr) <= 8 * abs(vr - vr2) + 0.03 * vr, (vg, vr, vr2)
    if name in ("c1", "c2", "c4", "zip", "poisson_big"):
        p0g, p0r = np.mean(cloud.observation == 0), np.mean(ref["obs"] == 0)
        assert abs(p0g - p0r) <= 6 * np.sqrt(2 * max(p0r, 1e-4) / N), (p0g, p0r)
    flt.close()


def test_forecast_chain_continues_from_the_forecast_cloud():
    """SimulateData.forecast (model/Data.scala:202-217) scans simStep over the forecast times: chain=True advances the
    forecast cloud itself.  Two chained steps of 0.3 and 0.4 agree in distribution with one step of 0.7 (the exact
    OU / Brownian transitions compose)."""
    mod = ALL["c5"]()
    N = 200000
    flt, s, orc = _filtered(mod, N, _abi.F32)
    h = s._handle
    one = h.forecast(s.t + 0.7, 0.975)
    h.forecast(s.t + 0.3, summarise=False)
    two = h.forecast(s.t + 0.7, 0.975, chain=True)
    c = h.forecast_cloud()
    sd = c["x"].std(axis=1)
    assert np.all(np.abs(one["mean"] - two["mean"]) <= 6 * sd * np.sqrt(2.0 / N))
    assert np.all(np.abs(one["upper"] - two["upper"]) <= 0.05 * sd + 1e-6)
    assert np.all(np.abs(one["lower"] - two["lower"]) <= 0.05 * sd + 1e-6)
    assert abs(one["obs"][0] - two["obs"][0]) <= 6 * c["obs2"].std() * np.sqrt(2.0 / N)
    h.step(s.t + 0.1, 0.3)                                       # the filter moves on: the forecast cloud is dropped
    with pytest.raises(cs._abi.CssmError):
        h.forecast(s.t + 0.9, 0.975, chain=True)
    flt.close()


@pytest.mark.parametrize("name,kind", [("c2", SYS), ("c4", STRAT), ("c1", MULTI)])
@pytest.mark.parametrize("dtype", [_abi.F64, _abi.F32])
def test_path_storage_follows_the_list_semantics_of_filter_interpolate(name, kind, dtype):
    """FilterInterpolate.stepInterpolate (model/ParticleFilter.scala:281-298): `x.head` is advanced and consed onto
    the path; an observed step resamples the PATHS, an unobserved one does not.  The lists are built here literally
    from what every device step returned (propagated cloud, ancestors) and compared, bit for bit, with the paths the
    device reads out of its ancestor tree."""
    mod = ALL[name]()
    orc = oracle.Oracle(mod)
    N, T, d = 1500, 7, mod.dimension
    missing = (2, 5)
    rng = np.random.default_rng(4)
    t, y, _ = orc.simulate(T, 0.1, 21)
    h = cs.GpuFilterHandle(mod, kind, N, dtype=dtype, seed=3)
    with pytest.raises(cs._abi.CssmError):
        h.get_paths()
    h.paths_enable(T)
    h.init_injected(t[0], rng.standard_normal((d, N)))
    x0 = h.get_particles()
    paths = [[x0[:, i]] for i in range(N)]                      # newest first, like List[State]
    np.testing.assert_array_equal(h.get_paths()[:, 0, :], x0.T)
    for s in range(T):
        obs = None if s in missing else float(y[s])
        g = h.step_injected(t[s], obs, rng.standard_normal((d, N)), rng.random(1 if kind == SYS else N))
        x1 = [[g["x_prop"][:, i]] + paths[i] for i in range(N)]
        paths = x1 if obs is None else [x1[a] for a in g["anc"]]
        assert h.paths_len() == s + 1
    got = h.get_paths()                                          # [N, T + 1, d], oldest first
    want = np.array([p[::-1] for p in paths])
    np.testing.assert_array_equal(got, want)
    some = np.array([N - 1, 0, 17, 17, 3])
    np.testing.assert_array_equal(h.get_paths(some), want[some])
    np.testing.assert_array_equal(got[:, -1, :], h.get_particles().T)
    with pytest.raises(cs._abi.CssmError):                       # the storage holds T steps
        h.step(t[-1] + 0.1, 1.0)
    h.ll_arrays(t, y)                                            # a whole-series call does not record ...
    assert h.paths_len() == -1                                   # ... and invalidates what was recorded
    with pytest.raises(cs._abi.CssmError):
        h.get_paths()
    h.init(t[0])                                                 # recording restarts with the next initialisation
    assert h.paths_len() == 0
    h.close()


def test_filter_interpolate_host_mirror():
    """FilterInterpolate.filterInterpolate (model/ParticleFilter.scala:300-310) as a generator transformer: same
    log-likelihood as the plain filter with the same seed, paths grow by one state per datum, the emitted states
    list their particles in reverse order (`s.particles.reverse`)."""
    from composablestatespacemodels_b200 import Filter, FilterInterpolate, Resampling, Data
    mod = ALL["c2"]()
    orc = oracle.Oracle(mod)
    T, N = 6, 2000
    t, y, _ = orc.simulate(T, 0.1, 5)
    data = [Data(t[k], None if k == 3 else y[k]) for k in range(T)]
    fi = FilterInterpolate(mod, Resampling.systematicResampling, max_steps=T, dtype=_abi.F32, seed=7)
    states = list(fi.filterInterpolate(t[0], N)(data))
    assert len(states) == T + 1 and states[0].ess == 0 and states[0].ll == 0.0
    last = states[-1]
    p = last.particles
    assert p.shape == (N, T + 1, mod.dimension)
    np.testing.assert_array_equal(p[::-1, 0, :], last._handle.get_particles().T)   # newest state first, particles reversed
    assert states[4].ll == states[3].ll and states[4].ess == states[3].ess           # the unobserved datum
    flt = Filter(mod, Resampling.systematicResampling, dtype=_abi.F32, seed=7)
    s = flt.initialiseState(N, t[0])
    for dd in data:
        s = flt.stepFilter(s, dd)
    assert s.ll == last.ll
    fi.close()
    flt.close()


def test_pilot_run_variances():
    """Streaming.pilotRun (model/Streaming.scala:19-41): the variance of the log-likelihood estimate per particle
    count falls roughly like 1/n and agrees with the oracle's own repeated filters within the F-distribution's
    99.9 % range for 40 repetitions each."""
    from composablestatespacemodels_b200 import Streaming, Resampling, Data
    mod = ALL["c1"]()
    orc = oracle.Oracle(mod)
    T, R = 60, 40
    t, y, _ = orc.simulate(T, 0.1, 3)
    data = [Data(a, b) for a, b in zip(t, y)]
    res = Streaming.pilotRun(data, mod, Resampling.systematicResampling, [100, 1600], R, dtype=_abi.F32, seed=11)
    assert [n for n, _ in res] == [100, 1600]
    v = dict(res)
    assert v[100] > 3 * v[1600] > 0
    for n in (100, 1600):
        ref = np.var(orc.filter_ll_many(n, SYS, t, y, seed=5, R=R, threads=4), ddof=1)
        assert ref / 3.2 < v[n] < ref * 3.2, (n, v[n], ref)


@pytest.mark.parametrize("dtype", [_abi.F32, _abi.F64])
def test_lgcp_philox_runs_agree_with_the_oracle(dtype):
    """FilterLgcp with the device's own noise (Philox, the call-by-call unrolled sub-step loop and its tail) against
    the oracle's independent runs of the same filter: the log-likelihood estimates agree within Monte-Carlo error
    (event gaps of 7, 13, 20 ... sub-steps, so full Philox calls and partial ones both occur)."""
    from composablestatespacemodels_b200 import FilterLgcp, Resampling, Data
    mod = c3(precision=2)
    orc = oracle.Oracle(mod)
    gaps = np.array([0.07, 0.13, 0.2, 0.05, 0.11, 0.3, 0.02, 0.09, 0.16, 0.04])
    t = np.concatenate([[0.0], np.cumsum(gaps)])
    y = np.ones_like(t)
    N, R = 20000, 8
    flt = FilterLgcp(mod, Resampling.stratifiedResampling, 2, dtype=dtype, seed=31)
    data = [Data(a, 1.0) for a in t]
    gpu = np.array([flt.llFilter(data, N) for _ in range(R)])
    flt.close()
    cpu = orc.filter_ll_many(N, STRAT, t, y, seed=9, R=R, threads=4)
    se = np.sqrt(gpu.var(ddof=1) / R + cpu.var(ddof=1) / R)
    assert np.all(np.isfinite(gpu))
    assert abs(gpu.mean() - cpu.mean()) < 5 * se + 2e-3, (gpu, cpu)
