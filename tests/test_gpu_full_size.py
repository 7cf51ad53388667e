"""GPU tests at BASELINE.json's full sizes, through properties that do not need the CPU oracle to finish a 2^24
particle run: an exact known answer (the Kalman-filter likelihood of the Normal compositions), the defining
property of systematic resampling (every offspring count within one of n*w/W), sortedness, determinism and
independence of the launch geometry."""
import math
import os

import numpy as np
import pytest
from scipy import special

import composablestatespacemodels_b200 as cs
from composablestatespacemodels_b200 import _abi
import oracle
from configs import SYS, STRAT, c2, c5
from test_oracle import kalman_loglik

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,N", [(_abi.F64, 1 << 18), (_abi.F32, 1 << 18), (_abi.F32, 1 << 24)])
def test_log_likelihood_against_the_exact_kalman_filter(dtype, N):
    """Normal + seasonal + OU (BASELINE configs[4]'s model) is linear-Gaussian: its marginal likelihood is known
    exactly.  The particle estimate exp(ll_hat) is unbiased, so on the likelihood scale independent Philox runs agree
    with it within Monte-Carlo error (north_star: "independent-RNG runs must agree within Monte Carlo standard
    error") -- and at 2^24 particles (the target cloud size) the estimate itself is within 0.02 of the exact value."""
    mod = c5()
    orc = oracle.Oracle(mod)
    T, R = 40, 6
    t, y, _ = orc.simulate(T, 0.1, 5)
    exact = kalman_loglik(mod, t, y)
    h = cs.GpuFilterHandle(mod, SYS, N, dtype=dtype, seed=21)
    h.load_series(t, y)
    est = np.array([h.ll_resident() for _ in range(R)])
    h.close()
    assert np.all(np.isfinite(est))
    lm = special.logsumexp(est) - math.log(R)
    se = np.std(np.exp(est - exact)) / math.sqrt(R)
    assert abs(math.exp(lm - exact) - 1) < 5 * se + 0.02, (est, exact)
    if N == 1 << 24:
        assert np.max(np.abs(est - exact)) < 0.02, (est, exact)


@pytest.mark.parametrize("kind", [SYS, STRAT])
def test_full_size_resampling_properties(kind):
    """cssm_resample at 2^24 weights: ancestors sorted, inside the cloud, and (systematic) every particle's number
    of offspring follows n*w/W (cumulatively within one), with the reference's repeated-key quirk."""
    from composablestatespacemodels_b200.resampling import ancestors
    n = 1 << 24
    rng = np.random.default_rng(12)
    w = np.exp(rng.normal(0.0, 2.0, n))
    w[rng.integers(0, n, 1000)] = 0.0
    u = rng.random(1 if kind == SYS else n)
    anc = ancestors(kind, w, u)
    assert anc.shape == (n,) and anc.min() >= 0 and anc.max() < n
    assert np.all(np.diff(anc) >= 0)
    counts = np.bincount(anc, minlength=n)
    # the defining property, on the cumulative counts: #{k_i <= C_j} is within one of n * C_j.  It is read where a
    # TreeMap key is not repeated: a zero weight repeats the key of its predecessor and the LAST particle of such a run
    # takes the run's offspring (model/Resampling.scala:55-57) -- the reference's quirk, kept
    nxt_zero = np.append(w[1:] == 0.0, False)
    cc, ce = np.cumsum(counts), np.cumsum(w) * (n / w.sum())
    assert np.all(np.abs(cc - ce)[~nxt_zero] <= 1.0 + 1e-3)
    assert cc[-1] == n
    assert np.all(counts[(w == 0.0) & nxt_zero] == 0)           # inside a run of zeros: nothing
    heirs = (w == 0.0) & ~nxt_zero                               # the last zero of a run inherits its predecessor's offspring
    assert counts[heirs].sum() > 0
    # and bit for bit against the oracle's sequential search (a plain loop over the 2^24 weights: about a second)
    np.testing.assert_array_equal(anc, oracle.resample(kind, w, u))


def test_full_size_filter_is_deterministic_and_geometry_independent(monkeypatch):
    """2^24 particles x 12 observations of the target model: the same seed gives the same bits twice, and the same
    bits again with 512-particle tiles instead of 2048 (exact integer sums: no result depends on the tiling)."""
    mod = c2()
    orc = oracle.Oracle(mod)
    t, y, _ = orc.simulate(12, 0.1, 1)
    N = 1 << 24

    def run():
        h = cs.GpuFilterHandle(mod, SYS, N, dtype=_abi.F32, seed=4)
        h.load_series(t, y)
        ll, lls, ess = h.ll_resident(steps=True)
        m = h.mean_state()
        h.close()
        return ll, lls, ess, m

    a = run()
    b = run()
    monkeypatch.setenv("CSSM_TILE_ITEMS", "2")
    c = run()
    for other in (b, c):
        assert a[0] == other[0]
        np.testing.assert_array_equal(a[1], other[1])
        np.testing.assert_array_equal(a[2], other[2])
        np.testing.assert_allclose(a[3], other[3], rtol=1e-12)   # meanState: fp64 atomics, the order of the adds varies
    assert np.isfinite(a[0]) and np.all(a[2] >= 1) and np.all(a[2] <= N)


def test_tie_rule_and_the_drift_of_the_reference_estimate():
    """What the reference's TreeMap rule does at GPU cloud sizes: under a repeated cumulative
    weight the LAST particle wins (model/Resampling.scala:52-58), so offspring of good particles land on particles of
    negligible weight; with the degenerate weights of the 7-dimensional Poisson model the log-likelihood estimate
    drifts DOWN as the cloud grows.  The library reproduces that by default (ancestors bit-exact with the oracle's
    restatement); CSSM_TIE_FIRST, the textbook rule, gives estimates that agree across cloud sizes."""
    mod = c2()
    orc = oracle.Oracle(mod)
    t, y, _ = orc.simulate(12, 0.1, 1)

    def est(N, rule, R=4):
        h = cs.GpuFilterHandle(mod, SYS, N, dtype=_abi.F32, seed=4)
        h.set_tie_rule(rule)
        h.load_series(t, y)
        v = np.array([h.ll_resident() for _ in range(R)])
        h.close()
        return v

    first_small, first_big = est(1 << 18, _abi.TIE_FIRST, 8), est(1 << 24, _abi.TIE_FIRST)
    ref_small, ref_big = est(1 << 18, _abi.TIE_REFERENCE, 8), est(1 << 24, _abi.TIE_REFERENCE)
    se = first_small.std(ddof=1) / math.sqrt(8)
    assert abs(first_small.mean() - first_big.mean()) < 6 * se + 0.01, (first_small, first_big)
    assert ref_big.mean() < ref_small.mean() - 0.05, (ref_small, ref_big)        # the reference rule: measured -0.12
    assert ref_big.mean() < first_big.mean() - 0.05


def test_bench_line_carries_the_contract_keys():
    """bench.py on a small workload (c1, one cooperative launch per llFilter): one JSON line with the keys the bench
    contract names -- roofline (bound, achieved, peak, unit, frac, traffic), e2e with the bytes copied per step,
    gpu_launches, clocks -- and a log-likelihood that is a number."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "c1", "--steps", "2", "--warmup", "3", "--no-cpu"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in j, k
    assert j["n_gpus"] == 1 and j["steps"] == 2 and j["warmup"] == 3 and j["value"] > 0 and j["gpu_launches"] > 0
    assert "workload" in j["config"] and "model" not in j["config"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in j["roofline"], k
    assert abs(j["roofline"]["frac"] - j["roofline"]["achieved"] / j["roofline"]["peak"]) < 1e-12
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in j["e2e"], k
    assert j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] > 0
    assert set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert math.isfinite(j["log_likelihood_mean"])


def test_injected_noise_step_parity_at_the_baseline_cloud_size():
    """BASELINE.json configs[1]: the composed Poisson + seasonal + OU model at 2^20 particles, fp32, systematic -- every
    stage of two stepFilters against the oracle on identical injected noise (states / log-weights 1e-5, w1, ESS and
    ancestors bit-exact), not only the resampler and the likelihood as the other full-size tests do."""
    from test_gpu_parity import run_steps
    run_steps(c2(), 1 << 20, 2, SYS, _abi.F32, seed=21)


def test_sharded_eight_virtual_ranks_with_the_large_tiles(monkeypatch):
    """BASELINE.json configs[4]'s model on eight (virtual) ranks at the tile size large clouds use (2048 particles, two-level
    sum tables): 8 x 2^19 particles, same bits as the unsharded filter of 2^22 and as the oracle's resampler."""
    from test_gpu_sharded import run_pair
    from configs import c5
    monkeypatch.setenv("CSSM_K3_FAST", "1")  # the certified scan also on clouds of 256 tiles per rank (default: more than 1024)
    stats = run_pair(c5(), 1 << 19, 8, 2, SYS, _abi.F32, seed=33)
    # round 2: the certified fp64 scan runs on every rank (tiles whose outputs fall into the rank's own slots); the last
    # tile of a rank and tiles that write into a peer's slots go to the exact path
    assert all(fast > 0 for fast, _ in stats), stats
    assert sum(f for f, _ in stats) > 4 * sum(e for _, e in stats), stats


@pytest.mark.parametrize("name", ["c2", "c5"])
def test_certified_scan_equals_the_exact_path(name):
    """The scan + search of large fp32 clouds first runs in fp64 with certified counts and hands undecided tiles to the
    exact 128-bit path (include/cssm.h, cssm_filter_scan_mode).  Same seeds through AUTO and through EXACT: the
    log-likelihood, every per-step value and the final cloud have the same bits; the fast path settles most tiles and the
    fallback is exercised too."""
    from configs import c5
    mod = c2() if name == "c2" else c5()
    t, y, _ = oracle.Oracle(mod).simulate(25, 0.1, 3)
    N = (1 << 21) + 4096 + 77
    out = []
    for mode in (_abi.SCAN_AUTO, _abi.SCAN_EXACT):
        h = cs.GpuFilterHandle(mod, SYS, N, dtype=_abi.F32, seed=6)
        h.scan_mode(mode)
        h.load_series(t, y)
        ll, lls, ess = h.ll_resident(steps=True)
        out.append((ll, lls, ess, h.get_particles(), h.scan_stats()))
        h.close()
    a, e = out
    assert a[0] == e[0]
    np.testing.assert_array_equal(a[1], e[1])
    np.testing.assert_array_equal(a[2], e[2])
    np.testing.assert_array_equal(a[3], e[3])
    fast, exact = a[4]
    assert e[4][0] == 0                                  # EXACT never takes the fast path
    assert fast > 5 * exact > 0, (fast, exact)           # most tiles certified, some undecided


def test_certified_scan_with_one_particle_that_takes_nearly_everything():
    """Observations far in the tail on a cloud large enough for the certified scan (more than 1024 tiles): one particle has
    millions of offspring -- a stretch of outputs without a head that both paths fill directly instead of window by window.
    AUTO and EXACT: same log-likelihood, ESS and final cloud, bit for bit; and the steps are not slower than ordinary ones
    by more than the fill costs (a cliff here was 3 s per step at 2^24 before the fill existed)."""
    import time
    from configs import c5
    mod = c5()
    N = (1 << 21) + 4096 + 77
    t = 0.1 * np.arange(6)
    y = np.array([0.3, 75.0, -60.0, 0.0, 40.0, 0.1])
    out = []
    for mode in (_abi.SCAN_AUTO, _abi.SCAN_EXACT):
        h = cs.GpuFilterHandle(mod, SYS, N, dtype=_abi.F32, seed=12)
        h.scan_mode(mode)
        h.load_series(t, y)
        h.ll_resident()
        t0 = time.perf_counter()
        ll, lls, ess = h.ll_resident(steps=True)
        dt = time.perf_counter() - t0
        out.append((ll, lls, ess, h.get_particles(), h.scan_stats(), dt))
        h.close()
    a, e = out
    assert a[4][0] > 0 and e[4][0] == 0
    np.testing.assert_array_equal(a[2], e[2])
    assert min(a[2]) <= 3, a[2]          # the steps really are degenerate
    assert a[5] < 0.5 and e[5] < 0.5, (a[5], e[5])
    h1 = cs.GpuFilterHandle(mod, SYS, N, dtype=_abi.F32, seed=12)
    h1.scan_mode(_abi.SCAN_AUTO)
    h1.load_series(t, y)
    h2 = cs.GpuFilterHandle(mod, SYS, N, dtype=_abi.F32, seed=12)
    h2.scan_mode(_abi.SCAN_EXACT)
    h2.load_series(t, y)
    r1, r2 = h1.ll_resident(steps=True), h2.ll_resident(steps=True)
    assert r1[0] == r2[0]
    np.testing.assert_array_equal(r1[1], r2[1])
    np.testing.assert_array_equal(h1.get_particles(), h2.get_particles())
    h1.close()
    h2.close()
