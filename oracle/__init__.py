"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see the header of cssm_oracle.cpp).

PARITY UNPINNED: the reference is Scala, no JVM exists here, and its tests pin no numeric value
of this path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from composablestatespacemodels_b200 import _abi  # descriptor struct definitions only

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcssm_oracle.so")
ORDER_REFERENCE, ORDER_DEVICE, ORDER_DEVICE_F32 = 0, 1, 2
TIE_FIRST = 8  # flag on the order of resample(): textbook inverse CDF, no TreeMap duplicate-key rule (not the reference)


def device_order(dtype):
    """The device definition that matches a filter dtype (F32 filters evaluate w1 in fp32)."""
    return ORDER_DEVICE_F32 if dtype == _abi.F32 else ORDER_DEVICE
_lib = None

dp, ip, i64p, u8p = _abi.c_double_p, _abi.c_int32_p, _abi.c_int64_p, _abi.c_uint8_p
MD = C.POINTER(_abi.ModelDesc)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        l = C.CDLL(_SO)
        sig = {
            "orc_exp_det": ([C.c_double], C.c_double),
            "orc_expf_det": ([C.c_float], C.c_float),
            "orc_fix96": ([C.c_double, _abi.c_uint64_p, _abi.c_uint64_p], None),
            "orc_unfix96": ([C.c_uint64, C.c_uint64], C.c_double),
            "orc_dbl128": ([C.c_uint64, C.c_uint64, C.c_int], C.c_double),
            "orc_dim": ([MD], C.c_int),
            "orc_init_state": ([MD, C.c_int64, dp, dp], None),
            "orc_propagate": ([MD, C.c_int64, C.c_double, dp, dp, dp], None),
            "orc_f": ([MD, C.c_int64, dp, C.c_double, dp], None),
            "orc_loglik_1": ([MD, C.c_double, C.c_double], C.c_double),
            "orc_loglik": ([MD, C.c_int64, dp, C.c_double, dp], None),
            "orc_max": ([C.c_int64, dp], C.c_double),
            "orc_w1": ([C.c_int64, dp, C.c_double, C.c_int, dp], None),
            "orc_total": ([C.c_int64, dp, C.c_int], C.c_double),
            "orc_ll_ess": ([C.c_int64, dp, C.c_double, C.c_int, dp, ip], None),
            "orc_resample": ([C.c_int, C.c_int, C.c_int64, dp, dp, ip, i64p], C.c_int),
            "orc_resample_treemap": ([C.c_int, C.c_int64, dp, dp, ip], C.c_int),
            "orc_gather": ([C.c_int64, C.c_int, dp, ip, dp], None),
            "orc_step_filter": ([MD, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, dp, dp, dp,
                                 dp, dp, dp, ip, dp, dp, ip], C.c_int),
            "orc_lgcp_nsub": ([C.c_double, C.c_int], C.c_int64),
            "orc_step_lgcp": ([MD, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, dp, dp, dp, dp, dp, dp, ip, dp, dp,
                               ip], C.c_int),
            "orc_filter_ll": ([MD, C.c_int64, C.c_int, C.c_int64, dp, dp, u8p, C.c_uint64, C.c_int, dp, ip], C.c_double),
            "orc_filter_ll_many": ([MD, C.c_int64, C.c_int, C.c_int64, dp, dp, u8p, C.c_uint64, C.c_int, C.c_int, C.c_int, dp],
                                   None),
            "orc_simulate": ([MD, C.c_int64, C.c_double, C.c_uint64, dp, dp, dp], None),
            "orc_link": ([MD, C.c_double], C.c_double),
            "orc_intervals": ([MD, C.c_int64, dp, C.c_double, C.c_double, dp, dp, dp, dp], C.c_int),
        }
        for name, (args, res) in sig.items():
            fn = getattr(l, name)
            fn.argtypes, fn.restype = args, res
        _lib = l
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """The reference algorithm for one parameterised model (a `Model` of the host package)."""

    def __init__(self, mod):
        self.mod = mod
        self.desc, self._keep = mod.desc()
        self.d = mod.dimension
        self.L = lib()

    def init_state(self, z0):
        z0 = f64(z0)
        N = z0.shape[1]
        x = np.empty((self.d, N))
        self.L.orc_init_state(C.byref(self.desc), N, _d(z0), _d(x))
        return x

    def propagate(self, x, z, dt):
        x, z = f64(x), f64(z)
        out = np.empty_like(x)
        self.L.orc_propagate(C.byref(self.desc), x.shape[1], dt, _d(x), _d(z), _d(out))
        return out

    def f(self, x, t):
        x = f64(x)
        g = np.empty(x.shape[1])
        self.L.orc_f(C.byref(self.desc), x.shape[1], _d(x), t, _d(g))
        return g

    def loglik(self, gamma, y):
        gamma = f64(gamma)
        w = np.empty_like(gamma)
        self.L.orc_loglik(C.byref(self.desc), gamma.size, _d(gamma), y, _d(w))
        return w

    def step(self, x, t_prev, t, y, z, u, resample_kind, order=ORDER_DEVICE):
        """One stepFilter (y None = no observation).  Returns a dict like the GPU hook does."""
        x, z = f64(x), f64(z)
        N = x.shape[1]
        u = None if u is None else f64(np.atleast_1d(u))
        xp, lw, w1 = np.empty_like(x), np.empty(N), np.empty(N)
        anc, xo = np.empty(N, dtype=np.int32), np.empty_like(x)
        ll, ess = C.c_double(getattr(self, "_ll", 0.0)), C.c_int32(getattr(self, "_ess", N))
        if self.mod.obs_kind == _abi.OBS_LGCP:
            self.L.orc_step_lgcp(C.byref(self.desc), N, resample_kind, order, t_prev, t, _d(x), _d(z), _d(u), _d(xp), _d(lw),
                                 _d(w1), _i(anc), _d(xo), C.byref(ll), C.byref(ess))
            has = True
        else:
            has = y is not None
            self.L.orc_step_filter(C.byref(self.desc), N, resample_kind, order, t_prev, t, 1 if has else 0,
                                   0.0 if y is None else y, _d(x), _d(z), _d(u), _d(xp), _d(lw), _d(w1), _i(anc), _d(xo),
                                   C.byref(ll), C.byref(ess))
        self._ll, self._ess = ll.value, ess.value
        return dict(x_prop=xp, logw=lw if has else None, w1=w1 if has else None, anc=anc if has else None, x_out=xo,
                    ll=ll.value, ess=ess.value)

    def reset(self, N):
        self._ll, self._ess = 0.0, N

    def filter_ll(self, N, resample_kind, t, y, has_obs=None, seed=1, variant=1, steps=False):
        t, y = f64(t), f64(y)
        h = None if has_obs is None else np.ascontiguousarray(has_obs, dtype=np.uint8)
        lls = np.empty(t.size) if steps else None
        ess = np.empty(t.size, dtype=np.int32) if steps else None
        ll = self.L.orc_filter_ll(C.byref(self.desc), N, resample_kind, t.size, _d(t), _d(y),
                                  None if h is None else h.ctypes.data_as(u8p), seed, variant, _d(lls), _i(ess))
        return (ll, lls, ess) if steps else ll

    def filter_ll_many(self, N, resample_kind, t, y, has_obs=None, seed=1, variant=1, R=1, threads=1):
        t, y = f64(t), f64(y)
        h = None if has_obs is None else np.ascontiguousarray(has_obs, dtype=np.uint8)
        out = np.empty(R)
        self.L.orc_filter_ll_many(C.byref(self.desc), N, resample_kind, t.size, _d(t), _d(y),
                                  None if h is None else h.ctypes.data_as(u8p), seed, variant, R, threads, _d(out))
        return out

    def intervals(self, x, t, interval=0.975):
        """ParticleFilter.getIntervals of the cloud x[d][N]: dict(mean, lower, upper, eta=(mean, lower, upper))."""
        x = f64(x)
        N = x.shape[1]
        mean, lo, up, eta = np.empty(self.d), np.empty(self.d), np.empty(self.d), np.empty(3)
        rc = self.L.orc_intervals(C.byref(self.desc), N, _d(x), t, interval, _d(mean), _d(lo), _d(up), _d(eta))
        if rc:
            raise IndexError("order-statistic index outside the cloud")
        return dict(mean=mean, lower=lo, upper=up, eta=tuple(eta))

    def observation(self, gamma, rng):
        """Model.observation(gamma).draw for every element of gamma (model/Model.scala:145-149,169-178,209-213,
        242-246,267,282-290,316,340-341), with NumPy's samplers."""
        m, k = self.mod, self.mod.obs_kind
        g = f64(gamma)
        if k == _abi.OBS_POISSON:
            return rng.poisson(np.exp(g)).astype(np.float64)
        if k == _abi.OBS_NEGBIN:
            size = np.exp(m.scale)
            mu = np.exp(g)
            prob = mu / (size + mu)
            return rng.poisson(rng.gamma(size, prob / (1 - prob))).astype(np.float64)
        if k == _abi.OBS_NORMAL:
            return rng.normal(g, np.exp(m.scale))
        if k == _abi.OBS_BERNOULLI:
            return (rng.random(g.size) < np.array([self.link(v) for v in g])).astype(np.float64)
        if k == _abi.OBS_STUDENT_T:
            return rng.standard_t(m.df, g.size) * np.exp(m.scale) + g
        if k == _abi.OBS_ZIP:
            p = np.exp(m.scale) / (1 + np.exp(m.scale))
            u, nz = rng.random(g.size), rng.poisson(np.exp(g))
            return np.where(u < p, 0.0, nz.astype(np.float64))
        if k == _abi.OBS_BETA:
            return rng.beta(np.exp(-g), m.scale)
        raise NotImplementedError("observation = ??? (model/Model.scala:364)")

    def forecast(self, x, t_from, t, rng, interval=0.975):
        """ParticleFilter.getForecast + getMeanForecast (model/ParticleFilter.scala:368-412) of the cloud x[d][N]."""
        x = f64(x)
        N = x.shape[1]
        x1 = self.propagate(x, rng.standard_normal(x.shape), t - t_from)
        gamma = self.f(x1, t)
        eta = np.array([self.link(v) for v in gamma])
        obs = self.observation(gamma, rng)
        obs2 = self.observation(gamma, rng)
        idx = int(np.floor(interval * N))

        def order_stat(v):  # getOrderStatistic :455-460
            o = np.sort(v)
            return float(o[N - idx]), float(o[idx])

        xs = np.sort(x1, axis=1)  # getCredibleInterval :498-503
        return dict(x=x1, gamma=gamma, eta=eta, obs=obs, mean=x1.mean(axis=1), lower=xs[:, N - idx - 1], upper=xs[:, idx - 1],
                    eta_summary=(float(eta.mean()),) + order_stat(eta), obs_summary=(float(obs2.mean()),) + order_stat(obs2))

    def link(self, g):
        return self.L.orc_link(C.byref(self.desc), float(g))

    def simulate(self, T, dt=0.1, seed=1):
        t, y, x = np.empty(T), np.empty(T), np.empty((T, self.d))
        self.L.orc_simulate(C.byref(self.desc), T, dt, seed, _d(t), _d(y), _d(x))
        return t, y, x


def lgcp_nsub(dt, precision):
    return lib().orc_lgcp_nsub(dt, precision)


def exp_det(x):
    return lib().orc_exp_det(float(x))


def expf_det(x):
    return lib().orc_expf_det(float(x))


def w1(logw, mx, order=ORDER_DEVICE):
    logw = f64(logw)
    out = np.empty_like(logw)
    lib().orc_w1(logw.size, _d(logw), mx, order, _d(out))
    return out


def total(w, order=ORDER_DEVICE):
    w = f64(w)
    return lib().orc_total(w.size, _d(w), order)


def ll_ess(w1v, mx, order=ORDER_DEVICE):
    w1v = f64(w1v)
    incr, ess = C.c_double(), C.c_int32()
    lib().orc_ll_ess(w1v.size, _d(w1v), mx, order, C.byref(incr), C.byref(ess))
    return incr.value, ess.value


def resample(kind, w, u, order=ORDER_DEVICE, return_clamped=False):
    w, u = f64(w), f64(np.atleast_1d(u))
    anc = np.empty(w.size, dtype=np.int32)
    nc = C.c_int64()
    lib().orc_resample(kind, order, w.size, _d(w), _d(u), _i(anc), C.byref(nc))
    return (anc, nc.value) if return_clamped else anc


def resample_treemap(kind, w, u):
    w, u = f64(w), f64(np.atleast_1d(u))
    anc = np.empty(w.size, dtype=np.int32)
    lib().orc_resample_treemap(kind, w.size, _d(w), _d(u), _i(anc))
    return anc
