"""The particle filter behind the reference's API (reference: model/ParticleFilter.scala).

`Filter`, `FilterLgcp`, `FilterInit` and the `ParticleFilter` constructors keep the reference's
names and argument meaning; every particle operation runs in libcssm_gpu.so on the GPU.  A
`PfState` does not hold the cloud on the host: `particles` copies it out of device memory on
demand, `ll` and `ess` are plain numbers.
"""
import ctypes as C

import numpy as np

from . import _abi
from .resampling import Resampling


class Data:
    """model/Data.scala: anything with a time and an optional observation."""

    def __init__(self, t, observation):
        self.t = float(t)
        self.observation = None if observation is None else float(observation)


TimedObservation = Data


class StateSpace:
    """model/Sde.scala:170: a state at a time."""

    def __init__(self, time, state):
        self.time, self.state = time, state


class CredibleInterval:
    """model/ParticleFilter.scala:20-22."""

    def __init__(self, lower, upper):
        self.lower, self.upper = lower, upper

    def __repr__(self):
        return f"{self.lower}, {self.upper}"


class PfOut:
    """model/ParticleFilter.scala:53-59."""

    def __init__(self, time, observation, eta, etaIntervals, state, stateIntervals):
        self.time, self.observation, self.eta, self.etaIntervals = time, observation, eta, etaIntervals
        self.state, self.stateIntervals = state, stateIntervals


class ForecastOut:
    """model/ParticleFilter.scala:71-78."""

    def __init__(self, t, obs, obsIntervals, eta, etaIntervals, state, stateIntervals):
        self.t, self.obs, self.obsIntervals, self.eta, self.etaIntervals = t, obs, obsIntervals, eta, etaIntervals
        self.state, self.stateIntervals = state, stateIntervals


class ObservationWithState:
    """model/Data.scala:27-36: one forecast / simulated datum.  In a forecast of N particles every field but `t` is
    an array over the particles (sdeState: [N, d])."""

    def __init__(self, t, observation, eta, gamma, sdeState):
        self.t, self.observation, self.eta, self.gamma, self.sdeState = t, observation, eta, gamma, sdeState


class StaleStateError(RuntimeError):
    """A PfState whose cloud the device handle no longer holds was read (see PfState)."""


class PfState:
    """model/ParticleFilter.scala:32-37.  `t`, `observation`, `ll` and `ess` are plain values; the cloud stays in
    device memory and `particles` copies it out when asked for.  The reference's PfState is an immutable value, the
    handle's cloud is not: every init / step / whole-series call replaces it.  A state therefore remembers the
    generation of the cloud it describes and reading the cloud of an older state (`particles`, getIntervals,
    getForecast, paths) raises StaleStateError instead of silently returning the handle's CURRENT cloud.  Call
    `materialise()` on a state that has to outlive the next step: it copies the cloud to the host once and the state
    then behaves like the reference's value."""

    def __init__(self, t, observation, handle, ll, ess):
        self.t, self.observation, self.ll, self.ess = t, observation, ll, ess
        self._handle = handle
        self._gen = handle.generation
        self._host = None

    def _live(self):
        if self._gen != self._handle.generation:
            raise StaleStateError(
                f"this PfState (t = {self.t}) describes cloud generation {self._gen}, the filter handle has moved on to "
                f"{self._handle.generation}: read a state before the next step, or keep it with materialise()")
        return self._handle

    def materialise(self):
        """Copy the cloud to the host now; afterwards `particles` no longer depends on the handle."""
        if self._host is None:
            self._host = self._live().get_particles().T.copy()
        return self

    @property
    def particles(self):
        """The resampled cloud as an array [N, d] (particle-major, like Vector[State])."""
        if self._host is not None:
            return self._host
        return self._live().get_particles().T.copy()


class PfStateInterpolate:
    """model/ParticleFilter.scala:39-44: the particles are paths.  `particles` reads them from the device when asked
    for: an array [N, len, d] with the NEWEST state first (the reference's List[State] conses new states at the head)."""

    def __init__(self, t, observation, handle, ll, ess, reverse=False, gen=None):
        self.t, self.observation, self.ll, self.ess = t, observation, ll, ess
        self._handle, self._reverse = handle, reverse
        self._gen = handle.generation if gen is None else gen

    def _live(self):
        if self._gen != self._handle.generation:
            raise StaleStateError(f"this PfStateInterpolate (t = {self.t}) is older than the handle's cloud; read it before the next step")
        return self._handle

    def paths(self, indices=None):
        """Paths of the given particles (all by default), [n, len, d], OLDEST state first."""
        return self._live().get_paths(indices)

    @property
    def particles(self):
        p = self._live().get_paths()[:, ::-1, :]
        return p[::-1] if self._reverse else p


class GpuFilterHandle:
    """Owner of one cssm_filter_t (AutoCloseable on the JVM side)."""

    def __init__(self, mod, resample_kind, n, dtype=_abi.F32, device=0, seed=0, stream_id=0, rank=0, world=1):
        """`world` > 1: this handle is rank `rank` of a filter sharded over `world` GPUs and `n` is
        the number of particles of THIS rank (see include/cssm.h, cssm_filter_create_sharded)."""
        self._lib = _abi.lib()
        self._h = C.c_void_p()
        self.mod, self.n, self.d = mod, int(n), mod.dimension
        self.rank, self.world = int(rank), int(world)
        desc, keep = mod.desc()
        if world > 1:
            _abi.check(self._lib.cssm_filter_create_sharded(C.byref(desc), self.n, resample_kind, dtype, device, seed,
                                                            stream_id, rank, world, C.byref(self._h)))
        else:
            _abi.check(self._lib.cssm_filter_create(C.byref(desc), self.n, resample_kind, dtype, device, seed, stream_id,
                                                    C.byref(self._h)))
        self.resample_kind, self.dtype = resample_kind, dtype
        self.generation = 0  # bumped by every call that replaces the cloud (PfState staleness check)

    # ---- sharding ---------------------------------------------------------------------------
    def shard_export(self):
        """This rank's connection blob (bytes) for the other ranks."""
        buf = C.create_string_buffer(_abi.SHARD_BLOB_BYTES)
        _abi.check(self._lib.cssm_filter_shard_export(self._h, buf))
        return buf.raw

    def shard_connect(self, blobs):
        """`blobs`: the blobs of all ranks, in rank order."""
        joined = b"".join(blobs)
        assert len(joined) == _abi.SHARD_BLOB_BYTES * self.world
        _abi.check(self._lib.cssm_filter_shard_connect(self._h, joined, self.world))

    def close(self):
        if self._h:
            self._lib.cssm_filter_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- life cycle -------------------------------------------------------------------------
    def set_params(self, mod):
        desc, keep = mod.desc()
        _abi.check(self._lib.cssm_filter_set_params(self._h, C.byref(desc)))
        self.mod = mod

    def reseed(self, seed, stream_id=0):
        _abi.check(self._lib.cssm_filter_reseed(self._h, seed, stream_id))

    def set_tie_rule(self, rule):
        """_abi.TIE_REFERENCE (default: the reference's TreeMap rule, a repeated cumulative weight selects the last
        particle inserted) or _abi.TIE_FIRST (textbook inverse CDF; see include/cssm.h)."""
        _abi.check(self._lib.cssm_filter_set_tie_rule(self._h, int(rule)))

    def scan_mode(self, mode):
        """_abi.SCAN_AUTO (certified fp64 scan with the exact path as fallback) or _abi.SCAN_EXACT (include/cssm.h)."""
        _abi.check(self._lib.cssm_filter_scan_mode(self._h, int(mode)))

    def scan_stats(self):
        """(tiles settled by the certified fp64 path, tiles handed to the exact path) since the last initialisation"""
        a, b = C.c_int64(), C.c_int64()
        _abi.check(self._lib.cssm_filter_scan_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_stream(self, cuda_stream):
        _abi.check(self._lib.cssm_filter_set_stream(self._h, cuda_stream))

    # ---- stepping ---------------------------------------------------------------------------
    def init(self, t0):
        self.generation += 1
        _abi.check(self._lib.cssm_filter_init(self._h, float(t0)))

    def init_state(self, t0, x0):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self.generation += 1
        _abi.check(self._lib.cssm_filter_init_state(self._h, float(t0), _abi.dptr(x0)))

    def init_injected(self, t0, z0):
        z0 = np.ascontiguousarray(z0, dtype=np.float64)
        assert z0.shape == (self.d, self.n)
        self.generation += 1
        _abi.check(self._lib.cssm_filter_init_injected(self._h, float(t0), _abi.dptr(z0)))

    def step(self, t, observation):
        ll, ess = C.c_double(), C.c_int32()
        has = 0 if observation is None else 1
        self.generation += 1
        _abi.check(self._lib.cssm_filter_step(self._h, float(t), has, 0.0 if observation is None else float(observation),
                                              C.byref(ll), C.byref(ess)))
        return ll.value, ess.value

    def n_substeps(self, dt):
        n = C.c_int64()
        _abi.check(self._lib.cssm_filter_n_substeps(self._h, float(dt), C.byref(n)))
        return n.value

    def step_injected(self, t, observation, z, u, want=("x_prop", "logw", "w1", "anc")):
        """One stepFilter with caller-provided noise; returns a dict of the device's results."""
        has = 0 if observation is None else 1
        z = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
        u = None if u is None else np.ascontiguousarray(np.atleast_1d(u), dtype=np.float64)
        out = {}
        xp = np.empty((self.d, self.n)) if "x_prop" in want else None
        lw = np.empty(self.n) if "logw" in want else None
        w1 = np.empty(self.n) if "w1" in want else None
        anc = np.empty(self.n, dtype=np.int32) if "anc" in want else None
        ll, ess = C.c_double(), C.c_int32()
        self.generation += 1
        _abi.check(self._lib.cssm_filter_step_injected(
            self._h, float(t), has, 0.0 if observation is None else float(observation), _abi.dptr(z), _abi.dptr(u),
            _abi.dptr(xp), _abi.dptr(lw), _abi.dptr(w1),
            None if anc is None else anc.ctypes.data_as(_abi.c_int32_p), C.byref(ll), C.byref(ess)))
        out.update(x_prop=xp, logw=lw, w1=w1, anc=anc, ll=ll.value, ess=ess.value)
        return out

    # ---- whole series -----------------------------------------------------------------------
    @staticmethod
    def _series(data):
        t = np.ascontiguousarray([d.t for d in data], dtype=np.float64)
        y = np.ascontiguousarray([0.0 if d.observation is None else d.observation for d in data], dtype=np.float64)
        h = np.ascontiguousarray([0 if d.observation is None else 1 for d in data], dtype=np.uint8)
        return t, y, h

    def ll_arrays(self, t, y, has_obs=None):
        t = np.ascontiguousarray(t, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        h = None if has_obs is None else np.ascontiguousarray(has_obs, dtype=np.uint8)
        ll = C.c_double()
        self.generation += 1
        _abi.check(self._lib.cssm_filter_ll(self._h, _abi.dptr(t), _abi.dptr(y),
                                            None if h is None else h.ctypes.data_as(_abi.c_uint8_p), t.size, C.byref(ll)))
        return ll.value

    def load_series(self, t, y, has_obs=None):
        t = np.ascontiguousarray(t, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        h = None if has_obs is None else np.ascontiguousarray(has_obs, dtype=np.uint8)
        _abi.check(self._lib.cssm_filter_load_series(self._h, _abi.dptr(t), _abi.dptr(y),
                                                     None if h is None else h.ctypes.data_as(_abi.c_uint8_p), t.size))

    def series_len(self):
        """Length of the series the handle holds, as the C side counts it (load_series, ll_arrays, run_arrays and
        set_params all (re)load one)."""
        n = C.c_int64()
        _abi.check(self._lib.cssm_filter_series_len(self._h, C.byref(n)))
        return n.value

    def ll_resident(self, steps=False):
        ll = C.c_double()
        self.generation += 1
        if not steps:
            _abi.check(self._lib.cssm_filter_ll_resident(self._h, C.byref(ll), None, None))
            return ll.value
        T = self.series_len()  # the C side writes exactly this many entries
        if T <= 0:
            raise _abi.CssmError(-5, "no series loaded")
        lls, ess = np.empty(T), np.empty(T, dtype=np.int32)
        _abi.check(self._lib.cssm_filter_ll_resident(self._h, C.byref(ll), _abi.dptr(lls), ess.ctypes.data_as(_abi.c_int32_p)))
        return ll.value, lls, ess

    def run_arrays(self, t, y, has_obs=None):
        t = np.ascontiguousarray(t, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        h = None if has_obs is None else np.ascontiguousarray(has_obs, dtype=np.uint8)
        ll = C.c_double()
        states = np.empty((t.size + 1, self.d))
        self.generation += 1
        _abi.check(self._lib.cssm_filter_run(self._h, _abi.dptr(t), _abi.dptr(y),
                                             None if h is None else h.ctypes.data_as(_abi.c_uint8_p), t.size, C.byref(ll),
                                             _abi.dptr(states)))
        return ll.value, states

    def last_elapsed_ms(self):
        ms = C.c_float()
        _abi.check(self._lib.cssm_filter_last_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def last_launches(self):
        n = C.c_int64()
        _abi.check(self._lib.cssm_filter_last_launches(self._h, C.byref(n)))
        return n.value

    KERNEL_CLASSES = ("propagate_weight", "weight_sums", "scan_search", "multinomial_search", "init", "series")

    def series_mode(self, mode):
        """How whole-series calls run: _abi.SERIES_AUTO (small clouds in one cooperative launch),
        SERIES_THREE_LAUNCH or SERIES_SINGLE_LAUNCH (include/cssm.h)."""
        _abi.check(self._lib.cssm_filter_series_mode(self._h, int(mode)))

    def profile(self, stride):
        _abi.check(self._lib.cssm_filter_profile(self._h, int(stride)))

    def profile_read(self):
        """{kernel class: (sum of device ms over the sampled launches, sampled launches)}"""
        ms, n = np.zeros(8), np.zeros(8, dtype=np.int64)
        _abi.check(self._lib.cssm_filter_profile_read(self._h, _abi.dptr(ms), n.ctypes.data_as(_abi.c_int64_p)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KERNEL_CLASSES)}

    # ---- reading the cloud ------------------------------------------------------------------
    def get_particles(self):
        x = np.empty((self.d, self.n))
        _abi.check(self._lib.cssm_filter_get_particles(self._h, _abi.dptr(x)))
        return x

    def sample_one(self):
        x = np.empty(self.d)
        _abi.check(self._lib.cssm_filter_sample_one(self._h, _abi.dptr(x)))
        return x

    def get_ll(self):
        ll, ess = C.c_double(), C.c_int32()
        _abi.check(self._lib.cssm_filter_get_ll(self._h, C.byref(ll), C.byref(ess)))
        return ll.value, ess.value

    def mean_state(self):
        m = np.empty(self.d)
        _abi.check(self._lib.cssm_filter_mean_state(self._h, _abi.dptr(m)))
        return m

    def intervals(self, t, interval=0.975):
        """Mean, per-coordinate credible intervals and the two order statistics of gamma = f(x, t) of
        the current cloud, computed on the device (cssm_filter_intervals)."""
        mean, lo, up, g = np.empty(self.d), np.empty(self.d), np.empty(self.d), np.empty(2)
        _abi.check(self._lib.cssm_filter_intervals(self._h, float(t), float(interval), _abi.dptr(mean), _abi.dptr(lo),
                                                   _abi.dptr(up), _abi.dptr(g)))
        return dict(mean=mean, lower=lo, upper=up, gamma=(float(g[0]), float(g[1])))


    # ---- paths (FilterInterpolate) ------------------------------------------------------------
    def paths_enable(self, max_steps):
        _abi.check(self._lib.cssm_filter_paths_enable(self._h, int(max_steps)))

    def paths_len(self):
        n = C.c_int64()
        _abi.check(self._lib.cssm_filter_paths_len(self._h, C.byref(n)))
        return n.value

    def get_paths(self, indices=None):
        """[n, len + 1, d], oldest state first, of the particles `indices` (all by default) of the current cloud."""
        ln = self.paths_len()
        if ln < 0:
            raise _abi.CssmError(-5, "no paths recorded")
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
        n = self.n if idx is None else idx.size
        out = np.empty((n, ln + 1, self.d))
        _abi.check(self._lib.cssm_filter_get_paths(self._h, None if idx is None else idx.ctypes.data_as(_abi.c_int32_p), n,
                                                   _abi.dptr(out)))
        return out

    def forecast(self, t, interval=0.975, chain=False, summarise=True):
        """getForecast / getMeanForecast (model/ParticleFilter.scala:368-412) on the device: every particle advanced
        to `t` (the filter itself is not touched), eta and two observation draws per particle; with `summarise` the
        means and credible intervals of cssm_filter_forecast.  `chain`: continue from the previous forecast cloud
        (SimulateData.forecast, model/Data.scala:202-217)."""
        if not summarise:
            _abi.check(self._lib.cssm_filter_forecast(self._h, float(t), float(interval), int(bool(chain)), None, None, None,
                                                      None, None))
            return None
        mean, lo, up, eta, obs = np.empty(self.d), np.empty(self.d), np.empty(self.d), np.empty(3), np.empty(3)
        _abi.check(self._lib.cssm_filter_forecast(self._h, float(t), float(interval), int(bool(chain)), _abi.dptr(mean),
                                                  _abi.dptr(lo), _abi.dptr(up), _abi.dptr(eta), _abi.dptr(obs)))
        return dict(mean=mean, lower=lo, upper=up, eta=tuple(eta), obs=tuple(obs))

    def forecast_cloud(self):
        """The last forecast cloud: x [d, N], gamma, eta, obs, obs2 [N]."""
        x, g, e, o, o2 = np.empty((self.d, self.n)), np.empty(self.n), np.empty(self.n), np.empty(self.n), np.empty(self.n)
        _abi.check(self._lib.cssm_filter_forecast_cloud(self._h, _abi.dptr(x), _abi.dptr(g), _abi.dptr(e), _abi.dptr(o),
                                                        _abi.dptr(o2)))
        return dict(x=x, gamma=g, eta=e, obs=o, obs2=o2)


class ShardedGroup:
    """R shards of ONE filter inside one process, driven in lock-step by the cssm_group_* entry
    points: virtual ranks on one GPU (devices all equal) or one process over several GPUs.  This
    is how the sharded path is tested without a multi-process launch; the multi-process form is
    one GpuFilterHandle(rank=r, world=R) per process plus shard_export / shard_connect.

    Virtual ranks (several shards on ONE device) are a test vehicle: the kernels of a sharded filter wait inside the
    kernel for flags their peers write, and CUDA does not co-schedule kernels of different streams.  The group
    drivers therefore put all shards of a device on ONE stream and launch every phase in rank order, so that each
    wait is already satisfied when its kernel starts (cssm_api.cu, check_group); never drive same-device shards from
    separate streams or threads.  One GPU per rank -- the production layout -- has no such restriction."""

    def __init__(self, mod, resample_kind, n_local, world, dtype=_abi.F32, devices=None, seed=0, stream_id=0):
        devices = [0] * world if devices is None else list(devices)
        self.world, self.n_local, self.n, self.d = world, int(n_local), int(n_local) * world, mod.dimension
        self.shards = [GpuFilterHandle(mod, resample_kind, n_local, dtype, devices[r], seed, stream_id, rank=r, world=world)
                       for r in range(world)]
        if world > 1:
            blobs = [s.shard_export() for s in self.shards]
            for s in self.shards:
                s.shard_connect(blobs)
        self._lib = _abi.lib()
        self._arr = (C.c_void_p * world)(*[s._h for s in self.shards])

    def close(self):
        for s in self.shards:
            s.close()

    def init(self, t0):
        self.generation += 1
        _abi.check(self._lib.cssm_group_init(self._arr, self.world, float(t0)))

    def init_injected(self, t0, z0):
        z0 = np.ascontiguousarray(z0, dtype=np.float64)
        assert z0.shape == (self.d, self.n)
        _abi.check(self._lib.cssm_group_init_injected(self._arr, self.world, float(t0), _abi.dptr(z0)))

    def step_injected(self, t, observation, z, u):
        has = 0 if observation is None else 1
        z = np.ascontiguousarray(z, dtype=np.float64)
        u = None if u is None else np.ascontiguousarray(np.atleast_1d(u), dtype=np.float64)
        xp, lw, w1 = np.empty((self.d, self.n)), np.empty(self.n), np.empty(self.n)
        anc = np.empty(self.n, dtype=np.int32)
        ll, ess = C.c_double(), C.c_int32()
        _abi.check(self._lib.cssm_group_step_injected(
            self._arr, self.world, float(t), has, 0.0 if observation is None else float(observation), _abi.dptr(z),
            _abi.dptr(u), _abi.dptr(xp), _abi.dptr(lw), _abi.dptr(w1), anc.ctypes.data_as(_abi.c_int32_p), C.byref(ll),
            C.byref(ess)))
        return dict(x_prop=xp, logw=lw, w1=w1, anc=anc, ll=ll.value, ess=ess.value)

    def get_particles(self):
        x = np.empty((self.d, self.n))
        _abi.check(self._lib.cssm_group_get_particles(self._arr, self.world, _abi.dptr(x)))
        return x

    def ll_arrays(self, t, y, has_obs=None):
        t = np.ascontiguousarray(t, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        h = None if has_obs is None else np.ascontiguousarray(has_obs, dtype=np.uint8)
        ll, ms = C.c_double(), C.c_float()
        _abi.check(self._lib.cssm_group_ll(self._arr, self.world, _abi.dptr(t), _abi.dptr(y),
                                           None if h is None else h.ctypes.data_as(_abi.c_uint8_p), t.size, C.byref(ll),
                                           C.byref(ms)))
        self.last_ms = ms.value
        return ll.value


class _ParticleFilterBase:
    """trait ParticleFilter[S], model/ParticleFilter.scala:96-167, on the GPU."""

    def __init__(self, mod, resample, dtype=_abi.F32, device=0, seed=0, stream_id=0):
        self.mod = mod
        self.resample = resample
        self.resample_kind = Resampling.kind_of(resample)
        self.dtype, self.device, self.seed, self.stream_id = dtype, device, seed, stream_id
        self._handle = None

    def _get(self, n):
        if self._handle is None or self._handle.n != n:
            if self._handle is not None:
                self._handle.close()
            self._handle = GpuFilterHandle(self.mod, self.resample_kind, n, self.dtype, self.device, self.seed, self.stream_id)
        return self._handle

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None

    def initialiseState(self, particles, t0):
        """model/ParticleFilter.scala:105-108"""
        h = self._get(particles)
        h.init(t0)
        return PfState(t0, None, h, 0.0, particles)

    def stepFilter(self, s, y):
        """model/ParticleFilter.scala:116-132 (FilterLgcp :210-226).  The cloud lives in the
        handle, so states must be stepped in order (as foldLeft / scan do)."""
        h = s._live()  # the cloud lives in the handle: only its newest state can be stepped
        ll, ess = h.step(y.t, y.observation)
        return PfState(y.t, y.observation, h, ll, ess)

    def llFilter(self, data, n):
        """model/ParticleFilter.scala:137-140"""
        h = self._get(n)
        t, y, ho = GpuFilterHandle._series(data)
        return h.ll_arrays(t, y, ho)

    def filter(self, data, particles):
        """model/ParticleFilter.scala:152-158: (ll, one sampled particle per time, T+1 of them)"""
        h = self._get(particles)
        t, y, ho = GpuFilterHandle._series(data)
        ll, states = h.run_arrays(t, y, ho)
        times = [float(t.min())] + [float(v) for v in t]
        return ll, [StateSpace(tt, st) for tt, st in zip(times, states)]

    def filterStream(self, t0, particles):
        """model/ParticleFilter.scala:163-166: Flow[Data].scan(init)(stepFilter) as a generator
        transformer: emits the initial state, then one PfState per datum."""
        def flow(source):
            s = self.initialiseState(particles, t0)
            yield s
            for y in source:
                s = self.stepFilter(s, y)
                yield s
        return flow


class Filter(_ParticleFilterBase):
    """model/ParticleFilter.scala:233-246"""


class FilterLgcp(_ParticleFilterBase):
    """model/ParticleFilter.scala:169-227: log-Gaussian Cox process, sub-step 10^-precision."""

    def __init__(self, mod, resample, precision, **kw):
        if mod.obs_kind != _abi.OBS_LGCP:
            raise Exception("FilterLgcp needs a model built with Model.lgcp")
        from .model import Model
        super().__init__(Model(mod.leaves, mod.step_mode, precision), resample, **kw)
        self.precision = precision


class FilterInit(_ParticleFilterBase):
    """model/ParticleFilter.scala:252-271: every particle starts at a given state."""

    def __init__(self, mod, resample, initState, **kw):
        super().__init__(mod, resample, **kw)
        self.initState = np.asarray(initState, dtype=np.float64)

    def initialiseState(self, particles, t0):
        h = self._get(particles)
        h.init_state(t0, self.initState)
        return PfState(t0, None, h, 0.0, particles)


class FilterInterpolate:
    """model/ParticleFilter.scala:273-311: the particle filter whose particles are whole paths -- an unobserved datum
    extends every path, an observed one resamples the paths.  The device keeps the propagated cloud of every step and
    the ancestors of every resampling (cssm_filter_paths_enable); a path is materialised by a walk through the
    ancestor tree only when `particles` / `paths` is read.  `max_steps` bounds the stored history."""

    def __init__(self, mod, resample, max_steps=1024, dtype=_abi.F32, device=0, seed=0, stream_id=0):
        self.mod, self.resample, self.max_steps = mod, resample, int(max_steps)
        self.resample_kind = Resampling.kind_of(resample)
        self.dtype, self.device, self.seed, self.stream_id = dtype, device, seed, stream_id
        self._handle = None

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None

    def initialise(self, particles, t0):
        """x0 = mod.sde.initialState.sample(particles) map (_ :: Nil); ll = 0, ess = 0 (sic, :302-303)"""
        if self._handle is None or self._handle.n != particles:
            self.close()
            self._handle = GpuFilterHandle(self.mod, self.resample_kind, particles, self.dtype, self.device, self.seed,
                                           self.stream_id)
            self._handle.paths_enable(self.max_steps)
        self._handle.init(t0)
        return PfStateInterpolate(t0, None, self._handle, 0.0, 0)

    def stepInterpolate(self, s, y):
        """model/ParticleFilter.scala:281-298"""
        h = s._live()
        ll, ess = h.step(y.t, y.observation)
        if y.observation is None:
            ll, ess = s.ll, s.ess
        return PfStateInterpolate(y.t, y.observation, h, ll, ess)

    def filterInterpolate(self, t0, particles):
        """model/ParticleFilter.scala:300-310: Flow[Data].scan(init)(stepInterpolate).map(s => s.copy(particles =
        s.particles.reverse)) as a generator transformer -- the emitted states list their particles in reverse order,
        as the reference's do."""
        def flow(source):
            s = self.initialise(particles, t0)
            yield PfStateInterpolate(s.t, s.observation, s._handle, s.ll, s.ess, reverse=True)
            for y in source:
                s = self.stepInterpolate(s, y)
                yield PfStateInterpolate(s.t, s.observation, s._handle, s.ll, s.ess, reverse=True)
        return flow


class ParticleFilter:
    """object ParticleFilter, model/ParticleFilter.scala:313-361: Readers from Model."""

    @staticmethod
    def filter(resample, t0, n, **kw):
        return lambda mod: Filter(mod, resample, **kw).filterStream(t0, n)

    @staticmethod
    def filterInit(resample, t0, n, initState, **kw):
        return lambda mod: FilterInit(mod, resample, initState, **kw).filterStream(t0, n)

    @staticmethod
    def filterLlState(data, resample, n, **kw):
        return lambda mod: Filter(mod, resample, **kw).filter(data, n)

    @staticmethod
    def likelihood(data, resample, n, **kw):
        return lambda mod: Filter(mod, resample, **kw).llFilter(data, n)

    @staticmethod
    def getIntervals(model, s, interval=0.975):
        """model/ParticleFilter.scala:415-424: PfState -> PfOut (mean state, state intervals, eta = link(f(mean)),
        eta intervals).  The cloud stays on the device: mean and order statistics come from
        cssm_filter_intervals; the (monotone) link is applied here, in fp64, to the two order statistics of gamma."""
        r = s._live().intervals(s.t, interval)
        lo, up = model.link(r["gamma"][0]), model.link(r["gamma"][1])
        if lo > up:  # decreasing link (Beta: exp(-x)): ascending eta is descending gamma
            lo, up = model.link(r["gamma"][1]), model.link(r["gamma"][0])
        eta = model.link(model.f(r["mean"], s.t))
        return PfOut(s.t, s.observation, eta, CredibleInterval(lo, up), r["mean"],
                     [CredibleInterval(a, b) for a, b in zip(r["lower"], r["upper"])])

    @staticmethod
    def getForecast(s, mod, t):
        """model/ParticleFilter.scala:368-383: the particles of `s` advanced to `t` with eta and a drawn observation
        each, as ONE ObservationWithState whose fields are arrays over the particles (the reference returns a
        Vector of N of them)."""
        s._live().forecast(t, summarise=False)
        c = s._handle.forecast_cloud()
        return ObservationWithState(t, c["obs"], c["eta"], c["gamma"], c["x"].T.copy())

    @staticmethod
    def getMeanForecast(s, mod, t, interval):
        """model/ParticleFilter.scala:394-412 -> ForecastOut; nothing but the 3(d + 2) summary numbers leaves the device."""
        r = s._live().forecast(t, interval)
        return ForecastOut(t, r["obs"][0], CredibleInterval(r["obs"][1], r["obs"][2]), r["eta"][0],
                           CredibleInterval(r["eta"][1], r["eta"][2]), r["mean"],
                           [CredibleInterval(a, b) for a, b in zip(r["lower"], r["upper"])])

    @staticmethod
    def effectiveSampleSize(weights):
        """model/ParticleFilter.scala:431-434 (host helper for small vectors)."""
        w = np.asarray(weights, dtype=np.float64)
        wn = w / w.sum()
        return int(np.floor(1.0 / np.sum(wn * wn)))
