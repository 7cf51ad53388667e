"""The inference helpers of model/Streaming.scala that sit on the particle-filter path (the JSON / file / Akka
plumbing of that file is out of scope)."""
import numpy as np

from . import _abi
from .filter import GpuFilterHandle
from .resampling import Resampling


def pilotRun(data, model, resample, particles, repetitions, dtype=_abi.F32, device=0, seed=0):
    """model/Streaming.scala:19-41: for every particle count, the variance of `repetitions` independent estimates of
    the marginal log-likelihood (breeze.stats.variance: the n - 1 form) -- the rule of thumb for choosing the number of
    particles of a PMMH run.  The reference maps the counts over four threads (`mapAsyncUnordered(4)`); here each count
    is one device-resident filter whose series is uploaded once and re-run `repetitions` times (every run re-initialises
    the cloud with fresh Philox streams).  Returns [(n, variance)] in the order of `particles`."""
    t, y, ho = GpuFilterHandle._series(data)
    kind = Resampling.kind_of(resample)
    out = []
    for i, n in enumerate(particles):
        with GpuFilterHandle(model, kind, int(n), dtype=dtype, device=device, seed=seed, stream_id=i) as h:
            h.load_series(t, y, ho)
            lls = [h.ll_resident() for _ in range(int(repetitions))]
        out.append((int(n), float(np.var(lls, ddof=1))))
    return out
