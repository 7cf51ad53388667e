"""The inference helpers of model/Streaming.scala that sit on the particle-filter path (the JSON / file / Akka
plumbing of that file is out of scope)."""
import numpy as np

from . import _abi
from .filter import GpuFilterHandle
from .resampling import Resampling


def pilotRun(data, model, resample, particles, repetitions, dtype=_abi.F32, device=0, seed=0, parallelism=4):
    """model/Streaming.scala:19-41: for every particle count, the variance of `repetitions` independent estimates of
    the marginal log-likelihood (breeze.stats.variance: the n - 1 form) -- the rule of thumb for choosing the number of
    particles of a PMMH run.  As in the reference the counts are mapped over four workers (`mapAsyncUnordered(4)`,
    model/Streaming.scala:39): each count is one device-resident filter on its own stream whose series is uploaded once
    and re-run `repetitions` times (every run re-initialises the cloud with fresh Philox streams); pilot-sized clouds
    leave most of the GPU idle, so four of them overlap.  Returns [(n, variance)] in the order of `particles`
    (`parallelism=1` runs them one after the other)."""
    from concurrent.futures import ThreadPoolExecutor
    t, y, ho = GpuFilterHandle._series(data)
    kind = Resampling.kind_of(resample)

    def one(arg):
        i, n = arg
        with GpuFilterHandle(model, kind, int(n), dtype=dtype, device=device, seed=seed, stream_id=i) as h:
            h.load_series(t, y, ho)
            lls = [h.ll_resident() for _ in range(int(repetitions))]
        return int(n), float(np.var(lls, ddof=1))

    jobs = list(enumerate(particles))
    if parallelism <= 1 or len(jobs) <= 1:
        return [one(j) for j in jobs]
    with ThreadPoolExecutor(max_workers=min(int(parallelism), len(jobs))) as pool:
        return list(pool.map(one, jobs))
