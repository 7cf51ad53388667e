"""A filter step in which one particle takes nearly all the weight, at 2^24 particles: time per llFilter of six observations
with mild and with extreme observations (c5's model; the heavy-particle fill of the scan + search kernels)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import composablestatespacemodels_b200 as cs
from composablestatespacemodels_b200 import _abi
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from configs import c5, SYS

mod = c5()
N = 1 << 24
t = 0.1 * np.arange(6)
for name, y in (("mild", np.array([0.3, 0.5, -0.2, 0.0, 0.4, 0.1])), ("extreme", np.array([0.3, 75.0, -60.0, 0.0, 40.0, 0.1]))):
    for mode, mname in ((_abi.SCAN_AUTO, "auto"), (_abi.SCAN_EXACT, "exact")):
        h = cs.GpuFilterHandle(mod, SYS, N, dtype=_abi.F32, seed=12)
        h.scan_mode(mode)
        h.load_series(t, y)
        h.ll_resident()
        t0 = time.perf_counter()
        ll, lls, ess = h.ll_resident(steps=True)
        dt = time.perf_counter() - t0
        print(name, mname, "ms per llFilter of 6 observations: %.2f" % (dt * 1e3), "ESS", list(ess), "tiles (certified, exact)", h.scan_stats())
        h.close()
