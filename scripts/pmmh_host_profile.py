"""Where does a PMMH iteration spend its time outside the series kernel?  (c4; per-phase wall clock of GpuBootstrapFilter._eval)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from composablestatespacemodels_b200 import MetropolisHastings, GpuBootstrapFilter, Data, perturb
from composablestatespacemodels_b200.resampling import Resampling

_, wl_model, N, T, resampler = bench.WORKLOADS["c4"]
um, p0 = bench.build_unparam(wl_model)
mod = um(p0)
t, y = bench.synth_series(mod, wl_model, T)
data = [Data(tt, yy) for tt, yy in zip(t, y)]
rng = np.random.default_rng(1)
pf = GpuBootstrapFilter(um, p0, data, Resampling.systematicResampling, N, seed=5)
prop = perturb(0.05, rng)
h = pf.handle
acc = dict(propose=0.0, model=0.0, set_params=0.0, ll=0.0, sample=0.0)
p = p0
K = 60
for i in range(K + 5):
    if i == 5:
        acc = {k: 0.0 for k in acc}
        w0 = time.perf_counter()
    a = time.perf_counter(); q = prop(p); b = time.perf_counter(); acc["propose"] += b - a
    m = pf._model(q); c = time.perf_counter(); acc["model"] += c - b
    h.set_params(m); d_ = time.perf_counter(); acc["set_params"] += d_ - c
    ll = h.ll_resident(); e = time.perf_counter(); acc["ll"] += e - d_
    s = h.sample_one(); f = time.perf_counter(); acc["sample"] += f - e
tot = time.perf_counter() - w0
print("per iteration, us:", {k: round(1e6 * v / K, 1) for k, v in acc.items()}, "total", round(1e6 * tot / K, 1), "device ms", h.last_elapsed_ms())
pf.close()
