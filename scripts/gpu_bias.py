"""Log-likelihood estimate against the number of particles, fp32 vs fp64 (bias check)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import composablestatespacemodels_b200 as cs
from composablestatespacemodels_b200 import _abi
import oracle
from configs import c2, c5, c1, SYS, STRAT
from test_oracle import kalman_loglik

for name, make in (("c2", c2), ("c5", c5), ("c1", c1)):
    mod = make()
    orc = oracle.Oracle(mod)
    t, y, _ = orc.simulate(12, 0.1, 1)
    if name == "c5":
        print("c5 exact", kalman_loglik(mod, t, y))
    for dtype, dn in ((_abi.F32, "f32"), (_abi.F64, "f64")):
        for lg in (14, 16, 18, 20, 22, 24):
            if dtype == _abi.F64 and lg > 22:
                continue
            N = 1 << lg
            row = []
            for rule in (_abi.TIE_REFERENCE, _abi.TIE_FIRST):
                h = cs.GpuFilterHandle(mod, SYS, N, dtype=dtype, seed=4)
                h.set_tie_rule(rule)
                h.load_series(t, y)
                est = np.array([h.ll_resident() for _ in range(4)])
                h.close()
                row.append("%.4f +- %.4f" % (est.mean(), est.std(ddof=1) / 2))
            print(name, dn, "2^%d" % lg, "| reference rule", row[0], "| first-index rule", row[1], flush=True)
