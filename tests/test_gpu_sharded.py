"""One particle cloud sharded over several ranks must give the SAME bits as
the unsharded filter: ancestors, states, log-weights, ESS and log-likelihood.  Here the ranks are
virtual -- R handles of one process on one GPU, driven in lock-step by the cssm_group_* entry
points -- so the whole exchange / scatter / gather logic of the kernels is exercised on the single
GPU the test tier has; the multi-process form (CUDA IPC over NVLink) is tests/mp_shard_worker.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

import composablestatespacemodels_b200 as cs
from composablestatespacemodels_b200 import _abi
import oracle
from configs import SYS, STRAT, c1, c2, c5

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_pair(mod, n_local, R, T, kind, dtype, seed, ys=None, missing=()):
    """The same injected noise through one filter of R*n_local particles and through R shards."""
    rng = np.random.default_rng(seed)
    N, d = n_local * R, mod.dimension
    orc = oracle.Oracle(mod)
    t, y, _ = orc.simulate(T, 0.1, seed + 3)
    if ys is not None:
        y = np.asarray(ys, dtype=np.float64)
    one = cs.GpuFilterHandle(mod, kind, N, dtype=dtype, seed=7)
    grp = cs.ShardedGroup(mod, kind, n_local, R, dtype=dtype, seed=7)
    z0 = rng.standard_normal((d, N))
    one.init_injected(t[0], z0)
    grp.init_injected(t[0], z0)
    np.testing.assert_array_equal(grp.get_particles(), one.get_particles())
    for s in range(T):
        obs = None if s in missing else float(y[s])
        z = rng.standard_normal((d, N))
        u = rng.random(1 if kind == SYS else N)
        a = one.step_injected(t[s], obs, z, u)
        b = grp.step_injected(t[s], obs, z, u)
        np.testing.assert_array_equal(b["x_prop"], a["x_prop"])
        if obs is not None:
            np.testing.assert_array_equal(b["logw"], a["logw"])
            np.testing.assert_array_equal(b["w1"], a["w1"])
            np.testing.assert_array_equal(b["anc"], a["anc"])
            assert b["ess"] == a["ess"] and b["ll"] == a["ll"]
            # and both are what the oracle says for these weights
            mx = float(np.max(a["logw"]))
            np.testing.assert_array_equal(a["anc"], oracle.resample(kind, oracle.w1(a["logw"], mx, oracle.device_order(dtype)), u))
        np.testing.assert_array_equal(grp.get_particles(), one.get_particles())
    stats = [sh.scan_stats() for sh in grp.shards]  # per rank: (tiles settled by the certified fp64 scan, by the exact path)
    one.close()
    grp.close()
    return stats


@pytest.mark.parametrize("R", [2, 4, 8])
@pytest.mark.parametrize("dtype", [_abi.F32, _abi.F64])
def test_sharded_equals_unsharded_systematic(R, dtype):
    run_pair(c5(), 1024, R, 5, SYS, dtype, seed=R)


def test_sharded_equals_unsharded_stratified_and_ragged():
    run_pair(c2(), 1500, 3, 4, STRAT, _abi.F32, seed=11)              # 3 ranks, rank size not a tile multiple
    run_pair(c2(), 700, 2, 5, SYS, _abi.F64, seed=12, missing=(1, 3))  # unobserved steps: K1-only progress exchange
    run_pair(c1(), 5, 4, 3, SYS, _abi.F64, seed=13)                    # tiny shards


def test_sharded_degenerate_weights_cross_rank_runs(monkeypatch):
    """An observation far in the tail: one particle carries everything and long runs of vanishing
    weights (duplicate TreeMap keys, model/Resampling.scala:55-57) cross tile and rank borders."""
    run_pair(c5(), 1024, 4, 4, SYS, _abi.F64, seed=21, ys=[60.0, -45.0, 80.0, 0.0])
    run_pair(c5(), 1024, 4, 3, STRAT, _abi.F32, seed=22, ys=[70.0, 0.5, -90.0])
    monkeypatch.setenv("CSSM_TILE_ITEMS", "8")  # the 2048-particle tiles large clouds use
    run_pair(c5(), 6000, 2, 3, SYS, _abi.F64, seed=23, ys=[70.0, 0.5, -90.0])
    monkeypatch.delenv("CSSM_TILE_ITEMS")
    # one particle with tens of thousands of offspring: the direct fill of the exact scan path, into the peers' slots too
    run_pair(c5(), 12000, 4, 3, SYS, _abi.F32, seed=24, ys=[70.0, 0.5, -90.0])


def test_sharded_philox_run_is_partition_invariant():
    """Device RNG (Philox keyed by the GLOBAL particle slot) + exact sums: a whole llFilter gives
    the same log-likelihood bits for 1, 2 and 4 ranks."""
    mod = c2()
    orc = oracle.Oracle(mod)
    t, y, _ = orc.simulate(30, 0.1, 5)
    N = 1 << 14
    lls = []
    for R in (1, 2, 4):
        if R == 1:
            h = cs.GpuFilterHandle(mod, SYS, N, dtype=_abi.F32, seed=9)
            lls.append(h.ll_arrays(t, y))
            h.close()
        else:
            g = cs.ShardedGroup(mod, SYS, N // R, R, dtype=_abi.F32, seed=9)
            lls.append(g.ll_arrays(t, y))
            g.close()
    assert lls[0] == lls[1] == lls[2], lls


def test_single_gpu_degenerate_weights_match_oracle():
    """Same tail observations on the plain single-GPU filter, checked against the oracle."""
    mod = c5()
    rng = np.random.default_rng(31)
    N, d = 9000, mod.dimension
    for kind in (SYS, STRAT):
        h = cs.GpuFilterHandle(mod, kind, N, dtype=_abi.F64, seed=1)
        h.init_injected(0.0, rng.standard_normal((d, N)))
        for s, yv in enumerate([75.0, -60.0, 0.0]):
            z, u = rng.standard_normal((d, N)), rng.random(1 if kind == SYS else N)
            g = h.step_injected(0.1 * s, yv, z, u)
            mx = float(np.max(g["logw"]))
            w1 = oracle.w1(g["logw"], mx)
            np.testing.assert_array_equal(g["w1"], w1)
            np.testing.assert_array_equal(g["anc"], oracle.resample(kind, w1, u))
            incr, ess = oracle.ll_ess(w1, mx)
            assert g["ess"] == ess
        h.close()


def test_multiprocess_ipc_two_gpus():
    """Two processes, two GPUs, CUDA IPC + NVLink: same bits as one GPU.  Skipped on a 1-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "mp_shard_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "mp_shard_worker ok" in out.stdout
