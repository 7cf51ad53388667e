"""Worker of the CPU (gloo) test of the N>1 path: launched with torchrun, world_size 2, no GPU.

What the ranks of a sharded filter exchange per observation is restated
here on the host with real inter-process communication: the max log-weight, the EXACT fixed-point
weight totals (python integers built from the oracle's 2^-96 quantisation) and the exclusive scan
over the rank totals.  Each rank then resamples only ITS particles' offspring ranges from the
global CDF; gathered together the ancestors must equal the oracle's single-process answer bit for
bit -- i.e. the result does not depend on how the cloud is partitioned.  Also covers the host
plumbing of composablestatespacemodels_b200.sharding (blob all-gather order, slot ranges)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fix96(lib, w):
    import ctypes as C
    lo, hi = C.c_uint64(), C.c_uint64()
    lib.orc_fix96(float(w), C.byref(lo), C.byref(hi))
    return (hi.value << 64) | lo.value


def dbl128(lib, e):
    return lib.orc_dbl128(e & 0xFFFFFFFFFFFFFFFF, e >> 64, 96)


def main():
    import torch
    import torch.distributed as dist
    from composablestatespacemodels_b200 import _abi, sharding
    import oracle

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = oracle.lib()

    # ---- host plumbing ---------------------------------------------------------------------------
    blob = bytes([rank]) * _abi.SHARD_BLOB_BYTES if hasattr(_abi, "SHARD_BLOB_BYTES") else bytes([rank]) * 256
    blobs = sharding.all_gather_blobs(blob)
    assert [b[0] for b in blobs] == list(range(world)) and all(len(b) == len(blob) for b in blobs)
    N = 6 * 512
    n_loc = sharding.local_count(N, world)
    lo, hi = sharding.slot_range(rank, world, N)
    assert (lo, hi) == (rank * n_loc, (rank + 1) * n_loc)
    for bad in (N + 1, 0):
        try:
            sharding.local_count(bad, world)
            raise AssertionError("uneven split accepted")
        except ValueError:
            pass

    # ---- one resampling step, sharded ---------------------------------------------------------------
    for seed, kind in ((1, _abi.RESAMPLE_SYSTEMATIC), (2, _abi.RESAMPLE_STRATIFIED)):
        rng = np.random.default_rng(seed)                  # the same stream on every rank
        logw_all = -3.0 * rng.standard_exponential(N)      # moderate spread: no vanishing weights
        u = rng.random(1 if kind == _abi.RESAMPLE_SYSTEMATIC else N)
        mine = logw_all[lo:hi]
        # exchange 1: max
        mx = torch.tensor([float(mine.max())], dtype=torch.float64)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        gmax = float(mx.item())
        w1 = oracle.w1(mine, gmax, oracle.ORDER_DEVICE)
        # exchange 2: exact rank totals (integers: addition is associative, so the order is irrelevant)
        fx = [fix96(lib, w) for w in w1]
        totals = [None] * world
        dist.all_gather_object(totals, sum(fx))
        total_int, excl = sum(totals), sum(totals[:rank])
        total = dbl128(lib, total_int)
        # local part of the global CDF and the offspring range of every local particle
        P, run = np.empty(n_loc), excl
        for j, e in enumerate(fx):
            run += e
            P[j] = dbl128(lib, run)
        idx = np.arange(N, dtype=np.float64)
        k = (u[0] + idx) / N if kind == _abi.RESAMPLE_SYSTEMATIC else (idx + u) / N
        keys = k * total                                   # fl(fl(k_i) * total), non-decreasing in i
        first = np.searchsorted(keys, dbl128(lib, excl), side="right") if rank else 0   # outputs owned by earlier ranks
        cnt = np.searchsorted(keys, P, side="right")       # #{i : key_i <= P_j}
        if rank == world - 1:
            cnt[-1] = N                                    # the last particle takes what rounding left over
        anc_part = np.full(N, -1, dtype=np.int64)
        prev = first
        for j in range(n_loc):
            anc_part[prev:cnt[j]] = lo + j
            prev = max(prev, cnt[j])
        parts = [None] * world
        dist.all_gather_object(parts, anc_part)
        if rank == 0:
            anc = np.max(np.stack(parts), axis=0)
            assert (np.stack(parts) >= 0).sum(axis=0).tolist() == [1] * N, "every output has exactly one owner rank"
            w1_all = oracle.w1(logw_all, float(logw_all.max()), oracle.ORDER_DEVICE)
            ref = oracle.resample(kind, w1_all, u, oracle.ORDER_DEVICE)
            np.testing.assert_array_equal(anc, ref)
            assert total == oracle.total(w1_all, oracle.ORDER_DEVICE)
    dist.barrier()
    if rank == 0:
        print("mp_gloo_worker ok", world, "ranks")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
