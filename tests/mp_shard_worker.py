"""Worker of the multi-process sharding test: launched with torchrun, one process per GPU.
Checks that a filter sharded over WORLD_SIZE processes (CUDA IPC peer pointers, exchanges inside
the kernels) returns the same log-likelihood bits as the same filter on one GPU, and that the
resampled clouds agree."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import composablestatespacemodels_b200 as cs
    from composablestatespacemodels_b200 import _abi, sharding, simulate
    from configs import c2, c5, SYS, STRAT

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for mod, kind, N, T in ((c2(), SYS, 1 << 16, 40), (c5(), STRAT, 3 * 4096 * world, 12), (c5(), SYS, 1 << 20, 25)):
        t, y, _ = simulate.simRegular(mod, 0.1, T, seed=4)
        h = sharding.create_sharded(mod, kind, N, dtype=_abi.F32, device=local, seed=5)
        ll = h.ll_arrays(t, y)
        x = h.get_particles()            # this rank's slots of the resampled cloud
        dist.barrier()
        ref_ll, ref_x = None, None
        if rank == 0:
            one = cs.GpuFilterHandle(mod, kind, N, dtype=_abi.F32, device=local, seed=5)
            ref_ll = one.ll_arrays(t, y)
            ref_x = one.get_particles()
            one.close()
        lls = [None] * world
        dist.all_gather_object(lls, ll)
        xs = [None] * world
        dist.all_gather_object(xs, x)
        if rank == 0:
            assert all(v == ref_ll for v in lls), (lls, ref_ll)
            np.testing.assert_array_equal(np.concatenate(xs, axis=1), ref_x)
        dist.barrier()
        h.close()
    if rank == 0:
        print("mp_shard_worker ok", world, "ranks")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
