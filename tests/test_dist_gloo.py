"""world_size-2 gloo test of the sharding + all-gather logic (CPU).  The per-rank alignment is
done by the C oracle here -- the test stands it in for the kernel because this box has no GPU;
the product path never does that."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_indices_partition():
    from icp_slam_b200 import dist as d
    for n, world, block in [(0, 2, 4), (1, 2, 4), (37, 2, 5), (1000, 8, 16), (4999, 4, 4096), (10, 3, 1)]:
        seen = np.concatenate([d.shard_indices(n, r, world, block) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(n))
        for r in range(world):
            k_first, k_block, k_stride, b = d.shard_all_pairs(n, r, world, block)
            idx = d.shard_indices(n, r, world, block)
            assert b == len(idx)
            enum = [k_first + (i // k_block) * k_stride + (i % k_block) for i in range(b)]
            assert enum == idx.tolist()


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from icp_slam_b200 import dist as d, synth
    from oracle import c_oracle
    scans, pairs, init, _, _ = synth.make_chain_workload(30, 120, seed=21)
    xy, off = c_oracle.pack(scans)

    class Res:
        pass

    def align(p, i, **kw):
        r = Res()
        r.T, r.error, r.iters = c_oracle.icp_batch(xy, off, p, i, n_threads=1, **kw)
        return r

    T, err, passes = d.icp_batch_sharded(align, pairs, init, block=4, epsilon=0.05, max_iters=100)
    np.savez(os.path.join(tmp, f"r{rank}.npz"), T=T, err=err, passes=passes)
    dist.destroy_process_group()


def test_sharded_gather_world2(tmp_path):
    port = 29500 + os.getpid() % 500
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from icp_slam_b200 import synth
    from oracle import c_oracle
    scans, pairs, init, _, _ = synth.make_chain_workload(30, 120, seed=21)
    xy, off = c_oracle.pack(scans)
    T, err, passes = c_oracle.icp_batch(xy, off, pairs, init, epsilon=0.05, max_iters=100)
    for r in range(2):
        z = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_array_equal(z["T"], T)
        np.testing.assert_array_equal(z["err"], err)
        np.testing.assert_array_equal(z["passes"], passes)
