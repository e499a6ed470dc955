"""The N > 1 path on CPU: world_size-2 gloo processes shard a batch by system index, each
"integrates" its slice (the oracle stands in for the GPU here -- test infrastructure), and the
final gather on rank 0 must equal the single-process result, in order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nsys, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dvode_b200 import batches, sharding
    import oracle_lib as O
    first, last = sharding.shard_range(nsys, rank, world)
    par = batches.robertson_batch(last - first, first=first)
    tol = batches.ROBERTSON_TOL_REF
    touts = batches.robertson_touts()[:3]
    r = O.batch_solve("robertson", par, batches.ROBERTSON_Y0, 0.0, touts, tol["rtol"], tol["atol"],
                      21, itol=2, nthreads=1)
    y = sharding.gather_to_rank0(torch.from_numpy(r["y"][:, -1].copy()), dist, rank, world, nsys)
    st = sharding.gather_to_rank0(torch.from_numpy(r["istats"][:, -1].copy()), dist, rank, world, nsys)
    if rank == 0:
        np.savez(out_path, y=y.numpy(), st=st.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_batch():
    from dvode_b200 import sharding
    for n in (1, 7, 64, 1000003):
        for w in (1, 2, 3, 8):
            edges = [sharding.shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_gloo_gather_matches_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from dvode_b200 import batches
    O.build()
    nsys, world, port = 101, 2, 29533  # odd size: ragged shards
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, port, nsys, out), nprocs=world, join=True)
    got = np.load(out)
    tol = batches.ROBERTSON_TOL_REF
    ref = O.batch_solve("robertson", batches.robertson_batch(nsys), batches.ROBERTSON_Y0, 0.0,
                        batches.robertson_touts()[:3], tol["rtol"], tol["atol"], 21, itol=2)
    assert np.array_equal(got["y"], ref["y"][:, -1])
    assert np.array_equal(got["st"], ref["istats"][:, -1])
