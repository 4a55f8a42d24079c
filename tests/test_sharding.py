"""Multi-GPU bookkeeping on the CPU: contiguous shards, and a world_size-2 gloo run that evaluates the
cycle per shard (with the oracle standing in for the device, since this container has no GPU) and checks that
gathering the shards reproduces the unsharded result -- the path has no data-path collective, the gather is
only the test's way to look at every shard."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_batch():
    from sai_primitives_b200.sharding import shard_of, shard_range
    for n in (1, 7, 8, 65536, 65537, 1000003):
        for w in (1, 2, 3, 4, 8):
            hi_prev = 0
            sizes = []
            for r in range(w):
                lo, hi = shard_range(n, r, w)
                assert lo == hi_prev and hi >= lo
                hi_prev = hi
                sizes.append(hi - lo)
            assert hi_prev == n and max(sizes) - min(sizes) <= 1
            for i in (0, n // 3, n - 1):
                lo, hi = shard_range(n, shard_of(i, n, w), w)
                assert lo <= i < hi
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from sai_primitives_b200.sharding import shard_range
    from tests.osc_testlib import TASK_POINTS, OracleBatch, sample_states
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    # every rank regenerates ITS slice from the counter-based generator (stream = robot index)
    q, dq, _ = sample_states("panda", hi - lo, first_index=lo)
    link, pt = TASK_POINTS["panda"]
    ob = OracleBatch("panda", hi - lo); ob.set_state(q, dq)
    ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt(); ob.finalize()
    for i, t in enumerate(ojt):
        t.setGoalPosition(q[i] + 0.1)
    tau = torch.from_numpy(ob.cycle())
    bufs = [None] * world          # shards may differ in size by one robot
    dist.all_gather_object(bufs, tau)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), torch.cat([torch.as_tensor(b) for b in bufs]).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_unsharded(tmp_path):
    import torch.multiprocessing as mp
    from tests.osc_testlib import TASK_POINTS, OracleBatch, sample_states
    n_total, world = 13, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    gathered = np.load(tmp_path / "gathered.npy")
    q, dq, _ = sample_states("panda", n_total)
    link, pt = TASK_POINTS["panda"]
    ob = OracleBatch("panda", n_total); ob.set_state(q, dq)
    ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt(); ob.finalize()
    for i, t in enumerate(ojt):
        t.setGoalPosition(q[i] + 0.1)
    assert np.array_equal(gathered, ob.cycle())     # same arithmetic, same inputs: bit-identical


def test_model_groups_route_rows():
    from sai_primitives_b200.sharding import ModelGroups

    class Fake:
        def __init__(self, name, n):
            self.name, self.n = name, n
        def set_state(self, q, dq):
            self.q = q
        def cycle(self):
            return self.q * 2.0
    models = ["panda", "rrrr", "panda", "puma_like", "rrrr"]
    dof = {"panda": 7, "rrrr": 4, "puma_like": 6}
    g = ModelGroups(models, lambda nm, n: Fake(nm, n))
    q = [np.full(dof[m], float(i)) for i, m in enumerate(models)]
    g.set_state(q, q)
    out = g.cycle()
    for i, m in enumerate(models):
        assert out[i].shape == (dof[m],) and (out[i] == 2.0 * i).all()
