"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard a batch, each 'solves' its shard (here: the C
oracle stands in for the device so that the test runs without a GPU -- it is the checker, not the product) and the
all-gather of the command block reproduces the single-process result in instance order."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libmpc_b200.sharding import gather_commands, shard_bounds
    from oracle import c_oracle
    from oracle.lmpc_formulation import quadrotor_formulation
    f = quadrotor_formulation(5)
    rng = np.random.default_rng(11)
    x0 = rng.uniform(-0.2, 0.2, (total, 12))
    lo, hi = shard_bounds(total, world, rank)
    if hi > lo:
        out = c_oracle.solve_batch(f, x0[lo:hi], np.zeros((hi - lo, 4)), c_oracle.default_params(max_iter=250), want_xy=False)
        local = torch.from_numpy(out["cmd"])
    else:
        local = torch.zeros((0, 4), dtype=torch.float64)
    allcmd = gather_commands(local, total)
    if rank == 0:
        q.put(allcmd.numpy())
    dist.destroy_process_group()


def test_shard_bounds_cover_batch():
    from libmpc_b200.sharding import shard_bounds
    for total, world in ((4096, 8), (10, 4), (3, 8), (65536, 8), (7, 2)):
        seen = []
        for r in range(world):
            lo, hi = shard_bounds(total, world, r)
            assert 0 <= lo <= hi <= total
            seen += list(range(lo, hi))
        assert seen == list(range(total))


@pytest.mark.parametrize("total", [6, 7])
def test_gloo_world2_allgather_matches_single_process(total):
    import torch.multiprocessing as mp
    import __graft_entry__ as g
    g.build()
    from oracle import c_oracle
    from oracle.lmpc_formulation import quadrotor_formulation
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f = quadrotor_formulation(5)
    rng = np.random.default_rng(11)
    x0 = rng.uniform(-0.2, 0.2, (total, 12))
    ref = c_oracle.solve_batch(f, x0, np.zeros((total, 4)), c_oracle.default_params(max_iter=250), want_xy=False)["cmd"]
    assert got.shape == (total, 4) and np.array_equal(got, ref)
