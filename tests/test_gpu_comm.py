"""The exchange step inside the product: b200mpc_comm_* + b200mpc_lmpc_allgather_cmd through the C ABI.
A one-rank communicator exercises the whole plumbing on a single GPU (unique id -> ncclCommInitRank -> ncclAllGather on the
solve stream); the N-rank path is the same call and is measured by bench.py under torchrun."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_allgather_cmd_single_rank():
    import torch
    import libmpc_b200 as L
    from libmpc_b200 import workloads as W
    B, ph = 64, 10
    c = W.build_quadrotor_controller(L, ph, B, 250)
    x0, r = W.quadrotor_inputs(0, B)
    yref = np.zeros((B, 12, ph)); yref[:, 2, :] = r[:, None]
    c.setReferences(yref, np.zeros((4, ph)), np.zeros((4, ph)))
    comm = L.Comm(1, 0, L.Comm.unique_id(), 0)
    out = torch.full((B, 4), float("nan"), dtype=torch.float64, device="cuda")
    c.solve_async(x0, np.zeros((B, 4)))
    c.allgather_cmd(comm, out.data_ptr())          # enqueued behind the solve on the same stream, no host sync in between
    res = c.fetch_result()
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), res.cmd)
    assert np.isfinite(res.cmd).all()
    comm.close()


def test_comm_rejects_bad_arguments():
    import libmpc_b200 as L
    with pytest.raises(ValueError):
        L.Comm(1, 0, b"short", 0)
    with pytest.raises(RuntimeError):
        L.Comm(2, 5, L.Comm.unique_id(), 0)        # rank outside [0, nranks)
