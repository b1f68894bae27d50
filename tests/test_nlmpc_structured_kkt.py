"""CPU: the bordered block-tridiagonal factorisation planned for the stage-structured NLMPC kernel
(tests/nlmpc_structured_kkt_reference.py) against a dense solve, on the reduced KKT matrices the SQP actually produces for the
reference's example systems (real finite-difference Jacobians at random points, a random SPD block-diagonal B)."""
import numpy as np
import pytest

from nlmpc_sqp_reference import stage_groups
from nlmpc_structured_kkt_reference import BorderedBlockTridiagonal, assemble_blocks, stage_partition
from oracle.nlmpc_formulation import oscnet_formulation, ugv_formulation, vanderpol_formulation


def _H(f, rng):
    z = rng.standard_normal(f.nz) * 0.5
    z[-1] = 0.05
    x0 = rng.uniform(-0.5, 0.5, f.nx)
    _, Je = f.state_eq(z, x0)
    _, Ji = f.ineq_con(z, x0)
    n = f.nz
    A = np.vstack([Je, Ji, np.eye(n)])
    B = np.zeros((n, n))
    for G in stage_groups(f):                                      # the block-diagonal quasi-Newton matrix of the spec
        M = rng.standard_normal((G.size, G.size))
        B[np.ix_(G, G)] = M @ M.T + G.size * np.eye(G.size)
    D = rng.uniform(0.5, 2.0, n); E = rng.uniform(0.5, 2.0, A.shape[0])
    rho = np.where(np.arange(A.shape[0]) < Je.shape[0], 100.0, 0.1)
    As = E[:, None] * A * D[None, :]
    return 0.7 * D[:, None] * B * D[None, :] + 1e-6 * np.eye(n) + As.T @ (rho[:, None] * As)


CASES = [("vanderpol ch<ph", lambda: vanderpol_formulation()), ("ugv ch==ph", lambda: ugv_formulation(10, 10, v_pref=(0.6, 0.8))),
         ("oscnet ch<ph", lambda: oscnet_formulation(4, 15, 8)), ("ugv Tph=30", lambda: ugv_formulation(30, 30, v_pref=(0.6, 0.8)))]


@pytest.mark.parametrize("name,make", CASES)
def test_bordered_block_tridiagonal_solve_matches_dense(name, make):
    f = make()
    rng = np.random.default_rng(len(name))
    H = _H(f, rng)
    groups, border = stage_partition(f)
    assert sorted(np.concatenate(groups + [border]).tolist()) == list(range(f.nz))          # a partition of z
    diag, sub, C, D, outside = assemble_blocks(H, groups, border)
    assert outside == 0.0, f"{name}: H has entries outside the bordered block-tridiagonal pattern ({outside})"
    F = BorderedBlockTridiagonal(diag, sub, C, D)
    r = rng.standard_normal(f.nz)
    xT, xb = F.solve([r[g] for g in groups], r[border])
    x = np.zeros(f.nz)
    for g, v in zip(groups, xT):
        x[g] = v
    x[border] = xb
    ref = np.linalg.solve(H, r)
    assert np.abs(x - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
    assert max(g.size for g in groups) <= f.nx + f.nu and border.size == f.nu + 1


def test_structured_sqp_end_to_end():
    """The whole SQP with block-diagonal BFGS AND the structured factorisation in every ADMM / polish solve: same major
    iterations and optimum as the same algorithm on dense linear algebra, same optimum as SLSQP."""
    from nlmpc_sqp_reference import QPADMM, sqp_solve
    from nlmpc_structured_kkt_reference import structured_factor
    from oracle import nlmpc_slsqp as S
    for f, hard, x0 in ((vanderpol_formulation(), True, np.array([0.0, 1.0])), (ugv_formulation(10, 10, v_pref=(0.6, 0.8)), False, np.zeros(4))):
        lb, ub = S.default_bounds(f, hard)
        if not hard:
            lb[-1] = 0.0
        z0 = S.initial_guess(f, x0, np.zeros(f.nu), lb=lb, ub=ub)
        dense = sqp_solve(f, x0, z0, lb, ub, bfgs_groups=stage_groups(f))
        struct = sqp_solve(f, x0, z0, lb, ub, bfgs_groups=stage_groups(f), qp=QPADMM(spd_factor=structured_factor(f)))
        ref = S.solve(f, x0, z0, lb, ub)
        assert abs(struct["nit"] - dense["nit"]) <= 2 and np.abs(struct["z"] - dense["z"]).max() < 2e-5      # finite-difference noise floor in the flat input directions
        assert ref["success"] and abs(struct["cost"] - ref["cost"]) < 1e-7 * max(1.0, abs(ref["cost"])) and struct["viol"] < 1e-8
