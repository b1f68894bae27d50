"""CPU: user-defined NLMPC systems (the CUDA-source counterpart of NLMPC::setStateSpaceFunction / setObjectiveFunction /
setIneqConFunction / setEqConFunction) compile with NVRTC for sm_100a into the engine's kernels; no GPU is needed for
the compile step.  The run-time behaviour is covered by tests/test_gpu_user_systems.py."""
import pytest

from user_systems import BROKEN_SRC, OUTPUT_MAP_SRC, PENDULUM_SRC, UNICYCLE_SRC, VANDERPOL_SRC


def test_user_system_compiles_eval_kernel():
    import libmpc_b200 as L
    assert L.compile_check(VANDERPOL_SRC, "UserVanDerPol", 0) > 10_000


def test_user_system_with_equality_constraints_compiles_solve_kernel():
    import libmpc_b200 as L
    assert L.compile_check(PENDULUM_SRC, "UserPendulum", 0) > 10_000
    assert L.compile_check(PENDULUM_SRC, "UserPendulum", 1) > 50_000       # warp-per-controller solve variant


def test_user_system_with_output_map_compiles():
    """setOutputFunction: cost / ineq read y = out(x, u) through nl_y (the map is re-applied inside every perturbation)."""
    import libmpc_b200 as L
    assert L.compile_check(OUTPUT_MAP_SRC, "UserWithOutput", 0) > 10_000


def test_baseline_unicycle_shape_compiles_for_the_kernel_it_would_run():
    """BASELINE.json configs[2] at its stated shape (nx3 nu2 Tph30, 62 obstacle inequalities): nz = 151 does not fit shared
    memory with its matrices, so the launch policy picks the packed-factor variant (kernel index 3)."""
    import libmpc_b200 as L
    assert L.compile_check(UNICYCLE_SRC, "UserUnicycle", 3) > 50_000


def test_compile_error_is_reported_with_the_nvrtc_log():
    import libmpc_b200 as L
    with pytest.raises(RuntimeError, match="NVRTC could not compile"):
        L.compile_check(BROKEN_SRC, "Broken", 0)


def test_user_system_compiles_plant_kernel():
    """The plant-step / RK4 kernel of the device closed loop (b200mpc_nlmpc_closed_loop, b200mpc_nlmpc_rk4) for a user system."""
    import libmpc_b200 as L
    assert L.compile_check(VANDERPOL_SRC, "UserVanDerPol", 5) > 2_000
    assert L.compile_check(UNICYCLE_SRC, "UserUnicycle", 5) > 2_000
