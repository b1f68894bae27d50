// nlmpc_launch_choice.h -- dense vs stage-structured NLMPC solver selection, shared by the built-in and the NVRTC launch paths
#pragma once
#include <cstdlib>
#include <cstring>

namespace b200mpc {
// Which solver a (system, horizon) gets: the stage-structured kernel (nlmpc_structured.cuh) whenever the system declares its
// inequality sparsity and has no user equality constraints and one controller fits shared memory; the dense kernel otherwise.
// B200MPC_NLMPC_SOLVER=dense|structured overrides (structured fails loudly when it is not applicable); B200MPC_NLS_THREADS=32|64|128.
int nl_solver_override();          // b200mpc_nlmpc_set_solver: 0 automatic, 1 dense, 2 structured (b200mpc_nlmpc.cu)
inline int nl_structured_choice(bool supported, size_t smem, int maxsm, int n, int* nt) {
    const char* e = getenv("B200MPC_NLMPC_SOLVER");
    const int ov = nl_solver_override();
    if (ov == 1) e = "dense";
    if (ov == 2) e = "structured";
    const char* t = getenv("B200MPC_NLS_THREADS");
    const int auto_nt = n >= 60 ? 128 : 64;       // measured: vanderpol (n = 26) 85k / 75k solves/s with 64 / 128 threads, ugv Tph=30 (n = 181) 1.6k / 2.3k
    *nt = t ? atoi(t) : auto_nt;
    if (*nt != 32 && *nt != 64 && *nt != 128) *nt = auto_nt;
    const bool fits = supported && smem <= (size_t)maxsm;
    if (e && !strcmp(e, "dense")) return 0;
    if (e && !strcmp(e, "structured")) return fits ? 1 : -1;
    return fits ? 1 : 0;
}

}  // namespace b200mpc
