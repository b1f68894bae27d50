// the NLMPC kernels of one built-in system (see nlmpc_launch.cuh)
#include "nlmpc_launch.cuh"
namespace b200mpc { B200MPC_INSTANTIATE_NL_SYSTEM(SysOscNet<6>) }
