// capi_common.h -- error plumbing shared by the translation units behind include/b200mpc.h
#pragma once
#include "../../include/b200mpc.h"
#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace b200mpc {
int fail(int code, const std::string& msg);          // records the message b200mpc_last_error() returns; returns code
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return b200mpc::fail(B200MPC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
