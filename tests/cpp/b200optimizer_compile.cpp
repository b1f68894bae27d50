// Type-checks include/mpc_b200/B200Optimizer.hpp against the REFERENCE's headers (mpc/IOptimizer.hpp and what it includes) with
// the quadrotor sizes of examples/quadrotor_ex.cpp and with dynamic sizes; see tests/test_b200optimizer_header.py.
#include <mpc_b200/B200Optimizer.hpp>

constexpr mpc::MPCSize kQuad(12, 4, 4, 12, 10, 10, 0, 0);
constexpr mpc::MPCSize kDyn(Eigen::Dynamic, Eigen::Dynamic, Eigen::Dynamic, Eigen::Dynamic, Eigen::Dynamic, Eigen::Dynamic, 0, 0);
template class mpc::B200Optimizer<kQuad>;
template class mpc::B200Optimizer<kDyn>;

int main() {
    mpc::B200Optimizer<kQuad> opt;
    mpc::IOptimizer<kQuad>* seam = &opt;          // the virtual seam LMPC<> holds (LMPC.hpp:731)
    opt.initialize();
    mpc::LParameters p;
    p.maximum_iteration = 250;
    seam->setParameters(p);
    mpc::cvec<12> x0; mpc::cvec<4> u0;
    x0.setZero(); u0.setZero();
    seam->run(x0, u0);
    return (int)seam->result.status;
}
