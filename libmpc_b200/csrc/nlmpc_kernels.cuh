// nlmpc_kernels.cuh -- batched NLMPC problem evaluation for sm_100a: everything libmpc++ hands to its NLP
// solver for one decision vector z, for a batch of independent controllers, one warp per instance:
//   * Mapping::unwrapVector                        include/mpc/NLMPC/Mapping.hpp:174-211  (move blocking via Iz2u :221-257)
//   * Objective::evaluate + computeGradient        include/mpc/NLMPC/Objective.hpp:91-187,198-265  (FORWARD differences)
//   * Constraints::getStateEqConstraints + computeStateEqJacobian + glueJacobian
//                                                  include/mpc/NLMPC/Constraints.hpp:490-628,844-905,455-482 (CENTRAL)
//   * Constraints::evaluateIneq + computeIneqJacobian  include/mpc/NLMPC/Constraints.hpp:211-316,641-721 (CENTRAL)
// including the reference's finite-difference quirks (linear-index step sizes, last-row pairing of U in the objective
// gradient but not in the inequality Jacobian; see oracle/nlmpc_formulation.py).
//
// The reference takes the model / cost / constraints as host std::function callbacks (IDimensionable.hpp:94-149), which
// cannot run on the device: here they are device functors selected by a system id, with the three systems the
// reference ships as examples built in (examples/vanderpol_ex.cpp, networked_oscillators_ex.cpp, ugv_ex.cpp).
// Lanes parallelise over the finite-difference perturbations: every lane evaluates the whole cost / one constraint
// component through an accessor that adds its own perturbation on the fly, so X and U are never copied.
#pragma once
#if !defined(__CUDACC__) && !defined(__CUDACC_RTC__)
// Host emulation (TEST INFRASTRUCTURE, tests/cpp/nl_structured_host.cpp): the per-controller device routines are written as
// `for (i = g.tid; i < N; i += G::nt) ... g.sync()` loops over a thread group, so a "group" of ONE thread runs them
// sequentially and exactly; plain g++ then compiles the same source and the CPU tests compare it with the Python specification
// before anything touches a GPU.  Kernels (__global__) and the warp-shuffle group are compiled out.
#define B200_HOST_EMU 1
#include <cmath>
#include <cstddef>
#include <cstdint>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
#define B200_INF (INFINITY)
using std::fabs; using std::fmax; using std::fmin; using std::sqrt; using std::fma;
#else
#ifndef __CUDACC_RTC__          // NVRTC (user-defined systems, b200mpc_nlmpc_register_system) has these built in
#include <cuda_runtime.h>
#include <math.h>
#endif
#define B200_INF (__longlong_as_double(0x7ff0000000000000LL))
#endif

namespace b200mpc {

// Optional user equality constraints (NLMPC::setEqConFunction, NLMPC.hpp:261-281): a system may define
//   __host__ __device__ static int neq(int ph);   __device__ static double eq(int r, const Acc&, int ph, const double* p);
template <class S, class = void> struct NlHasEq { static constexpr bool value = false; };
template <class S> struct NlHasEq<S, decltype((void)S::neq(1))> { static constexpr bool value = true; };
template <class S> __host__ __device__ inline int nl_neq(int ph) { if constexpr (NlHasEq<S>::value) return S::neq(ph); else return 0; }
// Optional sparsity hint:  static constexpr int ineq_per_stage = K  declares that inequality r reads only row r / K of (X, U)
// (K rows per stage, stage-major: true of every per-stage bound or obstacle constraint).  The reference re-evaluates the
// whole constraint vector for every perturbed variable (Constraints.hpp:641-721: 2 ph (nx+nu) evaluations of all Tineq
// rows); a row that does not read the perturbed variable gives (c - c) / (2 dx) = 0 exactly, so evaluating only the K rows
// of the perturbed stage changes nothing but the cost (ugv Tph=30: 22 320 -> 720 evaluations).  The loop over those K
// rows has the same trip count on every lane (a per-row `skip` test instead made the warp walk all rows anyway, one lane
// at a time: measured 3.5x slower on vanderpol_ex).
template <class S, class = void> struct NlIneqPerStage { static constexpr int value = 0; };
template <class S> struct NlIneqPerStage<S, decltype((void)S::ineq_per_stage)> { static constexpr int value = S::ineq_per_stage; };

// Optional output map (NLMPC::setOutputFunction, NLMPC.hpp:202): a system may define
//   __device__ static void out(double* y, const double* x, const double* u, int i, const double* p);
template <class S, class = void> struct NlHasOut { static constexpr bool value = false; };
template <class S> struct NlHasOut<S, decltype((void)&S::out)> { static constexpr bool value = true; };

// X [(ph+1) x nx], U [(ph+1) x nu] row-major in shared memory + one (or a pair of) perturbed entries
struct Acc {
    const double* X; const double* U; int nx, nu;
    int kind;      // 0 none, 1 X(row,col) += d, 2 U(row,col) += d, 3 U(row,col) and U(row+1,col) += d
    int row, col; double d;
    __device__ __forceinline__ double x(int i, int j) const { double v = X[i * nx + j]; return (kind == 1 && i == row && j == col) ? v + d : v; }
    __device__ __forceinline__ double u(int i, int j) const {
        double v = U[i * nu + j];
        if (kind == 2 && i == row && j == col) v += d;
        if (kind == 3 && (i == row || i == row + 1) && j == col) v += d;
        return v;
    }
};

// Output map (NLMPC::setOutputFunction, NLMPC.hpp:202-215 -> Model::getOutput, Model.hpp:72-96): the reference hands the
// objective and the inequality constraints Y(i,:) = out(X(i,:), U(i,:), i) next to X and U.  A system that has an output map
// defines   __device__ static void out(double* y, const double* x, const double* u, int i, const double* p)   and reads
// nl_y<Self>(a, i, j, p) inside cost / ineq: the map is applied to the (perturbed) row on the fly, exactly what the reference's
// re-evaluation of getOutput inside every finite-difference perturbation computes.
template <class S>
__device__ __forceinline__ double nl_y(const Acc& a, int i, int j, const double* p) {
    double x[S::nx], u[S::nu], y[S::ny];
    for (int k = 0; k < S::nx; ++k) x[k] = a.x(i, k);
    for (int k = 0; k < S::nu; ++k) u[k] = a.u(i, k);
    S::out(y, x, u, i, p);
    return y[j];
}

// ---- built-in systems ---------------------------------------------------------------------------------------------
// examples/vanderpol_ex.cpp:9-71 -- params: [Ts]
struct SysVanDerPol {
    static constexpr int id = 0;
    static constexpr int nx = 2, nu = 1, ny = 2, nparam = 1;
    static constexpr bool continuous = true;
    __device__ static double Ts(const double* p) { return p[0]; }
    __host__ __device__ static int nineq(int ph) { return ph + 1; }
    __device__ static void f(double* dx, const double* x, const double* u, int, const double*) {
        dx[0] = ((1.0 - (x[1] * x[1])) * x[0]) - x[1] + u[0];
        dx[1] = x[0];
    }
    __device__ static double cost(const Acc& a, double, int ph, const double*) {
        double sx = 0, su = 0;
        for (int j = 0; j < nx; ++j) for (int i = 0; i <= ph; ++i) { double v = a.x(i, j); sx += v * v; }   // column-major sum order (Eigen)
        for (int j = 0; j < nu; ++j) for (int i = 0; i <= ph; ++i) { double v = a.u(i, j); su += v * v; }
        return sx + su;
    }
    __device__ static double ineq(int r, const Acc& a, double, int, const double*) { return a.u(r, 0) - 0.5; }
    static constexpr int ineq_per_stage = 1;
};

// examples/networked_oscillators_ex.cpp:5-72 -- params: [Ts, mu, k]
template <int N>
struct SysOscNet {
    static constexpr int id = N == 4 ? 1 : 2;
    static constexpr int nx = 2 * N, nu = N, ny = 2 * N, nparam = 3;
    static constexpr bool continuous = true;
    __device__ static double Ts(const double* p) { return p[0]; }
    __host__ __device__ static int nineq(int ph) { return (ph + 1) * nu; }
    __device__ static void f(double* dx, const double* x, const double* u, int, const double* p) {
        const double mu = p[1], k = p[2];
        for (int i = 0; i < N; ++i) {
            dx[2 * i] = x[2 * i + 1];
            double v = mu * (1 - x[2 * i] * x[2 * i]) * x[2 * i + 1] - x[2 * i] + u[i];
            for (int j = 0; j < N; ++j) if (i != j) v += k * (x[2 * j] - x[2 * i]);
            dx[2 * i + 1] = v;
        }
    }
    __device__ static double cost(const Acc& a, double, int ph, const double*) {
        double sx = 0, su = 0;
        for (int j = 0; j < nx; ++j) for (int i = 0; i <= ph; ++i) { double v = a.x(i, j); sx += v * v; }
        for (int j = 0; j < nu; ++j) for (int i = 0; i <= ph; ++i) { double v = a.u(i, j); su += v * v; }
        return sx + su;
    }
    __device__ static double ineq(int r, const Acc& a, double, int, const double*) { return a.u(r / nu, r % nu) - 0.5; }
    static constexpr int ineq_per_stage = nu;
};

// examples/ugv_ex.cpp:12-126 -- discrete double integrator, 2 circular obstacles, soft constraints.
// params: [Ad(16) | Bd(8) | v_pref(2) | obs0(x,y,r) | obs1(x,y,r)]  (C = I, D = 0: y = x)
struct SysUgv {
    static constexpr int id = 3;
    static constexpr int nx = 4, nu = 2, ny = 4, nparam = 16 + 8 + 2 + 6, nobs = 2;
    static constexpr bool continuous = false;
    __device__ static double Ts(const double*) { return 0.0; }
    __host__ __device__ static int nineq(int ph) { return (ph + 1) * nobs; }
    __device__ static void f(double* xn, const double* x, const double* u, int, const double* p) {
        for (int r = 0; r < 4; ++r) {
            double v = 0;
            for (int c = 0; c < 4; ++c) v += p[r * 4 + c] * x[c];
            double w = 0;
            for (int c = 0; c < 2; ++c) w += p[16 + r * 2 + c] * u[c];
            xn[r] = v + w;
        }
    }
    __device__ static void out(double* y, const double* x, const double*, int, const double*) {      // ugv_ex.cpp:69-77, Cd = I, Dd = 0
        for (int r = 0; r < 4; ++r) y[r] = x[r];
    }
    __device__ static double cost(const Acc& a, double e, int ph, const double* p) {
        double cost = 0;
        for (int i = 0; i <= ph; ++i) {
            double d0 = a.x(i, 2) - p[24], d1 = a.x(i, 3) - p[25];
            cost += 1e3 * (d0 * d0 + d1 * d1);
            double u0 = a.u(i, 0), u1 = a.u(i, 1);
            cost += 1e-2 * (u0 * u0 + u1 * u1);
        }
        cost += 1e-5 * e * e;
        return cost;
    }
    __device__ static double ineq(int r, const Acc& a, double, int, const double* p) {
        int i = r / nobs, j = r % nobs;
        double dx = a.x(i, 0) - p[26 + 3 * j], dy = a.x(i, 1) - p[26 + 3 * j + 1];
        return p[26 + 3 * j + 2] - sqrt(dx * dx + dy * dy);
    }
    static constexpr int ineq_per_stage = nobs;
};

// ---- the threads that cooperate on one controller -------------------------------------------------------------------
// NT = 32: one warp (several controllers per CTA); NT > 32: the whole CTA works on one controller.
#ifdef B200_HOST_EMU
// a "thread group" of one: every cooperative loop runs sequentially (host emulation, see the top of this file)
struct NlGrpHost {
    static constexpr int nt = 1, nw = 1;
    int tid = 0, lane = 0, wid = 0;
    double* red = nullptr;
    void sync() const {}
    double sum(double v) const { return v; }
    double max(double v) const { return v; }
    bool any(bool b) const { return b; }
};
#else
template <int NT>
struct NlGrp {
    static constexpr int nt = NT, nw = NT / 32;
    int tid, lane, wid;
    double* red;                                   // nw doubles of shared scratch (unused when NT == 32)
    __device__ __forceinline__ void sync() const { if (NT == 32) __syncwarp(); else __syncthreads(); }
    __device__ __forceinline__ double sum(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (NT == 32) return v;
        __syncthreads();
        if (lane == 0) red[wid] = v;
        __syncthreads();
        double t = 0;
#pragma unroll
        for (int k = 0; k < nw; ++k) t += red[k];
        return t;
    }
    __device__ __forceinline__ double max(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (NT == 32) return v;
        __syncthreads();
        if (lane == 0) red[wid] = v;
        __syncthreads();
        double t = red[0];
#pragma unroll
        for (int k = 1; k < nw; ++k) t = fmax(t, red[k]);
        return t;
    }
    __device__ __forceinline__ bool any(bool b) const { return NT == 32 ? __any_sync(0xffffffffu, b) : (bool)__syncthreads_or(b); }
};
#endif

// Where nl_eval_instance puts a Jacobian entry.  Dense (the reference's layout): row-major with row stride ld.  Compact (the
// stage-structured solver, nlmpc_structured.cuh): only the columns a row can touch are stored --
//   dynamics row r of stage i = r / nx:      [ d/dX_i (nx) | d/dX_{i+1} (nx) | d/dU_i (nu) ]            (X_0 = x0: first run unused at i = 0)
//   inequality row r of stage i = r / K:     [ d/dX_i (nx) | d/dU_i (nu) | d/dslack ]                   (needs ineq_per_stage = K)
// with U_i standing for the control block stage i reads (move blocking, row ph duplicated).
struct NlDenseMap {
    int ld;
    __device__ __forceinline__ size_t je(int r, int col) const { return (size_t)r * ld + col; }
    __device__ __forceinline__ size_t ji(int r, int col) const { return (size_t)r * ld + col; }
    __device__ __forceinline__ size_t je_total(int me) const { return (size_t)me * ld; }
    __device__ __forceinline__ size_t ji_total(int mi) const { return (size_t)mi * ld; }
};
struct NlCompactMap {
    int ph, ch, nx, nu, K;
    __device__ __forceinline__ size_t je(int r, int col) const {
        const int i = r / nx, we = 2 * nx + nu;
        int q;
        if (col >= ph * nx) q = 2 * nx + (col - ph * nx) % nu;          // the block stage i reads (the only one its rows touch)
        else q = col - (i - 1) * nx;                                       // X_i -> [0, nx), X_{i+1} -> [nx, 2 nx)
        return (size_t)r * we + q;
    }
    __device__ __forceinline__ size_t ji(int r, int col) const {
        const int i = r / K, wi = nx + nu + 1;
        int q;
        if (col == ph * nx + ch * nu) q = nx + nu;                         // slack
        else if (col >= ph * nx) q = (col - ph * nx) % nu + nx;
        else q = col - (i - 1) * nx;                                       // X_i = z block i - 1
        return (size_t)r * wi + q;
    }
    __device__ __forceinline__ size_t je_total(int me) const { return (size_t)me * (2 * nx + nu); }
    __device__ __forceinline__ size_t ji_total(int mi) const { return (size_t)mi * (nx + nu + 1); }
};

struct NlEvalArgs {
    int ph, ch, batch;
    const double* z;        // [batch, nz]
    const double* x0;       // [batch, nx]
    const double* params;   // [batch or 1, nparam]
    long long param_stride;
    double* fval;           // [batch]
    double* grad;           // [batch, nz]
    double* ceq;            // [batch, ph*nx]
    double* Jeq;            // [batch, ph*nx, nz] row-major
    double* cin;            // [batch, nineq]
    double* Jin;            // [batch, nineq, nz] row-major
    double* cue;            // [batch, neq]       user equality constraints (Constraints::evaluateEq)
    double* Jue;            // [batch, neq, nz]
    const double* sx;       // [nx] state scaling  (NLMPC::setStateScale, Mapping.hpp:108-130); null = 1
    const double* su;       // [nu] input scaling  (NLMPC::setInputScale); null = 1
    double* yout;           // [batch, ph+1, ny]  Model::getOutput of the unwrapped sequences (Model.hpp:72-96); may be null
};

// ---- evaluation of one instance by one thread group (shared by the evaluation kernel and by the SQP kernel) ----------
// X,U: shared-memory scratch [(ph+1)*nx], [(ph+1)*nu].  Output pointers may be global or shared; Jacobians are row-major
// with row stride ldj.  Any output pointer may be null.  The whole group must call; ends with a group barrier.
// sx / su: state / input scaling (null = none).  The reference divides X (row 0 = x0 included) by the state scaling and
// multiplies U by the input scaling in unwrapVector (Mapping.hpp:174-211,221-257), divides the dynamics residual by the
// state scaling (Constraints.hpp:528,588), forms the dynamics blocks as I + h Sx A Tx (Constraints.hpp:553-575), scales the
// state columns of the user-constraint Jacobians but NOT of the objective gradient, and maps every input derivative
// through Iz2u (x input scaling).  cue/Jue: user equality constraints, rows [0, neq).
template <class S, class G, class MAP>
__device__ __forceinline__ void nl_eval_instance_map(const G& g_, int ph, int ch, const double* z, const double* x0, const double* p, double* X, double* U,
                                 double* fval, double* grad, double* ceq, double* Jeq, double* cin, double* Jin, int ldj,
                                 const double* sx, const double* su, double* cue, double* Jue, const MAP map) {
    constexpr int nx = S::nx, nu = S::nu;
    const int nz = ph * nx + ch * nu + 1;
    auto SX = [&](int j) { return sx ? sx[j] : 1.0; };
    auto SU = [&](int j) { return su ? su[j] : 1.0; };
    // no FP64 division on the unscaled path (it is the common one and these sit in the innermost loops)
    auto DIVX = [&](double v, int j) { return sx ? v / sx[j] : v; };
    auto RXX = [&](int q, int r) { return sx ? sx[q] / sx[r] : 1.0; };
    auto RUX = [&](int qu, int r) { return sx ? SU(qu) / sx[r] : SU(qu); };
    const double dv = 1.4901161193847656e-08;     // sqrt(DBL_EPSILON)  (Objective.hpp:283)
    // unwrapVector: X row 0 = x0, rows 1..ph from z; U row i = block min(i, ch-1), last row repeated
    for (int e = g_.tid; e < (ph + 1) * nx; e += G::nt) {
        int i = e / nx, j = e - i * nx;
        double v = i == 0 ? x0[j] : z[(i - 1) * nx + j];
        X[e] = sx ? v / sx[j] : v;
    }
    for (int e = g_.tid; e < (ph + 1) * nu; e += G::nt) {
        int i = e / nu, j = e - i * nu;
        int st = i < ph ? i : ph - 1;
        int blk = st < ch ? st : ch - 1;
        double v = z[ph * nx + blk * nu + j];
        U[e] = su ? su[j] * v : v;
    }
    const double slack = z[nz - 1];
    g_.sync();
    Acc base{X, U, nx, nu, 0, 0, 0, 0.0};
    double f0 = 0;
    if (fval || grad) {
        f0 = S::cost(base, slack, ph, p);
        if (fval && g_.tid == 0) *fval = f0;
    }
    if (grad) {
        double* g = grad;
        for (int e = g_.tid; e < nz; e += G::nt) g[e] = 0.0;
        g_.sync();
        for (int t = g_.tid; t < ph * nx; t += G::nt) {          // step uses Xa.array()(j): linear index j (column-major)
            int i = t / nx, j = t - i * nx;
            int lr = j % (ph + 1), lc = j / (ph + 1);
            double dx = dv * fmax(fabs(X[lr * nx + lc]), 1.0);
            Acc ac = base; ac.kind = 1; ac.row = i + 1; ac.col = j; ac.d = dx;
            g[i * nx + j] = (S::cost(ac, slack, ph, p) - f0) / dx;
        }
        for (int t = g_.tid; t < ph * nu; t += G::nt) {          // stage ph-1 moves together with the duplicated row ph
            int i = t / nu, j = t - i * nu;
            int lr = j % (ph + 1), lc = j / (ph + 1);
            double du = dv * fmax(fabs(U[lr * nu + lc]), 1.0);
            Acc ac = base; ac.kind = (i == ph - 1) ? 3 : 2; ac.row = i; ac.col = j; ac.d = du;
            double df = (S::cost(ac, slack, ph, p) - f0) / du;
            int blk = i < ch ? i : ch - 1;
            atomicAdd(&g[ph * nx + blk * nu + j], SU(j) * df);
        }
        if (g_.tid == 0) {
            double ea = fmax(dv, fabs(slack)), de = ea * dv;
            g[nz - 1] = (S::cost(base, slack + de, ph, p) - S::cost(base, slack - de, ph, p)) / (2 * de);
        }
    }
    if (ceq) {
        double* c = ceq;
        double* J = Jeq;
        if (J) for (size_t e = g_.tid; e < map.je_total(ph * nx); e += G::nt) J[e] = 0.0;
        g_.sync();
        const double h = S::Ts(p) / 2.0;
        for (int i = g_.tid; i < ph; i += G::nt) {
            double xk[nx], xk1[nx], uk[nu], fk[nx], fk1[nx];
            for (int j = 0; j < nx; ++j) { xk[j] = X[i * nx + j]; xk1[j] = X[(i + 1) * nx + j]; }
            for (int j = 0; j < nu; ++j) uk[j] = U[i * nu + j];
            if (S::continuous) {
                S::f(fk, xk, uk, i, p); S::f(fk1, xk1, uk, i, p);
                for (int j = 0; j < nx; ++j) c[i * nx + j] = DIVX(xk[j] + (h * (fk[j] + fk1[j])) - xk1[j], j);
            } else {
                S::f(fk, xk, uk, i, p);
                for (int j = 0; j < nx; ++j) c[i * nx + j] = DIVX(xk1[j] - fk[j], j);
            }
        }
        if (J) {
            const int per_stage = nx + nu;
            for (int t = g_.tid; t < ph * per_stage; t += G::nt) {
                int i = t / per_stage, q = t - i * per_stage;
                double xk[nx], xk1[nx], uk[nu], fp[nx], fm[nx];
                for (int j = 0; j < nx; ++j) { xk[j] = X[i * nx + j]; xk1[j] = X[(i + 1) * nx + j]; }
                for (int j = 0; j < nu; ++j) uk[j] = U[i * nu + j];
                int blk = i < ch ? i : ch - 1;
                if (q < nx) {
                    double dx = dv * fmax(fabs(xk[q]), 1.0), keep = xk[q];
                    xk[q] = keep + dx; S::f(fp, xk, uk, i, p);
                    xk[q] = keep - dx; S::f(fm, xk, uk, i, p);
                    xk[q] = keep;
                    if (S::continuous) {
                        if (i > 0) for (int r = 0; r < nx; ++r) J[map.je(i * nx + r, (i - 1) * nx + q)] = (r == q ? 1.0 : 0.0) + h * ((fp[r] - fm[r]) / (2 * dx)) * RXX(q, r);
                        double dx1 = dv * fmax(fabs(xk1[q]), 1.0), keep1 = xk1[q];
                        xk1[q] = keep1 + dx1; S::f(fp, xk1, uk, i, p);
                        xk1[q] = keep1 - dx1; S::f(fm, xk1, uk, i, p);
                        xk1[q] = keep1;
                        for (int r = 0; r < nx; ++r) J[map.je(i * nx + r, i * nx + q)] = (r == q ? -1.0 : 0.0) + h * ((fp[r] - fm[r]) / (2 * dx1)) * RXX(q, r);
                    } else {
                        if (i > 0) for (int r = 0; r < nx; ++r) J[map.je(i * nx + r, (i - 1) * nx + q)] = -((fp[r] - fm[r]) / (2 * dx)) * RXX(q, r);
                        for (int r = 0; r < nx; ++r) J[map.je(i * nx + r, i * nx + q)] = (r == q ? 1.0 : 0.0);
                    }
                } else {
                    int qu = q - nx;
                    double du = dv * fmax(fabs(uk[qu]), 1.0), keep = uk[qu];
                    uk[qu] = keep + du; S::f(fp, xk, uk, i, p);
                    uk[qu] = keep - du; S::f(fm, xk, uk, i, p);
                    uk[qu] = keep;
                    double Bk[nx];
                    for (int r = 0; r < nx; ++r) Bk[r] = (fp[r] - fm[r]) / (2 * du);
                    if (S::continuous) {
                        uk[qu] = keep + du; S::f(fp, xk1, uk, i, p);
                        uk[qu] = keep - du; S::f(fm, xk1, uk, i, p);
                        uk[qu] = keep;
                        for (int r = 0; r < nx; ++r) atomicAdd(&J[map.je(i * nx + r, ph * nx + blk * nu + qu)], h * (Bk[r] + (fp[r] - fm[r]) / (2 * du)) * RUX(qu, r));
                    } else {
                        for (int r = 0; r < nx; ++r) atomicAdd(&J[map.je(i * nx + r, ph * nx + blk * nu + qu)], -Bk[r] * RUX(qu, r));
                    }
                }
            }
        }
    }
    if (cin) {
        const int ni = S::nineq(ph);
        double* c = cin;
        for (int r = g_.tid; r < ni; r += G::nt) c[r] = S::ineq(r, base, slack, ph, p);
        if (Jin) {
            double* J = Jin;
            for (size_t e = g_.tid; e < map.ji_total(ni); e += G::nt) J[e] = 0.0;
            g_.sync();
            for (int t = g_.tid; t < ph * nx; t += G::nt) {
                int i = t / nx, j = t - i * nx;
                int lr = j % (ph + 1), lc = j / (ph + 1);
                double dx = dv * fmax(fabs(X[lr * nx + lc]), 1.0);
                Acc ap = base; ap.kind = 1; ap.row = i + 1; ap.col = j; ap.d = dx;
                Acc am = ap; am.d = -dx;
                constexpr int K = NlIneqPerStage<S>::value;             // K > 0: only the rows of stage i+1 read X(i+1, .)
                const int rlo = K ? (i + 1) * K : 0, rhi = K ? (i + 2) * K : ni;
                for (int r = rlo; r < rhi && r < ni; ++r)
                    J[map.ji(r, i * nx + j)] = (S::ineq(r, ap, slack, ph, p) - S::ineq(r, am, slack, ph, p)) / (2 * dx) * SX(j);
            }
            for (int t = g_.tid; t < ph * nu; t += G::nt) {     // every one of the ph rows alone (row ph is never perturbed)
                int i = t / nu, j = t - i * nu;
                int lr = j % (ph + 1), lc = j / (ph + 1);
                double du = dv * fmax(fabs(U[lr * nu + lc]), 1.0);
                Acc ap = base; ap.kind = 2; ap.row = i; ap.col = j; ap.d = du;
                Acc am = ap; am.d = -du;
                int blk = i < ch ? i : ch - 1;
                constexpr int K = NlIneqPerStage<S>::value;             // K > 0: only the rows of stage i read U(i, .)
                const int rlo = K ? i * K : 0, rhi = K ? (i + 1) * K : ni;
                for (int r = rlo; r < rhi && r < ni; ++r)
                    atomicAdd(&J[map.ji(r, ph * nx + blk * nu + j)], (S::ineq(r, ap, slack, ph, p) - S::ineq(r, am, slack, ph, p)) / (2 * du) * SU(j));
            }
            {
                double ea = fmax(dv, fabs(slack)), de = ea * dv;
                for (int r = g_.tid; r < ni; r += G::nt) J[map.ji(r, nz - 1)] = (S::ineq(r, base, slack + de, ph, p) - S::ineq(r, base, slack - de, ph, p)) / (2 * de);
            }
        }
    }
    if constexpr (NlHasEq<S>::value) {
        // Constraints::evaluateEq + computeEqJacobian (Constraints.hpp:365-442,731-832): c_eq(X,U) = 0, central differences;
        // steps use Xa(i+1,j) for states and Ua(ph-1,j) for EVERY input column (the reference's indexing); the last stage
        // moves together with the duplicated row ph; no slack column.
        if (cue) {
            const int ne = S::neq(ph);
            for (int r = g_.tid; r < ne; r += G::nt) cue[r] = S::eq(r, base, ph, p);
            if (Jue) {
                double* J = Jue;
                for (int e = g_.tid; e < ne * ldj; e += G::nt) J[e] = 0.0;
                g_.sync();
                for (int t = g_.tid; t < ph * nx; t += G::nt) {
                    int i = t / nx, j = t - i * nx;
                    double dx = dv * fmax(fabs(X[(i + 1) * nx + j]), 1.0);
                    Acc ap = base; ap.kind = 1; ap.row = i + 1; ap.col = j; ap.d = dx;
                    Acc am = ap; am.d = -dx;
                    for (int r = 0; r < ne; ++r) J[(size_t)r * ldj + i * nx + j] = (S::eq(r, ap, ph, p) - S::eq(r, am, ph, p)) / (2 * dx) * SX(j);
                }
                for (int t = g_.tid; t < ph * nu; t += G::nt) {
                    int i = t / nu, j = t - i * nu;
                    double du = dv * fmax(fabs(U[(ph - 1) * nu + j]), 1.0);
                    Acc ap = base; ap.kind = (i == ph - 1) ? 3 : 2; ap.row = i; ap.col = j; ap.d = du;
                    Acc am = ap; am.d = -du;
                    int blk = i < ch ? i : ch - 1;
                    for (int r = 0; r < ne; ++r)
                        atomicAdd(&J[(size_t)r * ldj + ph * nx + blk * nu + j], (S::eq(r, ap, ph, p) - S::eq(r, am, ph, p)) / (2 * du) * SU(j));
                }
            }
        }
    }
    g_.sync();
}

template <class S, class G>
__device__ __forceinline__ void nl_eval_instance_impl(const G& g_, int ph, int ch, const double* z, const double* x0, const double* p, double* X, double* U,
                                 double* fval, double* grad, double* ceq, double* Jeq, double* cin, double* Jin, int ldj,
                                 const double* sx, const double* su, double* cue, double* Jue) {
    nl_eval_instance_map<S>(g_, ph, ch, z, x0, p, X, U, fval, grad, ceq, Jeq, cin, Jin, ldj, sx, su, cue, Jue, NlDenseMap{ldj});
}

template <class S, class G>
__device__ __noinline__ void nl_eval_instance_call(G g_, int ph, int ch, const double* z, const double* x0, const double* p, double* X, double* U,
                                 double* fval, double* grad, double* ceq, double* Jeq, double* cin, double* Jin, int ldj,
                                 const double* sx, const double* su, double* cue, double* Jue) {
    nl_eval_instance_impl<S>(g_, ph, ch, z, x0, p, X, U, fval, grad, ceq, Jeq, cin, Jin, ldj, sx, su, cue, Jue);
}
// A warp per controller (tiny problems: the evaluation is a large share of the solve) inlines the evaluation at every call
// site; CTA-sized groups call one shared copy (4x less code to compile, and measurably faster for the shared-memory CTA kernel).
template <class S, class G>
__device__ __forceinline__ void nl_eval_instance(const G& g_, int ph, int ch, const double* z, const double* x0, const double* p, double* X, double* U,
                                 double* fval, double* grad, double* ceq, double* Jeq, double* cin, double* Jin, int ldj,
                                 const double* sx = nullptr, const double* su = nullptr, double* cue = nullptr, double* Jue = nullptr) {
    if constexpr (G::nt == 32) nl_eval_instance_impl<S>(g_, ph, ch, z, x0, p, X, U, fval, grad, ceq, Jeq, cin, Jin, ldj, sx, su, cue, Jue);
    else nl_eval_instance_call<S>(g_, ph, ch, z, x0, p, X, U, fval, grad, ceq, Jeq, cin, Jin, ldj, sx, su, cue, Jue);
}

#ifndef B200_HOST_EMU
template <class S>
__global__ void __launch_bounds__(128) nlmpc_eval_kernel(const NlEvalArgs a) {
    extern __shared__ __align__(16) double nl_smem[];
    constexpr int nx = S::nx, nu = S::nu;
    const int ph = a.ph, ch = a.ch;
    const int nz = ph * nx + ch * nu + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double* X = nl_smem + (size_t)warp * (ph + 1) * (nx + nu);
    double* U = X + (ph + 1) * nx;
    const int ni = S::nineq(ph), nue = nl_neq<S>(ph);
    const NlGrp<32> grp{lane, lane, 0, nullptr};
    for (int inst = blockIdx.x * wpb + warp; inst < a.batch; inst += gridDim.x * wpb) {
        nl_eval_instance<S>(grp, ph, ch, a.z + (size_t)inst * nz, a.x0 + (size_t)inst * nx, a.params + (size_t)inst * a.param_stride, X, U,
                            a.fval ? a.fval + inst : nullptr, a.grad ? a.grad + (size_t)inst * nz : nullptr,
                            a.ceq ? a.ceq + (size_t)inst * ph * nx : nullptr, a.Jeq ? a.Jeq + (size_t)inst * ph * nx * nz : nullptr,
                            a.cin ? a.cin + (size_t)inst * ni : nullptr, a.Jin ? a.Jin + (size_t)inst * ni * nz : nullptr, nz, a.sx, a.su,
                            a.cue ? a.cue + (size_t)inst * nue : nullptr, a.Jue ? a.Jue + (size_t)inst * nue * nz : nullptr);
        if (a.yout) {
            // OptSequence::output = Model::getOutput(Xmat, Umat) (NLOptimizer.hpp:596-611, Model.hpp:72-96): the output map of
            // every row of the unwrapped sequences; all zeros when the system has no output map, as in the reference.
            constexpr int ny = S::ny;
            double* Y = a.yout + (size_t)inst * (ph + 1) * ny;
            for (int i = lane; i <= ph; i += 32) {
                double y[ny > 0 ? ny : 1];
                for (int r = 0; r < ny; ++r) y[r] = 0.0;
                if constexpr (NlHasOut<S>::value) S::out(y, X + i * nx, U + i * nu, i, a.params + (size_t)inst * a.param_stride);
                for (int r = 0; r < ny; ++r) Y[i * ny + r] = y[r];
            }
            __syncwarp();
        }
    }
}

// ---- plant step / RK4 (SURVEY.md 8f N1 + N3) ---------------------------------------------------------------------------
// One thread per controller: x_out = step(x, u) with the system's own model S::f and u held over the step.
//   mode 0  discrete:   x+ = f(x, u)                                   (ugv_ex.cpp:143-166: the loop steps its discrete model)
//   mode 1  Euler:      x+ = x + h f(x, u)                             (vanderpol_ex.cpp:76-85: modelX += modeldX * ts)
//   mode 2  RK4:        `substeps` classical Runge-Kutta steps of size h (mpc::RK4<N>::run, Integrator.hpp:38-56; like the
//                       reference the time argument is NOT advanced between sub-steps: `stage` is passed unchanged)
// u is either given (u_in [batch, nu]) or read from a solved decision vector z (first control block x input scaling = row 0 of
// Umat = Result::cmd, NLOptimizer.hpp:596-606); u_out / status bookkeeping serve the closed loop.
struct NlPlantArgs {
    int batch, mode, substeps, stage;
    double h;
    const double* x;        // [batch, nx]
    const double* u_in;     // [batch, nu] or null
    const double* z;        // [batch, nz] or null (u = su * z[ph*nx .. +nu])
    int u_off, nz;
    const double* su;       // input scaling or null
    const double* params; long long param_stride;
    double* x_out;          // [batch, nx]
    double* u_out;          // [batch, nu] or null
};
template <class S>
__global__ void nlmpc_plant_kernel(const NlPlantArgs a) {
    constexpr int nx = S::nx, nu = S::nu;
    for (int inst = blockIdx.x * blockDim.x + threadIdx.x; inst < a.batch; inst += gridDim.x * blockDim.x) {
        const double* p = a.params + (size_t)inst * a.param_stride;
        double x[nx], u[nu > 0 ? nu : 1], k1[nx], k2[nx], k3[nx], k4[nx], t[nx];
        for (int j = 0; j < nx; ++j) x[j] = a.x[(size_t)inst * nx + j];
        for (int j = 0; j < nu; ++j) {
            double v = a.u_in ? a.u_in[(size_t)inst * nu + j] : a.z[(size_t)inst * a.nz + a.u_off + j] * (a.su ? a.su[j] : 1.0);
            u[j] = v;
            if (a.u_out) a.u_out[(size_t)inst * nu + j] = v;
        }
        if (a.mode == 0) {
            S::f(k1, x, u, a.stage, p);
            for (int j = 0; j < nx; ++j) x[j] = k1[j];
        } else if (a.mode == 1) {
            S::f(k1, x, u, a.stage, p);
            for (int j = 0; j < nx; ++j) x[j] += k1[j] * a.h;
        } else {
            const double h = a.h;
            for (int s = 0; s < a.substeps; ++s) {
                S::f(k1, x, u, a.stage, p);
                for (int j = 0; j < nx; ++j) t[j] = x[j] + (h / 2.0) * k1[j];
                S::f(k2, t, u, a.stage, p);
                for (int j = 0; j < nx; ++j) t[j] = x[j] + (h / 2.0) * k2[j];
                S::f(k3, t, u, a.stage, p);
                for (int j = 0; j < nx; ++j) t[j] = x[j] + h * k3[j];
                S::f(k4, t, u, a.stage, p);
                for (int j = 0; j < nx; ++j) x[j] += h * (k1[j] + 2.0 * k2[j] + 2.0 * k3[j] + k4[j]) / 6.0;
            }
        }
        for (int j = 0; j < nx; ++j) a.x_out[(size_t)inst * nx + j] = x[j];
    }
}

// ---- NLOptimizer::run's initial guess on the device (NLOptimizer.hpp:425-510) ---------------------------------------------
// zprev [batch, nz] is the optimisation vector the previous optimize() left (opt_vector); cold (first iteration or warm start
// off) tiles x0 / u0 over the horizons first; fixOptimalSolution (:705-716) replaces out-of-bound entries by (ub - lb) / 2;
// then the state rows and the per-stage controls (through Iz2u, move blocking) shift left by one stage, the last repeats,
// Iu2z picks the first stage of every block, and the slack entry carries currentSlack.  System independent: one thread per
// (controller, entry).
template <int UNUSED = 0>
__global__ void nlmpc_guess_kernel(int batch, int nx, int nu, int ph, int ch, int cold, const double* x0, const double* u0,
                                   const double* zprev, const double* slack, const double* lb, const double* ub, double* z0) {
    const int nz = ph * nx + ch * nu + 1;
    const long long total = (long long)batch * nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int inst = (int)(t / nz), e = (int)(t - (long long)inst * nz);
        auto src = [&](int k) -> double {                 // entry k of the repaired (cold-tiled or previous) vector
            double v;
            if (k == nz - 1) v = cold ? 0.0 : zprev[(size_t)inst * nz + k];
            else if (cold) v = k < ph * nx ? x0[(size_t)inst * nx + k % nx] : u0[(size_t)inst * nu + (k - ph * nx) % nu];
            else v = zprev[(size_t)inst * nz + k];
            if (v < lb[k] || v > ub[k]) v = (ub[k] - lb[k]) / 2.0;
            return v;
        };
        double v;
        if (e == nz - 1) v = slack ? slack[(size_t)inst * nz] : 0.0;     // `slack` points at entry nz-1 of controller 0's previous vector
        else if (e < ph * nx) {
            int i = e / nx, j = e - i * nx;
            v = src((i == ph - 1 ? i : i + 1) * nx + j);
        } else {
            int c = (e - ph * nx) / nu, j = (e - ph * nx) - c * nu;      // block c <- stage c (Iu2z), shifted stage c+1 (last repeats)
            int st = c == ph - 1 ? c : c + 1;
            int blk = st < ch ? st : ch - 1;                             // Iz2u
            v = src(ph * nx + blk * nu + j);
        }
        z0[t] = v;
    }
}
#endif  // !B200_HOST_EMU

}  // namespace b200mpc
