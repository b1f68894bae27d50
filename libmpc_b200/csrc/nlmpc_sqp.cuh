// nlmpc_sqp.cuh -- batched NLMPC solve for sm_100a: one thread group (a warp for tiny problems, a CTA of
// 4-8 warps otherwise) per controller, the whole NLP solve in one kernel.  Replaces what NLOptimizer::run hands to NLopt's SLSQP (include/mpc/NLMPC/NLOptimizer.hpp:412-638):
//
//     min f(z)   s.t.  c_eq(z) = 0 (multiple-shooting dynamics),  c_in(z) <= 0 (user inequalities),  lb <= z <= ub
//
// with f, its forward-difference gradient and the central-difference Jacobians evaluated exactly as the reference does
// (nl_eval_instance in nlmpc_kernels.cuh).  The solver is a damped-BFGS SQP (the algorithm family of Kraft's SLSQP): every
// major iteration solves   min 1/2 d'Bd + g'd  s.t.  J_eq d = -c_eq,  J_in d <= -c_in,  lb-z <= d <= ub-z
// with a dense OSQP-style ADMM (Ruiz equilibration, rho by row class, over-relaxation 1.6, adaptive rho with dense
// refactorisation) run to moderate accuracy and then polished (OSQP polish.c), globalised by an L1 merit function with
// backtracking and a quasi-Newton restart.  tests/nlmpc_sqp_reference.py is the executable specification; solution-level
// parity is against SciPy's SLSQP on the restated formulation (oracle/nlmpc_slsqp.py).  All vectors live in shared
// memory; the matrices (B, the KKT factor, both Jacobians) too when they fit (vanderpol nz=26, the shipped ugv nz=61),
// otherwise in a per-CTA HBM workspace that stays L2-resident.  The dense O(nz^3) factorisation is what bounds the
// large shapes; the stage-structured (block-tridiagonal) variant is the next step.
#pragma once
#include "nlmpc_kernels.cuh"
#include "nlmpc_structured.cuh"

namespace b200mpc {

struct NlSolveArgs {
    int ph, ch, batch;
    const double* z0;       // [batch, nz] initial decision vectors (NLOptimizer::run's optX0)
    const double* x0;       // [batch, nx]
    const double* params; long long param_stride;
    const double* lb; const double* ub;      // [nz] shared by the batch
    const double* sx; const double* su;      // state / input scaling ([nx], [nu]; null = 1)
    int max_sqp, max_qp;
    double tol, ftol, qp_eps, rho0;
    double* z_out;          // [batch, nz]
    double* cost;           // [batch]
    double* viol;           // [batch] sum |c_eq| + sum max(c_in,0) at the solution
    int* status;            // 0 converged, 1 iteration limit
    int* iters;             // SQP iterations
    int* qp_iters;          // total ADMM iterations
    double* mat_ws;         // MODE 1/2 kernels: per-CTA matrix workspace [grid][NlWs::gmem_doubles]
    int* counter;           // structured kernel: next controller to draw (zeroed by the launcher); null = static round robin
};

__device__ __forceinline__ double nl_wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double nl_wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double nl_lim(double v) { v = v < 1e-4 ? 1.0 : v; return v > 1e4 ? 1e4 : v; }

// Per-controller workspace.  Vectors always live in shared memory.  The KKT factor H is held PACKED (lower triangle,
// row-major: element (i,j), j <= i, at i(i+1)/2 + j -- row walks by consecutive threads are bank-conflict free because
// triangular numbers are a permutation mod 32, column walks are contiguous).  Residency (template parameter MODE):
//   0: B, J_eq, J_in and H in shared memory                       (vanderpol, the shipped ugv)
//   1: H in shared memory, B / J_eq / J_in in a per-CTA HBM workspace (L2-resident)   (ugv Tph=30, oscillator network N=4)
//   2: everything in the HBM workspace                            (nz > ~190)
// Matrix passes over B / J run with consecutive threads along the contiguous (column) index.
struct NlWs {
    int n, me, mi, m, ld, nx, nu, ph, ch;
    int mii;                                 // the first mii rows of the dense block (J_in) are inequalities, the rest user equalities
    double *B, *H, *Je, *Ji;                 // n x ld, packed n(n+1)/2, me x ld, mi x ld (row-major)
    double *z, *g, *g2, *d, *xs, *xt, *D, *gs, *rhs, *tmp, *glo, *sv, *zt2;     // n
    double *E, *ls, *us, *zs, *ys, *rho, *yq, *w, *pr, *pt;                     // m
    double *ce, *ci, *cet, *cit;                                                // me, mi, me, mi
    double *muv;                                                                // me + mi: L1-merit penalty per constraint row
    double *X, *U, *red;
    __host__ __device__ static int ldim(int n, bool global) { return global ? ((n + 3) & ~3) : (n | 1); }
    __host__ __device__ static size_t h_doubles(int n) { return ((size_t)n * (n + 1) / 2 + 1) & ~(size_t)1; }
    __host__ __device__ static size_t mat_doubles(int n, int me, int mi, bool global) { return (size_t)(n + me + mi) * ldim(n, global); }
    __host__ __device__ static size_t vec_doubles(int n, int me, int mi, int ph, int nx, int nu) {
        int m = me + mi + n;
        return 13 * (size_t)n + 10 * (size_t)m + 3 * (size_t)(me + mi) + (size_t)(ph + 1) * (nx + nu) + 16;
    }
    // shared / global doubles one controller needs in each residency mode
    __host__ __device__ static size_t smem_doubles(int mode, int n, int me, int mi, int ph, int nx, int nu) {
        return vec_doubles(n, me, mi, ph, nx, nu) + (mode <= 1 ? h_doubles(n) : 0) + (mode == 0 ? mat_doubles(n, me, mi, false) : 0);
    }
    __host__ __device__ static size_t gmem_doubles(int mode, int n, int me, int mi) {
        return (mode >= 1 ? mat_doubles(n, me, mi, true) : 0) + (mode == 2 ? h_doubles(n) : 0);
    }
    // what carve() was called with: the shared subroutines take this by value and rebuild their own NlWs, so that the caller's
    // copy never has its address taken and stays in registers (an NlWs passed by reference lives in local memory and every
    // pointer field becomes a local load)
    struct Key { double* sm; double* gm; int mode, n, me, mi, mii, ph, ch, nx, nu; };
    Key key;
    __device__ void carve(int mode, double* sm, double* gm, int n_, int me_, int mi_, int ph_, int ch_, int nx_, int nu_) {
        key = Key{sm, gm, mode, n_, me_, mi_, mi_, ph_, ch_, nx_, nu_};
        n = n_; me = me_; mi = mi_; m = me + mi + n; ld = ldim(n, mode >= 1); nx = nx_; nu = nu_; ph = ph_; ch = ch_;
        double*& pm = mode >= 1 ? gm : sm;
        B = pm; pm += (size_t)n * ld; Je = pm; pm += (size_t)me * ld; Ji = pm; pm += (size_t)mi * ld;
        double*& phh = mode == 2 ? gm : sm;
        H = phh; phh += h_doubles(n);
        double* p = sm;
        double** nv[] = {&z, &g, &g2, &d, &xs, &xt, &D, &gs, &rhs, &tmp, &glo, &sv, &zt2};
        for (auto q : nv) { *q = p; p += n; }
        double** mv[] = {&E, &ls, &us, &zs, &ys, &rho, &yq, &w, &pr, &pt};
        for (auto q : mv) { *q = p; p += m; }
        ce = p; p += me; ci = p; p += mi; cet = p; p += me; cit = p; p += mi; muv = p; p += me + mi;
        X = p; p += (ph + 1) * nx; U = p; p += (ph + 1) * nu; red = p;
    }
    __device__ __forceinline__ static size_t tri(int i) { return (size_t)i * (i + 1) / 2; }
    // Multiple-shooting sparsity of J_eq (Constraints.hpp:844-905): the rows of stage s touch X_{s-1}, X_s and the control
    // block of stage s; column j is touched by the rows of at most two stages (states) or of its block's stages (inputs).
    __device__ __forceinline__ void je_cols(int r, int& c0, int& c1, int& u0) const {
        int s = r / nx; c0 = (s > 0 ? s - 1 : 0) * nx; c1 = (s + 1) * nx; u0 = ph * nx + (s < ch ? s : ch - 1) * nu;
    }
    __device__ __forceinline__ void je_rows(int j, int& r0, int& r1) const {
        if (j < ph * nx) { int s = j / nx; r0 = s * nx; r1 = (s + 2) * nx; if (r1 > me) r1 = me; }
        else if (j < n - 1) { int b = (j - ph * nx) / nu; r0 = b * nx; r1 = (b < ch - 1) ? (b + 1) * nx : me; }
        else { r0 = r1 = 0; }
    }
};

// sum_r Je[r][j] v[r] + sum_r Ji[r][j] v[me + r]  for one column j
__device__ __forceinline__ double nl_col_dot(const NlWs& w, int j, const double* v) {
    int r0, r1; w.je_rows(j, r0, r1);
    double a = 0;
    for (int r = r0; r < r1; ++r) a = fma(w.Je[(size_t)r * w.ld + j], v[r], a);
#pragma unroll 4
    for (int r = 0; r < w.mi; ++r) a = fma(w.Ji[(size_t)r * w.ld + j], v[w.me + r], a);
    return a;
}

// H = c D B D + sigma I + (E A D)' diag(rho) (E A D)  -> Cholesky -> in-place inverse of the factor (packed lower).
template <class G>
__device__ __forceinline__ bool nl_factor_impl(const G& g, NlWs& w, double c, double sigma) {
    const int n = w.n, ld = w.ld, me = w.me, mi = w.mi, mc = me + mi;
    for (int r = g.tid; r < mc; r += G::nt) w.w[r] = w.rho[r] * w.E[r] * w.E[r];
    g.sync();
    for (int i = g.wid; i < n; i += G::nw) {            // row i of the lower triangle: a warp per row, lanes along j <= i
        int r0, r1; w.je_rows(i, r0, r1);
        const double di = w.D[i];
        double* hrow = w.H + NlWs::tri(i);
        for (int j = g.lane; j <= i; j += 32) {
            double acc = 0;
            for (int r = r0; r < r1; ++r) { const double* a = w.Je + (size_t)r * ld; acc = fma(w.w[r] * a[i], a[j], acc); }
            for (int r = 0; r < mi; ++r) {
                const double* a = w.Ji + (size_t)r * ld;
                double ai = a[i];
                if (ai != 0.0) acc = fma(w.w[me + r] * ai, a[j], acc);
            }
            double v = di * (c * w.B[(size_t)i * ld + j] + acc) * w.D[j];
            if (i == j) { int rb = mc + i; v += sigma + w.rho[rb] * w.E[rb] * w.E[rb] * di * di; }
            hrow[j] = v;
        }
    }
    g.sync();
    bool ok = true;
    for (int k = 0; k < n; ++k) {                       // right-looking Cholesky, in place; column k staged in tmp
        double dkk = w.H[NlWs::tri(k) + k];
        if (!(dkk > 0.0)) ok = false;
        double piv = sqrt(dkk), inv = 1.0 / piv;
        g.sync();
        for (int r = k + g.tid; r < n; r += G::nt) {
            double* e = w.H + NlWs::tri(r) + k;
            double v = (r == k) ? piv : *e * inv;
            *e = v; w.tmp[r] = v;
        }
        g.sync();
        for (int r = k + 1 + g.wid; r < n; r += G::nw) {
            const double lrk = w.tmp[r];
            double* hrow = w.H + NlWs::tri(r);
            for (int q = k + 1 + g.lane; q <= r; q += 32) hrow[q] -= lrk * w.tmp[q];
        }
        g.sync();
    }
    // in-place inverse of the lower-triangular factor, column by column from the right:
    // Linv[r][j] = -(1/L[j][j]) * sum_{q=j+1..r} Linv[r][q] L[q][j]
    for (int j = n - 1; j >= 0; --j) {
        const double ljj = 1.0 / w.H[NlWs::tri(j) + j];
        for (int q = j + 1 + g.tid; q < n; q += G::nt) w.tmp[q] = w.H[NlWs::tri(q) + j];     // L[:, j]
        g.sync();
        for (int r = j + 1 + g.tid; r < n; r += G::nt) {
            double* hrow = w.H + NlWs::tri(r);
            double a0 = 0, a1 = 0;
            int q = j + 1;
            for (; q + 1 <= r; q += 2) { a0 = fma(hrow[q], w.tmp[q], a0); a1 = fma(hrow[q + 1], w.tmp[q + 1], a1); }
            if (q <= r) a0 = fma(hrow[q], w.tmp[q], a0);
            hrow[j] = -ljj * (a0 + a1);
        }
        if (g.tid == 0) w.H[NlWs::tri(j) + j] = ljj;
        g.sync();
    }
    return !g.any(!ok);
}

template <class G>
__device__ __noinline__ bool nl_factor_call(G g, NlWs::Key k, double c, double sigma) {
    NlWs w; w.carve(k.mode, k.sm, k.gm, k.n, k.me, k.mi, k.ph, k.ch, k.nx, k.nu); w.mii = k.mii;
    return nl_factor_impl(g, w, c, sigma);
}
// a warp per controller inlines (tiny problems, measured 82k vs 60k solves/s on vanderpol_ex); CTA groups share one copy
template <class G>
__device__ __forceinline__ bool nl_factor(const G& g, NlWs& w, double c, double sigma) {
    if constexpr (G::nt == 32) return nl_factor_impl(g, w, c, sigma); else { NlWs::Key k = w.key; k.mii = w.mii; return nl_factor_call(g, k, c, sigma); }
}

// xt = (Linv' Linv) rhs ; optionally dxt = D .* xt (the input nl_As_core wants)
template <class G>
__device__ __forceinline__ void nl_kkt_apply(const G& g, NlWs& w, double* dxt = nullptr) {
    const int n = w.n;
    for (int i = g.tid; i < n; i += G::nt) {            // row walk: conflict-free across consecutive rows (packed storage)
        const double* hrow = w.H + NlWs::tri(i);
        double a0 = 0, a1 = 0;
        int q = 0;
        for (; q + 1 <= i; q += 2) { a0 = fma(hrow[q], w.rhs[q], a0); a1 = fma(hrow[q + 1], w.rhs[q + 1], a1); }
        if (q <= i) a0 = fma(hrow[q], w.rhs[q], a0);
        w.tmp[i] = a0 + a1;
    }
    g.sync();
    for (int i = g.tid; i < n; i += G::nt) {            // column walk: contiguous across threads
        double a0 = 0, a1 = 0;
        size_t t = NlWs::tri(i) + i;
        int q = i;
        for (; q + 1 < n; q += 2) { a0 = fma(w.H[t], w.tmp[q], a0); t += q + 1; a1 = fma(w.H[t], w.tmp[q + 1], a1); t += q + 2; }
        if (q < n) a0 = fma(w.H[t], w.tmp[q], a0);
        double v = a0 + a1;
        w.xt[i] = v;
        if (dxt) dxt[i] = w.D[i] * v;
    }
    g.sync();
}
// out_r = E_r * (A dx)_r for all m rows, dx = D.*x already formed: J_eq rows by their two column runs, J_in rows one
// warp per row.  Ends with a barrier.
template <class G>
__device__ __forceinline__ void nl_As_core(const G& g, NlWs& w, const double* dx, double* out) {
    const int n = w.n, me = w.me, mc = w.me + w.mi, ld = w.ld;
    for (int r = g.tid; r < me; r += G::nt) {
        int c0, c1, u0; w.je_cols(r, c0, c1, u0);
        const double* row = w.Je + (size_t)r * ld;
        double a = 0;
        for (int j = c0; j < c1; ++j) a = fma(row[j], dx[j], a);
        for (int j = u0; j < u0 + w.nu; ++j) a = fma(row[j], dx[j], a);
        out[r] = w.E[r] * a;
    }
    for (int r = g.wid; r < w.mi; r += G::nw) {
        const double* row = w.Ji + (size_t)r * ld;
        double a = 0;
        for (int j = g.lane; j < n; j += 32) a = fma(row[j], dx[j], a);
        a = nl_wsum(a);
        if (g.lane == 0) out[me + r] = w.E[me + r] * a;
    }
    for (int j = g.tid; j < n; j += G::nt) out[mc + j] = w.E[mc + j] * dx[j];
    g.sync();
}
template <class G>
__device__ __forceinline__ void nl_As(const G& g, NlWs& w, const double* x, double* out) {
    for (int i = g.tid; i < w.n; i += G::nt) w.tmp[i] = w.D[i] * x[i];
    g.sync();
    nl_As_core(g, w, w.tmp, out);
}
// out_j = D_j * (A' (E.v))_j
template <class G>
__device__ __forceinline__ void nl_Ats(const G& g, NlWs& w, const double* v, double* out) {
    const int n = w.n, mc = w.me + w.mi;
    for (int r = g.tid; r < w.m; r += G::nt) w.w[r] = w.E[r] * v[r];
    g.sync();
    for (int j = g.tid; j < n; j += G::nt) out[j] = w.D[j] * (w.w[mc + j] + nl_col_dot(w, j, w.w));
    g.sync();
}
// out_i = (c D B D x)_i ; B symmetric, read down its columns
template <class G>
__device__ __forceinline__ void nl_Ps(const G& g, NlWs& w, double c, const double* x, double* out) {
    const int n = w.n, ld = w.ld;
    for (int i = g.tid; i < n; i += G::nt) w.tmp[i] = w.D[i] * x[i];
    g.sync();
    for (int i = g.tid; i < n; i += G::nt) {
        double acc = 0;
#pragma unroll 4
        for (int j = 0; j < n; ++j) acc = fma(w.B[(size_t)j * ld + i], w.tmp[j], acc);
        out[i] = c * w.D[i] * acc;
    }
    g.sync();
}

// max bound violation of A x and max |P x + q + A' y| in the scaled problem (uses pt, rhs, zt2)
template <class G>
__device__ __forceinline__ void nl_qp_residuals(const G& g, NlWs& w, double c, const double* x, const double* y, double& pri, double& dua) {
    nl_As(g, w, x, w.pt);
    double p = 0, d = 0;
    for (int r = g.tid; r < w.m; r += G::nt) p = fmax(p, fmax(fmax(w.ls[r] - w.pt[r], w.pt[r] - w.us[r]), 0.0));
    nl_Ats(g, w, y, w.rhs);
    nl_Ps(g, w, c, x, w.zt2);
    for (int i = g.tid; i < w.n; i += G::nt) d = fmax(d, fabs(w.zt2[i] + w.gs[i] + w.rhs[i]));
    pri = g.max(p); dua = g.max(d);
}

// OSQP polish.c on the dense QP: guess the active set from (z, y), solve the equality-constrained QP on it through the
// same reduced system (rho = 1/delta on active rows, sigma = delta) with iterative refinement, keep the result if both
// residuals improve.  In: xs, ys, zs.  Out: xs, ys (replaced when accepted).  Destroys H, rho, zs.
template <class G>
__device__ __forceinline__ bool nl_qp_polish_impl(const G& g, NlWs& w, double c) {
    const int n = w.n, m = w.m;
    const double delta = 1e-6, idelta = 1e6;
    double pri_a, dua_a;
    nl_qp_residuals(g, w, c, w.xs, w.ys, pri_a, dua_a);
    for (int r = g.tid; r < m; r += G::nt) {
        bool lo = (w.zs[r] - w.ls[r]) < -w.ys[r], up = (w.us[r] - w.zs[r]) < w.ys[r];
        w.rho[r] = (lo || up) ? idelta : 0.0;
        w.zs[r] = lo ? w.ls[r] : (up ? w.us[r] : 0.0);              // b on the active rows
        w.yq[r] = 0.0;                                              // y_p
    }
    for (int i = g.tid; i < n; i += G::nt) w.g2[i] = 0.0;           // x_p
    g.sync();
    if (!nl_factor(g, w, c, delta)) return false;
    for (int it = 0; it <= 5; ++it) {                               // first pass = the plain solve, then 5 refinements
        nl_Ats(g, w, w.yq, w.rhs);
        nl_Ps(g, w, c, w.g2, w.zt2);
        for (int i = g.tid; i < n; i += G::nt) w.sv[i] = -w.gs[i] - w.zt2[i] - w.rhs[i];          // r1
        nl_As(g, w, w.g2, w.pt);
        for (int r = g.tid; r < m; r += G::nt) w.pr[r] = w.rho[r] > 0.0 ? (w.zs[r] - w.pt[r]) * idelta : 0.0;   // r2 / delta
        g.sync();
        nl_Ats(g, w, w.pr, w.rhs);
        for (int i = g.tid; i < n; i += G::nt) w.rhs[i] += w.sv[i];
        g.sync();
        nl_kkt_apply(g, w);
        nl_As(g, w, w.xt, w.pt);
        for (int r = g.tid; r < m; r += G::nt) if (w.rho[r] > 0.0) w.yq[r] += w.pt[r] * idelta - w.pr[r];
        for (int i = g.tid; i < n; i += G::nt) w.g2[i] += w.xt[i];
        g.sync();
    }
    double pri_p, dua_p;
    nl_qp_residuals(g, w, c, w.g2, w.yq, pri_p, dua_p);
    const bool ok = pri_p <= fmax(pri_a, 1e-10) && dua_p <= fmax(dua_a, 1e-10);
    if (ok) {
        for (int i = g.tid; i < n; i += G::nt) w.xs[i] = w.g2[i];
        for (int r = g.tid; r < m; r += G::nt) w.ys[r] = w.yq[r];
        g.sync();
    }
    return ok;
}

template <class G>
__device__ __noinline__ bool nl_qp_polish_call(G g, NlWs::Key k, double c) {
    NlWs w; w.carve(k.mode, k.sm, k.gm, k.n, k.me, k.mi, k.ph, k.ch, k.nx, k.nu); w.mii = k.mii;
    return nl_qp_polish_impl(g, w, c);
}
template <class G>
__device__ __forceinline__ bool nl_qp_polish(const G& g, NlWs& w, double c) {
    if constexpr (G::nt == 32) return nl_qp_polish_impl(g, w, c); else { NlWs::Key k = w.key; k.mii = w.mii; return nl_qp_polish_call(g, k, c); }
}

// Dense OSQP-style ADMM for the QP subproblem.  In: B, g, Je, Ji, ce, ci, z, lb, ub; warm dual yq (if have_y).
// Out: d (step), yq (multipliers, unscaled).  Returns ADMM iterations.
template <class G>
__device__ __forceinline__ int nl_qp_solve(const G& g, NlWs& w, const NlSolveArgs& a, bool have_y, int max_qp) {
    const int n = w.n, me = w.me, mi = w.mi, mc = me + mi, m = w.m, ld = w.ld;
    const double sigma = 1e-6, alpha = 1.6;
    // ---- Ruiz equilibration (10 passes) with cost normalisation
    for (int i = g.tid; i < n; i += G::nt) { w.D[i] = 1.0; w.gs[i] = w.g[i]; }
    for (int r = g.tid; r < m; r += G::nt) w.E[r] = 1.0;
    double c = 1.0;
    g.sync();
    for (int pass = 0; pass < 10; ++pass) {
        for (int j = g.tid; j < n; j += G::nt) {    // column norms -> xt ; uses old D,E
            double cn = 0;
#pragma unroll 4
            for (int i = 0; i < n; ++i) cn = fmax(cn, w.D[i] * fabs(w.B[(size_t)i * ld + j]));
            cn *= c * w.D[j];
            double an = 0;
            int r0, r1; w.je_rows(j, r0, r1);
            for (int r = r0; r < r1; ++r) an = fmax(an, w.E[r] * fabs(w.Je[(size_t)r * ld + j]));
#pragma unroll 4
            for (int r = 0; r < mi; ++r) an = fmax(an, w.E[me + r] * fabs(w.Ji[(size_t)r * ld + j]));
            an = fmax(an, w.E[mc + j]) * w.D[j];
            w.xt[j] = 1.0 / sqrt(nl_lim(fmax(cn, an)));
        }
        for (int r = g.tid; r < me; r += G::nt) {   // row norms -> w
            int c0, c1, u0; w.je_cols(r, c0, c1, u0);
            const double* row = w.Je + (size_t)r * ld;
            double rn = 0;
            for (int j = c0; j < c1; ++j) rn = fmax(rn, fabs(row[j]) * w.D[j]);
            for (int j = u0; j < u0 + w.nu; ++j) rn = fmax(rn, fabs(row[j]) * w.D[j]);
            w.w[r] = 1.0 / sqrt(nl_lim(rn * w.E[r]));
        }
        for (int r = g.wid; r < mi; r += G::nw) {
            const double* row = w.Ji + (size_t)r * ld;
            double rn = 0;
            for (int j = g.lane; j < n; j += 32) rn = fmax(rn, fabs(row[j]) * w.D[j]);
            rn = nl_wmax(rn);
            if (g.lane == 0) w.w[me + r] = 1.0 / sqrt(nl_lim(rn * w.E[me + r]));
        }
        for (int j = g.tid; j < n; j += G::nt) w.w[mc + j] = 1.0 / sqrt(nl_lim(w.E[mc + j] * w.D[j]));
        g.sync();
        for (int j = g.tid; j < n; j += G::nt) { w.D[j] *= w.xt[j]; w.gs[j] *= w.xt[j]; }
        for (int r = g.tid; r < m; r += G::nt) w.E[r] *= w.w[r];
        g.sync();
        double psum = 0, qmax = 0;
        for (int j = g.tid; j < n; j += G::nt) {
            double cn = 0;
#pragma unroll 4
            for (int i = 0; i < n; ++i) cn = fmax(cn, w.D[i] * fabs(w.B[(size_t)i * ld + j]));
            psum += c * cn * w.D[j];
            qmax = fmax(qmax, fabs(w.gs[j]));
        }
        psum = g.sum(psum); qmax = g.max(qmax);
        double ct = fmax(psum / n, nl_lim(qmax));
        ct = 1.0 / nl_lim(ct);
        c *= ct;
        for (int j = g.tid; j < n; j += G::nt) w.gs[j] *= ct;
        g.sync();
    }
    // ---- scaled bounds, rho per row
    double rho0 = a.rho0;
    for (int r = g.tid; r < m; r += G::nt) {
        double l, u;
        if (r < me) { l = u = -w.ce[r]; }
        else if (r < mc) { u = -w.ci[r - me]; l = (r - me) < w.mii ? -B200_INF : u; }
        else { int j = r - mc; l = a.lb[j] - w.z[j]; u = a.ub[j] - w.z[j]; }
        w.ls[r] = w.E[r] * l; w.us[r] = w.E[r] * u;
    }
    g.sync();
    auto set_rho = [&](double r0) {
        for (int r = g.tid; r < m; r += G::nt)     // OSQP's row classes: no bounds -> RHO_MIN, equality -> 1e3 rho
            w.rho[r] = (w.ls[r] < -1e20 && w.us[r] > 1e20) ? 1e-6 : ((w.us[r] - w.ls[r]) < 1e-9) ? 1e3 * r0 : r0;
        g.sync();
    };
    set_rho(rho0);
    nl_factor(g, w, c, sigma);
    // ---- start: x = 0, y = warm (scaled), z = clip(A x)
    for (int i = g.tid; i < n; i += G::nt) w.xs[i] = 0.0;
    for (int r = g.tid; r < m; r += G::nt) { w.ys[r] = have_y ? c * w.yq[r] / w.E[r] : 0.0; w.zs[r] = fmin(fmax(0.0, w.ls[r]), w.us[r]); }
    g.sync();
    int it = 0;
    for (it = 1; it <= max_qp; ++it) {
        // rhs = sigma x - q + A'(rho z - y)  (scaled), 2 barriers
        for (int r = g.tid; r < m; r += G::nt) w.w[r] = w.E[r] * (w.rho[r] * w.zs[r] - w.ys[r]);
        g.sync();
        for (int j = g.tid; j < n; j += G::nt)
            w.rhs[j] = w.D[j] * (w.w[mc + j] + nl_col_dot(w, j, w.w)) + sigma * w.xs[j] - w.gs[j];
        g.sync();
        nl_kkt_apply(g, w, w.zt2);                                                      // x~ in xt, D.*x~ in zt2
        nl_As_core(g, w, w.zt2, w.yq);                                                  // z~ in yq
        for (int i = g.tid; i < n; i += G::nt) w.xs[i] = alpha * w.xt[i] + (1 - alpha) * w.xs[i];
        for (int r = g.tid; r < m; r += G::nt) {
            double zr = alpha * w.yq[r] + (1 - alpha) * w.zs[r];
            double zn = fmin(fmax(zr + w.ys[r] / w.rho[r], w.ls[r]), w.us[r]);
            w.ys[r] += w.rho[r] * (zr - zn);
            w.zs[r] = zn;
        }
        g.sync();
        if (it % 25 == 0) {
            nl_As(g, w, w.xs, w.yq);                          // Ax
            double pri = 0, nz = 0, nAx = 0;
            for (int r = g.tid; r < m; r += G::nt) { pri = fmax(pri, fabs(w.yq[r] - w.zs[r])); nz = fmax(nz, fabs(w.zs[r])); nAx = fmax(nAx, fabs(w.yq[r])); }
            nl_Ats(g, w, w.ys, w.rhs);                        // A'y
            nl_Ps(g, w, c, w.xs, w.zt2);                      // Px
            double dua = 0, nq = 0, nAty = 0, nPx = 0;
            for (int i = g.tid; i < n; i += G::nt) {
                double px = w.zt2[i];
                dua = fmax(dua, fabs(px + w.gs[i] + w.rhs[i])); nq = fmax(nq, fabs(w.gs[i])); nAty = fmax(nAty, fabs(w.rhs[i])); nPx = fmax(nPx, fabs(px));
            }
            pri = g.max(pri); nz = g.max(nz); nAx = g.max(nAx); dua = g.max(dua); nq = g.max(nq); nAty = g.max(nAty); nPx = g.max(nPx);
            if (pri < a.qp_eps && dua < a.qp_eps) break;
            double pn = pri / (fmax(nz, nAx) + 1e-10), dn = dua / (fmax(fmax(nq, nAty), nPx) + 1e-10);
            double est = fmin(fmax(rho0 * sqrt(pn / (dn + 1e-10)), 1e-6), 1e6);
            if (est > 5 * rho0 || est < rho0 / 5) { rho0 = est; set_rho(rho0); nl_factor(g, w, c, sigma); }
        }
    }
    if (it > max_qp) it = max_qp;
    nl_qp_polish(g, w, c);
    for (int i = g.tid; i < n; i += G::nt) w.d[i] = w.D[i] * w.xs[i];
    for (int r = g.tid; r < m; r += G::nt) w.yq[r] = w.E[r] * w.ys[r] / c;
    g.sync();
    return it;
}

// One controller per thread group of NT threads: NT = 32 -> one warp, several controllers per CTA; NT > 32 -> the CTA.
template <class S, int MODE, int NT>
__global__ void __launch_bounds__(NT == 32 ? 64 : NT) nlmpc_solve_kernel(const NlSolveArgs a) {
    extern __shared__ __align__(16) double nls_smem[];
    constexpr int nx = S::nx, nu = S::nu;
    using G = NlGrp<NT>;
    const int ph = a.ph, ch = a.ch;
    const int mii = S::nineq(ph), mue = nl_neq<S>(ph);
    const int n = ph * nx + ch * nu + 1, me = ph * nx, mi = mii + mue;
    const int gpb = blockDim.x / NT, gi = threadIdx.x / NT;            // groups per CTA, this thread's group
    const size_t nsm = NlWs::smem_doubles(MODE, n, me, mi, ph, nx, nu), ngm = NlWs::gmem_doubles(MODE, n, me, mi);
    NlWs w;
    w.carve(MODE, nls_smem + (size_t)gi * nsm, MODE ? a.mat_ws + (size_t)(blockIdx.x * gpb + gi) * ngm : nullptr, n, me, mi, ph, ch, nx, nu);
    w.mii = mii;
    const G g{(int)(threadIdx.x % NT), (int)(threadIdx.x & 31), (int)((threadIdx.x % NT) >> 5), w.red};
    const int mc = me + mi, ld = w.ld;
    for (int inst = blockIdx.x * gpb + gi; inst < a.batch; inst += gridDim.x * gpb) {
        const double* p = a.params + (size_t)inst * a.param_stride;
        const double* x0 = a.x0 + (size_t)inst * nx;
        for (int i = g.tid; i < n; i += NT) w.z[i] = fmin(fmax(a.z0[(size_t)inst * n + i], a.lb[i]), a.ub[i]);
        for (int e = g.tid; e < n * ld; e += NT) { int i = e / ld, j = e - i * ld; w.B[e] = (i == j) ? 1.0 : 0.0; }
        g.sync();
        double fval = 0;
        nl_eval_instance<S>(g, ph, ch, w.z, x0, p, w.X, w.U, w.tmp, w.g, w.ce, w.Je, w.ci, w.Ji, ld, a.sx, a.su, w.ci + mii, w.Ji + (size_t)mii * ld);
        fval = w.tmp[0];
        g.sync();
        bool have_y = false;
        int k = 0, qp_total = 0, status = 1, resets = 0;
        bool just_reset = false;
        int qp_cap = a.max_qp;             // ADMM iteration cap of the QP subproblems; raised once when a step fails (below)
        auto violation = [&](const double* ce, const double* ci) {
            double v = 0;
            for (int r = g.tid; r < me; r += NT) v += fabs(ce[r]);
            for (int r = g.tid; r < mi; r += NT) v += r < mii ? fmax(ci[r], 0.0) : fabs(ci[r]);
            return g.sum(v);
        };
        // L1 merit with one penalty per constraint row, updated by Powell's rule as in Kraft's SLSQP:
        // mu_r <- max(|lambda_r|, (mu_r + |lambda_r|) / 2).  (A single monotone penalty max_r |lambda_r| -- the first version --
        // stays at the largest multiplier ever seen and forces tiny steps along curved dynamics: unicycle Tph=30, 300 iterations.)
        auto merit_violation = [&](const double* ce, const double* ci) {
            double v = 0;
            for (int r = g.tid; r < me; r += NT) v += w.muv[r] * fabs(ce[r]);
            for (int r = g.tid; r < mi; r += NT) v += w.muv[me + r] * (r < mii ? fmax(ci[r], 0.0) : fabs(ci[r]));
            return g.sum(v);
        };
        for (k = 0; k < a.max_sqp; ++k) {
            qp_total += nl_qp_solve(g, w, a, have_y, qp_cap);
            have_y = true;
            // L1 merit line search
            double v0 = violation(w.ce, w.ci);
            double gd = 0;
            for (int r = g.tid; r < mc; r += NT) {
                const double lam = fabs(w.yq[r]);
                w.muv[r] = (k == 0 || just_reset) ? lam : fmax(lam, 0.5 * (w.muv[r] + lam));
            }
            for (int i = g.tid; i < n; i += NT) gd += w.g[i] * w.d[i];
            gd = g.sum(gd);
            g.sync();
            const double mv0 = merit_violation(w.ce, w.ci);
            const double phi0 = fval + mv0, dphi = gd - mv0;
            // Kraft's first stopping test (|g'd| and the violation below the accuracy): nothing left to gain
            if (fabs(gd) < a.ftol * fmax(1.0, fabs(fval)) && v0 < 1e-8) { status = 0; ++k; break; }
            double t = 1.0, ft = fval;
            bool ls_ok = false;
            for (int ls = 0; ls < 25; ++ls) {
                for (int i = g.tid; i < n; i += NT) w.zt2[i] = w.z[i] + t * w.d[i];
                g.sync();
                nl_eval_instance<S>(g, ph, ch, w.zt2, x0, p, w.X, w.U, w.tmp, nullptr, w.cet, nullptr, w.cit, nullptr, ld, a.sx, a.su, w.cit + mii, nullptr);
                ft = w.tmp[0];
                g.sync();
                const double mvt = merit_violation(w.cet, w.cit);
                if (ft + mvt <= phi0 + 1e-4 * t * dphi) { ls_ok = true; break; }
                t *= 0.5;
            }
            if (!ls_ok) {
                // no decrease of the merit along d at any step length: restart the quasi-Newton matrix once (as SLSQP
                // does); failing again straight after the restart is the finite-difference noise floor.
                // First suspect: an inexact QP step (the ADMM stopped at its cap far from the QP's solution).  Re-solve this
                // subproblem -- and every later one -- with a 5x cap before touching B (unicycle Tph=30: 2 of 64 cold starts need
                // 650-1000 ADMM iterations in their first SQP iterations; a 1000 cap for everybody doubles the UGV solve time).
                if (qp_cap == a.max_qp) { qp_cap = 5 * a.max_qp; continue; }
                if (just_reset || resets >= 5) { status = v0 < 1e-8 ? 0 : 1; ++k; break; }
                for (int e = g.tid; e < n * ld; e += NT) { int i = e / ld, j = e - i * ld; w.B[e] = (i == j) ? 1.0 : 0.0; }
                g.sync();
                ++resets; just_reset = true;
                continue;
            }
            just_reset = false;
            // s = t d ; Lagrangian gradient at the old point with the new multipliers
            for (int i = g.tid; i < n; i += NT) w.sv[i] = t * w.d[i];
            for (int j = g.tid; j < n; j += NT) w.glo[j] = w.g[j] + nl_col_dot(w, j, w.yq);
            g.sync();
            for (int i = g.tid; i < n; i += NT) w.z[i] += w.sv[i];
            g.sync();
            nl_eval_instance<S>(g, ph, ch, w.z, x0, p, w.X, w.U, w.tmp, w.g2, w.ce, w.Je, w.ci, w.Ji, ld, a.sx, a.su, w.ci + mii, w.Ji + (size_t)mii * ld);
            fval = w.tmp[0];
            g.sync();
            // damped BFGS:  yk = gl_new - gl_old,  Bs = B s
            double sBs = 0, sy = 0;
            for (int j = g.tid; j < n; j += NT) {
                w.rhs[j] = (w.g2[j] + nl_col_dot(w, j, w.yq)) - w.glo[j];          // yk
                double bs = 0;
#pragma unroll 4
                for (int q = 0; q < n; ++q) bs = fma(w.B[(size_t)q * ld + j], w.sv[q], bs);      // symmetric B, column access
                w.xt[j] = bs;                                  // Bs
                sBs += w.sv[j] * bs; sy += w.sv[j] * w.rhs[j];
            }
            sBs = g.sum(sBs); sy = g.sum(sy);
            g.sync();
            if (sBs > 1e-300) {
                double theta = (sy >= 0.2 * sBs) ? 1.0 : 0.8 * sBs / (sBs - sy);
                double sr = 0;
                for (int j = g.tid; j < n; j += NT) { double r = theta * w.rhs[j] + (1 - theta) * w.xt[j]; w.rhs[j] = r; sr += w.sv[j] * r; }
                sr = g.sum(sr);
                g.sync();
                for (int e = g.tid; e < n * n; e += NT) {
                    int i = e / n, j = e - i * n;
                    w.B[(size_t)i * ld + j] += -w.xt[i] * w.xt[j] / sBs + w.rhs[i] * w.rhs[j] / sr;
                }
            }
            for (int i = g.tid; i < n; i += NT) w.g[i] = w.g2[i];
            g.sync();
            double step = 0, zmax = 0;
            // the full QP step d is small only at a KKT point (t*d can be small far from one)
            for (int i = g.tid; i < n; i += NT) { step = fmax(step, fabs(w.d[i])); zmax = fmax(zmax, fabs(w.z[i])); }
            step = g.max(step); zmax = g.max(zmax);
            double v1 = violation(w.ce, w.ci);
            if (step < a.tol * fmax(1.0, zmax) && v1 < 1e-8) { status = 0; ++k; break; }
        }
        double vf = violation(w.ce, w.ci);
        for (int i = g.tid; i < n; i += NT) a.z_out[(size_t)inst * n + i] = w.z[i];
        if (g.tid == 0) {
            a.cost[inst] = fval; a.viol[inst] = vf; a.status[inst] = status; a.iters[inst] = k; a.qp_iters[inst] = qp_total;
        }
        g.sync();
    }
}

// ---- the stage-structured variant (nlmpc_structured.cuh): one CTA of NT threads per controller, everything in shared memory ----
template <class S, int NT>
__global__ void __launch_bounds__(NT) nlmpc_structured_kernel(const NlSolveArgs a) {
    extern __shared__ __align__(16) double nls_smem2[];
    constexpr int nx = S::nx, nu = S::nu, K = NlIneqPerStage<S>::value > 0 ? NlIneqPerStage<S>::value : 1;
    NlSW<nx, nu, K> w;
    w.carve(nls_smem2, a.ph, a.ch);
    const NlGrp<NT> g{(int)threadIdx.x, (int)(threadIdx.x & 31), (int)(threadIdx.x >> 5), w.red};
    NlSParams sp;
    sp.max_sqp = a.max_sqp; sp.max_qp = a.max_qp; sp.tol = a.tol; sp.ftol = a.ftol; sp.qp_eps = a.qp_eps; sp.rho0 = a.rho0;
    sp.lb = a.lb; sp.ub = a.ub; sp.sx = a.sx; sp.su = a.su;
    const int n = w.n;
    __shared__ int next_inst;
    // controllers are drawn from a global counter: solve times spread over 15x (iteration counts), a static split would leave
    // the slowest CTA with several long ones
    for (int inst = blockIdx.x;; ) {
        if (a.counter) {
            if (threadIdx.x == 0) next_inst = atomicAdd(a.counter, 1);
            __syncthreads();
            inst = next_inst;
            __syncthreads();
        }
        if (inst >= a.batch) break;
        NlSResult r = nls_solve_instance<S>(g, w, sp, a.z0 + (size_t)inst * n, a.x0 + (size_t)inst * nx, a.params + (size_t)inst * a.param_stride,
                                            a.z_out + (size_t)inst * n);
        if (g.tid == 0) { a.cost[inst] = r.cost; a.viol[inst] = r.viol; a.status[inst] = r.status; a.iters[inst] = r.iters; a.qp_iters[inst] = r.qp_iters; }
        g.sync();
        if (!a.counter) inst += gridDim.x;
    }
#ifdef NLS_PROFILE
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const char* names[] = {"kkt par1", "kkt chain fwd", "kkt border+par2", "kkt chain bwd", "kkt scatter", "factor assemble", "factor chain", "ruiz", "admm rhs", "admm rows"};
        for (int k = 0; k < 10; ++k) printf("NLSPROF %-16s %lld\n", names[k], g_nls_prof[k]);
    }
#endif
}

}  // namespace b200mpc
