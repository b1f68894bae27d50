// nlmpc_structured.cuh -- stage-structured NLMPC solve for sm_100a (SURVEY.md K6/K7, round-2 kernel).
//
// Same algorithm as nlmpc_sqp.cuh (damped-BFGS SQP, OSQP-style ADMM QP + polish, L1 merit with per-row penalties), with every
// matrix held in the form the multiple-shooting structure gives it instead of densely -- what replaces, on the reference side,
// the dense Jacobian glue of Constraints::glueJacobian (include/mpc/NLMPC/Constraints.hpp:455-482, 844-905) and NLopt's dense
// least-squares machinery behind NLOptimizer.hpp:519 ("no exploitation of the stage structure"):
//   * J_eq row of stage i touches X_i, X_{i+1} and one control block            -> stored compact, 2 nx + nu entries per row
//   * J_in row of stage i touches X_i, one control block and the slack          -> nx + nu + 1 entries per row (the system's
//     ineq_per_stage declaration; systems without it, or with user equality constraints, use the dense kernel)
//   * the quasi-Newton matrix B is block diagonal over the stages (block s = [control block s ; X_{s+1}], the slack alone): the
//     damped BFGS update is applied block by block (tests/nlmpc_sqp_reference.py: stage_groups -- it reaches the same optimum in
//     FEWER major iterations than the dense update, profiles/r01_block_bfgs_experiment.txt)
//   * the reduced KKT matrix H = c D B D + sigma I + (E A D)' R (E A D) is then block tridiagonal over the stages plus a border
//     of nu + 1 columns (last control block -- which absorbs move blocking -- and the slack):
//     tests/nlmpc_structured_kkt_reference.py.  It is factorised by a block Cholesky whose diagonal blocks are inverted
//     explicitly (solves are mat-vecs) with the border carried along as the last block row:  O(ph (nx+nu)^3) instead of O(nz^3),
//     and a solve costs O(ph (nx+nu)^2) instead of O(nz^2).
// Everything of one controller -- vectors, compact Jacobians, B blocks, the factor -- lives in shared memory (unicycle Tph=30:
// 75 KB instead of the dense kernel's 184 KB + an L2-resident matrix workspace), so several controllers share an SM.
//
// The routines are written as cooperative loops over a thread group (`for (i = g.tid; i < N; i += G::nt) ... g.sync()`), the
// sequential block recurrences by the group's first warp.  With a group of one thread the same source runs on the host:
// tests/cpp/nl_structured_host.cpp + tests/test_nlmpc_structured_host.py check it against the Python specification on the CPU.
#pragma once
#include "nlmpc_kernels.cuh"

namespace b200mpc {

#ifdef NLS_PROFILE
__device__ long long g_nls_prof[16];
#define NLS_T0() nls_t0_ = clock64()
#define NLS_DECL() long long nls_t0_ = 0; (void)nls_t0_
#define NLS_ACC(slot) do { long long t_ = clock64(); if (threadIdx.x == 0 && blockIdx.x == 0) g_nls_prof[slot] += t_ - nls_t0_; nls_t0_ = t_; } while (0)
#else
#define NLS_T0() do {} while (0)
#define NLS_DECL() do {} while (0)
#define NLS_ACC(slot) do {} while (0)
#endif
__device__ __forceinline__ double nls_lim(double v) { v = v < 1e-4 ? 1.0 : v; return v > 1e4 ? 1e4 : v; }

struct NlSParams {
    int max_sqp, max_qp;
    double tol, ftol, qp_eps, rho0;
    const double* lb; const double* ub;      // [n]
    const double* sx; const double* su;      // state / input scaling or null
};
struct NlSResult { double cost, viol; int status, iters, qp_iters; };

// Shared-memory footprint of one controller, in doubles.
__host__ __device__ inline size_t nls_doubles(int ph, int ch, int nx, int nu, int K) {
    const int b = nx + nu, nb = nu + 1, n = ph * nx + ch * nu + 1, me = ph * nx, mi = (ph + 1) * K, m = me + mi + n;
    size_t v = 13 * (size_t)n + 10 * (size_t)m + 2 * (size_t)(me + mi) + (size_t)(me + mi) + (size_t)(ph + 1) * (nx + nu) + 16;
    size_t mats = (size_t)me * (2 * nx + nu) + (size_t)mi * (nx + nu + 1) + (size_t)ph * b * b + 2 + 2 * (size_t)ph * b * b +
                  (size_t)ph * nb * b + (size_t)nb * nb + (size_t)ph * b + 4 * (size_t)(ph + 1) + 2 * (size_t)nb + 8 +
                  (size_t)ph * b * b + (size_t)ph * b * nb + (size_t)ph * b + (size_t)nb * ph;
    return (v + mats + 1) & ~(size_t)1;
}

// Per-controller workspace (all shared memory) + the index algebra of the stage partition.  nx, nu and K (inequality rows per
// stage, ineq_per_stage) are compile-time: every inner loop has a constant trip count and the index divisions strength-reduce.
template <int NX, int NU, int KK>
struct NlSW {
    static constexpr int nx = NX, nu = NU, K = KK;
    static constexpr int b = NX + NU, nb = NU + 1, we = 2 * NX + NU, wi = NX + NU + 1;
    int ph, ch;
    int n, me, mi, mc, m;
    // vectors
    double *z, *g, *g2, *d, *xs, *xt, *D, *gs, *rhs, *tmp, *glo, *sv, *zt2;     // n
    double *E, *ls, *us, *zs, *ys, *rho, *yq, *w, *pr, *pt;                     // m
    double *ce, *ci, *cet, *cit, *muv;                                          // me, mi, me, mi, mc
    double *X, *U, *red;
    // matrices
    double *JeC, *JiC;                       // me x we, mi x wi
    double *Bb, *Bsl;                        // ph blocks b x b (local order [U ; X], U slots live iff s < ch) ; slack entry
    double *Li, *Ls, *Wb, *LSi;              // factor: ph x (b x b) inverse diagonal blocks, ph x (b x b) sub-diagonal blocks,
                                             // ph x (nb x b) border rows, nb x nb inverse of the border's Cholesky factor
    double *Nb, *Vb;                         // solve-time products: N_s = Li_s' Lsub_s' (b x b), V_s = Li_s' Wb_s' (b x nb); after the
                                             // factorisation Ls[s-1] holds M_s = Li_s Lsub_{s-1} (Lsub itself is no longer needed)
    double rho_keep;                         // ADMM penalty carried from one QP subproblem to the next (as OSQP does between re-solves)
    double *cy, *cq, *bp, *cs;               // chain vectors ph x b (y, then x), ph x b (gathered rhs / q), border partials nb x ph,
                                             // per-stage scalars 4 x (ph + 1)
    __device__ void carve(double* p, int ph_, int ch_) {
        ph = ph_; ch = ch_;
        n = ph * nx + ch * nu + 1; me = ph * nx; mi = (ph + 1) * K; mc = me + mi; m = mc + n;
        double** nv[] = {&z, &g, &g2, &d, &xs, &xt, &D, &gs, &rhs, &tmp, &glo, &sv, &zt2};
        for (auto q : nv) { *q = p; p += n; }
        double** mv[] = {&E, &ls, &us, &zs, &ys, &rho, &yq, &w, &pr, &pt};
        for (auto q : mv) { *q = p; p += m; }
        ce = p; p += me; ci = p; p += mi; cet = p; p += me; cit = p; p += mi; muv = p; p += mc;
        X = p; p += (ph + 1) * nx; U = p; p += (ph + 1) * nu; red = p; p += 16;
        JeC = p; p += (size_t)me * we; JiC = p; p += (size_t)mi * wi;
        Bb = p; p += (size_t)ph * b * b; Bsl = p; p += 2;
        Li = p; p += (size_t)ph * b * b; Ls = p; p += (size_t)ph * b * b; Wb = p; p += (size_t)ph * nb * b; LSi = p; p += nb * nb;
        cy = p; p += (size_t)ph * b; cs = p; p += 4 * (size_t)(ph + 1) + 2 * nb + 8;
        Nb = p; p += (size_t)ph * b * b; Vb = p; p += (size_t)ph * b * nb; cq = p; p += (size_t)ph * b; bp = p; p += (size_t)nb * ph;
    }
    // ---- index algebra --------------------------------------------------------------------------------------------------
    __device__ __forceinline__ int blk(int i) const { int st = i < ph ? i : ph - 1; return st < ch ? st : ch - 1; }   // control block stage i reads
    // KKT partition: group s = [U block s (live iff s < ch-1) ; X block s]; border = [U block ch-1 ; slack].  -1 = dead slot.
    __device__ __forceinline__ int gz(int s, int l) const { return l < nu ? (s < ch - 1 ? ph * nx + s * nu + l : -1) : s * nx + (l - nu); }
    __device__ __forceinline__ int bz(int l) const { return l < nu ? ph * nx + (ch - 1) * nu + l : n - 1; }
    // BFGS partition (tests/nlmpc_sqp_reference.py: stage_groups): block s = [U block s (live iff s < ch) ; X block s]; slack alone
    __device__ __forceinline__ int qz(int s, int l) const { return l < nu ? (s < ch ? ph * nx + s * nu + l : -1) : s * nx + (l - nu); }
    // coefficient of variable j in dynamics row r / inequality row r (0 when the row does not touch it)
    __device__ __forceinline__ double je_at(int r, int j) const {
        const int i = r / nx;
        if (j < ph * nx) { int q = j - (i - 1) * nx; return (q >= 0 && q < 2 * nx && (i > 0 || q >= nx)) ? JeC[(size_t)r * we + q] : 0.0; }
        if (j < n - 1) { int c = (j - ph * nx) / nu; return c == blk(i) ? JeC[(size_t)r * we + 2 * nx + (j - ph * nx) - c * nu] : 0.0; }
        return 0.0;
    }
    __device__ __forceinline__ double ji_at(int r, int j) const {
        const int i = r / K;
        if (j < ph * nx) { int q = j - (i - 1) * nx; return (i > 0 && q >= 0 && q < nx) ? JiC[(size_t)r * wi + q] : 0.0; }
        if (j < n - 1) { int c = (j - ph * nx) / nu; return c == blk(i) ? JiC[(size_t)r * wi + nx + (j - ph * nx) - c * nu] : 0.0; }
        return JiC[(size_t)r * wi + nx + nu];
    }
    // B entry by variable indices (0 outside the block-diagonal pattern)
    __device__ __forceinline__ double b_at(int j1, int j2) const {
        if (j1 == n - 1 || j2 == n - 1) return (j1 == j2) ? Bsl[0] : 0.0;
        int s1, l1, s2, l2;
        if (j1 < ph * nx) { s1 = j1 / nx; l1 = nu + j1 - s1 * nx; } else { s1 = (j1 - ph * nx) / nu; l1 = (j1 - ph * nx) - s1 * nu; }
        if (j2 < ph * nx) { s2 = j2 / nx; l2 = nu + j2 - s2 * nx; } else { s2 = (j2 - ph * nx) / nu; l2 = (j2 - ph * nx) - s2 * nu; }
        return s1 == s2 ? Bb[((size_t)s1 * b + l1) * b + l2] : 0.0;
    }
    // sum_r A[r][j] v[r] over the dynamics and inequality rows (the bound rows are the caller's)
    __device__ __forceinline__ double col_dot(int j, const double* v) const {
        double a = 0;
        if (j < ph * nx) {
            const int s = j / nx, q = j - s * nx;
            for (int r = s * nx; r < (s + 1) * nx; ++r) a = fma(JeC[(size_t)r * we + nx + q], v[r], a);
            if (s + 1 < ph) for (int r = (s + 1) * nx; r < (s + 2) * nx; ++r) a = fma(JeC[(size_t)r * we + q], v[r], a);
            for (int r = (s + 1) * K; r < (s + 2) * K; ++r) a = fma(JiC[(size_t)r * wi + q], v[me + r], a);
        } else if (j < n - 1) {
            const int c = (j - ph * nx) / nu, q = (j - ph * nx) - c * nu;
            const int i1 = c < ch - 1 ? c + 1 : ph, j1 = c < ch - 1 ? c + 1 : ph + 1;
            for (int r = c * nx; r < i1 * nx; ++r) a = fma(JeC[(size_t)r * we + 2 * nx + q], v[r], a);
            for (int r = c * K; r < j1 * K; ++r) a = fma(JiC[(size_t)r * wi + nx + q], v[me + r], a);
        } else {
            for (int r = 0; r < mi; ++r) a = fma(JiC[(size_t)r * wi + nx + nu], v[me + r], a);
        }
        return a;
    }
    // max_r e[r] |A[r][j]| over the dynamics and inequality rows
    __device__ __forceinline__ double col_max(int j, const double* e) const {
        double a = 0;
        if (j < ph * nx) {
            const int s = j / nx, q = j - s * nx;
            for (int r = s * nx; r < (s + 1) * nx; ++r) a = fmax(a, e[r] * fabs(JeC[(size_t)r * we + nx + q]));
            if (s + 1 < ph) for (int r = (s + 1) * nx; r < (s + 2) * nx; ++r) a = fmax(a, e[r] * fabs(JeC[(size_t)r * we + q]));
            for (int r = (s + 1) * K; r < (s + 2) * K; ++r) a = fmax(a, e[me + r] * fabs(JiC[(size_t)r * wi + q]));
        } else if (j < n - 1) {
            const int c = (j - ph * nx) / nu, q = (j - ph * nx) - c * nu;
            const int i1 = c < ch - 1 ? c + 1 : ph, j1 = c < ch - 1 ? c + 1 : ph + 1;
            for (int r = c * nx; r < i1 * nx; ++r) a = fmax(a, e[r] * fabs(JeC[(size_t)r * we + 2 * nx + q]));
            for (int r = c * K; r < j1 * K; ++r) a = fmax(a, e[me + r] * fabs(JiC[(size_t)r * wi + nx + q]));
        } else {
            for (int r = 0; r < mi; ++r) a = fmax(a, e[me + r] * fabs(JiC[(size_t)r * wi + nx + nu]));
        }
        return a;
    }
    // (A x)_r for a dynamics / inequality row
    __device__ __forceinline__ double je_row_dot(int r, const double* x) const {
        const int i = r / nx;
        const double* a = JeC + (size_t)r * we;
        double s = 0;
        if (i > 0) for (int q = 0; q < nx; ++q) s = fma(a[q], x[(i - 1) * nx + q], s);
        for (int q = 0; q < nx; ++q) s = fma(a[nx + q], x[i * nx + q], s);
        const int u0 = ph * nx + blk(i) * nu;
        for (int q = 0; q < nu; ++q) s = fma(a[2 * nx + q], x[u0 + q], s);
        return s;
    }
    __device__ __forceinline__ double ji_row_dot(int r, const double* x) const {
        const int i = r / K;
        const double* a = JiC + (size_t)r * wi;
        double s = a[nx + nu] * x[n - 1];
        if (i > 0) for (int q = 0; q < nx; ++q) s = fma(a[q], x[(i - 1) * nx + q], s);
        const int u0 = ph * nx + blk(i) * nu;
        for (int q = 0; q < nu; ++q) s = fma(a[nx + q], x[u0 + q], s);
        return s;
    }
    __device__ __forceinline__ double je_row_max(int r, const double* dsc) const {
        const int i = r / nx;
        const double* a = JeC + (size_t)r * we;
        double s = 0;
        if (i > 0) for (int q = 0; q < nx; ++q) s = fmax(s, fabs(a[q]) * dsc[(i - 1) * nx + q]);
        for (int q = 0; q < nx; ++q) s = fmax(s, fabs(a[nx + q]) * dsc[i * nx + q]);
        const int u0 = ph * nx + blk(i) * nu;
        for (int q = 0; q < nu; ++q) s = fmax(s, fabs(a[2 * nx + q]) * dsc[u0 + q]);
        return s;
    }
    __device__ __forceinline__ double ji_row_max(int r, const double* dsc) const {
        const int i = r / K;
        const double* a = JiC + (size_t)r * wi;
        double s = fabs(a[nx + nu]) * dsc[n - 1];
        if (i > 0) for (int q = 0; q < nx; ++q) s = fmax(s, fabs(a[q]) * dsc[(i - 1) * nx + q]);
        const int u0 = ph * nx + blk(i) * nu;
        for (int q = 0; q < nu; ++q) s = fmax(s, fabs(a[nx + q]) * dsc[u0 + q]);
        return s;
    }
    // (B x)_j and max_i dsc_i |B_ij| for variable j
    __device__ __forceinline__ double b_row_dot(int j, const double* x) const {
        if (j == n - 1) return Bsl[0] * x[j];
        int s, l;
        if (j < ph * nx) { s = j / nx; l = nu + j - s * nx; } else { s = (j - ph * nx) / nu; l = (j - ph * nx) - s * nu; }
        const double* row = Bb + ((size_t)s * b + l) * b;
        double a = 0;
        for (int k = 0; k < b; ++k) { int jz = qz(s, k); if (jz >= 0) a = fma(row[k], x[jz], a); }
        return a;
    }
    __device__ __forceinline__ double b_col_max(int j, const double* dsc) const {
        if (j == n - 1) return fabs(Bsl[0]) * dsc[j];
        int s, l;
        if (j < ph * nx) { s = j / nx; l = nu + j - s * nx; } else { s = (j - ph * nx) / nu; l = (j - ph * nx) - s * nu; }
        const double* row = Bb + ((size_t)s * b + l) * b;                    // symmetric: row = column
        double a = 0;
        for (int k = 0; k < b; ++k) { int jz = qz(s, k); if (jz >= 0) a = fmax(a, fabs(row[k]) * dsc[jz]); }
        return a;
    }
};

// out_r = E_r (A dx)_r for all m rows (dx = D .* x already formed)
template <class G, class WS>
__device__ __forceinline__ void nls_As_core(const G& g, WS& w, const double* dx, double* out) {
    for (int r = g.tid; r < w.me; r += G::nt) out[r] = w.E[r] * w.je_row_dot(r, dx);
    for (int r = g.tid; r < w.mi; r += G::nt) out[w.me + r] = w.E[w.me + r] * w.ji_row_dot(r, dx);
    for (int j = g.tid; j < w.n; j += G::nt) out[w.mc + j] = w.E[w.mc + j] * dx[j];
    g.sync();
}
template <class G, class WS>
__device__ __forceinline__ void nls_As(const G& g, WS& w, const double* x, double* out) {
    for (int i = g.tid; i < w.n; i += G::nt) w.tmp[i] = w.D[i] * x[i];
    g.sync();
    nls_As_core(g, w, w.tmp, out);
}
// out_j = D_j (A' (E .* v))_j
template <class G, class WS>
__device__ __forceinline__ void nls_Ats(const G& g, WS& w, const double* v, double* out) {
    for (int r = g.tid; r < w.m; r += G::nt) w.w[r] = w.E[r] * v[r];
    g.sync();
    for (int j = g.tid; j < w.n; j += G::nt) out[j] = w.D[j] * (w.w[w.mc + j] + w.col_dot(j, w.w));
    g.sync();
}
// out = c D B D x
template <class G, class WS>
__device__ __forceinline__ void nls_Ps(const G& g, WS& w, double c, const double* x, double* out) {
    for (int i = g.tid; i < w.n; i += G::nt) w.tmp[i] = w.D[i] * x[i];
    g.sync();
    for (int i = g.tid; i < w.n; i += G::nt) out[i] = c * w.D[i] * w.b_row_dot(i, w.tmp);
    g.sync();
}

// ---- H = c D B D + sigma I + (E A D)' diag(rho) (E A D), bordered block tridiagonal; factorisation ------------------------------
// Blocks are assembled straight into the factor storage (Li <- diagonal blocks, Ls[s] <- H[g_{s+1}, g_s], Wb[s] <- H[border, g_s]')
// and factorised in place by the group's first warp.  Dead slots (U slots of groups >= ch-1) are unit rows / columns.
template <class G, class WS>
__device__ __forceinline__ bool nls_factor(const G& g, WS& w, double c, double sigma) {
    constexpr int b = WS::b, nb = WS::nb, nx = WS::nx, K = WS::K;
    const int ph = w.ph, me = w.me, mc = w.mc;
    NLS_DECL(); NLS_T0();
    for (int r = g.tid; r < mc; r += G::nt) w.w[r] = w.rho[r] * w.E[r] * w.E[r];
    g.sync();
    // entry (ja, jb) of D (c B + A' W A) D from the candidate rows [e0,e1) of J_eq and [i0,i1) of J_in
    auto entry = [&](int ja, int jb, int e0, int e1, int i0, int i1) {
        double acc = c * w.b_at(ja, jb);
        for (int r = e0; r < e1; ++r) { double a = w.je_at(r, ja); if (a != 0.0) acc = fma(w.w[r] * a, w.je_at(r, jb), acc); }
        for (int r = i0; r < i1; ++r) { double a = w.ji_at(r, ja); if (a != 0.0) acc = fma(w.w[me + r] * a, w.ji_at(r, jb), acc); }
        return w.D[ja] * acc * w.D[jb];
    };
    auto diag_extra = [&](int j) { int rb = mc + j; return sigma + w.rho[rb] * w.E[rb] * w.E[rb] * w.D[j] * w.D[j]; };
    // diagonal blocks, sub-diagonal blocks, border rows: one pass over all (s, row, col) triples
    const int per = b * b + b * b + nb * b;
    for (int t = g.tid; t < ph * per; t += G::nt) {
        const int s = t / per, e = t - s * per;
        const int e0 = s * nx, e1 = (s + 2) * nx < me ? (s + 2) * nx : me;            // dynamics stages s, s+1
        const int i0 = s * K, i1 = (s + 2) * K;                                       // inequality stages s, s+1
        if (e < b * b) {
            const int la = e / b, lb = e - la * b, ja = w.gz(s, la), jb = w.gz(s, lb);
            double v;
            if (ja < 0 || jb < 0) v = (la == lb) ? 1.0 : 0.0;
            else { v = entry(ja, jb, e0, e1, i0, i1); if (la == lb) v += diag_extra(ja); }
            w.Li[((size_t)s * b + la) * b + lb] = v;
        } else if (e < 2 * b * b) {
            const int ee = e - b * b, la = ee / b, lb = ee - la * b;                  // row of group s+1, column of group s
            double v = 0.0;
            if (s + 1 < ph) {
                const int ja = w.gz(s + 1, la), jb = w.gz(s, lb);
                if (ja >= 0 && jb >= 0) v = entry(ja, jb, (s + 1) * nx, (s + 2) * nx, (s + 1) * K, (s + 2) * K);
            }
            w.Ls[((size_t)s * b + la) * b + lb] = v;
        } else {
            const int ee = e - 2 * b * b, la = ee / b, lb = ee - la * b;              // border row la, column lb of group s
            const int ja = w.bz(la), jb = w.gz(s, lb);
            w.Wb[((size_t)s * nb + la) * b + lb] = jb < 0 ? 0.0 : entry(ja, jb, e0, e1, i0, i1);
        }
    }
    // border block (nb x nb): every dynamics row from stage ch-1 on, every inequality row
    for (int e = g.tid; e < nb * nb; e += G::nt) {
        const int la = e / nb, lb = e - la * nb, ja = w.bz(la), jb = w.bz(lb);
        double v = entry(ja, jb, (w.ch - 1) * nx, me, 0, w.mi);
        if (la == lb) v += diag_extra(ja);
        w.LSi[e] = v;
    }
    g.sync();
    NLS_ACC(5);
    // ---- block Cholesky with the border as the last block row ------------------------------------------------------------------
    // Sequential part (first warp): A_s = H_ss - Lsub_{s-1} Lsub_{s-1}', its Cholesky factor, Li_s = L_s^-1, Lsub_s = H_{s+1,s} Li_s';
    // then the border recurrence Wb_s = (H_{b,s} - Wb_{s-1} Lsub_{s-1}') Li_s'.  Everything that is not a recurrence -- the solve-time
    // products N_s, V_s, M_s and the border's Schur complement -- is done afterwards by the whole group, all stages at once.
    bool ok = true;
    double* Sc = w.pt;                              // scratch for one block (m >= b*b: nls_supported)
    double* ipiv = w.cs;                            // reciprocal pivots of the current block (b <= 4 (ph+1) + 8: nls_supported)
    if (g.wid == 0) {
        const int lane = g.lane;
        constexpr int W = G::nt < 32 ? G::nt : 32;
        auto wsync = [&]() {
#ifndef B200_HOST_EMU
            __syncwarp();
#endif
        };
        for (int s = 0; s < ph; ++s) {
            double* A = w.Li + (size_t)s * b * b;
            if (s > 0) {
                const double* Lp = w.Ls + (size_t)(s - 1) * b * b;
                for (int e = lane; e < b * b; e += W) {
                    const int i = e / b, j = e - i * b;
                    if (j <= i) { double acc = 0; for (int q = 0; q < b; ++q) acc = fma(Lp[i * b + q], Lp[j * b + q], acc); A[e] -= acc; }
                }
                wsync();
            }
            // left-looking Cholesky, one column per step: lane r-k updates entry (r, k), the pivot comes from the first lane
            for (int k = 0; k < b; ++k) {
#ifndef B200_HOST_EMU
                const int r = k + lane;
                double v = 0;
                if (r < b) { v = A[r * b + k]; for (int q = 0; q < k; ++q) v = fma(-A[r * b + q], A[k * b + q], v); }
                const double dkk = __shfl_sync(0xffffffffu, v, 0);
                if (!(dkk > 0.0)) ok = false;
                const double piv = sqrt(dkk), inv = 1.0 / piv;
                if (r < b) A[r * b + k] = (r == k) ? piv : v * inv;
                if (lane == 0) ipiv[k] = inv;
                __syncwarp();
#else
                double dkk = A[k * b + k];
                for (int q = 0; q < k; ++q) dkk = fma(-A[k * b + q], A[k * b + q], dkk);
                if (!(dkk > 0.0)) ok = false;
                const double piv = sqrt(dkk), inv = 1.0 / piv;
                for (int r = k + 1; r < b; ++r) { double v = A[r * b + k]; for (int q = 0; q < k; ++q) v = fma(-A[r * b + q], A[k * b + q], v); A[r * b + k] = v * inv; }
                A[k * b + k] = piv; ipiv[k] = inv;
#endif
            }
            // explicit inverse of L (lower) into Sc, column by column: lane = column
            for (int col = lane; col < b; col += W) {
                for (int r = 0; r < b; ++r) {
                    double v;
                    if (r < col) v = 0.0;
                    else if (r == col) v = ipiv[r];
                    else { double acc = 0; for (int q = col; q < r; ++q) acc = fma(A[r * b + q], Sc[q * b + col], acc); v = -acc * ipiv[r]; }
                    Sc[r * b + col] = v;
                }
            }
            wsync();
            for (int e = lane; e < b * b; e += W) A[e] = Sc[e];                    // Li[s] = L^-1 (full square, upper part zero)
            wsync();
            // Lsub[s] = sub[s] Li[s]'  in place: lane = row, columns from the right (entry j needs the old entries q <= j only)
            if (s + 1 < ph) {
                double* Lsb = w.Ls + (size_t)s * b * b;
                for (int i = lane; i < b; i += W)
                    for (int j = b - 1; j >= 0; --j) { double acc = 0; for (int q = 0; q <= j; ++q) acc = fma(Lsb[i * b + q], A[j * b + q], acc); Lsb[i * b + j] = acc; }
                wsync();
            }
        }
        // border recurrence:  Wb[s] = (Wb[s] - Wb[s-1] Lsub[s-1]') Li[s]'   (lane = one of the nb x b entries; two steps per stage)
        for (int s = 0; s < ph; ++s) {
            const double* A = w.Li + (size_t)s * b * b;
            double* Wc = w.Wb + (size_t)s * nb * b;
            if (s > 0) {
                const double* Lp = w.Ls + (size_t)(s - 1) * b * b;
                const double* Wp = w.Wb + (size_t)(s - 1) * nb * b;
                for (int e = lane; e < nb * b; e += W) {
                    const int i = e / b, j = e - i * b;
                    double acc = Wc[e]; for (int q = 0; q < b; ++q) acc = fma(-Wp[i * b + q], Lp[j * b + q], acc);
                    Wc[e] = acc;
                }
                wsync();
            }
            for (int i = lane; i < nb; i += W)
                for (int j = b - 1; j >= 0; --j) { double acc = 0; for (int q = 0; q <= j; ++q) acc = fma(Wc[i * b + q], A[j * b + q], acc); Wc[i * b + j] = acc; }
            wsync();
        }
    }
    g.sync();
    // solve-time products, all stages in parallel:  V_s = Li_s' Wb_s'  (b x nb),  N_s = Li_s' Lsub_s'  (b x b)
    for (int t = g.tid; t < ph * b * (nb + b); t += G::nt) {
        const int s = t / (b * (nb + b)), e = t - s * b * (nb + b);
        const double* A = w.Li + (size_t)s * b * b;
        if (e < b * nb) {
            const int l = e / nb, q = e - l * nb;
            const double* Wc = w.Wb + (size_t)s * nb * b;
            double acc = 0; for (int k = l; k < b; ++k) acc = fma(A[k * b + l], Wc[q * b + k], acc);
            w.Vb[(size_t)s * b * nb + e] = acc;
        } else if (s + 1 < ph) {
            const int ee = e - b * nb, l = ee / b, q = ee - l * b;
            const double* Lsb = w.Ls + (size_t)s * b * b;
            double acc = 0; for (int k = l; k < b; ++k) acc = fma(A[k * b + l], Lsb[q * b + k], acc);
            w.Nb[(size_t)s * b * b + ee] = acc;
        }
    }
    // border Schur complement  S = H_bb - sum_s Wb[s] Wb[s]'   (nb x nb entries, one thread each)
    for (int e = g.tid; e < nb * nb; e += G::nt) {
        const int i = e / nb, j = e - i * nb;
        double acc = 0;
        for (int s = 0; s < ph; ++s) { const double* Wc = w.Wb + (size_t)s * nb * b; for (int q = 0; q < b; ++q) acc = fma(Wc[i * b + q], Wc[j * b + q], acc); }
        Sc[e] = w.LSi[e] - acc;
    }
    g.sync();
    // M_s = Li_s Lsub[s-1] in place of Lsub[s-1] (one thread per column: the whole column is read before it is written)
    for (int t = g.tid; t < (ph - 1) * b; t += G::nt) {
        const int s = 1 + t / b, q = t - (s - 1) * b;
        const double* A = w.Li + (size_t)s * b * b;
        double* Lp = w.Ls + (size_t)(s - 1) * b * b;
        double col[b];
#pragma unroll
        for (int k = 0; k < b; ++k) col[k] = Lp[k * b + q];
#pragma unroll
        for (int l = 0; l < b; ++l) { double acc = 0; for (int k = 0; k <= l; ++k) acc = fma(A[l * b + k], col[k], acc); Lp[l * b + q] = acc; }
    }
    if (g.tid == 0) {                                                           // nb <= ~7: Cholesky of S and its inverse by one thread
        double* Sb = w.LSi;
        for (int k = 0; k < nb; ++k) {
            double dkk = Sc[k * nb + k];
            for (int q = 0; q < k; ++q) dkk -= Sc[k * nb + q] * Sc[k * nb + q];
            if (!(dkk > 0.0)) ok = false;
            const double piv = sqrt(dkk);
            Sc[k * nb + k] = piv;
            for (int r = k + 1; r < nb; ++r) {
                double v = Sc[r * nb + k];
                for (int q = 0; q < k; ++q) v -= Sc[r * nb + q] * Sc[k * nb + q];
                Sc[r * nb + k] = v / piv;
            }
        }
        for (int col = 0; col < nb; ++col)
            for (int r = 0; r < nb; ++r) {
                double v;
                if (r < col) v = 0.0;
                else if (r == col) v = 1.0 / Sc[r * nb + r];
                else { double acc = 0; for (int q = col; q < r; ++q) acc += Sc[r * nb + q] * Sb[q * nb + col]; v = -acc / Sc[r * nb + r]; }
                Sb[r * nb + col] = v;
            }
    }
    g.sync();
    NLS_ACC(6);
    return !g.any(!ok);
}

// xt = H^-1 rhs through the factor; optionally dxt = D .* xt.
//   forward   y_s = Li_s r_s - M_s y_{s-1}            (Li_s r_s for all stages in parallel first; one mat-vec per stage on the chain)
//   border    y_b = LSi (r_b - sum_s Wb_s y_s),  x_b = LSi' y_b
//   backward  x_s = (Li_s' y_s - V_s x_b) - N_s x_{s+1}   (the bracket in parallel; one mat-vec per stage on the chain)
// The two recurrences run on the group's first warp, lane = row, one warp barrier per stage; everything else is group-parallel.
template <int BS, class G, class WS>
__device__ __forceinline__ void nls_kkt_apply(const G& g, WS& w, double* dxt = nullptr) {
    const int ph = w.ph;
    constexpr int b = BS, nb = WS::nb;
    constexpr int W = G::nt < 32 ? G::nt : 32;
    auto wsync = [&]() {
#ifndef B200_HOST_EMU
        __syncwarp();
#endif
    };
    NLS_DECL(); NLS_T0();
    for (int e = g.tid; e < ph * b; e += G::nt) { const int s = e / b, l = e - s * b, jz = w.gz(s, l); w.cq[e] = jz >= 0 ? w.rhs[jz] : 0.0; }
    g.sync();
    for (int e = g.tid; e < ph * b; e += G::nt) {
        const int s = e / b, l = e - s * b;
        const double* Lr = w.Li + (size_t)e * b;
        const double* r = w.cq + s * b;
        double v = 0;
#pragma unroll
        for (int q = 0; q < b; ++q) if (q <= l) v = fma(Lr[q], r[q], v);
        w.cy[e] = v;
    }
    g.sync();
    NLS_ACC(0);
    if (g.wid == 0) {
#ifndef B200_HOST_EMU
        // the running vector stays in registers (lane = row) and is broadcast with shuffles: no shared-memory round trip and no
        // barrier on the dependent chain; the matrix row and the parallel part of the next stage are loaded one stage ahead
        const int l = g.lane < b ? g.lane : b - 1;
        double y = w.cy[l];
        double mrow[b], pn = 0;
        if (ph > 1) {
            const double* M = w.Ls + (size_t)l * b;
#pragma unroll
            for (int q = 0; q < b; ++q) mrow[q] = M[q];
            pn = w.cy[b + l];
        }
        for (int s = 1; s < ph; ++s) {
            double a0 = pn, a1 = 0;
#pragma unroll
            for (int q = 0; q < b; ++q) {
                const double yq = __shfl_sync(0xffffffffu, y, q);
                if (q & 1) a1 = fma(-mrow[q], yq, a1); else a0 = fma(-mrow[q], yq, a0);
            }
            if (s + 1 < ph) {
                const double* M = w.Ls + ((size_t)s * b + l) * b;
#pragma unroll
                for (int q = 0; q < b; ++q) mrow[q] = M[q];
                pn = w.cy[(s + 1) * b + l];
            }
            y = a0 + a1;
            if (g.lane < b) w.cy[s * b + l] = y;
        }
        __syncwarp();
#else
        for (int s = 1; s < ph; ++s) {
            for (int l = g.lane; l < b; l += W) {
                const double* M = w.Ls + ((size_t)(s - 1) * b + l) * b;
                const double* yp = w.cy + (s - 1) * b;
                double v = w.cy[s * b + l];
                for (int q = 0; q < b; ++q) v = fma(-M[q], yp[q], v);
                w.cy[s * b + l] = v;
            }
            wsync();
        }
#endif
    }
    g.sync();
    NLS_ACC(1);
    for (int e = g.tid; e < nb * ph; e += G::nt) {
        const int l = e / ph, s = e - l * ph;
        const double* Wr = w.Wb + ((size_t)s * nb + l) * b;
        const double* ys = w.cy + s * b;
        double v = 0;
#pragma unroll
        for (int q = 0; q < b; ++q) v = fma(Wr[q], ys[q], v);
        w.bp[e] = v;
    }
    g.sync();
    double* yb = w.cs; double* xb = w.cs + nb; double* t = w.cs + 2 * nb;
    if (g.wid == 0) {
        for (int l = g.lane; l < nb; l += W) {
            double v0 = w.rhs[w.bz(l)], v1 = 0, v2 = 0, v3 = 0;
            const double* bpl = w.bp + l * ph;
            int s = 0;
            for (; s + 3 < ph; s += 4) { v0 -= bpl[s]; v1 -= bpl[s + 1]; v2 -= bpl[s + 2]; v3 -= bpl[s + 3]; }
            for (; s < ph; ++s) v0 -= bpl[s];
            t[l] = (v0 + v1) + (v2 + v3);
        }
        wsync();
        for (int l = g.lane; l < nb; l += W) { double v = 0; for (int q = 0; q <= l; ++q) v = fma(w.LSi[l * nb + q], t[q], v); yb[l] = v; }
        wsync();
        for (int l = g.lane; l < nb; l += W) { double v = 0; for (int q = l; q < nb; ++q) v = fma(w.LSi[q * nb + l], yb[q], v); xb[l] = v; w.xt[w.bz(l)] = v; }
        wsync();
    }
    g.sync();
    for (int e = g.tid; e < ph * b; e += G::nt) {
        const int s = e / b, l = e - s * b;
        const double* Lc = w.Li + (size_t)s * b * b;
        const double* ys = w.cy + s * b;
        double v = 0;
#pragma unroll
        for (int q = 0; q < b; ++q) if (q >= l) v = fma(Lc[q * b + l], ys[q], v);
        const double* Vr = w.Vb + (size_t)e * nb;
        for (int q = 0; q < nb; ++q) v = fma(-Vr[q], xb[q], v);
        w.cq[e] = v;
    }
    g.sync();
    NLS_ACC(2);
    if (g.wid == 0) {
#ifndef B200_HOST_EMU
        const int l = g.lane < b ? g.lane : b - 1;
        double x = w.cq[(ph - 1) * b + l];
        double nrow[b], qn = 0;
        if (ph > 1) {
            const double* N = w.Nb + ((size_t)(ph - 2) * b + l) * b;
#pragma unroll
            for (int q = 0; q < b; ++q) nrow[q] = N[q];
            qn = w.cq[(ph - 2) * b + l];
        }
        for (int s = ph - 2; s >= 0; --s) {
            double a0 = qn, a1 = 0;
#pragma unroll
            for (int q = 0; q < b; ++q) {
                const double xq = __shfl_sync(0xffffffffu, x, q);
                if (q & 1) a1 = fma(-nrow[q], xq, a1); else a0 = fma(-nrow[q], xq, a0);
            }
            if (s > 0) {
                const double* N = w.Nb + ((size_t)(s - 1) * b + l) * b;
#pragma unroll
                for (int q = 0; q < b; ++q) nrow[q] = N[q];
                qn = w.cq[(s - 1) * b + l];
            }
            x = a0 + a1;
            if (g.lane < b) w.cq[s * b + l] = x;
        }
        __syncwarp();
#else
        for (int s = ph - 2; s >= 0; --s) {
            for (int l = g.lane; l < b; l += W) {
                const double* N = w.Nb + ((size_t)s * b + l) * b;
                const double* xn = w.cq + (s + 1) * b;
                double v = w.cq[s * b + l];
                for (int q = 0; q < b; ++q) v = fma(-N[q], xn[q], v);
                w.cq[s * b + l] = v;
            }
            wsync();
        }
#endif
    }
    g.sync();
    NLS_ACC(3);
    for (int e = g.tid; e < ph * b; e += G::nt) { const int s = e / b, l = e - s * b, jz = w.gz(s, l); if (jz >= 0) w.xt[jz] = w.cq[e]; }
    g.sync();
    if (dxt) { for (int i = g.tid; i < w.n; i += G::nt) dxt[i] = w.D[i] * w.xt[i]; g.sync(); }
    NLS_ACC(4);
}

// max bound violation of A x and max |P x + q + A' y| in the scaled problem (uses pt, rhs, zt2)
template <class G, class WS>
__device__ __forceinline__ void nls_qp_residuals(const G& g, WS& w, double c, const double* x, const double* y, double& pri, double& dua) {
    nls_As(g, w, x, w.pt);
    double p = 0, d = 0;
    for (int r = g.tid; r < w.m; r += G::nt) p = fmax(p, fmax(fmax(w.ls[r] - w.pt[r], w.pt[r] - w.us[r]), 0.0));
    nls_Ats(g, w, y, w.rhs);
    nls_Ps(g, w, c, x, w.zt2);
    for (int i = g.tid; i < w.n; i += G::nt) d = fmax(d, fabs(w.zt2[i] + w.gs[i] + w.rhs[i]));
    pri = g.max(p); dua = g.max(d);
}

// OSQP polish.c on the structured QP (see nl_qp_polish_impl in nlmpc_sqp.cuh: identical steps, structured products / factor)
template <int BS, class G, class WS>
__device__ __forceinline__ bool nls_qp_polish(const G& g, WS& w, double c) {
    const int n = w.n, m = w.m;
    const double delta = 1e-6, idelta = 1e6;
    double pri_a, dua_a;
    nls_qp_residuals(g, w, c, w.xs, w.ys, pri_a, dua_a);
    for (int r = g.tid; r < m; r += G::nt) {
        bool lo = (w.zs[r] - w.ls[r]) < -w.ys[r], up = (w.us[r] - w.zs[r]) < w.ys[r];
        w.rho[r] = (lo || up) ? idelta : 0.0;
        w.zs[r] = lo ? w.ls[r] : (up ? w.us[r] : 0.0);
        w.yq[r] = 0.0;
    }
    for (int i = g.tid; i < n; i += G::nt) w.g2[i] = 0.0;
    g.sync();
    if (!nls_factor(g, w, c, delta)) return false;
    for (int it = 0; it <= 5; ++it) {
        nls_Ats(g, w, w.yq, w.rhs);
        nls_Ps(g, w, c, w.g2, w.zt2);
        for (int i = g.tid; i < n; i += G::nt) w.sv[i] = -w.gs[i] - w.zt2[i] - w.rhs[i];
        nls_As(g, w, w.g2, w.pt);
        for (int r = g.tid; r < m; r += G::nt) w.pr[r] = w.rho[r] > 0.0 ? (w.zs[r] - w.pt[r]) * idelta : 0.0;
        g.sync();
        nls_Ats(g, w, w.pr, w.rhs);
        for (int i = g.tid; i < n; i += G::nt) w.rhs[i] += w.sv[i];
        g.sync();
        nls_kkt_apply<BS>(g, w);
        nls_As(g, w, w.xt, w.pt);
        for (int r = g.tid; r < m; r += G::nt) if (w.rho[r] > 0.0) w.yq[r] += w.pt[r] * idelta - w.pr[r];
        for (int i = g.tid; i < n; i += G::nt) w.g2[i] += w.xt[i];
        g.sync();
    }
    double pri_p, dua_p;
    nls_qp_residuals(g, w, c, w.g2, w.yq, pri_p, dua_p);
    const bool ok = pri_p <= fmax(pri_a, 1e-10) && dua_p <= fmax(dua_a, 1e-10);
    if (ok) {
        for (int i = g.tid; i < n; i += G::nt) w.xs[i] = w.g2[i];
        for (int r = g.tid; r < m; r += G::nt) w.ys[r] = w.yq[r];
        g.sync();
    }
    return ok;
}

// OSQP-style ADMM for the QP subproblem on the structured storage.  In: Bb, g, JeC, JiC, ce, ci, z, lb, ub; warm dual yq (if have_y).
// Out: d (step), yq (multipliers, unscaled).  Returns ADMM iterations.  Step for step nl_qp_solve of nlmpc_sqp.cuh.
template <int BS, class G, class WS>
__device__ __forceinline__ int nls_qp_solve(const G& g, WS& w, const NlSParams& a, int mii, bool have_y, int max_qp) {
    const int n = w.n, me = w.me, mi = w.mi, mc = w.mc, m = w.m;
    const double sigma = 1e-6, alpha = 1.6;
    NLS_DECL(); NLS_T0();
    for (int i = g.tid; i < n; i += G::nt) { w.D[i] = 1.0; w.gs[i] = w.g[i]; }
    for (int r = g.tid; r < m; r += G::nt) w.E[r] = 1.0;
    double c = 1.0;
    g.sync();
    for (int pass = 0; pass < 10; ++pass) {
        for (int j = g.tid; j < n; j += G::nt) {
            double cn = c * w.D[j] * w.b_col_max(j, w.D);
            double an = fmax(w.col_max(j, w.E), w.E[mc + j]) * w.D[j];
            w.xt[j] = 1.0 / sqrt(nls_lim(fmax(cn, an)));
        }
        for (int r = g.tid; r < me; r += G::nt) w.w[r] = 1.0 / sqrt(nls_lim(w.je_row_max(r, w.D) * w.E[r]));
        for (int r = g.tid; r < mi; r += G::nt) w.w[me + r] = 1.0 / sqrt(nls_lim(w.ji_row_max(r, w.D) * w.E[me + r]));
        for (int j = g.tid; j < n; j += G::nt) w.w[mc + j] = 1.0 / sqrt(nls_lim(w.E[mc + j] * w.D[j]));
        g.sync();
        for (int j = g.tid; j < n; j += G::nt) { w.D[j] *= w.xt[j]; w.gs[j] *= w.xt[j]; }
        for (int r = g.tid; r < m; r += G::nt) w.E[r] *= w.w[r];
        g.sync();
        double psum = 0, qmax = 0;
        for (int j = g.tid; j < n; j += G::nt) { psum += c * w.D[j] * w.b_col_max(j, w.D); qmax = fmax(qmax, fabs(w.gs[j])); }
        psum = g.sum(psum); qmax = g.max(qmax);
        double ct = fmax(psum / n, nls_lim(qmax));
        ct = 1.0 / nls_lim(ct);
        c *= ct;
        for (int j = g.tid; j < n; j += G::nt) w.gs[j] *= ct;
        g.sync();
    }
    NLS_ACC(7);
    double rho0 = w.rho_keep;
    for (int r = g.tid; r < m; r += G::nt) {
        double l, u;
        if (r < me) { l = u = -w.ce[r]; }
        else if (r < mc) { u = -w.ci[r - me]; l = (r - me) < mii ? -B200_INF : u; }
        else { int j = r - mc; l = a.lb[j] - w.z[j]; u = a.ub[j] - w.z[j]; }
        w.ls[r] = w.E[r] * l; w.us[r] = w.E[r] * u;
    }
    g.sync();
    auto set_rho = [&](double r0) {
        for (int r = g.tid; r < m; r += G::nt)
            w.rho[r] = (w.ls[r] < -1e20 && w.us[r] > 1e20) ? 1e-6 : ((w.us[r] - w.ls[r]) < 1e-9) ? 1e3 * r0 : r0;
        g.sync();
    };
    set_rho(rho0);
    nls_factor(g, w, c, sigma);
    for (int i = g.tid; i < n; i += G::nt) w.xs[i] = 0.0;
    for (int r = g.tid; r < m; r += G::nt) { w.ys[r] = have_y ? c * w.yq[r] / w.E[r] : 0.0; w.zs[r] = fmin(fmax(0.0, w.ls[r]), w.us[r]); }
    g.sync();
    int it = 0;
    for (it = 1; it <= max_qp; ++it) {
        NLS_T0();
        for (int r = g.tid; r < m; r += G::nt) w.w[r] = w.E[r] * (w.rho[r] * w.zs[r] - w.ys[r]);
        g.sync();
        for (int j = g.tid; j < n; j += G::nt)
            w.rhs[j] = w.D[j] * (w.w[mc + j] + w.col_dot(j, w.w)) + sigma * w.xs[j] - w.gs[j];
        g.sync();
        NLS_ACC(8);
        nls_kkt_apply<BS>(g, w, w.zt2);
        NLS_T0();
        nls_As_core(g, w, w.zt2, w.yq);
        for (int i = g.tid; i < n; i += G::nt) w.xs[i] = alpha * w.xt[i] + (1 - alpha) * w.xs[i];
        for (int r = g.tid; r < m; r += G::nt) {
            double zr = alpha * w.yq[r] + (1 - alpha) * w.zs[r];
            double zn = fmin(fmax(zr + w.ys[r] / w.rho[r], w.ls[r]), w.us[r]);
            w.ys[r] += w.rho[r] * (zr - zn);
            w.zs[r] = zn;
        }
        g.sync();
        NLS_ACC(9);
        if (it % 25 == 0) {
            nls_As(g, w, w.xs, w.yq);
            double pri = 0, nz = 0, nAx = 0;
            for (int r = g.tid; r < m; r += G::nt) { pri = fmax(pri, fabs(w.yq[r] - w.zs[r])); nz = fmax(nz, fabs(w.zs[r])); nAx = fmax(nAx, fabs(w.yq[r])); }
            nls_Ats(g, w, w.ys, w.rhs);
            nls_Ps(g, w, c, w.xs, w.zt2);
            double dua = 0, nq = 0, nAty = 0, nPx = 0;
            for (int i = g.tid; i < n; i += G::nt) {
                double px = w.zt2[i];
                dua = fmax(dua, fabs(px + w.gs[i] + w.rhs[i])); nq = fmax(nq, fabs(w.gs[i])); nAty = fmax(nAty, fabs(w.rhs[i])); nPx = fmax(nPx, fabs(px));
            }
            pri = g.max(pri); nz = g.max(nz); nAx = g.max(nAx); dua = g.max(dua); nq = g.max(nq); nAty = g.max(nAty); nPx = g.max(nPx);
            if (pri < a.qp_eps && dua < a.qp_eps) break;
            double pn = pri / (fmax(nz, nAx) + 1e-10), dn = dua / (fmax(fmax(nq, nAty), nPx) + 1e-10);
            double est = fmin(fmax(rho0 * sqrt(pn / (dn + 1e-10)), 1e-6), 1e6);
            if (est > 5 * rho0 || est < rho0 / 5) { rho0 = est; set_rho(rho0); nls_factor(g, w, c, sigma); }
            NLS_T0();
        }
    }
    if (it > max_qp) it = max_qp;
    w.rho_keep = rho0;
    NLS_T0();
    nls_qp_polish<BS>(g, w, c);
    NLS_T0();
    for (int i = g.tid; i < n; i += G::nt) w.d[i] = w.D[i] * w.xs[i];
    for (int r = g.tid; r < m; r += G::nt) w.yq[r] = w.E[r] * w.ys[r] / c;
    g.sync();
    return it;
}

// The whole NLOptimizer::run core (NLOptimizer.hpp:519) for ONE controller on the structured storage.  z0: initial decision
// vector; z_out [n].  The SQP loop is nlmpc_solve_kernel's (nlmpc_sqp.cuh) with the block-diagonal BFGS update.
template <class S, class G, class WS>
__device__ __forceinline__ NlSResult nls_solve_instance(const G& g, WS& w, const NlSParams& a, const double* z0, const double* x0,
                                                        const double* p, double* z_out) {
    constexpr int nx = S::nx, nu = S::nu;
    constexpr int b = WS::b;
    const int ph = w.ph, ch = w.ch, n = w.n, me = w.me, mi = w.mi, mc = w.mc;
    const int mii = mi;                                   // no user equality constraints on this path
    const NlCompactMap map{ph, ch, nx, nu, WS::K};
    for (int i = g.tid; i < n; i += G::nt) w.z[i] = fmin(fmax(z0[i], a.lb[i]), a.ub[i]);
    auto reset_B = [&]() {
        for (int e = g.tid; e < ph * b * b; e += G::nt) { int l = (e / b) % b, k = e % b; w.Bb[e] = (l == k) ? 1.0 : 0.0; }
        if (g.tid == 0) w.Bsl[0] = 1.0;
        g.sync();
    };
    reset_B();
    w.rho_keep = a.rho0;
    double fval = 0;
    nl_eval_instance_map<S>(g, ph, ch, w.z, x0, p, w.X, w.U, w.tmp, w.g, w.ce, w.JeC, w.ci, w.JiC, 0, a.sx, a.su, (double*)nullptr, (double*)nullptr, map);
    fval = w.tmp[0];
    g.sync();
    bool have_y = false;
    int k = 0, qp_total = 0, status = 1, resets = 0;
    bool just_reset = false;
    int qp_cap = a.max_qp;
    auto violation = [&](const double* ce, const double* ci) {
        double v = 0;
        for (int r = g.tid; r < me; r += G::nt) v += fabs(ce[r]);
        for (int r = g.tid; r < mi; r += G::nt) v += r < mii ? fmax(ci[r], 0.0) : fabs(ci[r]);
        return g.sum(v);
    };
    auto merit_violation = [&](const double* ce, const double* ci) {
        double v = 0;
        for (int r = g.tid; r < me; r += G::nt) v += w.muv[r] * fabs(ce[r]);
        for (int r = g.tid; r < mi; r += G::nt) v += w.muv[me + r] * (r < mii ? fmax(ci[r], 0.0) : fabs(ci[r]));
        return g.sum(v);
    };
    for (k = 0; k < a.max_sqp; ++k) {
        qp_total += nls_qp_solve<nx + nu>(g, w, a, mii, have_y, qp_cap);
        have_y = true;
        double v0 = violation(w.ce, w.ci);
        double gd = 0;
        for (int r = g.tid; r < mc; r += G::nt) {
            const double lam = fabs(w.yq[r]);
            w.muv[r] = (k == 0 || just_reset) ? lam : fmax(lam, 0.5 * (w.muv[r] + lam));
        }
        for (int i = g.tid; i < n; i += G::nt) gd += w.g[i] * w.d[i];
        gd = g.sum(gd);
        g.sync();
        const double mv0 = merit_violation(w.ce, w.ci);
        const double phi0 = fval + mv0, dphi = gd - mv0;
        if (fabs(gd) < a.ftol * fmax(1.0, fabs(fval)) && v0 < 1e-8) { status = 0; ++k; break; }
        double t = 1.0, ft = fval;
        bool ls_ok = false;
        for (int ls = 0; ls < 25; ++ls) {
            for (int i = g.tid; i < n; i += G::nt) w.zt2[i] = w.z[i] + t * w.d[i];
            g.sync();
            nl_eval_instance_map<S>(g, ph, ch, w.zt2, x0, p, w.X, w.U, w.tmp, (double*)nullptr, w.cet, (double*)nullptr, w.cit, (double*)nullptr, 0, a.sx, a.su,
                                    (double*)nullptr, (double*)nullptr, map);
            ft = w.tmp[0];
            g.sync();
            const double mvt = merit_violation(w.cet, w.cit);
            if (ft + mvt <= phi0 + 1e-4 * t * dphi) { ls_ok = true; break; }
            t *= 0.5;
        }
        if (!ls_ok) {
            if (qp_cap == a.max_qp) { qp_cap = 5 * a.max_qp; continue; }
            if (just_reset || resets >= 5) { status = v0 < 1e-8 ? 0 : 1; ++k; break; }
            reset_B();
            ++resets; just_reset = true;
            continue;
        }
        just_reset = false;
        for (int i = g.tid; i < n; i += G::nt) w.sv[i] = t * w.d[i];
        for (int j = g.tid; j < n; j += G::nt) w.glo[j] = w.g[j] + w.col_dot(j, w.yq);
        g.sync();
        for (int i = g.tid; i < n; i += G::nt) w.z[i] += w.sv[i];
        g.sync();
        nl_eval_instance_map<S>(g, ph, ch, w.z, x0, p, w.X, w.U, w.tmp, w.g2, w.ce, w.JeC, w.ci, w.JiC, 0, a.sx, a.su, (double*)nullptr, (double*)nullptr, map);
        fval = w.tmp[0];
        g.sync();
        // block-diagonal damped BFGS: yk = gl_new - gl_old per block, Bs = B_G s_G
        for (int j = g.tid; j < n; j += G::nt) { w.rhs[j] = (w.g2[j] + w.col_dot(j, w.yq)) - w.glo[j]; w.xt[j] = w.b_row_dot(j, w.sv); }
        g.sync();
        double* sBs = w.cs; double* sy = w.cs + (ph + 1); double* sr = w.cs + 2 * (ph + 1); double* th = w.cs + 3 * (ph + 1);
        for (int s = g.tid; s <= ph; s += G::nt) {
            double a0 = 0, a1 = 0;
            if (s < ph) { for (int l = 0; l < b; ++l) { int jz = w.qz(s, l); if (jz >= 0) { a0 += w.sv[jz] * w.xt[jz]; a1 += w.sv[jz] * w.rhs[jz]; } } }
            else { a0 = w.sv[n - 1] * w.xt[n - 1]; a1 = w.sv[n - 1] * w.rhs[n - 1]; }
            sBs[s] = a0; sy[s] = a1;
            th[s] = (a1 >= 0.2 * a0) ? 1.0 : 0.8 * a0 / (a0 - a1);
        }
        g.sync();
        for (int j = g.tid; j < n; j += G::nt) {
            int s = j == n - 1 ? ph : (j < ph * nx ? j / nx : (j - ph * nx) / nu);
            w.rhs[j] = th[s] * w.rhs[j] + (1 - th[s]) * w.xt[j];
        }
        g.sync();
        for (int s = g.tid; s <= ph; s += G::nt) {
            double a0 = 0;
            if (s < ph) { for (int l = 0; l < b; ++l) { int jz = w.qz(s, l); if (jz >= 0) a0 += w.sv[jz] * w.rhs[jz]; } }
            else a0 = w.sv[n - 1] * w.rhs[n - 1];
            sr[s] = a0;
        }
        g.sync();
        for (int e = g.tid; e < ph * b * b; e += G::nt) {
            const int s = e / (b * b), l = (e / b) % b, q = e % b;
            const int jl = w.qz(s, l), jq = w.qz(s, q);
            if (jl >= 0 && jq >= 0 && sBs[s] > 1e-300) w.Bb[e] += -w.xt[jl] * w.xt[jq] / sBs[s] + w.rhs[jl] * w.rhs[jq] / sr[s];
        }
        if (g.tid == 0 && sBs[ph] > 1e-300) w.Bsl[0] += -w.xt[n - 1] * w.xt[n - 1] / sBs[ph] + w.rhs[n - 1] * w.rhs[n - 1] / sr[ph];
        for (int i = g.tid; i < n; i += G::nt) w.g[i] = w.g2[i];
        g.sync();
        double step = 0, zmax = 0;
        for (int i = g.tid; i < n; i += G::nt) { step = fmax(step, fabs(w.d[i])); zmax = fmax(zmax, fabs(w.z[i])); }
        step = g.max(step); zmax = g.max(zmax);
        double v1 = violation(w.ce, w.ci);
        if (step < a.tol * fmax(1.0, zmax) && v1 < 1e-8) { status = 0; ++k; break; }
        // overall work bound (the role NLopt's maxeval plays for the reference, NLOptimizer.hpp:135-147): 150 x max_qp ADMM iterations
        // per solve.  99 % of the unicycle Tph=30 cold starts need fewer than 26 k; without the bound the 1 % that do not converge
        // (300 major iterations x up to 1000 ADMM iterations) decide the time of the whole batch.
        if (qp_total > 150 * a.max_qp) { ++k; break; }
    }
    double vf = violation(w.ce, w.ci);
    for (int i = g.tid; i < n; i += G::nt) z_out[i] = w.z[i];
    g.sync();
    NlSResult r; r.cost = fval; r.viol = vf; r.status = status; r.iters = k; r.qp_iters = qp_total;
    return r;
}

// Can this (system, horizon) run on the structured path?  A declared ineq_per_stage = K with Tineq = (ph+1) K, no user equality
// constraints, a block that fits a warp, and scratch vectors long enough for a block.
template <class S>
__host__ __device__ inline bool nls_supported(int ph, int ch) {
    constexpr int K = NlIneqPerStage<S>::value;
    if (K <= 0 || NlHasEq<S>::value) return false;
    const int b = S::nx + S::nu, nb = S::nu + 1;
    const int n = ph * S::nx + ch * S::nu + 1, m = ph * S::nx + (ph + 1) * K + n;
    return S::nineq(ph) == (ph + 1) * K && b <= 32 && m >= b * b && m >= nb * b && m >= nb * nb && 4 * (ph + 1) + 8 >= b;
}

}  // namespace b200mpc
