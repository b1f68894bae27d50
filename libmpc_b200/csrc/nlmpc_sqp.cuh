// nlmpc_sqp.cuh -- batched NLMPC solve for sm_100a: one warp per controller, the whole NLP solve in
// shared memory.  Replaces what NLOptimizer::run hands to NLopt's SLSQP (include/mpc/NLMPC/NLOptimizer.hpp:412-638):
//
//     min f(z)   s.t.  c_eq(z) = 0 (multiple-shooting dynamics),  c_in(z) <= 0 (user inequalities),  lb <= z <= ub
//
// with f, its forward-difference gradient and the central-difference Jacobians evaluated exactly as the reference does
// (nl_eval_instance in nlmpc_kernels.cuh).  The solver is a damped-BFGS SQP (the algorithm family of Kraft's SLSQP): every
// major iteration solves   min 1/2 d'Bd + g'd  s.t.  J_eq d = -c_eq,  J_in d <= -c_in,  lb-z <= d <= ub-z
// with a dense OSQP-style ADMM (Ruiz equilibration, rho_eq = 1e3 rho, over-relaxation 1.6, adaptive rho with dense
// refactorisation), globalised by an L1 merit function with backtracking.  tests/nlmpc_sqp_reference.py is the
// executable specification; solution-level parity is against SciPy's SLSQP on the restated formulation
// (oracle/nlmpc_slsqp.py).  Everything (B, the KKT factor, both Jacobians, all vectors) lives in shared memory, which
// bounds the problem size: nz up to ~64 (vanderpol nz=26, the shipped ugv nz=61); larger systems need the stage-
// structured LTV kernel (next step).
#pragma once
#include "nlmpc_kernels.cuh"

namespace b200mpc {

struct NlSolveArgs {
    int ph, ch, batch;
    const double* z0;       // [batch, nz] initial decision vectors (NLOptimizer::run's optX0)
    const double* x0;       // [batch, nx]
    const double* params; long long param_stride;
    const double* lb; const double* ub;      // [nz] shared by the batch
    int max_sqp, max_qp;
    double tol, qp_eps, rho0;
    double* z_out;          // [batch, nz]
    double* cost;           // [batch]
    double* viol;           // [batch] sum |c_eq| + sum max(c_in,0) at the solution
    int* status;            // 0 converged, 1 iteration limit
    int* iters;             // SQP iterations
    int* qp_iters;          // total ADMM iterations
};

__device__ __forceinline__ double nl_wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double nl_wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double nl_lim(double v) { v = v < 1e-4 ? 1.0 : v; return v > 1e4 ? 1e4 : v; }

// Per-warp shared-memory view
struct NlWs {
    int n, me, mi, m, ld;
    double *B, *H, *Je, *Ji;                 // n x ld, n x ld, me x ld, mi x ld (row-major)
    double *z, *g, *g2, *d, *xs, *xt, *D, *gs, *rhs, *tmp, *glo, *sv, *zt2;     // n
    double *E, *ls, *us, *zs, *ys, *rho, *yq, *w;                               // m
    double *ce, *ci, *cet, *cit;                                                // me, mi, me, mi
    double *X, *U;
    __host__ __device__ static size_t doubles(int n, int me, int mi, int ph, int nx, int nu) {
        int ld = n | 1, m = me + mi + n;
        return (size_t)2 * n * ld + (size_t)(me + mi) * ld + 13 * (size_t)n + 8 * (size_t)m + 2 * (size_t)(me + mi) + (size_t)(ph + 1) * (nx + nu) + 8;
    }
    __device__ void carve(double* p, int n_, int me_, int mi_, int ph, int nx, int nu) {
        n = n_; me = me_; mi = mi_; m = me + mi + n; ld = n | 1;
        B = p; p += (size_t)n * ld; H = p; p += (size_t)n * ld; Je = p; p += (size_t)me * ld; Ji = p; p += (size_t)mi * ld;
        double** nv[] = {&z, &g, &g2, &d, &xs, &xt, &D, &gs, &rhs, &tmp, &glo, &sv, &zt2};
        for (auto q : nv) { *q = p; p += n; }
        double** mv[] = {&E, &ls, &us, &zs, &ys, &rho, &yq, &w};
        for (auto q : mv) { *q = p; p += m; }
        ce = p; p += me; ci = p; p += mi; cet = p; p += me; cit = p; p += mi;
        X = p; p += (ph + 1) * nx; U = p;
    }
};

// row r of the constraint matrix A = [Je; Ji; I] (first me+mi rows only)
__device__ __forceinline__ const double* nl_row(const NlWs& w, int r) { return r < w.me ? w.Je + (size_t)r * w.ld : w.Ji + (size_t)(r - w.me) * w.ld; }

// H = c D B D + sigma I + (E A D)' diag(rho) (E A D)  -> Cholesky -> inverse of the factor, all in w.H (lower triangle)
__device__ bool nl_factor(NlWs& w, int lane, double c, double sigma) {
    const int n = w.n, ld = w.ld, mc = w.me + w.mi;
    for (int r = lane; r < mc; r += 32) w.w[r] = w.rho[r] * w.E[r] * w.E[r];
    __syncwarp();
    const int npairs = n * (n + 1) / 2;
    for (int pidx = lane; pidx < npairs; pidx += 32) {
        int i = (int)((sqrt(8.0 * pidx + 1.0) - 1.0) * 0.5);
        while ((i + 1) * (i + 2) / 2 <= pidx) ++i;
        while (i * (i + 1) / 2 > pidx) --i;
        int j = pidx - i * (i + 1) / 2;
        double acc = 0;
        for (int r = 0; r < mc; ++r) { const double* a = nl_row(w, r); acc = fma(w.w[r] * a[i], a[j], acc); }
        double v = w.D[i] * (c * w.B[(size_t)i * ld + j] + acc) * w.D[j];
        if (i == j) { int rb = mc + i; v += sigma + w.rho[rb] * w.E[rb] * w.E[rb] * w.D[i] * w.D[i]; }
        w.H[(size_t)i * ld + j] = v;
    }
    __syncwarp();
    bool ok = true;
    for (int k = 0; k < n; ++k) {                       // right-looking Cholesky, lower, in place
        double dkk = w.H[(size_t)k * ld + k];
        if (!(dkk > 0.0)) ok = false;
        double piv = sqrt(dkk), inv = 1.0 / piv;
        __syncwarp();
        for (int r = k + lane; r < n; r += 32) w.H[(size_t)r * ld + k] = (r == k) ? piv : w.H[(size_t)r * ld + k] * inv;
        __syncwarp();
        for (int r = k + 1 + lane; r < n; r += 32) {
            double lrk = w.H[(size_t)r * ld + k];
            for (int q = k + 1; q <= r; ++q) w.H[(size_t)r * ld + q] -= lrk * w.H[(size_t)q * ld + k];
        }
        __syncwarp();
    }
    // in-place inverse of the lower-triangular factor (column by column from the right)
    for (int j = n - 1; j >= 0; --j) {
        double ljj = 1.0 / w.H[(size_t)j * ld + j];
        __syncwarp();
        // t = Linv[j+1:, j+1:] * L[j+1:, j]
        for (int r = j + 1 + lane; r < n; r += 32) {
            double acc = 0;
            for (int q = j + 1; q <= r; ++q) acc = fma(w.H[(size_t)r * ld + q], w.H[(size_t)q * ld + j], acc);
            w.tmp[r] = acc;
        }
        __syncwarp();
        for (int r = j + 1 + lane; r < n; r += 32) w.H[(size_t)r * ld + j] = -ljj * w.tmp[r];
        if (lane == 0) w.H[(size_t)j * ld + j] = ljj;
        __syncwarp();
    }
    return !__any_sync(0xffffffffu, !ok);
}

// xt = (Linv' Linv) rhs
__device__ void nl_kkt_apply(NlWs& w, int lane) {
    const int n = w.n, ld = w.ld;
    for (int i = lane; i < n; i += 32) {
        double acc = 0;
        for (int q = 0; q <= i; ++q) acc = fma(w.H[(size_t)i * ld + q], w.rhs[q], acc);
        w.tmp[i] = acc;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        double acc = 0;
        for (int q = i; q < n; ++q) acc = fma(w.H[(size_t)q * ld + i], w.tmp[q], acc);
        w.xt[i] = acc;
    }
    __syncwarp();
}
// out_r = E_r * (A (D.x))_r for all m rows
__device__ void nl_As(NlWs& w, int lane, const double* x, double* out) {
    const int n = w.n, mc = w.me + w.mi;
    for (int i = lane; i < n; i += 32) w.tmp[i] = w.D[i] * x[i];
    __syncwarp();
    for (int r = lane; r < w.m; r += 32) {
        double a;
        if (r < mc) { const double* row = nl_row(w, r); a = 0; for (int j = 0; j < n; ++j) a = fma(row[j], w.tmp[j], a); }
        else a = w.tmp[r - mc];
        out[r] = w.E[r] * a;
    }
    __syncwarp();
}
// out_j = D_j * (A' (E.v))_j
__device__ void nl_Ats(NlWs& w, int lane, const double* v, double* out) {
    const int n = w.n, mc = w.me + w.mi;
    for (int r = lane; r < w.m; r += 32) w.w[r] = w.E[r] * v[r];
    __syncwarp();
    for (int j = lane; j < n; j += 32) {
        double a = w.w[mc + j];
        for (int r = 0; r < mc; ++r) a = fma(nl_row(w, r)[j], w.w[r], a);
        out[j] = w.D[j] * a;
    }
    __syncwarp();
}

// Dense OSQP-style ADMM for the QP subproblem.  In: B, g, Je, Ji, ce, ci, z, lb, ub; warm dual yq (if have_y).
// Out: d (step), yq (multipliers, unscaled).  Returns ADMM iterations.
__device__ int nl_qp_solve(NlWs& w, int lane, const NlSolveArgs& a, bool have_y) {
    const int n = w.n, me = w.me, mi = w.mi, mc = me + mi, m = w.m, ld = w.ld;
    const double sigma = 1e-6, alpha = 1.6;
    // ---- Ruiz equilibration (10 passes) with cost normalisation
    for (int i = lane; i < n; i += 32) { w.D[i] = 1.0; w.gs[i] = w.g[i]; }
    for (int r = lane; r < m; r += 32) w.E[r] = 1.0;
    double c = 1.0;
    __syncwarp();
    for (int pass = 0; pass < 10; ++pass) {
        for (int j = lane; j < n; j += 32) {        // column norms -> xt ; uses old D,E
            double cn = 0;
            for (int i = 0; i < n; ++i) cn = fmax(cn, w.D[i] * fabs(w.B[(size_t)i * ld + j]));
            cn *= c * w.D[j];
            double an = 0;
            for (int r = 0; r < mc; ++r) an = fmax(an, w.E[r] * fabs(nl_row(w, r)[j]));
            an = fmax(an, w.E[mc + j]) * w.D[j];
            w.xt[j] = 1.0 / sqrt(nl_lim(fmax(cn, an)));
        }
        for (int r = lane; r < m; r += 32) {        // row norms -> w
            double rn;
            if (r < mc) { const double* row = nl_row(w, r); rn = 0; for (int j = 0; j < n; ++j) rn = fmax(rn, fabs(row[j]) * w.D[j]); rn *= w.E[r]; }
            else rn = w.E[r] * w.D[r - mc];
            w.w[r] = 1.0 / sqrt(nl_lim(rn));
        }
        __syncwarp();
        for (int j = lane; j < n; j += 32) { w.D[j] *= w.xt[j]; w.gs[j] *= w.xt[j]; }
        for (int r = lane; r < m; r += 32) w.E[r] *= w.w[r];
        __syncwarp();
        double psum = 0, qmax = 0;
        for (int j = lane; j < n; j += 32) {
            double cn = 0;
            for (int i = 0; i < n; ++i) cn = fmax(cn, w.D[i] * fabs(w.B[(size_t)i * ld + j]));
            psum += c * cn * w.D[j];
            qmax = fmax(qmax, fabs(w.gs[j]));
        }
        psum = nl_wsum(psum); qmax = nl_wmax(qmax);
        double ct = fmax(psum / n, nl_lim(qmax));
        ct = 1.0 / nl_lim(ct);
        c *= ct;
        for (int j = lane; j < n; j += 32) w.gs[j] *= ct;
        __syncwarp();
    }
    // ---- scaled bounds, rho per row
    double rho0 = a.rho0;
    for (int r = lane; r < m; r += 32) {
        double l, u;
        if (r < me) { l = u = -w.ce[r]; }
        else if (r < mc) { l = -INFINITY; u = -w.ci[r - me]; }
        else { int j = r - mc; l = a.lb[j] - w.z[j]; u = a.ub[j] - w.z[j]; }
        w.ls[r] = w.E[r] * l; w.us[r] = w.E[r] * u;
    }
    __syncwarp();
    auto set_rho = [&](double r0) {
        for (int r = lane; r < m; r += 32)         // OSQP's row classes: no bounds -> RHO_MIN, equality -> 1e3 rho
            w.rho[r] = (w.ls[r] < -1e20 && w.us[r] > 1e20) ? 1e-6 : ((w.us[r] - w.ls[r]) < 1e-9) ? 1e3 * r0 : r0;
        __syncwarp();
    };
    set_rho(rho0);
    nl_factor(w, lane, c, sigma);
    // ---- start: x = 0, y = warm (scaled), z = clip(A x)
    for (int i = lane; i < n; i += 32) w.xs[i] = 0.0;
    for (int r = lane; r < m; r += 32) { w.ys[r] = have_y ? c * w.yq[r] / w.E[r] : 0.0; w.zs[r] = fmin(fmax(0.0, w.ls[r]), w.us[r]); }
    __syncwarp();
    int it = 0;
    for (it = 1; it <= a.max_qp; ++it) {
        for (int r = lane; r < m; r += 32) w.yq[r] = w.rho[r] * w.zs[r] - w.ys[r];     // yq as temp
        __syncwarp();
        nl_Ats(w, lane, w.yq, w.rhs);
        for (int i = lane; i < n; i += 32) w.rhs[i] += sigma * w.xs[i] - w.gs[i];
        __syncwarp();
        nl_kkt_apply(w, lane);
        nl_As(w, lane, w.xt, w.yq);                                                    // z~ in yq
        for (int i = lane; i < n; i += 32) w.xs[i] = alpha * w.xt[i] + (1 - alpha) * w.xs[i];
        for (int r = lane; r < m; r += 32) {
            double zr = alpha * w.yq[r] + (1 - alpha) * w.zs[r];
            double zn = fmin(fmax(zr + w.ys[r] / w.rho[r], w.ls[r]), w.us[r]);
            w.ys[r] += w.rho[r] * (zr - zn);
            w.zs[r] = zn;
        }
        __syncwarp();
        if (it % 25 == 0) {
            nl_As(w, lane, w.xs, w.yq);                       // Ax
            double pri = 0, nz = 0, nAx = 0;
            for (int r = lane; r < m; r += 32) { pri = fmax(pri, fabs(w.yq[r] - w.zs[r])); nz = fmax(nz, fabs(w.zs[r])); nAx = fmax(nAx, fabs(w.yq[r])); }
            nl_Ats(w, lane, w.ys, w.rhs);                     // A'y
            double dua = 0, nq = 0, nAty = 0, nPx = 0;
            for (int i = lane; i < n; i += 32) {              // Px = c D B D x
                double acc = 0;
                for (int j = 0; j < n; ++j) acc = fma(w.B[(size_t)i * ld + j], w.D[j] * w.xs[j], acc);
                double px = c * w.D[i] * acc;
                dua = fmax(dua, fabs(px + w.gs[i] + w.rhs[i])); nq = fmax(nq, fabs(w.gs[i])); nAty = fmax(nAty, fabs(w.rhs[i])); nPx = fmax(nPx, fabs(px));
            }
            pri = nl_wmax(pri); nz = nl_wmax(nz); nAx = nl_wmax(nAx); dua = nl_wmax(dua); nq = nl_wmax(nq); nAty = nl_wmax(nAty); nPx = nl_wmax(nPx);
            if (pri < a.qp_eps && dua < a.qp_eps) break;
            double pn = pri / (fmax(nz, nAx) + 1e-10), dn = dua / (fmax(fmax(nq, nAty), nPx) + 1e-10);
            double est = fmin(fmax(rho0 * sqrt(pn / (dn + 1e-10)), 1e-6), 1e6);
            if (est > 5 * rho0 || est < rho0 / 5) { rho0 = est; set_rho(rho0); nl_factor(w, lane, c, sigma); }
        }
    }
    if (it > a.max_qp) it = a.max_qp;
    for (int i = lane; i < n; i += 32) w.d[i] = w.D[i] * w.xs[i];
    for (int r = lane; r < m; r += 32) w.yq[r] = w.E[r] * w.ys[r] / c;
    __syncwarp();
    return it;
}

template <class S>
__global__ void __launch_bounds__(64) nlmpc_solve_kernel(const NlSolveArgs a) {
    extern __shared__ __align__(16) double nls_smem[];
    constexpr int nx = S::nx, nu = S::nu;
    const int ph = a.ph, ch = a.ch;
    const int n = ph * nx + ch * nu + 1, me = ph * nx, mi = S::nineq(ph);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    NlWs w;
    w.carve(nls_smem + (size_t)warp * NlWs::doubles(n, me, mi, ph, nx, nu), n, me, mi, ph, nx, nu);
    const int mc = me + mi, ld = w.ld;
    for (int inst = blockIdx.x * wpb + warp; inst < a.batch; inst += gridDim.x * wpb) {
        const double* p = a.params + (size_t)inst * a.param_stride;
        const double* x0 = a.x0 + (size_t)inst * nx;
        for (int i = lane; i < n; i += 32) w.z[i] = fmin(fmax(a.z0[(size_t)inst * n + i], a.lb[i]), a.ub[i]);
        for (int e = lane; e < n * ld; e += 32) { int i = e / ld, j = e - i * ld; w.B[e] = (i == j) ? 1.0 : 0.0; }
        __syncwarp();
        double fval = 0;
        nl_eval_instance<S>(lane, ph, ch, w.z, x0, p, w.X, w.U, w.tmp, w.g, w.ce, w.Je, w.ci, w.Ji, ld);
        fval = w.tmp[0];
        __syncwarp();
        double mu = 1.0;
        bool have_y = false;
        int k = 0, qp_total = 0, status = 1;
        auto violation = [&](const double* ce, const double* ci) {
            double v = 0;
            for (int r = lane; r < me; r += 32) v += fabs(ce[r]);
            for (int r = lane; r < mi; r += 32) v += fmax(ci[r], 0.0);
            return nl_wsum(v);
        };
        for (k = 0; k < a.max_sqp; ++k) {
            qp_total += nl_qp_solve(w, lane, a, have_y);
            have_y = true;
            // L1 merit line search
            double v0 = violation(w.ce, w.ci);
            double ymax = 0, gd = 0;
            for (int r = lane; r < mc; r += 32) ymax = fmax(ymax, fabs(w.yq[r]));
            for (int i = lane; i < n; i += 32) gd += w.g[i] * w.d[i];
            ymax = nl_wmax(ymax); gd = nl_wsum(gd);
            mu = fmax(mu, 1.1 * ymax);
            const double phi0 = fval + mu * v0, dphi = gd - mu * v0;
            double t = 1.0, ft = fval;
            for (int ls = 0; ls < 25; ++ls) {
                for (int i = lane; i < n; i += 32) w.zt2[i] = w.z[i] + t * w.d[i];
                __syncwarp();
                nl_eval_instance<S>(lane, ph, ch, w.zt2, x0, p, w.X, w.U, w.tmp, nullptr, w.cet, nullptr, w.cit, nullptr, ld);
                ft = w.tmp[0];
                __syncwarp();
                double vt = violation(w.cet, w.cit);
                if (ft + mu * vt <= phi0 + 1e-4 * t * dphi) break;
                t *= 0.5;
            }
            // s = t d ; Lagrangian gradient at the old point with the new multipliers
            for (int i = lane; i < n; i += 32) { w.sv[i] = t * w.d[i]; }
            __syncwarp();
            for (int j = lane; j < n; j += 32) {
                double acc = w.g[j];
                for (int r = 0; r < mc; ++r) acc = fma(nl_row(w, r)[j], w.yq[r], acc);
                w.glo[j] = acc;
            }
            for (int i = lane; i < n; i += 32) w.z[i] += w.sv[i];
            __syncwarp();
            nl_eval_instance<S>(lane, ph, ch, w.z, x0, p, w.X, w.U, w.tmp, w.g2, w.ce, w.Je, w.ci, w.Ji, ld);
            fval = w.tmp[0];
            __syncwarp();
            // damped BFGS:  yk = gl_new - gl_old,  Bs = B s
            double sBs = 0, sy = 0;
            for (int j = lane; j < n; j += 32) {
                double acc = w.g2[j];
                for (int r = 0; r < mc; ++r) acc = fma(nl_row(w, r)[j], w.yq[r], acc);
                w.rhs[j] = acc - w.glo[j];                     // yk
                double bs = 0;
                for (int q = 0; q < n; ++q) bs = fma(w.B[(size_t)j * ld + q], w.sv[q], bs);
                w.xt[j] = bs;                                  // Bs
                sBs += w.sv[j] * bs; sy += w.sv[j] * w.rhs[j];
            }
            sBs = nl_wsum(sBs); sy = nl_wsum(sy);
            __syncwarp();
            if (sBs > 1e-300) {
                double theta = (sy >= 0.2 * sBs) ? 1.0 : 0.8 * sBs / (sBs - sy);
                double sr = 0;
                for (int j = lane; j < n; j += 32) { double r = theta * w.rhs[j] + (1 - theta) * w.xt[j]; w.rhs[j] = r; sr += w.sv[j] * r; }
                sr = nl_wsum(sr);
                __syncwarp();
                for (int e = lane; e < n * n; e += 32) {
                    int i = e / n, j = e - i * n;
                    w.B[(size_t)i * ld + j] += -w.xt[i] * w.xt[j] / sBs + w.rhs[i] * w.rhs[j] / sr;
                }
            }
            for (int i = lane; i < n; i += 32) w.g[i] = w.g2[i];
            __syncwarp();
            double step = 0, zmax = 0;
            // the full QP step d is small only at a KKT point (t*d can be small far from one)
            for (int i = lane; i < n; i += 32) { step = fmax(step, fabs(w.d[i])); zmax = fmax(zmax, fabs(w.z[i])); }
            step = nl_wmax(step); zmax = nl_wmax(zmax);
            double v1 = violation(w.ce, w.ci);
            if (step < a.tol * fmax(1.0, zmax) && v1 < 1e-8) { status = 0; ++k; break; }
        }
        double vf = violation(w.ce, w.ci);
        for (int i = lane; i < n; i += 32) a.z_out[(size_t)inst * n + i] = w.z[i];
        if (lane == 0) {
            a.cost[inst] = fval; a.viol[inst] = vf; a.status[inst] = status; a.iters[inst] = k; a.qp_iters[inst] = qp_total;
        }
        __syncwarp();
    }
}

}  // namespace b200mpc
