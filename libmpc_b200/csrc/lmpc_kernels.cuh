// lmpc_kernels.cuh -- batched linear-MPC solve for sm_100a: ONE persistent kernel, one warp per MPC instance.
//
// What it computes (reference anchors, paths relative to the libmpc++ repository):
//   * the per-step QP terms q,l,u of ProblemBuilder::get               include/mpc/LMPC/ProblemBuilder.hpp:528-633
//   * the time-invariant P,A of buildTimeInvariantTems -- never formed: include/mpc/LMPC/ProblemBuilder.hpp:642-825
//     every product with them is evaluated on the stage structure (rows/cols of one horizon stage at a time)
//   * OSQP v0.6.3 (scale_data, set_rho_vec, ADMM loop, adaptive rho, termination + infeasibility tests, polish) as
//     driven by LOptimizer::run                                        include/mpc/LMPC/LOptimizer.hpp:241-284
//   * the unpack / status map of LOptimizer::run                       include/mpc/LMPC/LOptimizer.hpp:292-361,386-415
//
// Data layout.  Variables are kept stage-major: w_i = [x_i ; xu_i ; du_i] (b = nx+2nu doubles, the last stage has no
// du).  Constraint rows are kept stage-major too: stage i owns [box(i) ; out(i) ; sc(i) ; eq(i+1) ; du(i)], and the ne
// rows eq(0) sit in front.  With that ownership one ADMM iteration is exactly one forward sweep (rhs assembly fused
// with the block forward substitution) and one backward sweep (back substitution fused with z~ = A x~, the relaxation,
// the projection and the dual update); nothing of size m x n is ever stored.
//
// The reduced KKT matrix  H = Pbar + sigma I + Abar' diag(rho) Abar  is block tridiagonal over the stages; it is
// factorised by a block Cholesky (diagonal blocks inverted explicitly, so the sweeps are mat-vecs, not substitutions).
// Per-instance state lives in a per-warp-slot workspace in global memory that is sized by the number of RESIDENT warps
// (not by the batch), so it stays L2 resident; stage factor blocks are staged through shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace b200mpc {

constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4;
constexpr double kOsqpInfty = 1e30, kMinScaling = 1e-4, kMaxScaling = 1e4;

// OSQP status_val (constants.h of v0.6.3)
enum { OSQP_DUAL_INFEASIBLE_INACCURATE = 4, OSQP_PRIMAL_INFEASIBLE_INACCURATE = 3, OSQP_SOLVED_INACCURATE = 2,
       OSQP_SOLVED = 1, OSQP_MAX_ITER_REACHED = -2, OSQP_PRIMAL_INFEASIBLE = -3, OSQP_DUAL_INFEASIBLE = -4,
       OSQP_NON_CVX = -7, OSQP_UNSOLVED = -10, B200_SETUP_ERROR = -1 };
// mpc::ResultStatus
enum { RS_SUCCESS = 0, RS_MAX_ITERATION = 1, RS_INFEASIBLE = 2, RS_ERROR = 3, RS_UNKNOWN = 4 };

struct Dm {
    int nx, nu, ndu, ny, ph, ch;
    int ne, b, n, m;        // ne=nx+nu, b=ne+nu
    int RS, RSL;            // rows owned by a stage (<ph) / by the last stage
    int oBOX, oOUT, oSC, oEQ, oDU;   // offsets inside a stage's row segment
    int ldG, ldC, ldb;      // odd leading dimensions (bank-conflict free column walks)
    int FS;                 // doubles per stage factor block: packed Linv (b(b+1)/2) + Lc (ne x ldb), even
    int oLc;                // offset of Lc inside a factor block
    // reference row offsets (ProblemBuilder.hpp:70-76)
    int M0, M1, M2, M3;
    __host__ __device__ void derive() {
        ne = nx + nu; b = ne + nu;
        n = (ph + 1) * ne + ph * nu;
        m = 2 * (ph + 1) * ne + (ph + 1) * ny + ph * nu + (ph + 1);
        oBOX = 0; oOUT = ne; oSC = ne + ny; oEQ = ne + ny + 1; oDU = oEQ + ne;
        RS = oDU + nu; RSL = oEQ;
        ldG = b | 1; ldC = nx | 1; ldb = b | 1;
        oLc = (b * (b + 1) / 2 + 1) & ~1;
        FS = (oLc + ne * ldb + 1) & ~1;
        M0 = (ph + 1) * ne; M1 = 2 * (ph + 1) * ne; M2 = M1 + (ph + 1) * ny; M3 = M2 + ph * nu;
    }
    __host__ __device__ int roff(int i) const { return ne + i * RS; }
    __host__ __device__ int rcount(int i) const { return i < ph ? RS : RSL; }
    __host__ __device__ int bcount(int i) const { return i < ph ? b : ne; }
    // workspace sizes (doubles)
    __host__ __device__ size_t ws_doubles() const {
        return (size_t)6 * n + (size_t)8 * m + (size_t)(ph + 1) * FS + ((m + 7) / 8) + 16;
    }
    // shared memory per warp (doubles)
    __host__ __device__ int smem_doubles() const {
        int model = ne * ldG + ny * ldC + ne;
        int vec = 5 * b + 2 * ne + RS + ny + 4;
        int fac = 2 * b * ldb + 2 * ne * ldb + nx * nx;
        int ring = 2 * FS;
        return (model + vec + (fac > ring ? fac : ring) + 3) & ~1;
    }
};

struct Arr { const double* p; long long stride; };

struct Prob {
    Arr A, B, C, Bd, Dd, OW, UW, DUW, XMin, XMax, YMin, YMax, UMin, UMax, SMin, SMax, SX, SU, yRef, uRef, duRef, uMeas;
    const double* x0; const double* u0;       // [batch*nx], [batch*nu]
    const double* warm_x; const double* warm_y; // reference order, [batch*n],[batch*m]; used iff warm!=0
    int warm;
};

struct Params {
    int max_iter, adaptive_rho, polish, scaling, check_termination, adaptive_rho_interval, polish_refine_iter;
    double alpha, rho, sigma, delta, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
};

struct Out {
    double* cmd; double* cost; int* status; int* solver_status; int* feasible; int* iters; int* rho_updates; int* polish;
    double* seq_state; double* seq_input; double* seq_output;   // may be null
    double* sol_x; double* sol_y;                               // reference order [batch*n],[batch*m]; may be null
    double* prev_cmd;                                           // [batch*nu] last command (failure semantics)
};

// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ bool wany(bool p) { return __any_sync(0xffffffffu, p); }
__device__ __forceinline__ double lim_scaling(double v) {
    v = v < kMinScaling ? 1.0 : v;
    return v > kMaxScaling ? kMaxScaling : v;
}
__device__ __forceinline__ double ldp(const Arr& a, int inst, int idx) { return __ldg(a.p + (long long)inst * a.stride + idx); }

struct InfoNorms {
    double pri, dua;                 // unscaled residual norms (termination)
    double nz, nAx, nq, nAty, nPx;   // unscaled normalisers
    double spri, sdua, snz, snAx, snq, snAty, snPx;   // scaled (rho estimate)
    double xPx, qx;
};

// Per-warp context ------------------------------------------------------------------------------------------
struct Ctx {
    const Dm& d; const Params& p; const Prob& pr; int inst; int lane;
    __device__ Ctx(const Dm& d_, const Params& p_, const Prob& pr_) : d(d_), p(p_), pr(pr_) {}
    // shared memory
    double *G, *Cm, *s;
    double *uxc, *uxn, *xn, *vrowA, *veqp, *tA, *tB, *vtmp, *yv;
    double *fb0, *fb1;               // factor ring (aliases the factor scratch)
    double *S, *Li, *Hc, *Lc, *Pblk; // factor scratch
    // workspace (global)
    double *D, *qs, *x, *t, *va, *px;            // variables (n each)
    double *E, *lo, *up, *z, *y, *ra, *rb, *rc;   // rows (m each)
    double *fac; int8_t* rtype;
    double c;                       // cost scaling
    double rsel[3], rinv[3];        // rho by row type
    // per-stage unscaled problem data helpers
    __device__ __forceinline__ int jcol(int i) const { return i > 0 ? i - 1 : 0; }
    __device__ __forceinline__ double wO(int i, int r) const { return ldp(pr.OW, inst, jcol(i) * d.ny + r); }
    __device__ __forceinline__ double wU(int i, int r) const { return ldp(pr.UW, inst, jcol(i) * d.nu + r); }
    __device__ __forceinline__ double wDU(int i, int r) const { return ldp(pr.DUW, inst, i * d.nu + r); }
    __device__ __forceinline__ int voff(int i) const { return i * d.b; }
};

// ---- model to shared memory ---------------------------------------------------------------------------------
__device__ void load_model(Ctx& c) {
    const Dm& d = c.d;
    for (int e = c.lane; e < d.ne * d.b; e += 32) {
        int r = e / d.b, k = e - r * d.b;
        double v;
        if (r < d.nx) {
            if (k < d.nx) v = ldp(c.pr.A, c.inst, r * d.nx + k);
            else if (k < d.ne) v = ldp(c.pr.B, c.inst, r * d.nu + (k - d.nx));
            else v = ldp(c.pr.B, c.inst, r * d.nu + (k - d.ne));
        } else {
            int j = r - d.nx;
            v = ((k >= d.nx && k < d.ne && k - d.nx == j) || (k >= d.ne && k - d.ne == j)) ? 1.0 : 0.0;
        }
        c.G[r * d.ldG + k] = v;
    }
    for (int e = c.lane; e < d.ny * d.nx; e += 32) {
        int r = e / d.nx, k = e - r * d.nx;
        c.Cm[r * d.ldC + k] = ldp(c.pr.C, c.inst, e);
    }
    for (int k = c.lane; k < d.ne; k += 32)
        c.s[k] = k < d.nx ? ldp(c.pr.SX, c.inst, k) : ldp(c.pr.SU, c.inst, k - d.nx);
    __syncwarp();
}

// ---- unscaled q of stage i, variable k (ProblemBuilder.hpp:586-595); needs yv = wO*(-yRef + Dd d) in smem ------
__device__ void stage_q_prepare(Ctx& c, int i) {
    const Dm& d = c.d;
    int j = c.jcol(i);
    for (int r = c.lane; r < d.ny; r += 32) {
        double acc = -ldp(c.pr.yRef, c.inst, j * d.ny + r);
        for (int q = 0; q < d.ndu; ++q) acc += ldp(c.pr.Dd, c.inst, r * d.ndu + q) * ldp(c.pr.uMeas, c.inst, j * d.ndu + q);
        c.yv[r] = c.wO(i, r) * acc;
    }
    __syncwarp();
}
__device__ __forceinline__ double stage_q(Ctx& c, int i, int k) {
    const Dm& d = c.d;
    int j = c.jcol(i);
    if (k < d.nx) {
        double acc = 0;
        for (int r = 0; r < d.ny; ++r) acc += c.Cm[r * d.ldC + k] * c.yv[r];
        return acc;
    } else if (k < d.ne) {
        int q = k - d.nx;
        return c.wU(i, q) * (-ldp(c.pr.uRef, c.inst, j * d.nu + q));
    } else {
        int q = k - d.ne;
        return -(c.wDU(i, q) * ldp(c.pr.duRef, c.inst, j * d.nu + q));
    }
}
// unscaled bounds of row r of stage i (ProblemBuilder.hpp:597-630,727-809); eq0 handled by caller
__device__ __forceinline__ void stage_bounds(Ctx& c, int i, int r, double& l, double& u) {
    const Dm& d = c.d;
    int j = c.jcol(i);
    const double inf = INFINITY;
    if (r < d.oOUT) {            // box(i): [minX(i); minU(min(i,ph-1))]
        if (r < d.nx) { l = ldp(c.pr.XMin, c.inst, j * d.nx + r); u = ldp(c.pr.XMax, c.inst, j * d.nx + r); }
        else { int q = r - d.nx; int col = i < d.ph ? i : d.ph - 1;
               l = ldp(c.pr.UMin, c.inst, col * d.nu + q); u = ldp(c.pr.UMax, c.inst, col * d.nu + q); }
    } else if (r < d.oSC) {      // out(i): minY(i) - Dd d
        int q = r - d.oOUT;
        double off = 0;
        for (int e = 0; e < d.ndu; ++e) off -= ldp(c.pr.Dd, c.inst, q * d.ndu + e) * ldp(c.pr.uMeas, c.inst, j * d.ndu + e);
        l = ldp(c.pr.YMin, c.inst, j * d.ny + q) + off; u = ldp(c.pr.YMax, c.inst, j * d.ny + q) + off;
    } else if (r < d.oEQ) {      // sc(i)
        l = ldp(c.pr.SMin, c.inst, j); u = ldp(c.pr.SMax, c.inst, j);
    } else if (r < d.oDU) {      // eq(i+1): -ssBv * d(i)   (stage index i+1 uses uMeas column i)
        int q = r - d.oEQ;
        double v = 0;
        if (q < d.nx) for (int e = 0; e < d.ndu; ++e) v -= ldp(c.pr.Bd, c.inst, q * d.ndu + e) * ldp(c.pr.uMeas, c.inst, i * d.ndu + e);
        l = u = v;
    } else {                      // du(i): frozen strictly after ch (ProblemBuilder.hpp:784-785)
        bool frozen = i > d.ch;
        l = frozen ? 0.0 : -inf; u = frozen ? 0.0 : inf;
    }
}

// ---- Pblk = C' diag(wO_i) C (nx x nx), cached across stages with identical weights -----------------------------
__device__ void stage_Pblk(Ctx& c, int i, bool& valid) {
    const Dm& d = c.d;
    bool same = valid && i > 0;
    if (same) {
        bool diff = false;
        for (int r = c.lane; r < d.ny; r += 32) diff |= (c.wO(i, r) != c.wO(i - 1, r));
        same = !wany(diff);
    }
    if (same) return;
    for (int r = c.lane; r < d.ny; r += 32) c.yv[r] = c.wO(i, r);
    __syncwarp();
    for (int e = c.lane; e < d.nx * d.nx; e += 32) {
        int a = e / d.nx, k = e - a * d.nx;
        double acc = 0;
        for (int r = 0; r < d.ny; ++r) acc += c.Cm[r * d.ldC + a] * c.yv[r] * c.Cm[r * d.ldC + k];
        c.Pblk[e] = acc;
    }
    __syncwarp();
    valid = true;
}
// column inf-norm of the (D-scaled, not yet c-scaled) P column k of stage i: max_j D_j |P_jk| D_k
__device__ __forceinline__ double Pcol_norm(Ctx& c, int i, int k, const double* dcur) {
    const Dm& d = c.d;
    if (k < d.nx) {
        double mx = 0;
        for (int j = 0; j < d.nx; ++j) mx = fmax(mx, dcur[j] * fabs(c.Pblk[j * d.nx + k]));
        return mx * dcur[k];
    } else if (k < d.ne) return dcur[k] * dcur[k] * fabs(c.wU(i, k - d.nx));
    return dcur[k] * dcur[k] * fabs(c.wDU(i, k - d.ne));
}

// ---- setup: q, Ruiz equilibration (scaling.c scale_data), scaled bounds, row types ------------------------------
__device__ bool setup_and_scale(Ctx& c) {
    const Dm& d = c.d;
    const int lane = c.lane;
    // unscaled q into qs, D=1, E=1
    for (int i = 0; i <= d.ph; ++i) {
        stage_q_prepare(c, i);
        int bi = d.bcount(i);
        for (int k = lane; k < bi; k += 32) { c.qs[c.voff(i) + k] = stage_q(c, i, k); c.D[c.voff(i) + k] = 1.0; }
        __syncwarp();
    }
    for (int g = lane; g < d.m; g += 32) c.E[g] = 1.0;
    __syncwarp();
    c.c = 1.0;
    double pending_c = 1.0;
    double* Dt = c.va; double* Et = c.ra;
    double* dcur = c.uxc; double* dnxt = c.uxn; double* erow = c.vrowA; double* eprev = c.veqp;
    for (int it = 0; it < c.p.scaling; ++it) {
        // pass A: norms with the current D,E
        bool pv = false;
        for (int i = 0; i <= d.ph; ++i) {
            int bi = d.bcount(i), rs = d.rcount(i), ro = d.roff(i), vo = c.voff(i);
            for (int k = lane; k < bi; k += 32) dcur[k] = c.D[vo + k];
            if (i < d.ph) for (int k = lane; k < d.ne; k += 32) dnxt[k] = c.D[vo + d.b + k];
            for (int r = lane; r < rs; r += 32) erow[r] = c.E[ro + r];
            if (i == 0) for (int r = lane; r < d.ne; r += 32) eprev[r] = c.E[r];
            __syncwarp();
            stage_Pblk(c, i, pv);
            for (int k = lane; k < bi; k += 32) {
                double cn = c.c * Pcol_norm(c, i, k, dcur);
                double dk = dcur[k];
                if (k < d.ne) {
                    cn = fmax(cn, eprev[k] * dk);
                    cn = fmax(cn, erow[d.oBOX + k] * dk);
                    if (k < d.nx) for (int r = 0; r < d.ny; ++r) cn = fmax(cn, erow[d.oOUT + r] * fabs(c.Cm[r * d.ldC + k]) * dk);
                    cn = fmax(cn, erow[d.oSC] * fabs(c.s[k]) * dk);
                } else cn = fmax(cn, erow[d.oDU + k - d.ne] * dk);
                if (i < d.ph) for (int r = 0; r < d.ne; ++r) cn = fmax(cn, erow[d.oEQ + r] * fabs(c.G[r * d.ldG + k]) * dk);
                Dt[vo + k] = 1.0 / sqrt(lim_scaling(cn));
            }
            for (int r = lane; r < rs; r += 32) {
                double e = erow[r], rn;
                if (r < d.oOUT) rn = e * dcur[r];
                else if (r < d.oSC) { rn = 0; int q = r - d.oOUT; for (int k = 0; k < d.nx; ++k) rn = fmax(rn, e * fabs(c.Cm[q * d.ldC + k]) * dcur[k]); }
                else if (r < d.oEQ) { rn = 0; for (int k = 0; k < d.ne; ++k) rn = fmax(rn, e * fabs(c.s[k]) * dcur[k]); }
                else if (r < d.oDU) { int q = r - d.oEQ; rn = e * dnxt[q]; for (int k = 0; k < d.b; ++k) rn = fmax(rn, e * fabs(c.G[q * d.ldG + k]) * dcur[k]); }
                else rn = e * dcur[d.ne + r - d.oDU];
                Et[ro + r] = 1.0 / sqrt(lim_scaling(rn));
            }
            if (i == 0) for (int r = lane; r < d.ne; r += 32) Et[r] = 1.0 / sqrt(lim_scaling(eprev[r] * dcur[r]));
            __syncwarp();
            if (i < d.ph) for (int r = lane; r < d.ne; r += 32) eprev[r] = erow[d.oEQ + r];
            __syncwarp();
        }
        // pass B: apply, accumulate cost-normalisation terms
        double psum = 0, qmax = 0;
        pv = false;
        for (int g = lane; g < d.m; g += 32) c.E[g] *= Et[g];
        for (int i = 0; i <= d.ph; ++i) {
            int bi = d.bcount(i), vo = c.voff(i);
            for (int k = lane; k < bi; k += 32) {
                double dn = c.D[vo + k] * Dt[vo + k];
                c.D[vo + k] = dn; dcur[k] = dn;
                double qv = (c.qs[vo + k] * pending_c) * Dt[vo + k];
                c.qs[vo + k] = qv; qmax = fmax(qmax, fabs(qv));
            }
            __syncwarp();
            stage_Pblk(c, i, pv);
            for (int k = lane; k < bi; k += 32) psum += c.c * Pcol_norm(c, i, k, dcur);
            __syncwarp();
        }
        psum = wsum(psum); qmax = wmax(qmax);
        double ct = psum / (double)d.n;
        double nq = lim_scaling(qmax);
        ct = fmax(ct, nq);
        ct = 1.0 / lim_scaling(ct);
        c.c *= ct; pending_c = ct;
    }
    // finalise: pending q scaling, scaled bounds, row types, validate l<=u
    bool bad = false;
    for (int k = lane; k < d.n; k += 32) c.qs[k] *= pending_c;
    for (int i = 0; i <= d.ph; ++i) {
        int rs = d.rcount(i), ro = d.roff(i);
        for (int r = lane; r < rs; r += 32) {
            double l, u; stage_bounds(c, i, r, l, u);
            bad |= (l > u);
            double e = c.E[ro + r];
            l *= e; u *= e;
            c.lo[ro + r] = l; c.up[ro + r] = u;
            int8_t ty = ((l < -kOsqpInfty * kMinScaling) && (u > kOsqpInfty * kMinScaling)) ? 0 : ((u - l < kRhoTol) ? 2 : 1);
            c.rtype[ro + r] = ty;
        }
    }
    for (int r = lane; r < d.ne; r += 32) {   // eq(0) = -[x0;u0]
        double v = r < d.nx ? -__ldg(c.pr.x0 + (long long)c.inst * d.nx + r) : -__ldg(c.pr.u0 + (long long)c.inst * d.nu + (r - d.nx));
        v *= c.E[r];
        c.lo[r] = v; c.up[r] = v; c.rtype[r] = 2;
    }
    __syncwarp();
    return !wany(bad);
}

__device__ __forceinline__ void set_rho(Ctx& c, double rho) {
    c.rsel[0] = kRhoMin; c.rsel[1] = rho; c.rsel[2] = kRhoEqOverIneq * rho;
    for (int k = 0; k < 3; ++k) c.rinv[k] = 1.0 / c.rsel[k];
}

// ---- block tridiagonal Cholesky of H = D (c P + A' R' A) D + sigma I ---------------------------------------------
// rsel[rtype] gives the row weight; polish passes {0, 1/delta, -} with rtype = activity.  Returns false on a
// non-positive pivot.
__device__ bool factorize(Ctx& c, double sigma) {
    const Dm& d = c.d;
    const int lane = c.lane, ldb = d.ldb;
    double* rw = c.vrowA;      // rho' = rho E^2 of this stage's rows
    double* rwp = c.veqp;      // rho' of eq(i) rows (owned by the previous stage)
    double* dw = c.uxc;        // D of this stage
    double* dn = c.uxn;        // D of e_{i+1}
    bool ok = true;
    for (int i = 0; i <= d.ph; ++i) {
        int bi = d.bcount(i), rs = d.rcount(i), ro = d.roff(i), vo = c.voff(i);
        int bprev = d.b;
        for (int r = lane; r < rs; r += 32) { double e = c.E[ro + r]; rw[r] = c.rsel[c.rtype[ro + r]] * e * e; }
        if (i == 0) for (int r = lane; r < d.ne; r += 32) { double e = c.E[r]; rwp[r] = c.rsel[c.rtype[r]] * e * e; }
        for (int k = lane; k < bi; k += 32) dw[k] = c.D[vo + k];
        if (i < d.ph) for (int k = lane; k < d.ne; k += 32) dn[k] = c.D[vo + d.b + k];
        for (int r = lane; r < d.ny; r += 32) c.yv[r] = 0.0;
        __syncwarp();
        for (int r = lane; r < d.ny; r += 32) c.yv[r] = c.c * c.wO(i, r) + rw[d.oOUT + r];
        __syncwarp();
        // lower triangle of S
        int npairs = bi * (bi + 1) / 2;
        for (int pidx = lane; pidx < npairs; pidx += 32) {
            // unrank (r,k), k<=r
            int r = (int)((sqrt(8.0 * pidx + 1.0) - 1.0) * 0.5);
            while ((r + 1) * (r + 2) / 2 <= pidx) ++r;
            while (r * (r + 1) / 2 > pidx) --r;
            int k = pidx - r * (r + 1) / 2;
            double v = 0;
            if (i < d.ph) for (int j = 0; j < d.ne; ++j) v += c.G[j * d.ldG + r] * rw[d.oEQ + j] * c.G[j * d.ldG + k];
            if (r < d.ne) {
                v += rw[d.oSC] * c.s[r] * c.s[k];
                if (r < d.nx) for (int j = 0; j < d.ny; ++j) v += c.Cm[j * d.ldC + r] * c.yv[j] * c.Cm[j * d.ldC + k];
                if (r == k) {
                    v += rwp[k] + rw[d.oBOX + k];
                    if (k >= d.nx) v += c.c * c.wU(i, k - d.nx);
                }
            } else if (r == k) v += c.c * c.wDU(i, k - d.ne) + rw[d.oDU + k - d.ne];
            v = dw[r] * v * dw[k];
            if (r == k) v += sigma;
            if (i > 0 && r < d.ne) {   // Schur complement of the previous stage: Lc_{i-1} Lc_{i-1}'
                double acc = 0;
                for (int q = 0; q < bprev; ++q) acc += c.Lc[r * ldb + q] * c.Lc[k * ldb + q];
                v -= acc;
            }
            c.S[r * ldb + k] = v;
        }
        __syncwarp();
        // Cholesky (right-looking), in place
        for (int k = 0; k < bi; ++k) {
            double dkk = c.S[k * ldb + k];
            if (!(dkk > 0.0)) ok = false;
            double piv = sqrt(dkk), inv = 1.0 / piv;
            __syncwarp();
            for (int r = k + lane; r < bi; r += 32) c.S[r * ldb + k] = (r == k) ? piv : c.S[r * ldb + k] * inv;
            __syncwarp();
            for (int r = k + 1 + lane; r < bi; r += 32) {
                double lrk = c.S[r * ldb + k];
                for (int q = k + 1; q <= r; ++q) c.S[r * ldb + q] -= lrk * c.S[q * ldb + k];
            }
            __syncwarp();
        }
        // inverse of L: lane = column
        for (int col = lane; col < bi; col += 32) {
            for (int r = 0; r < bi; ++r) {
                double v;
                if (r < col) v = 0.0;
                else if (r == col) v = 1.0 / c.S[r * ldb + r];
                else {
                    double acc = 0;
                    for (int q = col; q < r; ++q) acc += c.S[r * ldb + q] * c.Li[q * ldb + col];
                    v = -acc / c.S[r * ldb + r];
                }
                c.Li[r * ldb + col] = v;
            }
        }
        __syncwarp();
        double* fblk = c.fac + (size_t)i * d.FS;
        for (int pidx = lane; pidx < npairs; pidx += 32) {
            int r = (int)((sqrt(8.0 * pidx + 1.0) - 1.0) * 0.5);
            while ((r + 1) * (r + 2) / 2 <= pidx) ++r;
            while (r * (r + 1) / 2 > pidx) --r;
            int k = pidx - r * (r + 1) / 2;
            fblk[pidx] = c.Li[r * ldb + k];
        }
        if (i < d.ph) {
            // Hc = -(D_e(i+1) rho'_eq(i+1)) G Dw ;  Lc = Hc Li'
            for (int e = lane; e < d.ne * bi; e += 32) {
                int r = e / bi, k = e - r * bi;
                c.Hc[r * ldb + k] = -(dn[r] * rw[d.oEQ + r]) * c.G[r * d.ldG + k] * dw[k];
            }
            __syncwarp();
            for (int e = lane; e < d.ne * bi; e += 32) {
                int r = e / bi, k = e - r * bi;
                double acc = 0;
                for (int q = 0; q <= k; ++q) acc += c.Hc[r * ldb + q] * c.Li[k * ldb + q];
                c.Lc[r * ldb + k] = acc;
                fblk[d.oLc + r * ldb + k] = acc;
            }
            __syncwarp();
            for (int r = lane; r < d.ne; r += 32) rwp[r] = rw[d.oEQ + r];
        }
        __syncwarp();
    }
    return !wany(!ok);
}

// ---- factor block i -> shared memory ring slot ------------------------------------------------------------------
__device__ __forceinline__ void load_fblk(Ctx& c, int i, double* dst) {
    const double* src = c.fac + (size_t)i * c.d.FS;
    const double2* s2 = reinterpret_cast<const double2*>(src);
    double2* d2 = reinterpret_cast<double2*>(dst);
    int n2 = c.d.FS >> 1;
    for (int e = c.lane; e < n2; e += 32) d2[e] = s2[e];
}

// ---- one reduced-KKT solve fused with the ADMM updates (MODE 0) or with the polish bookkeeping (MODE 1) -----------
//  MODE 0: rhs = sigma x - q + A'(rho z - y);  x~ = H^-1 rhs;  then x,z,y updates of osqp.c (update_x/z/y)
//  MODE 1: rhs = r1 + A'(w r2) (w = act/delta); dx = H^-1 rhs; px += dx; pnu += w (A dx - r2)      [r1=va, r2=rc, pnu=rb]
template <int MODE>
__device__ void kkt_sweeps(Ctx& c, bool store_delta, bool first) {
    const Dm& d = c.d;
    const int lane = c.lane, ldb = d.ldb;
    const double sigma = c.p.sigma, alpha = c.p.alpha;
    double* vrow = c.vrowA; double* veqp = c.veqp;
    double* tprev = c.tA; double* tcur = c.tB;
    // ---------------- forward ----------------
    for (int r = lane; r < d.ne; r += 32) {
        int ty = c.rtype[r];
        veqp[r] = MODE == 0 ? c.E[r] * (c.rsel[ty] * c.z[r] - c.y[r]) : c.E[r] * (c.rsel[ty] * c.rc[r]);
    }
    for (int i = 0; i <= d.ph; ++i) {
        int bi = d.bcount(i), rs = d.rcount(i), ro = d.roff(i), vo = c.voff(i);
        double* fcur = (i & 1) ? c.fb1 : c.fb0;
        double* fprv = (i & 1) ? c.fb0 : c.fb1;
        load_fblk(c, i, fcur);
        for (int r = lane; r < rs; r += 32) {
            int g = ro + r; int ty = c.rtype[g];
            vrow[r] = MODE == 0 ? c.E[g] * (c.rsel[ty] * c.z[g] - c.y[g]) : c.E[g] * (c.rsel[ty] * c.rc[g]);
        }
        __syncwarp();
        for (int k = lane; k < bi; k += 32) {
            double au;
            if (k < d.ne) {
                au = vrow[d.oBOX + k] - veqp[k] + c.s[k] * vrow[d.oSC];
                if (k < d.nx) for (int j = 0; j < d.ny; ++j) au += c.Cm[j * d.ldC + k] * vrow[d.oOUT + j];
            } else au = vrow[d.oDU + k - d.ne];
            if (i < d.ph) for (int j = 0; j < d.ne; ++j) au += c.G[j * d.ldG + k] * vrow[d.oEQ + j];
            double rhs = MODE == 0 ? (sigma * c.x[vo + k] - c.qs[vo + k] + c.D[vo + k] * au) : (c.va[vo + k] + c.D[vo + k] * au);
            if (i > 0 && k < d.ne) {
                const double* Lcp = fprv + d.oLc + k * ldb;
                double acc = 0;
                for (int q = 0; q < d.b; ++q) acc += Lcp[q] * tprev[q];
                rhs -= acc;
            }
            c.vtmp[k] = rhs;
        }
        __syncwarp();
        for (int k = lane; k < bi; k += 32) {
            const double* Lr = fcur + k * (k + 1) / 2;
            double acc = 0;
            for (int q = 0; q <= k; ++q) acc += Lr[q] * c.vtmp[q];
            tcur[k] = acc; c.t[vo + k] = acc;
        }
        if (i < d.ph) for (int r = lane; r < d.ne; r += 32) veqp[r] = vrow[d.oEQ + r];
        __syncwarp();
        double* tt = tprev; tprev = tcur; tcur = tt;
    }
    // ---------------- backward ----------------
    double* uxc = c.uxc; double* uxn = c.uxn; double* xn = c.xn;
    for (int i = d.ph; i >= 0; --i) {
        int bi = d.bcount(i), ro = d.roff(i), vo = c.voff(i);
        double* fcur = (i & 1) ? c.fb1 : c.fb0;
        if (i != d.ph) load_fblk(c, i, fcur);   // block ph is still resident from the forward sweep
        __syncwarp();
        for (int k = lane; k < bi; k += 32) {
            double w = c.t[vo + k];
            if (i < d.ph) {
                const double* Lcc = fcur + d.oLc + k;
                double acc = 0;
                for (int r = 0; r < d.ne; ++r) acc += Lcc[r * ldb] * xn[r];
                w -= acc;
            }
            c.vtmp[k] = w;
        }
        __syncwarp();
        for (int k = lane; k < bi; k += 32) {
            double acc = 0;
            for (int r = k; r < bi; ++r) acc += fcur[r * (r + 1) / 2 + k] * c.vtmp[r];
            double xt = acc;
            if (MODE == 0) {
                double xo = c.x[vo + k];
                double xnew = alpha * xt + (1.0 - alpha) * xo;
                c.x[vo + k] = xnew;
                if (store_delta) c.va[vo + k] = xnew - xo;
            } else {
                c.px[vo + k] = first ? xt : c.px[vo + k] + xt;
            }
            uxc[k] = c.D[vo + k] * xt;
            c.tA[k] = xt;     // scaled x~ of this stage (tA/tB are free during the backward sweep)
        }
        __syncwarp();
        // rows owned by stage i
        auto row_update = [&](int g, double a) {
            int ty = c.rtype[g];
            if (MODE == 0) {
                double zt = c.E[g] * a;
                double zo = c.z[g];
                double zr = alpha * zt + (1.0 - alpha) * zo;
                double yo = c.y[g];
                double zn = fmin(fmax(zr + c.rinv[ty] * yo, c.lo[g]), c.up[g]);
                double dy = c.rsel[ty] * (zr - zn);
                c.y[g] = yo + dy; c.z[g] = zn;
                if (store_delta) c.ra[g] = dy;
            } else {
                double dnu = c.rsel[ty] * (c.E[g] * a - c.rc[g]);
                c.rb[g] = first ? dnu : c.rb[g] + dnu;
            }
        };
        for (int r = lane; r < d.ne; r += 32) row_update(ro + d.oBOX + r, uxc[r]);
        for (int r = lane; r < d.ny; r += 32) {
            double a = 0;
            for (int k = 0; k < d.nx; ++k) a += c.Cm[r * d.ldC + k] * uxc[k];
            row_update(ro + d.oOUT + r, a);
        }
        if (lane == 0) {
            double a = 0;
            for (int k = 0; k < d.ne; ++k) a += c.s[k] * uxc[k];
            row_update(ro + d.oSC, a);
        }
        if (i < d.ph) {
            for (int r = lane; r < d.ne; r += 32) {
                double a = -uxn[r];
                for (int k = 0; k < d.b; ++k) a += c.G[r * d.ldG + k] * uxc[k];
                row_update(ro + d.oEQ + r, a);
            }
            for (int r = lane; r < d.nu; r += 32) row_update(ro + d.oDU + r, uxc[d.ne + r]);
        }
        if (i == 0) for (int r = lane; r < d.ne; r += 32) row_update(r, -uxc[r]);
        __syncwarp();
        for (int r = lane; r < d.ne; r += 32) { uxn[r] = uxc[r]; xn[r] = c.tA[r]; }
        __syncwarp();
    }
}

// ---- generic structured products --------------------------------------------------------------------------------
// rows_pass: for every row g:  rowfn(g, a_g . ux)  with ux = D*xsrc (unscaled variable values)
template <class RowFn>
__device__ void rows_pass(Ctx& c, const double* xsrc, RowFn rowfn) {
    const Dm& d = c.d;
    const int lane = c.lane;
    double* uxc = c.uxc; double* uxn = c.uxn;
    for (int k = lane; k < d.bcount(0); k += 32) uxc[k] = c.D[k] * xsrc[k];
    __syncwarp();
    for (int i = 0; i <= d.ph; ++i) {
        int ro = d.roff(i), vo = c.voff(i);
        if (i < d.ph) for (int k = lane; k < d.bcount(i + 1); k += 32) uxn[k] = c.D[vo + d.b + k] * xsrc[vo + d.b + k];
        __syncwarp();
        for (int r = lane; r < d.ne; r += 32) rowfn(ro + d.oBOX + r, uxc[r]);
        for (int r = lane; r < d.ny; r += 32) {
            double a = 0;
            for (int k = 0; k < d.nx; ++k) a += c.Cm[r * d.ldC + k] * uxc[k];
            rowfn(ro + d.oOUT + r, a);
        }
        if (lane == 0) {
            double a = 0;
            for (int k = 0; k < d.ne; ++k) a += c.s[k] * uxc[k];
            rowfn(ro + d.oSC, a);
        }
        if (i < d.ph) {
            for (int r = lane; r < d.ne; r += 32) {
                double a = -uxn[r];
                for (int k = 0; k < d.b; ++k) a += c.G[r * d.ldG + k] * uxc[k];
                rowfn(ro + d.oEQ + r, a);
            }
            for (int r = lane; r < d.nu; r += 32) rowfn(ro + d.oDU + r, uxc[d.ne + r]);
        }
        if (i == 0) for (int r = lane; r < d.ne; r += 32) rowfn(r, -uxc[r]);
        __syncwarp();
        double* tt = uxc; uxc = uxn; uxn = tt;
    }
}
// cols_pass: for every variable kg: colfn(kg, D_k * sum_r a_r[k] v_r, c*D_k*(P ux)_k) with v_r = rowval(g) (already
// E-weighted) and ux = D*xsrc (xsrc may be null when WITHP is false)
template <bool WITHP, class RowVal, class ColFn>
__device__ void cols_pass(Ctx& c, const double* xsrc, RowVal rowval, ColFn colfn) {
    const Dm& d = c.d;
    const int lane = c.lane;
    double* vrow = c.vrowA; double* veqp = c.veqp; double* uxc = c.uxc;
    for (int r = lane; r < d.ne; r += 32) veqp[r] = rowval(r);
    for (int i = 0; i <= d.ph; ++i) {
        int bi = d.bcount(i), rs = d.rcount(i), ro = d.roff(i), vo = c.voff(i);
        for (int r = lane; r < rs; r += 32) vrow[r] = rowval(ro + r);
        if (WITHP) for (int k = lane; k < bi; k += 32) uxc[k] = c.D[vo + k] * xsrc[vo + k];
        __syncwarp();
        if (WITHP) {
            for (int r = lane; r < d.ny; r += 32) {
                double a = 0;
                for (int k = 0; k < d.nx; ++k) a += c.Cm[r * d.ldC + k] * uxc[k];
                c.yv[r] = c.wO(i, r) * a;
            }
            __syncwarp();
        }
        for (int k = lane; k < bi; k += 32) {
            double au, pu = 0;
            if (k < d.ne) {
                au = vrow[d.oBOX + k] - veqp[k] + c.s[k] * vrow[d.oSC];
                if (k < d.nx) for (int j = 0; j < d.ny; ++j) au += c.Cm[j * d.ldC + k] * vrow[d.oOUT + j];
            } else au = vrow[d.oDU + k - d.ne];
            if (i < d.ph) for (int j = 0; j < d.ne; ++j) au += c.G[j * d.ldG + k] * vrow[d.oEQ + j];
            if (WITHP) {
                if (k < d.nx) for (int j = 0; j < d.ny; ++j) pu += c.Cm[j * d.ldC + k] * c.yv[j];
                else if (k < d.ne) pu = c.wU(i, k - d.nx) * uxc[k];
                else pu = c.wDU(i, k - d.ne) * uxc[k];
                pu *= c.c * c.D[vo + k];
            }
            colfn(vo + k, c.D[vo + k] * au, pu);
        }
        __syncwarp();
        if (i < d.ph) for (int r = lane; r < d.ne; r += 32) veqp[r] = vrow[d.oEQ + r];
        __syncwarp();
    }
}

// update_info (auxil.c): residuals + every norm the termination test and the rho estimate need.
// zy(g, Ax, z, y) supplies the (z,y) pair of row g (ADMM iterate or the polished pair).
template <class ZY>
__device__ InfoNorms info_pass(Ctx& c, const double* xsrc, ZY zy) {
    InfoNorms I;
    double pri = 0, nz = 0, nAx = 0, spri = 0, snz = 0, snAx = 0;
    rows_pass(c, xsrc, [&](int g, double a) {
        double e = c.E[g], einv = 1.0 / e;
        double Ax = e * a, z, y;
        zy(g, Ax, z, y);
        c.rc[g] = e * y;            // E-weighted dual for the column pass
        double pv = Ax - z;
        spri = fmax(spri, fabs(pv)); snz = fmax(snz, fabs(z)); snAx = fmax(snAx, fabs(Ax));
        pri = fmax(pri, fabs(einv * pv)); nz = fmax(nz, fabs(einv * z)); nAx = fmax(nAx, fabs(einv * Ax));
    });
    __syncwarp();
    double dua = 0, nq = 0, nAty = 0, nPx = 0, sdua = 0, snq = 0, snAty = 0, snPx = 0, xPx = 0, qx = 0;
    cols_pass<true>(c, xsrc, [&](int g) { return c.rc[g]; }, [&](int kg, double aty, double px) {
        double dinv = 1.0 / c.D[kg];
        double q = c.qs[kg];
        double dv = q + px + aty;
        sdua = fmax(sdua, fabs(dv)); snq = fmax(snq, fabs(q)); snAty = fmax(snAty, fabs(aty)); snPx = fmax(snPx, fabs(px));
        dua = fmax(dua, fabs(dinv * dv)); nq = fmax(nq, fabs(dinv * q)); nAty = fmax(nAty, fabs(dinv * aty)); nPx = fmax(nPx, fabs(dinv * px));
        double xv = xsrc[kg];
        xPx += xv * px; qx += q * xv;
    });
    double cinv = 1.0 / c.c;
    I.pri = wmax(pri); I.nz = wmax(nz); I.nAx = wmax(nAx); I.spri = wmax(spri); I.snz = wmax(snz); I.snAx = wmax(snAx);
    I.dua = cinv * wmax(dua); I.nq = wmax(nq); I.nAty = wmax(nAty); I.nPx = wmax(nPx);
    I.sdua = wmax(sdua); I.snq = wmax(snq); I.snAty = wmax(snAty); I.snPx = wmax(snPx);
    I.xPx = wsum(xPx); I.qx = wsum(qx);
    return I;
}

// is_primal_infeasible (auxil.c); delta_y lives in ra
__device__ bool primal_infeasible(Ctx& c, double eps) {
    const Dm& d = c.d;
    double nd = 0, lhs = 0;
    for (int g = c.lane; g < d.m; g += 32) {
        double l = c.lo[g], u = c.up[g], dy = c.ra[g];
        if (u > kOsqpInfty * kMinScaling) {
            if (l < -kOsqpInfty * kMinScaling) dy = 0.0; else dy = fmin(dy, 0.0);
        } else if (l < -kOsqpInfty * kMinScaling) dy = fmax(dy, 0.0);
        c.ra[g] = dy;
        nd = fmax(nd, fabs(c.E[g] * dy));
        lhs += u * fmax(dy, 0.0) + l * fmin(dy, 0.0);   // IEEE: inf*0 = NaN, exactly as in the reference build
    }
    nd = wmax(nd); lhs = wsum(lhs);
    __syncwarp();
    if (nd > eps) {
        if (lhs < -eps * nd) {
            double mx = 0;
            cols_pass<false>(c, nullptr, [&](int g) { return c.E[g] * c.ra[g]; },
                             [&](int kg, double aty, double) { mx = fmax(mx, fabs(aty / c.D[kg])); });
            mx = wmax(mx);
            return mx < eps * nd;
        }
    }
    return false;
}
// is_dual_infeasible (auxil.c); delta_x lives in va
__device__ bool dual_infeasible(Ctx& c, double eps) {
    const Dm& d = c.d;
    double nd = 0, qd = 0;
    for (int k = c.lane; k < d.n; k += 32) { double dx = c.va[k]; nd = fmax(nd, fabs(c.D[k] * dx)); qd += c.qs[k] * dx; }
    nd = wmax(nd); qd = wsum(qd);
    double cs = c.c;
    if (nd > eps) {
        if (qd < -cs * eps * nd) {
            double mx = 0;
            cols_pass<true>(c, c.va, [&](int) { return 0.0; },
                            [&](int kg, double, double px) { mx = fmax(mx, fabs(px / c.D[kg])); });
            mx = wmax(mx);
            if (mx < cs * eps * nd) {
                bool bad = false;
                rows_pass(c, c.va, [&](int g, double a) {
                    double adx = a;   // Einv * (E a) = a
                    if (((c.up[g] < kOsqpInfty * kMinScaling) && (adx > eps * nd)) ||
                        ((c.lo[g] > -kOsqpInfty * kMinScaling) && (adx < -eps * nd))) bad = true;
                });
                return !wany(bad);
            }
        }
    }
    return false;
}

// check_termination (auxil.c).  Returns true when the loop must stop; status/obj updated.
__device__ bool check_termination(Ctx& c, const InfoNorms& I, bool approximate, int& status, double& obj) {
    double eps_abs = c.p.eps_abs, eps_rel = c.p.eps_rel, epi = c.p.eps_prim_inf, edi = c.p.eps_dual_inf;
    if (I.pri > kOsqpInfty || I.dua > kOsqpInfty) { status = OSQP_NON_CVX; obj = NAN; return true; }
    if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    double eps_prim = eps_abs + eps_rel * fmax(I.nz, I.nAx);
    if (I.pri < eps_prim) prim_ok = true; else prim_inf = primal_infeasible(c, epi);
    double eps_dual = eps_abs + eps_rel * (1.0 / c.c) * fmax(fmax(I.nq, I.nAty), I.nPx);
    if (I.dua < eps_dual) dual_ok = true; else dual_inf = dual_infeasible(c, edi);
    if (prim_ok && dual_ok) { status = approximate ? OSQP_SOLVED_INACCURATE : OSQP_SOLVED; return true; }
    if (prim_inf) { status = approximate ? OSQP_PRIMAL_INFEASIBLE_INACCURATE : OSQP_PRIMAL_INFEASIBLE; obj = kOsqpInfty; return true; }
    if (dual_inf) { status = approximate ? OSQP_DUAL_INFEASIBLE_INACCURATE : OSQP_DUAL_INFEASIBLE; obj = -kOsqpInfty; return true; }
    return false;
}

__device__ __forceinline__ int to_result_status(int st) {   // LOptimizer.hpp:386-415
    switch (st) {
    case OSQP_SOLVED: return RS_SUCCESS;
    case OSQP_MAX_ITER_REACHED: return RS_MAX_ITERATION;
    case OSQP_PRIMAL_INFEASIBLE: case OSQP_DUAL_INFEASIBLE: return RS_INFEASIBLE;
    case OSQP_SOLVED_INACCURATE: case OSQP_PRIMAL_INFEASIBLE_INACCURATE: case OSQP_DUAL_INFEASIBLE_INACCURATE: return RS_SUCCESS;
    case OSQP_NON_CVX: return RS_ERROR;
    default: return RS_UNKNOWN;
    }
}

// reference row index of internal row g / reference variable index of internal variable kg
__device__ __forceinline__ int ref_row(const Dm& d, int g) {
    if (g < d.ne) return g;
    int i = (g - d.ne) / d.RS, r = (g - d.ne) - i * d.RS;
    if (r < d.oOUT) return d.M0 + i * d.ne + r;
    if (r < d.oSC) return d.M1 + i * d.ny + (r - d.oOUT);
    if (r < d.oEQ) return d.M3 + i;
    if (r < d.oDU) return (i + 1) * d.ne + (r - d.oEQ);
    return d.M2 + i * d.nu + (r - d.oDU);
}
__device__ __forceinline__ int ref_var(const Dm& d, int kg) {
    int i = kg / d.b, k = kg - i * d.b;
    return k < d.ne ? i * d.ne + k : (d.ph + 1) * d.ne + i * d.nu + (k - d.ne);
}
__device__ __forceinline__ int int_var_e(const Dm& d, int i, int k) { return i * d.b + k; }

// ---- the whole LOptimizer::run for one instance -----------------------------------------------------------------
__device__ void solve_instance(Ctx& c, const Out& o) {
    const Dm& d = c.d;
    const int lane = c.lane, inst = c.inst;
    load_model(c);
    bool valid = setup_and_scale(c);
    int status = OSQP_UNSOLVED; double obj = 0; int iters = 0, rho_updates = 0, status_polish = 0;
    double rho = fmin(fmax(c.p.rho, kRhoMin), kRhoMax);
    set_rho(c, rho);
    if (valid) valid = factorize(c, c.p.sigma);
    if (!valid) {
        // osqp_setup would have failed (validate_data / factorisation): LOptimizer.hpp:348-361 failure semantics
        for (int k = lane; k < d.nu; k += 32) o.cmd[(long long)inst * d.nu + k] = o.prev_cmd[(long long)inst * d.nu + k];
        if (lane == 0) { o.cost[inst] = INFINITY; o.status[inst] = RS_ERROR; o.solver_status[inst] = B200_SETUP_ERROR;
                         o.feasible[inst] = 0; o.iters[inst] = 0; o.rho_updates[inst] = 0; o.polish[inst] = 0; }
        if (o.seq_state) for (int e = lane; e < (d.ph + 1) * d.nx; e += 32) o.seq_state[(long long)inst * (d.ph + 1) * d.nx + e] = 0;
        if (o.seq_input) for (int e = lane; e < (d.ph + 1) * d.nu; e += 32) o.seq_input[(long long)inst * (d.ph + 1) * d.nu + e] = 0;
        if (o.seq_output) for (int e = lane; e < (d.ph + 1) * d.ny; e += 32) o.seq_output[(long long)inst * (d.ph + 1) * d.ny + e] = 0;
        return;
    }
    // cold / warm start (osqp_warm_start: x <- Dinv x, y <- c Einv y, z <- A x)
    if (c.pr.warm) {
        for (int kg = lane; kg < d.n; kg += 32) c.x[kg] = __ldg(c.pr.warm_x + (long long)inst * d.n + ref_var(d, kg)) / c.D[kg];
        for (int g = lane; g < d.m; g += 32) c.y[g] = c.c * (__ldg(c.pr.warm_y + (long long)inst * d.m + ref_row(d, g)) / c.E[g]);
        __syncwarp();
        rows_pass(c, c.x, [&](int g, double a) { c.z[g] = c.E[g] * a; });
    } else {
        for (int k = lane; k < d.n; k += 32) c.x[k] = 0.0;
        for (int g = lane; g < d.m; g += 32) { c.z[g] = 0.0; c.y[g] = 0.0; }
    }
    __syncwarp();
    InfoNorms I; I.pri = I.dua = 0;
    bool can_check = false, done = false;
    auto admm_zy = [&](int g, double, double& z, double& y) { z = c.z[g]; y = c.y[g]; };
    int it = 1;
    for (; it <= c.p.max_iter; ++it) {
        can_check = c.p.check_termination && (it % c.p.check_termination == 0);
        bool can_adapt = c.p.adaptive_rho && c.p.adaptive_rho_interval && (it % c.p.adaptive_rho_interval == 0);
        kkt_sweeps<0>(c, can_check || it == c.p.max_iter, false);
        if (can_check || can_adapt) {
            I = info_pass(c, c.x, admm_zy);
            if (can_check && check_termination(c, I, false, status, obj)) { done = true; break; }
        }
        if (can_adapt) {
            // compute_rho_estimate (auxil.c) on the SCALED residuals
            double pr = I.spri / (fmax(I.snz, I.snAx) + 1e-10);
            double dr = I.sdua / (fmax(fmax(I.snq, I.snAty), I.snPx) + 1e-10);
            double est = rho * sqrt(pr / (dr + 1e-10));
            est = fmin(fmax(est, kRhoMin), kRhoMax);
            if (est > rho * c.p.adaptive_rho_tolerance || est < rho / c.p.adaptive_rho_tolerance) {
                rho = est; set_rho(c, rho);
                factorize(c, c.p.sigma);
                ++rho_updates;
            }
        }
    }
    iters = done ? it : c.p.max_iter;
    if (!done && !can_check) {
        I = info_pass(c, c.x, admm_zy);
        check_termination(c, I, false, status, obj);
    }
    bool has_solution = !(status == OSQP_PRIMAL_INFEASIBLE || status == OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
                          status == OSQP_DUAL_INFEASIBLE || status == OSQP_DUAL_INFEASIBLE_INACCURATE || status == OSQP_NON_CVX);
    if (has_solution) {
        obj = (0.5 * I.xPx + I.qx) / c.c;
    }
    if (status == OSQP_UNSOLVED) {
        if (!check_termination(c, I, true, status, obj)) status = OSQP_MAX_ITER_REACHED;
    }
    // ---------------- polish (polish.c) ----------------
    if (c.p.polish && status == OSQP_SOLVED) {
        const double dinv = 1.0 / c.p.delta;
        for (int g = lane; g < d.m; g += 32) {
            double z = c.z[g], y = c.y[g], l = c.lo[g], u = c.up[g];
            bool low = (z - l) < -y;
            bool upp = !low && ((u - z) < y);
            c.rtype[g] = (low || upp) ? 1 : 0;
            c.ra[g] = low ? l : (upp ? u : 0.0);     // b_act
        }
        c.rsel[0] = 0.0; c.rsel[1] = dinv; c.rsel[2] = 0.0;
        __syncwarp();
        if (factorize(c, c.p.delta)) {
            for (int k = lane; k < d.n; k += 32) c.va[k] = -c.qs[k];
            for (int g = lane; g < d.m; g += 32) c.rc[g] = c.ra[g];
            __syncwarp();
            kkt_sweeps<1>(c, false, true);
            for (int rf = 0; rf < c.p.polish_refine_iter; ++rf) {
                // r1 = -q - P px - A' pnu ; r2 = act (b - A px)
                cols_pass<true>(c, c.px, [&](int g) { return c.E[g] * c.rb[g]; },
                                [&](int kg, double aty, double pxv) { c.va[kg] = -c.qs[kg] - pxv - aty; });
                rows_pass(c, c.px, [&](int g, double a) { c.rc[g] = c.rtype[g] ? (c.ra[g] - c.E[g] * a) : 0.0; });
                __syncwarp();
                kkt_sweeps<1>(c, false, false);
            }
            // polished (z,y): z = A px, project_normalcone
            auto pol_zy = [&](int g, double Ax, double& z, double& y) {
                double t = Ax + c.rb[g];
                z = fmin(fmax(t, c.lo[g]), c.up[g]);
                y = t - z;
            };
            // info_pass overwrites rc (free now); keep pnu in rb
            InfoNorms P = info_pass(c, c.px, pol_zy);
            bool okp = (P.pri < I.pri && P.dua < I.dua) || (P.pri < I.pri && I.dua < 1e-10) || (P.dua < I.dua && I.pri < 1e-10);
            if (okp) {
                obj = (0.5 * P.xPx + P.qx) / c.c;
                status_polish = 1;
                for (int k = lane; k < d.n; k += 32) c.x[k] = c.px[k];
                rows_pass(c, c.px, [&](int g, double a) {
                    double z, y; pol_zy(g, c.E[g] * a, z, y);
                    c.z[g] = z; c.y[g] = y;
                });
            } else status_polish = -1;
        } else status_polish = -1;
        __syncwarp();
    }
    // ---------------- store_solution + LOptimizer unpack ----------------
    const double cinv = 1.0 / c.c;
    const long long ib = inst;
    for (int kg = lane; kg < d.n; kg += 32) {
        double xv = has_solution ? c.D[kg] * c.x[kg] : NAN;
        c.x[kg] = xv;   // unscaled from here on
        if (o.sol_x) o.sol_x[ib * d.n + ref_var(d, kg)] = xv;
    }
    if (o.sol_y) for (int g = lane; g < d.m; g += 32) o.sol_y[ib * d.m + ref_row(d, g)] = has_solution ? cinv * c.E[g] * c.y[g] : NAN;
    __syncwarp();
    for (int k = lane; k < d.nu; k += 32) {
        int st = d.ph >= 1 ? 1 : 0;    // sequence.input.row(0) = x_u(1)  (LOptimizer.hpp:316-327,341)
        double v = c.x[int_var_e(d, st, d.nx + k)];
        o.cmd[ib * d.nu + k] = v; o.prev_cmd[ib * d.nu + k] = v;
    }
    if (o.seq_state) for (int e = lane; e < (d.ph + 1) * d.nx; e += 32) {
        int i = e / d.nx, k = e - i * d.nx;
        o.seq_state[ib * (d.ph + 1) * d.nx + e] = c.x[int_var_e(d, i, k)];
    }
    if (o.seq_input) for (int e = lane; e < (d.ph + 1) * d.nu; e += 32) {
        int i = e / d.nu, k = e - i * d.nu;
        int st = (i + 1 < d.ph + 1) ? i + 1 : i;
        o.seq_input[ib * (d.ph + 1) * d.nu + e] = c.x[int_var_e(d, st, d.nx + k)];
    }
    if (o.seq_output) for (int e = lane; e < (d.ph + 1) * d.ny; e += 32) {
        int i = e / d.ny, r = e - i * d.ny;
        int j = i > 0 ? i - 1 : 0;
        double acc = 0;
        for (int k = 0; k < d.nx; ++k) acc += c.Cm[r * d.ldC + k] * c.x[int_var_e(d, i, k)];
        for (int q = 0; q < d.ndu; ++q) acc += ldp(c.pr.Dd, inst, r * d.ndu + q) * ldp(c.pr.uMeas, inst, j * d.ndu + q);
        o.seq_output[ib * (d.ph + 1) * d.ny + e] = acc;
    }
    if (lane == 0) {
        o.cost[inst] = obj; o.solver_status[inst] = status; o.status[inst] = to_result_status(status);
        o.feasible[inst] = (status == OSQP_SOLVED || status == OSQP_SOLVED_INACCURATE || status == OSQP_MAX_ITER_REACHED) ? 1 : 0;
        o.iters[inst] = iters; o.rho_updates[inst] = rho_updates; o.polish[inst] = status_polish;
    }
}

// ---- persistent kernel: warps pull instances from a global counter ------------------------------------------------
__global__ void __launch_bounds__(256) lmpc_solve_kernel(const __grid_constant__ Dm d, const __grid_constant__ Params p,
                                                        const __grid_constant__ Prob pr, const __grid_constant__ Out o,
                                                        int batch, double* workspace, size_t ws_stride, int* counter) {
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int slot = blockIdx.x * wpb + warp;
    Ctx c(d, p, pr);
    c.lane = lane;
    double* sm = smem + (size_t)warp * d.smem_doubles();
    c.G = sm; sm += d.ne * d.ldG;
    c.Cm = sm; sm += d.ny * d.ldC;
    c.s = sm; sm += d.ne;
    c.uxc = sm; sm += d.b; c.uxn = sm; sm += d.b; c.vtmp = sm; sm += d.b; c.tA = sm; sm += d.b; c.tB = sm; sm += d.b;
    c.xn = sm; sm += d.ne; c.veqp = sm; sm += d.ne;
    c.vrowA = sm; sm += d.RS;
    c.yv = sm; sm += d.ny;
    sm = (double*)(((uintptr_t)sm + 15) & ~(uintptr_t)15);
    // union: factor scratch | factor ring
    c.fb0 = sm; c.fb1 = sm + d.FS;
    c.S = sm; c.Li = c.S + d.b * d.ldb; c.Hc = c.Li + d.b * d.ldb; c.Lc = c.Hc + d.ne * d.ldb; c.Pblk = c.Lc + d.ne * d.ldb;
    double* ws = workspace + (size_t)slot * ws_stride;
    c.D = ws; ws += d.n; c.qs = ws; ws += d.n; c.x = ws; ws += d.n; c.t = ws; ws += d.n; c.va = ws; ws += d.n; c.px = ws; ws += d.n;
    c.E = ws; ws += d.m; c.lo = ws; ws += d.m; c.up = ws; ws += d.m; c.z = ws; ws += d.m; c.y = ws; ws += d.m;
    c.ra = ws; ws += d.m; c.rb = ws; ws += d.m; c.rc = ws; ws += d.m;
    c.fac = ws; ws += (size_t)(d.ph + 1) * d.FS;
    c.rtype = reinterpret_cast<int8_t*>(ws);
    for (;;) {
        int inst = 0;
        if (lane == 0) inst = atomicAdd(counter, 1);
        inst = __shfl_sync(0xffffffffu, inst, 0);
        if (inst >= batch) break;
        c.inst = inst;
        solve_instance(c, o);
        __syncwarp();
    }
}

}  // namespace b200mpc
