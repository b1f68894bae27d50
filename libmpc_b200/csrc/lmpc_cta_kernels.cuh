// lmpc_cta_kernels.cuh -- batched linear-MPC solve for sm_100a, engine 2: ONE CTA PER CONTROLLER, everything on chip.
//
// Same arithmetic contract as lmpc_kernels.cuh (the reference's LOptimizer::run driving OSQP v0.6.3 on the stage structure of
// ProblemBuilder's QP; reference anchors are listed there and in include/b200mpc.h) -- what changes is the mapping:
//
//   * One thread block works on one controller at a time and a persistent grid of one block per SM draws controllers from a
//     global counter.  Only ~148 controllers are in flight, so ALL of a controller's state -- the block-tridiagonal factor
//     (92 KB at quadrotor ph=20), the scaled problem vectors and the ADMM iterates (89 KB) -- lives in shared memory for the
//     whole solve.  Nothing streams: HBM sees the problem data once and the results once (the algorithmic bytes of SURVEY 8d).
//   * The reduced KKT solve is split into stage-PARALLEL mat-vecs (all warps, all stages at once) around a short stage-SERIAL
//     recurrence.  With L_i the Cholesky factor of the i-th pivot block, Lc_i the coupling block and W_i = Lc_i L_i^-1:
//         forward   rhat_i = L_i^-1 r_i,  g_i = W_i r_i                               (parallel)
//                   c_{i+1} = g_i - W_i[:, :ne] c_i                                     (serial: one ne x ne mat-vec per stage)
//                   t_i = rhat_i - L_i^-1[:, :ne] c_i                                   (parallel)
//         backward  s_i = L_i^-T t_i                                                    (parallel)
//                   xe_i = s_i[:ne] - W_i[:, :ne]^T xe_{i+1}                             (serial: one ne x ne mat-vec per stage)
//                   x_i = [xe_i ; s_i[ne:] - W_i[:, ne:]^T xe_{i+1}]                     (parallel)
//     which is algebraically the block forward / backward substitution of lmpc_kernels.cuh (tests/structured_reference.py)
//     with the two dependent mat-vecs per stage fused into one.
//   * L_i^-1 and W_i are stored as zero-padded rectangles with an even leading dimension whose half is odd: every dot product
//     has a compile-time length (no triangular predicates), rows are read with 128-bit loads and both the row walk and the
//     column walk are bank-conflict free.
//   * The factorisation of a stage is a register-resident LDL' by ONE warp (lane r = row r of the pivot block, lane q = column
//     q of the accumulated inverse; the pivot column is the only thing that goes through shared memory), followed by
//     stage-parallel dot products for Lc_i, the Schur complement of stage i+1 and W_i.
//   * The ADMM iteration is bulk-synchronous: the stage-parallel phases (right-hand side, L^-1 r / W r, forward substitution,
//     L^-T products, x update, row updates) are flat task lists over ALL stages spread over all threads; the two serial
//     recurrences run on warp 0 with two lanes per row (half a dot product each, one shuffle), the carried vector exchanged
//     through shared memory and the next stage's K half-row loaded before the step's barrier.  (A warp-specialised pipeline
//     ordered by shared-memory progress flags was measured 2x slower -- profiles/r02_engine_probe.jsonl -- and removed.)
//   * A x / A' y / P x never touch a matrix: rows and columns are evaluated from A, B, C, the weights and the scalar row.
//
// Used for batch sizes from 1 (a single mpc::LMPC<> object: the whole SM works on it) to any; the warp-per-controller engine
// of lmpc_kernels.cuh remains for problems whose vectors do not fit shared memory.
#pragma once
#include "lmpc_kernels.cuh"
#include <type_traits>

namespace b200mpc {

constexpr int OSQP_TIME_LIMIT_REACHED = -6;   // constants.h of OSQP v0.6.3; LOptimizer.hpp:386-415 maps it to UNKNOWN

struct CtaLayout {
    // offsets (doubles) into dynamic shared memory
    int oG, oC, oSV, oW, oD, oQ, oX, oT, oR, oE, oLO, oUP, oZ, oY, oV, oRT, oGT, oCAR, oYV, oRED, oAUG, oM, oLCS, oTB, oTAB, oXE, oFLAG, oFAC;
    int ldG, ldC, LD;    // leading dimensions of G (ne x b), C (ny x nx) and of the factor blocks: even, half odd
    int oWm, FS2;        // factor record of a stage: L^-1 (b x LD, zero above the diagonal) then W (ne x LD)
    int nP;              // padded length of the stage-major variable vectors, (ph+1) b (the last stage has no du block)
    int lda;             // leading dimension of the elimination matrix of the run-time-dimension path
    int n_ent, n_aug;    // elimination table of the run-time-dimension path
    int total;           // doubles of dynamic shared memory
    int fac_shared;      // factor records in shared memory (else in the CTA's global scratch)
    // global scratch of a CTA (doubles)
    int gVA, gPX, gRA, gRB, gRC;
    long long gFAC, gtotal;
};

// even leading dimension >= n whose half is odd: 16-byte aligned rows for 128-bit loads, and both a walk down the rows
// (stride ld) and across a row touch every bank group once
__host__ __device__ constexpr int cta_evld(int n) { int h = (n + 1) / 2; if ((h & 1) == 0) ++h; return 2 * h; }

template <class DM>
inline CtaLayout cta_layout(const DM& d, int fac_shared, int generic_scratch) {
    CtaLayout L;
    int o = 0;
    auto take = [&](int cnt) { int r = o; o += (cnt + 1) & ~1; return r; };
    L.ldG = cta_evld(d.b); L.ldC = cta_evld(d.nx); L.LD = cta_evld(d.b);
    L.oWm = d.b * L.LD; L.FS2 = (d.b + d.ne) * L.LD;
    L.nP = (d.ph + 1) * d.b;
    L.oG = take(d.ne * L.ldG); L.oC = take(d.ny * L.ldC); L.oSV = take(d.ne);
    L.oW = take((d.ph + 1) * (d.ny + 2 * d.nu));
    L.oD = take(L.nP); L.oQ = take(L.nP); L.oX = take(L.nP); L.oT = take(L.nP); L.oR = take(L.nP + d.b);
    L.oE = take(d.m); L.oLO = take(d.m); L.oUP = take(d.m); L.oZ = take(d.m); L.oY = take(d.m); L.oV = take(d.m);
    L.oRT = take((d.m + 7) / 8);
    L.oGT = take((d.ph + 1) * d.ne); L.oCAR = take((d.ph + 2) * d.ne); L.oXE = take((d.ph + 2) * d.ne);
    L.oYV = take((d.ph + 1) * d.ny);
    L.oRED = take(16 * 16);
    L.oLCS = take(d.ne * L.LD);
    L.oTB = take(d.ne * L.LD);
    L.oFLAG = o;
    L.lda = d.b | 1;
    L.n_aug = d.b * (d.b + 1) / 2;
    L.n_ent = L.n_aug + d.b * (d.b - 1) / 2;
    L.oAUG = L.oM = L.oTAB = o;
    if (generic_scratch) {
        L.oAUG = take(d.b * L.lda);
        L.oM = take(d.b * L.LD);
        L.oTAB = take((L.n_ent + 1) / 2);            // int32 table, 2 per double
    }
    L.oFAC = o;
    L.fac_shared = fac_shared;
    if (fac_shared) o += (d.ph + 1) * L.FS2;
    L.total = o;
    long long g = 0;
    auto gt = [&](long long cnt) { long long r = g; g += (cnt + 3) & ~3ll; return r; };
    L.gVA = (int)gt(L.nP); L.gPX = (int)gt(L.nP); L.gRA = (int)gt(d.m); L.gRB = (int)gt(d.m); L.gRC = (int)gt(d.m);
    L.gFAC = gt((long long)(d.ph + 1) * L.FS2);     // also the Ruiz scratch (P blocks of all stages)
    L.gtotal = g;
    return L;
}

#define CSM(name) (smem + L.o##name)

// compile-time loop (register arrays must only ever be indexed by constants)
template <int K, int N, class Fn>
__device__ __forceinline__ void cta_static_for(Fn&& fn) {
    if constexpr (K < N) { fn(std::integral_constant<int, K>{}); cta_static_for<K + 1, N>(fn); }
}

// compile-time dimensions of a dimension policy (0 = run-time): lengths of the unrolled dot products
template <class DM, bool S = DM::is_static> struct CtaDims { static constexpr int b = 0, ne = 0, nx = 0, ny = 0, LD = 0, ldG = 0, ldC = 0; };
template <class DM> struct CtaDims<DM, true> {
    static constexpr bool even = (DM::b % 2 == 0) && (DM::ne % 2 == 0) && (DM::nx % 2 == 0) && (DM::ny % 2 == 0);
    static constexpr int b = even ? DM::b : 0, ne = even ? DM::ne : 0, nx = even ? DM::nx : 0, ny = even ? DM::ny : 0;
    static constexpr int LD = cta_evld(DM::b), ldG = cta_evld(DM::b), ldC = cta_evld(DM::nx);   // immediates in every strided address
};

// dot product  sum_q a[q*sa] x[q].  NS > 0: compile-time length, fully unrolled, four accumulators, all loads issued ahead of
// the multiply-adds.  MODE 0: `a` is a contiguous 16-byte aligned row and `x` is 16-byte aligned (128-bit loads of both);
// MODE 1: `a` strided, `x` aligned;  MODE 2: `a` strided, `x` unaligned.  NS == 0: run-time length n.
template <int NS, int MODE>
__device__ __forceinline__ double cta_dot(const double* __restrict__ a, int sa, const double* __restrict__ x, int n) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    if constexpr (NS > 0) {
        if constexpr (MODE == 0) {
            const double2* a2 = reinterpret_cast<const double2*>(a); const double2* x2 = reinterpret_cast<const double2*>(x);
#pragma unroll
            for (int j = 0; j < NS / 2; ++j) {
                const double2 av = a2[j], xv = x2[j];
                if (j & 1) { s2 = fma(av.x, xv.x, s2); s3 = fma(av.y, xv.y, s3); } else { s0 = fma(av.x, xv.x, s0); s1 = fma(av.y, xv.y, s1); }
            }
        } else if constexpr (MODE == 1) {
            const double2* x2 = reinterpret_cast<const double2*>(x);
#pragma unroll
            for (int j = 0; j < NS / 2; ++j) {
                const double2 xv = x2[j];
                const double alo = a[(2 * j) * sa], ahi = a[(2 * j + 1) * sa];
                if (j & 1) { s2 = fma(alo, xv.x, s2); s3 = fma(ahi, xv.y, s3); } else { s0 = fma(alo, xv.x, s0); s1 = fma(ahi, xv.y, s1); }
            }
        } else {
#pragma unroll
            for (int j = 0; j < NS / 2; ++j) {
                const double alo = a[(2 * j) * sa], ahi = a[(2 * j + 1) * sa], xlo = x[2 * j], xhi = x[2 * j + 1];
                if (j & 1) { s2 = fma(alo, xlo, s2); s3 = fma(ahi, xhi, s3); } else { s0 = fma(alo, xlo, s0); s1 = fma(ahi, xhi, s1); }
            }
        }
    } else {
        int q = 0;
        for (; q + 1 < n; q += 2) { s0 = fma(a[q * sa], x[q], s0); s1 = fma(a[(q + 1) * sa], x[q + 1], s1); }
        if (q < n) s0 = fma(a[q * sa], x[q], s0);
    }
    return (s0 + s1) + (s2 + s3);
}

template <class DM, int NT, bool FSH>
struct CtaSolver {
    static constexpr int NW = NT / 32;
    static constexpr int SB = CtaDims<DM>::b, SNE = CtaDims<DM>::ne, SNX = CtaDims<DM>::nx, SNY = CtaDims<DM>::ny;
    const DM& d; const Params& p; const Prob& pr; const CtaLayout& L;
    // leading dimensions: compile-time for static dimension policies (the host layout computes the same values)
    __device__ __forceinline__ int fLD() const { if constexpr (DM::is_static) return CtaDims<DM>::LD; else return L.LD; }
    __device__ __forceinline__ int fldG() const { if constexpr (DM::is_static) return CtaDims<DM>::ldG; else return L.ldG; }
    __device__ __forceinline__ int fWm() const { return d.b * fLD(); }
    __device__ __forceinline__ int fldC() const { if constexpr (DM::is_static) return CtaDims<DM>::ldC; else return L.ldC; }
    int inst, tid, lane, warp;
    double* gws;
    double c;
    double rsel[3], rinv[3];
    double time_limit; long long t_start;
    long long pw[8];     // profiling aid (warp 0): cycles in [0] the forward recurrence, [1] the backward recurrence, all KKT solves of a run

    __device__ CtaSolver(const DM& d_, const Params& p_, const Prob& pr_, const CtaLayout& L_) : d(d_), p(p_), pr(pr_), L(L_) {}

    __device__ __forceinline__ double* fac(int i) const {
        if constexpr (FSH) return smem + L.oFAC + i * L.FS2; else return gws + L.gFAC + (size_t)i * L.FS2;
    }
    __device__ __forceinline__ double* pblk() const { if constexpr (FSH) return smem + L.oFAC; else return gws + L.gFAC; }
    __device__ __forceinline__ int jcol(int i) const { return i > 0 ? i - 1 : 0; }
    __device__ __forceinline__ double wO(int i, int r) const { return ldp(pr.OW, inst, jcol(i) * d.ny + r); }
    __device__ __forceinline__ double wU(int i, int r) const { return ldp(pr.UW, inst, jcol(i) * d.nu + r); }
    __device__ __forceinline__ double wDU(int i, int r) const { return ldp(pr.DUW, inst, i * d.nu + r); }
    __device__ __forceinline__ const int8_t* rtp() const { return reinterpret_cast<const int8_t*>(smem + L.oRT); }
    __device__ __forceinline__ double rho_of(int ty) const { return ty == 0 ? rsel[0] : (ty == 1 ? rsel[1] : rsel[2]); }
    __device__ __forceinline__ double rinv_of(int ty) const { return ty == 0 ? rinv[0] : (ty == 1 ? rinv[1] : rinv[2]); }
    __device__ __forceinline__ void set_rho(double rho) {
        rsel[0] = kRhoMin; rsel[1] = rho; rsel[2] = kRhoEqOverIneq * rho;
        for (int k = 0; k < 3; ++k) rinv[k] = 1.0 / rsel[k];
    }
    // row g -> (stage i, row r within the stage); i = -1: the eq(0) rows
    __device__ __forceinline__ void row_of(int g, int& i, int& r) const {
        if (g < d.ne) { i = -1; r = g; return; }
        int e = g - d.ne; i = e / d.RS; r = e - i * d.RS;
    }

    // ---- block reductions: v[j] <- op_j over the block (bit j of maxmask: max, else sum); identical in every thread ----
    template <int N>
    __device__ __forceinline__ void reduce(double (&v)[N], unsigned maxmask) {
        double* red = CSM(RED);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double x = v[j];
            if ((maxmask >> j) & 1u) { for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o)); }
            else { for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o); }
            if (lane == 0) red[warp * N + j] = x;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double x = red[j];
            if ((maxmask >> j) & 1u) { for (int w = 1; w < NW; ++w) x = fmax(x, red[w * N + j]); }
            else { for (int w = 1; w < NW; ++w) x += red[w * N + j]; }
            v[j] = x;
        }
        __syncthreads();
    }

    // ---- model to shared memory: G = [A B B; 0 I I] (ne x b), C, scalar row -------------------------------------------------
    __device__ void load_model() {
        double* G = CSM(G); double* Cm = CSM(C); double* sv = CSM(SV);
        for (int e = tid; e < d.ne * L.ldG; e += NT) {
            int r = e / L.ldG, k = e - r * L.ldG;
            double v = 0.0;
            if (k < d.b) {
                if (r < d.nx) {
                    if (k < d.nx) v = ldp(pr.A, inst, r * d.nx + k);
                    else if (k < d.ne) v = ldp(pr.B, inst, r * d.nu + (k - d.nx));
                    else v = ldp(pr.B, inst, r * d.nu + (k - d.ne));
                } else {
                    int j = r - d.nx;
                    v = ((k >= d.nx && k < d.ne && k - d.nx == j) || (k >= d.ne && k - d.ne == j)) ? 1.0 : 0.0;
                }
            }
            G[e] = v;
        }
        for (int e = tid; e < d.ny * L.ldC; e += NT) { int r = e / L.ldC, k = e - r * L.ldC; Cm[e] = k < d.nx ? ldp(pr.C, inst, r * d.nx + k) : 0.0; }
        for (int k = tid; k < d.ne; k += NT) sv[k] = k < d.nx ? ldp(pr.SX, inst, k) : ldp(pr.SU, inst, k - d.nx);
        __syncthreads();
    }

    // unscaled bounds of row r of stage i (ProblemBuilder.hpp:597-630,727-809)
    __device__ __forceinline__ void stage_bounds(int i, int r, double& l, double& u) const {
        int j = jcol(i);
        const double inf = INFINITY;
        if (r < d.oOUT) {
            if (r < d.nx) { l = ldp(pr.XMin, inst, j * d.nx + r); u = ldp(pr.XMax, inst, j * d.nx + r); }
            else { int q = r - d.nx; int col = i < d.ph ? i : d.ph - 1;
                   l = ldp(pr.UMin, inst, col * d.nu + q); u = ldp(pr.UMax, inst, col * d.nu + q); }
        } else if (r < d.oSC) {
            int q = r - d.oOUT;
            double off = 0;
            for (int e = 0; e < d.ndu; ++e) off -= ldp(pr.Dd, inst, q * d.ndu + e) * ldp(pr.uMeas, inst, j * d.ndu + e);
            l = ldp(pr.YMin, inst, j * d.ny + q) + off; u = ldp(pr.YMax, inst, j * d.ny + q) + off;
        } else if (r < d.oEQ) {
            l = ldp(pr.SMin, inst, j); u = ldp(pr.SMax, inst, j);
        } else if (r < d.oDU) {
            int q = r - d.oEQ;
            double v = 0;
            if (q < d.nx) for (int e = 0; e < d.ndu; ++e) v -= ldp(pr.Bd, inst, q * d.ndu + e) * ldp(pr.uMeas, inst, i * d.ndu + e);
            l = u = v;
        } else {
            bool frozen = i > d.ch;                                         // ProblemBuilder.hpp:784-785
            l = frozen ? 0.0 : -inf; u = frozen ? 0.0 : inf;
        }
    }

    // column inf-norm of the D-scaled (not yet c-scaled) P column k of stage i
    __device__ __forceinline__ double Pcol_norm(int i, int k) const {
        const double* Dv = CSM(D) + i * d.b; const double* Wst = CSM(W) + i * (d.ny + 2 * d.nu);
        if (k < d.nx) {
            const double* P = pblk() + (size_t)i * d.nx * d.nx;
            double mx = 0;
            for (int j = 0; j < d.nx; ++j) mx = fmax(mx, Dv[j] * fabs(P[j * d.nx + k]));
            return mx * Dv[k];
        } else if (k < d.ne) return Dv[k] * Dv[k] * fabs(Wst[d.ny + k - d.nx]);
        return Dv[k] * Dv[k] * fabs(Wst[d.ny + d.nu + k - d.ne]);
    }

    // ---- set-up: q (ProblemBuilder::get), Ruiz equilibration (scaling.c scale_data), scaled bounds, row types -----------------
    __device__ bool setup_and_scale() {
        double* G = CSM(G); double* Cm = CSM(C); double* sv = CSM(SV); double* Wr = CSM(W);
        double* D = CSM(D); double* Q = CSM(Q); double* E = CSM(E); double* YV = CSM(YV);
        double* Dt = CSM(T); double* Et = CSM(V);
        const int nw = d.ny + 2 * d.nu, ldG = fldG(), ldC = fldC();
        for (int e = tid; e < (d.ph + 1) * nw; e += NT) {
            int i = e / nw, r = e - i * nw;
            Wr[e] = r < d.ny ? wO(i, r) : (r < d.ny + d.nu ? wU(i, r - d.ny) : (i < d.ph ? wDU(i, r - d.ny - d.nu) : 0.0));
        }
        for (int e = tid; e < (d.ph + 1) * d.ny; e += NT) {
            int i = e / d.ny, r = e - i * d.ny, j = jcol(i);
            double acc = -ldp(pr.yRef, inst, j * d.ny + r);
            for (int q = 0; q < d.ndu; ++q) acc += ldp(pr.Dd, inst, r * d.ndu + q) * ldp(pr.uMeas, inst, j * d.ndu + q);
            YV[e] = wO(i, r) * acc;
        }
        // the du slots of the last stage do not exist: keep them zero in every stage-major vector
        for (int kg = d.n + tid; kg < L.nP + d.b; kg += NT) {
            if (kg < L.nP) { D[kg] = 0.0; Q[kg] = 0.0; CSM(X)[kg] = 0.0; CSM(T)[kg] = 0.0; }
            CSM(R)[kg] = 0.0;
        }
        __syncthreads();
        for (int kg = tid; kg < d.n; kg += NT) {
            int i = kg / d.b, k = kg - i * d.b, j = jcol(i);
            double v;
            if (k < d.nx) { v = 0; for (int r = 0; r < d.ny; ++r) v += Cm[r * ldC + k] * YV[i * d.ny + r]; }
            else if (k < d.ne) { int q = k - d.nx; v = Wr[i * nw + d.ny + q] * (-ldp(pr.uRef, inst, j * d.nu + q)); }
            else { int q = k - d.ne; v = -(Wr[i * nw + d.ny + d.nu + q] * ldp(pr.duRef, inst, j * d.nu + q)); }
            Q[kg] = v; D[kg] = 1.0;
        }
        for (int g = tid; g < d.m; g += NT) E[g] = 1.0;
        {   // P blocks C' diag(wO_i) C of every stage (scratch: the factor area, unused until the first factorisation)
            double* P = pblk();
            const int nn = d.nx * d.nx;
            for (int e = tid; e < (d.ph + 1) * nn; e += NT) {
                int i = e / nn, rem = e - i * nn, a = rem / d.nx, k = rem - a * d.nx;
                double acc = 0;
                for (int r = 0; r < d.ny; ++r) acc += Cm[r * ldC + a] * Wr[i * nw + r] * Cm[r * ldC + k];
                P[e] = acc;
            }
        }
        __syncthreads();
        c = 1.0;
        double pending_c = 1.0;
        for (int it = 0; it < p.scaling; ++it) {
            for (int kg = tid; kg < d.n; kg += NT) {
                int i = kg / d.b, k = kg - i * d.b;
                const double* erow = E + d.roff(i);
                const double* eprev = i == 0 ? E : E + d.roff(i - 1) + d.oEQ;
                double dk = D[kg];
                double cn = c * Pcol_norm(i, k);
                if (k < d.ne) {
                    cn = fmax(cn, eprev[k] * dk);
                    cn = fmax(cn, erow[d.oBOX + k] * dk);
                    if (k < d.nx) for (int r = 0; r < d.ny; ++r) cn = fmax(cn, erow[d.oOUT + r] * fabs(Cm[r * ldC + k]) * dk);
                    cn = fmax(cn, erow[d.oSC] * fabs(sv[k]) * dk);
                } else cn = fmax(cn, erow[d.oDU + k - d.ne] * dk);
                if (i < d.ph) for (int r = 0; r < d.ne; ++r) cn = fmax(cn, erow[d.oEQ + r] * fabs(G[r * ldG + k]) * dk);
                Dt[kg] = 1.0 / sqrt(lim_scaling(cn));
            }
            for (int g = tid; g < d.m; g += NT) {
                int i, r; row_of(g, i, r);
                double e = E[g], rn;
                if (i < 0) rn = e * D[r];
                else {
                    const double* dcur = D + i * d.b;
                    if (r < d.oOUT) rn = e * dcur[r];
                    else if (r < d.oSC) { rn = 0; int q = r - d.oOUT; for (int k = 0; k < d.nx; ++k) rn = fmax(rn, e * fabs(Cm[q * ldC + k]) * dcur[k]); }
                    else if (r < d.oEQ) { rn = 0; for (int k = 0; k < d.ne; ++k) rn = fmax(rn, e * fabs(sv[k]) * dcur[k]); }
                    else if (r < d.oDU) { int q = r - d.oEQ; rn = e * dcur[d.b + q]; for (int k = 0; k < d.b; ++k) rn = fmax(rn, e * fabs(G[q * ldG + k]) * dcur[k]); }
                    else rn = e * dcur[d.ne + r - d.oDU];
                }
                Et[g] = 1.0 / sqrt(lim_scaling(rn));
            }
            __syncthreads();
            double red[2] = {0.0, 0.0};    // [psum, qmax]
            for (int g = tid; g < d.m; g += NT) E[g] *= Et[g];
            for (int kg = tid; kg < d.n; kg += NT) {
                double dt = Dt[kg];
                D[kg] *= dt;
                double qv = (Q[kg] * pending_c) * dt;
                Q[kg] = qv; red[1] = fmax(red[1], fabs(qv));
            }
            __syncthreads();
            for (int kg = tid; kg < d.n; kg += NT) { int i = kg / d.b, k = kg - i * d.b; red[0] += c * Pcol_norm(i, k); }
            reduce<2>(red, 0x2u);
            double ct = red[0] / (double)d.n;
            double nq = lim_scaling(red[1]);
            ct = fmax(ct, nq);
            ct = 1.0 / lim_scaling(ct);
            c *= ct; pending_c = ct;
        }
        // scaled q, scaled bounds, row types; validate l <= u
        bool bad = false;
        double* LO = CSM(LO); double* UP = CSM(UP);
        int8_t* rt = reinterpret_cast<int8_t*>(smem + L.oRT);
        for (int kg = tid; kg < d.n; kg += NT) Q[kg] *= pending_c;
        for (int kg = d.n + tid; kg < L.nP; kg += NT) CSM(T)[kg] = 0.0;      // (T served as scratch)
        for (int g = tid; g < d.m; g += NT) {
            int i, r; row_of(g, i, r);
            double e = E[g];
            if (i < 0) {
                double v = r < d.nx ? -__ldg(pr.x0 + (long long)inst * d.nx + r) : -__ldg(pr.u0 + (long long)inst * d.nu + (r - d.nx));
                v *= e;
                LO[g] = v; UP[g] = v; rt[g] = 2;
            } else {
                double l, u; stage_bounds(i, r, l, u);
                bad |= (l > u);
                l *= e; u *= e;
                LO[g] = l; UP[g] = u;
                rt[g] = ((l < -kOsqpInfty * kMinScaling) && (u > kOsqpInfty * kMinScaling)) ? 0 : ((u - l < kRhoTol) ? 2 : 1);
            }
        }
        return !__syncthreads_or(bad ? 1 : 0);
    }

    // ---- structured products ---------------------------------------------------------------------------------------------
    // unscaled (A' v)[kg] for the E-weighted row values Vs
    __device__ __forceinline__ double col_atv(int i, int k, const double* Vs) const {
        const double* vj = Vs + d.roff(i);
        const double* vprev = i == 0 ? Vs : Vs + d.roff(i - 1) + d.oEQ;
        double au;
        if (k < d.ne) {
            au = vj[d.oBOX + k] - vprev[k] + CSM(SV)[k] * vj[d.oSC];
            if (k < d.nx) au += cta_dot<SNY, 2>(CSM(C) + k, fldC(), vj + d.oOUT, d.ny);
        } else au = vj[d.oDU + k - d.ne];
        if (i < d.ph) au += cta_dot<SNE, 2>(CSM(G) + k, fldG(), vj + d.oEQ, d.ne);
        return au;
    }
    // row tasks are ordered by row class so that the lanes of a warp run the same dot-product length:
    // [eq(i+1) rows | out rows | sc rows | box rows | du rows | eq(0) rows];  returns the row index g and (i, r)
    __device__ __forceinline__ int row_task(int t, int& i, int& r) const {
        const int nEQ = d.ph * d.ne, nOUT = (d.ph + 1) * d.ny, nSC = d.ph + 1, nBOX = (d.ph + 1) * d.ne, nDU = d.ph * d.nu;
        if (t < nEQ) { i = t / d.ne; r = d.oEQ + (t - i * d.ne); }
        else if ((t -= nEQ) < nOUT) { i = t / d.ny; r = d.oOUT + (t - i * d.ny); }
        else if ((t -= nOUT) < nSC) { i = t; r = d.oSC; }
        else if ((t -= nSC) < nBOX) { i = t / d.ne; r = d.oBOX + (t - i * d.ne); }
        else if ((t -= nBOX) < nDU) { i = t / d.nu; r = d.oDU + (t - i * d.nu); }
        else { t -= nDU; i = -1; r = t; return t; }
        return d.roff(i) + r;
    }
    // a_g . u for the unscaled variable values u = D x in R
    __device__ __forceinline__ double row_dot(int i, int r) const {
        const double* U = CSM(R);
        if (i < 0) return -U[r];
        const double* u = U + i * d.b;
        if (r < d.oOUT) return u[r];
        if (r < d.oSC) return cta_dot<SNX, 0>(CSM(C) + (r - d.oOUT) * fldC(), 1, u, d.nx);
        if (r < d.oEQ) return cta_dot<SNE, 0>(CSM(SV), 1, u, d.ne);
        if (r < d.oDU) { const int q = r - d.oEQ; return cta_dot<SB, 0>(CSM(G) + q * fldG(), 1, u, d.b) - u[d.b + q]; }
        return u[d.ne + r - d.oDU];
    }
    // c D_k (P u)_k of variable (i,k); needs YV_i = wO_i .* (C u_x) for k < nx
    __device__ __forceinline__ double col_pu(int i, int k) const {
        const double* Wst = CSM(W) + i * (d.ny + 2 * d.nu);
        const double* U = CSM(R) + i * d.b;
        double pu;
        if (k < d.nx) pu = cta_dot<SNY, 1>(CSM(C) + k, fldC(), CSM(YV) + i * d.ny, d.ny);
        else if (k < d.ne) pu = Wst[d.ny + k - d.nx] * U[k];
        else pu = Wst[d.ny + d.nu + k - d.ne] * U[k];
        return pu * c * CSM(D)[i * d.b + k];
    }
    __device__ __forceinline__ void fill_yv() {      // YV_i = wO_i .* (C u_x,i)
        const int nw = d.ny + 2 * d.nu;
        for (int e = tid; e < (d.ph + 1) * d.ny; e += NT) {
            int i = e / d.ny, r = e - i * d.ny;
            CSM(YV)[e] = CSM(W)[i * nw + r] * cta_dot<SNX, 0>(CSM(C) + r * fldC(), 1, CSM(R) + i * d.b, d.nx);
        }
    }
    __device__ __forceinline__ void refresh_V() {    // V = E (rho z - y): the row weights of the next right-hand side
        const double* E = CSM(E); const double* Z = CSM(Z); const double* Y = CSM(Y); double* V = CSM(V);
        const int8_t* rt = rtp();
        for (int g = tid; g < d.m; g += NT) V[g] = E[g] * (rho_of(rt[g]) * Z[g] - Y[g]);
        __syncthreads();
    }

    // ---- elimination table of the run-time-dimension path: (kind, row, col) of every entry, built once per launch ------------
    __device__ void build_table() {
        if constexpr (!DM::is_static) {
            int* tab = reinterpret_cast<int*>(smem + L.oTAB);
            for (int e = tid; e < L.n_ent; e += NT) {
                int kind = e < L.n_aug ? 0 : 1;
                int pidx = kind ? e - L.n_aug : e;
                int r = (int)((sqrt(8.0 * pidx + 1.0) - 1.0) * 0.5);
                while ((r + 1) * (r + 2) / 2 <= pidx) ++r;
                while (r * (r + 1) / 2 > pidx) --r;
                int k = pidx - r * (r + 1) / 2;
                if (kind) r += 1;                       // strict lower triangle of M: (r+1, k)
                tab[e] = (kind << 30) | (r << 15) | k;
            }
        }
        __syncthreads();
    }

    // ---- block-tridiagonal factorisation of H = D (c P + A' R A) D + sigma I ---------------------------------------------------
    // Per stage i: S_i = H_ii - Lc_{i-1} Lc_{i-1}',  S_i = L_i L_i',  Lc_i = H_{i+1,i} L_i^-T.  Stored: L_i^-1 (b x LD, zero above
    // the diagonal and in the rows / columns a short last stage does not have) and W_i = Lc_i L_i^-1 (ne x LD).
    __device__ bool factorize(double sigma) {
        const double* G = CSM(G); const double* Cm = CSM(C); const double* sv = CSM(SV); const double* Wr = CSM(W);
        const double* D = CSM(D); const double* E = CSM(E);
        double* RW = CSM(V);                         // rho' = rho E^2 per row (V is rebuilt afterwards)
        double* YV = CSM(YV);
        double* LCS = CSM(LCS); double* TB = CSM(TB);
        const int8_t* rt = rtp();
        const int nw = d.ny + 2 * d.nu, LD = fLD(), ldG = fldG(), ldC = fldC();
        for (int g = tid; g < d.m; g += NT) { double e = E[g]; RW[g] = rho_of(rt[g]) * e * e; }
        __syncthreads();
        for (int e = tid; e < (d.ph + 1) * d.ny; e += NT) { int i = e / d.ny, r = e - i * d.ny; YV[e] = c * Wr[i * nw + r] + RW[d.roff(i) + d.oOUT + r]; }
        __syncthreads();
        // H_ii of every stage (stage-parallel) into the L^-1 slot of the stage's record: lower triangle, zero elsewhere
        for (int t = tid; t < (d.ph + 1) * d.b * d.b; t += NT) {
            const int i = t / (d.b * d.b), rem = t - i * d.b * d.b, r = rem / d.b, k = rem - r * d.b;
            double v = 0;
            if (k <= r && r < d.bcount(i)) {
                const double* rw = RW + d.roff(i);
                const double* rwp = i == 0 ? RW : RW + d.roff(i - 1) + d.oEQ;
                const double* wst = Wr + i * nw; const double* yv = YV + i * d.ny;
                if (i < d.ph) for (int j = 0; j < d.ne; ++j) v += G[j * ldG + r] * rw[d.oEQ + j] * G[j * ldG + k];
                if (r < d.ne) {
                    v += rw[d.oSC] * sv[r] * sv[k];
                    if (r < d.nx) for (int j = 0; j < d.ny; ++j) v += Cm[j * ldC + r] * yv[j] * Cm[j * ldC + k];
                    if (r == k) {
                        v += rwp[k] + rw[d.oBOX + k];
                        if (k >= d.nx) v += c * wst[d.ny + k - d.nx];
                    }
                } else if (r == k) v += c * wst[d.ny + d.nu + k - d.ne] + rw[d.oDU + k - d.ne];
                v = D[i * d.b + r] * v * D[i * d.b + k];
                if (r == k) v += sigma;
            }
            fac(i)[r * LD + k] = v;
        }
        __syncthreads();
        bool ok = true;
        double* COL = CSM(RED);                      // 128 doubles of warp-0 scratch: three pivot-column buffers and the row scales
        for (int i = 0; i <= d.ph; ++i) {
            const int bi = d.bcount(i);
            double* F = fac(i);
            // ---- S_i = H_ii + TB (TB = -Lc_{i-1} Lc_{i-1}'),  S_i = L~ diag(dd) L~',  L_i^-1 = diag(dd)^-1/2 L~^-1 ----
            if constexpr (DM::is_static) {
                // One warp, registers only: lane r holds row r of S (lower part), lane q holds COLUMN q of L~^-1.  Per pivot k
                // the column k of S is published through shared memory (one store per lane), everything else is register math:
                //   a_rq -= (a_rk / d_k) a_qk  (q <= r),      m_rq -= (a_rk / d_k) m_kq  (q <= k < r).
                constexpr int B = DM::b;
                if (warp == 0) {
                    const int r = lane;
                    double a[B], m[B], dsave = 1.0;
                    cta_static_for<0, B>([&](auto qc) {
                        constexpr int q = decltype(qc)::value;
                        double v = 0.0;
                        if (r < bi && q <= r) { v = F[r * LD + q]; if (i > 0 && r < d.ne) v += TB[r * LD + q]; }
                        a[q] = v; m[q] = (q == r) ? 1.0 : 0.0;           // m[x]: entry (row x, column = lane)
                    });
                    // Look-ahead: column k+1 is updated and published FIRST in step k, so its trip through shared memory and the
                    // reciprocal of the next pivot overlap the rest of the trailing update (three column buffers: the one written in
                    // step k is read in step k+1 and overwritten in step k+3, with a warp barrier in between).
                    COL[lane] = a[0];
                    __syncwarp();
                    cta_static_for<0, B>([&](auto kc) {
                        constexpr int k = decltype(kc)::value;
                        if (k < bi) {
                            const double* cur = COL + 32 * (k % 3);
                            const double dk = cur[k];
                            if (!(dk > 0.0)) ok = false;
                            const double rd = 1.0 / dk;
                            if (r == k) dsave = dk;
                            const double l = (r > k && r < bi) ? a[k] * rd : 0.0;
                            const double mk = (r <= k) ? m[k] * rd : 0.0;
                            if constexpr (k + 1 < B) {
                                const double cq = cur[k + 1];
                                a[k + 1] = (k + 1 <= r) ? fma(-l, cq, a[k + 1]) : a[k + 1];
                                m[k + 1] = fma(-cq, mk, m[k + 1]);
                                COL[32 * ((k + 1) % 3) + lane] = a[k + 1];
                                __syncwarp();
                            }
                            cta_static_for<k + 2, B>([&](auto qc) {
                                constexpr int q = decltype(qc)::value;
                                const double cq = cur[q];
                                a[q] = (q <= r) ? fma(-l, cq, a[q]) : a[q];
                                m[q] = fma(-cq, mk, m[q]);
                            });
                        }
                    });
                    __syncwarp();
                    // row scales 1/sqrt(d_r), then L^-1(x, q) = m_q[x] rs_x : lane q writes column q (zeros where L^-1 has none)
                    COL[96 + lane] = rsqrt(dsave);
                    __syncwarp();
                    cta_static_for<0, B>([&](auto xc) {
                        constexpr int x = decltype(xc)::value;
                        if (lane < B) F[x * LD + lane] = (x >= lane && x < bi && lane < bi) ? m[x] * COL[96 + x] : 0.0;
                    });
                }
                __syncthreads();
            } else {
                // run-time dimensions: the same elimination by the whole block on shared memory, one barrier per pivot
                double* AUG = CSM(AUG); double* M = CSM(M);
                const int* tab = reinterpret_cast<const int*>(smem + L.oTAB);
                const int lda = L.lda;
                for (int e = tid; e < L.n_ent; e += NT) {
                    const int te = tab[e];
                    const int r = (te >> 15) & 0x7fff, k = te & 0x7fff;
                    if (te >> 30) M[r * LD + k] = 0.0;
                    else if (r < bi) { double v = F[r * LD + k]; if (i > 0 && r < d.ne) v += TB[r * LD + k]; AUG[r * lda + k] = v; }
                }
                __syncthreads();
                for (int k = 0; k < bi; ++k) {
                    const double dk = AUG[k * lda + k];
                    if (!(dk > 0.0)) ok = false;
                    const double rd = 1.0 / dk;
                    for (int e = tid; e < L.n_ent; e += NT) {
                        const int te = tab[e];
                        const int r = (te >> 15) & 0x7fff, q = te & 0x7fff;
                        if (r >= bi) continue;
                        if (te >> 30) {      // M(r,q), strict lower: active for q <= k < r
                            if (q <= k && k < r) M[r * LD + q] -= (AUG[r * lda + k] * rd) * ((k == q) ? 1.0 : M[k * LD + q]);
                        } else if (q > k) AUG[r * lda + q] -= (AUG[r * lda + k] * rd) * AUG[q * lda + k];
                    }
                    __syncthreads();
                }
                for (int e = tid; e < d.b * d.b; e += NT) {
                    const int r = e / d.b, q = e - r * d.b;
                    double v = 0.0;
                    if (r < bi && q <= r) v = (q == r ? 1.0 : M[r * LD + q]) * rsqrt(AUG[r * lda + r]);
                    F[r * LD + q] = v;
                }
                __syncthreads();
            }
            if (i == d.ph) break;
            // ---- Lc_i = Hc_i L_i^-T,  Hc_i = -(D_e(i+1) rho'_eq(i+1)) G D_i   (ne x b, stage-parallel dot products) ----
            for (int e = tid; e < d.ne * d.b; e += NT) {
                const int h = e / d.b, k = e - h * d.b;
                const double hs = -(D[(i + 1) * d.b + h] * RW[d.roff(i) + d.oEQ + h]);
                const double* gp = G + h * ldG; const double* dp = D + i * d.b; const double* li = F + k * LD;
                double a0 = 0, a1 = 0;
                if constexpr (SB > 0) {
#pragma unroll
                    for (int q = 0; q < SB; q += 2) { a0 = fma(gp[q] * dp[q], li[q], a0); a1 = fma(gp[q + 1] * dp[q + 1], li[q + 1], a1); }
                } else for (int q = 0; q <= k; ++q) a0 = fma(gp[q] * dp[q], li[q], a0);
                LCS[h * LD + k] = hs * (a0 + a1);
            }
            __syncthreads();
            // ---- TB = -Lc Lc' (full symmetric ne x ne) for the next stage's pivot block;  W_i = Lc_i L_i^-1 ----
            {
                const int nT = d.ne * d.ne;
                for (int e = tid; e < nT + d.ne * d.b; e += NT) {
                    if (e < nT) {
                        const int r = e / d.ne, k = e - r * d.ne;
                        TB[r * LD + k] = -cta_dot<SB, 0>(LCS + r * LD, 1, LCS + k * LD, d.b);
                    } else {
                        const int e2 = e - nT, h = e2 / d.b, cc = e2 - h * d.b;
                        F[fWm() + h * LD + cc] = cta_dot<SB, 1>(F + cc, LD, LCS + h * LD, d.b);
                    }
                }
            }
            __syncthreads();
        }
        return !__syncthreads_or(ok ? 0 : 1);
    }

    // ---- the two stage recurrences of the KKT solve (one warp) -----------------------------------------------------------------
    //   forward   c_{i+1} = g_i - K_i c_i      (K_i = W_i[:, :ne];  g in GT, c in CAR)
    //   backward  xe_i = s_i[:ne] - K_i' xe_{i+1}   (s in R, xe in XE)
    // The carried vector goes through shared memory (store, warp barrier, broadcast loads: 34 cycles).  Keeping it in registers with
    // shuffle broadcast and the K rows loaded one stage ahead was measured no faster (~440 vs ~410 cycles per step: a step is
    // bound by its ~100 instructions on one warp, 32 SHFL.32 cost more issue slots than 8 LDS.128; tools/ulat.cu has the latencies).
    __device__ __forceinline__ void chain_forward() {
        double* CAR = CSM(CAR); const double* GT = CSM(GT);
        const int LD = fLD();
        for (int r = lane; r < d.ne; r += 32) CAR[r] = 0.0;
        __syncwarp();
        if constexpr (SNE == 16) {
            // two lanes per row: lane = (half, row) computes half of the dot product, one shuffle adds the halves -- the step is bound
            // by the instructions one warp issues, and this halves them
            constexpr int NE = SNE, H = SNE / 2;
            const int r = lane & (NE - 1), hf = (lane >> 4) & 1;      // NE == 16: lanes 0..15 first half, 16..31 second half
            // the half row of K_{i+1} and g_{i+1} are loaded before the barrier of step i: only the carried vector is waited for
            double2 ak[H / 2];
            double gk = GT[r];
            {
                const double2* a2 = reinterpret_cast<const double2*>(fac(0) + fWm() + r * LD + hf * H);
#pragma unroll
                for (int j = 0; j < H / 2; ++j) ak[j] = a2[j];
            }
            for (int i = 0; i < d.ph; ++i) {
                const double2* x2 = reinterpret_cast<const double2*>(CAR + i * NE + hf * H);
                double s0 = 0, s1 = 0;
#pragma unroll
                for (int j = 0; j < H / 2; ++j) { const double2 xv = x2[j]; s0 = fma(ak[j].x, xv.x, s0); s1 = fma(ak[j].y, xv.y, s1); }
                const int st = i + 1 < d.ph ? i + 1 : i;
                const double2* a2 = reinterpret_cast<const double2*>(fac(st) + fWm() + r * LD + hf * H);
#pragma unroll
                for (int j = 0; j < H / 2; ++j) ak[j] = a2[j];
                const double gn = GT[st * NE + r];
                double sv_ = s0 + s1;
                sv_ += __shfl_xor_sync(0xffffffffu, sv_, 16);
                if (lane < NE) CAR[(i + 1) * NE + lane] = gk - sv_;
                gk = gn;
                __syncwarp();
            }
        } else {
            for (int i = 0; i < d.ph; ++i) {
                for (int r = lane; r < d.ne; r += 32) CAR[(i + 1) * d.ne + r] = GT[i * d.ne + r] - cta_dot<SNE, 0>(fac(i) + fWm() + r * LD, 1, CAR + i * d.ne, d.ne);
                __syncwarp();
            }
        }
    }
    __device__ __forceinline__ void chain_backward() {
        double* XE = CSM(XE); const double* R = CSM(R);
        const int LD = fLD();
        for (int r = lane; r < d.ne; r += 32) XE[d.ph * d.ne + r] = R[d.ph * d.b + r];
        __syncwarp();
        if constexpr (SNE == 16) {
            constexpr int NE = SNE, H = SNE / 2;
            const int cc = lane & (NE - 1), hf = (lane >> 4) & 1;
            for (int i = d.ph - 1; i >= 0; --i) {
                const double* a = fac(i) + fWm() + cc + hf * H * LD;
                const double2* x2 = reinterpret_cast<const double2*>(XE + (i + 1) * NE + hf * H);
                double s0 = 0, s1 = 0;
#pragma unroll
                for (int j = 0; j < H / 2; ++j) { const double2 xv = x2[j]; s0 = fma(a[(2 * j) * LD], xv.x, s0); s1 = fma(a[(2 * j + 1) * LD], xv.y, s1); }
                double sv_ = s0 + s1;
                sv_ += __shfl_xor_sync(0xffffffffu, sv_, 16);
                if (lane < NE) XE[i * NE + lane] = R[i * d.b + lane] - sv_;
                __syncwarp();
            }
        } else {
            for (int i = d.ph - 1; i >= 0; --i) {
                for (int cc = lane; cc < d.ne; cc += 32) XE[i * d.ne + cc] = R[i * d.b + cc] - cta_dot<SNE, 1>(fac(i) + fWm() + cc, LD, XE + (i + 1) * d.ne, d.ne);
                __syncwarp();
            }
        }
    }

    // ---- reduced KKT solve, bulk-synchronous (polish): in R = right-hand side, out T = solution (scaled), R = D .* solution ----
    template <class Epi>
    __device__ __forceinline__ void kkt_solve(Epi epilogue) {
        double* R = CSM(R); double* T = CSM(T); double* XE = CSM(XE); const double* D = CSM(D);
        const int LD = fLD();
        {   // rhat_j = L_j^-1 r_j -> T,  g_j = W_j r_j -> GT as ONE flat task list over the rows of all stages (the records of
            // consecutive stages are contiguous: row t of the list starts at fac(0) + t * LD) -- a warp per stage would run a second,
            // almost empty round for rows 32..35 of every stage
            const int nrow = d.b + d.ne, ntask = (d.ph + 1) * nrow - d.ne;      // the last stage has no W rows
            for (int t = tid; t < ntask; t += NT) {
                const int j = t / nrow, rr = t - j * nrow;
                const double v = cta_dot<SB, 0>(fac(0) + (size_t)t * LD, 1, R + j * d.b, d.b);
                if (rr < d.b) T[j * d.b + rr] = v; else CSM(GT)[j * d.ne + (rr - d.b)] = v;
            }
        }
        __syncthreads();
        if (warp == 0) { const long long q0 = clock64(); chain_forward(); pw[0] += clock64() - q0; }
        __syncthreads();
        // t_i = rhat_i - L_i^-1[:, :ne] c_i, then s_i = L_i^-T t_i: both flat over the variable index (two rounds instead of three
        // stage rounds at 20 of 32 lanes), one block barrier in between
        for (int kg = tid; kg < L.nP; kg += NT) {
            const int i = kg / d.b, r = kg - i * d.b;
            if (i > 0) T[kg] -= cta_dot<SNE, 0>(fac(i) + r * LD, 1, CSM(CAR) + i * d.ne, d.ne);
        }
        __syncthreads();
        for (int kg = tid; kg < L.nP; kg += NT) {
            const int i = kg / d.b, k = kg - i * d.b;
            R[kg] = cta_dot<SB, 1>(fac(i) + k, LD, T + i * d.b, d.b);
        }
        __syncthreads();
        if (warp == 0) { const long long q0 = clock64(); chain_backward(); pw[1] += clock64() - q0; }
        __syncthreads();
        for (int kg = tid; kg < d.n; kg += NT) {
            const int i = kg / d.b, k = kg - i * d.b;
            double xt;
            if (k < d.ne) xt = XE[i * d.ne + k];
            else xt = R[kg] - cta_dot<SNE, 1>(fac(i) + fWm() + k, LD, XE + (i + 1) * d.ne, d.ne);
            T[kg] = xt;
            R[kg] = D[kg] * xt;
            epilogue(kg, xt);
        }
        __syncthreads();
    }

    // z~ = A x~ of one row, relaxation, projection, dual update, row weight of the next right-hand side
    __device__ __forceinline__ void row_update(int g, int i, int r, bool store_delta) {
        const int ty = rtp()[g];
        const double alpha = p.alpha;
        const double e = CSM(E)[g];
        const double zt = e * row_dot(i, r);
        const double zo = CSM(Z)[g], yo = CSM(Y)[g];
        const double zr = alpha * zt + (1.0 - alpha) * zo;
        const double zn = fmin(fmax(zr + rinv_of(ty) * yo, CSM(LO)[g]), CSM(UP)[g]);
        const double dy = rho_of(ty) * (zr - zn);
        const double yn = yo + dy;
        CSM(Y)[g] = yn; CSM(Z)[g] = zn;
        CSM(V)[g] = e * (rho_of(ty) * zn - yn);
        if (store_delta) (gws + L.gRA)[g] = dy;
    }
    // One ADMM iteration, bulk-synchronous: every phase by all warps, the recurrences by warp 0 between barriers.
    __device__ __forceinline__ void bulk_iteration(bool store_delta) {
        {   // right-hand side of every stage, flat over the (padded) variable index
            const double sigma = p.sigma;
            for (int kg = tid; kg < L.nP; kg += NT) {
                const int j = kg / d.b, k = kg - j * d.b;
                CSM(R)[kg] = k < d.bcount(j) ? sigma * CSM(X)[kg] - CSM(Q)[kg] + CSM(D)[kg] * col_atv(j, k, CSM(V)) : 0.0;
            }
        }
        __syncthreads();
        const double alpha = p.alpha;
        double* va = gws + L.gVA;
        kkt_solve([&](int kg, double xt) {
            const double xo = CSM(X)[kg];
            const double xn = alpha * xt + (1.0 - alpha) * xo;
            CSM(X)[kg] = xn;
            if (store_delta) va[kg] = xn - xo;
        });
        for (int t = tid; t < d.m; t += NT) { int i, r; const int g = row_task(t, i, r); row_update(g, i, r, store_delta); }
        __syncthreads();
    }
    // ---- update_info (auxil.c): residuals and the norms of the termination test / rho estimate ---------------------------------
    // XSRC 0: x = X (ADMM iterate), 1: x = px (polish iterate, global).  ZY 0: (z, y) = (Z, Y); 1: polish pair
    // z = clip(Ax + pnu), y = Ax + pnu - z (project_normalcone).  Leaves rc = E y (global scratch: V, the row weights of the next right-hand side, must survive); clobbers R, YV.
    template <int XSRC, int ZYM>
    __device__ InfoNorms info_pass() {
        const double* D = CSM(D); const double* E = CSM(E); double* R = CSM(R);
        const double* LO = CSM(LO); const double* UP = CSM(UP); const double* Q = CSM(Q);
        const double* xs = XSRC ? gws + L.gPX : CSM(X);
        double* rc = gws + L.gRC; const double* rb = gws + L.gRB;
        for (int kg = tid; kg < d.n; kg += NT) R[kg] = D[kg] * xs[kg];
        __syncthreads();
        fill_yv();
        double a[16];
        for (int k = 0; k < 16; ++k) a[k] = 0.0;
        // a: 0 pri 1 nz 2 nAx 3 spri 4 snz 5 snAx | 6 dua 7 nq 8 nAty 9 nPx 10 sdua 11 snq 12 snAty 13 snPx | 14 xPx 15 qx
        for (int t = tid; t < d.m; t += NT) {
            int i, r; int g = row_task(t, i, r);
            double e = E[g], einv = 1.0 / e;
            double Ax = e * row_dot(i, r), z, y;
            if (ZYM == 0) { z = CSM(Z)[g]; y = CSM(Y)[g]; }
            else { double tt = Ax + rb[g]; z = fmin(fmax(tt, LO[g]), UP[g]); y = tt - z; }
            rc[g] = e * y;
            double pv = Ax - z;
            a[3] = fmax(a[3], fabs(pv)); a[4] = fmax(a[4], fabs(z)); a[5] = fmax(a[5], fabs(Ax));
            a[0] = fmax(a[0], fabs(einv * pv)); a[1] = fmax(a[1], fabs(einv * z)); a[2] = fmax(a[2], fabs(einv * Ax));
        }
        __syncthreads();
        for (int kg = tid; kg < d.n; kg += NT) {
            int i = kg / d.b, k = kg - i * d.b;
            double dk = D[kg], dinv = 1.0 / dk;
            double aty = dk * col_atv(i, k, rc), px = col_pu(i, k), q = Q[kg];
            double dv = q + px + aty;
            a[10] = fmax(a[10], fabs(dv)); a[11] = fmax(a[11], fabs(q)); a[12] = fmax(a[12], fabs(aty)); a[13] = fmax(a[13], fabs(px));
            a[6] = fmax(a[6], fabs(dinv * dv)); a[7] = fmax(a[7], fabs(dinv * q)); a[8] = fmax(a[8], fabs(dinv * aty)); a[9] = fmax(a[9], fabs(dinv * px));
            double xv = xs[kg];
            a[14] += xv * px; a[15] += q * xv;
        }
        reduce<16>(a, 0x3fffu);
        InfoNorms I;
        I.pri = a[0]; I.nz = a[1]; I.nAx = a[2]; I.spri = a[3]; I.snz = a[4]; I.snAx = a[5];
        I.dua = a[6] / c; I.nq = a[7]; I.nAty = a[8]; I.nPx = a[9]; I.sdua = a[10]; I.snq = a[11]; I.snAty = a[12]; I.snPx = a[13];
        I.xPx = a[14]; I.qx = a[15];
        return I;
    }

    // is_primal_infeasible (auxil.c); delta_y in ra (global).  Clobbers rb (global).
    __device__ bool primal_infeasible(double eps) {
        const double* E = CSM(E); const double* LO = CSM(LO); const double* UP = CSM(UP);
        double* ra = gws + L.gRA; double* V = gws + L.gRB;      // E dy: global scratch (the polish multipliers are not live here)
        double a[2] = {0.0, 0.0};      // [nd (max), lhs (sum)]
        for (int g = tid; g < d.m; g += NT) {
            double dy = ra[g], l = LO[g], u = UP[g];
            if (u > kOsqpInfty * kMinScaling) {
                if (l < -kOsqpInfty * kMinScaling) dy = 0.0; else dy = fmin(dy, 0.0);
            } else if (l < -kOsqpInfty * kMinScaling) dy = fmax(dy, 0.0);
            ra[g] = dy;
            a[0] = fmax(a[0], fabs(E[g] * dy));
            a[1] += u * fmax(dy, 0.0) + l * fmin(dy, 0.0);      // IEEE: inf*0 = NaN, exactly as in the reference build
            V[g] = E[g] * dy;
        }
        reduce<2>(a, 0x1u);
        if (a[0] > eps) {
            if (a[1] < -eps * a[0]) {
                double mx[1] = {0.0};
                for (int kg = tid; kg < d.n; kg += NT) { int i = kg / d.b, k = kg - i * d.b; mx[0] = fmax(mx[0], fabs(col_atv(i, k, V))); }   // D aty / D
                reduce<1>(mx, 0x1u);
                return mx[0] < eps * a[0];
            }
        }
        return false;
    }
    // is_dual_infeasible (auxil.c); delta_x in va (global).  Clobbers R, YV.
    __device__ bool dual_infeasible(double eps) {
        const double* D = CSM(D); const double* Q = CSM(Q); double* R = CSM(R);
        const double* LO = CSM(LO); const double* UP = CSM(UP);
        const double* va = gws + L.gVA;
        double a[2] = {0.0, 0.0};      // [nd (max), qd (sum)]
        for (int kg = tid; kg < d.n; kg += NT) { double dx = va[kg]; a[0] = fmax(a[0], fabs(D[kg] * dx)); a[1] += Q[kg] * dx; R[kg] = D[kg] * dx; }
        reduce<2>(a, 0x1u);
        if (a[0] > eps) {
            if (a[1] < -c * eps * a[0]) {
                fill_yv();
                __syncthreads();
                double mx[1] = {0.0};
                for (int kg = tid; kg < d.n; kg += NT) { int i = kg / d.b, k = kg - i * d.b; mx[0] = fmax(mx[0], fabs(col_pu(i, k) / D[kg])); }
                reduce<1>(mx, 0x1u);
                if (mx[0] < c * eps * a[0]) {
                    bool bad = false;
                    for (int t = tid; t < d.m; t += NT) {
                        int i, r; int g = row_task(t, i, r);
                        double adx = row_dot(i, r);
                        if (((UP[g] < kOsqpInfty * kMinScaling) && (adx > eps * a[0])) || ((LO[g] > -kOsqpInfty * kMinScaling) && (adx < -eps * a[0]))) bad = true;
                    }
                    return !__syncthreads_or(bad ? 1 : 0);
                }
            }
        }
        return false;
    }
    // check_termination (auxil.c)
    __device__ bool check_termination(const InfoNorms& I, bool approximate, int& status, double& obj) {
        double eps_abs = p.eps_abs, eps_rel = p.eps_rel, epi = p.eps_prim_inf, edi = p.eps_dual_inf;
        if (I.pri > kOsqpInfty || I.dua > kOsqpInfty) { status = OSQP_NON_CVX; obj = NAN; return true; }
        if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
        bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
        double eps_prim = eps_abs + eps_rel * fmax(I.nz, I.nAx);
        if (I.pri < eps_prim) prim_ok = true; else prim_inf = primal_infeasible(epi);
        double eps_dual = eps_abs + eps_rel * (1.0 / c) * fmax(fmax(I.nq, I.nAty), I.nPx);
        if (I.dua < eps_dual) dual_ok = true; else dual_inf = dual_infeasible(edi);
        if (prim_ok && dual_ok) { status = approximate ? OSQP_SOLVED_INACCURATE : OSQP_SOLVED; return true; }
        if (prim_inf) { status = approximate ? OSQP_PRIMAL_INFEASIBLE_INACCURATE : OSQP_PRIMAL_INFEASIBLE; obj = kOsqpInfty; return true; }
        if (dual_inf) { status = approximate ? OSQP_DUAL_INFEASIBLE_INACCURATE : OSQP_DUAL_INFEASIBLE; obj = -kOsqpInfty; return true; }
        return false;
    }

    // ---- the whole LOptimizer::run for one controller --------------------------------------------------------------------------
    __device__ void solve(const Out& o) {
        const long long ib = inst;
        double* X = CSM(X); double* Z = CSM(Z); double* Y = CSM(Y); double* V = CSM(V); double* R = CSM(R); double* T = CSM(T);
        const double* D = CSM(D); const double* Q = CSM(Q); const double* E = CSM(E); const double* LO = CSM(LO); const double* UP = CSM(UP);
        double* va = gws + L.gVA; double* px = gws + L.gPX; double* ra = gws + L.gRA; double* rb = gws + L.gRB; double* rc = gws + L.gRC;
        long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 8; ++k) pw[k] = 0;
        long long t0 = clock64(), t1;
#define PROF(slot) do { t1 = clock64(); pt[slot] += t1 - t0; t0 = t1; } while (0)
        bool valid = setup_and_scale();
        PROF(0);
        int status = OSQP_UNSOLVED; double obj = 0; int iters = 0, rho_updates = 0, status_polish = 0;
        double rho = fmin(fmax(p.rho, kRhoMin), kRhoMax);
        set_rho(rho);
        if (valid) valid = factorize(p.sigma);
        PROF(1);
        if (!valid) {   // osqp_setup would have failed: LOptimizer.hpp:348-361 failure semantics
            for (int k = tid; k < d.nu; k += NT) o.cmd[ib * d.nu + k] = o.prev_cmd[ib * d.nu + k];
            if (tid == 0) { o.cost[inst] = INFINITY; o.status[inst] = RS_ERROR; o.solver_status[inst] = B200_SETUP_ERROR;
                            o.feasible[inst] = 0; o.iters[inst] = 0; o.rho_updates[inst] = 0; o.polish[inst] = 0; }
            if (o.seq_state) for (int e = tid; e < (d.ph + 1) * d.nx; e += NT) o.seq_state[ib * (d.ph + 1) * d.nx + e] = 0;
            if (o.seq_input) for (int e = tid; e < (d.ph + 1) * d.nu; e += NT) o.seq_input[ib * (d.ph + 1) * d.nu + e] = 0;
            if (o.seq_output) for (int e = tid; e < (d.ph + 1) * d.ny; e += NT) o.seq_output[ib * (d.ph + 1) * d.ny + e] = 0;
            __syncthreads();
            return;
        }
        // cold / warm start (osqp_warm_start: x <- Dinv x, y <- c Einv y, z <- A x)
        if (pr.warm) {
            for (int kg = tid; kg < d.n; kg += NT) {
                int i = kg / d.b, k = kg - i * d.b;
                double v = (k < d.bcount(i)) ? __ldg(pr.warm_x + ib * d.n + ref_var(d, i, k)) / D[kg] : 0.0;
                X[kg] = v; R[kg] = D[kg] * v;
            }
            for (int g = tid; g < d.m; g += NT) { int i, r; row_of(g, i, r); Y[g] = c * (__ldg(pr.warm_y + ib * d.m + ref_row(d, i, r)) / E[g]); }
            __syncthreads();
            for (int t = tid; t < d.m; t += NT) { int i, r; int g = row_task(t, i, r); Z[g] = E[g] * row_dot(i, r); }
            __syncthreads();
        } else {
            for (int kg = tid; kg < d.n; kg += NT) X[kg] = 0.0;
            for (int g = tid; g < d.m; g += NT) { Z[g] = 0.0; Y[g] = 0.0; }
            __syncthreads();
        }
        refresh_V();
        InfoNorms I; I.pri = I.dua = 0; I.xPx = I.qx = 0;
        bool can_check = false, done = false;
        const double sigma = p.sigma;
        const int8_t* rt = rtp();
        int it = 1;
        for (; it <= p.max_iter; ++it) {
            if (time_limit > 0.0) {       // osqp.c: run time (set-up included) against settings->time_limit at the top of every iteration
                int* flag = reinterpret_cast<int*>(CSM(RED));
                if (tid == 0) {
                    long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    flag[0] = ((double)(now - t_start) * 1e-9 >= time_limit) ? 1 : 0;
                }
                __syncthreads();
                const int hit = flag[0];
                __syncthreads();
                if (hit) { status = OSQP_TIME_LIMIT_REACHED; break; }
            }
            can_check = p.check_termination && (it % p.check_termination == 0);
            const bool can_adapt = p.adaptive_rho && p.adaptive_rho_interval && (it % p.adaptive_rho_interval == 0);
            const bool store_delta = can_check || it == p.max_iter;
            bulk_iteration(store_delta);
            PROF(2);
            if (can_check || can_adapt) {
                I = info_pass<0, 0>();
                bool stop = can_check && check_termination(I, false, status, obj);
                PROF(3);
                if (stop) { done = true; break; }
                if (can_adapt) {   // compute_rho_estimate (auxil.c) on the SCALED residuals
                    double prn = I.spri / (fmax(I.snz, I.snAx) + 1e-10);
                    double drn = I.sdua / (fmax(fmax(I.snq, I.snAty), I.snPx) + 1e-10);
                    double est = rho * sqrt(prn / (drn + 1e-10));
                    est = fmin(fmax(est, kRhoMin), kRhoMax);
                    if (est > rho * p.adaptive_rho_tolerance || est < rho / p.adaptive_rho_tolerance) {
                        rho = est; set_rho(rho);
                        factorize(sigma);
                        refresh_V();           // new rho (and the factorisation used V as scratch)
                        PROF(1);
                        ++rho_updates;
                    }
                }
            }
        }
        const bool timed_out = status == OSQP_TIME_LIMIT_REACHED;
        iters = done ? it : (timed_out ? it - 1 : p.max_iter);
        if (!done && !can_check) {      // osqp.c: update_info + check_termination when the last iteration did not run them
            I = info_pass<0, 0>();
            check_termination(I, false, status, obj);
        }
        bool has_solution = !(status == OSQP_PRIMAL_INFEASIBLE || status == OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
                              status == OSQP_DUAL_INFEASIBLE || status == OSQP_DUAL_INFEASIBLE_INACCURATE || status == OSQP_NON_CVX);
        if (has_solution) obj = (0.5 * I.xPx + I.qx) / c;
        if (status == OSQP_UNSOLVED) {
            if (!check_termination(I, true, status, obj)) status = OSQP_MAX_ITER_REACHED;
        }
        if (status == OSQP_TIME_LIMIT_REACHED) {       // osqp.c: try the approximate test, otherwise the status stays TIME_LIMIT_REACHED
            if (!check_termination(I, true, status, obj)) status = OSQP_TIME_LIMIT_REACHED;
        }
        bool use_px = false;
        PROF(3);
        // ---------------- polish (polish.c) ----------------
        if (p.polish && status == OSQP_SOLVED) {
            int8_t* rtw = reinterpret_cast<int8_t*>(smem + L.oRT);
            for (int g = tid; g < d.m; g += NT) {
                double z = Z[g], y = Y[g], l = LO[g], u = UP[g];
                bool low = (z - l) < -y;
                bool upp = !low && ((u - z) < y);
                rtw[g] = (low || upp) ? 1 : 0;
                double ba = low ? l : (upp ? u : 0.0);
                ra[g] = ba; rc[g] = ba;
            }
            for (int kg = tid; kg < d.n; kg += NT) va[kg] = -Q[kg];
            rsel[0] = 0.0; rsel[1] = 1.0 / p.delta; rsel[2] = 0.0;
            __syncthreads();
            PROF(4);
            bool fok = factorize(p.delta);
            PROF(5);
            if (fok) {
                for (int rf = 0; rf <= p.polish_refine_iter; ++rf) {
                    const bool first = rf == 0;
                    if (!first) {
                        // r1 = -q - P px - A' pnu ;  r2 = act (b - A px)
                        for (int kg = tid; kg < d.n; kg += NT) R[kg] = D[kg] * px[kg];
                        for (int g = tid; g < d.m; g += NT) V[g] = E[g] * rb[g];
                        __syncthreads();
                        fill_yv();
                        for (int t = tid; t < d.m; t += NT) {
                            int i, r; int g = row_task(t, i, r);
                            rc[g] = rt[g] ? (ra[g] - E[g] * row_dot(i, r)) : 0.0;
                        }
                        __syncthreads();
                        for (int kg = tid; kg < d.n; kg += NT) {
                            int i = kg / d.b, k = kg - i * d.b;
                            if (k < d.bcount(i)) va[kg] = -Q[kg] - col_pu(i, k) - D[kg] * col_atv(i, k, V);
                        }
                        __syncthreads();
                    }
                    // rhs = r1 + D A'(E w r2), w = act / delta
                    for (int g = tid; g < d.m; g += NT) V[g] = E[g] * (rho_of(rt[g]) * rc[g]);
                    __syncthreads();
                    for (int kg = tid; kg < d.n; kg += NT) { int i = kg / d.b, k = kg - i * d.b; if (k < d.bcount(i)) R[kg] = va[kg] + D[kg] * col_atv(i, k, V); else R[kg] = 0.0; }
                    __syncthreads();
                    kkt_solve([&](int kg, double xt) { px[kg] = first ? xt : (px[kg] + xt); });
                    for (int t = tid; t < d.m; t += NT) {
                        int i, r; int g = row_task(t, i, r);
                        double dnu = rho_of(rt[g]) * (E[g] * row_dot(i, r) - rc[g]);
                        rb[g] = first ? dnu : rb[g] + dnu;
                    }
                    __syncthreads();
                }
                InfoNorms P = info_pass<1, 1>();
                bool okp = (P.pri < I.pri && P.dua < I.dua) || (P.pri < I.pri && I.dua < 1e-10) || (P.dua < I.dua && I.pri < 1e-10);
                if (okp) { obj = (0.5 * P.xPx + P.qx) / c; status_polish = 1; use_px = true; }
                else status_polish = -1;
            } else status_polish = -1;
            __syncthreads();
        }
        PROF(6);
        // ---------------- store_solution + LOptimizer unpack ----------------
        const double cinv = 1.0 / c;
        for (int kg = tid; kg < d.n; kg += NT) {
            int i = kg / d.b, k = kg - i * d.b;
            if (k >= d.bcount(i)) continue;
            double xs = use_px ? px[kg] : X[kg];
            double xv = has_solution ? D[kg] * xs : NAN;
            T[kg] = xv;                                  // unscaled solution, stage-major
            if (o.sol_x) o.sol_x[ib * d.n + ref_var(d, i, k)] = xv;
        }
        if (o.sol_y) for (int g = tid; g < d.m; g += NT) {
            int i, r; row_of(g, i, r);
            double yv = use_px ? rc[g] : E[g] * Y[g];
            o.sol_y[ib * d.m + ref_row(d, i, r)] = has_solution ? cinv * yv : NAN;
        }
        __syncthreads();
        const double* xu = T;
        for (int k = tid; k < d.nu; k += NT) {
            int st = d.ph >= 1 ? 1 : 0;                  // sequence.input.row(0) = x_u(1)  (LOptimizer.hpp:316-327,341)
            double v = xu[st * d.b + d.nx + k];
            o.cmd[ib * d.nu + k] = v; o.prev_cmd[ib * d.nu + k] = v;
        }
        if (o.seq_state) for (int e = tid; e < (d.ph + 1) * d.nx; e += NT) { int i = e / d.nx, k = e - i * d.nx; o.seq_state[ib * (d.ph + 1) * d.nx + e] = xu[i * d.b + k]; }
        if (o.seq_input) for (int e = tid; e < (d.ph + 1) * d.nu; e += NT) {
            int i = e / d.nu, k = e - i * d.nu;
            int st = (i + 1 < d.ph + 1) ? i + 1 : i;
            o.seq_input[ib * (d.ph + 1) * d.nu + e] = xu[st * d.b + d.nx + k];
        }
        if (o.seq_output) for (int e = tid; e < (d.ph + 1) * d.ny; e += NT) {
            int i = e / d.ny, r = e - i * d.ny, j = i > 0 ? i - 1 : 0;
            double acc = 0;
            for (int k = 0; k < d.nx; ++k) acc += CSM(C)[r * fldC() + k] * xu[i * d.b + k];
            for (int q = 0; q < d.ndu; ++q) acc += ldp(pr.Dd, inst, r * d.ndu + q) * ldp(pr.uMeas, inst, j * d.ndu + q);
            o.seq_output[ib * (d.ph + 1) * d.ny + e] = acc;
        }
        if (tid == 0) {
            o.cost[inst] = obj; o.solver_status[inst] = status; o.status[inst] = to_result_status(status);
            o.feasible[inst] = (status == OSQP_SOLVED || status == OSQP_SOLVED_INACCURATE || status == OSQP_MAX_ITER_REACHED) ? 1 : 0;
            o.iters[inst] = iters; o.rho_updates[inst] = rho_updates; o.polish[inst] = status_polish;
        }
        PROF(7);
        if (o.prof && tid == 0) { for (int k = 0; k < 8; ++k) o.prof[ib * 16 + k] = pt[k]; o.prof[ib * 16 + 8] = pw[0]; o.prof[ib * 16 + 9] = pw[1]; o.prof[ib * 16 + 10] = pw[2]; o.prof[ib * 16 + 11] = pw[5]; o.prof[ib * 16 + 12] = 0; }
        if (o.prof && warp == 1 && lane == 0) { o.prof[ib * 16 + 13] = pw[3]; o.prof[ib * 16 + 14] = pw[2]; o.prof[ib * 16 + 15] = pw[4]; }
#undef PROF
        __syncthreads();
    }
};

// ---- persistent kernel: one CTA per SM, CTAs draw controllers from a global counter (longest-expected first when `order`) ----
template <class DM, int NT, bool FSH>
__global__ void __launch_bounds__(NT, 1) lmpc_cta_kernel(const __grid_constant__ DM d, const __grid_constant__ Params p,
                                                         const __grid_constant__ Prob pr, const __grid_constant__ Out o,
                                                         const __grid_constant__ CtaLayout L, int batch, double* gscratch, int* counter,
                                                         int model_shared, const int* order, double time_limit) {
    __shared__ int s_next;
    CtaSolver<DM, NT, FSH> S(d, p, pr, L);
    S.tid = threadIdx.x; S.lane = threadIdx.x & 31; S.warp = threadIdx.x >> 5;
    S.gws = gscratch + (size_t)blockIdx.x * L.gtotal;
    S.time_limit = time_limit;
    S.inst = 0;
    S.build_table();
    if (model_shared) S.load_model();
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(counter, 1);
        __syncthreads();
        const int k = s_next;
        __syncthreads();
        if (k >= batch) break;
        S.inst = order ? order[k] : k;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(S.t_start));
        if (!model_shared) S.load_model();
        S.solve(o);
    }
}

#undef CSM

}  // namespace b200mpc
