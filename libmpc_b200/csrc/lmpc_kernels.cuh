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
// Data layout.  Variables are stage-major: w_i = [x_i ; xu_i ; du_i] (b = nx+2nu doubles, the last stage has no du).
// Constraint rows are stage-major too: stage i owns [box(i) ; out(i) ; sc(i) ; eq(i+1) ; du(i)]; the ne rows eq(0)
// are kept apart.  With that ownership one ADMM iteration is exactly one forward sweep (rhs assembly fused with the
// block forward substitution) and one backward sweep (back substitution fused with z~ = A x~, the relaxation, the
// projection and the dual update); nothing of size m x n is ever stored.
//
// The reduced KKT matrix  H = Pbar + sigma I + Abar' diag(rho) Abar  is block tridiagonal over the stages; it is
// factorised by a block Cholesky (diagonal blocks inverted explicitly, so the sweeps are mat-vecs, not substitutions).
//
// Memory.  Per-instance state lives in a per-warp-slot workspace in global memory sized by the number of RESIDENT
// warps (not by the batch).  Everything a sweep needs of stage i is packed in two contiguous records -- a static one
// (factor blocks, D, q, E, row types, bounds) and a dynamic one (x, z, y) -- which one elected lane pulls into a
// 3-slot shared-memory ring with cp.async.bulk (TMA) + mbarrier two stages ahead of the arithmetic, so no global
// load sits on the dependent chain of a sweep.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace b200mpc {

constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4;
constexpr double kOsqpInfty = 1e30, kMinScaling = 1e-4, kMaxScaling = 1e4;
#ifndef B200_DOT_UNROLL
#define B200_DOT_UNROLL 8   // trips of 4 FMAs unrolled together (full for b <= 32): measured 69.1k vs 61.9k solves/s for 1
#endif
constexpr int kDotUnroll = B200_DOT_UNROLL;
#ifndef B200_MAX_THREADS
#define B200_MAX_THREADS 384   // up to 12 warps per CTA -> at most 170 registers per thread
#endif
#ifndef B200_RING_V
#define B200_RING_V 3          // 3 slots: 18.1 KB of shared memory per warp (12 warps/SM).  Measured alternatives (gang
                               // scheduling, batch 4096 / 32768): 2 slots + 12 warps 73k / 85k, 2 slots + 13 warps 67k / 77k,
                               // 2 slots + 14 warps (448 threads, 144 regs) 76k / 82k, against 77k / 94k for this setting
#endif
constexpr int kRingF = 2;            // factor-block ring slots
constexpr int kRingV = B200_RING_V;  // vector-record ring slots

// OSQP status_val (constants.h of v0.6.3)
enum { OSQP_DUAL_INFEASIBLE_INACCURATE = 4, OSQP_PRIMAL_INFEASIBLE_INACCURATE = 3, OSQP_SOLVED_INACCURATE = 2,
       OSQP_SOLVED = 1, OSQP_MAX_ITER_REACHED = -2, OSQP_PRIMAL_INFEASIBLE = -3, OSQP_DUAL_INFEASIBLE = -4,
       OSQP_NON_CVX = -7, OSQP_UNSOLVED = -10, B200_SETUP_ERROR = -1 };
// mpc::ResultStatus
enum { RS_SUCCESS = 0, RS_MAX_ITERATION = 1, RS_INFEASIBLE = 2, RS_ERROR = 3, RS_UNKNOWN = 4 };


// Shared-memory (per warp) and workspace (per slot) layouts, as offsets in doubles.  Kept as plain functions of the
// dimensions so that device code can rebuild every pointer from two bases (the extern __shared__ symbol + warp offset,
// and the slot's workspace pointer): pointers derived that way stay in registers and shared accesses compile to LDS.
#define B200_LAYOUT_HOST_DEVICE                                                                                    \
    __host__ __device__ int sBARS() const { return ring_doubles(); }                                                \
    __host__ __device__ int model_doubles() const { return (ne * ldG + ny * ldC + ne + 1) & ~1; }                   \
    __host__ __device__ int sUXC() const { return sBARS() + kRingF + kRingV + 1; }                                                      \
    __host__ __device__ int sUXN() const { return sUXC() + b; }                                                     \
    __host__ __device__ int sVTMP() const { return sUXN() + b; }                                                    \
    __host__ __device__ int sTCUR() const { return sVTMP() + b; }                                                   \
    __host__ __device__ int sXCUR() const { return sTCUR() + b; }                                                   \
    __host__ __device__ int sXN() const { return sXCUR() + b; }                                                     \
    __host__ __device__ int sVEQP() const { return sXN() + ne; }                                                    \
    __host__ __device__ int sCARRY() const { return sVEQP() + ne; }                                                 \
    __host__ __device__ int sVROW() const { return sCARRY() + ne; }                                                 \
    __host__ __device__ int sYV() const { return sVROW() + RS; }                                                    \
    /* forward-only buffers alias backward-only ones: wbuf~vtmp, ux2~tcur, arow~vrow2 */                            \
    __host__ __device__ int sW() const { return sVTMP(); }                                                          \
    __host__ __device__ int sVROW2() const { return sYV() + ny; }                                                   \
    __host__ __device__ int sUX2() const { return sTCUR(); }                                                        \
    __host__ __device__ int sAROW() const { return sVROW2(); }                                                      \
    __host__ __device__ int sWST() const { return sVROW2() + RS; }                                                  \
    __host__ __device__ int smem_doubles() const { return (sWST() + ny + 2 * nu + 3) & ~1; }                                \
    __host__ __device__ size_t wVREC() const { return (size_t)(ph + 1) * FS; }                                       \
    __host__ __device__ size_t wE0() const { return wVREC() + (size_t)(ph + 1) * VSLOT; }                              \
    __host__ __device__ size_t wT0() const { return wE0() + 5 * (size_t)ne; }                                        \
    __host__ __device__ size_t wT() const { return wT0() + (ne + 7) / 8 + 1; }                                       \
    __host__ __device__ size_t wVA() const { return wT() + n; }                                                      \
    __host__ __device__ size_t wPX() const { return wVA() + n; }                                                     \
    __host__ __device__ size_t wRA() const { return wPX() + n; }                                                     \
    __host__ __device__ size_t wRB() const { return wRA() + m; }                                                     \
    __host__ __device__ size_t wRC() const { return wRB() + m; }                                                     \
    __host__ __device__ size_t wWREC() const { return wRC() + m; }                                                   \
    __host__ __device__ size_t wRUIZ() const { return wWREC() + (size_t)(ph + 1) * (ny + 2 * nu); }                                                   \
    __host__ __device__ size_t ws_doubles() const { return wRUIZ() + ruiz_doubles() + 32; }

#define B200_DERIVED_HOST_DEVICE                                                                                   \
    __host__ __device__ int roff(int i) const { return ne + i * RS; }                                              \
    __host__ __device__ int rcount(int i) const { return i < ph ? RS : RSL; }                                      \
    __host__ __device__ int bcount(int i) const { return i < ph ? b : ne; }                                        \
    __host__ __device__ size_t ruiz_doubles() const { return 2 * (size_t)n + (size_t)m + 8; }                        \
    __host__ __device__ int fac_doubles() const { return nx * nx + 2 * b * ldb + 2 * ne * ldb; }                    \
    __host__ __device__ int ringF_doubles() const { return kRingF * FS; }                                           \
    __host__ __device__ int ring_doubles() const { int r = kRingF * FS + kRingV * VSLOT, f = fac_doubles(); return ((r > f ? r : f) + 1) & ~1; } \
    B200_LAYOUT_HOST_DEVICE

// Runtime dimensions (any shape).
struct Dm {
    static constexpr bool is_static = false;
    int nx, nu, ndu, ny, ph, ch;
    int ne, b, n, m;
    int RS, RSL, oBOX, oOUT, oSC, oEQ, oDU;      // rows owned by a stage and the offsets of its groups
    int ldG, ldC, ldb;                           // odd leading dimensions (conflict-free column walks)
    int oLc, FS;                                 // factor block: packed Linv (b(b+1)/2) + Lc (ne x ldb)
    int oD, oQ, oE, oT, RT, oLO, oUP, VSS;       // static vector record [D | q | E | types | lo | up]
    int oZ, oY, DRS, VSLOT;                      // dynamic record ([x | z | y]); V-ring slot = VSS + DRS
    int M0, M1, M2, M3;                          // reference row offsets (ProblemBuilder.hpp:70-76)
    __host__ __device__ void derive() {
        ne = nx + nu; b = ne + nu;
        n = (ph + 1) * ne + ph * nu;
        m = 2 * (ph + 1) * ne + (ph + 1) * ny + ph * nu + (ph + 1);
        oBOX = 0; oOUT = ne; oSC = ne + ny; oEQ = ne + ny + 1; oDU = oEQ + ne;
        RS = oDU + nu; RSL = oEQ;
        ldG = b | 1; ldC = nx | 1; ldb = b | 1;
        oLc = (b * (b + 1) / 2 + 1) & ~1;
        FS = (oLc + ne * ldb + 1) & ~1;
        oD = 0; oQ = oD + b; oE = oQ + b; oT = oE + RS; RT = (RS + 7) / 8; oLO = oT + RT; oUP = oLO + RS;
        VSS = (oUP + RS + 1) & ~1;
        oZ = b; oY = b + RS; DRS = (b + 2 * RS + 1) & ~1; VSLOT = VSS + DRS;
        M0 = (ph + 1) * ne; M1 = 2 * (ph + 1) * ne; M2 = M1 + (ph + 1) * ny; M3 = M2 + ph * nu;
    }
    __host__ __device__ void from(const Dm& d) { *this = d; }
    B200_DERIVED_HOST_DEVICE
};

// Compile-time dimensions: identical member names, so the same device code instantiates with every inner loop bound
// known (full unrolling, constant-folded index arithmetic).  ph/ch stay runtime.
template <int NX_, int NU_, int NDU_, int NY_>
struct SDm {
    static constexpr bool is_static = true;
    static constexpr int nx = NX_, nu = NU_, ndu = NDU_, ny = NY_;
    static constexpr int ne = NX_ + NU_, b = NX_ + 2 * NU_;
    static constexpr int oBOX = 0, oOUT = ne, oSC = ne + ny, oEQ = ne + ny + 1, oDU = oEQ + ne;
    static constexpr int RS = oDU + nu, RSL = oEQ;
    static constexpr int ldG = b | 1, ldC = nx | 1, ldb = b | 1;
    static constexpr int oLc = (b * (b + 1) / 2 + 1) & ~1;
    static constexpr int FS = (oLc + ne * ldb + 1) & ~1;
    static constexpr int oD = 0, oQ = oD + b, oE = oQ + b, oT = oE + RS, RT = (RS + 7) / 8, oLO = oT + RT, oUP = oLO + RS;
    static constexpr int VSS = (oUP + RS + 1) & ~1;
    static constexpr int oZ = b, oY = b + RS, DRS = (b + 2 * RS + 1) & ~1, VSLOT = VSS + DRS;
    int ph, ch, n, m, M0, M1, M2, M3;
    __host__ __device__ void from(const Dm& d) { ph = d.ph; ch = d.ch; n = d.n; m = d.m; M0 = d.M0; M1 = d.M1; M2 = d.M2; M3 = d.M3; }
    B200_DERIVED_HOST_DEVICE
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
    long long* prof;                                            // optional [batch*8] cycle counters per phase (null = off)
};

// ---- warp helpers ---------------------------------------------------------------------------------------------
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

// ---- TMA (cp.async.bulk) + mbarrier -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

struct InfoNorms {
    double pri, dua;                 // unscaled residual norms (termination)
    double nz, nAx, nq, nAty, nPx;   // unscaled normalisers
    double spri, sdua, snz, snAx, snq, snAty, snPx;   // scaled (rho estimate)
    double xPx, qx;
};

// Per-warp context: only scalars + two bases.  Device functions rebuild their pointers locally with B200_LOCALS so that
// they live in registers (a pointer read back from this struct would be an LDL + a generic LD).
extern __shared__ __align__(16) double smem[];

template <class DM>
struct Ctx {
    const DM& d; const Params& p; const Prob& pr; int inst; int lane;
    __device__ Ctx(const DM& d_, const Params& p_, const Prob& pr_) : d(d_), p(p_), pr(pr_) {}
    int sb;                          // this warp's offset (doubles) into the dynamic shared memory
    int model_shared;
    int mb;                          // offset of the model block [G | C | s] (per warp, or one per CTA when the batch shares the model)
    double* ws;                      // this slot's workspace
    double *wD, *wq, *wE;            // Ruiz working arrays (shared memory when they fit, workspace otherwise)
    unsigned long long fres, vres;   // resident stage+1 per ring slot, 16 bits each (F ring: 3 slots, V ring: 4 slots)
    uint32_t rflags;                 // pending bits: F [0..2], V [3..6]; parity bits: F [8..10], V [11..14]
    double c;                        // cost scaling
    double rsel[3], rinv[3];         // rho by row type
    long long sp[8];                 // sweep-phase cycle counters (profiling aid)
    __device__ __forceinline__ int jcol(int i) const { return i > 0 ? i - 1 : 0; }
    __device__ __forceinline__ double wO(int i, int r) const { return ldp(pr.OW, inst, jcol(i) * d.ny + r); }
    __device__ __forceinline__ double wU(int i, int r) const { return ldp(pr.UW, inst, jcol(i) * d.nu + r); }
    __device__ __forceinline__ double wDU(int i, int r) const { return ldp(pr.DUW, inst, i * d.nu + r); }
    __device__ __forceinline__ int voff(int i) const { return i * d.b; }
};

#define B200_LOCALS(c)                                                                                             \
    const DM& d = (c).d;                                                                                           \
    const int lane = (c).lane;                                                                                     \
    double* const sm_ = smem + (c).sb;                                                                             \
    double* const ring = sm_;                                                                                      \
    double* const Pblk = sm_;                                                                                      \
    double* const Sf = Pblk + d.nx * d.nx;                                                                         \
    double* const Li = Sf + d.b * d.ldb;                                                                           \
    double* const Hc = Li + d.b * d.ldb;                                                                           \
    double* const Lcs = Hc + d.ne * d.ldb;                                                                         \
    uint64_t* const bars = reinterpret_cast<uint64_t*>(sm_ + d.sBARS());                                           \
    double* const G = smem + (c).mb;                                                                               \
    double* const Cm = G + d.ne * d.ldG;                                                                           \
    double* const sv = Cm + d.ny * d.ldC;                                                                          \
    double* uxc = sm_ + d.sUXC();                                                                                  \
    double* uxn = sm_ + d.sUXN();                                                                                  \
    double* const vtmp = sm_ + d.sVTMP();                                                                          \
    double* const tcur = sm_ + d.sTCUR();                                                                          \
    double* const xcur = sm_ + d.sXCUR();                                                                          \
    double* const xn = sm_ + d.sXN();                                                                              \
    double* const veqp = sm_ + d.sVEQP();                                                                          \
    double* const carry = sm_ + d.sCARRY();                                                                        \
    double* const vrow = sm_ + d.sVROW();                                                                          \
    double* const yv = sm_ + d.sYV();                                                                              \
    double* const wbuf = sm_ + d.sW();                                                                             \
    double* const vrow2 = sm_ + d.sVROW2();                                                                        \
    double* const ux2 = sm_ + d.sUX2();                                                                            \
    double* const arow = sm_ + d.sAROW();                                                                          \
    double* const wst = sm_ + d.sWST();                                                                            \
    double* const ws_ = (c).ws;                                                                                    \
    double* const frec = ws_;                                                                                      \
    double* const wrec = ws_ + d.wWREC();                                                                          \
    double* const srec = ws_ + d.wVREC();                                                                          \
    double* const drec = srec + d.VSS;                                                                             \
    double* const e0E = ws_ + d.wE0();                                                                             \
    double* const e0lo = e0E + d.ne;                                                                               \
    double* const e0up = e0lo + d.ne;                                                                              \
    double* const e0z = e0up + d.ne;                                                                               \
    double* const e0y = e0z + d.ne;                                                                                \
    int8_t* const e0t = reinterpret_cast<int8_t*>(ws_ + d.wT0());                                                  \
    double* const tg = ws_ + d.wT();                                                                               \
    double* const va = ws_ + d.wVA();                                                                              \
    double* const px = ws_ + d.wPX();                                                                              \
    double* const ra = ws_ + d.wRA();                                                                              \
    double* const rb = ws_ + d.wRB();                                                                              \
    double* const rc = ws_ + d.wRC();                                                                              \
    const double rs0 = (c).rsel[0], rs1 = (c).rsel[1], rs2 = (c).rsel[2];                                          \
    const double ri0 = (c).rinv[0], ri1 = (c).rinv[1], ri2 = (c).rinv[2];                                          \
    const double csc = (c).c;                                                                                      \
    (void)ring; (void)Pblk; (void)Sf; (void)Li; (void)Hc; (void)Lcs; (void)bars; (void)G; (void)Cm; (void)sv; (void)uxc; \
    (void)uxn; (void)vtmp; (void)tcur; (void)xcur; (void)xn; (void)veqp; (void)carry; (void)vrow; (void)yv; (void)srec;  \
    (void)drec; (void)frec; (void)wrec; (void)wbuf; (void)vrow2; (void)ux2; (void)arow; (void)wst; (void)e0E; (void)e0lo; (void)e0up; (void)e0z; (void)e0y; (void)e0t; (void)tg; (void)va; (void)px;       \
    (void)ra; (void)rb; (void)rc; (void)rs0; (void)rs1; (void)rs2; (void)ri0; (void)ri1; (void)ri2; (void)csc; (void)lane
#define SRP(i) (srec + (size_t)(i) * d.VSLOT)
#define FRP(i) (frec + (size_t)(i) * d.FS)
#define DRP(i) (drec + (size_t)(i) * d.VSLOT)
#define RHO_OF(ty) ((ty) == 0 ? rs0 : ((ty) == 1 ? rs1 : rs2))
#define RINV_OF(ty) ((ty) == 0 ? ri0 : ((ty) == 1 ? ri1 : ri2))

// ---- TMA rings (state packed in registers of the context) ------------------------------------------------------------
// F ring: kRingF slots of FS doubles (factor block of a stage).  V ring: kRingV slots of VSLOT doubles (static vector
// record + dynamic record of a stage).  Slot = stage mod ring size; a slot is re-armed only by the sweep that knows the
// stage it held is finished.
__device__ __forceinline__ int ring_res(unsigned long long w, int slot) { return (int)((w >> (16 * slot)) & 0xffffull) - 1; }
template <class DM>
__device__ __forceinline__ void ring_reset(Ctx<DM>& c) {
    // called after generic-proxy stores changed records: make them visible to the async proxy, drop residency
    fence_proxy_async();
    __syncwarp();
    c.fres = 0; c.vres = 0; c.rflags &= 0x7f00u;   // keep the parities, clear pending
}
template <class DM>
__device__ __forceinline__ void ringF_issue(Ctx<DM>& c, int i) {
    const DM& d = c.d;
    if (i < 0 || i > d.ph) return;
    int slot = i % kRingF;
    if (ring_res(c.fres, slot) == i) return;
    c.fres = (c.fres & ~(0xffffull << (16 * slot))) | ((unsigned long long)(i + 1) << (16 * slot));
    c.rflags |= (1u << slot);
    if (c.lane == 0) {
        double* const sm_ = smem + c.sb;
        uint64_t* const bars = reinterpret_cast<uint64_t*>(sm_ + d.sBARS());
        mbar_expect_tx(&bars[slot], (uint32_t)(d.FS * sizeof(double)));
        tma_load_1d(sm_ + slot * d.FS, c.ws + (size_t)i * d.FS, (uint32_t)(d.FS * sizeof(double)), &bars[slot]);
    }
}
template <class DM>
__device__ __forceinline__ double* ringF_acquire(Ctx<DM>& c, int i) {
    const DM& d = c.d;
    int slot = i % kRingF;
    if (ring_res(c.fres, slot) != i) ringF_issue(c, i);
    double* const sm_ = smem + c.sb;
    if (c.rflags & (1u << slot)) {
        uint64_t* const bars = reinterpret_cast<uint64_t*>(sm_ + d.sBARS());
        mbar_wait(&bars[slot], (c.rflags >> (8 + slot)) & 1u);
        c.rflags = (c.rflags ^ (1u << (8 + slot))) & ~(1u << slot);
    }
    return sm_ + slot * d.FS;
}
template <class DM>
__device__ __forceinline__ void ringV_issue(Ctx<DM>& c, int i) {
    const DM& d = c.d;
    if (i < 0 || i > d.ph) return;
    int slot = i % kRingV;
    if (ring_res(c.vres, slot) == i) return;
    c.vres = (c.vres & ~(0xffffull << (16 * slot))) | ((unsigned long long)(i + 1) << (16 * slot));
    c.rflags |= (1u << (3 + slot));
    if (c.lane == 0) {
        double* const sm_ = smem + c.sb;
        uint64_t* const bars = reinterpret_cast<uint64_t*>(sm_ + d.sBARS()) + kRingF;
        double* dst = sm_ + d.ringF_doubles() + slot * d.VSLOT;
        mbar_expect_tx(&bars[slot], (uint32_t)(d.VSLOT * sizeof(double)));
        tma_load_1d(dst, c.ws + d.wVREC() + (size_t)i * d.VSLOT, (uint32_t)(d.VSLOT * sizeof(double)), &bars[slot]);
    }
}
template <class DM>
__device__ __forceinline__ double* ringV_acquire(Ctx<DM>& c, int i) {
    const DM& d = c.d;
    int slot = i % kRingV;
    if (ring_res(c.vres, slot) != i) ringV_issue(c, i);
    double* const sm_ = smem + c.sb;
    if (c.rflags & (1u << (3 + slot))) {
        uint64_t* const bars = reinterpret_cast<uint64_t*>(sm_ + d.sBARS()) + kRingF;
        mbar_wait(&bars[slot], (c.rflags >> (11 + slot)) & 1u);
        c.rflags = (c.rflags ^ (1u << (11 + slot))) & ~(1u << (3 + slot));
    }
    return sm_ + d.ringF_doubles() + slot * d.VSLOT;
}

// ---- model to shared memory ---------------------------------------------------------------------------------
template <class DM>
__device__ void load_model(Ctx<DM>& c) {
    B200_LOCALS(c);
    for (int e = lane; e < d.ne * d.b; e += 32) {
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
        G[r * d.ldG + k] = v;
    }
    for (int e = lane; e < d.ny * d.nx; e += 32) {
        int r = e / d.nx, k = e - r * d.nx;
        Cm[r * d.ldC + k] = ldp(c.pr.C, c.inst, e);
    }
    for (int k = lane; k < d.ne; k += 32)
        sv[k] = k < d.nx ? ldp(c.pr.SX, c.inst, k) : ldp(c.pr.SU, c.inst, k - d.nx);
    __syncwarp();
}

// ---- unscaled q of stage i, variable k (ProblemBuilder.hpp:586-595); needs yv = wO*(-yRef + Dd d) in smem ------
template <class DM>
__device__ void stage_q_prepare(Ctx<DM>& c, int i) {
    B200_LOCALS(c);
    int j = c.jcol(i);
    for (int r = lane; r < d.ny; r += 32) {
        double acc = -ldp(c.pr.yRef, c.inst, j * d.ny + r);
        for (int q = 0; q < d.ndu; ++q) acc += ldp(c.pr.Dd, c.inst, r * d.ndu + q) * ldp(c.pr.uMeas, c.inst, j * d.ndu + q);
        yv[r] = c.wO(i, r) * acc;
    }
    __syncwarp();
}
template <class DM>
__device__ __forceinline__ double stage_q(Ctx<DM>& c, int i, int k) {
    B200_LOCALS(c);
    int j = c.jcol(i);
    if (k < d.nx) {
        double acc = 0;
        for (int r = 0; r < d.ny; ++r) acc += Cm[r * d.ldC + k] * yv[r];
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
template <class DM>
__device__ __forceinline__ void stage_bounds(Ctx<DM>& c, int i, int r, double& l, double& u) {
    B200_LOCALS(c);
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

// stage weights [wO | wU | wDU] of stage i -> shared memory (one batched global load; caller syncs)
template <class DM>
__device__ __forceinline__ void load_wst(Ctx<DM>& c, int i) {
    B200_LOCALS(c);
    for (int r = lane; r < d.ny + 2 * d.nu; r += 32) {
        double v;
        if (r < d.ny) v = c.wO(i, r);
        else if (r < d.ny + d.nu) v = c.wU(i, r - d.ny);
        else v = i < d.ph ? c.wDU(i, r - d.ny - d.nu) : 0.0;
        wst[r] = v;
    }
}

// ---- Pblk = C' diag(wO_i) C (nx x nx), cached across stages with identical weights -----------------------------
template <class DM>
__device__ void stage_Pblk(Ctx<DM>& c, int i, bool& valid) {
    B200_LOCALS(c);
    // wst holds this stage's weights (load_wst + sync done by the caller); yv holds the weights Pblk was built with
    bool same = valid;
    if (same) {
        bool diff = false;
        for (int r = lane; r < d.ny; r += 32) diff |= (wst[r] != yv[r]);
        same = !wany(diff);
    }
    if (same) return;
    __syncwarp();
    for (int r = lane; r < d.ny; r += 32) yv[r] = wst[r];
    __syncwarp();
    for (int e = lane; e < d.nx * d.nx; e += 32) {
        int a = e / d.nx, k = e - a * d.nx;
        double acc = 0;
        for (int r = 0; r < d.ny; ++r) acc += Cm[r * d.ldC + a] * yv[r] * Cm[r * d.ldC + k];
        Pblk[e] = acc;
    }
    __syncwarp();
    valid = true;
}
// column inf-norm of the (D-scaled, not yet c-scaled) P column k of stage i: max_j D_j |P_jk| D_k
template <class DM>
__device__ __forceinline__ double Pcol_norm(Ctx<DM>& c, int i, int k, const double* dcur) {
    B200_LOCALS(c);
    if (k < d.nx) {
        double mx = 0;
        for (int j = 0; j < d.nx; ++j) mx = fmax(mx, dcur[j] * fabs(Pblk[j * d.nx + k]));
        return mx * dcur[k];
    } else if (k < d.ne) return dcur[k] * dcur[k] * fabs(wst[d.ny + k - d.nx]);
    return dcur[k] * dcur[k] * fabs(wst[d.ny + d.nu + k - d.ne]);
}

// ---- setup: q, Ruiz equilibration (scaling.c scale_data), scaled bounds, row types ------------------------------
// Works on flat stage-major arrays wD[n], wq[n], wE[m] (shared memory when they fit beside Pblk, global otherwise) and
// scatters the result into the stage records.
template <class DM>
__device__ bool setup_and_scale(Ctx<DM>& c) {
    B200_LOCALS(c);
    double* wD = c.wD; double* wq = c.wq; double* wE = c.wE;
    for (int i = 0; i <= d.ph; ++i) {
        stage_q_prepare(c, i);
        int bi = d.bcount(i);
        for (int k = lane; k < bi; k += 32) { wq[c.voff(i) + k] = stage_q(c, i, k); wD[c.voff(i) + k] = 1.0; }
        __syncwarp();
    }
    for (int g = lane; g < d.m; g += 32) wE[g] = 1.0;
    __syncwarp();
    c.c = 1.0;
    double pending_c = 1.0;
    double* Dt = va; double* Et = ra;
    double* dcur = uxc; double* dnxt = uxn; double* erow = vrow; double* eprev = veqp;
    bool pv = false;
    for (int it = 0; it < c.p.scaling; ++it) {
        for (int i = 0; i <= d.ph; ++i) {
            int bi = d.bcount(i), rs = d.rcount(i), ro = d.roff(i), vo = c.voff(i);
            load_wst(c, i);
            for (int k = lane; k < bi; k += 32) dcur[k] = wD[vo + k];
            if (i < d.ph) for (int k = lane; k < d.ne; k += 32) dnxt[k] = wD[vo + d.b + k];
            for (int r = lane; r < rs; r += 32) erow[r] = wE[ro + r];
            if (i == 0) for (int r = lane; r < d.ne; r += 32) eprev[r] = wE[r];
            __syncwarp();
            stage_Pblk(c, i, pv);
            for (int k = lane; k < bi; k += 32) {
                double cn = c.c * Pcol_norm(c, i, k, dcur);
                double dk = dcur[k];
                if (k < d.ne) {
                    cn = fmax(cn, eprev[k] * dk);
                    cn = fmax(cn, erow[d.oBOX + k] * dk);
                    if (k < d.nx) for (int r = 0; r < d.ny; ++r) cn = fmax(cn, erow[d.oOUT + r] * fabs(Cm[r * d.ldC + k]) * dk);
                    cn = fmax(cn, erow[d.oSC] * fabs(sv[k]) * dk);
                } else cn = fmax(cn, erow[d.oDU + k - d.ne] * dk);
                if (i < d.ph) for (int r = 0; r < d.ne; ++r) cn = fmax(cn, erow[d.oEQ + r] * fabs(G[r * d.ldG + k]) * dk);
                Dt[vo + k] = 1.0 / sqrt(lim_scaling(cn));
            }
            for (int r = lane; r < rs; r += 32) {
                double e = erow[r], rn;
                if (r < d.oOUT) rn = e * dcur[r];
                else if (r < d.oSC) { rn = 0; int q = r - d.oOUT; for (int k = 0; k < d.nx; ++k) rn = fmax(rn, e * fabs(Cm[q * d.ldC + k]) * dcur[k]); }
                else if (r < d.oEQ) { rn = 0; for (int k = 0; k < d.ne; ++k) rn = fmax(rn, e * fabs(sv[k]) * dcur[k]); }
                else if (r < d.oDU) { int q = r - d.oEQ; rn = e * dnxt[q]; for (int k = 0; k < d.b; ++k) rn = fmax(rn, e * fabs(G[q * d.ldG + k]) * dcur[k]); }
                else rn = e * dcur[d.ne + r - d.oDU];
                Et[ro + r] = 1.0 / sqrt(lim_scaling(rn));
            }
            if (i == 0) for (int r = lane; r < d.ne; r += 32) Et[r] = 1.0 / sqrt(lim_scaling(eprev[r] * dcur[r]));
            __syncwarp();
            if (i < d.ph) for (int r = lane; r < d.ne; r += 32) eprev[r] = erow[d.oEQ + r];
            __syncwarp();
        }
        double psum = 0, qmax = 0;
        for (int g = lane; g < d.m; g += 32) wE[g] *= Et[g];
        for (int i = 0; i <= d.ph; ++i) {
            int bi = d.bcount(i), vo = c.voff(i);
            load_wst(c, i);
            for (int k = lane; k < bi; k += 32) {
                double dn = wD[vo + k] * Dt[vo + k];
                wD[vo + k] = dn; dcur[k] = dn;
                double qv = (wq[vo + k] * pending_c) * Dt[vo + k];
                wq[vo + k] = qv; qmax = fmax(qmax, fabs(qv));
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
    // scatter into the records: D, scaled q, E, scaled bounds, row types; validate l<=u
    bool bad = false;
    for (int i = 0; i <= d.ph; ++i) {
        int bi = d.bcount(i), rs = d.rcount(i), ro = d.roff(i), vo = c.voff(i);
        double* S = SRP(i);
        int8_t* rt = reinterpret_cast<int8_t*>(S + d.oT);
        for (int k = lane; k < bi; k += 32) { S[d.oD + k] = wD[vo + k]; S[d.oQ + k] = wq[vo + k] * pending_c; }
        for (int r = lane; r < d.ny + 2 * d.nu; r += 32) {
            double v;
            if (r < d.ny) v = c.wO(i, r);
            else if (r < d.ny + d.nu) v = c.wU(i, r - d.ny);
            else v = i < d.ph ? c.wDU(i, r - d.ny - d.nu) : 0.0;
            wrec[(size_t)i * (d.ny + 2 * d.nu) + r] = v;
        }
        for (int r = lane; r < rs; r += 32) {
            double l, u; stage_bounds(c, i, r, l, u);
            bad |= (l > u);
            double e = wE[ro + r];
            l *= e; u *= e;
            S[d.oE + r] = e; S[d.oLO + r] = l; S[d.oUP + r] = u;
            rt[r] = ((l < -kOsqpInfty * kMinScaling) && (u > kOsqpInfty * kMinScaling)) ? 0 : ((u - l < kRhoTol) ? 2 : 1);
        }
    }
    for (int r = lane; r < d.ne; r += 32) {   // eq(0) = -[x0;u0]
        double v = r < d.nx ? -__ldg(c.pr.x0 + (long long)c.inst * d.nx + r) : -__ldg(c.pr.u0 + (long long)c.inst * d.nu + (r - d.nx));
        double e = wE[r];
        v *= e;
        e0E[r] = e; e0lo[r] = v; e0up[r] = v; e0t[r] = 2;
    }
    __syncwarp();
    return !wany(bad);
}

template <class DM>
__device__ __forceinline__ void set_rho(Ctx<DM>& c, double rho) {
    c.rsel[0] = kRhoMin; c.rsel[1] = rho; c.rsel[2] = kRhoEqOverIneq * rho;
    for (int k = 0; k < 3; ++k) c.rinv[k] = 1.0 / c.rsel[k];
}

__device__ __forceinline__ void unrank_pair(int pidx, int& r, int& k) {
    r = (int)((sqrt(8.0 * pidx + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= pidx) ++r;
    while (r * (r + 1) / 2 > pidx) --r;
    k = pidx - r * (r + 1) / 2;
}

// ---- block tridiagonal Cholesky of H = D (c P + A' R' A) D + sigma I ---------------------------------------------
// rho_of(rtype) gives the row weight; polish passes {0, 1/delta, -} with rtype = activity.  Returns false on a
// non-positive pivot.  Uses the ring area as scratch, so the ring is reset on exit.
template <class DM>
__device__ bool factorize(Ctx<DM>& c, double sigma) {
    B200_LOCALS(c);
    const int ldb = d.ldb;
    double* rw = vrow;       // rho' = rho E^2 of this stage's rows
    double* rwp = veqp;      // rho' of eq(i) rows (owned by the previous stage)
    double* dw = uxc;        // D of this stage
    double* dn = uxn;        // D of e_{i+1}
    bool ok = true;
    for (int i = 0; i <= d.ph; ++i) {
        int bi = d.bcount(i), rs = d.rcount(i);
        const int bprev = d.b;
        double* Sg = SRP(i);
        double* Fg = FRP(i);
        const int8_t* rt = reinterpret_cast<const int8_t*>(Sg + d.oT);
        for (int r = lane; r < rs; r += 32) { double e = Sg[d.oE + r]; rw[r] = RHO_OF(rt[r]) * e * e; }
        if (i == 0) for (int r = lane; r < d.ne; r += 32) { double e = e0E[r]; rwp[r] = RHO_OF(e0t[r]) * e * e; }
        for (int k = lane; k < bi; k += 32) dw[k] = Sg[d.oD + k];
        for (int r = lane; r < d.ny + 2 * d.nu; r += 32) wst[r] = wrec[(size_t)i * (d.ny + 2 * d.nu) + r];
        if (i < d.ph) { const double* Sn = SRP(i + 1); for (int k = lane; k < d.ne; k += 32) dn[k] = Sn[d.oD + k]; }
        __syncwarp();
        for (int r = lane; r < d.ny; r += 32) yv[r] = csc * wst[r] + rw[d.oOUT + r];
        __syncwarp();
        int npairs = bi * (bi + 1) / 2;
        for (int pidx = lane; pidx < npairs; pidx += 32) {
            int r, k; unrank_pair(pidx, r, k);
            double v = 0;
            if (i < d.ph) for (int j = 0; j < d.ne; ++j) v += G[j * d.ldG + r] * rw[d.oEQ + j] * G[j * d.ldG + k];
            if (r < d.ne) {
                v += rw[d.oSC] * sv[r] * sv[k];
                if (r < d.nx) for (int j = 0; j < d.ny; ++j) v += Cm[j * d.ldC + r] * yv[j] * Cm[j * d.ldC + k];
                if (r == k) {
                    v += rwp[k] + rw[d.oBOX + k];
                    if (k >= d.nx) v += csc * wst[d.ny + k - d.nx];
                }
            } else if (r == k) v += csc * wst[d.ny + d.nu + k - d.ne] + rw[d.oDU + k - d.ne];
            v = dw[r] * v * dw[k];
            if (r == k) v += sigma;
            if (i > 0 && r < d.ne) {   // Schur complement of the previous stage: Lc_{i-1} Lc_{i-1}'
                double acc = 0;
                for (int q = 0; q < bprev; ++q) acc += Lcs[r * ldb + q] * Lcs[k * ldb + q];
                v -= acc;
            }
            Sf[r * ldb + k] = v;
        }
        __syncwarp();
        for (int k = 0; k < bi; ++k) {       // right-looking Cholesky, in place
            double dkk = Sf[k * ldb + k];
            if (!(dkk > 0.0)) ok = false;
            double piv = sqrt(dkk), inv = 1.0 / piv;
            __syncwarp();
            for (int r = k + lane; r < bi; r += 32) Sf[r * ldb + k] = (r == k) ? piv : Sf[r * ldb + k] * inv;
            __syncwarp();
            for (int r = k + 1 + lane; r < bi; r += 32) {
                double lrk = Sf[r * ldb + k];
                for (int q = k + 1; q <= r; ++q) Sf[r * ldb + q] -= lrk * Sf[q * ldb + k];
            }
            __syncwarp();
        }
        for (int col = lane; col < bi; col += 32) {     // inverse of L: lane = column
            for (int r = 0; r < bi; ++r) {
                double v;
                if (r < col) v = 0.0;
                else if (r == col) v = 1.0 / Sf[r * ldb + r];
                else {
                    double acc = 0;
                    for (int q = col; q < r; ++q) acc += Sf[r * ldb + q] * Li[q * ldb + col];
                    v = -acc / Sf[r * ldb + r];
                }
                Li[r * ldb + col] = v;
            }
        }
        __syncwarp();
        for (int pidx = lane; pidx < npairs; pidx += 32) {
            int r, k; unrank_pair(pidx, r, k);
            Fg[pidx] = Li[r * ldb + k];
        }
        if (i < d.ph) {
            // Hc = -(D_e(i+1) rho'_eq(i+1)) G Dw ;  Lc = Hc Li'
            for (int e = lane; e < d.ne * bi; e += 32) {
                int r = e / bi, k = e - r * bi;
                Hc[r * ldb + k] = -(dn[r] * rw[d.oEQ + r]) * G[r * d.ldG + k] * dw[k];
            }
            __syncwarp();
            for (int e = lane; e < d.ne * bi; e += 32) {
                int r = e / bi, k = e - r * bi;
                double acc = 0;
                for (int q = 0; q <= k; ++q) acc += Hc[r * ldb + q] * Li[k * ldb + q];
                Lcs[r * ldb + k] = acc;
                Fg[d.oLc + r * ldb + k] = acc;
            }
            __syncwarp();
            for (int r = lane; r < d.ne; r += 32) rwp[r] = rw[d.oEQ + r];
        }
        __syncwarp();
    }
    ring_reset(c);
    return !wany(!ok);
}

// dot product of a strided shared-memory vector with a contiguous one, two independent accumulators
__device__ __forceinline__ double sdot(const double* a, int sa, const double* x, int nn) {
    double a0 = 0, a1 = 0;
    int q = 0;
    for (; q + 1 < nn; q += 2) { a0 = fma(a[q * sa], x[q], a0); a1 = fma(a[(q + 1) * sa], x[q + 1], a1); }
    if (q < nn) a0 = fma(a[q * sa], x[q], a0);
    return a0 + a1;
}

// predicated dot product, 4 independent accumulators; `maxn` is a compile-time bound for static dimensions
__device__ __forceinline__ double dotp_inl(const double* a, int sa, const double* x, int cnt, int maxn) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll (kDotUnroll)
    for (int q = 0; q < maxn; q += 4) {
        if (q < cnt) a0 = fma(a[q * sa], x[q], a0);
        if (q + 1 < cnt) a1 = fma(a[(q + 1) * sa], x[q + 1], a1);
        if (q + 2 < cnt) a2 = fma(a[(q + 2) * sa], x[q + 2], a2);
        if (q + 3 < cnt) a3 = fma(a[(q + 3) * sa], x[q + 3], a3);
    }
    return (a0 + a1) + (a2 + a3);
}
#ifndef B200_DOT_CALL
#define B200_DOT_CALL 0   // 1: the unrolled dot products become shared subroutines (one copy per length class) instead of being
#endif                    //    inlined.  Measured (gang scheduling, batch 32768): 82k solves/s against 94k inlined -- the call overhead
                          //    costs more than the 0.5k instructions of footprint it saves; kept as a knob
template <int MAXN>
__device__ __noinline__ double dotp_call(const double* a, int sa, const double* x, int cnt) { return dotp_inl(a, sa, x, cnt, MAXN); }
template <class DM>
__device__ __forceinline__ double dotp(const double* a, int sa, const double* x, int cnt, int maxn) {
#if B200_DOT_CALL
    if constexpr (DM::is_static) {          // maxn is a constant after inlining: the chain folds to one call
        if (maxn <= 12) return dotp_call<12>(a, sa, x, cnt);
        if (maxn <= 16) return dotp_call<16>(a, sa, x, cnt);
        if (maxn <= 20) return dotp_call<20>(a, sa, x, cnt);
    }
#endif
    return dotp_inl(a, sa, x, cnt, maxn);
}

// ---- one reduced-KKT solve fused with the ADMM updates (MODE 0) or with the polish bookkeeping (MODE 1) -----------
//  MODE 0: rhs = sigma x - q + A'(rho z - y);  x~ = H^-1 rhs;  then x,z,y updates of osqp.c (update_x/z/y)
//  MODE 1: rhs = r1 + A'(w r2) (w = act/delta); dx = H^-1 rhs; px += dx; pnu += w (A dx - r2)      [r1=va, r2=rc, pnu=rb]
//
// Both sweeps are software pipelined so that only the block recurrence (two dependent mat-vecs per stage) is on the
// critical path: the forward sweep assembles the right-hand side of stage i+1 while stage i is substituted, the
// backward sweep applies A x~ / projection / dual update of stage i+1 while stage i is back-substituted.  Stage data
// arrives through the TMA rings (factor blocks: F ring; D,q,E,types,bounds,x,z,y: V ring).
template <class DM, int MODE>
__device__ void kkt_sweeps(Ctx<DM>& c, bool store_delta, bool first) {
    B200_LOCALS(c);
    const int ldb = d.ldb;
    const double sigma = c.p.sigma, alpha = c.p.alpha;
    long long q0 = 0, q1 = 0;
#ifdef B200_SWEEP_PROFILE
    q0 = clock64();
#endif
#ifdef B200_SWEEP_PROFILE
#define SPROF(slot) do { q1 = clock64(); c.sp[slot] += q1 - q0; q0 = q1; } while (0)
#else
#define SPROF(slot) do { (void)q0; (void)q1; } while (0)
#endif
    // =========================================== forward ===========================================
    for (int k = 0; k < kRingF; ++k) ringF_issue(c, k);
    for (int k = 0; k < kRingV; ++k) ringV_issue(c, k);
    double* vc = vrow; double* vn = vrow2;     // row weights v = E (rho z - y) of the current / next stage
    // row weights of a stage (elementwise)
    auto rows_v = [&](int j, const double* V, double* dst) {
        const double* Dy = V + d.VSS;
        const int8_t* rt = reinterpret_cast<const int8_t*>(V + d.oT);
        const int rsj = d.rcount(j), roj = d.roff(j);
        for (int r = lane; r < rsj; r += 32) {
            int ty = rt[r];
            dst[r] = MODE == 0 ? V[d.oE + r] * (RHO_OF(ty) * Dy[d.oZ + r] - Dy[d.oY + r]) : V[d.oE + r] * (RHO_OF(ty) * rc[roj + r]);
        }
    };
    // unscaled A' v for variable k of stage j: own rows in vj, the eq(j) rows (owned by stage j-1) in vprev
    auto col_au = [&](int j, int k, const double* vj, const double* vprev) {
        double au;
        if (k < d.ne) {
            au = vj[d.oBOX + k] - vprev[k] + sv[k] * vj[d.oSC];
            au += dotp<DM>(Cm + k, d.ldC, vj + d.oOUT, k < d.nx ? d.ny : 0, d.ny);
        } else au = vj[d.oDU + k - d.ne];
        if (j < d.ph) au += dotp<DM>(G + k, d.ldG, vj + d.oEQ, d.ne, d.ne);
        return au;
    };
    {   // prologue: right-hand side of stage 0 -> wbuf
        const double* V = ringV_acquire(c, 0);
        for (int r = lane; r < d.ne; r += 32) {
            int ty = e0t[r];
            veqp[r] = MODE == 0 ? e0E[r] * (RHO_OF(ty) * e0z[r] - e0y[r]) : e0E[r] * (RHO_OF(ty) * rc[r]);
        }
        rows_v(0, V, vc);
        __syncwarp();
        const double* Dy = V + d.VSS;
        for (int k = lane; k < d.bcount(0); k += 32) {
            double au = col_au(0, k, vc, veqp);
            wbuf[k] = MODE == 0 ? (sigma * Dy[k] - V[d.oQ + k] + V[d.oD + k] * au) : (va[k] + V[d.oD + k] * au);
        }
        __syncwarp();
    }
    SPROF(0);
    for (int i = 0; i <= d.ph; ++i) {
        const int bi = d.bcount(i), vo = c.voff(i);
        const double* F = ringF_acquire(c, i);
        const double* Vn = (i < d.ph) ? ringV_acquire(c, i + 1) : nullptr;
        SPROF(1);
        // [B || Q1]  t_i = Linv_i w   ||   row weights of stage i+1
        for (int k = lane; k < bi; k += 32) {
            double acc = dotp<DM>(F + k * (k + 1) / 2, 1, wbuf, k + 1, d.b);
            tcur[k] = acc; tg[vo + k] = acc;
        }
        if (i < d.ph) rows_v(i + 1, Vn, vn);
        __syncwarp();
        SPROF(2);
        // [C || Q2]  carry = Lc_i t_i   ||   assemble rhs of stage i+1;  w_{i+1} = rhs - carry
        if (i < d.ph) {
            const double* Dyn = Vn + d.VSS;
            const int bn = d.bcount(i + 1), von = c.voff(i + 1);
            for (int k = lane; k < bn; k += 32) {
                double au = col_au(i + 1, k, vn, vc + d.oEQ);
                double rhs = MODE == 0 ? (sigma * Dyn[k] - Vn[d.oQ + k] + Vn[d.oD + k] * au) : (va[von + k] + Vn[d.oD + k] * au);
                if (k < d.ne) rhs -= dotp<DM>(F + d.oLc + k * ldb, 1, tcur, d.b, d.b);
                wbuf[k] = rhs;
            }
        }
        __syncwarp();
        SPROF(3);
        double* tt = vc; vc = vn; vn = tt;
        ringF_issue(c, i + kRingF);
        ringV_issue(c, i + kRingV);
    }
    // =========================================== backward ==========================================
    double* u0 = uxc; double* u1 = uxn; double* u2 = ux2;      // D*x~ of stages i, i+1, i+2
    double* xa = xcur; double* xb = xn;                         // scaled x~ of stage i (being written) / i+1
    double tnext = 0;
    if (lane < d.bcount(d.ph)) tnext = tg[c.voff(d.ph) + lane];
    // deferred row work of stage j (uj = D x~_j, ujn = D x~_{j+1})
    auto rows_dot = [&](int j, const double* uj, const double* ujn) {       // R1: a_r . ux -> arow
        const int T = d.ny + 1 + (j < d.ph ? d.ne : 0);                      // dot rows: out, sc, eq
        for (int t = lane; t < T; t += 32) {
            const double* cp; int cnt, row; double extra = 0;
            if (t < d.ny) { cp = Cm + t * d.ldC; cnt = d.nx; row = d.oOUT + t; }
            else if (t == d.ny) { cp = sv; cnt = d.ne; row = d.oSC; }
            else { int r = t - d.ny - 1; cp = G + r * d.ldG; cnt = d.b; row = d.oEQ + r; extra = ujn[r]; }
            arow[row] = dotp<DM>(cp, 1, uj, cnt, d.b) - extra;
        }
        for (int r = lane; r < d.ne; r += 32) arow[d.oBOX + r] = uj[r];
        if (j < d.ph) for (int r = lane; r < d.nu; r += 32) arow[d.oDU + r] = uj[d.ne + r];
    };
    auto rows_upd = [&](int j, double* V) {                                  // R2: relax / project / dual update
        double* Dy = V + d.VSS;
        double* Dg = DRP(j);
        const int8_t* rt = reinterpret_cast<const int8_t*>(V + d.oT);
        const int rsj = d.rcount(j), roj = d.roff(j);
        for (int r = lane; r < rsj; r += 32) {
            int ty = rt[r];
            double a = arow[r];
            if (MODE == 0) {
                double zt = V[d.oE + r] * a;
                double zo = Dy[d.oZ + r];
                double zr = alpha * zt + (1.0 - alpha) * zo;
                double yo = Dy[d.oY + r];
                double zn = fmin(fmax(zr + RINV_OF(ty) * yo, V[d.oLO + r]), V[d.oUP + r]);
                double dy = RHO_OF(ty) * (zr - zn);
                double yn = yo + dy;
                Dy[d.oY + r] = yn; Dy[d.oZ + r] = zn; Dg[d.oY + r] = yn; Dg[d.oZ + r] = zn;
                if (store_delta) ra[roj + r] = dy;
            } else {
                double dnu = RHO_OF(ty) * (V[d.oE + r] * a - rc[roj + r]);
                rb[roj + r] = first ? dnu : rb[roj + r] + dnu;
            }
        }
    };
    SPROF(4);
    for (int i = d.ph; i >= 0; --i) {
        const int bi = d.bcount(i), vo = c.voff(i);
        const double* F = ringF_acquire(c, i);
        double* Vp = (i < d.ph) ? ringV_acquire(c, i + 1) : nullptr;
        double* Dg = DRP(i);
        SPROF(5);
        double tk = tnext;
        if (i > 0 && lane < d.b) tnext = tg[c.voff(i - 1) + lane];          // prefetch for the next step
        // [A || R1(i+1)]   u = t_i - Lc_i' x~_{i+1}   ||   row dots of stage i+1
        for (int k = lane; k < bi; k += 32) {
            double w = (k == lane) ? tk : tg[vo + k];
            if (i < d.ph) w -= dotp<DM>(F + d.oLc + k, ldb, xb, d.ne, d.ne);
            vtmp[k] = w;
        }
        if (i < d.ph) rows_dot(i + 1, u1, u2);
        __syncwarp();
        double* V = ringV_acquire(c, i);          // with a 2-slot ring this record was requested one phase ago
        double* Dy = V + d.VSS;
        SPROF(6);
        // [B || R2(i+1)]   x~_i = Linv_i' u ; x update   ||   z,y update of stage i+1
        for (int k = lane; k < bi; k += 32) {
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll (kDotUnroll)
            for (int r = 0; r < d.b; r += 4) {
                if (r >= k && r < bi) a0 = fma(F[r * (r + 1) / 2 + k], vtmp[r], a0);
                if (r + 1 >= k && r + 1 < bi) a1 = fma(F[(r + 1) * (r + 2) / 2 + k], vtmp[r + 1], a1);
                if (r + 2 >= k && r + 2 < bi) a2 = fma(F[(r + 2) * (r + 3) / 2 + k], vtmp[r + 2], a2);
                if (r + 3 >= k && r + 3 < bi) a3 = fma(F[(r + 3) * (r + 4) / 2 + k], vtmp[r + 3], a3);
            }
            double xt = (a0 + a1) + (a2 + a3);
            if (MODE == 0) {
                double xo = Dy[k];
                double xnew = alpha * xt + (1.0 - alpha) * xo;
                Dy[k] = xnew; Dg[k] = xnew;
                if (store_delta) va[vo + k] = xnew - xo;
            } else {
                px[vo + k] = first ? xt : (px[vo + k] + xt);
            }
            u0[k] = V[d.oD + k] * xt;
            xa[k] = xt;
        }
        if (i < d.ph) rows_upd(i + 1, Vp);
        __syncwarp();
        SPROF(7);
        { double* tt = u2; u2 = u1; u1 = u0; u0 = tt; }
        { double* tt = xa; xa = xb; xb = tt; }
        ringF_issue(c, i - kRingF);
        ringV_issue(c, i + 1 - kRingV);
    }
    // epilogue: rows of stage 0 and the eq(0) rows (u1 = D x~_0, u2 = D x~_1)
    {
        double* V = ringV_acquire(c, 0);
        rows_dot(0, u1, u2);
        __syncwarp();
        rows_upd(0, V);
        for (int r = lane; r < d.ne; r += 32) {
            int ty = e0t[r];
            double a = -u1[r];
            if (MODE == 0) {
                double zt = e0E[r] * a, zo = e0z[r];
                double zr = alpha * zt + (1.0 - alpha) * zo;
                double yo = e0y[r];
                double zn = fmin(fmax(zr + RINV_OF(ty) * yo, e0lo[r]), e0up[r]);
                double dy = RHO_OF(ty) * (zr - zn);
                e0y[r] = yo + dy; e0z[r] = zn;
                if (store_delta) ra[r] = dy;
            } else {
                double dnu = RHO_OF(ty) * (e0E[r] * a - rc[r]);
                rb[r] = first ? dnu : rb[r] + dnu;
            }
        }
    }
#undef SPROF
    // global x,z,y were rewritten by generic stores: order them before the TMA reads of the next sweep
    fence_proxy_async();
    __syncwarp();
}

// ---- generic structured products (vector records through the V ring) ---------------------------------------------
// rows_pass: for every row:  rowfn(i, r, g, S, Dy, a_g . ux)  with ux = D*xfn(i,k,Dy) (unscaled variable values);
// i = -1 marks the eq(0) rows (S, Dy null).
template <class DM, class XFn, class RowFn>
__device__ void rows_pass(Ctx<DM>& c, XFn xfn, RowFn rowfn) {
    B200_LOCALS(c);
    for (int k = 0; k < kRingV; ++k) ringV_issue(c, k);
    {
        const double* S0 = ringV_acquire(c, 0);
        for (int k = lane; k < d.bcount(0); k += 32) uxc[k] = S0[d.oD + k] * xfn(0, k, S0 + d.VSS);
    }
    __syncwarp();
    for (int i = 0; i <= d.ph; ++i) {
        const int rs = d.rcount(i), ro = d.roff(i);
        double* S = ringV_acquire(c, i);
        double* Dy = S + d.VSS;
        if (i < d.ph) {
            const double* Sn = ringV_acquire(c, i + 1);
            for (int k = lane; k < d.bcount(i + 1); k += 32) uxn[k] = Sn[d.oD + k] * xfn(i + 1, k, Sn + d.VSS);
        }
        __syncwarp();
        for (int r = lane; r < rs; r += 32) {
            double a;
            if (r < d.oOUT) a = uxc[r];
            else if (r < d.oSC) a = sdot(Cm + (r - d.oOUT) * d.ldC, 1, uxc, d.nx);
            else if (r < d.oEQ) a = sdot(sv, 1, uxc, d.ne);
            else if (r < d.oDU) a = sdot(G + (r - d.oEQ) * d.ldG, 1, uxc, d.b) - uxn[r - d.oEQ];
            else a = uxc[d.ne + r - d.oDU];
            rowfn(i, r, ro + r, S, Dy, a);
        }
        if (i == 0) for (int r = lane; r < d.ne; r += 32) rowfn(-1, r, r, (double*)nullptr, (double*)nullptr, -uxc[r]);
        __syncwarp();
        ringV_issue(c, i + kRingV);
        double* tt = uxc; uxc = uxn; uxn = tt;
    }
}
// cols_pass: for every variable: colfn(i, k, vo+k, S, Dy, D_k * sum_r a_r[k] v_r, c*D_k*(P ux)_k) with
// v_r = rowval(i, r, g, S, Dy) (already E-weighted) and ux = D*xfn(...) (only when WITHP)
template <class DM, bool WITHP, class XFn, class RowVal, class ColFn>
__device__ void cols_pass(Ctx<DM>& c, XFn xfn, RowVal rowval, ColFn colfn) {
    B200_LOCALS(c);
    for (int k = 0; k < kRingV; ++k) ringV_issue(c, k);
    for (int r = lane; r < d.ne; r += 32) veqp[r] = rowval(-1, r, r, (double*)nullptr, (double*)nullptr);
    for (int i = 0; i <= d.ph; ++i) {
        const int bi = d.bcount(i), rs = d.rcount(i), ro = d.roff(i), vo = c.voff(i);
        double* S = ringV_acquire(c, i);
        double* Dy = S + d.VSS;
        for (int r = lane; r < rs; r += 32) vrow[r] = rowval(i, r, ro + r, S, Dy);
        if (WITHP) for (int k = lane; k < bi; k += 32) uxc[k] = S[d.oD + k] * xfn(i, k, Dy);
        __syncwarp();
        if (WITHP) {
            const double* wr = wrec + (size_t)i * (d.ny + 2 * d.nu);
            for (int r = lane; r < d.ny + 2 * d.nu; r += 32) wst[r] = wr[r];
            __syncwarp();
            for (int r = lane; r < d.ny; r += 32) yv[r] = wst[r] * sdot(Cm + r * d.ldC, 1, uxc, d.nx);
            __syncwarp();
        }
        for (int k = lane; k < bi; k += 32) {
            double au, pu = 0;
            if (k < d.ne) {
                au = vrow[d.oBOX + k] - veqp[k] + sv[k] * vrow[d.oSC];
                if (k < d.nx) au += sdot(Cm + k, d.ldC, vrow + d.oOUT, d.ny);
            } else au = vrow[d.oDU + k - d.ne];
            if (i < d.ph) au += sdot(G + k, d.ldG, vrow + d.oEQ, d.ne);
            if (WITHP) {
                if (k < d.nx) pu = sdot(Cm + k, d.ldC, yv, d.ny);
                else if (k < d.ne) pu = wst[d.ny + k - d.nx] * uxc[k];
                else pu = wst[d.ny + d.nu + k - d.ne] * uxc[k];
                pu *= csc * S[d.oD + k];
            }
            colfn(i, k, vo + k, S, Dy, S[d.oD + k] * au, pu);
        }
        __syncwarp();
        if (i < d.ph) for (int r = lane; r < d.ne; r += 32) veqp[r] = vrow[d.oEQ + r];
        __syncwarp();
        ringV_issue(c, i + kRingV);
    }
}

// accessors usable with i == -1 (eq0 rows); e0 = base of the eq0 block [E | lo | up | z | y] in the workspace
template <class DM> __device__ __forceinline__ double rowE(Ctx<DM>& c, int i, int r, const double* S) { return i < 0 ? (c.ws + c.d.wE0())[r] : S[c.d.oE + r]; }
template <class DM> __device__ __forceinline__ double rowLO(Ctx<DM>& c, int i, int r, const double* S) { return i < 0 ? (c.ws + c.d.wE0() + c.d.ne)[r] : S[c.d.oLO + r]; }
template <class DM> __device__ __forceinline__ double rowUP(Ctx<DM>& c, int i, int r, const double* S) { return i < 0 ? (c.ws + c.d.wE0() + 2 * c.d.ne)[r] : S[c.d.oUP + r]; }
template <class DM> __device__ __forceinline__ double rowZ(Ctx<DM>& c, int i, int r, const double* Dy) { return i < 0 ? (c.ws + c.d.wE0() + 3 * c.d.ne)[r] : Dy[c.d.oZ + r]; }
template <class DM> __device__ __forceinline__ double rowY(Ctx<DM>& c, int i, int r, const double* Dy) { return i < 0 ? (c.ws + c.d.wE0() + 4 * c.d.ne)[r] : Dy[c.d.oY + r]; }
template <class DM> __device__ __forceinline__ int rowT(Ctx<DM>& c, int i, int r, const double* S) {
    return i < 0 ? reinterpret_cast<const int8_t*>(c.ws + c.d.wT0())[r] : reinterpret_cast<const int8_t*>(S + c.d.oT)[r];
}

// update_info (auxil.c): residuals + every norm the termination test and the rho estimate need.
// xfn supplies the scaled x; zy(i, r, g, S, Dy, Ax, z, y) supplies the (z,y) pair of a row.
template <class DM, class XFn, class ZY>
__device__ InfoNorms info_pass(Ctx<DM>& c, XFn xfn, ZY zy) {
    B200_LOCALS(c);
    InfoNorms I;
    double pri = 0, nz = 0, nAx = 0, spri = 0, snz = 0, snAx = 0;
    rows_pass(c, xfn, [&](int i, int r, int g, double* S, double* Dy, double a) {
        double e = rowE(c, i, r, S), einv = 1.0 / e;
        double Ax = e * a, z, y;
        zy(i, r, g, S, Dy, Ax, z, y);
        rc[g] = e * y;            // E-weighted dual for the column pass
        double pv = Ax - z;
        spri = fmax(spri, fabs(pv)); snz = fmax(snz, fabs(z)); snAx = fmax(snAx, fabs(Ax));
        pri = fmax(pri, fabs(einv * pv)); nz = fmax(nz, fabs(einv * z)); nAx = fmax(nAx, fabs(einv * Ax));
    });
    __syncwarp();
    double dua = 0, nq = 0, nAty = 0, nPx = 0, sdua = 0, snq = 0, snAty = 0, snPx = 0, xPx = 0, qx = 0;
    cols_pass<DM, true>(c, xfn, [&](int, int, int g, double*, double*) { return rc[g]; },
                        [&](int i, int k, int, double* S, double* Dy, double aty, double px) {
        double dinv = 1.0 / S[c.d.oD + k];
        double q = S[c.d.oQ + k];
        double dv = q + px + aty;
        sdua = fmax(sdua, fabs(dv)); snq = fmax(snq, fabs(q)); snAty = fmax(snAty, fabs(aty)); snPx = fmax(snPx, fabs(px));
        dua = fmax(dua, fabs(dinv * dv)); nq = fmax(nq, fabs(dinv * q)); nAty = fmax(nAty, fabs(dinv * aty)); nPx = fmax(nPx, fabs(dinv * px));
        double xv = xfn(i, k, Dy);
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
template <class DM>
__device__ bool primal_infeasible(Ctx<DM>& c, double eps) {
    B200_LOCALS(c);
    double nd = 0, lhs = 0;
    auto one = [&](int g, double l, double u, double e) {
        double dy = ra[g];
        if (u > kOsqpInfty * kMinScaling) {
            if (l < -kOsqpInfty * kMinScaling) dy = 0.0; else dy = fmin(dy, 0.0);
        } else if (l < -kOsqpInfty * kMinScaling) dy = fmax(dy, 0.0);
        ra[g] = dy;
        nd = fmax(nd, fabs(e * dy));
        lhs += u * fmax(dy, 0.0) + l * fmin(dy, 0.0);   // IEEE: inf*0 = NaN, exactly as in the reference build
    };
    for (int r = lane; r < d.ne; r += 32) one(r, e0lo[r], e0up[r], e0E[r]);
    for (int i = 0; i <= d.ph; ++i) {
        const double* S = SRP(i);
        int ro = d.roff(i);
        for (int r = lane; r < d.rcount(i); r += 32) one(ro + r, S[d.oLO + r], S[d.oUP + r], S[d.oE + r]);
    }
    nd = wmax(nd); lhs = wsum(lhs);
    __syncwarp();
    if (nd > eps) {
        if (lhs < -eps * nd) {
            double mx = 0;
            cols_pass<DM, false>(c, [&](int, int, const double*) { return 0.0; },
                                 [&](int i, int r, int g, double* S, double*) { return rowE(c, i, r, S) * ra[g]; },
                                 [&](int, int k, int, double* S, double*, double aty, double) { mx = fmax(mx, fabs(aty / S[c.d.oD + k])); });
            mx = wmax(mx);
            return mx < eps * nd;
        }
    }
    return false;
}
// is_dual_infeasible (auxil.c); delta_x lives in va
template <class DM>
__device__ bool dual_infeasible(Ctx<DM>& c, double eps) {
    B200_LOCALS(c);
    double nd = 0, qd = 0;
    for (int i = 0; i <= d.ph; ++i) {
        const double* S = SRP(i);
        int vo = c.voff(i);
        for (int k = lane; k < d.bcount(i); k += 32) { double dx = va[vo + k]; nd = fmax(nd, fabs(S[d.oD + k] * dx)); qd += S[d.oQ + k] * dx; }
    }
    nd = wmax(nd); qd = wsum(qd);
    double cs = c.c;
    auto dxfn = [&](int i, int k, const double*) { return va[c.voff(i) + k]; };
    if (nd > eps) {
        if (qd < -cs * eps * nd) {
            double mx = 0;
            cols_pass<DM, true>(c, dxfn, [&](int, int, int, double*, double*) { return 0.0; },
                                [&](int, int k, int, double* S, double*, double, double px) { mx = fmax(mx, fabs(px / S[c.d.oD + k])); });
            mx = wmax(mx);
            if (mx < cs * eps * nd) {
                bool bad = false;
                rows_pass(c, dxfn, [&](int i, int r, int, double* S, double*, double a) {
                    double adx = a;   // Einv * (E a) = a
                    if (((rowUP(c, i, r, S) < kOsqpInfty * kMinScaling) && (adx > eps * nd)) ||
                        ((rowLO(c, i, r, S) > -kOsqpInfty * kMinScaling) && (adx < -eps * nd))) bad = true;
                });
                return !wany(bad);
            }
        }
    }
    return false;
}

// check_termination (auxil.c).  Returns true when the loop must stop; status/obj updated.
template <class DM>
__device__ bool check_termination(Ctx<DM>& c, const InfoNorms& I, bool approximate, int& status, double& obj) {
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

// reference row index of a row (i,r) (i=-1: eq0) / reference variable index of (i,k)
template <class DM>
__device__ __forceinline__ int ref_row(const DM& d, int i, int r) {
    if (i < 0) return r;
    if (r < d.oOUT) return d.M0 + i * d.ne + r;
    if (r < d.oSC) return d.M1 + i * d.ny + (r - d.oOUT);
    if (r < d.oEQ) return d.M3 + i;
    if (r < d.oDU) return (i + 1) * d.ne + (r - d.oEQ);
    return d.M2 + i * d.nu + (r - d.oDU);
}
template <class DM>
__device__ __forceinline__ int ref_var(const DM& d, int i, int k) {
    return k < d.ne ? i * d.ne + k : (d.ph + 1) * d.ne + i * d.nu + (k - d.ne);
}

// ---- the whole LOptimizer::run for one instance -----------------------------------------------------------------
template <class DM>
__device__ void solve_instance(Ctx<DM>& c, const Out& o) {
    B200_LOCALS(c);
    const int inst = c.inst;
    const long long ib = inst;
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 8; ++k) c.sp[k] = 0;
    long long t0 = clock64(), t1;
#define PROF(slot) do { t1 = clock64(); pt[slot] += t1 - t0; t0 = t1; } while (0)
    if (!c.model_shared) load_model(c);
    bool valid = setup_and_scale(c);
    PROF(0);
    int status = OSQP_UNSOLVED; double obj = 0; int iters = 0, rho_updates = 0, status_polish = 0;
    double rho = fmin(fmax(c.p.rho, kRhoMin), kRhoMax);
    set_rho(c, rho);
    if (valid) valid = factorize(c, c.p.sigma);
    PROF(1);
    if (!valid) {
        // osqp_setup would have failed (validate_data / factorisation): LOptimizer.hpp:348-361 failure semantics
        ring_reset(c);
        for (int k = lane; k < d.nu; k += 32) o.cmd[ib * d.nu + k] = o.prev_cmd[ib * d.nu + k];
        if (lane == 0) { o.cost[inst] = INFINITY; o.status[inst] = RS_ERROR; o.solver_status[inst] = B200_SETUP_ERROR;
                         o.feasible[inst] = 0; o.iters[inst] = 0; o.rho_updates[inst] = 0; o.polish[inst] = 0; }
        if (o.seq_state) for (int e = lane; e < (d.ph + 1) * d.nx; e += 32) o.seq_state[ib * (d.ph + 1) * d.nx + e] = 0;
        if (o.seq_input) for (int e = lane; e < (d.ph + 1) * d.nu; e += 32) o.seq_input[ib * (d.ph + 1) * d.nu + e] = 0;
        if (o.seq_output) for (int e = lane; e < (d.ph + 1) * d.ny; e += 32) o.seq_output[ib * (d.ph + 1) * d.ny + e] = 0;
        return;
    }
    auto xrec = [&](int, int k, const double* Dy) { return Dy[k]; };
    // cold / warm start (osqp_warm_start: x <- Dinv x, y <- c Einv y, z <- A x)
    if (c.pr.warm) {
        for (int i = 0; i <= d.ph; ++i) {
            double* S = SRP(i); double* Dg = DRP(i);
            for (int k = lane; k < d.bcount(i); k += 32) Dg[k] = __ldg(c.pr.warm_x + ib * d.n + ref_var(d, i, k)) / S[d.oD + k];
            for (int r = lane; r < d.rcount(i); r += 32) Dg[d.oY + r] = c.c * (__ldg(c.pr.warm_y + ib * d.m + ref_row(d, i, r)) / S[d.oE + r]);
        }
        for (int r = lane; r < d.ne; r += 32) e0y[r] = c.c * (__ldg(c.pr.warm_y + ib * d.m + r) / e0E[r]);
        ring_reset(c);
        rows_pass(c, xrec, [&](int i, int r, int, double* S, double*, double a) {
            if (i < 0) e0z[r] = e0E[r] * a; else DRP(i)[d.oZ + r] = S[d.oE + r] * a;
        });
        ring_reset(c);
    } else {
        for (int i = 0; i <= d.ph; ++i) {
            double* Dg = DRP(i);
            for (int e = lane; e < d.DRS; e += 32) Dg[e] = 0.0;
        }
        for (int r = lane; r < d.ne; r += 32) { e0z[r] = 0.0; e0y[r] = 0.0; }
        ring_reset(c);
    }
    InfoNorms I; I.pri = I.dua = 0;
    bool can_check = false, done = false;
    auto admm_zy = [&](int i, int r, int, double*, double* Dy, double, double& z, double& y) { z = rowZ(c, i, r, Dy); y = rowY(c, i, r, Dy); };
    int it = 1;
    for (; it <= c.p.max_iter; ++it) {
        can_check = c.p.check_termination && (it % c.p.check_termination == 0);
        bool can_adapt = c.p.adaptive_rho && c.p.adaptive_rho_interval && (it % c.p.adaptive_rho_interval == 0);
        kkt_sweeps<DM, 0>(c, can_check || it == c.p.max_iter, false);
        PROF(2);
        if (can_check || can_adapt) {
            I = info_pass(c, xrec, admm_zy);
            bool stop = can_check && check_termination(c, I, false, status, obj);
            PROF(3);
            if (stop) { done = true; break; }
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
                PROF(1);
                ++rho_updates;
            }
        }
    }
    iters = done ? it : c.p.max_iter;
    if (!done && !can_check) {
        I = info_pass(c, xrec, admm_zy);
        check_termination(c, I, false, status, obj);
    }
    bool has_solution = !(status == OSQP_PRIMAL_INFEASIBLE || status == OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
                          status == OSQP_DUAL_INFEASIBLE || status == OSQP_DUAL_INFEASIBLE_INACCURATE || status == OSQP_NON_CVX);
    if (has_solution) obj = (0.5 * I.xPx + I.qx) / c.c;
    if (status == OSQP_UNSOLVED) {
        if (!check_termination(c, I, true, status, obj)) status = OSQP_MAX_ITER_REACHED;
    }
    bool use_px = false;
    PROF(3);
    // ---------------- polish (polish.c) ----------------
    if (c.p.polish && status == OSQP_SOLVED) {
        const double dinv = 1.0 / c.p.delta;
        auto mark = [&](double z, double y, double l, double u, int8_t& ty, double& bact) {
            bool low = (z - l) < -y;
            bool upp = !low && ((u - z) < y);
            ty = (low || upp) ? 1 : 0;
            bact = low ? l : (upp ? u : 0.0);
        };
        for (int i = 0; i <= d.ph; ++i) {
            double* S = SRP(i); const double* Dg = DRP(i);
            int8_t* rt = reinterpret_cast<int8_t*>(S + d.oT);
            int ro = d.roff(i);
            for (int r = lane; r < d.rcount(i); r += 32) {
                int8_t ty; double ba;
                mark(Dg[d.oZ + r], Dg[d.oY + r], S[d.oLO + r], S[d.oUP + r], ty, ba);
                rt[r] = ty; ra[ro + r] = ba; rc[ro + r] = ba;
            }
            for (int k = lane; k < d.bcount(i); k += 32) va[c.voff(i) + k] = -S[d.oQ + k];
        }
        for (int r = lane; r < d.ne; r += 32) {
            int8_t ty; double ba;
            mark(e0z[r], e0y[r], e0lo[r], e0up[r], ty, ba);
            e0t[r] = ty; ra[r] = ba; rc[r] = ba;
        }
        c.rsel[0] = 0.0; c.rsel[1] = dinv; c.rsel[2] = 0.0;
        __syncwarp();
        PROF(4);
        bool fok = factorize(c, c.p.delta);
        PROF(5);
        if (fok) {
            auto pxfn = [&](int i, int k, const double*) { return px[c.voff(i) + k]; };
            kkt_sweeps<DM, 1>(c, false, true);
            for (int rf = 0; rf < c.p.polish_refine_iter; ++rf) {
                // r1 = -q - P px - A' pnu ; r2 = act (b - A px)
                cols_pass<DM, true>(c, pxfn, [&](int i, int r, int g, double* S, double*) { return rowE(c, i, r, S) * rb[g]; },
                                    [&](int, int k, int kg, double* S, double*, double aty, double pxv) { va[kg] = -S[c.d.oQ + k] - pxv - aty; });
                rows_pass(c, pxfn, [&](int i, int r, int g, double* S, double*, double a) {
                    rc[g] = rowT(c, i, r, S) ? (ra[g] - rowE(c, i, r, S) * a) : 0.0;
                });
                __syncwarp();
                kkt_sweeps<DM, 1>(c, false, false);
            }
            // polished (z,y): z = A px, project_normalcone
            auto pol_zy = [&](int i, int r, int g, double* S, double*, double Ax, double& z, double& y) {
                double t = Ax + rb[g];
                z = fmin(fmax(t, rowLO(c, i, r, S)), rowUP(c, i, r, S));
                y = t - z;
            };
            InfoNorms P = info_pass(c, pxfn, pol_zy);
            bool okp = (P.pri < I.pri && P.dua < I.dua) || (P.pri < I.pri && I.dua < 1e-10) || (P.dua < I.dua && I.pri < 1e-10);
            if (okp) {
                obj = (0.5 * P.xPx + P.qx) / c.c;
                status_polish = 1;
                use_px = true;     // x <- px; the polished dual is rc = E*y_pol left by info_pass
            } else status_polish = -1;
        } else status_polish = -1;
        __syncwarp();
    }
    PROF(6);
    // ---------------- store_solution + LOptimizer unpack ----------------
    const double cinv = 1.0 / c.c;
    for (int i = 0; i <= d.ph; ++i) {
        const double* S = SRP(i); const double* Dg = DRP(i);
        int vo = c.voff(i), ro = d.roff(i);
        for (int k = lane; k < d.bcount(i); k += 32) {
            double xs = use_px ? px[vo + k] : Dg[k];
            double xv = has_solution ? S[d.oD + k] * xs : NAN;
            tg[vo + k] = xv;     // unscaled solution, flat stage-major
            if (o.sol_x) o.sol_x[ib * d.n + ref_var(d, i, k)] = xv;
        }
        if (o.sol_y) for (int r = lane; r < d.rcount(i); r += 32) {
            double yv = use_px ? rc[ro + r] : S[d.oE + r] * Dg[d.oY + r];
            o.sol_y[ib * d.m + ref_row(d, i, r)] = has_solution ? cinv * yv : NAN;
        }
    }
    if (o.sol_y) for (int r = lane; r < d.ne; r += 32) {
        double yv = use_px ? rc[r] : e0E[r] * e0y[r];
        o.sol_y[ib * d.m + r] = has_solution ? cinv * yv : NAN;
    }
    __syncwarp();
    const double* xu = tg;
    for (int k = lane; k < d.nu; k += 32) {
        int st = d.ph >= 1 ? 1 : 0;    // sequence.input.row(0) = x_u(1)  (LOptimizer.hpp:316-327,341)
        double v = xu[c.voff(st) + d.nx + k];
        o.cmd[ib * d.nu + k] = v; o.prev_cmd[ib * d.nu + k] = v;
    }
    if (o.seq_state) for (int e = lane; e < (d.ph + 1) * d.nx; e += 32) {
        int i = e / d.nx, k = e - i * d.nx;
        o.seq_state[ib * (d.ph + 1) * d.nx + e] = xu[c.voff(i) + k];
    }
    if (o.seq_input) for (int e = lane; e < (d.ph + 1) * d.nu; e += 32) {
        int i = e / d.nu, k = e - i * d.nu;
        int st = (i + 1 < d.ph + 1) ? i + 1 : i;
        o.seq_input[ib * (d.ph + 1) * d.nu + e] = xu[c.voff(st) + d.nx + k];
    }
    if (o.seq_output) for (int e = lane; e < (d.ph + 1) * d.ny; e += 32) {
        int i = e / d.ny, r = e - i * d.ny;
        int j = i > 0 ? i - 1 : 0;
        double acc = 0;
        for (int k = 0; k < d.nx; ++k) acc += Cm[r * d.ldC + k] * xu[c.voff(i) + k];
        for (int q = 0; q < d.ndu; ++q) acc += ldp(c.pr.Dd, inst, r * d.ndu + q) * ldp(c.pr.uMeas, inst, j * d.ndu + q);
        o.seq_output[ib * (d.ph + 1) * d.ny + e] = acc;
    }
    if (lane == 0) {
        o.cost[inst] = obj; o.solver_status[inst] = status; o.status[inst] = to_result_status(status);
        o.feasible[inst] = (status == OSQP_SOLVED || status == OSQP_SOLVED_INACCURATE || status == OSQP_MAX_ITER_REACHED) ? 1 : 0;
        o.iters[inst] = iters; o.rho_updates[inst] = rho_updates; o.polish[inst] = status_polish;
    }
    ring_reset(c);
    PROF(7);
    if (o.prof && lane == 0) { for (int k = 0; k < 8; ++k) { o.prof[ib * 16 + k] = pt[k]; o.prof[ib * 16 + 8 + k] = c.sp[k]; } }
#undef PROF
}

// ---- persistent kernel: warps pull instances from a global counter ------------------------------------------------
template <class DM>
__global__ void __launch_bounds__(B200_MAX_THREADS, 1) lmpc_solve_kernel(const __grid_constant__ DM d, const __grid_constant__ Params p,
                                                        const __grid_constant__ Prob pr, const __grid_constant__ Out o,
                                                        int batch, double* workspace, size_t ws_stride, int* counter, int model_shared,
                                                        int gang, const int* order) {
    __shared__ int gang_base, gang_take;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int slot = blockIdx.x * wpb + warp;
    Ctx<DM> c(d, p, pr);
    c.lane = lane;
    c.sb = warp * d.smem_doubles();
    c.model_shared = model_shared;
    c.mb = wpb * d.smem_doubles() + (model_shared ? 0 : warp * d.model_doubles());
    c.ws = workspace + (size_t)slot * ws_stride;
    // Ruiz working arrays: in shared memory (the ring area past Pblk) when they fit, else in the workspace
    {
        size_t need = d.ruiz_doubles();
        if ((size_t)(d.nx * d.nx) + need <= (size_t)d.ring_doubles()) c.wD = smem + c.sb + d.nx * d.nx;
        else c.wD = c.ws + d.wRUIZ();
        c.wq = c.wD + d.n; c.wE = c.wq + d.n;
    }
    if (lane == 0) {
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem + c.sb + d.sBARS());
        for (int k = 0; k < kRingF + kRingV; ++k) mbar_init(&bars[k], 1);
        fence_mbar_init();
    }
    c.fres = 0; c.vres = 0; c.rflags = 0;
    fence_proxy_async();
    __syncwarp();
    if (model_shared) {            // the whole batch shares A,B,C and the scalar row: one copy per CTA
        c.inst = 0;
        if (warp == 0) load_model(c);
        __syncthreads();
    }
    // Gang scheduling: the warps of a CTA draw their instances together and start them together, so that at any time
    // most of them execute the same phase of the (large) program and share its instruction-cache footprint; a CTA waits
    // for its slowest member before drawing again.  Free scheduling: every warp draws on its own.
    int gang_round = 0;
    int sub = 0;                      // gang > 1: every member runs `gang` instances back to back between two CTA barriers
    for (;;) {
        int inst = 0;
        if (gang) {
            if (sub == 0) {
                __syncthreads();
                if (threadIdx.x == 0) {
                    // full gangs for the rounds every CTA can fill; the remainder is split evenly over the CTAs instead of
                    // leaving most of them idle in the last round
                    const int per_round = (int)gridDim.x * wpb * gang;
                    int take = wpb * gang;
                    if (gang_round >= batch / per_round) {
                        take = (batch % per_round + (int)gridDim.x - 1) / (int)gridDim.x;
                        take = take < 1 ? 1 : (take > wpb * gang ? wpb * gang : take);
                    }
                    gang_take = take;
                    gang_base = atomicAdd(counter, take);
                }
                ++gang_round;
                __syncthreads();
                if (gang_base >= batch) break;
            }
            const int k = sub * wpb + warp;
            inst = k < gang_take ? gang_base + k : batch;
            if (++sub == gang) sub = 0;
        } else {
            if (lane == 0) inst = atomicAdd(counter, 1);
            inst = __shfl_sync(0xffffffffu, inst, 0);
            if (inst >= batch) break;
        }
        if (inst < batch) {
            c.inst = order ? order[inst] : inst;      // drawing order (longest-expected first), see order_by_history_kernel
            solve_instance(c, o);
        }
        __syncwarp();
    }
}

// Drawing order for the next solve: instances sorted by the iteration count of their PREVIOUS solve, longest first
// (counting sort on iters / check_termination; one CTA).  Consecutive MPC steps of a controller need similar iteration
// counts, so (i) the members of a gang finish together instead of waiting for one straggler and (ii) the long instances
// start first (longest-processing-time-first).  Results do not depend on the order.
static __global__ void order_by_history_kernel(const int* iters, int batch, int bucket, int* order) {
    constexpr int NB = 1024;
    __shared__ int hist[NB];
    for (int k = threadIdx.x; k < NB; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    auto key = [&](int it) { int b = (it < 0 ? 0 : it) / bucket; return NB - 1 - (b >= NB ? NB - 1 : b); };    // descending
    for (int i = threadIdx.x; i < batch; i += blockDim.x) atomicAdd(&hist[key(iters[i])], 1);
    __syncthreads();
    if (threadIdx.x == 0) { int acc = 0; for (int k = 0; k < NB; ++k) { int v = hist[k]; hist[k] = acc; acc += v; } }
    __syncthreads();
    for (int i = threadIdx.x; i < batch; i += blockDim.x) order[atomicAdd(&hist[key(iters[i])], 1)] = i;
}

}  // namespace b200mpc
