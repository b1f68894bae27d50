/*
 * lmpc_oracle.c -- CPU oracle (TEST INFRASTRUCTURE / CPU BASELINE ONLY), plain C restatement of the reference's LMPC
 * solve path.  Never linked into the product (libmpc_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg call it.
 *
 * It does the same work per control step as the reference does (paths relative to /root/reference):
 *   1. ProblemBuilder::buildTimeInvariantTems: dense P (n x n) and A (m x n)      include/mpc/LMPC/ProblemBuilder.hpp:642-825
 *      (done at set-up, like the reference's setters)
 *   2. ProblemBuilder::get: per-step q,l,u                                          include/mpc/LMPC/ProblemBuilder.hpp:528-633
 *   3. Problem::getSparse + createOsqpSparseMatrix: per-step dense -> CSC scan       ProblemBuilder.hpp:54-67, LOptimizer.hpp:425-478
 *   4. osqp_setup / osqp_solve (OSQP v0.6.3, restated: Ruiz scaling, KKT assembly, sparse LDL', ADMM, adaptive rho,
 *      termination / infeasibility tests, polish)                                    include/mpc/LMPC/LOptimizer.hpp:241-284
 *   5. unpack                                                                       include/mpc/LMPC/LOptimizer.hpp:292-347
 *
 * OSQP and its QDLDL factorisation are third-party code that is NOT under /root/reference (pinned v0.6.3 by
 * configure.sh:36-38); both are restated here from their published algorithms (Stellato et al. 2020; the up-looking
 * sparse LDL' with elimination tree of Davis' "Direct methods for sparse linear systems"), not copied.  Two stated
 * deviations: (a) adaptive_rho_interval is pinned to 25 (v0.6.3 derives it from wall-clock time), (b) the fill-reducing
 * ordering is a stage-interleaved permutation computed in O(n+m) instead of AMD (cheaper than AMD, so the baseline is
 * if anything favoured).  Pinned against the numpy oracle and the reference's golden vector in tests/test_c_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_EQ_OVER_RHO_INEQ 1e3
#define RHO_TOL 1e-4
#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4

enum { ST_DUAL_INF_INACC = 4, ST_PRIM_INF_INACC = 3, ST_SOLVED_INACC = 2, ST_SOLVED = 1, ST_MAX_ITER = -2, ST_PRIM_INF = -3,
       ST_DUAL_INF = -4, ST_NON_CVX = -7, ST_UNSOLVED = -10 };

typedef struct {
    int nx, nu, ndu, ny, ph, ch;
} dims_t;

typedef struct {
    int max_iter, adaptive_rho, polish, scaling, check_termination, adaptive_rho_interval, polish_refine_iter, warm_start;
    double alpha, rho, sigma, delta, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
} params_t;

/* problem description, same layouts as include/b200mpc.h (row-major model, stage-major horizon matrices [ph][dim],
 * input bounds already expanded to [ph][nu]) */
typedef struct {
    const double *A, *B, *C, *Bd, *Dd, *OW, *UW, *DUW, *XMin, *XMax, *YMin, *YMax, *UMin, *UMax, *SMin, *SMax, *SX, *SU,
        *yRef, *uRef, *duRef, *uMeas;
} prob_t;

typedef struct {
    double cost;
    int status, solver_status, is_feasible, iters, rho_updates, status_polish;
} result_t;

/* ------------------------------------------------------------------------------------------------------------ */
typedef struct { int m, n, nnz; int *p, *i; double* x; } csc_t;

static csc_t* csc_alloc(int m, int n, int nnz) {
    csc_t* M = (csc_t*)malloc(sizeof(csc_t));
    M->m = m; M->n = n; M->nnz = nnz;
    M->p = (int*)malloc(sizeof(int) * (n + 1));
    M->i = (int*)malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
    M->x = (double*)malloc(sizeof(double) * (nnz > 0 ? nnz : 1));
    return M;
}
static void csc_free(csc_t* M) { if (M) { free(M->p); free(M->i); free(M->x); free(M); } }

/* dense (column-major) -> CSC, optionally upper triangle only: the per-step sparseView() scan of the reference */
static csc_t* dense_to_csc(const double* D, int m, int n, int upper) {
    int nnz = 0;
    for (int j = 0; j < n; ++j) { int lim = upper ? (j + 1 < m ? j + 1 : m) : m; for (int i = 0; i < lim; ++i) if (D[(size_t)j * m + i] != 0.0) ++nnz; }
    csc_t* M = csc_alloc(m, n, nnz);
    int k = 0;
    for (int j = 0; j < n; ++j) {
        M->p[j] = k;
        int lim = upper ? (j + 1 < m ? j + 1 : m) : m;
        for (int i = 0; i < lim; ++i) { double v = D[(size_t)j * m + i]; if (v != 0.0) { M->i[k] = i; M->x[k] = v; ++k; } }
    }
    M->p[n] = k;
    return M;
}

/* ---- the formulation (ProblemBuilder) ------------------------------------------------------------------------ */
typedef struct {
    dims_t d; int ne, n, m, meq;
    double *ssA, *ssB, *ssC, *ssBv, *ssDv;    /* row-major: ssA ne x ne, ssB ne x nu, ssC (ny+nu) x ne, ssBv ne x ndu, ssDv (ny+nu) x ndu */
    double *P, *Amat;                          /* dense column-major n x n, m x n */
    double *lineq, *uineq;                     /* m_ineq */
} form_t;

static int jcol(int i) { return i > 0 ? i - 1 : 0; }

static form_t* form_build(const dims_t* dd, const prob_t* pr) {
    form_t* f = (form_t*)calloc(1, sizeof(form_t));
    f->d = *dd;
    int nx = dd->nx, nu = dd->nu, ndu = dd->ndu, ny = dd->ny, ph = dd->ph, ch = dd->ch;
    int ne = nx + nu; f->ne = ne;
    int n = (ph + 1) * ne + ph * nu, meq = (ph + 1) * ne;
    int mineq = (ph + 1) * ne + (ph + 1) * ny + ph * nu + (ph + 1);
    int m = meq + mineq;
    f->n = n; f->m = m; f->meq = meq;
    f->ssA = (double*)calloc((size_t)ne * ne, 8); f->ssB = (double*)calloc((size_t)ne * nu + 1, 8);
    f->ssC = (double*)calloc((size_t)(ny + nu) * ne + 1, 8); f->ssBv = (double*)calloc((size_t)ne * ndu + 1, 8);
    f->ssDv = (double*)calloc((size_t)(ny + nu) * ndu + 1, 8);
    for (int r = 0; r < nx; ++r) {
        for (int k = 0; k < nx; ++k) f->ssA[r * ne + k] = pr->A[r * nx + k];
        for (int k = 0; k < nu; ++k) { f->ssA[r * ne + nx + k] = pr->B[r * nu + k]; f->ssB[r * nu + k] = pr->B[r * nu + k]; }
        for (int k = 0; k < ndu; ++k) f->ssBv[r * ndu + k] = pr->Bd[r * ndu + k];
    }
    for (int j = 0; j < nu; ++j) { f->ssA[(nx + j) * ne + nx + j] = 1.0; f->ssB[(nx + j) * nu + j] = 1.0; f->ssC[(ny + j) * ne + nx + j] = 1.0; }
    for (int r = 0; r < ny; ++r) {
        for (int k = 0; k < nx; ++k) f->ssC[r * ne + k] = pr->C[r * nx + k];
        for (int k = 0; k < ndu; ++k) f->ssDv[r * ndu + k] = pr->Dd[r * ndu + k];
    }
    f->P = (double*)calloc((size_t)n * n, 8);
    f->Amat = (double*)calloc((size_t)m * n, 8);
    f->lineq = (double*)calloc(mineq, 8); f->uineq = (double*)calloc(mineq, 8);
#define PM(r, c) f->P[(size_t)(c) * n + (r)]
#define AM(r, c) f->Amat[(size_t)(c) * m + (r)]
    double* w = (double*)malloc(sizeof(double) * (ny + nu));
    for (int i = 0; i <= ph; ++i) {
        int j = jcol(i);
        for (int r = 0; r < ny; ++r) w[r] = pr->OW[j * ny + r];
        for (int r = 0; r < nu; ++r) w[ny + r] = pr->UW[j * nu + r];
        for (int a = 0; a < ne; ++a) for (int b = 0; b < ne; ++b) {
            double acc = 0;
            for (int r = 0; r < ny + nu; ++r) acc += f->ssC[r * ne + a] * w[r] * f->ssC[r * ne + b];
            PM(i * ne + a, i * ne + b) = acc;
        }
        if (i < ph) for (int r = 0; r < nu; ++r) { int o = (ph + 1) * ne + i * nu + r; PM(o, o) = pr->DUW[i * nu + r]; }
    }
    free(w);
    for (int i = 0; i <= ph; ++i) {
        for (int r = 0; r < ne; ++r) AM(i * ne + r, i * ne + r) = -1.0;
        if (i > 0) {
            for (int r = 0; r < ne; ++r) {
                for (int k = 0; k < ne; ++k) AM(i * ne + r, (i - 1) * ne + k) += f->ssA[r * ne + k];
                for (int k = 0; k < nu; ++k) AM(i * ne + r, (ph + 1) * ne + (i - 1) * nu + k) = f->ssB[r * nu + k];
            }
        }
    }
    int r0 = meq, r1 = r0 + (ph + 1) * ne, r2 = r1 + (ph + 1) * ny, r3 = r2 + ph * nu;
    for (int k = 0; k < (ph + 1) * ne; ++k) AM(r0 + k, k) = 1.0;
    for (int i = 0; i <= ph; ++i) for (int r = 0; r < ny; ++r) for (int k = 0; k < ne; ++k) AM(r1 + i * ny + r, i * ne + k) = f->ssC[r * ne + k];
    for (int k = 0; k < ph * nu; ++k) AM(r2 + k, (ph + 1) * ne + k) = 1.0;
    for (int i = 0; i <= ph; ++i) for (int k = 0; k < ne; ++k) AM(r3 + i, i * ne + k) = k < nx ? pr->SX[k] : pr->SU[k - nx];
    const double inf = INFINITY;
    for (int i = 0; i <= ph; ++i) {
        int j = jcol(i), col = i < ph ? i : ph - 1;
        for (int r = 0; r < nx; ++r) { f->lineq[i * ne + r] = pr->XMin[j * nx + r]; f->uineq[i * ne + r] = pr->XMax[j * nx + r]; }
        for (int r = 0; r < nu; ++r) { f->lineq[i * ne + nx + r] = pr->UMin[col * nu + r]; f->uineq[i * ne + nx + r] = pr->UMax[col * nu + r]; }
        for (int r = 0; r < ny; ++r) { f->lineq[(ph + 1) * ne + i * ny + r] = pr->YMin[j * ny + r]; f->uineq[(ph + 1) * ne + i * ny + r] = pr->YMax[j * ny + r]; }
        f->lineq[(ph + 1) * ne + (ph + 1) * ny + ph * nu + i] = pr->SMin[j];
        f->uineq[(ph + 1) * ne + (ph + 1) * ny + ph * nu + i] = pr->SMax[j];
    }
    for (int i = 0; i < ph; ++i) for (int r = 0; r < nu; ++r) {
        int o = (ph + 1) * ne + (ph + 1) * ny + i * nu + r;
        int frozen = i > ch;
        f->lineq[o] = frozen ? 0.0 : -inf; f->uineq[o] = frozen ? 0.0 : inf;
    }
    return f;
}
static void form_free(form_t* f) {
    if (!f) return;
    free(f->ssA); free(f->ssB); free(f->ssC); free(f->ssBv); free(f->ssDv); free(f->P); free(f->Amat); free(f->lineq); free(f->uineq); free(f);
}

/* ProblemBuilder::get */
static void form_qlu(const form_t* f, const prob_t* pr, const double* yRef, const double* x0, const double* u0, double* q, double* l, double* u) {
    int nx = f->d.nx, nu = f->d.nu, ndu = f->d.ndu, ny = f->d.ny, ph = f->d.ph, ne = f->ne, n = f->n, m = f->m, meq = f->meq;
    memset(q, 0, sizeof(double) * n);
    double* tmp = (double*)malloc(sizeof(double) * (ny + nu));
    for (int g = 0; g < m; ++g) { l[g] = 0; u[g] = 0; }
    for (int g = meq; g < m; ++g) { l[g] = f->lineq[g - meq]; u[g] = f->uineq[g - meq]; }
    for (int i = 0; i <= ph; ++i) {
        int j = jcol(i);
        const double* d = pr->uMeas + (size_t)j * ndu;
        for (int r = 0; r < ny; ++r) {
            double acc = -yRef[j * ny + r];
            for (int k = 0; k < ndu; ++k) acc += f->ssDv[r * ndu + k] * d[k];
            tmp[r] = pr->OW[j * ny + r] * acc;
        }
        for (int r = 0; r < nu; ++r) tmp[ny + r] = pr->UW[j * nu + r] * (-pr->uRef[j * nu + r]);
        for (int a = 0; a < ne; ++a) {
            double acc = 0;
            for (int r = 0; r < ny + nu; ++r) acc += f->ssC[r * ne + a] * tmp[r];
            q[i * ne + a] = acc;
        }
        if (i < ph) for (int r = 0; r < nu; ++r) q[(ph + 1) * ne + i * nu + r] = -(pr->DUW[i * nu + r] * pr->duRef[j * nu + r]);
        if (i > 0) for (int r = 0; r < ne; ++r) {
            double acc = 0;
            for (int k = 0; k < ndu; ++k) acc -= f->ssBv[r * ndu + k] * d[k];
            l[i * ne + r] = acc; u[i * ne + r] = acc;
        }
        for (int r = 0; r < ny; ++r) {
            double off = 0;
            for (int k = 0; k < ndu; ++k) off -= f->ssDv[r * ndu + k] * d[k];
            int g = meq + (ph + 1) * ne + i * ny + r;
            l[g] += off; u[g] += off;
        }
    }
    for (int r = 0; r < nx; ++r) { l[r] = -x0[r]; u[r] = -x0[r]; }
    for (int r = 0; r < nu; ++r) { l[nx + r] = -u0[r]; u[nx + r] = -u0[r]; }
    free(tmp);
}

/* ---- sparse LDL' (up-looking, elimination tree) of a symmetric quasi-definite matrix given by its upper triangle -- */
typedef struct {
    int n; int *Lp, *Li, *etree, *Lnz, *iw; double *Lx, *D, *Dinv, *fw; unsigned char* bw;
} ldl_t;

static int ldl_symbolic(ldl_t* F, const csc_t* K) {
    int n = K->n; F->n = n;
    F->etree = (int*)malloc(sizeof(int) * n); F->Lnz = (int*)malloc(sizeof(int) * n); F->iw = (int*)malloc(sizeof(int) * 3 * n);
    int* work = F->iw;
    for (int i = 0; i < n; ++i) { work[i] = 0; F->Lnz[i] = 0; F->etree[i] = -1; }
    for (int j = 0; j < n; ++j) {
        work[j] = j;
        for (int p = K->p[j]; p < K->p[j + 1]; ++p) {
            int i = K->i[p];
            if (i > j) return -1;
            while (work[i] != j) {
                if (F->etree[i] == -1) F->etree[i] = j;
                F->Lnz[i]++; work[i] = j; i = F->etree[i];
            }
        }
    }
    int tot = 0;
    for (int i = 0; i < n; ++i) tot += F->Lnz[i];
    F->Lp = (int*)malloc(sizeof(int) * (n + 1)); F->Li = (int*)malloc(sizeof(int) * (tot > 0 ? tot : 1));
    F->Lx = (double*)malloc(sizeof(double) * (tot > 0 ? tot : 1));
    F->D = (double*)malloc(sizeof(double) * n); F->Dinv = (double*)malloc(sizeof(double) * n);
    F->fw = (double*)malloc(sizeof(double) * n); F->bw = (unsigned char*)malloc(n);
    return tot;
}
static int ldl_numeric(ldl_t* F, const csc_t* K) {
    int n = F->n;
    int* yIdx = F->iw; int* elim = F->iw + n; int* next = F->iw + 2 * n;
    double* yv = F->fw; unsigned char* mark = F->bw;
    F->Lp[0] = 0;
    for (int i = 0; i < n; ++i) { F->Lp[i + 1] = F->Lp[i] + F->Lnz[i]; mark[i] = 0; yv[i] = 0; F->D[i] = 0; next[i] = F->Lp[i]; }
    for (int k = 0; k < n; ++k) {
        int nnzY = 0;
        for (int p = K->p[k]; p < K->p[k + 1]; ++p) {
            int b = K->i[p];
            if (b == k) { F->D[k] = K->x[p]; continue; }
            yv[b] = K->x[p];
            int nx = b;
            if (!mark[nx]) {
                mark[nx] = 1; elim[0] = nx; int nE = 1;
                nx = F->etree[b];
                while (nx != -1 && nx < k) {
                    if (mark[nx]) break;
                    mark[nx] = 1; elim[nE++] = nx; nx = F->etree[nx];
                }
                while (nE) yIdx[nnzY++] = elim[--nE];
            }
        }
        for (int i = nnzY - 1; i >= 0; --i) {
            int c = yIdx[i];
            int t = next[c];
            double yc = yv[c];
            for (int j = F->Lp[c]; j < t; ++j) yv[F->Li[j]] -= F->Lx[j] * yc;
            F->Li[t] = k; F->Lx[t] = yc * F->Dinv[c];
            F->D[k] -= yc * F->Lx[t];
            next[c]++;
            yv[c] = 0; mark[c] = 0;
        }
        if (F->D[k] == 0.0) return -1;
        F->Dinv[k] = 1.0 / F->D[k];
    }
    return 0;
}
static void ldl_solve(const ldl_t* F, double* x) {
    int n = F->n;
    for (int i = 0; i < n; ++i) { double xi = x[i]; for (int j = F->Lp[i]; j < F->Lp[i + 1]; ++j) x[F->Li[j]] -= F->Lx[j] * xi; }
    for (int i = 0; i < n; ++i) x[i] *= F->Dinv[i];
    for (int i = n - 1; i >= 0; --i) { double xi = x[i]; for (int j = F->Lp[i]; j < F->Lp[i + 1]; ++j) xi -= F->Lx[j] * x[F->Li[j]]; x[i] = xi; }
}
static void ldl_free(ldl_t* F) { free(F->Lp); free(F->Li); free(F->etree); free(F->Lnz); free(F->iw); free(F->Lx); free(F->D); free(F->Dinv); free(F->fw); free(F->bw); }

/* ---- KKT assembly: upper triangle of  perm' [P+sigma I, A'; A, -diag(1/rho)] perm  ------------------------------ */
typedef struct { int n, m, N; int* perm; int* iperm; csc_t* K; int* rhoIdx; ldl_t F; int factored; } kkt_t;

/* stage-interleaved ordering: eq(0) rows, then for each stage: its local rows, its variables, the dynamics rows that
 * couple it to the next stage */
static void build_perm(const dims_t* d, int n, int m, int* perm) {
    int nx = d->nx, nu = d->nu, ny = d->ny, ph = d->ph, ne = nx + nu;
    int M0 = (ph + 1) * ne, M1 = 2 * (ph + 1) * ne, M2 = M1 + (ph + 1) * ny, M3 = M2 + ph * nu;
    int k = 0;
    for (int r = 0; r < ne; ++r) perm[k++] = n + r;
    for (int i = 0; i <= ph; ++i) {
        for (int r = 0; r < ne; ++r) perm[k++] = n + M0 + i * ne + r;
        for (int r = 0; r < ny; ++r) perm[k++] = n + M1 + i * ny + r;
        perm[k++] = n + M3 + i;
        if (i < ph) for (int r = 0; r < nu; ++r) perm[k++] = n + M2 + i * nu + r;
        for (int r = 0; r < ne; ++r) perm[k++] = i * ne + r;
        if (i < ph) for (int r = 0; r < nu; ++r) perm[k++] = (ph + 1) * ne + i * nu + r;
        if (i < ph) for (int r = 0; r < ne; ++r) perm[k++] = n + (i + 1) * ne + r;
    }
    (void)m;
}

typedef struct { int r, c; double v; int rhoidx; } trip_t;
static int trip_cmp(const void* a, const void* b) {
    const trip_t* x = (const trip_t*)a; const trip_t* y = (const trip_t*)b;
    if (x->c != y->c) return x->c - y->c;
    return x->r - y->r;
}
/* Pu: upper triangle CSC of P (n x n); Ac: CSC (mred x n); diag1 added to the P diagonal, diag2[j] = value of the (2,2)
 * diagonal (negative) */
static int kkt_build(kkt_t* S, const csc_t* Pu, const csc_t* Ac, double diag1, const double* diag2, const int* perm_in) {
    int n = Pu->n, m = Ac->m, N = n + m;
    S->n = n; S->m = m; S->N = N;
    S->perm = (int*)malloc(sizeof(int) * N); S->iperm = (int*)malloc(sizeof(int) * N);
    for (int k = 0; k < N; ++k) { S->perm[k] = perm_in ? perm_in[k] : k; }
    for (int k = 0; k < N; ++k) S->iperm[S->perm[k]] = k;
    int cap = Pu->nnz + n + Ac->nnz + m;
    trip_t* T = (trip_t*)malloc(sizeof(trip_t) * cap);
    int t = 0;
    unsigned char* hasdiag = (unsigned char*)calloc(n, 1);
    for (int j = 0; j < n; ++j) for (int p = Pu->p[j]; p < Pu->p[j + 1]; ++p) {
        int i = Pu->i[p]; double v = Pu->x[p];
        if (i == j) { v += diag1; hasdiag[j] = 1; }
        int a = S->iperm[i], b = S->iperm[j];
        T[t].r = a < b ? a : b; T[t].c = a < b ? b : a; T[t].v = v; T[t].rhoidx = -1; ++t;
    }
    for (int j = 0; j < n; ++j) if (!hasdiag[j]) { int a = S->iperm[j]; T[t].r = a; T[t].c = a; T[t].v = diag1; T[t].rhoidx = -1; ++t; }
    free(hasdiag);
    for (int j = 0; j < n; ++j) for (int p = Ac->p[j]; p < Ac->p[j + 1]; ++p) {
        int a = S->iperm[n + Ac->i[p]], b = S->iperm[j];
        T[t].r = a < b ? a : b; T[t].c = a < b ? b : a; T[t].v = Ac->x[p]; T[t].rhoidx = -1; ++t;
    }
    for (int g = 0; g < m; ++g) { int a = S->iperm[n + g]; T[t].r = a; T[t].c = a; T[t].v = diag2[g]; T[t].rhoidx = g; ++t; }
    qsort(T, t, sizeof(trip_t), trip_cmp);
    S->K = csc_alloc(N, N, t);
    S->rhoIdx = (int*)malloc(sizeof(int) * (m > 0 ? m : 1));
    int col = 0; S->K->p[0] = 0;
    for (int k = 0; k < t; ++k) {
        while (col < T[k].c) S->K->p[++col] = k;
        S->K->i[k] = T[k].r; S->K->x[k] = T[k].v;
        if (T[k].rhoidx >= 0) S->rhoIdx[T[k].rhoidx] = k;
    }
    while (col < N) S->K->p[++col] = t;
    free(T);
    if (ldl_symbolic(&S->F, S->K) < 0) return -1;
    S->factored = 1;
    return ldl_numeric(&S->F, S->K);
}
static int kkt_update_diag2(kkt_t* S, const double* diag2) {
    for (int g = 0; g < S->m; ++g) S->K->x[S->rhoIdx[g]] = diag2[g];
    return ldl_numeric(&S->F, S->K);
}
static void kkt_solve(const kkt_t* S, double* b, double* work) {   /* b has N entries in natural order */
    int N = S->N;
    for (int k = 0; k < N; ++k) work[k] = b[S->perm[k]];
    ldl_solve(&S->F, work);
    for (int k = 0; k < N; ++k) b[S->perm[k]] = work[k];
}
static void kkt_free(kkt_t* S) { free(S->perm); free(S->iperm); csc_free(S->K); free(S->rhoIdx); if (S->factored) ldl_free(&S->F); }

/* ---- sparse helpers ------------------------------------------------------------------------------------------------ */
static void spmv(const csc_t* A, const double* x, double* y, int accumulate) {      /* y (+)= A x */
    if (!accumulate) for (int i = 0; i < A->m; ++i) y[i] = 0;
    for (int j = 0; j < A->n; ++j) { double xj = x[j]; for (int p = A->p[j]; p < A->p[j + 1]; ++p) y[A->i[p]] += A->x[p] * xj; }
}
static void spmtv(const csc_t* A, const double* x, double* y, int accumulate) {     /* y (+)= A' x */
    for (int j = 0; j < A->n; ++j) { double acc = accumulate ? y[j] : 0; for (int p = A->p[j]; p < A->p[j + 1]; ++p) acc += A->x[p] * x[A->i[p]]; y[j] = acc; }
}
static void sym_spmv_upper(const csc_t* Pu, const double* x, double* y) {             /* y = P x, P given by its upper triangle */
    for (int i = 0; i < Pu->n; ++i) y[i] = 0;
    for (int j = 0; j < Pu->n; ++j) for (int p = Pu->p[j]; p < Pu->p[j + 1]; ++p) {
        int i = Pu->i[p]; double v = Pu->x[p];
        y[i] += v * x[j];
        if (i != j) y[j] += v * x[i];
    }
}
static double norm_inf(const double* v, int n) { double mx = 0; for (int i = 0; i < n; ++i) { double a = fabs(v[i]); if (a > mx) mx = a; } return mx; }
static double limit_scaling(double v) { v = v < MIN_SCALING ? 1.0 : v; return v > MAX_SCALING ? MAX_SCALING : v; }

/* ---- OSQP v0.6.3 restated on CSC data -------------------------------------------------------------------------------- */
typedef struct {
    int n, m; csc_t *P, *A; double *q, *l, *u;          /* scaled in place */
    double *D, *E, *Dinv, *Einv, c, cinv;
    double *rho_vec, *rho_inv; int* ctype; double rho;
    double *x, *z, *y, *xprev, *zprev, *dx, *dy, *Ax, *Px, *Aty, *xz, *work, *pv, *dv;
    kkt_t K; const int* perm;
} osqp_t;

static void scale_data(osqp_t* w, int iters) {
    int n = w->n, m = w->m;
    double* Dt = (double*)malloc(sizeof(double) * n); double* Et = (double*)malloc(sizeof(double) * m);
    for (int i = 0; i < n; ++i) w->D[i] = 1;
    for (int i = 0; i < m; ++i) w->E[i] = 1;
    w->c = 1.0;
    for (int it = 0; it < iters; ++it) {
        for (int j = 0; j < n; ++j) Dt[j] = 0;
        for (int i = 0; i < m; ++i) Et[i] = 0;
        for (int j = 0; j < n; ++j) for (int p = w->P->p[j]; p < w->P->p[j + 1]; ++p) {
            double a = fabs(w->P->x[p]); int i = w->P->i[p];
            if (a > Dt[j]) Dt[j] = a;
            if (i != j && a > Dt[i]) Dt[i] = a;
        }
        for (int j = 0; j < n; ++j) for (int p = w->A->p[j]; p < w->A->p[j + 1]; ++p) {
            double a = fabs(w->A->x[p]); int i = w->A->i[p];
            if (a > Dt[j]) Dt[j] = a;
            if (a > Et[i]) Et[i] = a;
        }
        for (int j = 0; j < n; ++j) Dt[j] = 1.0 / sqrt(limit_scaling(Dt[j]));
        for (int i = 0; i < m; ++i) Et[i] = 1.0 / sqrt(limit_scaling(Et[i]));
        for (int j = 0; j < n; ++j) for (int p = w->P->p[j]; p < w->P->p[j + 1]; ++p) w->P->x[p] = Dt[w->P->i[p]] * w->P->x[p] * Dt[j];
        for (int j = 0; j < n; ++j) for (int p = w->A->p[j]; p < w->A->p[j + 1]; ++p) w->A->x[p] = Et[w->A->i[p]] * w->A->x[p] * Dt[j];
        for (int j = 0; j < n; ++j) { w->q[j] *= Dt[j]; w->D[j] *= Dt[j]; }
        for (int i = 0; i < m; ++i) w->E[i] *= Et[i];
        for (int j = 0; j < n; ++j) Dt[j] = 0;
        for (int j = 0; j < n; ++j) for (int p = w->P->p[j]; p < w->P->p[j + 1]; ++p) {
            double a = fabs(w->P->x[p]); int i = w->P->i[p];
            if (a > Dt[j]) Dt[j] = a;
            if (i != j && a > Dt[i]) Dt[i] = a;
        }
        double ct = 0; for (int j = 0; j < n; ++j) ct += Dt[j]; ct /= n;
        double nq = limit_scaling(norm_inf(w->q, n));
        ct = ct > nq ? ct : nq;
        ct = 1.0 / limit_scaling(ct);
        for (int p = 0; p < w->P->nnz; ++p) w->P->x[p] *= ct;
        for (int j = 0; j < n; ++j) w->q[j] *= ct;
        w->c *= ct;
    }
    w->cinv = 1.0 / w->c;
    for (int j = 0; j < n; ++j) w->Dinv[j] = 1.0 / w->D[j];
    for (int i = 0; i < m; ++i) { w->Einv[i] = 1.0 / w->E[i]; w->l[i] *= w->E[i]; w->u[i] *= w->E[i]; }
    free(Dt); free(Et);
}
static void update_rho_vec(osqp_t* w) {
    for (int i = 0; i < w->m; ++i) {
        w->rho_vec[i] = w->ctype[i] == -1 ? RHO_MIN : (w->ctype[i] == 1 ? RHO_EQ_OVER_RHO_INEQ * w->rho : w->rho);
        w->rho_inv[i] = 1.0 / w->rho_vec[i];
    }
}
typedef struct { double pri, dua; } res_t;
static res_t update_info(osqp_t* w, const double* x, const double* z, const double* y) {
    int n = w->n, m = w->m; res_t r;
    spmv(w->A, x, w->Ax, 0);
    double mx = 0;
    for (int i = 0; i < m; ++i) { w->pv[i] = w->Ax[i] - z[i]; double a = fabs(w->Einv[i] * w->pv[i]); if (a > mx) mx = a; }
    r.pri = mx;
    sym_spmv_upper(w->P, x, w->Px);
    spmtv(w->A, y, w->Aty, 0);
    mx = 0;
    for (int j = 0; j < n; ++j) { w->dv[j] = w->q[j] + w->Px[j] + w->Aty[j]; double a = fabs(w->Dinv[j] * w->dv[j]); if (a > mx) mx = a; }
    r.dua = w->cinv * mx;
    return r;
}
static int primal_infeasible(osqp_t* w, double eps) {
    int m = w->m, n = w->n; double nd = 0, lhs = 0;
    for (int i = 0; i < m; ++i) {
        double dy = w->dy[i];
        if (w->u[i] > OSQP_INFTY * MIN_SCALING) { if (w->l[i] < -OSQP_INFTY * MIN_SCALING) dy = 0; else dy = dy < 0 ? dy : 0; }
        else if (w->l[i] < -OSQP_INFTY * MIN_SCALING) dy = dy > 0 ? dy : 0;
        w->dy[i] = dy;
        double a = fabs(w->E[i] * dy); if (a > nd) nd = a;
    }
    if (nd > eps) {
        for (int i = 0; i < m; ++i) lhs += w->u[i] * (w->dy[i] > 0 ? w->dy[i] : 0.0) + w->l[i] * (w->dy[i] < 0 ? w->dy[i] : 0.0);  /* inf*0 = NaN on purpose */
        if (lhs < -eps * nd) {
            spmtv(w->A, w->dy, w->work, 0);
            double mx = 0; for (int j = 0; j < n; ++j) { double a = fabs(w->Dinv[j] * w->work[j]); if (a > mx) mx = a; }
            return mx < eps * nd;
        }
    }
    return 0;
}
static int dual_infeasible(osqp_t* w, double eps) {
    int m = w->m, n = w->n; double nd = 0, qd = 0;
    for (int j = 0; j < n; ++j) { double a = fabs(w->D[j] * w->dx[j]); if (a > nd) nd = a; qd += w->q[j] * w->dx[j]; }
    if (nd > eps && qd < -w->c * eps * nd) {
        sym_spmv_upper(w->P, w->dx, w->work);
        double mx = 0; for (int j = 0; j < n; ++j) { double a = fabs(w->Dinv[j] * w->work[j]); if (a > mx) mx = a; }
        if (mx < w->c * eps * nd) {
            spmv(w->A, w->dx, w->xz, 0);
            for (int i = 0; i < m; ++i) {
                double adx = w->Einv[i] * w->xz[i];
                if ((w->u[i] < OSQP_INFTY * MIN_SCALING && adx > eps * nd) || (w->l[i] > -OSQP_INFTY * MIN_SCALING && adx < -eps * nd)) return 0;
            }
            return 1;
        }
    }
    return 0;
}
static int check_termination(osqp_t* w, const params_t* p, res_t r, int approx, int* status, double* obj) {
    int n = w->n, m = w->m;
    double ea = p->eps_abs, er = p->eps_rel, epi = p->eps_prim_inf, edi = p->eps_dual_inf;
    if (r.pri > OSQP_INFTY || r.dua > OSQP_INFTY) { *status = ST_NON_CVX; *obj = NAN; return 1; }
    if (approx) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
    double nz = 0, nAx = 0;
    for (int i = 0; i < m; ++i) { double a = fabs(w->Einv[i] * w->z[i]); if (a > nz) nz = a; a = fabs(w->Einv[i] * w->Ax[i]); if (a > nAx) nAx = a; }
    double nq = 0, nAty = 0, nPx = 0;
    for (int j = 0; j < n; ++j) { double a = fabs(w->Dinv[j] * w->q[j]); if (a > nq) nq = a; a = fabs(w->Dinv[j] * w->Aty[j]); if (a > nAty) nAty = a; a = fabs(w->Dinv[j] * w->Px[j]); if (a > nPx) nPx = a; }
    int pok = 0, dok = 0, pinf = 0, dinf = 0;
    double eps_prim = ea + er * (nz > nAx ? nz : nAx);
    if (r.pri < eps_prim) pok = 1; else pinf = primal_infeasible(w, epi);
    double mxd = nq > nAty ? nq : nAty; mxd = mxd > nPx ? mxd : nPx;
    double eps_dual = ea + er * w->cinv * mxd;
    if (r.dua < eps_dual) dok = 1; else dinf = dual_infeasible(w, edi);
    if (pok && dok) { *status = approx ? ST_SOLVED_INACC : ST_SOLVED; return 1; }
    if (pinf) { *status = approx ? ST_PRIM_INF_INACC : ST_PRIM_INF; *obj = OSQP_INFTY; return 1; }
    if (dinf) { *status = approx ? ST_DUAL_INF_INACC : ST_DUAL_INF; *obj = -OSQP_INFTY; return 1; }
    return 0;
}
static double obj_val(osqp_t* w, const double* x) {
    sym_spmv_upper(w->P, x, w->work);
    double s = 0; for (int j = 0; j < w->n; ++j) s += 0.5 * x[j] * w->work[j] + w->q[j] * x[j];
    return s * w->cinv;
}

typedef struct { int key, node; } kn_t;
static int kn_cmp(const void* a, const void* c2) { return ((const kn_t*)a)->key - ((const kn_t*)c2)->key; }

static void polish(osqp_t* w, const params_t* p, res_t info, double* obj, int* status_polish) {
    int n = w->n, m = w->m;
    int* rows = (int*)malloc(sizeof(int) * 2 * m); double* b = (double*)malloc(sizeof(double) * 2 * m);
    int nlow = 0, mred = 0;
    for (int i = 0; i < m; ++i) if (w->z[i] - w->l[i] < -w->y[i]) { rows[mred] = i; b[mred] = w->l[i]; ++mred; }
    nlow = mred;
    for (int i = 0; i < m; ++i) if (w->u[i] - w->z[i] < w->y[i]) { rows[mred] = i; b[mred] = w->u[i]; ++mred; }
    (void)nlow;
    /* Ared (mred x n) in CSC */
    int* rowmap = (int*)malloc(sizeof(int) * m * 2); int* cnt = (int*)calloc(m, sizeof(int));
    for (int k = 0; k < mred; ++k) { rowmap[rows[k] * 2 + cnt[rows[k]]] = k; cnt[rows[k]]++; }
    int nnz = 0;
    for (int pp = 0; pp < w->A->nnz; ++pp) nnz += cnt[w->A->i[pp]];
    csc_t* Ar = csc_alloc(mred, n, nnz);
    int t = 0;
    for (int j = 0; j < n; ++j) {
        Ar->p[j] = t;
        /* keep row indices sorted inside the column: two passes (lower rows first is not required by our kernels) */
        for (int pp = w->A->p[j]; pp < w->A->p[j + 1]; ++pp) { int i = w->A->i[pp]; for (int c = 0; c < cnt[i]; ++c) { Ar->i[t] = rowmap[i * 2 + c]; Ar->x[t] = w->A->x[pp]; ++t; } }
    }
    Ar->p[n] = t;
    double* d2 = (double*)malloc(sizeof(double) * (mred > 0 ? mred : 1));
    for (int k = 0; k < mred; ++k) d2[k] = -p->delta;
    /* ordering for the reduced KKT: keep the stage-interleaved order restricted to the kept rows */
    int N = n + mred; int* perm = (int*)malloc(sizeof(int) * N);
    {
        int k = 0;
        int* pos = (int*)malloc(sizeof(int) * (n + m));       /* position of every original node in w->perm */
        for (int q = 0; q < n + m; ++q) pos[w->perm[q]] = q;
        /* emit nodes in the original order, duplicating nothing: rows kept (possibly twice) */
        kn_t* arr = (kn_t*)malloc(sizeof(kn_t) * N);
        for (int j = 0; j < n; ++j) { arr[k].key = pos[j] * 2; arr[k].node = j; ++k; }
        for (int r = 0; r < mred; ++r) { arr[k].key = pos[n + rows[r]] * 2 + (r >= nlow ? 1 : 0); arr[k].node = n + r; ++k; }
        /* insertion-free sort: counting by key range would do; N is small, use qsort */
        qsort(arr, N, sizeof(kn_t), kn_cmp);
        for (int q = 0; q < N; ++q) perm[q] = arr[q].node;
        free(arr); free(pos);
    }
    kkt_t K; memset(&K, 0, sizeof(K));
    int rc = kkt_build(&K, w->P, Ar, p->delta, d2, perm);
    if (rc == 0) {
        double* rhs = (double*)malloc(sizeof(double) * N); double* sol = (double*)malloc(sizeof(double) * N);
        double* r2 = (double*)malloc(sizeof(double) * N); double* wk = (double*)malloc(sizeof(double) * N);
        for (int j = 0; j < n; ++j) rhs[j] = -w->q[j];
        for (int k = 0; k < mred; ++k) rhs[n + k] = b[k];
        memcpy(sol, rhs, sizeof(double) * N);
        kkt_solve(&K, sol, wk);
        for (int it = 0; it < p->polish_refine_iter; ++it) {
            memcpy(r2, rhs, sizeof(double) * N);
            sym_spmv_upper(w->P, sol, wk);
            for (int j = 0; j < n; ++j) r2[j] -= wk[j];
            spmtv(Ar, sol + n, wk, 0);
            for (int j = 0; j < n; ++j) r2[j] -= wk[j];
            spmv(Ar, sol, wk, 0);
            for (int k = 0; k < mred; ++k) r2[n + k] -= wk[k];
            kkt_solve(&K, r2, wk);
            for (int k = 0; k < N; ++k) sol[k] += r2[k];
        }
        double* pz = (double*)malloc(sizeof(double) * m); double* py = (double*)calloc(m, sizeof(double));
        spmv(w->A, sol, pz, 0);
        for (int k = mred - 1; k >= 0; --k) py[rows[k]] = sol[n + k];   /* a row kept twice keeps the LOWER multiplier (get_ypol_from_yred) */
        for (int i = 0; i < m; ++i) { double tt = pz[i] + py[i]; double zz = tt < w->l[i] ? w->l[i] : (tt > w->u[i] ? w->u[i] : tt); pz[i] = zz; py[i] = tt - zz; }
        /* update_info overwrites Ax/Px/Aty: fine, the ADMM info is already consumed */
        res_t pr = update_info(w, sol, pz, py);
        int ok = (pr.pri < info.pri && pr.dua < info.dua) || (pr.pri < info.pri && info.dua < 1e-10) || (pr.dua < info.dua && info.pri < 1e-10);
        if (ok) {
            *obj = obj_val(w, sol); *status_polish = 1;
            memcpy(w->x, sol, sizeof(double) * n); memcpy(w->z, pz, sizeof(double) * m); memcpy(w->y, py, sizeof(double) * m);
        } else *status_polish = -1;
        free(pz); free(py); free(rhs); free(sol); free(r2); free(wk);
    } else *status_polish = -1;
    kkt_free(&K);
    free(perm); free(d2); csc_free(Ar); free(rowmap); free(cnt); free(rows); free(b);
}

/* one LOptimizer::run.  x_out[n], y_out[m] (unscaled solution), warm (x,y) may be NULL */
static void osqp_run(const dims_t* dd, csc_t* Pu, csc_t* Ac, double* q, double* l, double* u, const params_t* p,
                     const double* warm_x, const double* warm_y, double* x_out, double* y_out, result_t* res) {
    osqp_t W; memset(&W, 0, sizeof(W));
    osqp_t* w = &W;
    int n = Pu->n, m = Ac->m, N = n + m;
    w->n = n; w->m = m; w->P = Pu; w->A = Ac; w->q = q; w->l = l; w->u = u;
#define AL(ptr, cnt) w->ptr = (double*)calloc((cnt) > 0 ? (cnt) : 1, sizeof(double))
    AL(D, n); AL(E, m); AL(Dinv, n); AL(Einv, m); AL(rho_vec, m); AL(rho_inv, m); AL(x, n); AL(z, m); AL(y, m); AL(xprev, n); AL(zprev, m);
    AL(dx, n); AL(dy, m); AL(Ax, m); AL(Px, n); AL(Aty, n); AL(xz, N); AL(work, N); AL(pv, m); AL(dv, n);
    w->ctype = (int*)malloc(sizeof(int) * (m > 0 ? m : 1));
    int* perm = (int*)malloc(sizeof(int) * N);
    build_perm(dd, n, m, perm);
    w->perm = perm;
    if (p->scaling) scale_data(w, p->scaling); else { for (int j = 0; j < n; ++j) w->D[j] = w->Dinv[j] = 1; for (int i = 0; i < m; ++i) w->E[i] = w->Einv[i] = 1; w->c = w->cinv = 1; }
    w->rho = p->rho < RHO_MIN ? RHO_MIN : (p->rho > RHO_MAX ? RHO_MAX : p->rho);
    for (int i = 0; i < m; ++i) {
        if (l[i] < -OSQP_INFTY * MIN_SCALING && u[i] > OSQP_INFTY * MIN_SCALING) w->ctype[i] = -1;
        else if (u[i] - l[i] < RHO_TOL) w->ctype[i] = 1; else w->ctype[i] = 0;
    }
    update_rho_vec(w);
    double* d2 = (double*)malloc(sizeof(double) * (m > 0 ? m : 1));
    for (int i = 0; i < m; ++i) d2[i] = -w->rho_inv[i];
    memset(&w->K, 0, sizeof(w->K));
    kkt_build(&w->K, Pu, Ac, p->sigma, d2, perm);
    int status = ST_UNSOLVED; double obj = 0; res->rho_updates = 0; res->status_polish = 0;
    if (p->warm_start && warm_x && warm_y) {
        for (int j = 0; j < n; ++j) w->x[j] = w->Dinv[j] * warm_x[j];
        for (int i = 0; i < m; ++i) w->y[i] = w->c * w->Einv[i] * warm_y[i];
        spmv(Ac, w->x, w->z, 0);
    }
    res_t info = {0, 0};
    int can_check = 0, it = 0, done = 0;
    for (it = 1; it <= p->max_iter; ++it) {
        double* tx = w->x; w->x = w->xprev; w->xprev = tx;
        double* tz = w->z; w->z = w->zprev; w->zprev = tz;
        for (int j = 0; j < n; ++j) w->xz[j] = p->sigma * w->xprev[j] - q[j];
        for (int i = 0; i < m; ++i) w->xz[n + i] = w->zprev[i] - w->rho_inv[i] * w->y[i];
        double* rhs_z = w->pv;      /* keep the z part of the rhs: z~ = rhs_z + rho_inv * nu */
        for (int i = 0; i < m; ++i) rhs_z[i] = w->xz[n + i];
        kkt_solve(&w->K, w->xz, w->work);
        for (int j = 0; j < n; ++j) { double xn = p->alpha * w->xz[j] + (1.0 - p->alpha) * w->xprev[j]; w->x[j] = xn; w->dx[j] = xn - w->xprev[j]; }
        for (int i = 0; i < m; ++i) {
            double zt = rhs_z[i] + w->rho_inv[i] * w->xz[n + i];
            double zr = p->alpha * zt + (1.0 - p->alpha) * w->zprev[i];
            double zz = zr + w->rho_inv[i] * w->y[i];
            zz = zz < l[i] ? l[i] : (zz > u[i] ? u[i] : zz);
            w->z[i] = zz;
            w->dy[i] = w->rho_vec[i] * (zr - zz);
            w->y[i] += w->dy[i];
        }
        can_check = p->check_termination && (it % p->check_termination == 0);
        int can_adapt = p->adaptive_rho && p->adaptive_rho_interval && (it % p->adaptive_rho_interval == 0);
        if (can_check || can_adapt) {
            info = update_info(w, w->x, w->z, w->y);
            if (can_check && check_termination(w, p, info, 0, &status, &obj)) { done = 1; break; }
        }
        if (can_adapt) {
            double pr = norm_inf(w->pv, m), dr = norm_inf(w->dv, n);
            double a = norm_inf(w->z, m), b2 = norm_inf(w->Ax, m);
            pr /= ((a > b2 ? a : b2) + 1e-10);
            double c1 = norm_inf(q, n), c2 = norm_inf(w->Aty, n), c3 = norm_inf(w->Px, n);
            double mx = c1 > c2 ? c1 : c2; mx = mx > c3 ? mx : c3;
            dr /= (mx + 1e-10);
            double est = w->rho * sqrt(pr / (dr + 1e-10));
            est = est < RHO_MIN ? RHO_MIN : (est > RHO_MAX ? RHO_MAX : est);
            if (est > w->rho * p->adaptive_rho_tolerance || est < w->rho / p->adaptive_rho_tolerance) {
                w->rho = est; update_rho_vec(w);
                for (int i = 0; i < m; ++i) d2[i] = -w->rho_inv[i];
                kkt_update_diag2(&w->K, d2);
                res->rho_updates++;
            }
        }
    }
    res->iters = done ? it : p->max_iter;
    if (!done && !can_check) { info = update_info(w, w->x, w->z, w->y); check_termination(w, p, info, 0, &status, &obj); }
    int has_solution = !(status == ST_PRIM_INF || status == ST_PRIM_INF_INACC || status == ST_DUAL_INF || status == ST_DUAL_INF_INACC || status == ST_NON_CVX);
    if (has_solution) obj = obj_val(w, w->x);
    if (status == ST_UNSOLVED) { if (!check_termination(w, p, info, 1, &status, &obj)) status = ST_MAX_ITER; }
    if (p->polish && status == ST_SOLVED) polish(w, p, info, &obj, &res->status_polish);
    for (int j = 0; j < n; ++j) x_out[j] = has_solution ? w->D[j] * w->x[j] : NAN;
    for (int i = 0; i < m; ++i) y_out[i] = has_solution ? w->cinv * w->E[i] * w->y[i] : NAN;
    res->solver_status = status; res->cost = obj;
    res->is_feasible = (status == ST_SOLVED || status == ST_SOLVED_INACC || status == ST_MAX_ITER);
    switch (status) {
    case ST_SOLVED: case ST_SOLVED_INACC: case ST_PRIM_INF_INACC: case ST_DUAL_INF_INACC: res->status = 0; break;
    case ST_MAX_ITER: res->status = 1; break;
    case ST_PRIM_INF: case ST_DUAL_INF: res->status = 2; break;
    case ST_NON_CVX: res->status = 3; break;
    default: res->status = 4;
    }
    kkt_free(&w->K);
    free(d2); free(perm); free(w->ctype);
    free(w->D); free(w->E); free(w->Dinv); free(w->Einv); free(w->rho_vec); free(w->rho_inv); free(w->x); free(w->z); free(w->y);
    free(w->xprev); free(w->zprev); free(w->dx); free(w->dy); free(w->Ax); free(w->Px); free(w->Aty); free(w->xz); free(w->work); free(w->pv); free(w->dv);
}

/* ---- public entry points (ctypes) ------------------------------------------------------------------------------------- */
typedef struct {
    const dims_t* dd; const params_t* par; const prob_t* pr; form_t* f; int batch; const double *x0, *u0; int yref_per_instance;
    const double *warm_x, *warm_y; double* cmd; result_t* results; double *sol_x, *sol_y; int next;
} job_t;

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    const dims_t* dd = J->dd; form_t* f = J->f;
    int n = f->n, m = f->m, nx = dd->nx, nu = dd->nu, ne = nx + nu;
    double* q = (double*)malloc(sizeof(double) * n); double* l = (double*)malloc(sizeof(double) * m); double* u = (double*)malloc(sizeof(double) * m);
    double* xo = (double*)malloc(sizeof(double) * n); double* yo = (double*)malloc(sizeof(double) * m);
    for (;;) {
        int b = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
        if (b >= J->batch) break;
        const double* yr = J->yref_per_instance ? J->pr->yRef + (size_t)b * dd->ph * dd->ny : J->pr->yRef;
        form_qlu(f, J->pr, yr, J->x0 + (size_t)b * nx, J->u0 + (size_t)b * nu, q, l, u);   /* per-step get() */
        csc_t* Pu = dense_to_csc(f->P, n, n, 1);                                          /* per-step getSparse() */
        csc_t* Ac = dense_to_csc(f->Amat, m, n, 0);
        osqp_run(dd, Pu, Ac, q, l, u, J->par, J->warm_x ? J->warm_x + (size_t)b * n : NULL, J->warm_y ? J->warm_y + (size_t)b * m : NULL,
                 xo, yo, &J->results[b]);
        int st = dd->ph >= 1 ? 1 : 0;
        for (int k = 0; k < nu; ++k) J->cmd[(size_t)b * nu + k] = xo[st * ne + nx + k];
        if (J->sol_x) memcpy(J->sol_x + (size_t)b * n, xo, sizeof(double) * n);
        if (J->sol_y) memcpy(J->sol_y + (size_t)b * m, yo, sizeof(double) * m);
        csc_free(Pu); csc_free(Ac);
    }
    free(q); free(l); free(u); free(xo); free(yo);
    return NULL;
}

/* Solve `batch` instances that share the problem description `pr` except x0[batch*nx], u0[batch*nu] and, when
 * yref_per_instance != 0, yRef[batch*ph*ny].  Outputs: cmd[batch*nu], results[batch], optional sol_x[batch*n],
 * sol_y[batch*m], optional warm starts warm_x/warm_y (same shapes).  nthreads <= 1: serial (the reference is
 * single-threaded); >1: one independent instance stream per pthread. */
int lmpc_oracle_solve_batch(const dims_t* dd, const params_t* par, const prob_t* pr, int batch, const double* x0, const double* u0,
                            int yref_per_instance, const double* warm_x, const double* warm_y, double* cmd, result_t* results,
                            double* sol_x, double* sol_y, int nthreads) {
    job_t J;
    J.dd = dd; J.par = par; J.pr = pr; J.batch = batch; J.x0 = x0; J.u0 = u0; J.yref_per_instance = yref_per_instance;
    J.warm_x = warm_x; J.warm_y = warm_y; J.cmd = cmd; J.results = results; J.sol_x = sol_x; J.sol_y = sol_y; J.next = 0;
    J.f = form_build(dd, pr);          /* the reference builds dense P,A in its setters, once */
    if (nthreads <= 1) worker(&J);
    else {
        if (nthreads > 256) nthreads = 256;
        pthread_t th[256];
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, &J);
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    }
    form_free(J.f);
    return 0;
}

int lmpc_oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
