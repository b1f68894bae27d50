/* nlmpc_oracle.c -- CPU oracle (TEST INFRASTRUCTURE / CPU BASELINE ONLY), plain C.
 *
 * Restates, for the reference's three example systems, everything libmpc++'s NLMPC hands to its NLP solver for one decision
 * vector z -- exactly what oracle/nlmpc_formulation.py restates in numpy, compiled, so that a CPU baseline of the NLMPC
 * path does not spend its time in Python callbacks:
 *   Mapping::unwrapVector                            include/mpc/NLMPC/Mapping.hpp:174-211 (move blocking :221-257)
 *   Objective::evaluate + computeGradient            include/mpc/NLMPC/Objective.hpp:91-187,198-265   (FORWARD differences)
 *   Constraints::getStateEqConstraints + Jacobian    include/mpc/NLMPC/Constraints.hpp:490-628,844-905 (CENTRAL)
 *   Constraints::evaluateIneq + computeIneqJacobian  include/mpc/NLMPC/Constraints.hpp:211-316,641-721 (CENTRAL)
 * including the reference's step rules (linear column-major index of the step reference value; the objective gradient
 * moves rows ph-1 and ph of U together, the inequality Jacobian does not) and its cost: every perturbation re-evaluates the
 * whole cost / constraint vector, as the reference does.  Systems: 0 vanderpol_ex, 1 / 2 networked_oscillators_ex with
 * N = 4 / 6, 3 ugv_ex; `params` as in include/b200mpc.h.  No scaling, no user equalities (the examples use neither).
 * Pinned by tests/test_nlmpc_c_oracle.py against the numpy oracle (itself pinned to the reference's known-answer tests).
 * Only tests/ and the CPU-baseline tooling may load this. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define DV 1.4901161193847656e-08 /* sqrt(DBL_EPSILON), Objective.hpp:283 */

typedef struct {
    int system, nx, nu, ph, ch, nineq, N, continuous;
    const double* p;
} Sys;

static int sys_init(Sys* s, int system, int ph, int ch, const double* params) {
    s->system = system; s->ph = ph; s->ch = ch; s->p = params; s->N = 0;
    switch (system) {
    case 0: s->nx = 2; s->nu = 1; s->nineq = ph + 1; s->continuous = 1; return 0;
    case 1: s->N = 4; s->nx = 8; s->nu = 4; s->nineq = (ph + 1) * 4; s->continuous = 1; return 0;
    case 2: s->N = 6; s->nx = 12; s->nu = 6; s->nineq = (ph + 1) * 6; s->continuous = 1; return 0;
    case 3: s->nx = 4; s->nu = 2; s->nineq = (ph + 1) * 2; s->continuous = 0; return 0;
    }
    return -1;
}

static void field(const Sys* s, double* out, const double* x, const double* u) {
    const double* p = s->p;
    if (s->system == 0) {                         /* examples/vanderpol_ex.cpp:36-43 */
        out[0] = ((1.0 - (x[1] * x[1])) * x[0]) - x[1] + u[0];
        out[1] = x[0];
    } else if (s->system == 3) {                  /* examples/ugv_ex.cpp: x+ = Ad x + Bd u */
        for (int r = 0; r < 4; ++r) {
            double v = 0, w = 0;
            for (int c = 0; c < 4; ++c) v += p[r * 4 + c] * x[c];
            for (int c = 0; c < 2; ++c) w += p[16 + r * 2 + c] * u[c];
            out[r] = v + w;
        }
    } else {                                      /* examples/networked_oscillators_ex.cpp:17-32 */
        const int N = s->N; const double mu = p[1], k = p[2];
        for (int i = 0; i < N; ++i) {
            out[2 * i] = x[2 * i + 1];
            double v = mu * (1 - x[2 * i] * x[2 * i]) * x[2 * i + 1] - x[2 * i] + u[i];
            for (int j = 0; j < N; ++j) if (i != j) v += k * (x[2 * j] - x[2 * i]);
            out[2 * i + 1] = v;
        }
    }
}

static double cost(const Sys* s, const double* X, const double* U, double e) {
    const int ph = s->ph, nx = s->nx, nu = s->nu;
    if (s->system == 3) {
        const double* p = s->p;
        double c = 0;
        for (int i = 0; i <= ph; ++i) {
            double d0 = X[i * nx + 2] - p[24], d1 = X[i * nx + 3] - p[25];
            c += 1e3 * (d0 * d0 + d1 * d1);
            c += 1e-2 * (U[i * nu] * U[i * nu] + U[i * nu + 1] * U[i * nu + 1]);
        }
        return c + 1e-5 * e * e;
    }
    double sx = 0, su = 0;                        /* column-major summation order (Eigen .array().square().sum()) */
    for (int j = 0; j < nx; ++j) for (int i = 0; i <= ph; ++i) sx += X[i * nx + j] * X[i * nx + j];
    for (int j = 0; j < nu; ++j) for (int i = 0; i <= ph; ++i) su += U[i * nu + j] * U[i * nu + j];
    return sx + su;
}

static void ineq(const Sys* s, const double* X, const double* U, double e, double* out) {
    const int ph = s->ph, nx = s->nx, nu = s->nu;
    (void)e;
    if (s->system == 0) { for (int i = 0; i <= ph; ++i) out[i] = U[i * nu] - 0.5; }
    else if (s->system == 3) {
        const double* p = s->p;
        for (int i = 0; i <= ph; ++i) for (int j = 0; j < 2; ++j) {
            double dx = X[i * nx] - p[26 + 3 * j], dy = X[i * nx + 1] - p[26 + 3 * j + 1];
            out[i * 2 + j] = p[26 + 3 * j + 2] - sqrt(dx * dx + dy * dy);
        }
    } else { for (int i = 0; i <= ph; ++i) for (int j = 0; j < nu; ++j) out[i * nu + j] = U[i * nu + j] - 0.5; }
}

static void unwrap(const Sys* s, const double* z, const double* x0, double* X, double* U, double* e) {
    const int ph = s->ph, ch = s->ch, nx = s->nx, nu = s->nu;
    for (int j = 0; j < nx; ++j) X[j] = x0[j];
    memcpy(X + nx, z, sizeof(double) * ph * nx);
    for (int i = 0; i <= ph; ++i) {
        int st = i < ph ? i : ph - 1, blk = st < ch ? st : ch - 1;
        for (int j = 0; j < nu; ++j) U[i * nu + j] = z[ph * nx + blk * nu + j];
    }
    *e = z[ph * nx + ch * nu];
}

/* the step reference value of column j: Xa.array()(j), the LINEAR column-major index of the (ph+1) x n matrix */
static double step_ref(const double* M, int n, int ph, int j) { return fmax(fabs(M[(j % (ph + 1)) * n + j / (ph + 1)]), 1.0); }

/* any output may be NULL.  Jacobians row-major [rows x nz].  Returns 0, or -1 for an unknown system / out of memory. */
int nlmpc_oracle_eval(int system, int ph, int ch, const double* z, const double* x0, const double* params, double* fval,
                      double* grad, double* ceq, double* Jeq, double* cin, double* Jin) {
    Sys s;
    if (sys_init(&s, system, ph, ch, params)) return -1;
    const int nx = s.nx, nu = s.nu, nz = ph * nx + ch * nu + 1, ni = s.nineq;
    double* X = (double*)malloc(sizeof(double) * ((ph + 1) * (nx + nu) + 2 * ni + 8 * (nx + nu)));
    if (!X) return -1;
    double* U = X + (ph + 1) * nx;
    double* cp = U + (ph + 1) * nu; double* cm = cp + ni;
    double* w = cm + ni;                           /* scratch: xk, xk1, uk, fp, fm ... */
    double e;
    unwrap(&s, z, x0, X, U, &e);
    const double f0 = (fval || grad) ? cost(&s, X, U, e) : 0.0;
    if (fval) *fval = f0;
    if (grad) {
        memset(grad, 0, sizeof(double) * nz);
        for (int i = 0; i < ph; ++i) for (int j = 0; j < nx; ++j) {
            double dx = DV * step_ref(X, nx, ph, j), keep = X[(i + 1) * nx + j];
            X[(i + 1) * nx + j] = keep + dx;
            grad[i * nx + j] = (cost(&s, X, U, e) - f0) / dx;
            X[(i + 1) * nx + j] = keep;
        }
        for (int i = 0; i < ph; ++i) for (int j = 0; j < nu; ++j) {
            double du = DV * step_ref(U, nu, ph, j), k0 = U[i * nu + j], k1 = U[ph * nu + j];
            U[i * nu + j] = k0 + du;
            if (i == ph - 1) U[ph * nu + j] = k1 + du;             /* the duplicated last row moves with stage ph-1 */
            double df = (cost(&s, X, U, e) - f0) / du;
            U[i * nu + j] = k0; U[ph * nu + j] = k1;
            grad[ph * nx + (i < ch ? i : ch - 1) * nu + j] += df;
        }
        double de = fmax(DV, fabs(e)) * DV;
        grad[nz - 1] = (cost(&s, X, U, e + de) - cost(&s, X, U, e - de)) / (2 * de);
    }
    if (ceq) {
        const double h = s.continuous ? params[0] / 2.0 : 0.0;
        double *xk = w, *xk1 = xk + nx, *uk = xk1 + nx, *fp = uk + nu, *fm = fp + nx, *fk = fm + nx, *fk1 = fk + nx;
        if (Jeq) memset(Jeq, 0, sizeof(double) * (size_t)ph * nx * nz);
        for (int i = 0; i < ph; ++i) {
            memcpy(xk, X + i * nx, sizeof(double) * nx); memcpy(xk1, X + (i + 1) * nx, sizeof(double) * nx);
            memcpy(uk, U + i * nu, sizeof(double) * nu);
            const int blk = i < ch ? i : ch - 1;
            if (s.continuous) {
                field(&s, fk, xk, uk); field(&s, fk1, xk1, uk);
                for (int j = 0; j < nx; ++j) ceq[i * nx + j] = xk[j] + (h * (fk[j] + fk1[j])) - xk1[j];
            } else {
                field(&s, fk, xk, uk);
                for (int j = 0; j < nx; ++j) ceq[i * nx + j] = xk1[j] - fk[j];
            }
            if (!Jeq) continue;
            for (int q = 0; q < nx; ++q) {
                double dx = DV * fmax(fabs(xk[q]), 1.0), keep = xk[q];
                xk[q] = keep + dx; field(&s, fp, xk, uk); xk[q] = keep - dx; field(&s, fm, xk, uk); xk[q] = keep;
                if (s.continuous) {
                    if (i > 0) for (int r = 0; r < nx; ++r) Jeq[(size_t)(i * nx + r) * nz + (i - 1) * nx + q] = (r == q ? 1.0 : 0.0) + h * ((fp[r] - fm[r]) / (2 * dx));
                    double dx1 = DV * fmax(fabs(xk1[q]), 1.0), keep1 = xk1[q];
                    xk1[q] = keep1 + dx1; field(&s, fp, xk1, uk); xk1[q] = keep1 - dx1; field(&s, fm, xk1, uk); xk1[q] = keep1;
                    for (int r = 0; r < nx; ++r) Jeq[(size_t)(i * nx + r) * nz + i * nx + q] = (r == q ? -1.0 : 0.0) + h * ((fp[r] - fm[r]) / (2 * dx1));
                } else {
                    if (i > 0) for (int r = 0; r < nx; ++r) Jeq[(size_t)(i * nx + r) * nz + (i - 1) * nx + q] = -((fp[r] - fm[r]) / (2 * dx));
                    for (int r = 0; r < nx; ++r) Jeq[(size_t)(i * nx + r) * nz + i * nx + q] = (r == q ? 1.0 : 0.0);
                }
            }
            for (int q = 0; q < nu; ++q) {
                double du = DV * fmax(fabs(uk[q]), 1.0), keep = uk[q];
                uk[q] = keep + du; field(&s, fp, xk, uk); uk[q] = keep - du; field(&s, fm, xk, uk); uk[q] = keep;
                double* col = Jeq + ph * nx + blk * nu + q;
                if (s.continuous) {
                    for (int r = 0; r < nx; ++r) fk[r] = (fp[r] - fm[r]) / (2 * du);          /* B_k */
                    uk[q] = keep + du; field(&s, fp, xk1, uk); uk[q] = keep - du; field(&s, fm, xk1, uk); uk[q] = keep;
                    for (int r = 0; r < nx; ++r) col[(size_t)(i * nx + r) * nz] += h * (fk[r] + (fp[r] - fm[r]) / (2 * du));
                } else {
                    for (int r = 0; r < nx; ++r) col[(size_t)(i * nx + r) * nz] += -((fp[r] - fm[r]) / (2 * du));
                }
            }
        }
    }
    if (cin) {
        ineq(&s, X, U, e, cin);
        if (Jin) {
            memset(Jin, 0, sizeof(double) * (size_t)ni * nz);
            for (int i = 0; i < ph; ++i) for (int j = 0; j < nx; ++j) {
                double dx = DV * step_ref(X, nx, ph, j), keep = X[(i + 1) * nx + j];
                X[(i + 1) * nx + j] = keep + dx; ineq(&s, X, U, e, cp);
                X[(i + 1) * nx + j] = keep - dx; ineq(&s, X, U, e, cm);
                X[(i + 1) * nx + j] = keep;
                for (int r = 0; r < ni; ++r) Jin[(size_t)r * nz + i * nx + j] = (cp[r] - cm[r]) / (2 * dx);
            }
            for (int i = 0; i < ph; ++i) for (int j = 0; j < nu; ++j) {     /* every one of the ph rows alone */
                double du = DV * step_ref(U, nu, ph, j), keep = U[i * nu + j];
                U[i * nu + j] = keep + du; ineq(&s, X, U, e, cp);
                U[i * nu + j] = keep - du; ineq(&s, X, U, e, cm);
                U[i * nu + j] = keep;
                const int blk = i < ch ? i : ch - 1;
                for (int r = 0; r < ni; ++r) Jin[(size_t)r * nz + ph * nx + blk * nu + j] += (cp[r] - cm[r]) / (2 * du);
            }
            double de = fmax(DV, fabs(e)) * DV;
            ineq(&s, X, U, e + de, cp); ineq(&s, X, U, e - de, cm);
            for (int r = 0; r < ni; ++r) Jin[(size_t)r * nz + nz - 1] = (cp[r] - cm[r]) / (2 * de);
        }
    }
    free(X);
    return 0;
}
