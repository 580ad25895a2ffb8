/*
 * acoc_oracle.c -- CPU restatement of the reference's regularized-Newton optimal-control path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under aircraftoptimalcontrol_b200/ links, loads or calls this
 * file; it exists so that tests/ (and bench.py's cpu_baseline / --impl reference legs) can check and
 * time the CUDA path against an independent statement of the reference's algorithm.  It is written
 * dense and literal on purpose (7x7 augmented LQ, general 2x2 inverse by pivoted LU, no sparsity
 * tricks) so that it does NOT share structure with the CUDA kernels it checks.
 *
 * Reference files restated (paths into MohamedAtwan/AirCraftOptimalControl):
 *   aircraft_simplified.py:25-69   Cost.stagecost          -> orc_stagecost
 *   aircraft_simplified.py:71-97   Cost.termcost           -> orc_termcost
 *   aircraft_simplified.py:212-261 dragForce / liftForce   -> inside orc_step
 *   aircraft_simplified.py:263-393 Dynamics.step           -> orc_step
 *   aircraft_simplified.py:397-404 tensorCont              -> inside orc_step (lmbd != NULL)
 *   aircraft_simplified.py:126-148 get_initial_trajectory  -> orc_initial_trajectory (float64 arithmetic)
 *   optcon.py:176-200              get_update              -> orc_rollout
 *   optcon.py:204-273,327          armijo_stepsize         -> orc_armijo
 *   optcon.py:341-505              NewtonMethod.optimize   -> orc_newton
 *   optcon.py:27-174               GradientMethod.optimize -> orc_gradient (line-search call repaired, see there)
 *   optcon.py:533-771              ltv_LQR                 -> orc_ltv_lqr
 *   lqr_tracking.py:245-283        lqr_tracking            -> orc_lqr_tracking
 *
 * Parity pin: tests/test_oracle_golden.py checks every function here against fixtures generated from
 * the live, unmodified Python reference (oracle/gen_golden.py -> tests/golden/*.npz).
 *
 * Numerics: IEEE double, no FMA contraction (compile with -ffp-contract=off), libm sin/cos/pow --
 * numpy's float64 scalar sin/cos/** resolve to the same glibc routines, which makes the next-state
 * formulas (aircraft_simplified.py:303-310) bit-identical to the reference, including its rounding of
 * the next state to float32 (aircraft_simplified.py:300) when quant_f32 != 0.
 *
 * Array conventions (all C-contiguous double): a trajectory of one instance is stored exactly like the
 * reference stores it, component-major: xx[i*TT + t] (shape (6,TT)), uu[j*TT + t] (shape (2,TT)).
 * Matrices are row-major.  A = fx.T and B = fu.T, i.e. A[i][j] = d f_i / d x_j.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NS 6
#define NI 2
#define NA 7 /* augmented state dimension, optcon.py:657 */

/* params = {cd0, cda, cla, m, g, S, rho, J, dt}  (aircraft_simplified.py:108-118) */
enum { P_CD0, P_CDA, P_CLA, P_M, P_G, P_S, P_RHO, P_J, P_DT, P_COUNT };

static inline double q32(double v, int quant_f32) { return quant_f32 ? (double)(float)v : v; }

/* ------------------------------------------------------------------------------------------------ */
/* Dynamics.step                                                                                     */
/* ------------------------------------------------------------------------------------------------ */
/*
 * xxp[6]; A[36] = fx.T; B[12] = fu.T (6x2 row-major).
 * If lmbd == NULL: fxx[216] with fxx[(i*6+j)*6+k] = d2 f_k / dx_i dx_j  (aircraft_simplified.py:368-371),
 *                  fux[72]  with fux[(a*6+j)*6+k] = d2 f_k / du_a dx_j  (:375-379).
 * If lmbd != NULL: fxx[36], fux[12] hold the contractions sum_k lmbd[k]*T[:,:,k] (:384-388, :397-404).
 * Any output pointer may be NULL.  fuu is identically zero (:382) and is not produced.
 */
void orc_step(const double *prm, const double *x, const double *u, const double *lmbd, int quant_f32,
              double *xxp, double *A, double *B, double *fxx, double *fux)
{
    const double Cd0 = prm[P_CD0], Cda = prm[P_CDA], Cla = prm[P_CLA], m = prm[P_M], g = prm[P_G];
    const double S = prm[P_S], rho = prm[P_RHO], J = prm[P_J], dt = prm[P_DT];
    const double V = x[2], th = x[3], q = x[4], gam = x[5], T = u[0], M = u[1];
    const double alpha = th - gam;                       /* :295 */
    const double sg = sin(gam), cg = cos(gam), sa = sin(alpha), ca = cos(alpha);
    /* :228 and :253, evaluated left to right exactly as python does */
    const double V2 = pow(V, 2.0), a2 = pow(alpha, 2.0);
    const double D = 0.5 * rho * V2 * S * (Cd0 + Cda * a2);
    const double L = 0.5 * rho * V2 * S * Cla * alpha;

    if (xxp) { /* :303-310 */
        xxp[0] = q32(x[0] + dt * V * cg, quant_f32);
        xxp[1] = q32(x[1] - dt * V * sg, quant_f32);
        xxp[2] = q32(V + (dt / m) * (-D - m * g * sg + T * ca), quant_f32);
        xxp[3] = q32(th + dt * q, quant_f32);
        xxp[4] = q32(q + dt * (M / J), quant_f32);
        xxp[5] = q32(gam + (dt / (m * V)) * (L - m * g * cg + T * sa), quant_f32);
    }

    const double k = S * rho;          /* the derivatives are written in terms of k = rho*S */
    const double dtm = dt / m;
    if (A) { /* :316-322 */
        memset(A, 0, 36 * sizeof(double));
        A[0 * 6 + 0] = 1.0; A[0 * 6 + 2] = dt * cg;  A[0 * 6 + 5] = -dt * V * sg;
        A[1 * 6 + 1] = 1.0; A[1 * 6 + 2] = -dt * sg; A[1 * 6 + 5] = -dt * V * cg;
        A[2 * 6 + 2] = 1.0 - dtm * k * V * (Cd0 + Cda * alpha * alpha);
        A[2 * 6 + 3] = -dtm * (Cda * k * alpha * V * V + T * sa);
        A[2 * 6 + 5] = dtm * (Cda * k * alpha * V * V + T * sa - g * m * cg);
        A[3 * 6 + 3] = 1.0; A[3 * 6 + 4] = dt;
        A[4 * 6 + 4] = 1.0;
        A[5 * 6 + 2] = Cla * k * dt * alpha / m - dt * (0.5 * Cla * k * alpha * V * V + T * sa - g * m * cg) / (m * V * V);
        A[5 * 6 + 3] = dt * (0.5 * Cla * k * V * V + T * ca) / (m * V);
        A[5 * 6 + 5] = 1.0 - dt * (0.5 * Cla * k * V * V + T * ca - g * m * sg) / (m * V);
    }
    if (B) { /* :324-325 */
        memset(B, 0, 12 * sizeof(double));
        B[2 * 2 + 0] = dt * ca / m;
        B[4 * 2 + 1] = dt / J;
        B[5 * 2 + 0] = dt * sa / (m * V);
    }
    if (!fxx && !fux) return;

    /* second-order terms: four non-zero Hessians (f_X, f_Z, f_V, f_gamma), :339-371 */
    double H[4][36];
    static const int Hk[4] = {0, 1, 2, 5};
    memset(H, 0, sizeof(H));
#define SYM(h, i, j, v) do { (h)[(i) * 6 + (j)] = (v); (h)[(j) * 6 + (i)] = (v); } while (0)
    SYM(H[0], 2, 5, -dt * sg);              H[0][5 * 6 + 5] = -dt * V * cg;
    SYM(H[1], 2, 5, -dt * cg);              H[1][5 * 6 + 5] = dt * V * sg;
    {
        const double c1 = Cda * k * dt * V * (2.0 * alpha) / m;
        const double c2 = dt * (Cda * k * V * V + T * ca) / m;
        H[2][2 * 6 + 2] = -(k * dt * (Cd0 + Cda * alpha * alpha)) / m;
        SYM(H[2], 2, 3, -c1); SYM(H[2], 2, 5, c1);
        H[2][3 * 6 + 3] = -c2; SYM(H[2], 3, 5, c2);
        H[2][5 * 6 + 5] = -(dt * (Cda * k * V * V + T * ca - g * m * sg)) / m;
    }
    {
        const double lift = 0.5 * Cla * k * V * V;
        const double e23 = Cla * k * dt / m - dt * (lift + T * ca) / (m * V * V);
        const double e25 = dt * (lift + T * ca - g * m * sg) / (m * V * V) - Cla * k * dt / m;
        const double e33 = dt * T * sa / (m * V);
        H[3][2 * 6 + 2] = 2.0 * dt * (lift * alpha + T * sa - g * m * cg) / (m * V * V * V) - Cla * k * dt * alpha / (m * V);
        SYM(H[3], 2, 3, e23); SYM(H[3], 2, 5, e25);
        H[3][3 * 6 + 3] = -e33; SYM(H[3], 3, 5, e33);
        H[3][5 * 6 + 5] = -(dt * (T * sa - g * m * cg)) / (m * V);
    }
#undef SYM
    /* mixed terms, thrust row only: :375-379 */
    double U2[6] = {0, 0, 0, -dt * sa / m, 0, dt * sa / m};                                   /* f_V */
    double U5[6] = {0, 0, -dt * sa / (m * V * V), dt * ca / (m * V), 0, -dt * ca / (m * V)};  /* f_gamma */

    if (lmbd) { /* tensorCont, :397-404: T = sum_i P[:,:,i]*a[i], i ascending from a zero array */
        if (fxx) {
            for (int e = 0; e < 36; ++e) { /* slices 3 and 4 are all-zero and contribute +0 */
                double acc = 0.0;
                acc += H[0][e] * lmbd[0]; acc += H[1][e] * lmbd[1]; acc += H[2][e] * lmbd[2]; acc += H[3][e] * lmbd[5];
                fxx[e] = acc;
            }
        }
        if (fux) {
            for (int j = 0; j < 6; ++j) {
                double acc = 0.0;
                acc += U2[j] * lmbd[2]; acc += U5[j] * lmbd[5];
                fux[j] = acc;
                fux[6 + j] = 0.0;
            }
        }
    } else {
        if (fxx) {
            memset(fxx, 0, 216 * sizeof(double));
            for (int c = 0; c < 4; ++c)
                for (int e = 0; e < 36; ++e) fxx[e * 6 + Hk[c]] = H[c][e];
        }
        if (fux) {
            memset(fux, 0, 72 * sizeof(double));
            for (int j = 0; j < 6; ++j) { fux[j * 6 + 2] = U2[j]; fux[j * 6 + 5] = U5[j]; }
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Cost.stagecost / Cost.termcost (dense Q, R)                                                       */
/* ------------------------------------------------------------------------------------------------ */
double orc_stagecost(const double *Q, const double *R, const double *x, const double *u,
                     const double *xr, const double *ur, double *lx, double *lu)
{
    double dx[NS], du[NI], Qdx[NS], Rdu[NI];
    for (int i = 0; i < NS; ++i) dx[i] = x[i] - xr[i];
    for (int i = 0; i < NI; ++i) du[i] = u[i] - ur[i];
    for (int i = 0; i < NS; ++i) { double a = 0; for (int j = 0; j < NS; ++j) a += Q[i * NS + j] * dx[j]; Qdx[i] = a; }
    for (int i = 0; i < NI; ++i) { double a = 0; for (int j = 0; j < NI; ++j) a += R[i * NI + j] * du[j]; Rdu[i] = a; }
    double sx = 0, su = 0;
    for (int i = 0; i < NS; ++i) sx += dx[i] * Qdx[i];
    for (int i = 0; i < NI; ++i) su += du[i] * Rdu[i];
    if (lx) memcpy(lx, Qdx, sizeof(Qdx));
    if (lu) memcpy(lu, Rdu, sizeof(Rdu));
    return 0.5 * sx + 0.5 * su; /* :61 */
}

double orc_termcost(const double *QT, const double *x, const double *xr, double *lTx)
{
    double dx[NS], Qdx[NS];
    for (int i = 0; i < NS; ++i) dx[i] = x[i] - xr[i];
    for (int i = 0; i < NS; ++i) { double a = 0; for (int j = 0; j < NS; ++j) a += QT[i * NS + j] * dx[j]; Qdx[i] = a; }
    /* :92 is (0.5*dx.T@QT)@dx ; the 0.5 is applied before the second product */
    double s = 0;
    for (int j = 0; j < NS; ++j) {
        double a = 0;
        for (int i = 0; i < NS; ++i) a += (0.5 * dx[i]) * QT[i * NS + j];
        s += a * dx[j];
    }
    if (lTx) memcpy(lTx, Qdx, sizeof(Qdx));
    return s;
}

/* cost of a whole trajectory, optcon.py:419-424: stage costs t = 0..TT-2 (ascending) then terminal */
double orc_traj_cost(const double *Q, const double *R, const double *QT, int TT,
                     const double *xx, const double *uu, const double *xr, const double *ur)
{
    double J = 0.0, x[NS], u[NI], r[NS], v[NI];
    for (int t = 0; t < TT - 1; ++t) {
        for (int i = 0; i < NS; ++i) { x[i] = xx[i * TT + t]; r[i] = xr[i * TT + t]; }
        for (int i = 0; i < NI; ++i) { u[i] = uu[i * TT + t]; v[i] = ur[i * TT + t]; }
        J += orc_stagecost(Q, R, x, u, r, v, NULL, NULL);
    }
    for (int i = 0; i < NS; ++i) { x[i] = xx[i * TT + TT - 1]; r[i] = xr[i * TT + TT - 1]; }
    J += orc_termcost(QT, x, r, NULL);
    return J;
}

/* ------------------------------------------------------------------------------------------------ */
/* small dense helpers                                                                               */
/* ------------------------------------------------------------------------------------------------ */
/* C(m x n) = A(m x k) @ B(k x n), plain triple loop, k ascending */
static void mm(int m, int k, int n, const double *A, const double *B, double *C)
{
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double a = 0.0;
            for (int l = 0; l < k; ++l) a += A[i * k + l] * B[l * n + j];
            C[i * n + j] = a;
        }
}
static void tr(int m, int n, const double *A, double *At)
{
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) At[j * m + i] = A[i * n + j];
}
/* inverse of a general 2x2 by LU with partial pivoting followed by two triangular solves against the
 * identity -- the sequence LAPACK's gesv runs for np.linalg.inv (optcon.py:728, :751). */
static void inv2(const double *Min, double *X)
{
    double a[4] = {Min[0], Min[1], Min[2], Min[3]};
    int piv = (fabs(a[2]) > fabs(a[0])) ? 1 : 0;
    if (piv) { double t0 = a[0], t1 = a[1]; a[0] = a[2]; a[1] = a[3]; a[2] = t0; a[3] = t1; }
    const double l = a[2] / a[0];
    const double u11 = a[3] - l * a[1];
    for (int c = 0; c < 2; ++c) {
        double b0 = (c == 0) ? 1.0 : 0.0, b1 = (c == 1) ? 1.0 : 0.0;
        if (piv) { double t = b0; b0 = b1; b1 = t; }
        b1 = b1 - l * b0;
        const double x1 = b1 / u11;
        const double x0 = (b0 - a[1] * x1) / a[0];
        X[0 * 2 + c] = x0; X[1 * 2 + c] = x1;
    }
}
/* "np.all(np.linalg.eigvals(MM) > 0)" for a real 2x2 (optcon.py:745).  Complex pairs compare by real part. */
static int all_eig_positive2(const double *M)
{
    const double a = M[0], b = M[1], c = M[2], d = M[3];
    const double h = 0.5 * (a - d), disc = h * h + b * c, mid = 0.5 * (a + d);
    if (disc >= 0.0) { const double r = sqrt(disc); return (mid - r > 0.0) && (mid + r > 0.0); }
    if (disc < 0.0) return mid > 0.0;
    return 0; /* NaN: every comparison is false */
}

/* ------------------------------------------------------------------------------------------------ */
/* ltv_LQR, optcon.py:533-771 (ns = 6, ni = 2).  Time-major inputs:                                  */
/*   A[t][6][6], B[t][6][2], Q[t][6][6], R[t][2][2], S[t][2][6], Qf[6][6], x0[6]                     */
/*   affine: q[t][6], r[t][2], qf[6] or all NULL (non-augmented branch :699-714)                     */
/* Outputs (time-major): K[t][2][n] with n = 7 (augmented) or 6, P[t][n][n] (may be NULL),           */
/*   xout[t][6], uout[t][2].  Entries at t = TT-1 of K and uout stay zero like the reference's.       */
/* ------------------------------------------------------------------------------------------------ */
int orc_ltv_lqr(int TT, const double *A, const double *B, const double *Q, const double *R, const double *S,
                const double *Qf, const double *x0, const double *q, const double *r, const double *qf,
                double *K, double *P, double *xout, double *uout, int *n_regularized)
{
    const int aug = (q != NULL) || (r != NULL) || (qf != NULL);
    const int n = aug ? NA : NS, off = aug ? 1 : 0;
    double *PP = (double *)calloc((size_t)TT * n * n, sizeof(double));
    double *AAt = (double *)calloc((size_t)n * n, sizeof(double)), *BBt = (double *)calloc((size_t)n * NI, sizeof(double));
    double *QQt = (double *)calloc((size_t)n * n, sizeof(double)), *SSt = (double *)calloc((size_t)NI * n, sizeof(double));
    double RRt[4], G[4], Gi[4], MM[4], MMi[4];
    double AT[NA * NA], BT[NI * NA], t1[NA * NA], t2[NA * NA], Mx[NI * NA], MxT[NA * NI], BPB[4], t3[NA * NI], t4[NA * NA], t5[NI * NA];
    if (!PP || !AAt || !BBt || !QQt || !SSt) return -1;
    if (n_regularized) *n_regularized = 0;

#define BUILD(t) do { \
        memset(AAt, 0, sizeof(double) * n * n); memset(BBt, 0, sizeof(double) * n * NI); \
        memset(QQt, 0, sizeof(double) * n * n); memset(SSt, 0, sizeof(double) * NI * n); \
        if (aug) { AAt[0] = 1.0; \
            for (int i = 0; i < NS; ++i) { const double hq = q ? 0.5 * q[(t) * NS + i] : 0.0; \
                QQt[(i + 1) * n + 0] = hq; QQt[0 * n + (i + 1)] = hq; } \
            for (int a = 0; a < NI; ++a) SSt[a * n + 0] = r ? 0.5 * r[(t) * NI + a] : 0.0; } \
        for (int i = 0; i < NS; ++i) for (int j = 0; j < NS; ++j) { \
            AAt[(i + off) * n + (j + off)] = A[((t) * NS + i) * NS + j]; \
            QQt[(i + off) * n + (j + off)] = Q[((t) * NS + i) * NS + j]; } \
        for (int i = 0; i < NS; ++i) for (int a = 0; a < NI; ++a) BBt[(i + off) * NI + a] = B[((t) * NS + i) * NI + a]; \
        for (int a = 0; a < NI; ++a) for (int j = 0; j < NS; ++j) SSt[a * n + (j + off)] = S[((t) * NI + a) * NS + j]; \
        for (int e = 0; e < 4; ++e) RRt[e] = R[(t) * 4 + e]; \
        tr(n, n, AAt, AT); tr(n, NI, BBt, BT); \
    } while (0)

    /* terminal condition, :688-690, :716 */
    {
        double *Pf = PP + (size_t)(TT - 1) * n * n;
        if (aug) for (int i = 0; i < NS; ++i) { const double h = qf ? 0.5 * qf[i] : 0.0; Pf[(i + 1) * n] = h; Pf[i + 1] = h; }
        for (int i = 0; i < NS; ++i) for (int j = 0; j < NS; ++j) Pf[(i + off) * n + (j + off)] = Qf[i * NS + j];
    }
    /* Riccati, :719-728 */
    for (int t = TT - 2; t >= 0; --t) {
        const double *Pn = PP + (size_t)(t + 1) * n * n;
        double *Pt = PP + (size_t)t * n * n;
        BUILD(t);
        mm(n, n, n, AT, Pn, t1); mm(n, n, n, t1, AAt, t2);            /* A'PA */
        mm(NI, n, n, BT, Pn, t5); mm(NI, n, n, t5, AAt, Mx);           /* B'PA */
        for (int e = 0; e < NI * n; ++e) Mx[e] += SSt[e];
        mm(NI, n, NI, t5, BBt, BPB);
        for (int e = 0; e < 4; ++e) G[e] = RRt[e] + BPB[e];
        inv2(G, Gi);
        tr(NI, n, Mx, MxT);
        mm(n, NI, NI, MxT, Gi, t3); mm(n, NI, n, t3, Mx, t4);
        for (int e = 0; e < n * n; ++e) Pt[e] = QQt[e] + t2[e] - t4[e];
    }
    /* gains, :732-751 */
    const int kn = n;
    memset(K, 0, sizeof(double) * (size_t)TT * NI * kn);
    for (int t = 0; t < TT - 1; ++t) {
        const double *Pn = PP + (size_t)(t + 1) * n * n;
        BUILD(t);
        mm(NI, n, n, BT, Pn, t5); mm(NI, n, NI, t5, BBt, BPB);
        for (int e = 0; e < 4; ++e) MM[e] = RRt[e] + BPB[e];
        if (!all_eig_positive2(MM)) { MM[0] += 0.5; MM[3] += 0.5; if (n_regularized) ++*n_regularized; }
        inv2(MM, MMi);
        mm(NI, n, n, t5, AAt, Mx);
        for (int e = 0; e < NI * n; ++e) Mx[e] += SSt[e];
        double nMi[4] = {-MMi[0], -MMi[1], -MMi[2], -MMi[3]};
        mm(NI, NI, n, nMi, Mx, K + (size_t)t * NI * kn);
    }
    /* forward pass, :756-762 */
    {
        double xa[NA], xn[NA], ua[NI];
        memset(xa, 0, sizeof(xa));
        if (aug) xa[0] = 1.0;
        for (int i = 0; i < NS; ++i) xa[i + off] = x0[i];
        memset(uout, 0, sizeof(double) * (size_t)TT * NI);
        memset(xout, 0, sizeof(double) * (size_t)TT * NS);
        for (int i = 0; i < NS; ++i) xout[i] = x0[i];
        for (int t = 0; t < TT - 1; ++t) {
            BUILD(t);
            const double *Kt = K + (size_t)t * NI * kn;
            for (int a = 0; a < NI; ++a) { double s = 0; for (int j = 0; j < n; ++j) s += Kt[a * kn + j] * xa[j]; ua[a] = s; }
            for (int i = 0; i < n; ++i) {
                double s1 = 0, s2 = 0;
                for (int j = 0; j < n; ++j) s1 += AAt[i * n + j] * xa[j];
                for (int a = 0; a < NI; ++a) s2 += BBt[i * NI + a] * ua[a];
                xn[i] = s1 + s2;
            }
            memcpy(xa, xn, sizeof(double) * n);
            if (aug) xa[0] = 1.0; /* the reference fills row 0 with ones (:696); A~[0,0]=1 keeps it 1 */
            for (int a = 0; a < NI; ++a) uout[t * NI + a] = ua[a];
            for (int i = 0; i < NS; ++i) xout[(t + 1) * NS + i] = xa[i + off];
        }
    }
#undef BUILD
    if (P) memcpy(P, PP, sizeof(double) * (size_t)TT * n * n);
    free(PP); free(AAt); free(BBt); free(QQt); free(SSt);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* get_update (optcon.py:176-200) and the Armijo candidate evaluation (:247-264)                     */
/* ------------------------------------------------------------------------------------------------ */
/* x' trajectory for u' = u + s*du; xx_out (6,TT), uu_out (2,TT) (last input column zero). Returns J(x',u')
 * when Q != NULL (refs needed), else 0. */
double orc_rollout(const double *prm, int quant_f32, int TT, const double *x0, const double *uu, const double *du,
                   double s, const double *Q, const double *R, const double *QT, const double *xr, const double *ur,
                   double *xx_out, double *uu_out)
{
    double x[NS], xn[NS], u[NI], r[NS], v[NI], J = 0.0;
    memcpy(x, x0, sizeof(x));
    for (int t = 0; t < TT - 1; ++t) {
        for (int a = 0; a < NI; ++a) u[a] = uu[a * TT + t] + s * du[a * TT + t];
        if (xx_out) for (int i = 0; i < NS; ++i) xx_out[i * TT + t] = x[i];
        if (uu_out) for (int a = 0; a < NI; ++a) uu_out[a * TT + t] = u[a];
        if (Q) {
            for (int i = 0; i < NS; ++i) r[i] = xr[i * TT + t];
            for (int a = 0; a < NI; ++a) v[a] = ur[a * TT + t];
            J += orc_stagecost(Q, R, x, u, r, v, NULL, NULL);
        }
        orc_step(prm, x, u, NULL, quant_f32, xn, NULL, NULL, NULL, NULL);
        memcpy(x, xn, sizeof(x));
    }
    if (xx_out) for (int i = 0; i < NS; ++i) xx_out[i * TT + TT - 1] = x[i];
    if (uu_out) for (int a = 0; a < NI; ++a) uu_out[a * TT + TT - 1] = 0.0;
    if (Q) { for (int i = 0; i < NS; ++i) r[i] = xr[i * TT + TT - 1]; J += orc_termcost(QT, x, r, NULL); }
    return J;
}

/* armijo_stepsize, optcon.py:240-273,327.  costs[i] receives the cost of every candidate tried.
 * Returns the step; *n_tried = number of rollouts; *accepted = 0 when the search ran out (the returned
 * step stepsize_0*beta^maxiters has then NOT been tested, as in the reference). */
double orc_armijo(const double *prm, int quant_f32, int TT, const double *x0, const double *uu, const double *du,
                  const double *Q, const double *R, const double *QT, const double *xr, const double *ur,
                  double JP, double descent, double stepsize_0, double cc, double beta, int maxiters,
                  double *costs, int *n_tried, int *accepted)
{
    double s = stepsize_0;
    int ok = 0, ii;
    for (ii = 0; ii < maxiters; ++ii) {
        const double Jt = orc_rollout(prm, quant_f32, TT, x0, uu, du, s, Q, R, QT, xr, ur, NULL, NULL);
        if (costs) costs[ii] = Jt;
        if (Jt > JP + cc * s * descent) s = beta * s; else { ok = 1; ++ii; break; }
    }
    if (n_tried) *n_tried = ii;
    if (accepted) *accepted = ok;
    return s;
}

/* ------------------------------------------------------------------------------------------------ */
/* NewtonMethod.optimize, optcon.py:341-505                                                          */
/* ------------------------------------------------------------------------------------------------ */
/*
 * Inputs: xx_ref (6,TT), uu_ref (2,TT), xx_init (6,TT), uu_init (2,TT).
 * hist_J / hist_descent / hist_step / hist_ncand (length >= max_iters) receive, per executed loop body kk,
 * JJ[kk], descent[kk], the Armijo step and the number of candidates tried.  *iters = bodies executed.
 * xx_star/uu_star = what optimize returns (slot max_iters-1 after the break logic, :499-505);
 * xx_last/uu_last (optional) = last iterate produced by get_update.
 * exact_after: exact Hessian iff kk > exact_after (8 in the reference, :443).
 * n_iters_cap: stop after this many bodies even if not converged (<=0: no cap) -- used by the timing legs.
 * Returns 0, or -1 on allocation failure.
 */
int orc_newton(const double *prm, int quant_f32, int TT, const double *Q, const double *R, const double *QT,
               const double *xx_ref, const double *uu_ref, const double *xx_init, const double *uu_init,
               int max_iters, double stepsize_0, double cc, double beta, int armijo_maxiters, int exact_after,
               double term_cond, int n_iters_cap,
               double *hist_J, double *hist_descent, double *hist_step, int *hist_ncand, int *iters,
               double *xx_star, double *uu_star, double *xx_last, double *uu_last, int *n_regularized)
{
    const size_t nx = (size_t)NS * TT, nu = (size_t)NI * TT;
    /* three rotating iterate slots are enough: kk-1, kk, kk+1 */
    double *X[3], *U[3];
    for (int s = 0; s < 3; ++s) { X[s] = (double *)calloc(nx, sizeof(double)); U[s] = (double *)calloc(nu, sizeof(double)); }
    double *A = (double *)calloc((size_t)TT * 36, sizeof(double)), *B = (double *)calloc((size_t)TT * 12, sizeof(double));
    double *Qt = (double *)calloc((size_t)TT * 36, sizeof(double)), *Rt = (double *)calloc((size_t)TT * 4, sizeof(double));
    double *St = (double *)calloc((size_t)TT * 12, sizeof(double)), *qq = (double *)calloc((size_t)TT * 6, sizeof(double));
    double *rr = (double *)calloc((size_t)TT * 2, sizeof(double)), *lam = (double *)calloc((size_t)TT * 6, sizeof(double));
    double *K = (double *)calloc((size_t)TT * NI * NA, sizeof(double)), *dxo = (double *)calloc((size_t)TT * 6, sizeof(double));
    double *duo = (double *)calloc((size_t)TT * 2, sizeof(double)), *du = (double *)calloc(nu, sizeof(double));
    if (!X[2] || !U[2] || !A || !B || !Qt || !Rt || !St || !qq || !rr || !lam || !K || !dxo || !duo || !du) return -1;
    memcpy(X[0], xx_init, nx * sizeof(double));
    memcpy(U[0], uu_init, nu * sizeof(double));
    double x0[NS], zero6[NS] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < NS; ++i) x0[i] = xx_init[i * TT]; /* :398 */
    int stop_at = -1, kk, nreg_total = 0;
    const int bodies = max_iters - 1;
    for (kk = 0; kk < bodies; ++kk) {
        if (n_iters_cap > 0 && kk >= n_iters_cap) break;
        const double *xx = X[kk % 3], *uu = U[kk % 3];
        const double JJ = orc_traj_cost(Q, R, QT, TT, xx, uu, xx_ref, uu_ref); /* :417-424 */
        double x[NS], u[NI], xr[NS], ur[NI], lx[NS], lu[NI], fxxc[36], fuxc[12];
        for (int i = 0; i < NS; ++i) { x[i] = xx[i * TT + TT - 1]; xr[i] = xx_ref[i * TT + TT - 1]; }
        orc_termcost(QT, x, xr, lam + (size_t)(TT - 1) * 6); /* :429-432 */
        for (int t = TT - 2; t >= 0; --t) { /* :434-464 */
            for (int i = 0; i < NS; ++i) { x[i] = xx[i * TT + t]; xr[i] = xx_ref[i * TT + t]; }
            for (int a = 0; a < NI; ++a) { u[a] = uu[a * TT + t]; ur[a] = uu_ref[a * TT + t]; }
            orc_stagecost(Q, R, x, u, xr, ur, lx, lu);
            const double *ln = lam + (size_t)(t + 1) * 6;
            orc_step(prm, x, u, ln, quant_f32, NULL, A + (size_t)t * 36, B + (size_t)t * 12, fxxc, fuxc);
            const int exact = kk > exact_after;
            for (int e = 0; e < 36; ++e) Qt[(size_t)t * 36 + e] = Q[e] + (exact ? fxxc[e] : 0.0);
            for (int e = 0; e < 4; ++e) Rt[(size_t)t * 4 + e] = R[e];
            for (int e = 0; e < 12; ++e) St[(size_t)t * 12 + e] = exact ? fuxc[e] : 0.0;
            for (int i = 0; i < NS; ++i) qq[(size_t)t * 6 + i] = lx[i];
            for (int a = 0; a < NI; ++a) rr[(size_t)t * 2 + a] = lu[a];
            const double *At = A + (size_t)t * 36;
            for (int i = 0; i < NS; ++i) { /* lambda_t = A' lambda_{t+1} + q, :461 */
                double s = 0.0;
                for (int j = 0; j < NS; ++j) s += At[j * 6 + i] * ln[j];
                lam[(size_t)t * 6 + i] = s + lx[i];
            }
        }
        int nreg = 0;
        orc_ltv_lqr(TT, A, B, Qt, Rt, St, QT, zero6, qq, rr, lam + (size_t)(TT - 1) * 6, K, NULL, dxo, duo, &nreg); /* :468-470 */
        nreg_total += nreg;
        double descent = 0.0;
        for (int t = TT - 2; t >= 0; --t) { /* :474-477 */
            const double *Bt = B + (size_t)t * 12, *ln = lam + (size_t)(t + 1) * 6;
            double tmp = 0.0;
            for (int a = 0; a < NI; ++a) {
                double gsum = 0.0;
                for (int i = 0; i < NS; ++i) gsum += Bt[i * 2 + a] * ln[i];
                tmp += (gsum + rr[(size_t)t * 2 + a]) * duo[t * 2 + a];
            }
            descent += tmp;
        }
        for (int t = 0; t < TT; ++t) for (int a = 0; a < NI; ++a) du[a * TT + t] = duo[t * 2 + a];
        int ntried = 0, acc = 0;
        const double s = orc_armijo(prm, quant_f32, TT, x0, uu, du, Q, R, QT, xx_ref, uu_ref, JJ, descent,
                                    stepsize_0, cc, beta, armijo_maxiters, NULL, &ntried, &acc); /* :482 */
        orc_rollout(prm, quant_f32, TT, x0, uu, du, s, NULL, NULL, NULL, NULL, NULL, X[(kk + 1) % 3], U[(kk + 1) % 3]); /* :488-491 */
        if (hist_J) hist_J[kk] = JJ;
        if (hist_descent) hist_descent[kk] = descent;
        if (hist_step) hist_step[kk] = s;
        if (hist_ncand) hist_ncand[kk] = ntried;
        if (descent >= term_cond) { stop_at = kk; ++kk; break; } /* :499-501 */
    }
    const int executed = kk;
    if (iters) *iters = executed;
    if (n_regularized) *n_regularized = nreg_total;
    if (xx_last) memcpy(xx_last, X[executed % 3], nx * sizeof(double));
    if (uu_last) memcpy(uu_last, U[executed % 3], nu * sizeof(double));
    /* result slot, :503-505: max_iters-1 with max_iters := kk on convergence */
    if (stop_at == 0) { /* slot -1 of the history array: never written, all zeros */
        memset(xx_star, 0, nx * sizeof(double)); memset(uu_star, 0, nu * sizeof(double));
    } else if (stop_at > 0) {
        memcpy(xx_star, X[(stop_at - 1) % 3], nx * sizeof(double));
        memcpy(uu_star, U[(stop_at - 1) % 3], nu * sizeof(double));
    } else { /* never converged (or capped): slot max_iters-1 is the last iterate written */
        memcpy(xx_star, X[executed % 3], nx * sizeof(double));
        memcpy(uu_star, U[executed % 3], nu * sizeof(double));
    }
    for (int a = 0; a < NI; ++a) uu_star[a * TT + TT - 1] = uu_star[a * TT + TT - 2]; /* :505 */
    for (int s = 0; s < 3; ++s) { free(X[s]); free(U[s]); }
    free(A); free(B); free(Qt); free(Rt); free(St); free(qq); free(rr); free(lam); free(K); free(dxo); free(duo); free(du);
    return 0;
}

/* batch driver: instance n uses xx_ref + n*ref_stride_x etc. (stride 0 = shared).  OpenMP over instances. */
int orc_newton_batch(int N, int n_threads, const double *prm, int quant_f32, int TT, const double *Q, const double *R, const double *QT,
                     const double *xx_ref, long ref_stride_x, const double *uu_ref, long ref_stride_u,
                     const double *xx_init, const double *uu_init,
                     int max_iters, double stepsize_0, double cc, double beta, int armijo_maxiters, int exact_after,
                     double term_cond, int n_iters_cap,
                     double *hist_J, double *hist_descent, double *hist_step, int *hist_ncand, int *iters,
                     double *xx_star, double *uu_star)
{
    int rc = 0;
    const size_t nx = (size_t)NS * TT, nu = (size_t)NI * TT;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int n = 0; n < N; ++n) {
        int r = orc_newton(prm, quant_f32, TT, Q, R, QT, xx_ref + (size_t)n * ref_stride_x, uu_ref + (size_t)n * ref_stride_u,
                           xx_init + n * nx, uu_init + n * nu, max_iters, stepsize_0, cc, beta, armijo_maxiters, exact_after,
                           term_cond, n_iters_cap,
                           hist_J ? hist_J + (size_t)n * max_iters : NULL, hist_descent ? hist_descent + (size_t)n * max_iters : NULL,
                           hist_step ? hist_step + (size_t)n * max_iters : NULL, hist_ncand ? hist_ncand + (size_t)n * max_iters : NULL,
                           iters ? iters + n : NULL, xx_star + n * nx, uu_star + n * nu, NULL, NULL, NULL);
        if (r) rc = r;
    }
    return rc;
}

/* ------------------------------------------------------------------------------------------------ */
/* GradientMethod.optimize, optcon.py:27-174 (steepest descent), with its line-search call repaired   */
/* ------------------------------------------------------------------------------------------------ */
/*
 * The reference's loop body, literally: cost (:87-93), terminal costate (:98-99), backward costate sweep with
 * deltau_t = -B'lam_{t+1} - lu and descent += deltau'deltau (:101-118), Armijo, get_update (:131), stop when
 * descent <= 1e-6 (:52, :157).  The call at :125 passes 8 of armijo_stepsize's 9 arguments (TypeError in the
 * reference); the repair restated here -- the same one oracle/pyref.py::run_gradient applies to the live
 * reference through a call adapter -- is JP = JJ[kk] and slope = -descent[kk] for the test of :268.
 * hist_descent receives the reference's descent[kk] = sum |deltau|^2 (positive).  Everything else as orc_newton.
 */
int orc_gradient(const double *prm, int quant_f32, int TT, const double *Q, const double *R, const double *QT,
                 const double *xx_ref, const double *uu_ref, const double *xx_init, const double *uu_init,
                 int max_iters, double stepsize_0, double cc, double beta, int armijo_maxiters, double term_cond,
                 double *hist_J, double *hist_descent, double *hist_step, int *hist_ncand, int *iters,
                 double *xx_star, double *uu_star, double *xx_last, double *uu_last, double *du_first)
{
    const size_t nx = (size_t)NS * TT, nu = (size_t)NI * TT;
    double *X[3], *U[3];
    for (int s = 0; s < 3; ++s) { X[s] = (double *)calloc(nx, sizeof(double)); U[s] = (double *)calloc(nu, sizeof(double)); }
    double *du = (double *)calloc(nu, sizeof(double));
    if (!X[2] || !U[2] || !du) return -1;
    memcpy(X[0], xx_init, nx * sizeof(double));
    memcpy(U[0], uu_init, nu * sizeof(double));
    double x0[NS];
    for (int i = 0; i < NS; ++i) x0[i] = xx_init[i * TT]; /* :69 */
    int stop_at = -1, kk;
    for (kk = 0; kk < max_iters - 1; ++kk) { /* :85 */
        const double *xx = X[kk % 3], *uu = U[kk % 3];
        const double JJ = orc_traj_cost(Q, R, QT, TT, xx, uu, xx_ref, uu_ref); /* :87-93 */
        double x[NS], u[NI], xr[NS], ur[NI], lx[NS], lu[NI], lam[NS], lamn[NS], A[36], B[12];
        for (int i = 0; i < NS; ++i) { x[i] = xx[i * TT + TT - 1]; xr[i] = xx_ref[i * TT + TT - 1]; }
        orc_termcost(QT, x, xr, lam); /* :98-99 */
        double descent = 0.0;
        for (int t = TT - 2; t >= 0; --t) { /* :101-118 */
            for (int i = 0; i < NS; ++i) { x[i] = xx[i * TT + t]; xr[i] = xx_ref[i * TT + t]; }
            for (int a = 0; a < NI; ++a) { u[a] = uu[a * TT + t]; ur[a] = uu_ref[a * TT + t]; }
            orc_stagecost(Q, R, x, u, xr, ur, lx, lu);
            orc_step(prm, x, u, NULL, quant_f32, NULL, A, B, NULL, NULL);
            for (int i = 0; i < NS; ++i) { /* lmbd_temp = AA.T@lmbd + aa, :110 */
                double s = 0.0;
                for (int j = 0; j < NS; ++j) s += A[j * 6 + i] * lam[j];
                lamn[i] = s + lx[i];
            }
            double sq = 0.0;
            for (int a = 0; a < NI; ++a) { /* deltau_temp = -BB.T@lmbd - bb, :111 */
                double s = 0.0;
                for (int i = 0; i < NS; ++i) s += B[i * 2 + a] * lam[i];
                const double d = -s - lu[a];
                du[a * TT + t] = d;
                sq += d * d;
            }
            descent += sq; /* :118 */
            memcpy(lam, lamn, sizeof(lam));
        }
        for (int a = 0; a < NI; ++a) du[a * TT + TT - 1] = 0.0;
        if (kk == 0 && du_first) memcpy(du_first, du, nu * sizeof(double));
        int ntried = 0, acc = 0;
        const double s = orc_armijo(prm, quant_f32, TT, x0, uu, du, Q, R, QT, xx_ref, uu_ref, JJ, -descent,
                                    stepsize_0, cc, beta, armijo_maxiters, NULL, &ntried, &acc); /* :125, repaired */
        orc_rollout(prm, quant_f32, TT, x0, uu, du, s, NULL, NULL, NULL, NULL, NULL, X[(kk + 1) % 3], U[(kk + 1) % 3]); /* :131 */
        if (hist_J) hist_J[kk] = JJ;
        if (hist_descent) hist_descent[kk] = descent;
        if (hist_step) hist_step[kk] = s;
        if (hist_ncand) hist_ncand[kk] = ntried;
        if (descent <= term_cond) { stop_at = kk; ++kk; break; } /* :157-161 */
    }
    const int executed = kk;
    if (iters) *iters = executed;
    if (xx_last) memcpy(xx_last, X[executed % 3], nx * sizeof(double));
    if (uu_last) memcpy(uu_last, U[executed % 3], nu * sizeof(double));
    if (stop_at == 0) { /* :163 with max_iters = 0: slot -1, never written */
        memset(xx_star, 0, nx * sizeof(double)); memset(uu_star, 0, nu * sizeof(double));
    } else if (stop_at > 0) {
        memcpy(xx_star, X[(stop_at - 1) % 3], nx * sizeof(double));
        memcpy(uu_star, U[(stop_at - 1) % 3], nu * sizeof(double));
    } else {
        memcpy(xx_star, X[executed % 3], nx * sizeof(double));
        memcpy(uu_star, U[executed % 3], nu * sizeof(double));
    }
    for (int a = 0; a < NI; ++a) uu_star[a * TT + TT - 1] = uu_star[a * TT + TT - 2]; /* :165 */
    for (int s = 0; s < 3; ++s) { free(X[s]); free(U[s]); }
    free(du);
    return 0;
}

int orc_gradient_batch(int N, int n_threads, const double *prm, int quant_f32, int TT, const double *Q, const double *R, const double *QT,
                       const double *xx_ref, long ref_stride_x, const double *uu_ref, long ref_stride_u,
                       const double *xx_init, const double *uu_init,
                       int max_iters, double stepsize_0, double cc, double beta, int armijo_maxiters, double term_cond,
                       double *hist_J, double *hist_descent, double *hist_step, int *hist_ncand, int *iters,
                       double *xx_star, double *uu_star)
{
    int rc = 0;
    const size_t nx = (size_t)NS * TT, nu = (size_t)NI * TT;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int n = 0; n < N; ++n) {
        int r = orc_gradient(prm, quant_f32, TT, Q, R, QT, xx_ref + (size_t)n * ref_stride_x, uu_ref + (size_t)n * ref_stride_u,
                             xx_init + n * nx, uu_init + n * nu, max_iters, stepsize_0, cc, beta, armijo_maxiters, term_cond,
                             hist_J ? hist_J + (size_t)n * max_iters : NULL, hist_descent ? hist_descent + (size_t)n * max_iters : NULL,
                             hist_step ? hist_step + (size_t)n * max_iters : NULL, hist_ncand ? hist_ncand + (size_t)n * max_iters : NULL,
                             iters ? iters + n : NULL, xx_star + n * nx, uu_star + n * nu, NULL, NULL, NULL);
        if (r) rc = r;
    }
    return rc;
}

/* ------------------------------------------------------------------------------------------------ */
/* lqr_tracking, lqr_tracking.py:245-283                                                             */
/* ------------------------------------------------------------------------------------------------ */
/* Nominal (xx_opt, uu_opt) of shape (6,TT)/(2,TT); N perturbations delta[n][6]; outputs xx_reg[n] (6,TT),
 * uu_reg[n] (2,TT).  K_out (optional) receives the shared gains K[t][2][6]. */
int orc_lqr_tracking(const double *prm, int quant_f32, int TT, const double *Q, const double *R, const double *QT,
                     const double *xx_opt, const double *uu_opt, int N, const double *delta, int n_threads,
                     double *xx_reg, double *uu_reg, double *K_out)
{
    double *A = (double *)calloc((size_t)TT * 36, sizeof(double)), *B = (double *)calloc((size_t)TT * 12, sizeof(double));
    double *Qt = (double *)calloc((size_t)TT * 36, sizeof(double)), *Rt = (double *)calloc((size_t)TT * 4, sizeof(double));
    double *St = (double *)calloc((size_t)TT * 12, sizeof(double)), *K = (double *)calloc((size_t)TT * 12, sizeof(double));
    double *xo = (double *)calloc((size_t)TT * 6, sizeof(double)), *uo = (double *)calloc((size_t)TT * 2, sizeof(double));
    if (!A || !B || !Qt || !Rt || !St || !K || !xo || !uo) return -1;
    for (int t = 0; t < TT; ++t) { /* :268-273 (linearises at all TT points) */
        double x[NS], u[NI];
        for (int i = 0; i < NS; ++i) x[i] = xx_opt[i * TT + t];
        for (int a = 0; a < NI; ++a) u[a] = uu_opt[a * TT + t];
        orc_step(prm, x, u, NULL, quant_f32, NULL, A + (size_t)t * 36, B + (size_t)t * 12, NULL, NULL);
        memcpy(Qt + (size_t)t * 36, Q, 36 * sizeof(double));
        memcpy(Rt + (size_t)t * 4, R, 4 * sizeof(double));
    }
    /* gains do not depend on x0 (:276 passes delta_xx, only the unused forward pass sees it) */
    double d0[NS] = {0, 0, 0, 0, 0, 0};
    orc_ltv_lqr(TT, A, B, Qt, Rt, St, QT, delta ? delta : d0, NULL, NULL, NULL, K, NULL, xo, uo, NULL);
    if (K_out) memcpy(K_out, K, (size_t)TT * 12 * sizeof(double));
    const size_t nx = (size_t)NS * TT, nu = (size_t)NI * TT;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int n = 0; n < N; ++n) { /* :279-281 */
        double x[NS], xn[NS], u[NI];
        double *xr = xx_reg + n * nx, *ur = uu_reg + n * nu;
        for (int i = 0; i < NS; ++i) x[i] = xx_opt[i * TT] + delta[n * NS + i];
        for (int t = 0; t < TT - 1; ++t) {
            const double *Kt = K + (size_t)t * 12;
            for (int a = 0; a < NI; ++a) {
                double s = 0.0;
                for (int j = 0; j < NS; ++j) s += Kt[a * 6 + j] * (x[j] - xx_opt[j * TT + t]);
                u[a] = uu_opt[a * TT + t] + s;
            }
            for (int i = 0; i < NS; ++i) xr[i * TT + t] = x[i];
            for (int a = 0; a < NI; ++a) ur[a * TT + t] = u[a];
            orc_step(prm, x, u, NULL, quant_f32, xn, NULL, NULL, NULL, NULL);
            memcpy(x, xn, sizeof(x));
        }
        for (int i = 0; i < NS; ++i) xr[i * TT + TT - 1] = x[i];
        for (int a = 0; a < NI; ++a) ur[a * TT + TT - 1] = 0.0;
    }
    free(A); free(B); free(Qt); free(Rt); free(St); free(K); free(xo); free(uo);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* get_initial_trajectory, aircraft_simplified.py:126-148, in float64 arithmetic                     */
/* ------------------------------------------------------------------------------------------------ */
/* Under NumPy >= 2 the reference runs this loop partly in float32 (the float32 xxp is fed back into step,
 * :145), so this restatement agrees with it only to ~1e-5; see DESIGN.md "initial guess". */
void orc_initial_trajectory(const double *prm, int quant_f32, int TT, const double *xx_ref, double kp, double kt,
                            double *xx, double *uu)
{
    double x[NS], xn[NS], u[NI];
    for (int i = 0; i < NS; ++i) x[i] = xx_ref[i * TT];
    for (int t = 0; t < TT - 1; ++t) {
        u[0] = kp * ((x[0] - xx_ref[0 * TT + t + 1]) + (x[1] - xx_ref[1 * TT + t + 1]));
        u[1] = kt * ((x[3] - xx_ref[3 * TT + t + 1]) + (x[5] - xx_ref[5 * TT + t + 1]));
        for (int i = 0; i < NS; ++i) xx[i * TT + t] = x[i];
        for (int a = 0; a < NI; ++a) uu[a * TT + t] = u[a];
        orc_step(prm, x, u, NULL, quant_f32, xn, NULL, NULL, NULL, NULL);
        memcpy(x, xn, sizeof(x));
    }
    for (int i = 0; i < NS; ++i) xx[i * TT + TT - 1] = x[i];
    for (int a = 0; a < NI; ++a) uu[a * TT + TT - 1] = 0.0;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
