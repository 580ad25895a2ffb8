// acoc_math.cuh -- per-time-step math of the aircraft OCP hot path, written once for the sm_100a kernels.
//
// Everything here is a __host__ __device__ inline function of scalars held in registers; the kernels in
// acoc_kernels.cuh call them once per (instance, time step).  The host flavour exists only so that
// tests/host_emul can replay the kernels' arithmetic on a CPU (no GPU in the build container); the product
// library never runs it.
//
// What is restated (reference = MohamedAtwan/AirCraftOptimalControl):
//   next_state()   aircraft_simplified.py:295-310 (+ dragForce :228, liftForce :253), operation order and
//                  float32 rounding of the next state (:300) preserved so the rollouts match bit for bit
//   linearize()    aircraft_simplified.py:316-325 -- only the 10 non-constant entries of A = fx.T and the
//                  2 non-constant entries of B = fu.T
//   hess_contract() aircraft_simplified.py:339-388, 397-404 -- sum_k lambda_k * d2f_k, which lives on the
//                  {V,theta,gamma} block (6 numbers) and the thrust row of d2f/dudx (3 numbers)
//   stage/terminal cost: aircraft_simplified.py:61-64, :92-94
//
// Floating-point contract: the translation unit is compiled with -fmad=false (nvcc) / -ffp-contract=off
// (gcc), so a*b+c is two roundings unless written as fma_().  next_state() and the cost sums use plain
// operators in the reference's order; the linearisation / Riccati code uses fma_() explicitly.
//
// Everything is a template on the arithmetic type F.  F = double is the parity path (all statements above are
// about it); F = float is the optional FP32 mode of the library (no bit-level contract, tolerance in DESIGN.md).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define ACOC_HD __host__ __device__ __forceinline__
#else
#define ACOC_HD inline
#endif

namespace acoc {

constexpr int NS = 6;  // state  [X, Z, V, theta, q, gamma]   aircraft_simplified.py:116, :264
constexpr int NI = 2;  // input  [T, M]                       aircraft_simplified.py:117

// Model constants (aircraft_simplified.py:108-118) plus products that the reference forms from python
// scalars before touching the state; they are computed once on the host with the same operations.
template <typename F>
struct ModelT {
    F cd0, cda, cla, m, g, S, rho, J, dt;
    F half_rho;  // 0.5*rho          (:228, :253)
    F mg;        // m*g              (:306, :310)
    F dt_m;      // dt/m             (:306)
    F k;         // S*rho
    F cdk, clk;  // cda*k, cla*k
    F b41;       // dt/J             (:324)
    F gm;        // g*m              (:318, :321)
    F rJ;        // RN(1/J): u/J is formed as a correctly rounded division from it (div_const)
};
using Model = ModelT<double>;

inline Model make_model(const double* p)
{
    Model M;
    M.cd0 = p[0]; M.cda = p[1]; M.cla = p[2]; M.m = p[3]; M.g = p[4]; M.S = p[5]; M.rho = p[6]; M.J = p[7]; M.dt = p[8];
    M.half_rho = 0.5 * M.rho;
    M.mg = M.m * M.g;
    M.dt_m = M.dt / M.m;
    M.k = M.S * M.rho;
    M.cdk = M.cda * M.k;
    M.clk = M.cla * M.k;
    M.b41 = M.dt / M.J;
    M.gm = M.g * M.m;
    M.rJ = 1.0 / M.J;
    return M;
}

// the FP32 mode rounds the float64 constants once
template <typename F>
inline ModelT<F> model_as(const Model& m)
{
    ModelT<F> o;
    o.cd0 = (F)m.cd0; o.cda = (F)m.cda; o.cla = (F)m.cla; o.m = (F)m.m; o.g = (F)m.g; o.S = (F)m.S; o.rho = (F)m.rho; o.J = (F)m.J; o.dt = (F)m.dt;
    o.half_rho = (F)m.half_rho; o.mg = (F)m.mg; o.dt_m = (F)m.dt_m; o.k = (F)m.k; o.cdk = (F)m.cdk; o.clk = (F)m.clk; o.b41 = (F)m.b41;
    o.gm = (F)m.gm; o.rJ = (F)m.rJ;
    return o;
}

// Quadratic weights.  diag != 0 promises that Q, R and QT are diagonal (every shipped configuration,
// main_newton_method.py:52-63); the dense path keeps the API honest (aircraft_simplified.py:61 takes any matrix).
template <typename F>
struct WeightsT {
    F Q[36], R[4], QT[36];
    int diag;
};
using Weights = WeightsT<double>;

template <typename F>
inline WeightsT<F> weights_as(const Weights& w)
{
    WeightsT<F> o;
    for (int e = 0; e < 36; ++e) { o.Q[e] = (F)w.Q[e]; o.QT[e] = (F)w.QT[e]; }
    for (int e = 0; e < 4; ++e) o.R[e] = (F)w.R[e];
    o.diag = w.diag;
    return o;
}

ACOC_HD double fma_(double a, double b, double c) { return fma(a, b, c); }
ACOC_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }
ACOC_HD double sqrt_(double a) { return sqrt(a); }
ACOC_HD float sqrt_(float a) { return sqrtf(a); }
ACOC_HD double abs_(double a) { return fabs(a); }
ACOC_HD float abs_(float a) { return fabsf(a); }

// ---- sin/cos -------------------------------------------------------------------------------------------
// Own implementation instead of CUDA's sincos(): same algorithm class (3-term Cody-Waite reduction by pi/2 with
// FMAs, degree-13/12 minimax polynomials (the classic fdlibm k_sin/k_cos coefficient sets), quadrant fix-up), but
// the coefficients are read as constant-bank operands, which removes ~100 UMOV/IMAD.MOV/spill instructions per
// time step that nvcc emits to materialise the library routine's 64-bit immediates.  Error <= 1 ulp against glibc
// on |x| < 1e5 (2 ulp in isolated cases); the float32 rounding of the next state absorbs it (tests: rollouts stay
// bit-identical to the oracle).  The same code runs in the host replay, so replay and GPU agree bit for bit.
#define ACOC_MATH_CONSTANTS                                                                                      \
    {0x1.45f306dc9c883p-1, /* 0: 2/pi */                                                                          \
     0x1.921fb54442d18p+0, 0x1.1a62633145c00p-54, 0x1.b839a252049c0p-104, /* 1-3: pi/2 in three parts */          \
     -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, /* 4-9: S1..S6 */      \
     2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,                         \
     4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, /* 10-15: C1..C6 */     \
     -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11}
#if defined(__CUDACC__)
__constant__ double c_math[16] = ACOC_MATH_CONSTANTS;
#endif
static const double h_math[16] = ACOC_MATH_CONSTANTS;
#if defined(__CUDA_ARCH__)
#define ACOC_K(i) c_math[i]
#else
#define ACOC_K(i) h_math[i]
#endif

// huge, Inf or NaN arguments: library routine, kept out of line so that its code (Payne-Hanek reduction, 64-bit
// immediates) does not sit inside the time loops; never taken on a sane trajectory.  Returned by value so that the
// caller's sin/cos stay in registers on the fast path.
struct SinCos { double s, c; };
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
SinCos sincos_slow_(double x)
{
    SinCos r;
#if defined(__CUDA_ARCH__)
    sincos(x, &r.s, &r.c);
#else
    r.s = sin(x); r.c = cos(x);
#endif
    return r;
}

ACOC_HD void sincos_(double x, double& s, double& c)
{
    if (!(fabs(x) < 1.0e5)) { const SinCos r = sincos_slow_(x); s = r.s; c = r.c; return; }
    const double jd = rint(x * ACOC_K(0));
    const int j = (int)jd;
    double r = fma_(-jd, ACOC_K(1), x);
    r = fma_(-jd, ACOC_K(2), r);
    r = fma_(-jd, ACOC_K(3), r);
    const double z = r * r;
    double ps = fma_(z, ACOC_K(9), ACOC_K(8));
    ps = fma_(z, ps, ACOC_K(7)); ps = fma_(z, ps, ACOC_K(6)); ps = fma_(z, ps, ACOC_K(5)); ps = fma_(z, ps, ACOC_K(4));
    const double sr = fma_(z * ps, r, r);
    double pc = fma_(z, ACOC_K(15), ACOC_K(14));
    pc = fma_(z, pc, ACOC_K(13)); pc = fma_(z, pc, ACOC_K(12)); pc = fma_(z, pc, ACOC_K(11)); pc = fma_(z, pc, ACOC_K(10));
    const double cr = fma_(z, fma_(z, pc, -0.5), 1.0);
    double a = (j & 1) ? cr : sr, b = (j & 1) ? sr : cr;
    if (j & 2) a = -a;
    if ((j + 1) & 2) b = -b;
    s = a; c = b;
}

// FP32 mode: same structure in float (Cody-Waite by pi/2 in three parts with FMAs, the cephes sinf/cosf minimax
// polynomials on [-pi/4, pi/4]); ~1e-7 absolute error.  Huge/NaN arguments: library routine.
ACOC_HD void sincos_(float x, float& s, float& c)
{
    if (!(fabsf(x) < 1.0e4f)) {
#if defined(__CUDA_ARCH__)
        sincosf(x, &s, &c);
#else
        s = sinf(x); c = cosf(x);
#endif
        return;
    }
    const float jf = rintf(x * 0x1.45f306p-1f);
    const int j = (int)jf;
    float r = fma_(-jf, 0x1.921fb6p+0f, x);
    r = fma_(-jf, -0x1.777a5cp-25f, r);
    r = fma_(-jf, -0x1.ee59dap-50f, r);
    const float z = r * r;
    const float ps = fma_(z, fma_(z, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f);
    const float sr = fma_(z * ps, r, r);
    const float pc = fma_(z, fma_(z, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f);
    const float cr = fma_(z, fma_(z, pc, -0.5f), 1.0f);
    float a = (j & 1) ? cr : sr, b = (j & 1) ? sr : cr;
    if (j & 2) a = -a;
    if ((j + 1) & 2) b = -b;
    s = a; c = b;
}

// a / d for a divisor whose correctly rounded reciprocal rd = RN(1/d) is known: q = RN(a*rd), r = a - q*d (exact, FMA),
// RN(q + r*rd) is the correctly rounded quotient (Markstein) -- three FMA-pipe operations instead of a division routine.
template <typename F>
ACOC_HD F div_const(F a, F d, F rd)
{
    const F q = a * rd;
    const F r = fma_(-q, d, a);
    return fma_(r, rd, q);
}

// aircraft_simplified.py:300 -- the reference stores the next state in a float32 array.  (FP32 mode: identity.)
// (The round trip is two F2F instructions at ~7.5 cycles per warp each on B200 -- tools/microbench/f2f.cu -- but they run beside the
// FP64 pipe, which is the busy one in the rollouts: replacing them by the two-addition rounding trick (v + 1.5*2^(e+29)) - 1.5*2^(e+29),
// bit-identical on 1.4e9 test values, made the candidate kernel 9 % SLOWER.  Measured in round 2, not kept.)
template <bool Q32>
ACOC_HD double quant_(double v) { return Q32 ? (double)(float)v : v; }
template <bool Q32>
ACOC_HD float quant_(float v) { return v; }

template <typename F>
struct Trig { F sg, cg, sa, ca, alpha; };

template <typename F>
ACOC_HD Trig<F> make_trig(const F* x)
{
    Trig<F> t;
    t.alpha = x[3] - x[5];  // :295
    sincos_(x[5], t.sg, t.cg);
    sincos_(t.alpha, t.sa, t.ca);
    return t;
}

// One forward-Euler step, aircraft_simplified.py:303-310, in the reference's evaluation order.
template <bool Q32, typename F>
ACOC_HD void next_state(const ModelT<F>& M, const F* x, const F* u, const Trig<F>& t, F* xn)
{
    const F V = x[2];
    const F V2 = V * V, a2 = t.alpha * t.alpha;
    const F hv = M.half_rho * V2 * M.S;                       // 0.5*rho*V**2*S
    const F D = hv * (M.cd0 + M.cda * a2);                    // :228
    const F L = hv * M.cla * t.alpha;                         // :253
    const F dtV = M.dt * V;
    xn[0] = quant_<Q32>(x[0] + dtV * t.cg);
    xn[1] = quant_<Q32>(x[1] - dtV * t.sg);
    xn[2] = quant_<Q32>(V + M.dt_m * (-D - M.mg * t.sg + u[0] * t.ca));
    xn[3] = quant_<Q32>(x[3] + M.dt * x[4]);
    xn[4] = quant_<Q32>(x[4] + M.dt * div_const(u[1], M.J, M.rJ));
    xn[5] = quant_<Q32>(x[5] + (M.dt / (M.m * V)) * (L - M.mg * t.cg + u[0] * t.sa));
}

// Non-constant entries of A = df/dx and B = df/du.  Constant ones: A00=A11=A33=A44=1, A34=dt, B41=dt/J.
template <typename F>
struct Lin {
    F a02, a05, a12, a15, a22, a23, a25, a52, a53, a55, b20, b50;
    // by-products reused by hess_contract()
    F iV, drag_c, lift, liftT, tsa, tca;
};

template <typename F>
ACOC_HD Lin<F> linearize(const ModelT<F>& M, const F* x, const F* u, const Trig<F>& t)
{
    Lin<F> l;
    const F V = x[2], T = u[0], al = t.alpha;
    const F iV = F(1.0) / V;
    const F V2 = V * V;
    const F dc = fma_(M.cda * al, al, M.cd0);          // Cd0 + Cda*alpha^2
    const F tsa = T * t.sa, tca = T * t.ca;
    const F cav = M.cdk * al * V2;                     // Cda*k*alpha*V^2
    const F lift = F(0.5) * M.clk * V2;                // 0.5*Cla*k*V^2
    const F dtV = M.dt * V;
    l.a02 = M.dt * t.cg;
    l.a05 = -dtV * t.sg;
    l.a12 = -M.dt * t.sg;
    l.a15 = -dtV * t.cg;
    l.a22 = fma_(-M.dt_m * M.k * V, dc, F(1.0));
    l.a23 = -M.dt_m * (cav + tsa);
    l.a25 = M.dt_m * (cav + tsa - M.gm * t.cg);
    const F w = fma_(lift, al, tsa) - M.gm * t.cg;     // lift*alpha + T sin(alpha) - g m cos(gamma)
    l.a52 = M.dt_m * fma_(-w * iV, iV, M.clk * al);
    l.a53 = M.dt_m * (lift + tca) * iV;
    l.a55 = fma_(-M.dt_m * (lift + tca - M.gm * t.sg), iV, F(1.0));
    l.b20 = M.dt_m * t.ca;
    l.b50 = M.dt_m * t.sa * iV;
    l.iV = iV; l.drag_c = dc; l.lift = lift; l.liftT = lift + tca; l.tsa = tsa; l.tca = tca;
    return l;
}

// lambda-contracted second-order terms (optcon.py:437 with aircraft_simplified.py:384-388):
// symmetric block on {V,theta,gamma} = indices {2,3,5} and the thrust row of fux.
template <typename F>
struct Hess { F h22, h23, h25, h33, h35, h55, s2, s3, s5; };

template <typename F>
ACOC_HD Hess<F> hess_contract(const ModelT<F>& M, const F* x, const F* u, const Trig<F>& t, const Lin<F>& l, const F* lam)
{
    Hess<F> h;
    const F V = x[2], al = t.alpha, iV = l.iV, iV2 = iV * iV;
    const F l0 = lam[0], l1 = lam[1], l2 = lam[2], l5 = lam[5];
    const F cv2 = M.cdk * V * V;
    // d2 f_V
    const F v22 = -M.dt_m * M.k * l.drag_c;
    const F v23 = -M.dt_m * M.cdk * V * (F(2.0) * al);
    const F v33 = -M.dt_m * (cv2 + l.tca);
    const F v55 = -M.dt_m * (cv2 + l.tca - M.gm * t.sg);
    // d2 f_gamma
    const F w = fma_(l.lift, al, l.tsa) - M.gm * t.cg;
    const F clkdtm = M.clk * M.dt_m;
    const F g22 = fma_(F(2.0) * M.dt_m * w * iV2, iV, -clkdtm * al * iV);
    const F g23 = fma_(-M.dt_m * l.liftT, iV2, clkdtm);
    const F g25 = fma_(M.dt_m * (l.liftT - M.gm * t.sg), iV2, -clkdtm);
    const F g33 = -M.dt_m * l.tsa * iV;
    const F g55 = -M.dt_m * (l.tsa - M.gm * t.cg) * iV;
    const F dtV = M.dt * V;
    h.h22 = fma_(l2, v22, l5 * g22);
    h.h23 = fma_(l2, v23, l5 * g23);
    h.h25 = fma_(l0, -M.dt * t.sg, fma_(l1, -M.dt * t.cg, fma_(l2, -v23, l5 * g25)));
    h.h33 = fma_(l2, v33, l5 * g33);
    h.h35 = -h.h33;
    h.h55 = fma_(l0, -dtV * t.cg, fma_(l1, dtV * t.sg, fma_(l2, v55, l5 * g55)));
    h.s2 = l5 * (-M.dt_m * t.sa * iV2);
    h.s3 = fma_(l2, -M.dt_m * t.sa, l5 * (M.dt_m * t.ca * iV));
    h.s5 = -h.s3;
    return h;
}

// hess_contract() cut at lambda: hess_pre() forms everything that does not depend on the costate (one warp of k_backward_cols runs it
// ahead of time, several time steps in parallel), hess_post() the fused multiply-adds with lambda_{t+1} (the warp that carries the
// costate recurrence).  Same expressions as hess_contract(), so hess_post(hess_pre(..), lam) is bit-identical to it (host replay test).
template <typename F>
struct HessPre { F v22, v23, v33, v55, g22, g23, g25, g33, g55, c0, c1, c2, c3, e2, e3a, e3b; };

template <typename F>
ACOC_HD HessPre<F> hess_pre(const ModelT<F>& M, const F* x, const Trig<F>& t, const Lin<F>& l)
{
    HessPre<F> a;
    const F V = x[2], al = t.alpha, iV = l.iV, iV2 = iV * iV;
    const F cv2 = M.cdk * V * V;
    a.v22 = -M.dt_m * M.k * l.drag_c;
    a.v23 = -M.dt_m * M.cdk * V * (F(2.0) * al);
    a.v33 = -M.dt_m * (cv2 + l.tca);
    a.v55 = -M.dt_m * (cv2 + l.tca - M.gm * t.sg);
    const F w = fma_(l.lift, al, l.tsa) - M.gm * t.cg;
    const F clkdtm = M.clk * M.dt_m;
    a.g22 = fma_(F(2.0) * M.dt_m * w * iV2, iV, -clkdtm * al * iV);
    a.g23 = fma_(-M.dt_m * l.liftT, iV2, clkdtm);
    a.g25 = fma_(M.dt_m * (l.liftT - M.gm * t.sg), iV2, -clkdtm);
    a.g33 = -M.dt_m * l.tsa * iV;
    a.g55 = -M.dt_m * (l.tsa - M.gm * t.cg) * iV;
    const F dtV = M.dt * V;
    a.c0 = -M.dt * t.sg; a.c1 = -M.dt * t.cg; a.c2 = -dtV * t.cg; a.c3 = dtV * t.sg;
    a.e2 = -M.dt_m * t.sa * iV2; a.e3a = -M.dt_m * t.sa; a.e3b = M.dt_m * t.ca * iV;
    return a;
}

template <typename F>
ACOC_HD Hess<F> hess_post(const HessPre<F>& a, const F* lam)
{
    Hess<F> h;
    const F l0 = lam[0], l1 = lam[1], l2 = lam[2], l5 = lam[5];
    h.h22 = fma_(l2, a.v22, l5 * a.g22);
    h.h23 = fma_(l2, a.v23, l5 * a.g23);
    h.h25 = fma_(l0, a.c0, fma_(l1, a.c1, fma_(l2, -a.v23, l5 * a.g25)));
    h.h33 = fma_(l2, a.v33, l5 * a.g33);
    h.h35 = -h.h33;
    h.h55 = fma_(l0, a.c2, fma_(l1, a.c3, fma_(l2, a.v55, l5 * a.g55)));
    h.s2 = l5 * a.e2;
    h.s3 = fma_(l2, a.e3a, l5 * a.e3b);
    h.s5 = -h.s3;
    return h;
}

// DG: what the caller knows about the weights at compile time -- 1: diagonal, 0: dense, -1: look at W.diag at run time.  The hot
// kernels are instantiated for DG = 1 (every shipped configuration) and DG = 0, which removes the dense arm, its branches and the
// additions of structural zeros from their time loops; same arithmetic on the entries that exist.
template <int DG, typename F>
ACOC_HD bool weights_diag(const WeightsT<F>& W) { return DG < 0 ? W.diag != 0 : DG != 0; }

// Stage cost (aircraft_simplified.py:61): 0.5*dx'(Q dx) + 0.5*du'(R du), sums in ascending index order.
template <int DG = -1, typename F>
ACOC_HD F stage_cost(const WeightsT<F>& W, const F* dx, const F* du)
{
    F sx = F(0.0), su = F(0.0);
    if (weights_diag<DG>(W)) {
#pragma unroll
        for (int i = 0; i < NS; ++i) sx += dx[i] * (W.Q[i * 7] * dx[i]);
#pragma unroll
        for (int i = 0; i < NI; ++i) su += du[i] * (W.R[i * 3] * du[i]);
    } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            F a = F(0.0);
#pragma unroll
            for (int j = 0; j < NS; ++j) a += W.Q[i * 6 + j] * dx[j];
            sx += dx[i] * a;
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            F a = F(0.0);
#pragma unroll
            for (int j = 0; j < NI; ++j) a += W.R[i * 2 + j] * du[j];
            su += du[i] * a;
        }
    }
    return F(0.5) * sx + F(0.5) * su;
}

// Terminal cost (aircraft_simplified.py:92): ((0.5*dx') QT) dx.
template <int DG = -1, typename F>
ACOC_HD F term_cost(const WeightsT<F>& W, const F* dx)
{
    F s = F(0.0);
    if (weights_diag<DG>(W)) {
#pragma unroll
        for (int j = 0; j < NS; ++j) s += ((F(0.5) * dx[j]) * W.QT[j * 7]) * dx[j];
    } else {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            F a = F(0.0);
#pragma unroll
            for (int i = 0; i < NS; ++i) a += (F(0.5) * dx[i]) * W.QT[i * 6 + j];
            s += a * dx[j];
        }
    }
    return s;
}

// v = W (dx) for a 6x6 weight (gradient lx = Q dx, aircraft_simplified.py:63 / :94)
template <typename F>
ACOC_HD void wmul6(const F* Wm, int diag, const F* dx, F* out)
{
    if (diag) {
#pragma unroll
        for (int i = 0; i < NS; ++i) out[i] = Wm[i * 7] * dx[i];
    } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            F a = F(0.0);
#pragma unroll
            for (int j = 0; j < NS; ++j) a = fma_(Wm[i * 6 + j], dx[j], a);
            out[i] = a;
        }
    }
}

template <typename F>
ACOC_HD void wmul2(const F* Wm, int diag, const F* du, F* out)
{
    if (diag) { out[0] = Wm[0] * du[0]; out[1] = Wm[3] * du[1]; }
    else { out[0] = fma_(Wm[1], du[1], Wm[0] * du[0]); out[1] = fma_(Wm[3], du[1], Wm[2] * du[0]); }
}

}  // namespace acoc
