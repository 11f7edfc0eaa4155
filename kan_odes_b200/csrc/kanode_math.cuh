// kanode_math.cuh — scalar device math shared by every kernel of the KAN-ODE hot path (sm_100a).
//
// Reference semantics (file:line into the reference tree; [EXT] = un-vendored Julia package):
//   normalizers / swish       [EXT NNlib 0.9.24] tanh_fast, softsign, sigmoid_fast, swish and their d/dx rules
//   rbf / rswaf / iqf         Lotka-Volterra/src/utils.jl:8-62 (forward :13,:33,:54; reverse rules :18,:40,:59)
//   fastpower                 [EXT FastPower 1.1.0]  exp2(Float32(y) * fastlog2(Float32(x)))
//   Tsit5 tableau, b(theta)   [EXT OrdinaryDiffEqTsit5 1.1.0]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kanode {

enum { NORM_TANH = 0, NORM_SOFTSIGN = 1, NORM_SIGMOID = 2 };
enum { BASIS_RBF = 0, BASIS_RSWAF = 1, BASIS_IQF = 2 };
enum { RET_SUCCESS = 0, RET_MAXITERS = 1, RET_DTMIN = 2, RET_UNSTABLE = 3, RET_OVERFLOW = 4 };

// ---- Tsit5 constants (stage s uses row s; row 6 == b) ------------------------------------------------
// static __constant__ with initialisers: every translation unit carries its own copy (no -rdc needed).
#define KANODE_TSIT5_A                                                                                                   \
    {{0, 0, 0, 0, 0, 0, 0, 0},                                                                                           \
     {0.161, 0, 0, 0, 0, 0, 0, 0},                                                                                       \
     {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0, 0, 0},                                                       \
     {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0, 0, 0},                                        \
     {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0, 0, 0},                     \
     {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0, 0, 0},      \
     {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774, 0, 0}}
#define KANODE_TSIT5_C {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0, 0.0}
#define KANODE_TSIT5_BT                                                                                                  \
    {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629, 0.5823571654525552,     \
     -0.45808210592918697, 0.015151515151515152, 0.0}
struct Tsit5Tab { double a[7][8]; double c[8]; double bt[8]; };
struct Tsit5TabF { float a[7][8]; float c[8]; float bt[8]; };
static __constant__ Tsit5Tab c_tab_d = {KANODE_TSIT5_A, KANODE_TSIT5_C, KANODE_TSIT5_BT};
static __constant__ Tsit5TabF c_tab_f = {KANODE_TSIT5_A, KANODE_TSIT5_C, KANODE_TSIT5_BT};

template <class T> struct Tab;
template <> struct Tab<double> {
    static __device__ __forceinline__ double a(int s, int j) { return c_tab_d.a[s][j]; }
    static __device__ __forceinline__ double bt(int j) { return c_tab_d.bt[j]; }
    static __device__ __forceinline__ double b(int j) { return c_tab_d.a[6][j]; }
};
template <> struct Tab<float> {
    static __device__ __forceinline__ float a(int s, int j) { return c_tab_f.a[s][j]; }
    static __device__ __forceinline__ float bt(int j) { return c_tab_f.bt[j]; }
    static __device__ __forceinline__ float b(int j) { return c_tab_f.a[6][j]; }
};
__device__ __forceinline__ double tab_c(int s) { return c_tab_d.c[s]; }

// dense-output weights b_i(theta)
template <class T> __device__ __forceinline__ void interp_weights(T th, T (&b)[7]) {
    const T th2 = th * th;
    b[0] = th * (T(1.0) + th * (T(-2.763706197274826) + th * (T(2.9132554618219126) + th * T(-1.0530884977290216))));
    b[1] = th2 * (T(0.13169999999999998) + th * (T(-0.2234) + th * T(0.1017)));
    b[2] = th2 * (T(3.9302962368947516) + th * (T(-5.941033872131505) + th * T(2.490627285651253)));
    b[3] = th2 * (T(-12.411077166933676) + th * (T(30.33818863028232) + th * T(-16.548102889244902)));
    b[4] = th2 * (T(37.50931341651104) + th * (T(-88.1789048947664) + th * T(47.37952196281928)));
    b[5] = th2 * (T(-27.896526289197286) + th * (T(65.09189467479366) + th * T(-34.87065786149661)));
    b[6] = th2 * (T(1.5) + th * (T(-4.0) + th * T(2.5)));
}

// ---- elementary functions -------------------------------------------------------------------------
__device__ __forceinline__ float kexp(float x) { return __expf(x); }      // ex2.approx(x*log2e), 2 ulp
__device__ __forceinline__ double kexp(double x) { return exp(x); }
__device__ __forceinline__ float kdiv(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double kdiv(double a, double b) { return a / b; }
__device__ __forceinline__ float kabs(float x) { return fabsf(x); }
__device__ __forceinline__ double kabs(double x) { return fabs(x); }
__device__ __forceinline__ float kmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double kmax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float ksqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double ksqrt(double a) { return sqrt(a); }

// ex2.approx: one MUFU, 2 ulp
__device__ __forceinline__ float kex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Gaussian RBF on a pre-scaled argument: exp(-a^2) with t = a*KRbfScale<T>.  fp32 folds sqrt(log2 e) into the
// argument so the basis is FFMA + FMUL + MUFU.EX2; fp64 uses libdevice exp (scale 1).
template <class T> struct KRbfScale;
template <> struct KRbfScale<float> { static constexpr double value = 1.2011224087864498; };   // sqrt(log2(e))
template <> struct KRbfScale<double> { static constexpr double value = 1.0; };
__device__ __forceinline__ float krbf_scaled(float t) { return kex2(-t * t); }
__device__ __forceinline__ double krbf_scaled(double t) { return exp(-t * t); }

// tanh.  The reference's tanh_fast(::Float64) is exp-based ((e-1)/(e+1), e = exp(2x)); fp32 here uses the same form
// with two MUFU ops, 1 - 2/(1+e): absolute error <= ~2e-7 (what matters: the result feeds (xn - grid)/h), never NaN
// (e=inf -> 1, e=0 -> -1).  NNlib's Float32 rational tanh_fast (3e-7) costs 3x the instructions.  fp64: libdevice.
__device__ __forceinline__ float krcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ktanh(float x) {
    const float e = kex2(x * 2.8853900817779268f);                     // exp(2x)
    return fmaf(-2.0f, krcp(1.0f + e), 1.0f);
}
__device__ __forceinline__ double ktanh(double x) { return tanh(x); }

template <class T> __device__ __forceinline__ T ksigmoid(T x) {
    const T t = kexp(-kabs(x));
    const T r = kdiv(T(1), T(1) + t);
    return x >= T(0) ? r : t * r;
}
// fp32: exp(-|x|) as one ex2.approx.ftz (no denormal fix-up code) and one rcp.approx
template <> __device__ __forceinline__ float ksigmoid<float>(float x) {
    const float t = kex2(-1.4426950408889634f * fabsf(x));
    const float r = krcp(1.0f + t);
    return x >= 0.0f ? r : t * r;
}

template <int NORM, class T> __device__ __forceinline__ T normalize(T x) {
    if (NORM == NORM_TANH) return ktanh(x);
    if (NORM == NORM_SOFTSIGN) return kdiv(x, T(1) + kabs(x));
    return ksigmoid(x);
}
template <int NORM, class T> __device__ __forceinline__ T normalize_deriv(T omega) {
    if (NORM == NORM_TANH) return T(1) - omega * omega;
    if (NORM == NORM_SOFTSIGN) { const T a = T(1) - kabs(omega); return a * a; }
    return omega * (T(1) - omega);
}
// runtime-dispatched versions for the generic kernels
template <class T> __device__ __forceinline__ T normalize_rt(int kind, T x) {
    return kind == NORM_TANH ? normalize<NORM_TANH>(x) : (kind == NORM_SOFTSIGN ? normalize<NORM_SOFTSIGN>(x) : normalize<NORM_SIGMOID>(x));
}
template <class T> __device__ __forceinline__ T normalize_deriv_rt(int kind, T w) {
    return kind == NORM_TANH ? normalize_deriv<NORM_TANH>(w) : (kind == NORM_SOFTSIGN ? normalize_deriv<NORM_SOFTSIGN>(w) : normalize_deriv<NORM_SIGMOID>(w));
}

// swish(x) = x*sigmoid(x) and d/dx = s + sigmoid(x)*(1 - s)
template <class T> __device__ __forceinline__ void swish_fwd(T x, T& s) { s = x * ksigmoid(x); }
template <class T> __device__ __forceinline__ void swish_both(T x, T& s, T& ds) {
    const T sg = ksigmoid(x);
    s = x * sg;
    ds = s + sg * (T(1) - s);
}

// basis value y(a) and dy/da as the reference's reverse rule states it (utils.jl:18,40,59 — iqf verbatim)
template <class T> __device__ __forceinline__ T basis_val(int kind, T a) {
    if (kind == BASIS_RBF) return kexp(-a * a);
    if (kind == BASIS_RSWAF) { const T t = ktanh(a); return T(1) - t * t; }
    return kdiv(T(1), T(1) + a * a);
}
template <class T> __device__ __forceinline__ void basis_both(int kind, T a, T& y, T& dy) {
    if (kind == BASIS_RBF) { y = kexp(-a * a); dy = T(-2) * a * y; }
    else if (kind == BASIS_RSWAF) { const T t = ktanh(a); y = T(1) - t * t; dy = T(-2) * t * y; }
    else { y = kdiv(T(1), T(1) + a * a); dy = T(-2) * a * y; }
}

// ---- packed fp32x2 FMA (Blackwell FFMA2: two IEEE fmas per issue slot) -----------------------------------
// d0 += a0*b0, d1 += a1*b1.  The float overloads emit fma.rn.f32x2 (bit-identical to two scalar FFMAs; a broadcast
// second operand costs no extra instruction); every other type falls back to two scalar FMAs.
__device__ __forceinline__ void kfma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rd) : "f"(d0), "f"(d1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rd) : "l"(ra), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(rd));
}
template <class T> __device__ __forceinline__ void kfma2(T& d0, T& d1, T a0, T a1, T b0, T b1) { d0 += a0 * b0; d1 += a1 * b1; }
template <class T> __device__ __forceinline__ void kfma2b(T& d0, T& d1, T a0, T a1, T c) { kfma2(d0, d1, a0, a1, c, c); }
// packed multiply (FMUL2): d0 = a0*b0, d1 = a1*b1
__device__ __forceinline__ void kmul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b0), "f"(b1));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(rd));
}
template <class T> __device__ __forceinline__ void kmul2(T& d0, T& d1, T a0, T a1, T b0, T b1) { d0 = a0 * b0; d1 = a1 * b1; }

// ---- PI controller helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float fastlog2f(float x) {
    const uint32_t bits = __float_as_uint(x);
    const float e = (float)((bits & 0x7F800000u) >> 23);
    float s, fe;
    if (bits & 0x00400000u) { s = __uint_as_float((bits & 0x007FFFFFu) | 0x3f000000u) - 1.0f; fe = e - 126.0f; }
    else                    { s = __uint_as_float((bits & 0x007FFFFFu) | 0x3f800000u) - 1.0f; fe = e - 127.0f; }
    return fe + s * (0.338953f * s + 2.198599f) / (s + 1.523692f);
}
__device__ __forceinline__ float fastpower(double x, float y) {
    if (x == 0.0) return 0.0f;
    return exp2f(y * fastlog2f((float)x));
}
__device__ __forceinline__ double eps_of(double x) {
    const double ax = fabs(x);
    return __longlong_as_double(__double_as_longlong(ax) + 1) - ax;
}

struct Ctrl {   // [EXT OrdinaryDiffEqCore 1.9.0] defaults for Tsit5
    static constexpr float beta1 = 0.14f, beta2 = 0.08f;
    static constexpr double gamma = 0.9, qmin = 0.2, qmax = 10.0, qoldinit = 1e-4;
};

// One controller decision.  Returns q; updates q11.  (stepsize_controller!(::PIController))
__device__ __forceinline__ double pi_q(double EEst, double qold, double& q11) {
    if (EEst == 0.0) return 1.0 / Ctrl::qmax;
    q11 = (double)fastpower(EEst, Ctrl::beta1);
    double q = q11 / (double)fastpower(qold, Ctrl::beta2);
    return fmax(1.0 / Ctrl::qmax, fmin(1.0 / Ctrl::qmin, q / Ctrl::gamma));
}

}  // namespace kanode
