// kanode_small.cuh — ensemble kernels for small 2-layer KAN-ODEs ([I,H,I], e.g. Lotka-Volterra [2,10,2] G=5).
//
// Mapping: ONE THREAD PER TRAJECTORY.  The whole ODE state, the 7 Tsit5 stage vectors and the
// per-trajectory step controller live in registers; the KAN weights (960 B for LV) are passed as a
// __grid_constant__ kernel parameter so every weight is an immediate constant-bank operand of an FFMA (no load
// instruction at all); the RBF features are produced and consumed in registers and never touch memory.
// The forward kernels (solve, dense solve + loss) live here; the adjoint (backward) kernel is the lane-group kernel of
// kanode_small_lg.cuh, which reads the dense record written here.
//
// Reference semantics: see kanode_math.cuh and oracle/kanode_oracle.cpp (same algorithm, CPU).
//   KDense forward            Lotka-Volterra/src/kdense.jl:109-130
//   rbf reverse rule          Lotka-Volterra/src/utils.jl:15-21
//   NeuralODE / loss / grad   Lotka-Volterra/LV_driver_KANODE.jl:180-184,197-203,284
#pragma once
#include "../../include/kanode.h"
#include <type_traits>

#include "kanode_math.cuh"

// tuning knobs (defaults chosen from B200 measurements, see profiles/)
#ifndef KANODE_UNROLL_J
#define KANODE_UNROLL_J 2      // hidden units processed per iteration of the rolled unit loops
#endif
#define KANODE_PRAGMA(x) _Pragma(#x)
#define KANODE_UNROLL(n) KANODE_PRAGMA(unroll n)

namespace kanode {

template <class T, int I_, int H_, int G_>
struct SmallParams {
    static constexpr int I = I_, H = H_, G = G_;
    static constexpr int OC1 = 0, OW1 = H * G * I, OC2 = OW1 + H * I, OW2 = OC2 + I * G * H, NP = OW2 + I * H;
    // rank-1 factor record of one backward stage: [hbar(H) | b1(I*G) sw1(I) | lam(I) | b2(H*G) sw2(H)]
    static constexpr int F_HBAR = 0, F_CA = H, QA = I * G + I, F_LAM = F_CA + QA, F_CB = F_LAM + I, QB = H * G + H,
                         NF = F_CB + QB;
    // per-hidden-unit packed weights (shared-memory copy used by the hot kernels):
    //   [C1[(i,g),o] (I*G) | W1[i,o] (I) | C2[(o,g),oo] (G*I, index g*I+oo) | W2[o,oo] (I)]
    static constexpr int NQ = I * (G + 1), UW = ((2 * NQ + 3) / 4) * 4, WPK = UW * H;
    T w[NP];
    T gs[G];      // grid[g] * hs
    T hs;         // Float32(1/h) * KRbfScale<T>  : basis = krbf_scaled(xn*hs - gs[g])
    T dk;         // d(basis)/d(xn) = dk * t * basis,  dk = -2*hs/KRbfScale^2   (utils.jl:18 times 1/h)
};

// ------------------------------------------------------------------------------------------------------
// KAN right-hand side, one sample, everything in registers
// ------------------------------------------------------------------------------------------------------
template <int NORM, class T, class P>
__device__ __forceinline__ void small_rhs(const P& p, const T (&u)[P::I], T (&du)[P::I]) {
    constexpr int I = P::I, H = P::H, G = P::G;
    T h[H];
#pragma unroll
    for (int o = 0; o < H; ++o) h[o] = T(0);
#pragma unroll
    for (int i = 0; i < I; ++i) {
        const T xn = normalize<NORM>(u[i]);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const T y = krbf_scaled(xn * p.hs - p.gs[g]);
#pragma unroll
            for (int o = 0; o < H; ++o) h[o] += p.w[P::OC1 + (i * G + g) * H + o] * y;
        }
        T s; swish_fwd(u[i], s);
#pragma unroll
        for (int o = 0; o < H; ++o) h[o] += p.w[P::OW1 + i * H + o] * s;
    }
#pragma unroll
    for (int o = 0; o < I; ++o) du[o] = T(0);
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const T xn = normalize<NORM>(h[i]);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const T y = krbf_scaled(xn * p.hs - p.gs[g]);
#pragma unroll
            for (int o = 0; o < I; ++o) du[o] += p.w[P::OC2 + (i * G + g) * I + o] * y;
        }
        T s; swish_fwd(h[i], s);
#pragma unroll
        for (int o = 0; o < I; ++o) du[o] += p.w[P::OW2 + i * I + o] * s;
    }
}

// ------------------------------------------------------------------------------------------------------
// fused forward-recompute + VJP: ubar = (df/du)^T lam, and the rank-1 factors of (df/dp)^T lam via st(idx, v).
// The layer-2 contraction is never computed (the adjoint does not need f(y)).
// ------------------------------------------------------------------------------------------------------
template <int NORM, class T, class P, class Store>
__device__ __forceinline__ void small_vjp(const P& p, const T (&y)[P::I], const T (&lam)[P::I], T (&ubar)[P::I],
                                          Store&& st) {
    constexpr int I = P::I, H = P::H, G = P::G;
    T h[H], xn1[I], db1[I * G], dsw1[I];
#pragma unroll
    for (int o = 0; o < H; ++o) h[o] = T(0);
#pragma unroll
    for (int i = 0; i < I; ++i) {
        xn1[i] = normalize<NORM>(y[i]);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const T a = xn1[i] * p.hs - p.gs[g];
            const T b = krbf_scaled(a);
            db1[i * G + g] = p.dk * a * b;                               // utils.jl:18 times d(arg)/d(xn)
            st(P::F_CA + i * G + g, b);
#pragma unroll
            for (int o = 0; o < H; ++o) h[o] += p.w[P::OC1 + (i * G + g) * H + o] * b;
        }
        T s; swish_both(y[i], s, dsw1[i]);
        st(P::F_CA + I * G + i, s);
#pragma unroll
        for (int o = 0; o < H; ++o) h[o] += p.w[P::OW1 + i * H + o] * s;
    }
    T hbar[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const T xn = normalize<NORM>(h[i]);
        T xnbar = T(0);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const T a = xn * p.hs - p.gs[g];
            const T b = krbf_scaled(a);
            st(P::F_CB + i * G + g, b);
            T bbar = T(0);
#pragma unroll
            for (int o = 0; o < I; ++o) bbar += p.w[P::OC2 + (i * G + g) * I + o] * lam[o];
            xnbar += (p.dk * a * b) * bbar;
        }
        T s, ds; swish_both(h[i], s, ds);
        st(P::F_CB + H * G + i, s);
        T sbar = T(0);
#pragma unroll
        for (int o = 0; o < I; ++o) sbar += p.w[P::OW2 + i * I + o] * lam[o];
        hbar[i] = xnbar * normalize_deriv<NORM>(xn) + sbar * ds;
        st(P::F_HBAR + i, hbar[i]);
    }
#pragma unroll
    for (int i = 0; i < I; ++i) {
        T xnbar = T(0);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            T bbar = T(0);
#pragma unroll
            for (int o = 0; o < H; ++o) bbar += p.w[P::OC1 + (i * G + g) * H + o] * hbar[o];
            xnbar += db1[i * G + g] * bbar;
        }
        T sbar = T(0);
#pragma unroll
        for (int o = 0; o < H; ++o) sbar += p.w[P::OW1 + i * H + o] * hbar[o];
        ubar[i] = xnbar * normalize_deriv<NORM>(xn1[i]) + sbar * dsw1[i];
        st(P::F_LAM + i, lam[i]);
    }
}

// ------------------------------------------------------------------------------------------------------
// TMA (1-D bulk async copy) staging of the packed weights into shared memory + mbarrier completion
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// all threads of the block call this once; afterwards wsm holds the WPK packed weights
template <class T, int WPK>
__device__ __forceinline__ void stage_weights(T* wsm, uint64_t* bar, const T* __restrict__ wpk) {
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)(sizeof(T) * WPK));
        tma_load_1d(wsm, wpk, (uint32_t)(sizeof(T) * WPK), bar);
    }
    mbar_wait(bar, 0);
}

// Gaussian RBF of one normalised input on all G grid points, two grid points per packed instruction (FFMA2 / FMUL2):
//   na[g] = gs[g] - xn*hs  (the NEGATED scaled argument: its square and the products below are bit-identical to the scalar form),
//   b[g] = exp2(-na[g]^2);  ndb[g] = (dk * a) * b = ((-dk) * na) * b when WITH_DB
template <bool WITH_DB, class T, class P>
__device__ __forceinline__ void rbf_eval(const P& p, T xn, T (&b)[P::G], T (&db)[P::G]) {
    constexpr int G = P::G;
    if constexpr (sizeof(T) == 4) {
        const T nhs = -p.hs, ndk = -p.dk;
        T na[G], t[G];
#pragma unroll
        for (int g = 0; g + 1 < G; g += 2) {
            na[g] = p.gs[g]; na[g + 1] = p.gs[g + 1];
            kfma2b(na[g], na[g + 1], nhs, nhs, xn);
            kmul2(t[g], t[g + 1], na[g], na[g + 1], na[g], na[g + 1]);
        }
        if constexpr (G % 2 == 1) { na[G - 1] = fmaf(nhs, xn, p.gs[G - 1]); t[G - 1] = na[G - 1] * na[G - 1]; }
#pragma unroll
        for (int g = 0; g < G; ++g) b[g] = kex2(-t[g]);
        if constexpr (WITH_DB) {
#pragma unroll
            for (int g = 0; g + 1 < G; g += 2) {
                T m0, m1;
                kmul2(m0, m1, ndk, ndk, na[g], na[g + 1]);
                kmul2(db[g], db[g + 1], m0, m1, b[g], b[g + 1]);
            }
            if constexpr (G % 2 == 1) db[G - 1] = (ndk * na[G - 1]) * b[G - 1];
        }
    } else {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const T a = xn * p.hs - p.gs[g];
            b[g] = krbf_scaled(a);
            if constexpr (WITH_DB) db[g] = p.dk * a * b[g];
        }
    }
}

// features of the state components: f[i*G+g] = basis_g(norm(u_i)), f[I*G+i] = swish(u_i)  (layout of the packed C1|W1)
template <int NORM, class T, class P>
__device__ __forceinline__ void input_features(const P& p, const T (&u)[P::I], T (&f)[P::NQ]) {
    constexpr int I = P::I, G = P::G;
#pragma unroll
    for (int i = 0; i < I; ++i) {
        const T xn = normalize<NORM>(u[i]);
        T bb[G], dummy[G];
        rbf_eval<false>(p, xn, bb, dummy);
#pragma unroll
        for (int g = 0; g < G; ++g) f[i * G + g] = bb[g];
        swish_fwd(u[i], f[I * G + i]);
    }
}

// KAN right-hand side with the weights in shared memory: one ROLLED loop over the hidden units (small code)
template <int NORM, class T, class P>
__device__ __forceinline__ void small_rhs_sm(const P& p, const T* __restrict__ wsm, const T (&u)[P::I], T (&du)[P::I]) {
    constexpr int I = P::I, H = P::H, G = P::G, NQ = P::NQ;
    T f[NQ];
    input_features<NORM>(p, u, f);
#pragma unroll
    for (int o = 0; o < I; ++o) du[o] = T(0);
    KANODE_UNROLL(KANODE_UNROLL_J)
    for (int j = 0; j < H; ++j) {
        const T* w = wsm + j * P::UW;
        T h = T(0);
#pragma unroll
        for (int q = 0; q < NQ; ++q) h += w[q] * f[q];
        const T xn = normalize<NORM>(h);
        T bb[G], dummy[G];
        rbf_eval<false>(p, xn, bb, dummy);
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
            for (int o = 0; o < I; ++o) du[o] += w[NQ + g * I + o] * bb[g];
        }
        T s; swish_fwd(h, s);
#pragma unroll
        for (int o = 0; o < I; ++o) du[o] += w[NQ + G * I + o] * s;
    }
}

// ---- dense forward record, array of structures: [t (fp64) | dt | u(I) | k1..k7 (7*I) | pad] in units of T ----
template <class T, int I_> struct RecLayout {
    static constexpr int OT = 8 / (int)sizeof(T);                     // T slots taken by the fp64 start time
    static constexpr int DT = OT, U = OT + 1, K = OT + 1 + I_;
    static constexpr int V = 16 / (int)sizeof(T);                     // T per 16-byte vector
    static constexpr int RS = ((OT + 1 + 8 * I_ + V - 1) / V) * V;    // record stride (T)
};
__device__ __forceinline__ double rec_get_time(const float* r) { return __hiloint2double(__float_as_int(r[1]), __float_as_int(r[0])); }
__device__ __forceinline__ double rec_get_time(const double* r) { return r[0]; }
__device__ __forceinline__ void rec_put_time(float* r, double t) { r[0] = __int_as_float(__double2loint(t)); r[1] = __int_as_float(__double2hiint(t)); }
__device__ __forceinline__ void rec_put_time(double* r, double t) { r[0] = t; }

// 16-byte vector copies between memory (shared or global, 16-byte aligned) and register arrays
template <int N> __device__ __forceinline__ void ldv(const float* p, float (&v)[N]) {
    static_assert(N % 4 == 0, "vector length");
#pragma unroll
    for (int k = 0; k < N; k += 4) { const float4 q = *reinterpret_cast<const float4*>(p + k); v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w; }
}
template <int N> __device__ __forceinline__ void ldv(const double* p, double (&v)[N]) {
    static_assert(N % 2 == 0, "vector length");
#pragma unroll
    for (int k = 0; k < N; k += 2) { const double2 q = *reinterpret_cast<const double2*>(p + k); v[k] = q.x; v[k + 1] = q.y; }
}
template <int N> __device__ __forceinline__ void stv(float* p, const float (&v)[N]) {
    static_assert(N % 4 == 0, "vector length");
#pragma unroll
    for (int k = 0; k < N; k += 4) *reinterpret_cast<float4*>(p + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
}
template <int N> __device__ __forceinline__ void stv(double* p, const double (&v)[N]) {
    static_assert(N % 2 == 0, "vector length");
#pragma unroll
    for (int k = 0; k < N; k += 2) *reinterpret_cast<double2*>(p + k) = make_double2(v[k], v[k + 1]);
}
__device__ __forceinline__ float shfl_t(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_t(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// ------------------------------------------------------------------------------------------------------
// argument blocks
// ------------------------------------------------------------------------------------------------------
template <class T> struct SmallFwdArgs {
    const T* wpk;           // packed per-unit weights (global), staged to smem by TMA
    const T* u0;            // [B][I]
    int64_t B;
    double t0, t1;
    const double* saveat;   // device, ascending, inside [t0,t1]
    int nsave;
    T abstol, reltol;
    int maxiters;
    T* out;                 // [B][nsave][I] or null
    kanode_stats* stats;    // [B] or null
    // dense record for the adjoint (DENSE kernels)
    double* rec_t;          // [cap][B]   start time of each accepted step
    T* rec;                 // [cap][1 + 8*I][B]: dt, u, k1..k7
    int cap;
    int* nsteps;            // [B]
    int* retcode;           // [B]
    // loss pieces (DENSE kernels)
    const T* target;        // [B][nsave][I]
    T* dg;                  // dL/du(t_s): [nsave][I][B], or [B][nsave][I] in the AOS kernels
    double* loss_sum;       // scalar accumulator
    // dt-replay (parity tooling): end times of the accepted steps of another run; the controller is bypassed
    const double* rp_t;     // [B][rp_cap] ascending, NaN-padded; or null
    int rp_cap;
};

// ------------------------------------------------------------------------------------------------------
// forward: adaptive Tsit5 with saveat interpolation; DENSE additionally records every accepted step and
// evaluates the loss / dL/du at the save times.
// ------------------------------------------------------------------------------------------------------
// AOS: the dense record is one RecLayout structure per accepted step ([B][cap][RS]) and dg is [B][nsave][I] — the layouts
// the lane-group backward kernel reads (kanode_small_lg.cuh).
template <class T, class P, int NORM, bool DENSE, bool AOS = false>
__global__ void __launch_bounds__(64) small_forward_kernel(const __grid_constant__ P prm, const SmallFwdArgs<T> a) {
    constexpr int I = P::I;
    __shared__ __align__(16) T wsm[P::WPK];
    __shared__ uint64_t wbar;
    stage_weights<T, P::WPK>(wsm, &wbar, a.wpk);
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = b < a.B;
    double lsum = 0.0;
    if (active) {
        const int64_t B = a.B;
        T u[I], uprev[I], k[7][I];
#pragma unroll
        for (int i = 0; i < I; ++i) { u[i] = a.u0[b * I + i]; uprev[i] = u[i]; }
        // stages 2..7 are multiplied by zero tableau entries before they are first written: they must not hold
        // NaN/Inf bit patterns
#pragma unroll
        for (int j = 1; j < 7; ++j)
#pragma unroll
            for (int i = 0; i < I; ++i) k[j][i] = T(0);
        small_rhs_sm<NORM>(prm, wsm, u, k[0]);
        int nf = 1, naccept = 0, nreject = 0, ret = RET_SUCCESS;
        const double t0 = a.t0, t1 = a.t1, dtmax = fabs(t1 - t0);
        const double dtmin0 = fmax(eps_of(t0), eps_of(t1));
        const T abstol = a.abstol, reltol = a.reltol;
        double t = t0, dt;
        {   // ---- ode_determine_initdt (Hairer) ----
            T s0 = T(0), s1 = T(0), sk[I], u1[I], f1[I];
#pragma unroll
            for (int i = 0; i < I; ++i) {
                sk[i] = abstol + kabs(u[i]) * reltol;
                const T x0 = u[i] / sk[i], x1 = k[0][i] / sk[i];
                s0 += x0 * x0; s1 += x1 * x1;
            }
            const double d0 = sqrt((double)s0 / I), d1 = sqrt((double)s1 / I);
            double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
            dt0 = fmin(dt0, dtmax);
#pragma unroll
            for (int i = 0; i < I; ++i) u1[i] = u[i] + (T)dt0 * k[0][i];
            small_rhs_sm<NORM>(prm, wsm, u1, f1);
            nf += 2;
            T s2 = T(0);
#pragma unroll
            for (int i = 0; i < I; ++i) { const T x = (f1[i] - k[0][i]) / sk[i]; s2 += x * x; }
            const double d2 = sqrt((double)s2 / I) / dt0;
            const double mx = fmax(d1, d2);
            const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
            dt = fmax(dtmin0, fmin(fmin(100.0 * dt0, dt1), dtmax));
        }
        double qold = Ctrl::qoldinit, q11 = 1.0, dtpropose = dt;
        bool accept = false;
        int iter = 0, sidx = 0, nrec = 0;
        const double* rp = a.rp_t ? a.rp_t + b * (int64_t)a.rp_cap : nullptr;
        if (t0 == t1) {   // degenerate span: outputs are u0
            for (; sidx < a.nsave; ++sidx)
#pragma unroll
                for (int i = 0; i < I; ++i) if (a.out) a.out[(b * a.nsave + sidx) * I + i] = u[i];
        }
        while (t < t1) {
            // ---- loopheader! ----
            if (iter > 0) {
                if (!accept) dt = dt / fmin(1.0 / Ctrl::qmin, q11 / Ctrl::gamma);
                else dt = dtpropose;
            }
            ++iter;
            const double dtmin_t = fmax(eps_of(t), dtmin0);
            dt = fmin(fmax(fmin(fabs(dt), dtmax), dtmin_t), t1 - t);
            double rp_next = t1;
            if (rp) {                                                  // replay: the step ends where the recorded one ended
                rp_next = naccept < a.rp_cap ? rp[naccept] : t1;
                if (!(rp_next > t) || !(rp_next <= t1)) rp_next = t1;
                dt = rp_next - t;
            }
            if (iter > a.maxiters) { ret = RET_MAXITERS; break; }
            if (!rp && !(dt > dtmin_t) && (t + dt < t1 || !accept) && iter > 1) { ret = RET_DTMIN; break; }
            if (dt != dt) { ret = RET_UNSTABLE; break; }
            // ---- perform_step! ----
            const T h = (T)dt;
            T unew[I];
#pragma unroll
            for (int i = 0; i < I; ++i) unew[i] = uprev[i];
#pragma unroll 1
            for (int s = 1; s < 7; ++s) {
                T us[I], ks[I];
#pragma unroll
                for (int i = 0; i < I; ++i) {
                    T acc = T(0);
#pragma unroll
                    for (int j = 0; j < 6; ++j) acc += Tab<T>::a(s, j) * k[j][i];
                    us[i] = uprev[i] + h * acc;
                }
                small_rhs_sm<NORM>(prm, wsm, us, ks);
#pragma unroll
                for (int j = 1; j < 7; ++j)
                    if (j == s) {
#pragma unroll
                        for (int i = 0; i < I; ++i) k[j][i] = ks[i];
                    }
                if (s == 6) {
#pragma unroll
                    for (int i = 0; i < I; ++i) unew[i] = us[i];
                }
            }
            nf += 6;
            T es = T(0);
            bool bad = false;
#pragma unroll
            for (int i = 0; i < I; ++i) {
                T ut = T(0);
#pragma unroll
                for (int j = 0; j < 7; ++j) ut += Tab<T>::bt(j) * k[j][i];
                ut *= h;
                const T sc = abstol + kmax(kabs(uprev[i]), kabs(unew[i])) * reltol;
                const T r = ut / sc;
                es += r * r;
                bad |= (unew[i] != unew[i]);
            }
            const double EEst = (double)ksqrt(es / T(I));
#ifdef KANODE_DEBUG
            if (b == 0) printf("fwd iter %d t %g dt %g es %g EEst %g unew %g %g k0 %g k6 %g abstol %g reltol %g\n", iter, t, dt, (double)es, EEst, (double)unew[0], (double)unew[1], (double)k[0][0], (double)k[6][0], (double)abstol, (double)reltol);
#endif
            if (EEst != EEst || bad) { ret = RET_UNSTABLE; break; }
            // ---- loopfooter!: PI controller ----
            const double q = pi_q(EEst, qold, q11);
            accept = rp ? true : (EEst <= 1.0);
            if (accept) {
                ++naccept;
                qold = fmax(EEst, Ctrl::qoldinit);
                const double dtnew = dt / q;
                double tnew = t + dt;
                if (rp) tnew = rp_next;
                else if (fabs(tnew - t1) < 100.0 * eps_of(fmax(fabs(t), fabs(t1)))) tnew = t1;
                dtpropose = fmax(fmin(dtmax, fabs(dtnew)), fmax(eps_of(tnew), dtmin0));
                if (DENSE) {
                    if (nrec >= a.cap) { ret = RET_OVERFLOW; break; }
                    if constexpr (AOS) {
                        using RL = RecLayout<T, I>;
                        T tmp[RL::RS];
#pragma unroll
                        for (int f = 0; f < RL::RS; ++f) tmp[f] = T(0);
                        rec_put_time(tmp, t);
                        tmp[RL::DT] = h;
#pragma unroll
                        for (int i = 0; i < I; ++i) tmp[RL::U + i] = uprev[i];
#pragma unroll
                        for (int j = 0; j < 7; ++j)
#pragma unroll
                            for (int i = 0; i < I; ++i) tmp[RL::K + j * I + i] = k[j][i];
                        stv(a.rec + (b * (int64_t)a.cap + nrec) * RL::RS, tmp);
                    } else {
                        a.rec_t[(int64_t)nrec * B + b] = t;
                        T* r = a.rec + (int64_t)nrec * (1 + 8 * I) * B + b;
                        r[0] = h;
#pragma unroll
                        for (int i = 0; i < I; ++i) r[(int64_t)(1 + i) * B] = uprev[i];
#pragma unroll
                        for (int j = 0; j < 7; ++j)
#pragma unroll
                            for (int i = 0; i < I; ++i) r[(int64_t)(1 + I + j * I + i) * B] = k[j][i];
                    }
                    ++nrec;
                }
                // saveat: dense output inside (t, tnew]  (and t0 itself on the first step)
                while (sidx < a.nsave && a.saveat[sidx] <= tnew) {
                    const T th = (T)((a.saveat[sidx] - t) / dt);
                    T bw[7]; interp_weights(th, bw);
#pragma unroll
                    for (int i = 0; i < I; ++i) {
                        T acc = T(0);
#pragma unroll
                        for (int j = 0; j < 7; ++j) acc += bw[j] * k[j][i];
                        const T v = uprev[i] + h * acc;
                        if (a.out) a.out[(b * a.nsave + sidx) * I + i] = v;
                        if (DENSE && a.target) {                     // target == null: the caller supplies dL/du(t_s) (a.dg is pre-filled)
                            const T e = v - a.target[(b * a.nsave + sidx) * I + i];
                            lsum += (double)e * (double)e;
                            a.dg[AOS ? (b * a.nsave + sidx) * I + i : ((int64_t)sidx * I + i) * B + b] = (T(2) / (T)((double)I * a.nsave)) * e;
                        }
                    }
                    ++sidx;
                }
                t = tnew;
#pragma unroll
                for (int i = 0; i < I; ++i) { uprev[i] = unew[i]; u[i] = unew[i]; k[0][i] = k[6][i]; }
            } else {
                ++nreject;
            }
        }
        if (ret != RET_SUCCESS) {   // leave NaN in the unsaved outputs of a failed trajectory
            for (; sidx < a.nsave; ++sidx)
#pragma unroll
                for (int i = 0; i < I; ++i) {
                    if (a.out) a.out[(b * a.nsave + sidx) * I + i] = T(NAN);
                    if (DENSE) a.dg[AOS ? (b * a.nsave + sidx) * I + i : ((int64_t)sidx * I + i) * B + b] = T(0);
                }
        }
        if (a.stats) a.stats[b] = kanode_stats{naccept, nreject, nf, ret};
        if (DENSE) { a.nsteps[b] = nrec; a.retcode[b] = ret; }
    }
    if (DENSE) {   // block-level loss reduction, one atomic per warp
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
        if ((threadIdx.x & 31) == 0 && lsum != 0.0) atomicAdd(a.loss_sum, lsum);
    }
}

// ------------------------------------------------------------------------------------------------------
// rank-1 structure of the parameter gradient of one RHS evaluation (used by the batch VJP kernel)
// ------------------------------------------------------------------------------------------------------
template <class T, class P> struct GPhase {
    // visit every parameter-gradient component j with the per-stage derivative factors a_s (by o) and c_s (by q):
    // k_s[j] = -(a_s[o] * c_s[q]).  fn(j, av[NS], cv[NS]) for the NS stage slots starting at slot0.
    template <int NS, class Fn>
    static __device__ __forceinline__ void for_each(const T* fac, int64_t B, int slot0, Fn&& fn) {
        constexpr int H = P::H, I = P::I;
        // block A: a = hbar (H), c = [b1; sw1] (QA), j = OC1 + q*H + o   (C1 then W1 are contiguous)
#pragma unroll 1
        for (int o = 0; o < H; ++o) {
            T av[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) av[s] = fac[((int64_t)(slot0 + s) * P::NF + P::F_HBAR + o) * B];
#pragma unroll 1
            for (int q = 0; q < P::QA; ++q) {
                T cv[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) cv[s] = fac[((int64_t)(slot0 + s) * P::NF + P::F_CA + q) * B];
                fn(P::OC1 + q * H + o, av, cv);
            }
        }
        // block B: a = lam (I), c = [b2; sw2] (QB), j = OC2 + q*I + o   (C2 then W2 are contiguous)
        T aw[I][NS];
#pragma unroll
        for (int o = 0; o < I; ++o)
#pragma unroll
            for (int s = 0; s < NS; ++s) aw[o][s] = fac[((int64_t)(slot0 + s) * P::NF + P::F_LAM + o) * B];
#pragma unroll 1
        for (int q = 0; q < P::QB; ++q) {
            T cv[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) cv[s] = fac[((int64_t)(slot0 + s) * P::NF + P::F_CB + q) * B];
#pragma unroll
            for (int o = 0; o < I; ++o) fn(P::OC2 + q * I + o, aw[o], cv);
        }
    }
};

// ------------------------------------------------------------------------------------------------------
// batch RHS / VJP (kanode_rhs, kanode_vjp)
// ------------------------------------------------------------------------------------------------------
template <class T, class P, int NORM>
__global__ void __launch_bounds__(128) small_rhs_kernel(const __grid_constant__ P prm, const T* u, T* du, int64_t B) {
    constexpr int I = P::I;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T x[I], y[I];
#pragma unroll
    for (int i = 0; i < I; ++i) x[i] = u[b * I + i];
    small_rhs<NORM>(prm, x, y);
#pragma unroll
    for (int i = 0; i < I; ++i) du[b * I + i] = y[i];
}

// pbar_rows[j][b] = ((df/dp)^T lam_b)[j]; reduced over b by reduce_rows_kernel
template <class T, class P, int NORM>
__global__ void __launch_bounds__(64) small_vjp_kernel(const __grid_constant__ P prm, const T* u, const T* lam, T* ubar,
                                                       T* fac /*[1][NF][B]*/, T* pbar_rows /*[NP][B]*/, int64_t B) {
    constexpr int I = P::I;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T x[I], l[I], ub[I];
#pragma unroll
    for (int i = 0; i < I; ++i) { x[i] = u[b * I + i]; l[i] = lam[b * I + i]; }
    T* f = fac + b;
    small_vjp<NORM>(prm, x, l, ub, [&](int idx, T v) { f[(int64_t)idx * B] = v; });
#pragma unroll
    for (int i = 0; i < I; ++i) ubar[b * I + i] = ub[i];
    GPhase<T, P>::template for_each<1>(f, B, 0, [&](int j, const T (&av)[1], const T (&cv)[1]) {
        pbar_rows[(int64_t)j * B + b] = av[0] * cv[0];
    });
}

// out[r] = scale * sum_b rows[r][b]   (double accumulation, one block per row)
template <class T, class OutT>
__global__ void __launch_bounds__(256) reduce_rows_kernel(const T* rows, int64_t B, OutT* out, double scale) {
    const T* src = rows + (int64_t)blockIdx.x * B;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) acc += (double)src[i];
    __shared__ double sh[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
        out[blockIdx.x] = (OutT)(s * scale);
    }
}

}  // namespace kanode
