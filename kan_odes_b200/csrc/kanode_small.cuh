// kanode_small.cuh — ensemble kernels for small 2-layer KAN-ODEs ([I,H,I], e.g. Lotka-Volterra [2,10,2] G=5).
//
// Mapping: ONE THREAD PER TRAJECTORY.  The whole ODE state, the 7 Tsit5 stage vectors and the
// per-trajectory step controller live in registers; the KAN weights (960 B for LV) are passed as a
// __grid_constant__ kernel parameter so every weight is an immediate constant-bank operand of an FFMA (no load
// instruction at all); the RBF features are produced and consumed in registers and never touch memory.
// The backward kernel integrates z=[lambda; g] per trajectory exactly like the reference's InterpolatingAdjoint:
// lambda and its stages stay in registers; the parameter-gradient part g (NP values per trajectory) is never
// materialised per stage — each stage stores only the rank-1 factors of dg/dt (NF = 84 values for LV) and the
// step-end pass rebuilds sum_s b_s k_s[j] and the error-estimate term per component.
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
#ifndef KANODE_BWD_BT
#define KANODE_BWD_BT 64       // trajectories (threads) per block of the backward kernel
#endif
#ifndef KANODE_BWD_MINB
#define KANODE_BWD_MINB 4      // resident blocks per SM the backward kernel is compiled for
#endif
#ifndef KANODE_UNROLL_S
#define KANODE_UNROLL_S 1      // stages per iteration of the rolled stage loops of the step-end gradient pass
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

// fused forward-recompute + VJP with shared-memory weights; h_j and hbar_j go straight to the stage record
// (rec = this thread's slot base, element f at rec[f*nthr])
template <int NORM, int UJ, class T, class P>
__device__ __forceinline__ void small_vjp_sm(const P& p, const T* __restrict__ wsm, const T (&y)[P::I], const T (&lam)[P::I],
                                             T (&ubar)[P::I], T* rec, int nthr, int off_h, int off_hbar) {
    constexpr int I = P::I, H = P::H, G = P::G, NQ = P::NQ;
    T f[NQ], df[NQ], bb[NQ];                  // features, their derivatives wrt the input, and sum_j w1[q][j]*hbar_j
#pragma unroll
    for (int i = 0; i < I; ++i) {
        const T xn = normalize<NORM>(y[i]);
        const T dn = normalize_deriv<NORM>(xn);
        T rb[G], rdb[G];
        rbf_eval<true>(p, xn, rb, rdb);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            f[i * G + g] = rb[g];
            df[i * G + g] = rdb[g] * dn;                                 // utils.jl:18 * d(arg)/d(xn) * norm'
        }
        swish_both(y[i], f[I * G + i], df[I * G + i]);
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) bb[q] = T(0);
#pragma unroll UJ
    for (int j = 0; j < H; ++j) {
        const T* w = wsm + j * P::UW;
        T wl[NQ];
        T h = T(0);
        if constexpr (sizeof(T) == 4 && NQ % 2 == 0) {          // two partial sums, one FFMA2 per pair of features
            T h1 = T(0);
#pragma unroll
            for (int q = 0; q < NQ; q += 2) { wl[q] = w[q]; wl[q + 1] = w[q + 1]; kfma2(h, h1, wl[q], wl[q + 1], f[q], f[q + 1]); }
            h += h1;
        } else {
#pragma unroll
            for (int q = 0; q < NQ; ++q) { wl[q] = w[q]; h += wl[q] * f[q]; }
        }
        const T xn = normalize<NORM>(h);
        T xnbar = T(0);
        T rb[G], rdb[G];                                                 // (the outer bb[] accumulates sum_j w1[q][j] * hbar_j)
        rbf_eval<true>(p, xn, rb, rdb);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            T bbar = T(0);
#pragma unroll
            for (int o = 0; o < I; ++o) bbar += w[NQ + g * I + o] * lam[o];
            xnbar += rdb[g] * bbar;
        }
        T s, ds; swish_both(h, s, ds);
        T sbar = T(0);
#pragma unroll
        for (int o = 0; o < I; ++o) sbar += w[NQ + G * I + o] * lam[o];
        const T hb = xnbar * normalize_deriv<NORM>(xn) + sbar * ds;
        rec[(off_h + j) * nthr] = h;
        rec[(off_hbar + j) * nthr] = hb;
        if constexpr (NQ % 2 == 0) {
#pragma unroll
            for (int q = 0; q < NQ; q += 2) kfma2b(bb[q], bb[q + 1], wl[q], wl[q + 1], hb);
        } else {
#pragma unroll
            for (int q = 0; q < NQ; ++q) bb[q] += wl[q] * hb;
        }
    }
#pragma unroll
    for (int i = 0; i < I; ++i) {
        T xb = bb[I * G + i] * df[I * G + i];
#pragma unroll
        for (int g = 0; g < G; ++g) xb += bb[i * G + g] * df[i * G + g];
        ubar[i] = xb;
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

template <class T> struct SmallBwdArgs {
    const T* wpk;           // packed per-unit weights (global), staged to smem by TMA
    int64_t B;
    double t0, t1;
    const double* saveat;
    int nsave;
    T abstol, reltol;
    int maxiters;
    const double* rec_t; const T* rec; int cap; const int* nsteps; const int* retcode;
    const T* dg;            // [nsave][I][B]
    T* fac;                 // [7][NF][B]  stage factors
    T* g;                   // [2][NP][B]  double-buffered gradient state; result ends in buffer 0
    T* du0;                 // [B][I] or null
    kanode_stats* stats;    // [B] or null
    // scheduling (optional): trajectories predicted to be long are listed in long_list and run in a separate launch
    const int* long_list;   // launch of the long ones: position -> trajectory; null in the bulk launch
    const int* long_count;  // device scalar: entries of long_list (clamped to the launch size by the kernel)
    const unsigned char* long_flag;   // bulk launch: [B] 1 = handled by the long launch (skip), or null
    int* attempts;          // [B] step attempts of this backward solve (feeds the next call's prediction), or null
    unsigned long long* attempts_sum; // device scalar accumulating sum of attempts, or null
    int64_t gidn;           // launch positions of this launch
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
                        if (DENSE) {
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
// backward: interpolating adjoint on z = [lambda(I); g(NP)], T -> t0, tstops + jumps at the save times
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

// Stage record kept in SHARED memory for the 7 Tsit5 stages of one backward step (per trajectory):
//   [ y(I) | h(H) | hbar(H) | lam(I) ]  — the forward state, the hidden pre-activations, the hidden cotangent and
// the stage adjoint.  Every rank-1 factor of dg/dt is a cheap function of these: the step-end pass recomputes the
// activations (tanh/RBF/SiLU) per unit instead of storing the 84 factors per stage in memory.
template <class P> struct StageRec {
    static constexpr int Y = 0, HH = P::I, HBAR = P::I + P::H, LAM = P::I + 2 * P::H, N = 2 * P::I + 2 * P::H;
};

// fused forward-recompute + VJP returning the hidden pre-activation h and its cotangent hbar (registers)
template <int NORM, class T, class P>
__device__ __forceinline__ void small_vjp_h(const P& p, const T (&y)[P::I], const T (&lam)[P::I], T (&ubar)[P::I],
                                            T (&h)[P::H], T (&hbar)[P::H]) {
    constexpr int I = P::I, H = P::H, G = P::G;
    T xn1[I], db1[I * G], dsw1[I];
#pragma unroll
    for (int o = 0; o < H; ++o) h[o] = T(0);
#pragma unroll
    for (int i = 0; i < I; ++i) {
        xn1[i] = normalize<NORM>(y[i]);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const T a = xn1[i] * p.hs - p.gs[g];
            const T b = krbf_scaled(a);
            db1[i * G + g] = p.dk * a * b;
#pragma unroll
            for (int o = 0; o < H; ++o) h[o] += p.w[P::OC1 + (i * G + g) * H + o] * b;
        }
        T s; swish_both(y[i], s, dsw1[i]);
#pragma unroll
        for (int o = 0; o < H; ++o) h[o] += p.w[P::OW1 + i * H + o] * s;
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const T xn = normalize<NORM>(h[i]);
        T xnbar = T(0);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const T a = xn * p.hs - p.gs[g];
            const T b = krbf_scaled(a);
            T bbar = T(0);
#pragma unroll
            for (int o = 0; o < I; ++o) bbar += p.w[P::OC2 + (i * G + g) * I + o] * lam[o];
            xnbar += (p.dk * a * b) * bbar;
        }
        T s, ds; swish_both(h[i], s, ds);
        T sbar = T(0);
#pragma unroll
        for (int o = 0; o < I; ++o) sbar += p.w[P::OW2 + i * I + o] * lam[o];
        hbar[i] = xnbar * normalize_deriv<NORM>(xn) + sbar * ds;
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
    }
}

// c-vector of one KDense input unit: [basis(G); swish]  (the rank-1 factor shared by all outputs of that unit)
template <int NORM, class T, class P>
__device__ __forceinline__ void unit_features(const P& p, T x, T (&c)[P::G + 1]) {
    const T xn = normalize<NORM>(x);
    T bb[P::G], dummy[P::G];
    rbf_eval<false>(p, xn, bb, dummy);
#pragma unroll
    for (int g = 0; g < P::G; ++g) c[g] = bb[g];
    swish_fwd(x, c[P::G]);
}

// LAT = 0: throughput build (rolled loops, small code, 4 blocks/SM).  LAT = 1: latency build for the launch of the
// predicted-long trajectories (one warp owns an SM): unit and stage loops fully unrolled for instruction-level parallelism.
template <class T, class P, int NORM, int LAT = 0>
__global__ void __launch_bounds__(KANODE_BWD_BT, LAT ? 1 : KANODE_BWD_MINB) small_backward_kernel(const __grid_constant__ P prm, const SmallBwdArgs<T> a) {
    constexpr int I = P::I, H = P::H, G = P::G, NP = P::NP, NZ = I + NP, RS = 1 + 8 * I;
    constexpr int US = LAT ? 7 : KANODE_UNROLL_S;                  // stage-loop unroll of the gradient pass
    using SR = StageRec<P>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sm = reinterpret_cast<T*>(smem_raw) + threadIdx.x;          // element (slot, f) at sm[(slot*SR::N + f)*nthr]
    const int nthr = blockDim.x;
    T* wsm = reinterpret_cast<T*>(smem_raw) + 7 * SR::N * nthr;    // packed weights behind the stage records
    uint64_t* wbar = reinterpret_cast<uint64_t*>(wsm + P::WPK);
    stage_weights<T, P::WPK>(wsm, wbar, a.wpk);
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= a.gidn) return;
    int64_t b = gid;
    if (a.long_list) {                                            // launch of the predicted-long trajectories
        if (gid >= *a.long_count) return;
        b = a.long_list[gid];
    } else if (a.long_flag && a.long_flag[b]) return;             // bulk launch: that one runs in the long launch
    const int64_t B = a.B;
    T* gbuf = a.g + b;           // element (buf, j) at gbuf[(buf*NP + j)*B]
#pragma unroll 1
    for (int j = 0; j < NP; ++j) gbuf[(int64_t)j * B] = T(0);
    T lam[I], lprev[I], kl[7][I];
#pragma unroll
    for (int i = 0; i < I; ++i) { lam[i] = T(0); lprev[i] = T(0); }
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int i = 0; i < I; ++i) kl[j][i] = T(0);   // zero-weighted stages must be finite (see forward kernel)
    int nf = 0, naccept = 0, nreject = 0, ret = a.retcode[b];
    const int nsteps = a.nsteps[b];
    if (ret != RET_SUCCESS || nsteps <= 0) {
        if (a.stats) a.stats[b] = kanode_stats{0, 0, 0, ret};
        if (a.du0) for (int i = 0; i < I; ++i) a.du0[b * I + i] = T(0);
        if (a.attempts) a.attempts[b] = 0;
        return;
    }
    // cached forward record (dense output of the forward solve)
    int ridx = nsteps - 1;
    double rt, rt_next;           // record covers [rt, rt_next]
    T rdt, ru[I], rk[7][I];
    auto load_rec = [&](int idx) {
        rt = a.rec_t[(int64_t)idx * B + b];
        const T* r = a.rec + (int64_t)idx * RS * B + b;
        rdt = r[0];
#pragma unroll
        for (int i = 0; i < I; ++i) ru[i] = r[(int64_t)(1 + i) * B];
#pragma unroll
        for (int j = 0; j < 7; ++j)
#pragma unroll
            for (int i = 0; i < I; ++i) rk[j][i] = r[(int64_t)(1 + I + j * I + i) * B];
        rt_next = (idx + 1 < nsteps) ? a.rec_t[(int64_t)(idx + 1) * B + b] : a.t1;
        ridx = idx;
    };
    load_rec(ridx);
    auto eval_y = [&](double t, T (&y)[I]) {      // y = sol(t), right-continuous at step boundaries
        while (t < rt && ridx > 0) load_rec(ridx - 1);
        while (t >= rt_next && ridx + 1 < nsteps) load_rec(ridx + 1);
        const T th = (T)((t - rt) / (double)rdt);
        T bw[7]; interp_weights(th, bw);
#pragma unroll
        for (int i = 0; i < I; ++i) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < 7; ++j) acc += bw[j] * rk[j][i];
            y[i] = ru[i] + rdt * acc;
        }
    };
    // one adjoint RHS evaluation at (t, l): dl = -(df/du)^T l; stage record -> shared-memory slot
    auto adj_eval = [&](double t, const T (&l)[I], T (&dl)[I], int slot) {
        T y[I], ub[I];
        eval_y(t, y);
        T* s = sm + slot * SR::N * nthr;
        small_vjp_sm<NORM, LAT ? P::H : KANODE_UNROLL_J>(prm, wsm, y, l, ub, s, nthr, SR::HH, SR::HBAR);
#pragma unroll
        for (int i = 0; i < I; ++i) { s[(SR::Y + i) * nthr] = y[i]; s[(SR::LAM + i) * nthr] = l[i]; dl[i] = -ub[i]; }
        ++nf;
    };
    // Visit every parameter-gradient component of the NS stage slots: fn(j, kv[NS]) with kv[s] = (df/dp)^T lam
    // of stage s at component j (dg/dt = -kv).  Activations are recomputed per unit from the stage records.
    auto for_each_g = [&](auto ns_tag, auto&& fn) {
        constexpr int NS = decltype(ns_tag)::value;
        // layer 1: unit = state component i; a = hbar (H outputs), c = features(y_i); C1 then W1
#pragma unroll 1
        for (int i = 0; i < I; ++i) {
            T c[NS][G + 1];
#pragma unroll
            for (int s = 0; s < NS; ++s) unit_features<NORM>(prm, sm[(s * SR::N + SR::Y + i) * nthr], c[s]);
#pragma unroll 1
            for (int o = 0; o < H; ++o) {
                T av[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) av[s] = sm[(s * SR::N + SR::HBAR + o) * nthr];
#pragma unroll
                for (int q = 0; q <= G; ++q) {
                    T kv[NS];
#pragma unroll
                    for (int s = 0; s < NS; ++s) kv[s] = av[s] * c[s][q];
                    fn(q < G ? P::OC1 + (i * G + q) * H + o : P::OW1 + i * H + o, kv);
                }
            }
        }
        // layer 2: unit = hidden unit i; a = lam (I outputs), c = features(h_i); C2 then W2
        T al[NS][I];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int o = 0; o < I; ++o) al[s][o] = sm[(s * SR::N + SR::LAM + o) * nthr];
#pragma unroll 1
        for (int i = 0; i < H; ++i) {
            T c[NS][G + 1];
#pragma unroll
            for (int s = 0; s < NS; ++s) unit_features<NORM>(prm, sm[(s * SR::N + SR::HH + i) * nthr], c[s]);
#pragma unroll
            for (int q = 0; q <= G; ++q)
#pragma unroll
                for (int o = 0; o < I; ++o) {
                    T kv[NS];
#pragma unroll
                    for (int s = 0; s < NS; ++s) kv[s] = al[s][o] * c[s][q];
                    fn(q < G ? P::OC2 + (i * G + q) * I + o : P::OW2 + i * I + o, kv);
                }
        }
    };

    const double t0 = a.t0, t1 = a.t1, dtmax = fabs(t1 - t0);
    const double dtmin0 = fmax(eps_of(t0), eps_of(t1));
    const T abstol = a.abstol, reltol = a.reltol;
    double t = t1;
    int sp = a.nsave - 1;                        // next preset (save) time, descending
    auto apply_jumps = [&](double tt) {
        bool mod = false;
        while (sp >= 0 && a.saveat[sp] == tt) {
#pragma unroll
            for (int i = 0; i < I; ++i) lam[i] += a.dg[((int64_t)sp * I + i) * B + b];
            --sp; mod = true;
        }
        return mod;
    };
    apply_jumps(t1);                              // PresetTimeCallback fires at init when t_end is a save time
#pragma unroll
    for (int i = 0; i < I; ++i) lprev[i] = lam[i];
    adj_eval(t, lam, kl[0], 0);                   // FSAL
    double dt;                                    // |dt|; integration runs in -t
    {   // ---- initdt on the augmented state (g(T) = 0 so its scale is abstol) ----
        T sk[I], s0 = T(0), s1 = T(0);
#pragma unroll
        for (int i = 0; i < I; ++i) {
            sk[i] = abstol + kabs(lam[i]) * reltol;
            const T x0 = lam[i] / sk[i], x1 = kl[0][i] / sk[i];
            s0 += x0 * x0; s1 += x1 * x1;
        }
        for_each_g(std::integral_constant<int, 1>{}, [&](int, const T (&kv)[1]) {
            const T x = kv[0] / abstol; s1 += x * x;
        });
        const double d0 = sqrt((double)s0 / NZ), d1 = sqrt((double)s1 / NZ);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
        dt0 = fmin(dt0, dtmax);
        T l1[I], f1[I];
#pragma unroll
        for (int i = 0; i < I; ++i) l1[i] = lam[i] - (T)dt0 * kl[0][i];
        adj_eval(t - dt0, l1, f1, 1);
        ++nf;                                     // the package evaluates f0 again; counted like the package does
        T s2 = T(0);
#pragma unroll
        for (int i = 0; i < I; ++i) { const T x = (f1[i] - kl[0][i]) / sk[i]; s2 += x * x; }
        for_each_g(std::integral_constant<int, 2>{}, [&](int, const T (&kv)[2]) {
            const T x = (kv[1] - kv[0]) / abstol; s2 += x * x;
        });
        const double d2 = sqrt((double)s2 / NZ) / dt0;
        const double mx = fmax(d1, d2);
        const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        dt = fmax(dtmin0, fmin(fmin(100.0 * dt0, dt1), dtmax));
    }
    double qold = Ctrl::qoldinit, q11 = 1.0, dtpropose = dt;
    bool accept = false, modified = false;
    int iter = 0, cur = 0;
    while (t > t0) {
        // ---- loopheader! ----
        if (iter > 0) {
            if (!accept) dt = dt / fmin(1.0 / Ctrl::qmin, q11 / Ctrl::gamma);
            else {
                dt = dtpropose;
                if (!modified) {                                   // FSAL: stage 7 of the last step is stage 1
#pragma unroll
                    for (int i = 0; i < I; ++i) kl[0][i] = kl[6][i];
#pragma unroll
                    for (int f = 0; f < SR::N; ++f) sm[f * nthr] = sm[(6 * SR::N + f) * nthr];
                }
            }
        }
        ++iter;
        const double tstop = (sp >= 0) ? fmax(a.saveat[sp], t0) : t0;
        const double dtmin_t = fmax(eps_of(t), dtmin0);
        dt = fmin(fmax(fmin(fabs(dt), dtmax), dtmin_t), t - tstop);
        if (iter > a.maxiters) { ret = RET_MAXITERS; break; }
        if (!(dt > dtmin_t) && (t - dt > tstop || !accept) && iter > 1) { ret = RET_DTMIN; break; }
        if (dt != dt) { ret = RET_UNSTABLE; break; }
        // ---- perform_step! on lambda (registers); g stages exist only as stage records ----
        const T h = (T)(-dt);
        T lnew[I];
#pragma unroll
        for (int i = 0; i < I; ++i) lnew[i] = lprev[i];
        // after a jump the FSAL stage is re-evaluated: that is stage 0 of the same loop (row 0 of the tableau is
        // zero, c[0] = 0), so the hot loop contains exactly one copy of the fused forward+VJP code
#pragma unroll 1
        for (int s = modified ? 0 : 1; s < 7; ++s) {
            T ls[I], ks[I];
#pragma unroll
            for (int i = 0; i < I; ++i) {
                T acc = T(0);
#pragma unroll
                for (int j = 0; j < 6; ++j) acc += Tab<T>::a(s, j) * kl[j][i];
                ls[i] = lprev[i] + h * acc;
            }
            adj_eval(t - tab_c(s) * dt, ls, ks, s);
#pragma unroll
            for (int j = 0; j < 7; ++j)
                if (j == s) {
#pragma unroll
                    for (int i = 0; i < I; ++i) kl[j][i] = ks[i];
                }
            if (s == 6) {
#pragma unroll
                for (int i = 0; i < I; ++i) lnew[i] = ls[i];
            }
        }
        modified = false;
        // ---- error estimate over all I + NP components; g_new goes to the other buffer ----
        T es = T(0);
        bool bad = false;
#pragma unroll
        for (int i = 0; i < I; ++i) {
            T ut = T(0);
#pragma unroll
            for (int j = 0; j < 7; ++j) ut += Tab<T>::bt(j) * kl[j][i];
            ut *= h;
            const T sc = abstol + kmax(kabs(lprev[i]), kabs(lnew[i])) * reltol;
            const T r = ut / sc;
            es += r * r;
            bad |= (lnew[i] != lnew[i]);
        }
        {   // step-end pass over the NP gradient components: g1 = g0 - h*sum_s b_s kv_s, error term with btilde.
            // Per unit: prefetch its g values, then a ROLLED loop over the 7 stages recomputes the unit's features
            // from the stage record and accumulates (small code: the hot loop must stay inside the I-cache).
            const T* gold = gbuf + (int64_t)cur * NP * B;
            T* gnew = gbuf + (int64_t)(cur ^ 1) * NP * B;
            const T mh = -h;                                               // dg/dt = -kv
            auto finalize = [&](int j, T g0, T vb, T vt) {
                const T g1 = g0 + mh * vb;
                const T sc = abstol + kmax(kabs(g0), kabs(g1)) * reltol;
                const T r = kdiv(mh * vt, sc);
                es += r * r;
                gnew[(int64_t)j * B] = g1;
            };
            // layer 2: unit = hidden unit i, outputs o < I, features q <= G (C2 rows then the W2 row)
#pragma unroll 1
            for (int i = 0; i < H; ++i) {
                T g0[G + 1][I], vb[G + 1][I], vt[G + 1][I];
#pragma unroll
                for (int q = 0; q <= G; ++q)
#pragma unroll
                    for (int o = 0; o < I; ++o) {
                        const int j = q < G ? P::OC2 + (i * G + q) * I + o : P::OW2 + i * I + o;
                        g0[q][o] = gold[(int64_t)j * B]; vb[q][o] = T(0); vt[q][o] = T(0);
                    }
#pragma unroll US
                for (int s = 0; s < 7; ++s) {
                    const T* rec = sm + s * SR::N * nthr;
                    T c[G + 1];
                    unit_features<NORM>(prm, rec[(SR::HH + i) * nthr], c);
                    const T wb = Tab<T>::b(s), wt = Tab<T>::bt(s);
#pragma unroll
                    for (int o = 0; o < I; ++o) {
                        const T l = rec[(SR::LAM + o) * nthr];
                        const T ab = wb * l, at = wt * l;
#pragma unroll
                        for (int q = 0; q <= G; ++q) kfma2b(vb[q][o], vt[q][o], ab, at, c[q]);     // (vb, vt) += (ab, at) * c: one FFMA2
                    }
                }
#pragma unroll
                for (int q = 0; q <= G; ++q)
#pragma unroll
                    for (int o = 0; o < I; ++o)
                        finalize(q < G ? P::OC2 + (i * G + q) * I + o : P::OW2 + i * I + o, g0[q][o], vb[q][o], vt[q][o]);
            }
            // layer 1: unit = state component i, outputs o < H in chunks of OC, features q <= G (C1 rows, W1 row)
            constexpr int OC = (H % 5 == 0) ? 5 : (H % 4 == 0 ? 4 : (H % 2 == 0 ? 2 : 1));
#pragma unroll 1
            for (int io = 0; io < I * (H / OC); ++io) {
                const int i = io / (H / OC), o0 = (io % (H / OC)) * OC;
                T g0[G + 1][OC], vb[G + 1][OC], vt[G + 1][OC];
#pragma unroll
                for (int q = 0; q <= G; ++q)
#pragma unroll
                    for (int oo = 0; oo < OC; ++oo) {
                        const int j = (q < G ? P::OC1 + (i * G + q) * H : P::OW1 + i * H) + o0 + oo;
                        g0[q][oo] = gold[(int64_t)j * B]; vb[q][oo] = T(0); vt[q][oo] = T(0);
                    }
#pragma unroll US
                for (int s = 0; s < 7; ++s) {
                    const T* rec = sm + s * SR::N * nthr;
                    T c[G + 1];
                    unit_features<NORM>(prm, rec[(SR::Y + i) * nthr], c);
                    const T wb = Tab<T>::b(s), wt = Tab<T>::bt(s);
#pragma unroll
                    for (int oo = 0; oo < OC; ++oo) {
                        const T hb = rec[(SR::HBAR + o0 + oo) * nthr];
                        const T ab = wb * hb, at = wt * hb;
#pragma unroll
                        for (int q = 0; q <= G; ++q) kfma2b(vb[q][oo], vt[q][oo], ab, at, c[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q <= G; ++q)
#pragma unroll
                    for (int oo = 0; oo < OC; ++oo)
                        finalize((q < G ? P::OC1 + (i * G + q) * H : P::OW1 + i * H) + o0 + oo, g0[q][oo], vb[q][oo], vt[q][oo]);
            }
        }
        const double EEst = (double)ksqrt(es / T(NZ));
        if (EEst != EEst || bad) { ret = RET_UNSTABLE; break; }
        const double q = pi_q(EEst, qold, q11);
        accept = EEst <= 1.0;
        if (accept) {
            ++naccept;
            qold = fmax(EEst, Ctrl::qoldinit);
            const double dtnew = dt / q;
            double tnew = t - dt;
            if (fabs(tnew - tstop) < 100.0 * eps_of(fmax(fabs(t), fabs(tstop)))) tnew = tstop;
            dtpropose = fmax(fmin(dtmax, fabs(dtnew)), fmax(eps_of(tnew), dtmin0));
            t = tnew;
            cur ^= 1;
#pragma unroll
            for (int i = 0; i < I; ++i) lam[i] = lnew[i];
            modified = apply_jumps(t);
#pragma unroll
            for (int i = 0; i < I; ++i) lprev[i] = lam[i];
        } else {
            ++nreject;
        }
    }
    if (cur == 1) {   // result always in buffer 0
#pragma unroll 1
        for (int j = 0; j < NP; ++j) gbuf[(int64_t)j * B] = gbuf[((int64_t)NP + j) * B];
    }
    if (ret != RET_SUCCESS) {
#pragma unroll 1
        for (int j = 0; j < NP; ++j) gbuf[(int64_t)j * B] = T(0);
    }
    if (a.du0)
#pragma unroll
        for (int i = 0; i < I; ++i) a.du0[b * I + i] = lam[i];
    if (a.stats) a.stats[b] = kanode_stats{naccept, nreject, nf, ret};
    if (a.attempts) {
        a.attempts[b] = naccept + nreject;
        if (a.attempts_sum) atomicAdd(a.attempts_sum, (unsigned long long)(naccept + nreject));
    }
}

// Prediction for the NEXT backward solve from the step attempts of the last one (training steps repeat with slowly
// changing parameters, so the same trajectories are the long ones): a trajectory with more than mean+4 attempts is
// appended to long_list (up to `cap` of them) and flagged so that the bulk launch skips it.
// sched[0] = sum of attempts of the previous call, sched[1] = entries in long_list (may exceed cap; clamp when used).
static __global__ void __launch_bounds__(256) mark_long_kernel(const int* __restrict__ attempts, int64_t B, const unsigned long long* sched_sum,
                                                        int* long_count, int cap, int* __restrict__ long_list,
                                                        unsigned char* __restrict__ long_flag) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int thr = (int)(*sched_sum / (unsigned long long)B) + 4;
    unsigned char f = 0;
    if (attempts[b] > thr) {
        const int slot = atomicAdd(long_count, 1);
        if (slot < cap) { long_list[slot] = (int)b; f = 1; }
    }
    long_flag[b] = f;
}
static __global__ void clamp_count_kernel(int* c, int cap) { if (*c > cap) *c = cap; }

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
