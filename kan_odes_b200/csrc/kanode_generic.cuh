// kanode_generic.cuh — block-per-trajectory kernels for arbitrary KDense chains (PDE surrogates [n,10,n], any
// normalizer / basis / layer count) and for the hidden-source model (periodic Laplacian + pointwise KAN).
//
// Mapping: ONE THREAD BLOCK PER TRAJECTORY (initial condition).  State and stage vectors of length n live in a
// per-trajectory global workspace (L2-resident for the BASELINE sizes) and are walked with coalesced, block-strided
// loops; the narrow hidden vectors (width <= 1024) live in shared memory; the RBF features are produced and consumed
// on the fly (registers / shared memory) and never written to HBM.  A layer is evaluated with one of two mappings:
//   reduce  (I >= O, e.g. [n -> 10]): thread per input unit, 16 register accumulators, deterministic block reduction
//   expand  (O >  I, e.g. [10 -> n]): features of all inputs in shared memory, thread per output, coalesced weight rows
// The backward kernel integrates z = [lambda; g] like the reference's InterpolatingAdjoint.  dg/dt is never
// materialised: each stage stores the layer inputs x_l and output cotangents ybar_l (a "stage record", 2n+2H floats
// for [n,H,n]); the step-end pass rebuilds sum_s b_s ybar_s[o] * feature_s[i][q] per parameter.
//
// Reference semantics as in kanode_small.cuh / oracle/kanode_oracle.cpp:
//   KDense forward  Lotka-Volterra/src/kdense.jl:109-130;  reverse rules  Lotka-Volterra/src/utils.jl:15-21,36-43,56-62
//   surrogate drivers  "PDE examples/Burgers_Surrogate.jl":82-107,  Schrodinger_Surrogate.jl:89-114
//   source drivers     "PDE examples/Allen-Cahn_Source.jl":50-54,90-99,  Fisher-KPP_Source.jl:55-59,95-104
#pragma once
#include "kanode_host.h"
#include "kanode_math.cuh"

namespace kanode {

constexpr int GEN_BT = 256;      // threads per block
constexpr int GEN_ACT = 1024;    // widest intermediate layer interface kept in shared memory
constexpr int GEN_FEAT = 2048;   // feature scratch (elements)
constexpr int GEN_ACC = 16;      // register accumulators per thread in the reduce mappings
constexpr int GEN_PW = 8;        // widest layer of a pointwise (source-term) chain

template <class T> struct GSmem {
    T red[32 * GEN_ACC];
    T res[GEN_ACC];
    T act[2][GEN_ACT];
    T feat[GEN_FEAT];
};

// deterministic block reduction of NV values; results land in sm.res[0..NV) (and are returned in v)
template <class T, int NV> __device__ __forceinline__ void block_reduce(T (&v)[NV], GSmem<T>& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();                       // previous readers of red/res are done
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        if (lane == 0) sm.red[warp * GEN_ACC + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        T s = T(0);
        for (int w = 0; w < nw; ++w) s += sm.red[w * GEN_ACC + threadIdx.x];
        sm.res[threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = sm.res[k];
}
template <class T> __device__ __forceinline__ T block_sum(T x, GSmem<T>& sm) {
    T v[1] = {x};
    block_reduce<T, 1>(v, sm);
    return v[0];
}

// base-branch activation of a layer input: swish for KDense (kdense.jl:123); identity / tanh for Dense layers, where the output
// activation of the layer before acts on this layer's inputs (Lux.Dense(2 => 50, tanh), LV_driver_MLP.jl:61)
enum { ACT_SWISH = 0, ACT_IDENTITY = 1, ACT_TANH = 2 };
template <class T> __device__ __forceinline__ void base_fwd(int act, T x, T& s) {
    if (act == ACT_SWISH) swish_fwd(x, s);
    else if (act == ACT_IDENTITY) s = x;
    else s = ktanh(x);
}
template <class T> __device__ __forceinline__ void base_both(int act, T x, T& s, T& ds) {
    if (act == ACT_SWISH) swish_both(x, s, ds);
    else if (act == ACT_IDENTITY) { s = x; ds = T(1); }
    else { s = ktanh(x); ds = T(1) - s * s; }
}
// inputs of the feature x weight loops: the I units plus, for a Dense layer, the virtual bias unit (feature 1)
__device__ __forceinline__ int g_inputs(const GenericLayer& L) { return L.I + L.bias; }

// features of one input unit: c[q] = basis_q(x) for q < G, c[G] = base activation of x (0 without the base branch)
template <class T>
__device__ __forceinline__ void g_features(const GenericLayer& L, const float* grid, T xi, T* c) {
    if (L.G > 0) {
        const T xn = normalize_rt(L.norm, xi);
        const T inv_h = (T)L.inv_h;
        for (int g = 0; g < L.G; ++g) c[g] = basis_val(L.basis, (xn - (T)grid[g]) * inv_h);
    }
    T s = T(0);
    if (L.use_base) base_fwd(L.act, xi, s);
    c[L.G] = s;
}
// ... of unit i of the layer input vector x (i == L.I: the bias unit, which has no entry in x)
template <class T>
__device__ __forceinline__ void g_features_at(const GenericLayer& L, const float* grid, const T* x, int i, T* c) {
    if (i >= L.I) { for (int g = 0; g < L.G; ++g) c[g] = T(0); c[L.G] = T(1); return; }
    g_features(L, grid, x[i], c);
}

// ---------------------------------------------------------------------------------------------------------
// one KDense layer, forward:  y[O] = C * basis(norm(x)) + W * swish(x)           (kdense.jl:113-124)
// ---------------------------------------------------------------------------------------------------------
template <class T>
__device__ void g_layer_forward(const GenericModel& m, int l, const T* __restrict__ p, const T* x, T* y, GSmem<T>& sm) {
    const GenericLayer& L = m.L[l];
    const float* grid = m.grid + L.goff;
    const int I = L.I, O = L.O, G = L.G, GP = L.G + 1;
    const T* C = p + L.offC;
    const T* W = p + L.offW;
    const int IB = g_inputs(L);
    if (O > I && (long long)IB * GP <= GEN_FEAT) {
        __syncthreads();
        for (int i = threadIdx.x; i < IB; i += blockDim.x) g_features_at(L, grid, x, i, sm.feat + i * GP);
        __syncthreads();
        for (int o = threadIdx.x; o < O; o += blockDim.x) {
            T acc = T(0);
            for (int i = 0; i < IB; ++i) {
                const T* f = sm.feat + i * GP;
                for (int g = 0; g < G; ++g) acc += C[((long long)i * G + g) * O + o] * f[g];
                if (L.use_base) acc += W[(long long)i * O + o] * f[G];
            }
            y[o] = acc;
        }
        __syncthreads();
        return;
    }
    const T inv_h = (T)L.inv_h;
    for (int o0 = 0; o0 < O; o0 += GEN_ACC) {
        const int oc = min(GEN_ACC, O - o0);
        T acc[GEN_ACC];
#pragma unroll
        for (int k = 0; k < GEN_ACC; ++k) acc[k] = T(0);
        for (int i = threadIdx.x; i < IB; i += blockDim.x) {
            const T xi = i < I ? x[i] : T(0);
            if (G > 0 && i < I) {
                const T xn = normalize_rt(L.norm, xi);
                for (int g = 0; g < G; ++g) {
                    const T b = basis_val(L.basis, (xn - (T)grid[g]) * inv_h);
                    const T* col = C + ((long long)i * G + g) * O + o0;
#pragma unroll
                    for (int k = 0; k < GEN_ACC; ++k) if (k < oc) acc[k] += col[k] * b;
                }
            }
            if (L.use_base) {
                T s = T(1);                                            // the bias unit's feature
                if (i < I) base_fwd(L.act, xi, s);
                const T* col = W + (long long)i * O + o0;
#pragma unroll
                for (int k = 0; k < GEN_ACC; ++k) if (k < oc) acc[k] += col[k] * s;
            }
        }
        block_reduce<T, GEN_ACC>(acc, sm);
        if (threadIdx.x < oc) y[o0 + threadIdx.x] = sm.res[threadIdx.x];
    }
    __syncthreads();
}

// chain forward.  rec != null: store every layer's input x_l into the stage record and stop before the last
// layer's contraction (the adjoint does not need f(y)).
template <class T>
__device__ void g_chain_forward(const GenericModel& m, const T* p, const T* x0, T* yout, GSmem<T>& sm, T* rec) {
    const T* x = x0;
    for (int l = 0; l < m.n_layers; ++l) {
        const GenericLayer& L = m.L[l];
        if (rec) {
            for (int i = threadIdx.x; i < L.I; i += blockDim.x) rec[L.rx + i] = x[i];
            if (l == m.n_layers - 1) break;
        }
        T* y = (l == m.n_layers - 1) ? yout : sm.act[l & 1];
        g_layer_forward(m, l, p, x, y, sm);
        x = y;
    }
    __syncthreads();
}

// one layer, reverse: xbar[I] = (dy/dx)^T ybar      (rrule(_rbf) utils.jl:15-21 + activation rules)
template <class T>
__device__ void g_layer_reverse(const GenericModel& m, int l, const T* __restrict__ p, const T* x, const T* ybar, T* xbar,
                                GSmem<T>& sm) {
    const GenericLayer& L = m.L[l];
    const float* grid = m.grid + L.goff;
    const int I = L.I, O = L.O, G = L.G, GP = L.G + 1;
    const T* C = p + L.offC;
    const T* W = p + L.offW;
    const T inv_h = (T)L.inv_h;
    if (I >= O || GP > GEN_ACC) {               // thread per input unit
        for (int i = threadIdx.x; i < I; i += blockDim.x) {
            const T xi = x[i];
            const T xn = normalize_rt(L.norm, xi);
            T xnbar = T(0);
            for (int g = 0; g < G; ++g) {
                T b, db; basis_both(L.basis, (xn - (T)grid[g]) * inv_h, b, db);
                const T* col = C + ((long long)i * G + g) * O;
                T bbar = T(0);
                for (int o = 0; o < O; ++o) bbar += col[o] * ybar[o];
                xnbar += db * inv_h * bbar;
            }
            T xb = xnbar * normalize_deriv_rt(L.norm, xn);
            if (L.use_base) {
                T s, ds; base_both(L.act, xi, s, ds);
                const T* col = W + (long long)i * O;
                T sbar = T(0);
                for (int o = 0; o < O; ++o) sbar += col[o] * ybar[o];
                xb += sbar * ds;
            }
            xbar[i] = xb;
        }
        __syncthreads();
        return;
    }
    // O > I: chunks of inputs; each thread accumulates over its outputs, then a block reduction per chunk
    const int ic = GEN_ACC / GP;
    for (int i0 = 0; i0 < I; i0 += ic) {
        const T* row[GEN_ACC];
        T acc[GEN_ACC];
#pragma unroll
        for (int k = 0; k < GEN_ACC; ++k) {
            const int ii = k / GP, q = k - ii * GP;
            acc[k] = T(0);
            row[k] = nullptr;
            if (ii < ic && i0 + ii < I) {
                if (q < G) row[k] = C + ((long long)(i0 + ii) * G + q) * O;
                else if (L.use_base) row[k] = W + (long long)(i0 + ii) * O;
            }
        }
        for (int o = threadIdx.x; o < O; o += blockDim.x) {
            const T yb = ybar[o];
#pragma unroll
            for (int k = 0; k < GEN_ACC; ++k) if (row[k]) acc[k] += row[k][o] * yb;
        }
        block_reduce<T, GEN_ACC>(acc, sm);
        if (threadIdx.x < ic && i0 + threadIdx.x < I) {
            const int i = i0 + threadIdx.x;
            const T xi = x[i];
            const T xn = normalize_rt(L.norm, xi);
            T xnbar = T(0);
            for (int g = 0; g < G; ++g) {
                T b, db; basis_both(L.basis, (xn - (T)grid[g]) * inv_h, b, db);
                xnbar += db * inv_h * sm.res[threadIdx.x * GP + g];
            }
            T xb = xnbar * normalize_deriv_rt(L.norm, xn);
            if (L.use_base) { T s, ds; base_both(L.act, xi, s, ds); xb += sm.res[threadIdx.x * GP + G] * ds; }
            xbar[i] = xb;
        }
    }
    __syncthreads();
}

// fused forward-recompute + VJP of the chain: ubar = (df/du)^T lam; the stage record receives x_l and ybar_l
template <class T>
__device__ void g_chain_vjp(const GenericModel& m, const T* p, const T* y, const T* lam, T* ubar, GSmem<T>& sm, T* rec) {
    g_chain_forward<T>(m, p, y, nullptr, sm, rec);
    const T* yb = lam;
    for (int l = m.n_layers - 1; l >= 0; --l) {
        const GenericLayer& L = m.L[l];
        for (int o = threadIdx.x; o < L.O; o += blockDim.x) rec[L.ry + o] = yb[o];
        __syncthreads();
        T* xb = (l == 0) ? ubar : sm.act[l & 1];
        g_layer_reverse<T>(m, l, p, rec + L.rx, rec + L.ry, xb, sm);
        yb = xb;
    }
}

// ---------------------------------------------------------------------------------------------------------
// pointwise chain for the hidden-source model: kan1_.(u)  (Allen-Cahn_Source.jl:91) evaluated per grid node
// ---------------------------------------------------------------------------------------------------------
template <class T> __device__ T pw_forward(const GenericModel& m, const T* __restrict__ p, T u, T* xs /*[nl][GEN_PW] or null*/) {
    T cur[GEN_PW], nxt[GEN_PW];
    cur[0] = u;
    for (int l = 0; l < m.n_layers; ++l) {
        const GenericLayer& L = m.L[l];
        const float* grid = m.grid + L.goff;
        for (int o = 0; o < L.O; ++o) nxt[o] = T(0);
        for (int i = 0; i < L.I; ++i) {
            if (xs) xs[l * GEN_PW + i] = cur[i];
            const T xn = normalize_rt(L.norm, cur[i]);
            for (int g = 0; g < L.G; ++g) {
                const T b = basis_val(L.basis, (xn - (T)grid[g]) * (T)L.inv_h);
                for (int o = 0; o < L.O; ++o) nxt[o] += p[L.offC + ((long long)i * L.G + g) * L.O + o] * b;
            }
            if (L.use_base) { T s; swish_fwd(cur[i], s); for (int o = 0; o < L.O; ++o) nxt[o] += p[L.offW + (long long)i * L.O + o] * s; }
        }
        for (int o = 0; o < L.O; ++o) cur[o] = nxt[o];
    }
    return cur[0];
}
// reverse of pw_forward for output cotangent yb: returns du-bar, accumulates pbar[j] += w * (dkan/dp_j) via fn(j, v)
template <class T, class Fn> __device__ T pw_reverse(const GenericModel& m, const T* __restrict__ p, const T* xs, T yb, Fn&& fn) {
    T cur[GEN_PW], nxt[GEN_PW];
    cur[0] = yb;
    for (int l = m.n_layers - 1; l >= 0; --l) {
        const GenericLayer& L = m.L[l];
        const float* grid = m.grid + L.goff;
        for (int i = 0; i < L.I; ++i) {
            const T xi = xs[l * GEN_PW + i];
            const T xn = normalize_rt(L.norm, xi);
            T xnbar = T(0);
            for (int g = 0; g < L.G; ++g) {
                T b, db; basis_both(L.basis, (xn - (T)grid[g]) * (T)L.inv_h, b, db);
                T bbar = T(0);
                for (int o = 0; o < L.O; ++o) {
                    const long long j = L.offC + ((long long)i * L.G + g) * L.O + o;
                    bbar += p[j] * cur[o];
                    fn(j, cur[o] * b);
                }
                xnbar += db * (T)L.inv_h * bbar;
            }
            T xb = xnbar * normalize_deriv_rt(L.norm, xn);
            if (L.use_base) {
                T s, ds; swish_both(xi, s, ds);
                T sbar = T(0);
                for (int o = 0; o < L.O; ++o) { const long long j = L.offW + (long long)i * L.O + o; sbar += p[j] * cur[o]; fn(j, cur[o] * s); }
                xb += sbar * ds;
            }
            nxt[i] = xb;
        }
        for (int i = 0; i < L.I; ++i) cur[i] = nxt[i];
    }
    return cur[0];
}

// du = f(u) for either rhs kind (block-wide)
template <class T>
__device__ void g_rhs(const GenericModel& m, const T* p, const T* u, T* du, GSmem<T>& sm) {
    if (m.rhs_kind != KANODE_RHS_SOURCE_LAPLACIAN) { g_chain_forward<T>(m, p, u, du, sm, nullptr); return; }
    const int n = m.n;
    const T ls = (T)m.lap_scale;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const T um = u[(j + n - 1) % n], up = u[(j + 1) % n], uj = u[j];        // periodic corners AC_Source:53-54
        du[j] = ls * (um - T(2) * uj + up) + pw_forward<T>(m, p, uj, nullptr);   // AC_Source:92
    }
    __syncthreads();
}

// ubar = (df/du)^T lam and the parameter part: chain -> stage record; source -> kg[np] (block-reduced, small np)
template <class T>
__device__ void g_vjp(const GenericModel& m, const T* p, const T* y, const T* lam, T* ubar, GSmem<T>& sm, T* rec) {
    if (m.rhs_kind != KANODE_RHS_SOURCE_LAPLACIAN) { g_chain_vjp<T>(m, p, y, lam, ubar, sm, rec); return; }
    // source model: rec holds kg[np] = sum_j lam_j dkan(y_j)/dp ; np <= GEN_FEAT
    const int n = m.n, np = (int)m.np;
    const T ls = (T)m.lap_scale;
    if (np <= GEN_ACC) {
        // the reference's shape (1 -> 1, np = G + 1 = 11): every thread accumulates its nodes' contributions privately and the
        // block adds them once, in a fixed order (shared-memory atomics on 11 addresses serialised the whole block)
        T acc[GEN_ACC];
#pragma unroll
        for (int k = 0; k < GEN_ACC; ++k) acc[k] = T(0);
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            T xs[KANODE_MAX_LAYERS * GEN_PW];
            pw_forward<T>(m, p, y[j], xs);
            const T lj = lam[j];
            const T xb = pw_reverse<T>(m, p, xs, lj, [&](long long jj, T v) {
#pragma unroll
                for (int k = 0; k < GEN_ACC; ++k) if (k == (int)jj) acc[k] += v;
            });
            const T lm = lam[(j + n - 1) % n], lp = lam[(j + 1) % n];               // the Laplacian is symmetric
            ubar[j] = ls * (lm - T(2) * lj + lp) + xb;
        }
        block_reduce<T, GEN_ACC>(acc, sm);
        if (threadIdx.x < np) rec[threadIdx.x] = sm.res[threadIdx.x];
        __syncthreads();
        return;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < np; j += blockDim.x) sm.feat[j] = T(0);
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        T xs[KANODE_MAX_LAYERS * GEN_PW];
        pw_forward<T>(m, p, y[j], xs);
        const T lj = lam[j];
        const T xb = pw_reverse<T>(m, p, xs, lj, [&](long long jj, T v) { atomicAdd(&sm.feat[jj], v); });
        const T lm = lam[(j + n - 1) % n], lp = lam[(j + 1) % n];               // the Laplacian is symmetric
        ubar[j] = ls * (lm - T(2) * lj + lp) + xb;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < np; j += blockDim.x) rec[j] = sm.feat[j];
    __syncthreads();
}

// per-edge activations of layer l (LV/Activation_getter.jl:22-31,44-54): thread per (sample k, input i);
// act[k][i][o] = sum_g C[o,(i,g)] basis_g(norm(x_i)) + W[o,i] swish(x_i) — the fused basis expansion without the sum over i
template <class T>
__global__ void __launch_bounds__(128) edge_activation_kernel(const __grid_constant__ GenericModel m, int l, const T* __restrict__ p,
                                                              const T* __restrict__ x, T* __restrict__ act, int64_t K) {
    const GenericLayer& L = m.L[l];
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * L.I) return;
    const int i = (int)(idx % L.I);
    T c[GEN_MAX_G + 1];
    g_features(L, m.grid + L.goff, x[idx], c);
    T* a = act + idx * L.O;
    for (int o = 0; o < L.O; ++o) {
        T v = T(0);
        for (int g = 0; g < L.G; ++g) v += p[L.offC + ((long long)i * L.G + g) * L.O + o] * c[g];
        if (L.use_base) v += p[L.offW + (long long)i * L.O + o] * c[L.G];
        a[o] = v;
    }
}

// reg_loss(p) = act_reg * sum|p| + entropy_reg * (-sum e log e), e = |p| / sum|p|  (LV_driver_KANODE.jl:187-194), three passes:
// acc[0] = S = sum|p|;  acc[1] = sum |p| log|p| (so that -sum e log e = log S - acc[1]/S);  then the gradient
//   d reg / d p_j = sign(p_j) * (act_reg - entropy_reg * (log e_j + E) / S)
template <class T> __global__ void __launch_bounds__(256) reg_sums_kernel(const T* __restrict__ p, size_t n, double* __restrict__ acc) {
    double s = 0.0, sl = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double a = fabs((double)p[i]);
        s += a; if (a > 0.0) sl += a * log(a);
    }
    for (int d = 16; d > 0; d >>= 1) { s += __shfl_down_sync(0xffffffffu, s, d); sl += __shfl_down_sync(0xffffffffu, sl, d); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&acc[0], s); atomicAdd(&acc[1], sl); }
}
// loss_out (optional): += loss_scale * reg;  grad (optional): += grad_scale * d reg/d p
template <class T>
__global__ void __launch_bounds__(256) reg_apply_kernel(const T* __restrict__ p, size_t n, const double* __restrict__ acc, double act_reg,
                                                        double ent_reg, double* loss_out, double loss_scale, T* __restrict__ grad, double grad_scale) {
    const double S = acc[0], E = S > 0.0 ? log(S) - acc[1] / S : 0.0;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && loss_out) *loss_out += loss_scale * (S * act_reg + E * ent_reg);
    if (i >= n || !grad) return;
    const double v = (double)p[i], a = fabs(v);
    if (a == 0.0) return;
    const double dE = -(log(a / S) + E) / S;
    grad[i] += (T)(grad_scale * (v > 0.0 ? 1.0 : -1.0) * (act_reg + ent_reg * dE));
}

// ---------------------------------------------------------------------------------------------------------
// argument blocks (device pointers); per-trajectory workspace slices are carved inside the kernels
// ---------------------------------------------------------------------------------------------------------
template <class T> struct GenFwdArgs {
    const T* p; const T* u0; int64_t B;
    double t0, t1; const double* saveat; int nsave;
    T abstol, reltol; int maxiters;
    T* out; kanode_stats* stats;
    T* work;                 // [B][10n]: u, uprev, tmp, k1..k7
    // dense record
    double* rec_t; T* rec_dt; T* rec; int cap; int* nsteps; int* retcode;   // rec: [B][cap][8n] = uprev, k1..k7
    const T* target; T* dg /*[B][nsave][n]*/; double* loss_sum;
};
template <class T> struct GenBwdArgs {
    const T* p; int64_t B;
    double t0, t1; const double* saveat; int nsave;
    T abstol, reltol; int maxiters;
    const double* rec_t; const T* rec_dt; const T* rec; int cap; const int* nsteps; const int* retcode;
    const T* dg;
    T* work;                 // [B][(11n + 7*rec_len)]: lam, lprev, ls, y, k1..k7 (7n), stage records
    T* g;                    // [B][2][np]
    T* du0; kanode_stats* stats;
};

template <class T> __global__ void __launch_bounds__(GEN_BT) generic_rhs_kernel(const __grid_constant__ GenericModel m, const T* p, const T* u, T* du) {
    __shared__ GSmem<T> sm;
    g_rhs<T>(m, p, u + (int64_t)blockIdx.x * m.n, du + (int64_t)blockIdx.x * m.n_out, sm);
}

// ubar per sample; pbar accumulated over the batch with atomics (utility entry point, not the training path)
template <class T>
__global__ void __launch_bounds__(GEN_BT) generic_vjp_kernel(const __grid_constant__ GenericModel m, const T* p, const T* u, const T* lam,
                                                              T* ubar, T* pbar, T* recs) {
    __shared__ GSmem<T> sm;
    const int64_t b = blockIdx.x;
    T* rec = recs + b * m.rec_len;
    g_vjp<T>(m, p, u + b * m.n, lam + b * m.n_out, ubar + b * m.n, sm, rec);
    if (m.rhs_kind == KANODE_RHS_SOURCE_LAPLACIAN) {
        for (int j = threadIdx.x; j < (int)m.np; j += blockDim.x) atomicAdd(&pbar[j], rec[j]);
        return;
    }
    for (int l = 0; l < m.n_layers; ++l) {
        const GenericLayer& L = m.L[l];
        const float* grid = m.grid + L.goff;
        for (int i = threadIdx.x; i < g_inputs(L); i += blockDim.x) {
            T c[GEN_MAX_G + 1];
            g_features_at(L, grid, rec + L.rx, i, c);
            for (int o = 0; o < L.O; ++o) {
                const T a = rec[L.ry + o];
                for (int g = 0; g < L.G; ++g) atomicAdd(&pbar[L.offC + ((long long)i * L.G + g) * L.O + o], a * c[g]);
                if (L.use_base) atomicAdd(&pbar[L.offW + (long long)i * L.O + o], a * c[L.G]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// forward: adaptive Tsit5, one block per trajectory
// ---------------------------------------------------------------------------------------------------------
template <class T, bool DENSE>
__global__ void __launch_bounds__(GEN_BT) generic_forward_kernel(const __grid_constant__ GenericModel m, const GenFwdArgs<T> a) {
    __shared__ GSmem<T> sm;
    const int64_t b = blockIdx.x;
    const int n = m.n, tid = threadIdx.x, bt = blockDim.x;
    T* u = a.work + b * 10 * (int64_t)n;
    T* uprev = u + n; T* tmp = uprev + n; T* k = tmp + n;            // k[j] at k + j*n
    const T* p = a.p;
    for (int i = tid; i < n; i += bt) { const T v = a.u0[b * n + i]; u[i] = v; uprev[i] = v; }
    for (int i = tid; i < 7 * n; i += bt) k[i] = T(0);
    __syncthreads();
    g_rhs<T>(m, p, u, k, sm);
    int nf = 1, naccept = 0, nreject = 0, ret = RET_SUCCESS;
    const double t0 = a.t0, t1 = a.t1, dtmax = fabs(t1 - t0), dtmin0 = fmax(eps_of(t0), eps_of(t1));
    const T abstol = a.abstol, reltol = a.reltol;
    double t = t0, dt;
    {   // ---- initdt ----
        T v[2] = {T(0), T(0)};
        for (int i = tid; i < n; i += bt) {
            const T sk = abstol + kabs(u[i]) * reltol;
            const T x0 = u[i] / sk, x1 = k[i] / sk;
            v[0] += x0 * x0; v[1] += x1 * x1;
        }
        block_reduce<T, 2>(v, sm);
        const double d0 = sqrt((double)v[0] / n), d1 = sqrt((double)v[1] / n);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
        dt0 = fmin(dt0, dtmax);
        for (int i = tid; i < n; i += bt) tmp[i] = u[i] + (T)dt0 * k[i];
        __syncthreads();
        g_rhs<T>(m, p, tmp, k + n, sm);
        nf += 2;
        T s2 = T(0);
        for (int i = tid; i < n; i += bt) { const T sk = abstol + kabs(u[i]) * reltol; const T x = (k[n + i] - k[i]) / sk; s2 += x * x; }
        s2 = block_sum<T>(s2, sm);
        const double d2 = sqrt((double)s2 / n) / dt0, mx = fmax(d1, d2);
        const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        dt = fmax(dtmin0, fmin(fmin(100.0 * dt0, dt1), dtmax));
        __syncthreads();
        for (int i = tid; i < n; i += bt) k[n + i] = T(0);
        __syncthreads();
    }
    double qold = Ctrl::qoldinit, q11 = 1.0, dtpropose = dt, lsum = 0.0;
    bool accept = false;
    int iter = 0, sidx = 0, nrec = 0;
    if (t0 == t1 && a.out)
        for (; sidx < a.nsave; ++sidx) for (int i = tid; i < n; i += bt) a.out[(b * a.nsave + sidx) * n + i] = u[i];
    while (t < t1) {
        if (iter > 0) { if (!accept) dt = dt / fmin(1.0 / Ctrl::qmin, q11 / Ctrl::gamma); else dt = dtpropose; }
        ++iter;
        const double dtmin_t = fmax(eps_of(t), dtmin0);
        dt = fmin(fmax(fmin(fabs(dt), dtmax), dtmin_t), t1 - t);
        if (iter > a.maxiters) { ret = RET_MAXITERS; break; }
        if (!(dt > dtmin_t) && (t + dt < t1 || !accept) && iter > 1) { ret = RET_DTMIN; break; }
        if (dt != dt) { ret = RET_UNSTABLE; break; }
        const T h = (T)dt;
        for (int s = 1; s < 7; ++s) {
            T as[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) as[j] = Tab<T>::a(s, j);
            for (int i = tid; i < n; i += bt) {
                T acc = T(0);
#pragma unroll
                for (int j = 0; j < 6; ++j) acc += as[j] * k[(int64_t)j * n + i];
                tmp[i] = uprev[i] + h * acc;
            }
            __syncthreads();
            g_rhs<T>(m, p, tmp, k + (int64_t)s * n, sm);
        }
        nf += 6;                                                     // tmp now holds u_new (stage 7 input)
        T es = T(0);
        for (int i = tid; i < n; i += bt) {
            T ut = T(0);
#pragma unroll
            for (int j = 0; j < 7; ++j) ut += Tab<T>::bt(j) * k[(int64_t)j * n + i];
            const T sc = abstol + kmax(kabs(uprev[i]), kabs(tmp[i])) * reltol;
            const T r = (h * ut) / sc;
            es += r * r;
        }
        es = block_sum<T>(es, sm);
        const double EEst = (double)ksqrt(es / T(n));
        if (EEst != EEst) { ret = RET_UNSTABLE; break; }
        const double q = pi_q(EEst, qold, q11);
        accept = EEst <= 1.0;
        if (accept) {
            ++naccept;
            qold = fmax(EEst, Ctrl::qoldinit);
            const double dtnew = dt / q;
            double tnew = t + dt;
            if (fabs(tnew - t1) < 100.0 * eps_of(fmax(fabs(t), fabs(t1)))) tnew = t1;
            dtpropose = fmax(fmin(dtmax, fabs(dtnew)), fmax(eps_of(tnew), dtmin0));
            if (DENSE) {
                if (nrec >= a.cap) { ret = RET_OVERFLOW; break; }
                if (tid == 0) { a.rec_t[b * a.cap + nrec] = t; a.rec_dt[b * a.cap + nrec] = h; }
                T* r = a.rec + (b * a.cap + nrec) * 8 * (int64_t)n;
                for (int i = tid; i < n; i += bt) r[i] = uprev[i];
                for (int i = tid; i < 7 * n; i += bt) r[n + i] = k[i];
                ++nrec;
            }
            while (sidx < a.nsave && a.saveat[sidx] <= tnew) {
                const T th = (T)((a.saveat[sidx] - t) / dt);
                T bw[7]; interp_weights(th, bw);
                for (int i = tid; i < n; i += bt) {
                    T acc = T(0);
#pragma unroll
                    for (int j = 0; j < 7; ++j) acc += bw[j] * k[(int64_t)j * n + i];
                    const T v = uprev[i] + h * acc;
                    const int64_t o = (b * a.nsave + sidx) * n + i;
                    if (a.out) a.out[o] = v;
                    if (DENSE && a.target) {                        // target == null: a.dg holds the caller's cotangents
                        const T e = v - a.target[o];
                        lsum += (double)e * (double)e;
                        a.dg[o] = (T(2) / (T)((double)n * a.nsave)) * e;
                    }
                }
                ++sidx;
            }
            t = tnew;
            __syncthreads();
            for (int i = tid; i < n; i += bt) { uprev[i] = tmp[i]; u[i] = tmp[i]; k[i] = k[(int64_t)6 * n + i]; }
            __syncthreads();
        } else {
            ++nreject;
        }
    }
    if (ret != RET_SUCCESS)
        for (; sidx < a.nsave; ++sidx)
            for (int i = tid; i < n; i += bt) {
                if (a.out) a.out[(b * a.nsave + sidx) * n + i] = T(NAN);
                if (DENSE) a.dg[(b * a.nsave + sidx) * n + i] = T(0);
            }
    if (tid == 0) {
        if (a.stats) a.stats[b] = kanode_stats{naccept, nreject, nf, ret};
        if (DENSE) { a.nsteps[b] = nrec; a.retcode[b] = ret; }
    }
    if (DENSE) {
        T ls = (T)lsum;      // per-thread partial; block-reduce in double via two-float split is overkill here
        double tot = (double)block_sum<T>(ls, sm);
        if (tid == 0 && tot != 0.0) atomicAdd(a.loss_sum, tot);
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward: interpolating adjoint, one block per trajectory.  GP >= G+1 of every layer (feature registers).
// ---------------------------------------------------------------------------------------------------------
template <class T, int GP>
__device__ void g_gphase_layer(const GenericModel& m, int l, const T* recs, const T (&wb)[7], const T (&wbt)[7],
                               const T* gold, T* gnew, T abstol, T reltol, T& es, GSmem<T>& sm) {
    const GenericLayer& L = m.L[l];
    const float* grid = m.grid + L.goff;
    const int I = L.I, O = L.O, G = L.G, GQ = L.G + 1;
    const long long RL = m.rec_len;
    auto finalize = [&](long long j, T vb, T vt) {
        const T g0 = gold[j];
        const T g1 = g0 + vb;
        const T sc = abstol + kmax(kabs(g0), kabs(g1)) * reltol;
        const T r = vt / sc;
        es += r * r;
        gnew[j] = g1;
    };
    const int IB = g_inputs(L);
    if (O > I && 7LL * IB * GQ <= GEN_FEAT) {         // thread per output; features of all inputs in smem
        __syncthreads();
        for (int idx = threadIdx.x; idx < 7 * IB; idx += blockDim.x) {
            const int s = idx / IB, i = idx - s * IB;
            g_features_at(L, grid, recs + s * RL + L.rx, i, sm.feat + (long long)idx * GQ);
        }
        __syncthreads();
        for (int o = threadIdx.x; o < O; o += blockDim.x) {
            T ab[7], at[7];
#pragma unroll
            for (int s = 0; s < 7; ++s) { const T yb = recs[s * RL + L.ry + o]; ab[s] = wb[s] * yb; at[s] = wbt[s] * yb; }
            for (int i = 0; i < IB; ++i)
                for (int q = 0; q < GQ; ++q) {
                    if (q == G ? !L.use_base : i >= I) continue;
                    T vb = T(0), vt = T(0);
#pragma unroll
                    for (int s = 0; s < 7; ++s) { const T c = sm.feat[((long long)s * IB + i) * GQ + q]; vb += ab[s] * c; vt += at[s] * c; }
                    finalize(q < G ? L.offC + ((long long)i * G + q) * O + o : L.offW + (long long)i * O + o, vb, vt);
                }
        }
        __syncthreads();
        return;
    }
    for (int i = threadIdx.x; i < IB; i += blockDim.x) {    // thread per input unit (incl. the bias unit); features in registers
        T c[7][GP];
#pragma unroll
        for (int s = 0; s < 7; ++s) {
            const T xi = i < I ? recs[s * RL + L.rx + i] : T(0);
            const T xn = (G > 0 && i < I) ? normalize_rt(L.norm, xi) : T(0);
#pragma unroll
            for (int q = 0; q < GP - 1; ++q) c[s][q] = (q < G && i < I) ? basis_val(L.basis, (xn - (T)grid[q < G ? q : 0]) * (T)L.inv_h) : T(0);
            T sw = i < I ? T(0) : T(1);
            if (L.use_base && i < I) base_fwd(L.act, xi, sw);
            c[s][GP - 1] = sw;                                 // base activation kept in the last slot
        }
        for (int o = 0; o < O; ++o) {
            T ab[7], at[7];
#pragma unroll
            for (int s = 0; s < 7; ++s) { const T yb = recs[s * RL + L.ry + o]; ab[s] = wb[s] * yb; at[s] = wbt[s] * yb; }
#pragma unroll
            for (int q = 0; q < GP; ++q) {
                const bool is_sw = (q == GP - 1);
                if (is_sw ? !L.use_base : (q >= G || i >= I)) continue;
                T vb = T(0), vt = T(0);
#pragma unroll
                for (int s = 0; s < 7; ++s) { vb += ab[s] * c[s][q]; vt += at[s] * c[s][q]; }
                finalize(is_sw ? L.offW + (long long)i * O + o : L.offC + ((long long)i * G + q) * O + o, vb, vt);
            }
        }
    }
    __syncthreads();
}

template <class T, int GP>
__global__ void __launch_bounds__(GEN_BT) generic_backward_kernel(const __grid_constant__ GenericModel m, const GenBwdArgs<T> a) {
    __shared__ GSmem<T> sm;
    const int64_t b = blockIdx.x;
    const int n = m.n, tid = threadIdx.x, bt = blockDim.x;
    const long long np = m.np, RL = m.rec_len;
    const bool chain = m.rhs_kind == KANODE_RHS_CHAIN;
    const int NZ = n + (int)np;
    T* lam = a.work + b * (11 * (long long)n + 7 * RL);
    T* lprev = lam + n; T* ls = lprev + n; T* y = ls + n; T* kl = y + n; T* recs = kl + 7 * (long long)n;
    T* gbuf = a.g + b * 2 * np;
    const T* p = a.p;
    for (long long j = tid; j < 2 * np; j += bt) gbuf[j] = T(0);
    for (int i = tid; i < n; i += bt) { lam[i] = T(0); lprev[i] = T(0); }
    for (int i = tid; i < 7 * n; i += bt) kl[i] = T(0);
    int nf = 0, naccept = 0, nreject = 0, ret = a.retcode[b];
    const int nsteps = a.nsteps[b];
    if (ret != RET_SUCCESS || nsteps <= 0) {
        if (tid == 0 && a.stats) a.stats[b] = kanode_stats{0, 0, 0, ret};
        if (a.du0) for (int i = tid; i < n; i += bt) a.du0[b * n + i] = T(0);
        return;
    }
    __syncthreads();
    const double* rt = a.rec_t + b * a.cap;
    const T* rdt = a.rec_dt + b * a.cap;
    const T* rec = a.rec + b * a.cap * 8 * (int64_t)n;
    int ridx = nsteps - 1;
    auto eval_y = [&](double t) {                       // y = sol(t) from the dense forward record
        while (t < rt[ridx] && ridx > 0) --ridx;
        while (ridx + 1 < nsteps && t >= rt[ridx + 1]) ++ridx;
        const T hd = rdt[ridx];
        const T th = (T)((t - rt[ridx]) / (double)hd);
        T bw[7]; interp_weights(th, bw);
        const T* r = rec + (int64_t)ridx * 8 * n;
        for (int i = tid; i < n; i += bt) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < 7; ++j) acc += bw[j] * r[(int64_t)(1 + j) * n + i];
            y[i] = r[i] + hd * acc;
        }
        __syncthreads();
    };
    // adjoint RHS at (t, l): dl = -(df/du)^T l; stage record (or kg for the source model) -> slot
    auto adj_eval = [&](double t, const T* l, T* dl, int slot) {
        eval_y(t);
        g_vjp<T>(m, p, y, l, dl, sm, recs + slot * RL);
        for (int i = tid; i < n; i += bt) dl[i] = -dl[i];
        __syncthreads();
        ++nf;
    };
    // sum over the gradient components of fn(kv0, kv1) for slots 0/1 (initdt only; chain -> rank-1 records)
    auto g_norms = [&](T& s1, T& s2, bool two) {
        if (!chain) {
            for (int j = tid; j < (int)np; j += bt) {
                const T k0 = recs[j], k1 = two ? recs[RL + j] : T(0);
                const T x1 = k0 / a.abstol, x2 = (k1 - k0) / a.abstol;
                s1 += x1 * x1; s2 += x2 * x2;
            }
            return;
        }
        for (int l = 0; l < m.n_layers; ++l) {
            const GenericLayer& L = m.L[l];
            const float* grid = m.grid + L.goff;
            for (int i = tid; i < g_inputs(L); i += bt) {
                T c0[GEN_MAX_G + 1], c1[GEN_MAX_G + 1];
                g_features_at(L, grid, recs + L.rx, i, c0);
                if (two) g_features_at(L, grid, recs + RL + L.rx, i, c1);
                for (int o = 0; o < L.O; ++o) {
                    const T a0 = recs[L.ry + o], a1 = two ? recs[RL + L.ry + o] : T(0);
                    for (int q = 0; q <= L.G; ++q) {
                        if (q == L.G ? !L.use_base : i >= L.I) continue;
                        const T k0 = a0 * c0[q], k1 = two ? a1 * c1[q] : T(0);
                        const T x1 = k0 / a.abstol, x2 = (k1 - k0) / a.abstol;
                        s1 += x1 * x1; s2 += x2 * x2;
                    }
                }
            }
        }
    };

    const double t0 = a.t0, t1 = a.t1, dtmax = fabs(t1 - t0), dtmin0 = fmax(eps_of(t0), eps_of(t1));
    const T abstol = a.abstol, reltol = a.reltol;
    double t = t1;
    int sp = a.nsave - 1;
    auto apply_jumps = [&](double tt) {
        bool mod = false;
        while (sp >= 0 && a.saveat[sp] == tt) {
            for (int i = tid; i < n; i += bt) lam[i] += a.dg[(b * a.nsave + sp) * n + i];
            --sp; mod = true;
        }
        __syncthreads();
        return mod;
    };
    apply_jumps(t1);                                      // PresetTimeCallback fires at init when t_end is a save time
    for (int i = tid; i < n; i += bt) lprev[i] = lam[i];
    __syncthreads();
    adj_eval(t, lam, kl, 0);
    double dt;
    {   // ---- initdt on the augmented state ----
        T v[3] = {T(0), T(0), T(0)};
        for (int i = tid; i < n; i += bt) {
            const T sk = abstol + kabs(lam[i]) * reltol;
            const T x0 = lam[i] / sk, x1 = kl[i] / sk;
            v[0] += x0 * x0; v[1] += x1 * x1;
        }
        g_norms(v[1], v[2], false);
        v[2] = T(0);
        block_reduce<T, 3>(v, sm);
        const double d0 = sqrt((double)v[0] / NZ), d1 = sqrt((double)v[1] / NZ);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
        dt0 = fmin(dt0, dtmax);
        for (int i = tid; i < n; i += bt) ls[i] = lam[i] - (T)dt0 * kl[i];
        __syncthreads();
        adj_eval(t - dt0, ls, kl + n, 1);
        ++nf;
        T w[2] = {T(0), T(0)};
        for (int i = tid; i < n; i += bt) { const T sk = abstol + kabs(lam[i]) * reltol; const T x = (kl[n + i] - kl[i]) / sk; w[1] += x * x; }
        g_norms(w[0], w[1], true);
        block_reduce<T, 2>(w, sm);
        const double d2 = sqrt((double)w[1] / NZ) / dt0, mx = fmax(d1, d2);
        const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        dt = fmax(dtmin0, fmin(fmin(100.0 * dt0, dt1), dtmax));
        __syncthreads();
        for (int i = tid; i < n; i += bt) kl[n + i] = T(0);
        __syncthreads();
    }
    double qold = Ctrl::qoldinit, q11 = 1.0, dtpropose = dt;
    bool accept = false, modified = false;
    int iter = 0, cur = 0;
    while (t > t0) {
        if (iter > 0) {
            if (!accept) dt = dt / fmin(1.0 / Ctrl::qmin, q11 / Ctrl::gamma);
            else {
                dt = dtpropose;
                if (!modified) {                            // FSAL
                    for (int i = tid; i < n; i += bt) kl[i] = kl[(int64_t)6 * n + i];
                    for (long long i = tid; i < RL; i += bt) recs[i] = recs[6 * RL + i];
                    __syncthreads();
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
        const T h = (T)(-dt);
        for (int s = modified ? 0 : 1; s < 7; ++s) {
            T as[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) as[j] = Tab<T>::a(s, j);
            for (int i = tid; i < n; i += bt) {
                T acc = T(0);
#pragma unroll
                for (int j = 0; j < 6; ++j) acc += as[j] * kl[(int64_t)j * n + i];
                ls[i] = lprev[i] + h * acc;
            }
            __syncthreads();
            adj_eval(t - tab_c(s) * dt, ls, kl + (int64_t)s * n, s);
        }
        modified = false;                                    // ls now holds lambda_new
        T es = T(0);
        for (int i = tid; i < n; i += bt) {
            T ut = T(0);
#pragma unroll
            for (int j = 0; j < 7; ++j) ut += Tab<T>::bt(j) * kl[(int64_t)j * n + i];
            const T sc = abstol + kmax(kabs(lprev[i]), kabs(ls[i])) * reltol;
            const T r = (h * ut) / sc;
            es += r * r;
        }
        {
            T wb[7], wbt[7];
#pragma unroll
            for (int s = 0; s < 7; ++s) { wb[s] = -h * Tab<T>::b(s); wbt[s] = -h * Tab<T>::bt(s); }
            const T* gold = gbuf + (long long)cur * np;
            T* gnew = gbuf + (long long)(cur ^ 1) * np;
            if (chain) {
                for (int l = 0; l < m.n_layers; ++l) g_gphase_layer<T, GP>(m, l, recs, wb, wbt, gold, gnew, abstol, reltol, es, sm);
            } else {
                for (int j = tid; j < (int)np; j += bt) {
                    T vb = T(0), vt = T(0);
#pragma unroll
                    for (int s = 0; s < 7; ++s) { const T kv = recs[s * RL + j]; vb += wb[s] * kv; vt += wbt[s] * kv; }
                    const T g0 = gold[j], g1 = g0 + vb;
                    const T sc = abstol + kmax(kabs(g0), kabs(g1)) * reltol;
                    const T r = vt / sc;
                    es += r * r;
                    gnew[j] = g1;
                }
            }
        }
        es = block_sum<T>(es, sm);
        const double EEst = (double)ksqrt(es / T(NZ));
        if (EEst != EEst) { ret = RET_UNSTABLE; break; }
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
            for (int i = tid; i < n; i += bt) lam[i] = ls[i];
            __syncthreads();
            modified = apply_jumps(t);
            for (int i = tid; i < n; i += bt) lprev[i] = lam[i];
            __syncthreads();
        } else {
            ++nreject;
        }
    }
    __syncthreads();
    if (cur == 1) for (long long j = tid; j < np; j += bt) gbuf[j] = gbuf[np + j];
    if (ret != RET_SUCCESS) for (long long j = tid; j < np; j += bt) gbuf[j] = T(0);
    if (a.du0) for (int i = tid; i < n; i += bt) a.du0[b * n + i] = lam[i];
    if (tid == 0 && a.stats) a.stats[b] = kanode_stats{naccept, nreject, nf, ret};
}

// out[j] = sum_b g[b][0][j]   (per-trajectory gradients are [B][2][np]; buffer 0 holds the result)
template <class T>
__global__ void __launch_bounds__(256) generic_grad_reduce_kernel(const T* g, long long np, int64_t B, T* out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= np) return;
    double acc = 0.0;
    for (int64_t b = 0; b < B; ++b) acc += (double)g[b * 2 * np + j];
    out[j] = (T)acc;
}

// =========================================================================================================
// host front-ends
// =========================================================================================================
inline int generic_init(kanode_handle* h) {
    GenericModel& m = h->gm;
    const kanode_desc& d = h->desc;
    m = GenericModel{};
    m.n_layers = d.n_layers; m.rhs_kind = d.rhs_kind; m.n = d.n_state; m.np = (long long)h->np;
    m.n_out = d.rhs_kind == KANODE_RHS_MAP ? d.layers[d.n_layers - 1].out_dims : d.n_state;
    m.lap_scale = d.rhs_kind == KANODE_RHS_SOURCE_LAPLACIAN ? d.lap_coef / (d.dx * d.dx) : 0.0;
    long long off = 0, roff = 0;
    int goff = 0;
    for (int l = 0; l < d.n_layers; ++l) {
        const kanode_layer_desc& s = d.layers[l];
        GenericLayer& L = m.L[l];
        L.I = s.in_dims; L.O = s.out_dims; L.goff = goff;
        if (s.kind == KANODE_LAYER_DENSE) {
            // y = act(W x + b): no spline part; the base branch carries W with the activation of the layer BEFORE on the inputs
            // (identity on the first layer), and a virtual input unit with feature 1 carries b — [vec(W); b] is an O x (I+1) block
            L.G = 0; L.norm = 0; L.basis = 0; L.use_base = 1; L.bias = 1; L.inv_h = 1.0f;
            L.act = (l > 0 && d.layers[l - 1].kind == KANODE_LAYER_DENSE && d.layers[l - 1].dense_act == KANODE_ACT_TANH) ? 2 : 1;
            L.offC = off; L.offW = off; off += (long long)L.O * (L.I + 1);
        } else {
            L.G = s.grid_len; L.norm = s.normalizer; L.basis = s.basis; L.use_base = s.use_base_act; L.act = 0; L.bias = 0;
            L.inv_h = 1.0f / s.denominator;
            if (s.grid_len <= GEN_MAX_G) for (int g = 0; g < s.grid_len; ++g) m.grid[goff + g] = grid_point(s, g);
            goff += s.grid_len <= GEN_MAX_G ? s.grid_len : 0;
            L.offC = off; off += (long long)L.O * L.G * L.I;
            L.offW = off; if (L.use_base) off += (long long)L.O * L.I;
        }
        L.rx = roff; roff += L.I;
        L.ry = roff; roff += L.O;
    }
    m.rec_len = d.rhs_kind != KANODE_RHS_SOURCE_LAPLACIAN ? roff : (long long)h->np;
    return 0;
}

// does the generic path cover this descriptor?  (message in h->err otherwise)
inline int generic_supported(kanode_handle* h) {
    const kanode_desc& d = h->desc;
    for (int l = 0; l < d.n_layers; ++l) {
        if (d.layers[l].kind == KANODE_LAYER_KDENSE && d.layers[l].grid_len > GEN_MAX_G) return fail(h, KANODE_ERR_UNSUPPORTED, "grid_len > %d", GEN_MAX_G);
        if (d.rhs_kind != KANODE_RHS_SOURCE_LAPLACIAN && l > 0 && d.layers[l].in_dims > GEN_ACT)
            return fail(h, KANODE_ERR_UNSUPPORTED, "hidden width %d > %d", d.layers[l].in_dims, GEN_ACT);
        if (d.rhs_kind == KANODE_RHS_SOURCE_LAPLACIAN && d.layers[l].kind != KANODE_LAYER_KDENSE)
            return fail(h, KANODE_ERR_UNSUPPORTED, "the hidden-source model takes KDense layers");
        if (d.rhs_kind == KANODE_RHS_SOURCE_LAPLACIAN && (d.layers[l].in_dims > GEN_PW || d.layers[l].out_dims > GEN_PW))
            return fail(h, KANODE_ERR_UNSUPPORTED, "pointwise chain wider than %d", GEN_PW);
    }
    if (d.rhs_kind == KANODE_RHS_SOURCE_LAPLACIAN && h->np > (size_t)GEN_FEAT)
        return fail(h, KANODE_ERR_UNSUPPORTED, "source model with more than %d parameters", GEN_FEAT);
    return 0;
}

inline int generic_upload_params(kanode_handle* h) {
    float* pf = nullptr; double* pd = nullptr;
    ENSURE(h, W_PARAMS, sizeof(float) * h->np, pf);
    ENSURE(h, W_PARAMS64, sizeof(double) * h->np, pd);
    std::vector<float> tmp(h->np);
    for (size_t i = 0; i < h->np; ++i) tmp[i] = (float)h->params[i];
    CK(h, cudaMemcpyAsync(pf, tmp.data(), sizeof(float) * h->np, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(pd, h->params.data(), sizeof(double) * h->np, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}
template <class T> const T* generic_params(kanode_handle* h);
template <> inline const float* generic_params<float>(kanode_handle* h) { return (const float*)h->ws[kanode_handle::W_PARAMS].p; }
template <> inline const double* generic_params<double>(kanode_handle* h) { return (const double*)h->ws[kanode_handle::W_PARAMS64].p; }

template <class T> int generic_rhs(kanode_handle* h, const T* d_u, T* d_du, int64_t B) {
    if (int rc = generic_supported(h)) return rc;
    generic_rhs_kernel<T><<<(unsigned)B, GEN_BT, 0, h->stream>>>(h->gm, generic_params<T>(h), d_u, d_du);
    ++h->launches;
    CK(h, cudaGetLastError());
    return 0;
}

template <class T> int generic_vjp(kanode_handle* h, const T* d_u, const T* d_lam, T* d_ubar, T* d_pbar, int64_t B) {
    if (int rc = generic_supported(h)) return rc;
    T* recs = nullptr;
    ENSURE(h, W_GEN, sizeof(T) * (size_t)h->gm.rec_len * B, recs);
    CK(h, cudaMemsetAsync(d_pbar, 0, sizeof(T) * h->np, h->stream));
    generic_vjp_kernel<T><<<(unsigned)B, GEN_BT, 0, h->stream>>>(h->gm, generic_params<T>(h), d_u, d_lam, d_ubar, d_pbar, recs);
    ++h->launches;
    CK(h, cudaGetLastError());
    return 0;
}

template <class T>
int generic_solve(kanode_handle* h, const T* d_u0, int64_t B, double t0, double t1, const double* d_saveat, int nsave,
                  double abstol, double reltol, T* d_out, kanode_stats* d_stats) {
    if (int rc = generic_supported(h)) return rc;
    GenFwdArgs<T> a{};
    a.p = generic_params<T>(h); a.u0 = d_u0; a.B = B; a.t0 = t0; a.t1 = t1; a.saveat = d_saveat; a.nsave = nsave;
    a.abstol = (T)abstol; a.reltol = (T)reltol; a.maxiters = 100000; a.out = d_out; a.stats = d_stats;
    ENSURE(h, W_GEN, sizeof(T) * (size_t)10 * h->n * B, a.work);
    generic_forward_kernel<T, false><<<(unsigned)B, GEN_BT, 0, h->stream>>>(h->gm, a);
    ++h->launches;
    CK(h, cudaGetLastError());
    return 0;
}

template <class T>
int generic_loss_grad(kanode_handle* h, const T* d_u0, int64_t B, double t0, double t1, const double* d_saveat, int nsave,
                      const T* d_target, double abstol, double reltol, double* d_loss_sum, T* d_grad_sum, T* d_du0,
                      kanode_stats* d_fst, kanode_stats* d_bst, T* d_out_opt) {
    if (int rc = generic_supported(h)) return rc;
    const GenericModel& m = h->gm;
    const size_t n = (size_t)h->n, np = h->np;
    const int cap = h->rec_cap;
    GenFwdArgs<T> a{};
    a.p = generic_params<T>(h); a.u0 = d_u0; a.B = B; a.t0 = t0; a.t1 = t1; a.saveat = d_saveat; a.nsave = nsave;
    a.abstol = (T)abstol; a.reltol = (T)reltol; a.maxiters = 100000; a.out = d_out_opt; a.stats = d_fst;
    a.cap = cap; a.target = d_target; a.loss_sum = d_loss_sum;
    const size_t fwd_work = 10 * n, bwd_work = 11 * n + 7 * (size_t)m.rec_len;
    T* work = nullptr;
    ENSURE(h, W_GEN, sizeof(T) * (fwd_work > bwd_work ? fwd_work : bwd_work) * B, work);
    a.work = work;
    ENSURE(h, W_REC_T, sizeof(double) * (size_t)cap * B, a.rec_t);
    ENSURE(h, W_GEN2, sizeof(T) * (size_t)cap * B, a.rec_dt);
    ENSURE(h, W_REC, sizeof(T) * (size_t)cap * 8 * n * B, a.rec);
    ENSURE(h, W_NSTEPS, sizeof(int) * (size_t)B, a.nsteps);
    ENSURE(h, W_RET, sizeof(int) * (size_t)B, a.retcode);
    ENSURE(h, W_DG, sizeof(T) * (size_t)nsave * n * B, a.dg);
    T* g = nullptr;
    ENSURE(h, W_G, sizeof(T) * 2 * np * B, g);
    cudaEventRecord(h->ev[0], h->stream);
    generic_forward_kernel<T, true><<<(unsigned)B, GEN_BT, 0, h->stream>>>(m, a);
    cudaEventRecord(h->ev[1], h->stream);
    GenBwdArgs<T> bw{};
    bw.p = a.p; bw.B = B; bw.t0 = t0; bw.t1 = t1; bw.saveat = d_saveat; bw.nsave = nsave;
    bw.abstol = a.abstol; bw.reltol = a.reltol; bw.maxiters = 100000;
    bw.rec_t = a.rec_t; bw.rec_dt = a.rec_dt; bw.rec = a.rec; bw.cap = cap; bw.nsteps = a.nsteps; bw.retcode = a.retcode;
    bw.dg = a.dg; bw.work = work; bw.g = g; bw.du0 = d_du0; bw.stats = d_bst;
    int gmax = 0;
    for (int l = 0; l < m.n_layers; ++l) gmax = gmax > m.L[l].G ? gmax : m.L[l].G;
    if (gmax + 1 <= 6) generic_backward_kernel<T, 6><<<(unsigned)B, GEN_BT, 0, h->stream>>>(m, bw);
    else if (gmax + 1 <= 11) generic_backward_kernel<T, 11><<<(unsigned)B, GEN_BT, 0, h->stream>>>(m, bw);
    else generic_backward_kernel<T, GEN_MAX_G + 1><<<(unsigned)B, GEN_BT, 0, h->stream>>>(m, bw);
    cudaEventRecord(h->ev[2], h->stream);
    generic_grad_reduce_kernel<T><<<(unsigned)((np + 255) / 256), 256, 0, h->stream>>>(g, (long long)np, B, d_grad_sum);
    cudaEventRecord(h->ev[3], h->stream);
    h->ev_valid = true;
    h->launches += 3;
    CK(h, cudaGetLastError());
    return 0;
}

}  // namespace kanode
