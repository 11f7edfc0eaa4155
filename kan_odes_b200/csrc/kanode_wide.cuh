// kanode_wide.cuh — batched LOCKSTEP engine for the wide PDE surrogates: two-layer KDense chains [n, H, n]
// (Burgers-1024, Allen-Cahn-4096, Schrodinger-32768 of BASELINE.json; H = 10 in every reference script).
//
// Why a second engine: the block-per-trajectory kernels (kanode_generic.cuh) stream all np = 2*n*H*(G+1) weights
// through one block per right-hand-side evaluation, so a batch of B initial conditions re-reads the weights B times
// from L2 and keeps only B SMs busy.  The RHS is autonomous, so all ICs can evaluate "stage s of their current step
// attempt" at the same time whatever their own t and dt are: the batch advances in lockstep over step ATTEMPTS,
// every IC keeps its own PI controller, accept/reject decision, dense record and tstops (identical per-IC
// arithmetic to the generic path / oracle), and finished ICs are masked.  Every contraction then runs as a kernel
// over the whole batch with the weights of one input/output unit held in REGISTERS and reused across the ICs:
//
//   wide_l1_fwd   thread = input unit i   w1[(G+1)*H] regs   hidden[b][:] += w1 . features(x[b][i])   (reduce over n)
//   wide_l2_fwd   thread = output unit o  w2[H*(G+1)] regs   k[b][o] = w2 . features(hidden[b][:])    (expand)
//   wide_l2_vjp   thread = output unit o  w2 regs            hbar[b][j] += lam[b][o] * (w2_j . dfeat_j(hidden[b]))
//   wide_l1_vjp   thread = input unit i   w1 regs            lambda-dot[b][i] = -dfeat(x[b][i]) . (w1 . hbar[b])
//
// The RBF features are produced and consumed in registers / shared memory and never reach HBM.  Cross-block sums
// (over the n units) are two-level and deterministic: per-block partials, then the last block to finish (device
// counter) adds them in a fixed order.  The adjoint integrates z = [lambda; g] per IC exactly like the reference's
// InterpolatingAdjoint: dg/dt is kept as rank-1 stage records (layer inputs x_l, output cotangents ybar_l) and the
// step-end pass streams g[b] once (read g_old, write g_new, error-norm contribution) — that pass is the HBM-bound
// part of a wide step (8 bytes per parameter per IC per attempt).
//
// Reference semantics: KDense forward Lotka-Volterra/src/kdense.jl:109-130, reverse rules src/utils.jl:15-21;
// drivers "PDE examples/Burgers_Surrogate.jl":82-107, Allen-Cahn_Surrogate.jl:80-107, Schrodinger_Surrogate.jl:89-114;
// Tsit5 / controller / adjoint: [EXT OrdinaryDiffEqTsit5 1.1.0, OrdinaryDiffEqCore 1.9.0, SciMLSensitivity 7.69.0].
#pragma once
#include "kanode_host.h"
#include "kanode_math.cuh"
#include "kanode_wide_api.h"
#include <algorithm>
#include <cstdlib>

namespace kanode {

constexpr int W_BT = 128;      // threads per block of the unit-per-thread kernels
constexpr int W_ET = 256;      // threads per block of the elementwise kernels
constexpr int W_MAXCH = 64;    // most partials per IC in the reduce kernels
constexpr int W_HP = 16;       // padded hidden width of the [B][16] arrays
constexpr int W_PT = 32;       // ICs per block in the kernels without a reduction

struct WideModel {
    int n, norm1, norm2;
    float inv_h1, inv_h2;
    float grid1[16], grid2[16];
    long long offC1, offW1, offC2, offW2, np;
};

// per-IC solver state, structure of arrays (device)
struct WideCtl {
    double *t, *dt, *dtpropose, *qold, *q11, *told, *dtold, *d0, *d1, *dt0;
    int *iter, *accept, *acc_now, *fail_now, *active, *modified, *do_s0, *sidx, *s_lo, *s_hi, *nrec, *naccept, *nreject, *nf,
        *ret, *cur;
    int* ridx;   // [7][B] dense-record index of every backward stage time
    int* mask7;  // [7][B] backward: which (stage, IC) pairs are evaluated in this attempt (stage 0 only after a jump)
};
constexpr int W_NCTL_D = 10, W_NCTL_I = 16 + 7 + 7;

template <class T> __device__ __forceinline__ T ld_cg(const T* p) { return __ldcg(p); }
// ratio for the error norm: fp32 uses rcp.approx (1 ulp; it only feeds EEst), fp64 the IEEE quotient
__device__ __forceinline__ float wdiv(float a, float b) { return a * krcp(b); }
__device__ __forceinline__ double wdiv(double a, double b) { return a / b; }

template <class T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// deterministic block sum (blockDim multiple of 32, <= 1024); result valid in every thread
template <class T> __device__ __forceinline__ T wblock_sum(T v, T* sred /*[33]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sred[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) { T s = T(0); for (int w = 0; w < nw; ++w) s += sred[w]; sred[32] = s; }
    __syncthreads();
    return sred[32];
}

// features of one scalar: c[q] = rbf_q(norm(x)) (q < G), c[G] = swish(x)          (kdense.jl:116-123)
template <class T, int G> __device__ __forceinline__ void w_features(int norm, T inv_h, const float* grid, T x, T (&c)[G + 1]) {
    const T xn = normalize_rt(norm, x);
#pragma unroll
    for (int g = 0; g < G; ++g) { const T a = (xn - (T)grid[g]) * inv_h; c[g] = kexp(-a * a); }
    swish_fwd(x, c[G]);
}
// d c[q] / dx  (utils.jl:15-21 + NNlib activation rules)
template <class T, int G> __device__ __forceinline__ void w_dfeatures(int norm, T inv_h, const float* grid, T x, T (&d)[G + 1]) {
    const T xn = normalize_rt(norm, x);
    const T dn = normalize_deriv_rt(norm, xn);
#pragma unroll
    for (int g = 0; g < G; ++g) { const T a = (xn - (T)grid[g]) * inv_h; const T y = kexp(-a * a); d[g] = (T(-2) * a * y) * inv_h * dn; }
    T s, ds; swish_both(x, s, ds);
    d[G] = ds;
}

// stage input of a reduce kernel.  MODE 0: x = base + hs[b] * sum_j coef[j] * ks[j]   (Tsit5 stage combination)
//                                  MODE 1: x = sol(t_stage) from the dense forward record (backward pass)
template <class T> struct WideIn {
    const T* base; const T* ks; const T* hs; T coef[7]; int ncoef;
    const T* rec; int cap; const int* ridx; const T* th; const T* hd;
    int brec;            // > 0: b indexes (stage, IC) pairs laid out [7][brec]; the dense record of pair b belongs to IC b % brec
    T* xstore;           // optional copy of x: [B][n]
    const int* mask;     // per-IC, 0 = skip (may be null)
};

// Per-tile table of the dense-output data of its ICs (MODE 1): interpolation weights b_1..b_7(theta), the step size and the
// record row, computed once per block instead of once per (unit, IC) pair.
template <class T, int NB> struct WideInterpTab { T bw[NB][8]; const T* row[NB]; };
template <class T, int NB>
__device__ __forceinline__ void wide_interp_tab(const WideIn<T>& in, int b0, int b1, int n, WideInterpTab<T, NB>& tab) {
    if (threadIdx.x < NB) {
        const int bl = threadIdx.x, b = b0 + bl < b1 ? b0 + bl : b0;
        T bw[7]; interp_weights(in.th[b], bw);
#pragma unroll
        for (int j = 0; j < 7; ++j) tab.bw[bl][j] = bw[j];
        tab.bw[bl][7] = in.hd[b];
        const int brow = in.brec > 0 ? b % in.brec : b;
        tab.row[bl] = in.rec + ((int64_t)brow * in.cap + in.ridx[b]) * 8 * (int64_t)n;
    }
    __syncthreads();
}

// Inputs of unit i for the NB ICs b0..b0+NB of a tile.  All global loads are issued before the first dependent FMA (the
// loop is written in phases and fully unrolled), so one memory latency covers the whole tile.
template <class T, int MODE, int NB>
__device__ __forceinline__ void wide_inputs(const WideIn<T>& in, int b0, int b1, int i, bool valid, int n, int64_t B, T (&xs)[NB], bool (&on)[NB],
                                            const WideInterpTab<T, NB>* tab = nullptr) {
#pragma unroll
    for (int bl = 0; bl < NB; ++bl) { const int b = b0 + bl; on[bl] = valid && b < b1 && (!in.mask || in.mask[b]); }
    if (MODE == 0) {
        T base[NB], hs[NB], kv[NB][6];
        const int Bn = (int)B * n;
#pragma unroll
        for (int bl = 0; bl < NB; ++bl) {
            const int e = (b0 + bl) * n + i;
            base[bl] = on[bl] ? in.base[e] : T(0);
            hs[bl] = on[bl] && in.ncoef > 0 ? in.hs[b0 + bl] : T(0);
#pragma unroll
            for (int j = 0; j < 6; ++j) kv[bl][j] = (on[bl] && j < in.ncoef) ? in.ks[(int64_t)j * Bn + e] : T(0);
        }
#pragma unroll
        for (int bl = 0; bl < NB; ++bl) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < 6; ++j) acc += in.coef[j] * kv[bl][j];
            xs[bl] = in.ncoef > 0 ? base[bl] + hs[bl] * acc : base[bl];
        }
    } else {
        T raw[NB][8];
#pragma unroll
        for (int bl = 0; bl < NB; ++bl) {
            const T* r = tab->row[bl] + i;
#pragma unroll
            for (int j = 0; j < 8; ++j) raw[bl][j] = on[bl] ? r[j * n] : T(0);
        }
#pragma unroll
        for (int bl = 0; bl < NB; ++bl) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < 7; ++j) acc += tab->bw[bl][j] * raw[bl][1 + j];
            xs[bl] = raw[bl][0] + tab->bw[bl][7] * acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// shared tail of the two reduce kernels: red[v][tid] column sums -> part[chunk][b][H]; last block -> dst[b][W_HP]
// ---------------------------------------------------------------------------------------------------------
// last block of a b-tile (device counter) adds the per-block partials part[chunk][b][H] in a fixed order -> dst[b][W_HP]
template <class T, int H>
__device__ __forceinline__ void wide_finalize(int b0, int b1, int64_t B, T* part, T* dst, unsigned* counters, int* s_last, const int* mask) {
    const int tid = threadIdx.x, nchunk = gridDim.x;
    const int nv = (b1 - b0) * H;
    __threadfence();
    __syncthreads();
    if (tid == 0) *s_last = (atomicAdd(&counters[blockIdx.y], 1u) == (unsigned)(nchunk - 1));
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    for (int v = tid; v < nv; v += blockDim.x) {      // thread per value: coalesced over v, 8 independent loads in flight
        const T* src = part + (int64_t)b0 * H + v;
        T s = T(0);
        int c = 0;
        for (; c + 8 <= nchunk; c += 8) {
            T t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = ld_cg(src + (int64_t)(c + k) * B * H);
#pragma unroll
            for (int k = 0; k < 8; ++k) s += t[k];
        }
        for (; c < nchunk; ++c) s += ld_cg(src + (int64_t)c * B * H);
        const int b = b0 + v / H;
        if (!mask || mask[b]) dst[(int64_t)b * W_HP + v % H] = s;   // masked ICs keep their (FSAL-shifted) record
    }
    if (tid == 0) counters[blockIdx.y] = 0u;   // re-armed for the next launch on this stream
}
template <class T, int H>
__device__ __forceinline__ void wide_reduce_tail(T (*red)[W_BT + 1], int b0, int b1, int64_t B, T* part, T* dst, unsigned* counters, int* s_last,
                                                 const int* mask) {
    const int tid = threadIdx.x, chunk = blockIdx.x;
    const int nv = (b1 - b0) * H;
    __syncthreads();
    for (int v = tid; v < nv; v += W_BT) {
        T s = T(0);
        for (int k = 0; k < W_BT; ++k) s += red[v][k];
        part[((int64_t)chunk * B + b0 + v / H) * H + v % H] = s;
    }
    wide_finalize<T, H>(b0, b1, B, part, dst, counters, s_last, mask);
}

// dst[b][o] = sum_c part[c][b][o] as its own launch (tensor-core reduce kernel: many chunks, no last-block tail):
// 32 lanes per value, each adds every 32nd chunk, then a fixed-order shuffle tree
template <class T, int H>
__global__ void __launch_bounds__(128) wide_sum_partials_kernel(const T* part, int nchunk, int64_t B, T* dst, const int* mask) {
    const int64_t v = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int g = threadIdx.x & 31;
    T s = T(0);
    if (v < B * H) for (int c = g; c < nchunk; c += 32) s += ld_cg(part + (int64_t)c * B * H + v);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (g == 0 && v < B * H) { const int64_t b = v / H; if (!mask || mask[b]) dst[b * W_HP + (v - b * H)] = s; }
}

// ---------------------------------------------------------------------------------------------------------
// layer 1 forward (reduce over the n input units): hidden[b][o] = sum_i sum_q w1[i][q][o] * c_q(x[b][i])
// grid (nchunk, nbt); block W_BT; each block walks P passes of W_BT units and <= GB ICs
// ---------------------------------------------------------------------------------------------------------
template <class T, int H, int G, int MODE>
__global__ void __launch_bounds__(W_BT, 3) wide_l1_fwd_kernel(const __grid_constant__ WideModel m, const T* __restrict__ w1t, const WideIn<T> in,
                                                           int64_t B, int P, int btile, T* part, T* hidden, unsigned* counters) {
    constexpr int NQ = G + 1, NW = H * NQ, GB = sizeof(T) == 4 ? 8 : 4;
    __shared__ T red[GB * H][W_BT + 1];
    __shared__ int s_last;
    const int tid = threadIdx.x, n = m.n;
    const int b0 = blockIdx.y * btile, b1 = (int)min((int64_t)(b0 + btile), B);
    const T inv_h = (T)m.inv_h1;
    __shared__ WideInterpTab<T, GB> itab;
    if (MODE == 1) wide_interp_tab<T, GB>(in, b0, b1, n, itab);
    for (int pass = 0; pass < P; ++pass) {
        const int i = (blockIdx.x * P + pass) * W_BT + tid;
        const bool valid = i < n;
        T xs[GB]; bool on[GB];
        wide_inputs<T, MODE, GB>(in, b0, b1, i, valid, n, B, xs, on, &itab);
        T w[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) w[k] = valid ? w1t[k * n + i] : T(0);               // [NW][n]: coalesced over the units
#pragma unroll
        for (int bl = 0; bl < GB; ++bl) {
            if (b0 + bl >= b1) break;
            T acc[H];
#pragma unroll
            for (int o = 0; o < H; ++o) acc[o] = T(0);
            if (on[bl]) {
                if (in.xstore) in.xstore[(int64_t)(b0 + bl) * n + i] = xs[bl];
                T c[NQ];
                w_features<T, G>(m.norm1, inv_h, m.grid1, xs[bl], c);
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int o = 0; o < H; o += 2) kfma2b(acc[o], acc[o + 1], w[q * H + o], w[q * H + o + 1], c[q]);   // FFMA2
            }
#pragma unroll
            for (int o = 0; o < H; ++o) {
                if (pass == 0) red[bl * H + o][tid] = acc[o];
                else red[bl * H + o][tid] += acc[o];
            }
        }
    }
    wide_reduce_tail<T, H>(red, b0, b1, B, part, hidden, counters, &s_last, in.mask);
}

// ---------------------------------------------------------------------------------------------------------
// layer 2 forward (expand): out[b][o] = sum_j ( sum_q C2[(j,q)][o] * c_q(hidden[b][j]) + W2[j][o] * swish(hidden[b][j]) )
// grid (ceil(n / W_BT), nbt); weights of output o in registers, features of the ICs in shared memory
// ---------------------------------------------------------------------------------------------------------
template <class T, int H, int G> __device__ __forceinline__ void wide_load_w2(const WideModel& m, const T* __restrict__ p, int o, T (&w)[H * (G + 1)]) {
    constexpr int NQ = G + 1;
#pragma unroll
    const T* c2 = p + m.offC2 + o;
    const T* w2 = p + m.offW2 + o;
    for (int j = 0; j < H; ++j) {
#pragma unroll
        for (int q = 0; q < G; ++q) w[j * NQ + q] = c2[(j * G + q) * m.n];
        w[j * NQ + G] = w2[j * m.n];
    }
}

template <class T, int H, int G>
__global__ void __launch_bounds__(W_BT, 3) wide_l2_fwd_kernel(const __grid_constant__ WideModel m, const T* __restrict__ p, const T* hidden, T* out,
                                                           const int* mask, int64_t B, int btile) {
    constexpr int NQ = G + 1, NW = H * NQ;
    __shared__ __align__(16) T f2[NW][W_PT];              // features, IC index fastest: one vector load feeds 4 ICs
    const int tid = threadIdx.x, n = m.n;
    const int b0 = blockIdx.y * btile, b1 = (int)min((int64_t)(b0 + btile), B);
    for (int v = tid; v < W_PT * H; v += W_BT) {
        const int bl = v / H, j = v % H;
        T c[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) c[q] = T(0);
        if (b0 + bl < b1) w_features<T, G>(m.norm2, (T)m.inv_h2, m.grid2, hidden[(int64_t)(b0 + bl) * W_HP + j], c);
#pragma unroll
        for (int q = 0; q < NQ; ++q) f2[j * NQ + q][bl] = c[q];
    }
    __syncthreads();
    const int o = blockIdx.x * W_BT + tid;
    if (o >= n) return;
    T w[NW];
    wide_load_w2<T, H, G>(m, p, o, w);
    for (int bl0 = 0; b0 + bl0 < b1; bl0 += 4) {
        T acc[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
        for (int r = 0; r < NW; ++r) {
            T f[4];
            if constexpr (sizeof(T) == 4) { const float4 v = *reinterpret_cast<const float4*>(&f2[r][bl0]); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
            else { const double2 v0 = *reinterpret_cast<const double2*>(&f2[r][bl0]), v1 = *reinterpret_cast<const double2*>(&f2[r][bl0 + 2]); f[0] = v0.x; f[1] = v0.y; f[2] = v1.x; f[3] = v1.y; }
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] += w[r] * f[k];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int b = b0 + bl0 + k;
            if (b < b1 && (!mask || mask[b])) out[(int64_t)b * n + o] = acc[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// layer 2 forward on the 5th-generation tensor cores (fp32 instantiation): the expand contraction
//     out[o][b] = sum_r w2[o][r] * f2[b][r]        M = 128 output units per CTA, N = 32 ICs, K = H*(G+1) padded to 112
// is a dense GEMM tile.  tcgen05.mma kind::tf32 (FP32 accumulate in TMEM) with the 3xTF32 split
//     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo ,  x_hi = x with the low 13 mantissa bits cleared, x_lo = x - x_hi
// keeps ~22 mantissa bits (the fp32 parity budget is 1e-5).  The weight operand comes from a device image that is
// already in the UMMA shared-memory layout (K-major, no swizzle: [K/4][128 rows][4] = 8x16-byte core matrices), so one
// bulk async copy (TMA, mbarrier completion) stages it; the feature operand is produced by the CTA's threads in the
// same layout.  One elected thread issues the 14 x 3 MMAs, tcgen05.commit signals the epilogue, each warp reads its
// 32 TMEM lanes (tcgen05.ld 32x32b) and stores out[b][o] coalesced over o.
// ---------------------------------------------------------------------------------------------------------
constexpr int TC_M = 128, TC_N = 32;
__device__ __forceinline__ uint32_t w_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void w_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(w_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void w_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(w_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void w_tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(w_smem_u32(dst)), "l"(src), "r"(bytes), "r"(w_smem_u32(bar)) : "memory");
}
// bounded wait: a mis-programmed pipeline traps instead of hanging the GPU
__device__ __forceinline__ void w_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(w_smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && spin > (1u << 26)) __trap();
    }
}
// UMMA shared-memory descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B; lbo = byte distance between the two
// 16-byte K chunks of one MMA, sbo = byte distance between 8-row groups   [cute/arch/mma_sm100_desc.hpp: SmemDescriptor]
__device__ __forceinline__ uint64_t w_umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void w_umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float w_tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// device image of the layer-2 weights for the tensor-core kernel: [ceil(n/128)][hi|lo][KC][128][4]
template <int H, int G>
__global__ void __launch_bounds__(256) wide_w2_image_kernel(const __grid_constant__ WideModel m, const float* __restrict__ p, float* __restrict__ img) {
    constexpr int NQ = G + 1, NW = H * NQ, KC = (NW + 7) / 8 * 2;
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int nch = (m.n + TC_M - 1) / TC_M;
    if (idx >= (int64_t)nch * KC * TC_M * 4) return;
    const int kk = (int)(idx & 3), r = (int)((idx >> 2) % TC_M), kc = (int)((idx / (4 * TC_M)) % KC), cm = (int)(idx / ((int64_t)4 * TC_M * KC));
    const int o = cm * TC_M + r, k = kc * 4 + kk;
    float v = 0.f;
    if (o < m.n && k < NW) { const int j = k / NQ, q = k - j * NQ; v = q < G ? p[m.offC2 + (int64_t)(j * G + q) * m.n + o] : p[m.offW2 + (int64_t)j * m.n + o]; }
    const float hi = w_tf32_hi(v);
    float* base = img + (int64_t)cm * 2 * KC * TC_M * 4;
    base[((int64_t)kc * TC_M + r) * 4 + kk] = hi;
    base[(int64_t)KC * TC_M * 4 + ((int64_t)kc * TC_M + r) * 4 + kk] = v - hi;
}

template <int H, int G>
__global__ void __launch_bounds__(128) wide_l2_fwd_tc_kernel(const __grid_constant__ WideModel m, const float* __restrict__ img, const float* hidden,
                                                             float* out, const int* mask, int64_t B) {
    constexpr int NQ = G + 1, NW = H * NQ, KC = (NW + 7) / 8 * 2, KSTEPS = KC / 2;
    constexpr uint32_t A_BYTES = KC * TC_M * 16, B_BYTES = KC * TC_N * 16;
    extern __shared__ __align__(128) unsigned char tc_smem[];
    float* a_hi = reinterpret_cast<float*>(tc_smem);
    float* a_lo = reinterpret_cast<float*>(tc_smem + A_BYTES);
    float* b_hi = reinterpret_cast<float*>(tc_smem + 2 * A_BYTES);
    float* b_lo = reinterpret_cast<float*>(tc_smem + 2 * A_BYTES + B_BYTES);
    uint64_t* bar_a = reinterpret_cast<uint64_t*>(tc_smem + 2 * A_BYTES + 2 * B_BYTES);
    uint64_t* bar_mma = bar_a + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_a + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n = m.n;
    const int b0 = blockIdx.y * TC_N;
    if (tid == 0) {
        w_mbar_init(bar_a, 1); w_mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {        // TMEM: 32 fp32 accumulator columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(w_smem_u32(tmem_slot)), "n"(TC_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {         // weights: the image is already in the operand layout -> bulk copies, one mbarrier
        const float* src = img + (int64_t)blockIdx.x * 2 * KC * TC_M * 4;
        w_mbar_expect_tx(bar_a, 2 * A_BYTES);
        constexpr uint32_t PIECE = A_BYTES / 2;
#pragma unroll
        for (int c = 0; c < 4; ++c) w_tma_load_1d(tc_smem + c * PIECE, reinterpret_cast<const unsigned char*>(src) + c * PIECE, PIECE, bar_a);
    }
    // feature operand [KC][TC_N][4], hi and lo
    for (int v = tid; v < 2 * B_BYTES / 4; v += 128) b_hi[v] = 0.f;          // (b_lo follows b_hi)
    __syncthreads();
    for (int v = tid; v < TC_N * H; v += 128) {
        const int bl = v / H, j = v - bl * H;
        if (b0 + bl >= B) continue;
        float c[NQ];
        w_features<float, G>(m.norm2, m.inv_h2, m.grid2, hidden[(int64_t)(b0 + bl) * W_HP + j], c);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int k = j * NQ + q;
            const float hi = w_tf32_hi(c[q]);
            b_hi[((k >> 2) * TC_N + bl) * 4 + (k & 3)] = hi;
            b_lo[((k >> 2) * TC_N + bl) * 4 + (k & 3)] = c[q] - hi;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (warp == 0) {
        w_mbar_wait(bar_a, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
            // instruction descriptor: D=F32, A=B=TF32, both K-major, N>>3, M>>4   [cute/arch/mma_sm100_desc.hpp: InstrDescriptor]
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            const uint32_t ah = w_smem_u32(a_hi), al = w_smem_u32(a_lo), bh = w_smem_u32(b_hi), bl_ = w_smem_u32(b_lo);
#pragma unroll 1
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const uint32_t ao = ks * 2 * TC_M * 16, bo = ks * 2 * TC_N * 16;
                const uint64_t dah = w_umma_desc(ah + ao, TC_M * 16, 128), dal = w_umma_desc(al + ao, TC_M * 16, 128);
                const uint64_t dbh = w_umma_desc(bh + bo, TC_N * 16, 128), dbl = w_umma_desc(bl_ + bo, TC_N * 16, 128);
                w_umma_tf32(tmem, dal, dbh, idesc, ks > 0 ? 1u : 0u);        // small terms first
                w_umma_tf32(tmem, dah, dbl, idesc, 1u);
                w_umma_tf32(tmem, dah, dbh, idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(w_smem_u32(bar_mma)) : "memory");
        }
        __syncwarp();
    }
    w_mbar_wait(bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[TC_N];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                   "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int o = blockIdx.x * TC_M + tid;
    if (o < n) {
#pragma unroll
        for (int c = 0; c < TC_N; ++c) {
            const int b = b0 + c;
            if (b < B && (!mask || mask[b])) out[(int64_t)b * n + o] = __uint_as_float(v[c]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_N) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// layer 2 reverse (reduce over the n output units): hbar[b][j] = sum_o lam[b][o] * sum_q w2[o][j][q] * d_q(hidden[b][j])
// lam = lprev + h * sum a_sj kl_j is formed here (and stored: it is ybar of layer 2 in the stage record)
// ---------------------------------------------------------------------------------------------------------
template <class T, int H, int G>
__global__ void __launch_bounds__(W_BT, 3) wide_l2_vjp_kernel(const __grid_constant__ WideModel m, const T* __restrict__ p, const T* hidden,
                                                           const WideIn<T> in, int64_t B, int P, int btile, T* part, T* hbar, unsigned* counters) {
    constexpr int NQ = G + 1, NW = H * NQ, GB = sizeof(T) == 4 ? 8 : 4;
    __shared__ T red[GB * H][W_BT + 1];
    __shared__ __align__(16) T d2[NW][GB];                 // d c_q / d hidden_j, IC index fastest
    __shared__ int s_last;
    const int tid = threadIdx.x, n = m.n;
    const int b0 = blockIdx.y * btile, b1 = (int)min((int64_t)(b0 + btile), B);
    for (int v = tid; v < GB * H; v += W_BT) {
        const int bl = v / H, j = v % H;
        T d[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) d[q] = T(0);
        if (b0 + bl < b1) w_dfeatures<T, G>(m.norm2, (T)m.inv_h2, m.grid2, hidden[(int64_t)(b0 + bl) * W_HP + j], d);
#pragma unroll
        for (int q = 0; q < NQ; ++q) d2[j * NQ + q][bl] = d[q];
    }
    __syncthreads();
    for (int pass = 0; pass < P; ++pass) {
        const int o = (blockIdx.x * P + pass) * W_BT + tid;
        const bool valid = o < n;
        T lam[GB]; bool on[GB];
        wide_inputs<T, 0, GB>(in, b0, b1, o, valid, n, B, lam, on);
        T w[NW];
        if (valid) wide_load_w2<T, H, G>(m, p, o, w);
        else {
#pragma unroll
            for (int k = 0; k < NW; ++k) w[k] = T(0);
        }
        if (in.xstore) {
#pragma unroll
            for (int bl = 0; bl < GB; ++bl) if (on[bl]) in.xstore[(b0 + bl) * n + o] = lam[bl];
        }
#pragma unroll
        for (int j = 0; j < H; ++j) {
            T inner[GB];
#pragma unroll
            for (int bl = 0; bl < GB; ++bl) inner[bl] = T(0);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const T wv = w[j * NQ + q];
                T d[GB];
                if constexpr (sizeof(T) == 4) {
                    const float4 v0 = *reinterpret_cast<const float4*>(&d2[j * NQ + q][0]), v1 = *reinterpret_cast<const float4*>(&d2[j * NQ + q][4]);
                    d[0] = v0.x; d[1] = v0.y; d[2] = v0.z; d[3] = v0.w; d[4] = v1.x; d[5] = v1.y; d[6] = v1.z; d[7] = v1.w;
                } else {
                    const double2 v0 = *reinterpret_cast<const double2*>(&d2[j * NQ + q][0]), v1 = *reinterpret_cast<const double2*>(&d2[j * NQ + q][2]);
                    d[0] = v0.x; d[1] = v0.y; d[2] = v1.x; d[3] = v1.y;
                }
#pragma unroll
                for (int bl = 0; bl < GB; bl += 2) kfma2b(inner[bl], inner[bl + 1], d[bl], d[bl + 1], wv);
            }
#pragma unroll
            for (int bl = 0; bl < GB; ++bl) {
                const T a = lam[bl] * inner[bl];
                if (pass == 0) red[bl * H + j][tid] = a;
                else red[bl * H + j][tid] += a;
            }
        }
    }
    wide_reduce_tail<T, H>(red, b0, b1, B, part, hbar, counters, &s_last, in.mask);
}

// ---------------------------------------------------------------------------------------------------------
// layer 1 forward on tcgen05 (fp32, n % 64 == 0) for the (stage, IC) pairs of a backward attempt — the fused
// basis-expansion + contraction kernel:  hidden[v][o] = sum_i sum_q c_q(y_v[i]) * w1[i][q][o]
//   M = 128 rows v (one thread each), N = 16 (H padded), K = 8 units x KU features per pass (KU = G+1 padded to a multiple of 4).
// Every thread forms sol(t_s) of its row for the pass's 8 units (vector loads from the dense record), expands the RBF / SiLU
// features in registers and writes them — TF32 hi/lo split — straight into the K-major UMMA operand layout in shared memory
// (one conflict-free STS.128 per 4 features); the features never exist in HBM.  The weight operand of the pass (12 KB) arrives
// by TMA from an image in the same layout.  D accumulates in TMEM over the block's passes; partial sums over the unit blocks
// go to wide_sum_partials_kernel.  Two blocks per SM overlap one block's feature generation with the other's MMAs.
// ---------------------------------------------------------------------------------------------------------
constexpr int TC1_UC = 8, TC1_N = 16;
template <int H, int G>
__global__ void __launch_bounds__(256) wide_w1_image_kernel(const __grid_constant__ WideModel m, const float* __restrict__ p, float* __restrict__ img) {
    constexpr int NQ = G + 1, KU = (NQ + 3) / 4 * 4, KC = TC1_UC * KU / 4;
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int nblk = m.n / TC1_UC;
    if (idx >= (int64_t)nblk * KC * TC1_N * 4) return;
    const int kk = (int)(idx & 3), o = (int)((idx >> 2) % TC1_N), kc = (int)((idx / (4 * TC1_N)) % KC), ub = (int)(idx / ((int64_t)4 * TC1_N * KC));
    const int k = kc * 4 + kk, u = k / KU, q = k - u * KU;
    const int64_t i = (int64_t)ub * TC1_UC + u;
    float v = 0.f;
    if (o < H && q < NQ) v = q < G ? p[m.offC1 + (i * G + q) * H + o] : p[m.offW1 + i * H + o];
    const float hi = w_tf32_hi(v);
    float* base = img + (int64_t)ub * 2 * KC * TC1_N * 4;
    base[((int64_t)kc * TC1_N + o) * 4 + kk] = hi;
    base[(int64_t)KC * TC1_N * 4 + ((int64_t)kc * TC1_N + o) * 4 + kk] = v - hi;
}

// 256 threads: thread = (row v, half of the pass's 8 units).  Two operand buffers: the features of pass p+1 are generated while
// the MMAs of pass p run (tcgen05.commit -> one mbarrier per buffer guards the overwrite two passes later).
template <int H, int G>
__global__ void __launch_bounds__(256, 1) wide_l1_fwd_tc_kernel(const __grid_constant__ WideModel m, const float* __restrict__ img, const WideIn<float> in,
                                                                int64_t BV, int P, float* part) {
    constexpr int NQ = G + 1, KU = (NQ + 3) / 4 * 4, K = TC1_UC * KU, KC = K / 4, KSTEPS = KC / 2, CU = KU / 4, UH = TC1_UC / 2;
    constexpr uint32_t A_BYTES = KC * TC_M * 16, B_BYTES = KC * TC1_N * 16, STAGE = 2 * A_BYTES + 2 * B_BYTES;
    extern __shared__ __align__(128) unsigned char tc_smem[];
    uint64_t* bar_b = reinterpret_cast<uint64_t*>(tc_smem + 2 * STAGE);      // [2]
    uint64_t* bar_mma = bar_b + 2;                                          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_b + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n = m.n;
    const int row = tid & (TC_M - 1), uh = tid >> 7;                        // row of the tile, which 4 of the 8 units
    const int64_t v = (int64_t)blockIdx.y * TC_M + row;                      // this thread's (stage, IC) pair
    const bool on = v < BV && (!in.mask || in.mask[v]);
    const int nblk = n / TC1_UC;
    const int npass = min(P, nblk - (int)blockIdx.x * P);
    if (tid == 0) {
        w_mbar_init(bar_b, 1); w_mbar_init(bar_b + 1, 1); w_mbar_init(bar_mma, 1); w_mbar_init(bar_mma + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(w_smem_u32(tmem_slot)), "n"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // row constants: dense-record row and interpolation weights of sol(t_s)
    float bw[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, hd = 0.f;
    const float* rrow = in.rec;
    if (on) {
        interp_weights(in.th[v], bw);
        hd = in.hd[v];
        const int64_t brow = in.brec > 0 ? v % in.brec : v;
        rrow = in.rec + (brow * in.cap + in.ridx[v]) * 8 * (int64_t)n + uh * UH;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const float inv_h = m.inv_h1;
    float4 r0[8];                                                            // dense-record values of this thread's 4 units (8 arrays)
    auto load_raw = [&](int pass) {
        const int i0 = (blockIdx.x * P + pass) * TC1_UC;
#pragma unroll
        for (int j = 0; j < 8; ++j) r0[j] = *reinterpret_cast<const float4*>(rrow + (int64_t)j * n + i0);
    };
    if (on) load_raw(0);
    for (int pass = 0; pass < npass; ++pass) {
        const int sb = pass & 1, ub = blockIdx.x * P + pass, i0 = ub * TC1_UC + uh * UH;
        float* a_hi = reinterpret_cast<float*>(tc_smem + sb * STAGE);
        float* a_lo = reinterpret_cast<float*>(tc_smem + sb * STAGE + A_BYTES);
        float* b_hi = reinterpret_cast<float*>(tc_smem + sb * STAGE + 2 * A_BYTES);
        float x[UH];
        if (on) {
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 7; ++j) { a0.x += bw[j] * r0[1 + j].x; a0.y += bw[j] * r0[1 + j].y; a0.z += bw[j] * r0[1 + j].z; a0.w += bw[j] * r0[1 + j].w; }
            x[0] = r0[0].x + hd * a0.x; x[1] = r0[0].y + hd * a0.y; x[2] = r0[0].z + hd * a0.z; x[3] = r0[0].w + hd * a0.w;
            if (pass + 1 < npass) load_raw(pass + 1);                        // in flight behind the feature expansion below
        }
        if (pass >= 2) { w_mbar_wait(bar_mma + sb, ((pass >> 1) - 1) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        if (tid == 0) {
            w_mbar_expect_tx(bar_b + sb, 2 * B_BYTES);
            w_tma_load_1d(b_hi, img + (int64_t)ub * 2 * KC * TC1_N * 4, 2 * B_BYTES, bar_b + sb);
        }
        if (on) {
            if (in.xstore) *reinterpret_cast<float4*>(in.xstore + v * n + i0) = make_float4(x[0], x[1], x[2], x[3]);
#pragma unroll
            for (int u = 0; u < UH; ++u) {
                float c[KU];
                {
                    float cc[NQ];
                    w_features<float, G>(m.norm1, inv_h, m.grid1, x[u], cc);
#pragma unroll
                    for (int q = 0; q < KU; ++q) c[q] = q < NQ ? cc[q] : 0.f;
                }
#pragma unroll
                for (int cq = 0; cq < CU; ++cq) {
                    const float4 f = make_float4(c[4 * cq], c[4 * cq + 1], c[4 * cq + 2], c[4 * cq + 3]);
                    const float4 hi = make_float4(w_tf32_hi(f.x), w_tf32_hi(f.y), w_tf32_hi(f.z), w_tf32_hi(f.w));
                    const float4 lo = make_float4(f.x - hi.x, f.y - hi.y, f.z - hi.z, f.w - hi.w);
                    const int kc = (uh * UH + u) * CU + cq;
                    *reinterpret_cast<float4*>(a_hi + (kc * TC_M + row) * 4) = hi;
                    *reinterpret_cast<float4*>(a_lo + (kc * TC_M + row) * 4) = lo;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp == 0) {
            w_mbar_wait(bar_b + sb, (pass >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC1_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
                const uint32_t ah = w_smem_u32(a_hi), al = w_smem_u32(a_lo), bh = w_smem_u32(b_hi), blo = bh + B_BYTES;
#pragma unroll 1
                for (int ks = 0; ks < KSTEPS; ++ks) {
                    const uint32_t ao = ks * 2 * TC_M * 16, bo = ks * 2 * TC1_N * 16;
                    const uint64_t dah = w_umma_desc(ah + ao, TC_M * 16, 128), dal = w_umma_desc(al + ao, TC_M * 16, 128);
                    const uint64_t dbh = w_umma_desc(bh + bo, TC1_N * 16, 128), dbl = w_umma_desc(blo + bo, TC1_N * 16, 128);
                    w_umma_tf32(tmem, dal, dbh, idesc, (pass > 0 || ks > 0) ? 1u : 0u);
                    w_umma_tf32(tmem, dah, dbl, idesc, 1u);
                    w_umma_tf32(tmem, dah, dbh, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(w_smem_u32(bar_mma + sb)) : "memory");
            }
            __syncwarp();
        }
    }
    w_mbar_wait(bar_mma + ((npass - 1) & 1), ((npass - 1) >> 1) & 1);       // the last commit covers every earlier MMA
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        uint32_t d[TC1_N];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]), "=r"(d[10]),
                       "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (v < BV) {
#pragma unroll
            for (int o = 0; o < H; ++o) part[((int64_t)blockIdx.x * BV + v) * H + o] = __uint_as_float(d[o]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(32) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// layer 2 reverse on tcgen05 (fp32, n % 128 == 0):  S[r][b] = sum_o W2t[r][o] * lam[b][o]   (M = 128 rows r = (j,q), 110 used;
// N = 32 ICs; K = o, 128 per pass, accumulated in TMEM over the P passes of the block), then hbar[b][j] = sum_q S[(j,q)][b] * d2[b][(j,q)].
// A (weights) arrives by TMA from an image in the UMMA layout; B (lambda at the stage = lam + h*sum a*kl, also recorded as
// ybar of layer 2) is formed by the threads with coalesced 16-byte loads and written hi/lo in the K-major core-matrix layout.
// ---------------------------------------------------------------------------------------------------------
constexpr int TC_KB = 64;            // output units (K of the reverse contraction) per block
template <int H, int G>
__global__ void __launch_bounds__(256) wide_w2t_image_kernel(const __grid_constant__ WideModel m, const float* __restrict__ p, float* __restrict__ img) {
    constexpr int NQ = G + 1, NW = H * NQ, KC = TC_KB / 4;
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int nblk = m.n / TC_KB;
    if (idx >= (int64_t)nblk * KC * TC_M * 4) return;
    const int kk = (int)(idx & 3), r = (int)((idx >> 2) % TC_M), kc = (int)((idx / (4 * TC_M)) % KC), ob = (int)(idx / ((int64_t)4 * TC_M * KC));
    const int o = ob * TC_KB + kc * 4 + kk;
    float v = 0.f;
    if (r < NW) { const int j = r / NQ, q = r - j * NQ; v = q < G ? p[m.offC2 + (int64_t)(j * G + q) * m.n + o] : p[m.offW2 + (int64_t)j * m.n + o]; }
    const float hi = w_tf32_hi(v);
    float* base = img + (int64_t)ob * 2 * KC * TC_M * 4;
    base[((int64_t)kc * TC_M + r) * 4 + kk] = hi;
    base[(int64_t)KC * TC_M * 4 + ((int64_t)kc * TC_M + r) * 4 + kk] = v - hi;
}

// grid (n / 64, ceil(B / 32)); 94 KB of shared memory -> two blocks per SM overlap each other's load and MMA phases
template <int H, int G>
__global__ void __launch_bounds__(128) wide_l2_vjp_tc_kernel(const __grid_constant__ WideModel m, const float* __restrict__ img, const float* hidden,
                                                             const WideIn<float> in, int64_t B, float* part) {
    constexpr int NQ = G + 1, NW = H * NQ, NWP = (NW + 3) / 4 * 4, KC = TC_KB / 4, KSTEPS = KC / 2;
    constexpr uint32_t A_BYTES = KC * TC_M * 16, B_BYTES = KC * TC_N * 16;
    extern __shared__ __align__(128) unsigned char tc_smem[];
    float* a_hi = reinterpret_cast<float*>(tc_smem);
    float* a_lo = reinterpret_cast<float*>(tc_smem + A_BYTES);
    float* b_hi = reinterpret_cast<float*>(tc_smem + 2 * A_BYTES);
    float* b_lo = reinterpret_cast<float*>(tc_smem + 2 * A_BYTES + B_BYTES);
    float* d2 = reinterpret_cast<float*>(tc_smem + 2 * A_BYTES + 2 * B_BYTES);                 // [TC_N][NWP]
    uint64_t* bar_a = reinterpret_cast<uint64_t*>(tc_smem + 2 * A_BYTES + 2 * B_BYTES + TC_N * NWP * 4);
    uint64_t* bar_mma = bar_a + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_a + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n = m.n;
    const int b0 = blockIdx.y * TC_N, b1 = (int)min((int64_t)(b0 + TC_N), B);
    const int ob = blockIdx.x;
    if (tid == 0) {
        w_mbar_init(bar_a, 1); w_mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(w_smem_u32(tmem_slot)), "n"(TC_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(img + (int64_t)ob * 2 * KC * TC_M * 4);
        w_mbar_expect_tx(bar_a, 2 * A_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c) w_tma_load_1d(tc_smem + c * (A_BYTES / 2), src + c * (A_BYTES / 2), A_BYTES / 2, bar_a);
    }
    {   // lambda operand: lane -> (IC, 16-byte chunk): 64 B contiguous per IC in global memory, conflict-free STS.128
        const int bl_l = lane & 7, c_l = lane >> 3;
        const int Bn = (int)B * n;
        float4 base[4], kv[4][6]; float hs[4]; bool on[4]; int bl[4];
        const int ch = c_l + 4 * warp;                                                  // 0..15
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bl[k] = bl_l + 8 * k;
            const int b = b0 + bl[k];
            on[k] = b < b1 && (!in.mask || in.mask[b]);
            const int e = b * n + ob * TC_KB + ch * 4;
            base[k] = on[k] ? *reinterpret_cast<const float4*>(in.base + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            hs[k] = on[k] && in.ncoef > 0 ? in.hs[b] : 0.f;
#pragma unroll
            for (int j = 0; j < 6; ++j)
                kv[k][j] = (on[k] && j < in.ncoef) ? *reinterpret_cast<const float4*>(in.ks + (int64_t)j * Bn + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // per-IC derivative table while the loads are in flight
        for (int v = tid; v < TC_N * H; v += 128) {
            const int bq = v / H, j = v - bq * H;
            float d[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) d[q] = 0.f;
            if (b0 + bq < b1) w_dfeatures<float, G>(m.norm2, m.inv_h2, m.grid2, hidden[(int64_t)(b0 + bq) * W_HP + j], d);
#pragma unroll
            for (int q = 0; q < NQ; ++q) d2[bq * NWP + j * NQ + q] = d[q];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 6; ++j) { const float cj = in.coef[j]; acc.x += cj * kv[k][j].x; acc.y += cj * kv[k][j].y; acc.z += cj * kv[k][j].z; acc.w += cj * kv[k][j].w; }
            float4 lam = base[k];
            if (in.ncoef > 0) { lam.x = base[k].x + hs[k] * acc.x; lam.y = base[k].y + hs[k] * acc.y; lam.z = base[k].z + hs[k] * acc.z; lam.w = base[k].w + hs[k] * acc.w; }
            if (on[k] && in.xstore) *reinterpret_cast<float4*>(in.xstore + (b0 + bl[k]) * n + ob * TC_KB + ch * 4) = lam;
            const float4 hi = make_float4(w_tf32_hi(lam.x), w_tf32_hi(lam.y), w_tf32_hi(lam.z), w_tf32_hi(lam.w));
            const float4 lo = make_float4(lam.x - hi.x, lam.y - hi.y, lam.z - hi.z, lam.w - hi.w);
            *reinterpret_cast<float4*>(b_hi + (ch * TC_N + bl[k]) * 4) = hi;
            *reinterpret_cast<float4*>(b_lo + (ch * TC_N + bl[k]) * 4) = lo;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (warp == 0) {
        w_mbar_wait(bar_a, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            const uint32_t ah = w_smem_u32(a_hi), al = w_smem_u32(a_lo), bh = w_smem_u32(b_hi), blo = w_smem_u32(b_lo);
#pragma unroll 1
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const uint32_t ao = ks * 2 * TC_M * 16, bo = ks * 2 * TC_N * 16;
                const uint64_t dah = w_umma_desc(ah + ao, TC_M * 16, 128), dal = w_umma_desc(al + ao, TC_M * 16, 128);
                const uint64_t dbh = w_umma_desc(bh + bo, TC_N * 16, 128), dbl = w_umma_desc(blo + bo, TC_N * 16, 128);
                w_umma_tf32(tmem, dal, dbh, idesc, ks > 0 ? 1u : 0u);
                w_umma_tf32(tmem, dah, dbl, idesc, 1u);
                w_umma_tf32(tmem, dah, dbh, idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(w_smem_u32(bar_mma)) : "memory");
        }
        __syncwarp();
    }
    w_mbar_wait(bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[TC_N];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                   "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    // S[r][b] * d2[b][r] staged in shared memory (the operand buffers are free now), then summed over q per hidden unit
    float* ps = a_hi;                                             // [128][33]
    if (tid < NW) {
#pragma unroll
        for (int c = 0; c < TC_N; ++c) ps[tid * 33 + c] = __uint_as_float(v[c]) * d2[c * NWP + tid];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    for (int idx = tid; idx < TC_N * H; idx += 128) {
        const int j = idx / TC_N, bq = idx - j * TC_N;
        if (b0 + bq >= b1) continue;
        float sacc = 0.f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) sacc += ps[(j * NQ + q) * 33 + bq];
        part[((int64_t)ob * B + b0 + bq) * H + j] = sacc;
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_N) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// layer 1 reverse: dl[b][i] = -( dnorm * sum_g db_g/h * (w1[g][:] . hbar[b]) + dswish * (W1[i][:] . hbar[b]) )
// ---------------------------------------------------------------------------------------------------------
template <class T, int H, int G>
__global__ void __launch_bounds__(W_BT, 2) wide_l1_vjp_kernel(const __grid_constant__ WideModel m, const T* __restrict__ w1t, const T* x1, const T* hbar,
                                                           T* dl, const int* mask, int64_t B, int btile) {
    constexpr int NW = H * (G + 1), PF = 8;
    __shared__ T hb[W_PT][W_HP];
    const int tid = threadIdx.x, n = m.n;
    const int b0 = blockIdx.y * btile, b1 = (int)min((int64_t)(b0 + btile), B);
    for (int v = tid; v < W_PT * W_HP; v += W_BT) {
        const int bl = v / W_HP, o = v % W_HP;
        hb[bl][o] = (o < H && b0 + bl < b1) ? hbar[(int64_t)(b0 + bl) * W_HP + o] : T(0);
    }
    __syncthreads();
    const int i = blockIdx.x * W_BT + tid;
    if (i >= n) return;
    T w[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) w[k] = w1t[(int64_t)k * n + i];
    const T inv_h = (T)m.inv_h1;
    for (int bl0 = 0; b0 + bl0 < b1; bl0 += PF) {
        T xs[PF]; bool on[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int b = b0 + bl0 + k;
            on[k] = b < b1 && (!mask || mask[b]);
            xs[k] = on[k] ? x1[(int64_t)b * n + i] : T(0);
        }
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            if (!on[k]) continue;
            const T* yb = hb[bl0 + k];
            const T x = xs[k];
            const T xn = normalize_rt(m.norm1, x);
            T xnbar = T(0);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const T a = (xn - (T)m.grid1[g]) * inv_h;
                const T y = kexp(-a * a);
                const T db = T(-2) * a * y;
                T bbar = T(0);
                if constexpr (sizeof(T) == 4) {
                    T b1 = T(0);
#pragma unroll
                    for (int o = 0; o < H; o += 2) kfma2(bbar, b1, w[g * H + o], w[g * H + o + 1], yb[o], yb[o + 1]);
                    bbar += b1;
                } else {
#pragma unroll
                    for (int o = 0; o < H; ++o) bbar += w[g * H + o] * yb[o];
                }
                xnbar += db * inv_h * bbar;
            }
            T xb = xnbar * normalize_deriv_rt(m.norm1, xn);
            T sw, ds; swish_both(x, sw, ds);
            T sbar = T(0);
            if constexpr (sizeof(T) == 4) {
                T s1 = T(0);
#pragma unroll
                for (int o = 0; o < H; o += 2) kfma2(sbar, s1, w[G * H + o], w[G * H + o + 1], yb[o], yb[o + 1]);
                sbar += s1;
            } else {
#pragma unroll
                for (int o = 0; o < H; ++o) sbar += w[G * H + o] * yb[o];
            }
            xb += sbar * ds;
            dl[(int64_t)(b0 + bl0 + k) * n + i] = -xb;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// step-end pass over the gradient state g[b] (and the initdt norms of dg/dt)
//   MODE 0: g_new = g_old + sum_s wb_s k_s, es += (sum_s wbt_s k_s / (abstol + max(|g_old|,|g_new|) reltol))^2
//   MODE 1: es += (k_0 / abstol)^2                 MODE 2: es += ((k_1 - k_0) / abstol)^2
// with k_s[(i,q),o] = ybar_s[o] * c_q(x_s[i]) rebuilt from the stage records.
// ---------------------------------------------------------------------------------------------------------
template <class T> struct WideGp {
    const T* x1;    // [7][B][n]     layer-1 inputs  (y at the stage times)
    const T* yb1;   // [7][B][W_HP]  layer-1 output cotangents
    const T* x2;    // [7][B][W_HP]  layer-2 inputs  (hidden)
    const T* yb2;   // [7][B][n]     layer-2 output cotangents (lambda at the stages)
    T* g;           // [B][2][np]
    const int* cur; const T* h; const int* mask;
    T abstol, reltol;
    T* es_part; int npart, off;   // es_part[b * npart + off + block]
};

template <class T> __device__ __forceinline__ void gp_finalize(const T* gold, T* gnew, int64_t j, T vb, T vt, T abstol, T reltol, T& es) {
    const T g0 = gold[j];
    const T g1 = g0 + vb;
    const T sc = abstol + kmax(kabs(g0), kabs(g1)) * reltol;
    const T r = vt / sc;
    es += r * r;
    gnew[j] = g1;
}

// Layer 1: its slice of g is one contiguous array of (n*G + n) rows x H (C1 rows (i,q), then the W1 rows i), streamed with
// coalesced 2-element accesses.  A thread's output pair o0 = (2*tid) % H is the same in every iteration (2*W_GT % H == 0),
// so its 2 x 7 x 2 weighted cotangents stay in registers; the 7 stage features of a row come from shared memory.
constexpr int W_GT = 320, W_GK = 10;      // threads per block, iterations per block (tile = W_GK * 2 * W_GT elements)
template <class T> struct alignas(2 * sizeof(T)) WVec2 { T x, y; };

// grid.x = nblkC blocks over the C1 rows (tiles aligned to whole input units) followed by the blocks over the W1 rows
template <class T, int H, int G, int MODE>
__global__ void __launch_bounds__(W_GT) wide_gp1_kernel(const __grid_constant__ WideModel m, const WideGp<T> a, int64_t B, int nblkC) {
    static_assert(H % 2 == 0 && (2 * W_GT) % H == 0, "pair mapping");
    constexpr int NS = MODE == 0 ? 7 : (MODE == 1 ? 1 : 2);
    constexpr int ROWS = W_GK * 2 * W_GT / H, UNITS = ROWS / G;
    static_assert(ROWS % G == 0, "tiles hold whole input units");
    __shared__ T C[NS][ROWS];
    __shared__ T sred[33];
    const int b = blockIdx.y, tid = threadIdx.x, n = m.n;
    if (a.mask && !a.mask[b]) return;
    const bool segW = (int)blockIdx.x >= nblkC;
    const int row0 = (segW ? (int)blockIdx.x - nblkC : (int)blockIdx.x) * ROWS;       // row within the segment
    const int nrows = segW ? n : n * G;
    const T inv_h = (T)m.inv_h1;
    // g_old of this thread's 5 element pairs first: the loads are in flight while the block builds its feature tile
    const int64_t seg = segW ? m.offW1 : m.offC1;
    const T* gold = nullptr; T* gnew = nullptr;
    if (MODE == 0) { const int cur = a.cur[b]; gold = a.g + ((int64_t)b * 2 + cur) * m.np + seg + (int64_t)row0 * H; gnew = a.g + ((int64_t)b * 2 + (cur ^ 1)) * m.np + seg + (int64_t)row0 * H; }
    const int E = (nrows - row0) * H;             // elements of this tile that exist
    WVec2<T> g0[W_GK];
    if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < W_GK; ++k) {
            const int el = 2 * tid + k * 2 * W_GT;
            g0[k] = el < E ? *reinterpret_cast<const WVec2<T>*>(gold + el) : WVec2<T>{T(0), T(0)};
        }
    }
    if (!segW) {
        const int i0 = row0 / G;
        for (int idx = tid; idx < NS * UNITS; idx += W_GT) {
            const int s = idx / UNITS, il = idx - s * UNITS, i = i0 + il;
            if (i < n) {
                const T xn = normalize_rt(m.norm1, a.x1[((int64_t)s * B + b) * n + i]);
#pragma unroll
                for (int q = 0; q < G; ++q) { const T aa = (xn - (T)m.grid1[q]) * inv_h; C[s][il * G + q] = kexp(-aa * aa); }
            } else {
#pragma unroll
                for (int q = 0; q < G; ++q) C[s][il * G + q] = T(0);
            }
        }
    } else {
        for (int idx = tid; idx < NS * ROWS; idx += W_GT) {
            const int s = idx / ROWS, rl = idx - s * ROWS, i = row0 + rl;
            T c = T(0);
            if (i < n) swish_fwd(a.x1[((int64_t)s * B + b) * n + i], c);
            C[s][rl] = c;
        }
    }
    const int o0 = (2 * tid) % H;
    T ab[NS][2], at[NS][2];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const T yb = a.yb1[((int64_t)s * B + b) * W_HP + o0 + k];
            if (MODE == 0) { const T hh = a.h[b]; ab[s][k] = (-hh * Tab<T>::b(s)) * yb; at[s][k] = (-hh * Tab<T>::bt(s)) * yb; }
            else { ab[s][k] = yb; at[s][k] = T(0); }
        }
    __syncthreads();
    T es = T(0);
    const int rl0 = (2 * tid) / H;
#pragma unroll
    for (int k = 0; k < W_GK; ++k) {
        const int el = 2 * tid + k * 2 * W_GT;
        if (el >= E) break;
        const int rl = rl0 + k * (2 * W_GT / H);              // == el / H: a thread's rows are 64 apart
        if (MODE == 0) {
            T vb0 = T(0), vb1 = T(0), vt0 = T(0), vt1 = T(0);
#pragma unroll
            for (int s = 0; s < 7; ++s) { const T c = C[s][rl]; kfma2b(vb0, vb1, ab[s][0], ab[s][1], c); kfma2b(vt0, vt1, at[s][0], at[s][1], c); }
            WVec2<T> g1;
            g1.x = g0[k].x + vb0; g1.y = g0[k].y + vb1;
            const T r0 = wdiv(vt0, a.abstol + kmax(kabs(g0[k].x), kabs(g1.x)) * a.reltol);
            const T r1 = wdiv(vt1, a.abstol + kmax(kabs(g0[k].y), kabs(g1.y)) * a.reltol);
            es += r0 * r0; es += r1 * r1;
            *reinterpret_cast<WVec2<T>*>(gnew + el) = g1;
        } else if (MODE == 1) {
            const T x0 = (ab[0][0] * C[0][rl]) / a.abstol, x1 = (ab[0][1] * C[0][rl]) / a.abstol;
            es += x0 * x0; es += x1 * x1;
        } else {
            const T x0 = (ab[NS - 1][0] * C[NS - 1][rl] - ab[0][0] * C[0][rl]) / a.abstol;
            const T x1 = (ab[NS - 1][1] * C[NS - 1][rl] - ab[0][1] * C[0][rl]) / a.abstol;
            es += x0 * x0; es += x1 * x1;
        }
    }
    es = wblock_sum<T>(es, sred);
    if (tid == 0) a.es_part[(int64_t)b * a.npart + a.off + blockIdx.x] = es;
}

// Layer 2: rows (j,q) of length n; thread = output unit o walks all H*(G+1) rows (coalesced over o); the 7 stage features of a
// row are uniform over the block (shared memory, vector loads), the weighted cotangents of o stay in registers.
template <class T, int H, int G, int MODE>
__global__ void __launch_bounds__(W_BT) wide_gp2_kernel(const __grid_constant__ WideModel m, const WideGp<T> a, int64_t B) {
    constexpr int NQ = G + 1, NW = H * NQ, NS = MODE == 0 ? 7 : (MODE == 1 ? 1 : 2);
    __shared__ __align__(16) T F[NW][8];
    __shared__ T sred[33];
    const int b = blockIdx.y, tid = threadIdx.x, n = m.n;
    if (a.mask && !a.mask[b]) return;
    for (int v = tid; v < 8 * H; v += W_BT) {
        const int s = v / H, j = v % H;
        T c[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) c[q] = T(0);
        if (s < NS) w_features<T, G>(m.norm2, (T)m.inv_h2, m.grid2, a.x2[((int64_t)s * B + b) * W_HP + j], c);
#pragma unroll
        for (int q = 0; q < NQ; ++q) F[j * NQ + q][s] = c[q];
    }
    __syncthreads();
    const int o = blockIdx.x * W_BT + tid;
    T es = T(0);
    if (o < n) {
        T al[NS], alt[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const T yb = a.yb2[((int64_t)s * B + b) * n + o];
            if (MODE == 0) { const T hh = a.h[b]; al[s] = (-hh * Tab<T>::b(s)) * yb; alt[s] = (-hh * Tab<T>::bt(s)) * yb; }
            else { al[s] = yb; alt[s] = T(0); }
        }
        const T* gold = nullptr; T* gnew = nullptr;
        if (MODE == 0) { const int cur = a.cur[b]; gold = a.g + ((int64_t)b * 2 + cur) * m.np + m.offC2 + o; gnew = a.g + ((int64_t)b * 2 + (cur ^ 1)) * m.np + m.offC2 + o; }
#pragma unroll 2
        for (int j = 0; j < H; ++j) {
            T g0[NQ];
            if (MODE == 0) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) g0[q] = gold[(int64_t)(q < G ? j * G + q : H * G + j) * n];
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const T* f = F[j * NQ + q];
                if (MODE == 0) {
                    T fs[8];
                    if constexpr (sizeof(T) == 4) {
                        const float4 v0 = *reinterpret_cast<const float4*>(f), v1 = *reinterpret_cast<const float4*>(f + 4);
                        fs[0] = v0.x; fs[1] = v0.y; fs[2] = v0.z; fs[3] = v0.w; fs[4] = v1.x; fs[5] = v1.y; fs[6] = v1.z; fs[7] = v1.w;
                    } else {
#pragma unroll
                        for (int s = 0; s < 8; ++s) fs[s] = f[s];
                    }
                    T vb = T(0), vt = T(0);
#pragma unroll
                    for (int s = 0; s < 7; ++s) kfma2b(vb, vt, al[s], alt[s], fs[s]);
                    const T g1 = g0[q] + vb;
                    const T r = wdiv(vt, a.abstol + kmax(kabs(g0[q]), kabs(g1)) * a.reltol);
                    es += r * r;
                    gnew[(int64_t)(q < G ? j * G + q : H * G + j) * n] = g1;
                } else if (MODE == 1) {
                    const T x1 = (al[0] * f[0]) / a.abstol;
                    es += x1 * x1;
                } else {
                    const T x2 = (al[NS - 1] * f[NS - 1] - al[0] * f[0]) / a.abstol;
                    es += x2 * x2;
                }
            }
        }
    }
    es = wblock_sum<T>(es, sred);
    if (tid == 0) a.es_part[(int64_t)b * a.npart + a.off + blockIdx.x] = es;
}

// w1t[(q*H + o)][i]: layer-1 weights transposed so that a warp of consecutive input units loads them coalesced
template <class T, int H, int G>
__global__ void __launch_bounds__(256) wide_transpose_w1_kernel(const __grid_constant__ WideModel m, const T* __restrict__ p, T* __restrict__ w1t) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;       // over the source layout (coalesced reads)
    const int64_t nC = (int64_t)m.n * G * H, nW = (int64_t)m.n * H;
    if (idx < nC) {
        const int64_t i = idx / (G * H); const int r = (int)(idx - i * G * H);       // r = q*H + o
        w1t[(int64_t)r * m.n + i] = p[m.offC1 + idx];
    } else if (idx < nC + nW) {
        const int64_t k = idx - nC; const int64_t i = k / H; const int o = (int)(k - i * H);
        w1t[(int64_t)(G * H + o) * m.n + i] = p[m.offW1 + k];
    }
}

// ---------------------------------------------------------------------------------------------------------
// elementwise kernels over [B][n]; grid (ceil(n / W_ET), B)
// ---------------------------------------------------------------------------------------------------------
// norms of initdt: MODE 1: v0 = (u/sk)^2, v1 = (k0/sk)^2 ; MODE 2: v1 = ((k1-k0)/sk)^2    sk = abstol + |u| reltol
template <class T, int MODE>
__global__ void __launch_bounds__(W_ET) wide_norm_kernel(const T* u, const T* k0, const T* k1, int n, T abstol, T reltol, const int* mask,
                                                         T* part0, T* part1, int npart, int off) {
    __shared__ T sred[33];
    const int b = blockIdx.y, i = blockIdx.x * W_ET + threadIdx.x;
    if (mask && !mask[b]) return;
    T v0 = T(0), v1 = T(0);
    if (i < n) {
        const int64_t e = (int64_t)b * n + i;
        const T sk = abstol + kabs(u[e]) * reltol;
        if (MODE == 1) { const T x0 = u[e] / sk, x1 = k0[e] / sk; v0 = x0 * x0; v1 = x1 * x1; }
        else { const T x = (k1[e] - k0[e]) / sk; v1 = x * x; }
    }
    if (MODE == 1) { v0 = wblock_sum<T>(v0, sred); if (threadIdx.x == 0) part0[(int64_t)b * npart + off + blockIdx.x] = v0; }
    v1 = wblock_sum<T>(v1, sred);
    if (threadIdx.x == 0) part1[(int64_t)b * npart + off + blockIdx.x] = v1;
}

// embedded error of the state part: es += ((h sum_j bt_j k_j) / (abstol + max(|uprev|,|unew|) reltol))^2
template <class T>
__global__ void __launch_bounds__(W_ET) wide_err_kernel(const T* uprev, const T* unew, const T* k, int n, int64_t B, const T* h, T abstol, T reltol,
                                                        const int* mask, T* es_part, int npart, int off) {
    __shared__ T sred[33];
    const int b = blockIdx.y, i = blockIdx.x * W_ET + threadIdx.x;
    if (mask && !mask[b]) return;
    T es = T(0);
    if (i < n) {
        const int64_t e = (int64_t)b * n + i;
        T ut = T(0);
#pragma unroll
        for (int j = 0; j < 7; ++j) ut += Tab<T>::bt(j) * k[(int64_t)j * B * n + e];
        const T sc = abstol + kmax(kabs(uprev[e]), kabs(unew[e])) * reltol;
        const T r = (h[b] * ut) / sc;
        es = r * r;
    }
    es = wblock_sum<T>(es, sred);
    if (threadIdx.x == 0) es_part[(int64_t)b * npart + off + blockIdx.x] = es;
}

template <class T> struct WideFwd {
    const T* u0; T* uprev; T* unew; T* k;           // k: [7][B][n]
    T* h;                                            // [B] current step
    double t0, t1; const double* saveat; int nsave; T abstol, reltol; int maxiters;
    T* out; const T* target; T* dg; double* loss_sum;
    double* rec_t; T* rec_dt; T* rec; int cap;      // dense record (null rec: forward-only solve)
    T* part0; T* part1; int npart;                   // [B][npart]
    kanode_stats* stats; int* nsteps; int* retcode;
};

// block per IC: sum the partials in a fixed order (result in thread 0)
template <class T> __device__ __forceinline__ T sum_partials(const T* part, int cnt, T* sred) {
    T s = T(0);
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) s += part[k];
    return wblock_sum<T>(s, sred);
}

template <class T> __device__ void wf_begin(const WideCtl& c, const WideFwd<T>& a, int b) {
    if (!c.active[b]) return;
    const double t = c.t[b], t0 = a.t0, t1 = a.t1;
    if (!(t < t1)) { c.active[b] = 0; return; }
    const double dtmax = fabs(t1 - t0), dtmin0 = fmax(eps_of(t0), eps_of(t1));
    double dt = c.dt[b];
    int iter = c.iter[b];
    const bool accept = c.accept[b] != 0;
    if (iter > 0) { if (!accept) dt = dt / fmin(1.0 / Ctrl::qmin, c.q11[b] / Ctrl::gamma); else dt = c.dtpropose[b]; }
    ++iter;
    const double dtmin_t = fmax(eps_of(t), dtmin0);
    dt = fmin(fmax(fmin(fabs(dt), dtmax), dtmin_t), t1 - t);
    int ret = RET_SUCCESS;
    if (iter > a.maxiters) ret = RET_MAXITERS;
    else if (!(dt > dtmin_t) && (t + dt < t1 || !accept) && iter > 1) ret = RET_DTMIN;
    else if (dt != dt) ret = RET_UNSTABLE;
    c.iter[b] = iter; c.dt[b] = dt;
    if (ret != RET_SUCCESS) { c.ret[b] = ret; c.active[b] = 0; c.fail_now[b] = 1; return; }
    a.h[b] = (T)dt;
}

// forward control: finish the attempt just evaluated (error norm -> accept / reject), then open the next one
template <class T>
__global__ void __launch_bounds__(128) wide_fwd_ctl_kernel(const WideCtl c, const WideFwd<T> a, int n, int phase) {
    __shared__ T sred[33];
    const int b = blockIdx.x;
    if (phase == 0) {            // after k1 = f(u0): d0, d1 -> dt0
        const T v0 = sum_partials<T>(a.part0 + (int64_t)b * a.npart, a.npart, sred);
        const T v1 = sum_partials<T>(a.part1 + (int64_t)b * a.npart, a.npart, sred);
        if (threadIdx.x) return;
        const double d0 = sqrt((double)v0 / n), d1 = sqrt((double)v1 / n), dtmax = fabs(a.t1 - a.t0);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
        dt0 = fmin(dt0, dtmax);
        c.d1[b] = d1; c.dt0[b] = dt0;
        a.h[b] = (T)dt0;
        return;
    }
    if (phase == 1) {            // after k2 = f(u0 + dt0 k1): d2 -> dt; open the first attempt
        const T s2 = sum_partials<T>(a.part1 + (int64_t)b * a.npart, a.npart, sred);
        if (threadIdx.x) return;
        const double dt0 = c.dt0[b], d1 = c.d1[b], dtmax = fabs(a.t1 - a.t0), dtmin0 = fmax(eps_of(a.t0), eps_of(a.t1));
        const double d2 = sqrt((double)s2 / n) / dt0, mx = fmax(d1, d2);
        const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        c.dt[b] = fmax(dtmin0, fmin(fmin(100.0 * dt0, dt1), dtmax));
        c.nf[b] = 3;
        wf_begin<T>(c, a, b);
        return;
    }
    if (!c.active[b]) { if (threadIdx.x == 0) { c.acc_now[b] = 0; c.fail_now[b] = 0; } return; }
    const T es = sum_partials<T>(a.part0 + (int64_t)b * a.npart, a.npart, sred);
    if (threadIdx.x) return;
    c.acc_now[b] = 0; c.fail_now[b] = 0;
    c.nf[b] += 6;
    const double EEst = (double)ksqrt(es / T(n));
    if (EEst != EEst) { c.ret[b] = RET_UNSTABLE; c.active[b] = 0; c.fail_now[b] = 1; return; }
    double q11 = c.q11[b];
    const double q = pi_q(EEst, c.qold[b], q11);
    c.q11[b] = q11;
    const bool accept = EEst <= 1.0;
    c.accept[b] = accept;
    if (accept) {
        const double t = c.t[b], dt = c.dt[b], t1 = a.t1, dtmax = fabs(a.t1 - a.t0), dtmin0 = fmax(eps_of(a.t0), eps_of(a.t1));
        ++c.naccept[b];
        c.qold[b] = fmax(EEst, Ctrl::qoldinit);
        const double dtnew = dt / q;
        double tnew = t + dt;
        if (fabs(tnew - t1) < 100.0 * eps_of(fmax(fabs(t), fabs(t1)))) tnew = t1;
        c.dtpropose[b] = fmax(fmin(dtmax, fabs(dtnew)), fmax(eps_of(tnew), dtmin0));
        if (a.rec) {
            const int nrec = c.nrec[b];
            if (nrec >= a.cap) { c.ret[b] = RET_OVERFLOW; c.active[b] = 0; c.fail_now[b] = 1; return; }
            a.rec_t[(int64_t)b * a.cap + nrec] = t;
            a.rec_dt[(int64_t)b * a.cap + nrec] = a.h[b];
            c.nrec[b] = nrec + 1;
        }
        int sidx = c.sidx[b];
        c.s_lo[b] = sidx;
        while (sidx < a.nsave && a.saveat[sidx] <= tnew) ++sidx;
        c.s_hi[b] = sidx; c.sidx[b] = sidx;
        c.told[b] = t; c.dtold[b] = dt; c.t[b] = tnew;
        c.acc_now[b] = 1;
    } else {
        ++c.nreject[b];
    }
    wf_begin<T>(c, a, b);
}

// initial state of every IC
template <class T>
__global__ void __launch_bounds__(W_ET) wide_fwd_init_kernel(const WideCtl c, const WideFwd<T> a, int n) {
    const int b = blockIdx.y, i = blockIdx.x * W_ET + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c.t[b] = a.t0; c.dt[b] = 0.0; c.dtpropose[b] = 0.0; c.qold[b] = Ctrl::qoldinit; c.q11[b] = 1.0;
        c.iter[b] = 0; c.accept[b] = 0; c.acc_now[b] = 0; c.fail_now[b] = 0; c.active[b] = a.t0 < a.t1 ? 1 : 0;
        c.sidx[b] = a.t0 < a.t1 ? 0 : a.nsave; c.s_lo[b] = 0; c.s_hi[b] = 0; c.nrec[b] = 0; c.naccept[b] = 0; c.nreject[b] = 0; c.nf[b] = 1; c.ret[b] = RET_SUCCESS;
    }
    if (i >= n) return;
    const T v = a.u0[(int64_t)b * n + i];
    a.uprev[(int64_t)b * n + i] = v;
    if (a.t0 == a.t1 && a.out) for (int s = 0; s < a.nsave; ++s) a.out[((int64_t)b * a.nsave + s) * n + i] = v;
}

// after the control decision: dense record, outputs at the save times passed, loss, state shift
template <class T>
__global__ void __launch_bounds__(W_ET) wide_fwd_accept_kernel(const WideCtl c, const WideFwd<T> a, int n, int64_t B) {
    __shared__ double dred[33];
    const int b = blockIdx.y, i = blockIdx.x * W_ET + threadIdx.x;
    const bool acc = c.acc_now[b] != 0, failed = c.fail_now[b] != 0;
    if (!acc && !failed) return;
    double lsum = 0.0;
    if (i < n) {
        const int64_t e = (int64_t)b * n + i;
        if (acc) {
            const double told = c.told[b], dtold = c.dtold[b];
            const T h = (T)dtold;
            const T up = a.uprev[e];
            T kk[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) kk[j] = a.k[(int64_t)j * B * n + e];
            if (a.rec) {
                T* r = a.rec + ((int64_t)b * a.cap + (c.nrec[b] - 1)) * 8 * (int64_t)n;
                r[i] = up;
#pragma unroll
                for (int j = 0; j < 7; ++j) r[(int64_t)(1 + j) * n + i] = kk[j];
            }
            for (int s = c.s_lo[b]; s < c.s_hi[b]; ++s) {
                const T th = (T)((a.saveat[s] - told) / dtold);
                T bw[7]; interp_weights(th, bw);
                T accv = T(0);
#pragma unroll
                for (int j = 0; j < 7; ++j) accv += bw[j] * kk[j];
                const T v = up + h * accv;
                const int64_t o = ((int64_t)b * a.nsave + s) * n + i;
                if (a.out) a.out[o] = v;
                if (a.rec && a.target) {                            // target == null: a.dg holds the caller's cotangents
                    const T d = v - a.target[o];
                    lsum += (double)d * (double)d;
                    a.dg[o] = (T(2) / (T)((double)n * a.nsave)) * d;
                }
            }
            a.uprev[e] = a.unew[e];
            a.k[e] = kk[6];
        }
        if (failed)
            for (int s = c.sidx[b]; s < a.nsave; ++s) {
                const int64_t o = ((int64_t)b * a.nsave + s) * n + i;
                if (a.out) a.out[o] = T(NAN);
                if (a.rec) a.dg[o] = T(0);
            }
    }
    if (a.rec && acc && c.s_hi[b] > c.s_lo[b]) {
        const double tot = wblock_sum<double>(lsum, dred);
        if (threadIdx.x == 0 && tot != 0.0) atomicAdd(a.loss_sum, tot);
    }
}

template <class T>
__global__ void wide_fwd_finish_kernel(const WideCtl c, const WideFwd<T> a, int64_t B) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (a.stats) a.stats[b] = kanode_stats{c.naccept[b], c.nreject[b], c.nf[b], c.ret[b]};
    if (a.nsteps) { a.nsteps[b] = c.nrec[b]; a.retcode[b] = c.ret[b]; }
}

// ---------------------------------------------------------------------------------------------------------
// backward (interpolating adjoint) control
// ---------------------------------------------------------------------------------------------------------
template <class T> struct WideBwd {
    T* lam;            // [B][n] lambda at the start of the attempt (== lprev of the generic kernel)
    T* kl;             // [7][B][n]
    T* x1; T* yb1; T* x2; T* yb2;      // stage records, 7 slots each
    T* h; T* th; T* hd;                // [B], [7][B], [7][B]
    double t0, t1; const double* saveat; int nsave; T abstol, reltol; int maxiters;
    const double* rec_t; const T* rec_dt; const T* rec; int cap; const int* nsteps; const int* retcode;
    const T* dg; T* g; long long np;
    T* part0; T* part1; int npart;
    T* du0; kanode_stats* stats;
    // hidden-source model (kanode_wsrc.cuh): dg/dt of the few pointwise-KAN parameters is kept per stage as block partials
    const T* kg_part;  // [7][B][nkg][W_HP] or null
    int nkg;
    T* gsrc;           // [B][W_HP] gradient state of the source model
};

// hidden-source model: g terms of the control decisions.  kg[s][q] = sum of the block partials of stage s; the block's
// threads add them into shared memory, thread 0 then evaluates the few parameters.
template <class T> __device__ __forceinline__ void wsrc_sum_kg(const WideBwd<T>& a, int b, int64_t B, int nslots, T (*skg)[W_HP]) {
    for (int v = threadIdx.x; v < nslots * W_HP; v += blockDim.x) {
        const int s = v / W_HP, q = v - s * W_HP;
        const T* src = a.kg_part + (((int64_t)s * B + b) * a.nkg) * W_HP + q;
        T t = T(0);
        for (int ch = 0; ch < a.nkg; ++ch) t += src[(int64_t)ch * W_HP];
        skg[s][q] = t;
    }
    __syncthreads();
}

template <class T> __device__ __forceinline__ void wb_stage_time(const WideCtl& c, const WideBwd<T>& a, int b, int64_t B, int s, double ts) {
    const double* rt = a.rec_t + (int64_t)b * a.cap;
    const int nsteps = a.nsteps[b];
    int ridx = nsteps - 1;
    while (ts < rt[ridx] && ridx > 0) --ridx;
    while (ridx + 1 < nsteps && ts >= rt[ridx + 1]) ++ridx;
    const T hd = a.rec_dt[(int64_t)b * a.cap + ridx];
    c.ridx[(int64_t)s * B + b] = ridx;
    a.hd[(int64_t)s * B + b] = hd;
    a.th[(int64_t)s * B + b] = (T)((ts - rt[ridx]) / (double)hd);
}

template <class T> __device__ void wb_begin(const WideCtl& c, const WideBwd<T>& a, int b, int64_t B) {
    c.do_s0[b] = 0;
    for (int s = 0; s < 7; ++s) c.mask7[(int64_t)s * B + b] = 0;
    if (!c.active[b]) return;
    const double t = c.t[b], t0 = a.t0, t1 = a.t1;
    if (!(t > t0)) { c.active[b] = 0; return; }
    const double dtmax = fabs(t1 - t0), dtmin0 = fmax(eps_of(t0), eps_of(t1));
    double dt = c.dt[b];
    int iter = c.iter[b];
    const bool accept = c.accept[b] != 0;
    if (iter > 0) { if (!accept) dt = dt / fmin(1.0 / Ctrl::qmin, c.q11[b] / Ctrl::gamma); else dt = c.dtpropose[b]; }
    ++iter;
    const int sp = c.sidx[b];
    const double tstop = (sp >= 0) ? fmax(a.saveat[sp], t0) : t0;
    const double dtmin_t = fmax(eps_of(t), dtmin0);
    dt = fmin(fmax(fmin(fabs(dt), dtmax), dtmin_t), t - tstop);
    int ret = RET_SUCCESS;
    if (iter > a.maxiters) ret = RET_MAXITERS;
    else if (!(dt > dtmin_t) && (t - dt > tstop || !accept) && iter > 1) ret = RET_DTMIN;
    else if (dt != dt) ret = RET_UNSTABLE;
    c.iter[b] = iter; c.dt[b] = dt;
    if (ret != RET_SUCCESS) { c.ret[b] = ret; c.active[b] = 0; return; }
    a.h[b] = (T)(-dt);
    c.do_s0[b] = c.modified[b];
    c.mask7[b] = c.modified[b];
    for (int s = 1; s < 7; ++s) c.mask7[(int64_t)s * B + b] = 1;
    c.modified[b] = 0;
    for (int s = 0; s < 7; ++s) wb_stage_time<T>(c, a, b, B, s, t - tab_c(s) * dt);
}

template <class T>
__global__ void __launch_bounds__(128) wide_bwd_ctl_kernel(const WideCtl c, const WideBwd<T> a, int n, int64_t B, int phase) {
    __shared__ T sred[33];
    __shared__ T skg[7][W_HP];
    const int b = blockIdx.x;
    const int NZ = n + (int)a.np;
    const bool src = a.kg_part != nullptr;
    if (phase == -1) {           // state at t1: jumps at the end time (PresetTimeCallback fires at init), stage-0 time
        if (threadIdx.x) return;
        const int ok = a.retcode[b] == RET_SUCCESS && a.nsteps[b] > 0;
        c.t[b] = a.t1; c.dt[b] = 0.0; c.dtpropose[b] = 0.0; c.qold[b] = Ctrl::qoldinit; c.q11[b] = 1.0;
        c.iter[b] = 0; c.accept[b] = 0; c.acc_now[b] = 0; c.fail_now[b] = 0; c.active[b] = ok; c.modified[b] = 0; c.do_s0[b] = ok;
        c.nrec[b] = 0; c.naccept[b] = 0; c.nreject[b] = 0; c.nf[b] = 0; c.ret[b] = a.retcode[b]; c.cur[b] = 0;
        int sp = a.nsave - 1;
        c.s_hi[b] = sp;
        while (sp >= 0 && a.saveat[sp] == a.t1) --sp;
        c.s_lo[b] = sp + 1; c.sidx[b] = sp;
        if (ok) { wb_stage_time<T>(c, a, b, B, 0, a.t1); a.h[b] = T(0); }
        return;
    }
    if (phase == 0) {            // d0, d1 of initdt on the augmented state -> dt0, stage-1 time
        if (!c.active[b]) return;
        const T v0 = sum_partials<T>(a.part0 + (int64_t)b * a.npart, a.npart, sred);
        T v1 = sum_partials<T>(a.part1 + (int64_t)b * a.npart, a.npart, sred);
        if (src) wsrc_sum_kg<T>(a, b, B, 1, skg);
        if (threadIdx.x) return;
        if (src) for (int q = 0; q < (int)a.np; ++q) { const T x1 = skg[0][q] / a.abstol; v1 += x1 * x1; }
        const double d0 = sqrt((double)v0 / NZ), d1 = sqrt((double)v1 / NZ), dtmax = fabs(a.t1 - a.t0);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
        dt0 = fmin(dt0, dtmax);
        c.d1[b] = d1; c.dt0[b] = dt0;
        a.h[b] = -(T)dt0;
        wb_stage_time<T>(c, a, b, B, 1, a.t1 - dt0);
        return;
    }
    if (phase == 1) {
        if (!c.active[b]) return;
        T s2 = sum_partials<T>(a.part1 + (int64_t)b * a.npart, a.npart, sred);
        if (src) wsrc_sum_kg<T>(a, b, B, 2, skg);
        if (threadIdx.x) return;
        if (src) for (int q = 0; q < (int)a.np; ++q) { const T x2 = (skg[1][q] - skg[0][q]) / a.abstol; s2 += x2 * x2; }
        const double dt0 = c.dt0[b], d1 = c.d1[b], dtmax = fabs(a.t1 - a.t0), dtmin0 = fmax(eps_of(a.t0), eps_of(a.t1));
        const double d2 = sqrt((double)s2 / NZ) / dt0, mx = fmax(d1, d2);
        const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        c.dt[b] = fmax(dtmin0, fmin(fmin(100.0 * dt0, dt1), dtmax));
        c.nf[b] = 3;
        wb_begin<T>(c, a, b, B);
        return;
    }
    if (!c.active[b]) { if (threadIdx.x == 0) { c.acc_now[b] = 0; c.do_s0[b] = 0; } return; }
    T es = sum_partials<T>(a.part0 + (int64_t)b * a.npart, a.npart, sred);
    if (src) wsrc_sum_kg<T>(a, b, B, 7, skg);
    if (threadIdx.x) return;
    T gnew[W_HP];
    if (src) {                   // the few parameters of the pointwise KAN: g_new = g_old + sum_s wb_s kg_s and its error terms
        const T hh = a.h[b];
        for (int q = 0; q < (int)a.np; ++q) {
            T vb = T(0), vt = T(0);
#pragma unroll
            for (int st = 0; st < 7; ++st) { vb += (-hh * Tab<T>::b(st)) * skg[st][q]; vt += (-hh * Tab<T>::bt(st)) * skg[st][q]; }
            const T g0 = a.gsrc[(int64_t)b * W_HP + q], g1 = g0 + vb;
            const T r = vt / (a.abstol + kmax(kabs(g0), kabs(g1)) * a.reltol);
            es += r * r;
            gnew[q] = g1;
        }
    }
    c.acc_now[b] = 0;
    c.nf[b] += c.do_s0[b] ? 7 : 6;
    const double EEst = (double)ksqrt(es / T(NZ));
    if (EEst != EEst) { c.ret[b] = RET_UNSTABLE; c.active[b] = 0; c.do_s0[b] = 0; return; }
    double q11 = c.q11[b];
    const double q = pi_q(EEst, c.qold[b], q11);
    c.q11[b] = q11;
    const bool accept = EEst <= 1.0;
    c.accept[b] = accept;
    if (accept) {
        const double t = c.t[b], dt = c.dt[b], t0 = a.t0, dtmax = fabs(a.t1 - a.t0), dtmin0 = fmax(eps_of(a.t0), eps_of(a.t1));
        int sp = c.sidx[b];
        const double tstop = (sp >= 0) ? fmax(a.saveat[sp], t0) : t0;
        ++c.naccept[b];
        c.qold[b] = fmax(EEst, Ctrl::qoldinit);
        const double dtnew = dt / q;
        double tnew = t - dt;
        if (fabs(tnew - tstop) < 100.0 * eps_of(fmax(fabs(t), fabs(tstop)))) tnew = tstop;
        c.dtpropose[b] = fmax(fmin(dtmax, fabs(dtnew)), fmax(eps_of(tnew), dtmin0));
        c.t[b] = tnew;
        c.cur[b] ^= 1;
        if (src) for (int q = 0; q < (int)a.np; ++q) a.gsrc[(int64_t)b * W_HP + q] = gnew[q];
        c.s_hi[b] = sp;
        while (sp >= 0 && a.saveat[sp] == tnew) --sp;
        c.s_lo[b] = sp + 1; c.sidx[b] = sp;
        c.modified[b] = c.s_hi[b] >= c.s_lo[b];
        c.acc_now[b] = 1;
    } else {
        ++c.nreject[b];
    }
    wb_begin<T>(c, a, b, B);
}

// lambda at t1 (jumps of the save times equal to t1); g[b][0] is zeroed by the host
template <class T>
__global__ void __launch_bounds__(W_ET) wide_bwd_init_kernel(const WideCtl c, const WideBwd<T> a, int n) {
    const int b = blockIdx.y, i = blockIdx.x * W_ET + threadIdx.x;
    if (i >= n) return;
    T l = T(0);
    if (c.active[b]) for (int sp = c.s_hi[b]; sp >= c.s_lo[b]; --sp) l += a.dg[((int64_t)b * a.nsave + sp) * n + i];
    a.lam[(int64_t)b * n + i] = l;
}

// accepted attempt: lambda <- lambda_new (+ jumps at the save time just reached); FSAL shift of slot 6 -> slot 0
template <class T>
__global__ void __launch_bounds__(W_ET) wide_bwd_accept_kernel(const WideCtl c, const WideBwd<T> a, int n, int64_t B) {
    const int b = blockIdx.y, i = blockIdx.x * W_ET + threadIdx.x;
    if (!c.acc_now[b]) return;
    const bool modified = c.modified[b] != 0;
    if (!modified && blockIdx.x == 0 && threadIdx.x < W_HP) {
        a.x2[(int64_t)b * W_HP + threadIdx.x] = a.x2[((int64_t)6 * B + b) * W_HP + threadIdx.x];
        a.yb1[(int64_t)b * W_HP + threadIdx.x] = a.yb1[((int64_t)6 * B + b) * W_HP + threadIdx.x];
    }
    if (i >= n) return;
    const int64_t e = (int64_t)b * n + i, e6 = ((int64_t)6 * B + b) * n + i;
    const T lnew = a.yb2[e6];
    T l = lnew;
    for (int sp = c.s_hi[b]; sp >= c.s_lo[b]; --sp) l += a.dg[((int64_t)b * a.nsave + sp) * n + i];
    a.lam[e] = l;
    if (!modified) { a.kl[e] = a.kl[e6]; a.x1[e] = a.x1[e6]; a.yb2[e] = lnew; }
}

template <class T>
__global__ void __launch_bounds__(W_ET) wide_bwd_finish_kernel(const WideCtl c, const WideBwd<T> a, int n) {
    const int b = blockIdx.y, i = blockIdx.x * W_ET + threadIdx.x;
    const bool ran = a.retcode[b] == RET_SUCCESS && a.nsteps[b] > 0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.stats)
        a.stats[b] = ran ? kanode_stats{c.naccept[b], c.nreject[b], c.nf[b], c.ret[b]} : kanode_stats{0, 0, 0, a.retcode[b]};
    if (i < n && a.du0) a.du0[(int64_t)b * n + i] = ran ? a.lam[(int64_t)b * n + i] : T(0);
}

// out[j] = sum over the ICs whose adjoint succeeded of g[b][cur[b]][j]
template <class T>
__global__ void __launch_bounds__(256) wide_grad_reduce_kernel(const T* g, const int* cur, const int* ret, long long np, int64_t B, T* out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= np) return;
    double acc = 0.0;
    for (int64_t b = 0; b < B; ++b)
        if (ret[b] == RET_SUCCESS) acc += (double)g[(b * 2 + cur[b]) * np + j];
    out[j] = (T)acc;
}

// =========================================================================================================
// host side
// =========================================================================================================
inline WideModel wide_model(const kanode_handle* h) {
    const kanode_desc& d = h->desc;
    WideModel m{};
    const kanode_layer_desc &a = d.layers[0], &b = d.layers[1];
    m.n = d.n_state; m.norm1 = a.normalizer; m.norm2 = b.normalizer;
    m.inv_h1 = 1.0f / a.denominator; m.inv_h2 = 1.0f / b.denominator;
    for (int g = 0; g < a.grid_len; ++g) { m.grid1[g] = grid_point(a, g); m.grid2[g] = grid_point(b, g); }
    const long long n = m.n, H = a.out_dims, G = a.grid_len;
    m.offC1 = 0; m.offW1 = n * G * H; m.offC2 = m.offW1 + n * H; m.offW2 = m.offC2 + H * G * n;
    m.np = (long long)h->np;
    return m;
}

struct WideLaunch { int P, nchunk, bt_red, nbt_red, bt_par, nbt_par, uc, ec; };
inline WideLaunch wide_launch(int n, int64_t B, int GB) {
    WideLaunch L{};
    static const int maxch = [] { const char* e = std::getenv("KANODE_WIDE_MAXCH"); return e ? std::atoi(e) : W_MAXCH; }();   // tuning experiments
    L.uc = (n + W_BT - 1) / W_BT;
    L.ec = (n + W_ET - 1) / W_ET;
    L.P = (L.uc + maxch - 1) / maxch;
    L.nchunk = (L.uc + L.P - 1) / L.P;
    int bt = (int)((B * L.nchunk + 295) / 296);
    L.bt_red = bt < 1 ? 1 : (bt > GB ? GB : bt);
    L.nbt_red = (int)((B + L.bt_red - 1) / L.bt_red);
    bt = (int)((B * L.uc + 591) / 592);
    L.bt_par = bt < 1 ? 1 : (bt > W_PT ? W_PT : bt);
    L.nbt_par = (int)((B + L.bt_par - 1) / L.bt_par);
    return L;
}

// carve a byte arena
struct Arena {
    char* p; size_t off = 0;
    template <class U> U* take(size_t count) { off = (off + 255) / 256 * 256; U* r = reinterpret_cast<U*>(p + off); off += sizeof(U) * count; return r; }
};

inline size_t wide_ctl_bytes(int64_t B) { return (size_t)(W_NCTL_D * 8 + W_NCTL_I * 4) * B + 64 * 256; }
inline WideCtl wide_ctl_carve(Arena& A, int64_t B) {
    WideCtl c{};
    double** dd[] = {&c.t, &c.dt, &c.dtpropose, &c.qold, &c.q11, &c.told, &c.dtold, &c.d0, &c.d1, &c.dt0};
    for (auto pp : dd) *pp = A.take<double>(B);
    int** ii[] = {&c.iter, &c.accept, &c.acc_now, &c.fail_now, &c.active, &c.modified, &c.do_s0, &c.sidx, &c.s_lo, &c.s_hi, &c.nrec,
                  &c.naccept, &c.nreject, &c.nf, &c.ret, &c.cur};
    for (auto pp : ii) *pp = A.take<int>(B);
    c.ridx = A.take<int>(7 * B);
    c.mask7 = A.take<int>(7 * B);
    return c;
}

// Graph of one step attempt: `body` enqueues the attempt's kernels on h->stream.  The instantiated graph is kept per
// (kind, dtype) and reused while the signature (every pointer / size / scalar the kernels receive) is unchanged.
template <class Body>
inline int wide_attempt_graph(kanode_handle* h, int slot, const std::vector<char>& sig, Body&& body, cudaGraphExec_t* out) {
    kanode_handle::WideGraph& g = h->wide_graphs[slot];
    if (!g.exec || g.sig != sig) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        const int64_t l0 = h->launches;
        CK(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const int64_t n0 = body();
        cudaGraph_t graph = nullptr;
        // always end the capture: an error inside the body must not leave the caller's stream in capture mode
        const cudaError_t ec = cudaStreamEndCapture(h->stream, &graph);
        const int64_t extra = h->launches - l0;          // launches the helpers counted on the handle directly
        h->launches = l0;
        if (n0 < 0 || ec != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            (void)cudaGetLastError();
            return fail(h, KANODE_ERR_CUDA, "capturing a step attempt failed: %s", ec != cudaSuccess ? cudaGetErrorString(ec) : "a kernel of the attempt could not be enqueued");
        }
        const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) { g.exec = nullptr; return fail(h, KANODE_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ei)); }
        g.sig = sig; g.nodes = (int)(n0 + extra);
    }
    *out = g.exec;
    return 0;
}
template <class U> inline void sig_add(std::vector<char>& sig, const U& v) { const char* q = reinterpret_cast<const char*>(&v); sig.insert(sig.end(), q, q + sizeof(U)); }

// wait until no IC is active; polls the device flags (the host only decides how many more attempts to enqueue)
inline int wide_any_active(kanode_handle* h, const int* d_active, int64_t B, bool& any) {
    std::vector<int> act((size_t)B);
    CK(h, cudaMemcpyAsync(act.data(), d_active, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    any = false;
    for (int64_t b = 0; b < B; ++b) any |= act[(size_t)b] != 0;
    return 0;
}

// layer-1 weights in the transposed device layout, refreshed when the parameters changed
template <class T, int H, int G> int wide_w1t(kanode_handle* h, const WideModel& m, const T* p, const T** out) {
    T* d = nullptr;
    const size_t cnt = (size_t)m.n * H * (G + 1);
    const int slot = sizeof(T) == 4 ? 0 : 1;
    if (slot == 0) ENSURE(h, W_W1T32, sizeof(T) * cnt, d); else ENSURE(h, W_W1T64, sizeof(T) * cnt, d);
    if (h->wide_w1t_version[slot] != h->params_version) {
        wide_transpose_w1_kernel<T, H, G><<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(m, p, d);
        ++h->launches;
        CK(h, cudaGetLastError());
        h->wide_w1t_version[slot] = h->params_version;
    }
    *out = d;
    return 0;
}

// tensor-core image of the layer-2 weights (fp32 only), refreshed when the parameters changed
template <int H, int G> int wide_w2_image(kanode_handle* h, const WideModel& m, const float* p, const float** out) {
    constexpr int KC = (H * (G + 1) + 7) / 8 * 2;
    float* d = nullptr;
    const size_t cnt = (size_t)((m.n + TC_M - 1) / TC_M) * 2 * KC * TC_M * 4;
    ENSURE(h, W_W2IMG, sizeof(float) * cnt, d);
    if (h->wide_w2img_version != h->params_version) {
        wide_w2_image_kernel<H, G><<<(unsigned)((cnt / 2 + 255) / 256), 256, 0, h->stream>>>(m, p, d);
        ++h->launches;
        CK(h, cudaGetLastError());
        h->wide_w2img_version = h->params_version;
    }
    *out = d;
    return 0;
}

template <int H, int G> int wide_w2t_image(kanode_handle* h, const WideModel& m, const float* p, const float** out) {
    float* d = nullptr;
    const size_t cnt = (size_t)(m.n / TC_KB) * 2 * (TC_KB / 4) * TC_M * 4;
    ENSURE(h, W_W2TIMG, sizeof(float) * cnt, d);
    if (h->wide_w2timg_version != h->params_version) {
        wide_w2t_image_kernel<H, G><<<(unsigned)((cnt / 2 + 255) / 256), 256, 0, h->stream>>>(m, p, d);
        ++h->launches;
        CK(h, cudaGetLastError());
        h->wide_w2timg_version = h->params_version;
    }
    *out = d;
    return 0;
}

template <int H, int G> int wide_w1_image(kanode_handle* h, const WideModel& m, const float* p, const float** out) {
    constexpr int KU = (G + 1 + 3) / 4 * 4, KC = TC1_UC * KU / 4;
    float* d = nullptr;
    const size_t cnt = (size_t)(m.n / TC1_UC) * 2 * KC * TC1_N * 4;
    ENSURE(h, W_W1IMG, sizeof(float) * cnt, d);
    if (h->wide_w1img_version != h->params_version) {
        wide_w1_image_kernel<H, G><<<(unsigned)((cnt / 2 + 255) / 256), 256, 0, h->stream>>>(m, p, d);
        ++h->launches;
        CK(h, cudaGetLastError());
        h->wide_w1img_version = h->params_version;
    }
    *out = d;
    return 0;
}

// partial-sum rows the reduce kernels may need: CUDA-core kernels <= W_MAXCH, the tensor-core reverse kernel n / 64
inline size_t wide_part_rows(int n, const WideLaunch& L) { return (size_t)std::max(std::max(L.nchunk, n / TC_KB + 1), 600); }

// hbar = layer-2 reverse of lambda_s: tcgen05 kernel for fp32 when n is a multiple of 128, CUDA cores otherwise
template <class T, int H, int G>
int wide_l2_reverse(kanode_handle* h, const WideModel& m, const T* p, const T* hidden, const WideIn<T>& in, int64_t B, const WideLaunch& L,
                    T* part, T* hbar, unsigned* counters) {
    if constexpr (sizeof(T) == 4) {
        if (h->wide_tc && m.n % TC_M == 0) {
            constexpr int NWP = (H * (G + 1) + 3) / 4 * 4, KC = TC_KB / 4;
            constexpr size_t smem = 2 * (size_t)KC * TC_M * 16 + 2 * (size_t)KC * TC_N * 16 + (size_t)TC_N * NWP * 4 + 64;
            const float* img = nullptr;
            if (int rc = wide_w2t_image<H, G>(h, m, p, &img)) return rc;
            constexpr unsigned abit = 4u << (G == 5 ? 0 : 1);           // once per handle (= per device) and instantiation
            if (!(h->attr_done & abit)) { CK(h, cudaFuncSetAttribute(wide_l2_vjp_tc_kernel<H, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); h->attr_done |= abit; }
            const int nblk = m.n / TC_KB;
            wide_l2_vjp_tc_kernel<H, G><<<dim3(nblk, (unsigned)((B + TC_N - 1) / TC_N)), 128, smem, h->stream>>>(m, img, hidden, in, B, part);
            wide_sum_partials_kernel<float, H><<<(unsigned)((B * H + 3) / 4), 128, 0, h->stream>>>(part, nblk, B, hbar, in.mask);
            ++h->launches;
            return 0;
        }
    }
    wide_l2_vjp_kernel<T, H, G><<<dim3(L.nchunk, L.nbt_red), W_BT, 0, h->stream>>>(m, p, hidden, in, B, L.P, L.bt_red, part, hbar, counters);
    return 0;
}

// k = layer2(hidden): tensor-core kernel for fp32 (KANODE_WIDE_TC=0 selects the CUDA-core kernel), CUDA cores for fp64
template <class T, int H, int G>
int wide_l2_forward(kanode_handle* h, const WideModel& m, const T* p, const T* hidden, T* out, const int* mask, int64_t B, const WideLaunch& L) {
    if constexpr (sizeof(T) == 4) {
        if (h->wide_tc) {
            constexpr int KC = (H * (G + 1) + 7) / 8 * 2;
            constexpr size_t smem = 2 * (size_t)KC * TC_M * 16 + 2 * (size_t)KC * TC_N * 16 + 64;
            const float* img = nullptr;
            if (int rc = wide_w2_image<H, G>(h, m, p, &img)) return rc;
            constexpr unsigned abit = 16u << (G == 5 ? 0 : 1);
            if (!(h->attr_done & abit)) { CK(h, cudaFuncSetAttribute(wide_l2_fwd_tc_kernel<H, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); h->attr_done |= abit; }
            const dim3 g((m.n + TC_M - 1) / TC_M, (unsigned)((B + TC_N - 1) / TC_N));
            wide_l2_fwd_tc_kernel<H, G><<<g, 128, smem, h->stream>>>(m, img, hidden, out, mask, B);
            return 0;
        }
    }
    wide_l2_fwd_kernel<T, H, G><<<dim3(L.uc, L.nbt_par), W_BT, 0, h->stream>>>(m, p, hidden, out, mask, B, L.bt_par);
    return 0;
}

// du = chain(u) for a batch (kanode_rhs): the two forward contraction kernels of the engine
template <class T, int H, int G>
int wide_rhs_t(kanode_handle* h, const T* p, const T* d_u, T* d_du, int64_t B) {
    constexpr int GB = sizeof(T) == 4 ? 8 : 4;
    const WideModel m = wide_model(h);
    const WideLaunch L = wide_launch(m.n, B, GB);
    const T* w1t = nullptr;
    if (int rc = wide_w1t<T, H, G>(h, m, p, &w1t)) return rc;
    char* base = nullptr;
    ENSURE(h, W_WIDE_R, sizeof(T) * ((size_t)B * W_HP + (size_t)L.nchunk * B * H) + sizeof(unsigned) * (size_t)B + 4 * 256, base);
    Arena A{base};
    T* hidden = A.take<T>((size_t)B * W_HP);
    T* part = A.take<T>((size_t)L.nchunk * B * H);
    unsigned* counters = A.take<unsigned>(B);
    CK(h, cudaMemsetAsync(counters, 0, sizeof(unsigned) * (size_t)B, h->stream));
    WideIn<T> in{};
    in.base = d_u;
    wide_l1_fwd_kernel<T, H, G, 0><<<dim3(L.nchunk, L.nbt_red), W_BT, 0, h->stream>>>(m, w1t, in, B, L.P, L.bt_red, part, hidden, counters);
    if (int rc = wide_l2_forward<T, H, G>(h, m, p, hidden, d_du, nullptr, B, L)) return rc;
    h->launches += 2;
    CK(h, cudaGetLastError());
    return 0;
}

template <class T, int H, int G>
int wide_forward(kanode_handle* h, const WideModel& m, const T* p, WideFwd<T> a, int64_t B, bool dense, WideCtl* ctl_out) {
    constexpr int GB = sizeof(T) == 4 ? 8 : 4;
    const int n = m.n;
    const WideLaunch L = wide_launch(n, B, GB);
    const T* w1t = nullptr;
    if (int rc = wide_w1t<T, H, G>(h, m, p, &w1t)) return rc;
    a.npart = L.ec;
    const size_t nB = (size_t)n * B;
    size_t bytes = wide_ctl_bytes(B) + sizeof(T) * (9 * nB + (size_t)B * (1 + W_HP + 2 * L.ec) + (size_t)L.nchunk * B * H) + sizeof(unsigned) * (size_t)B + 16 * 256;
    char* base = nullptr;
    ENSURE(h, W_WIDE_F, bytes, base);
    Arena A{base};
    WideCtl c = wide_ctl_carve(A, B);
    a.uprev = A.take<T>(nB); a.unew = A.take<T>(nB); a.k = A.take<T>(7 * nB); a.h = A.take<T>(B);
    T* hidden = A.take<T>((size_t)B * W_HP);
    a.part0 = A.take<T>((size_t)B * L.ec); a.part1 = A.take<T>((size_t)B * L.ec);
    T* part = A.take<T>((size_t)L.nchunk * B * H);
    unsigned* counters = A.take<unsigned>(B);
    if (!h->wide_counters_zeroed[0] || h->wide_counters_ptr[0] != counters) {
        CK(h, cudaMemsetAsync(counters, 0, sizeof(unsigned) * (size_t)B, h->stream));
        h->wide_counters_zeroed[0] = true; h->wide_counters_ptr[0] = counters;
    }
    if (!dense) { a.rec = nullptr; a.rec_t = nullptr; a.rec_dt = nullptr; }
    cudaStream_t st = h->stream;
    const dim3 ge(L.ec, (unsigned)B), gr(L.nchunk, L.nbt_red), gp(L.uc, L.nbt_par);
    int64_t launches = 0;
    auto rhs = [&](int ncoef, const T* coef, int kslot, T* xstore) {     // k[kslot] = f(uprev + h * sum coef_j k_j)
        WideIn<T> in{};
        in.base = a.uprev; in.ks = a.k; in.hs = a.h; in.ncoef = ncoef;
        for (int j = 0; j < ncoef; ++j) in.coef[j] = coef[j];
        in.xstore = xstore; in.mask = c.active;
        wide_l1_fwd_kernel<T, H, G, 0><<<gr, W_BT, 0, st>>>(m, w1t, in, B, L.P, L.bt_red, part, hidden, counters);
        wide_l2_forward<T, H, G>(h, m, p, hidden, a.k + (size_t)kslot * nB, c.active, B, L);
        launches += 2;
    };
    wide_fwd_init_kernel<T><<<ge, W_ET, 0, st>>>(c, a, n);
    rhs(0, nullptr, 0, nullptr);
    wide_norm_kernel<T, 1><<<ge, W_ET, 0, st>>>(a.uprev, a.k, nullptr, n, a.abstol, a.reltol, c.active, a.part0, a.part1, L.ec, 0);
    wide_fwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, a, n, 0);
    const T one[1] = {T(1)};
    rhs(1, one, 1, nullptr);
    wide_norm_kernel<T, 2><<<ge, W_ET, 0, st>>>(a.uprev, a.k, a.k + nB, n, a.abstol, a.reltol, c.active, a.part0, a.part1, L.ec, 0);
    wide_fwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, a, n, 1);
    launches += 5;
    static const double A_[7][8] = KANODE_TSIT5_A;
    int iters = 0;
    const int slot = dense ? 1 : 0;
    const int expect = h->wide_iters[slot];          // attempts the previous call of this kind needed
    bool any = a.t0 < a.t1;
    auto attempt = [&]() -> int64_t {
        const int64_t l0 = launches;
        for (int s = 1; s < 7; ++s) {
            T coef[7];
            for (int j = 0; j < s; ++j) coef[j] = (T)A_[s][j];
            rhs(s, coef, s, s == 6 ? a.unew : nullptr);
        }
        wide_err_kernel<T><<<ge, W_ET, 0, st>>>(a.uprev, a.unew, a.k, n, B, a.h, a.abstol, a.reltol, c.active, a.part0, L.ec, 0);
        wide_fwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, a, n, 2);
        wide_fwd_accept_kernel<T><<<ge, W_ET, 0, st>>>(c, a, n, B);
        launches += 3;
        return launches - l0;
    };
    cudaGraphExec_t gexec = nullptr;
    if (h->wide_graph && n <= h->wide_graph_maxn && any) {
        std::vector<char> sig;
        sig_add(sig, m); sig_add(sig, a); sig_add(sig, c); sig_add(sig, L); sig_add(sig, B); sig_add(sig, p); sig_add(sig, w1t);
        sig_add(sig, hidden); sig_add(sig, part); sig_add(sig, counters); sig_add(sig, h->wide_tc); sig_add(sig, h->ws[kanode_handle::W_W2IMG].p);
        const int64_t l0 = launches;
        if (int rc = wide_attempt_graph(h, (dense ? 1 : 0) * 2 + (sizeof(T) == 8), sig, attempt, &gexec)) return rc;
        launches = l0;
    }
    while (any) {
        if (gexec) { CK(h, cudaGraphLaunch(gexec, st)); launches += h->wide_graphs[(dense ? 1 : 0) * 2 + (sizeof(T) == 8)].nodes; }
        else attempt();
        ++iters;
        if (iters >= expect) {
            if (int rc = wide_any_active(h, c.active, B, any)) return rc;
            if (!any && iters == expect && iters > 1) --iters;      // possibly overshot: poll one attempt earlier next time
        }
        if (iters > a.maxiters + 2) break;
    }
    h->wide_iters[slot] = iters;
    wide_fwd_finish_kernel<T><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(c, a, B);
    h->launches += launches + 1;
    CK(h, cudaGetLastError());
    if (ctl_out) *ctl_out = c;
    return 0;
}

template <class T, int H, int G>
int wide_solve_t(kanode_handle* h, const T* p, const T* d_u0, int64_t B, double t0, double t1, const double* d_saveat, int nsave,
                 double abstol, double reltol, T* d_out, kanode_stats* d_stats) {
    const WideModel m = wide_model(h);
    WideFwd<T> a{};
    a.u0 = d_u0; a.t0 = t0; a.t1 = t1; a.saveat = d_saveat; a.nsave = nsave; a.abstol = (T)abstol; a.reltol = (T)reltol;
    a.maxiters = 100000; a.out = d_out; a.stats = d_stats;
    return wide_forward<T, H, G>(h, m, p, a, B, false, nullptr);
}

template <class T, int H, int G>
int wide_loss_grad_t(kanode_handle* h, const T* p, const T* d_u0, int64_t B, double t0, double t1, const double* d_saveat, int nsave,
                     const T* d_target, double abstol, double reltol, double* d_loss_sum, T* d_grad_sum, T* d_du0,
                     kanode_stats* d_fst, kanode_stats* d_bst, T* d_out_opt) {
    constexpr int GB = sizeof(T) == 4 ? 8 : 4;
    const WideModel m = wide_model(h);
    const int n = m.n;
    const size_t nB = (size_t)n * B, np = h->np;
    const int cap = h->rec_cap;
    cudaStream_t st = h->stream;
    // ---- forward, dense ----
    WideFwd<T> a{};
    a.u0 = d_u0; a.t0 = t0; a.t1 = t1; a.saveat = d_saveat; a.nsave = nsave; a.abstol = (T)abstol; a.reltol = (T)reltol;
    a.maxiters = 100000; a.out = d_out_opt; a.stats = d_fst; a.target = d_target; a.loss_sum = d_loss_sum; a.cap = cap;
    ENSURE(h, W_REC_T, sizeof(double) * (size_t)cap * B, a.rec_t);
    ENSURE(h, W_GEN2, sizeof(T) * (size_t)cap * B, a.rec_dt);
    ENSURE(h, W_REC, sizeof(T) * (size_t)cap * 8 * nB, a.rec);
    ENSURE(h, W_NSTEPS, sizeof(int) * (size_t)B, a.nsteps);
    ENSURE(h, W_RET, sizeof(int) * (size_t)B, a.retcode);
    ENSURE(h, W_DG, sizeof(T) * (size_t)nsave * nB, a.dg);
    T* g = nullptr;
    ENSURE(h, W_G, sizeof(T) * 2 * np * B, g);
    cudaEventRecord(h->ev[0], st);
    if (int rc = wide_forward<T, H, G>(h, m, p, a, B, true, nullptr)) return rc;
    cudaEventRecord(h->ev[1], st);
    // ---- backward ----
    const WideLaunch L = wide_launch(n, B, GB);
    const T* w1t = nullptr;
    if (int rc = wide_w1t<T, H, G>(h, m, p, &w1t)) return rc;
    constexpr int ROWS1 = W_GK * 2 * W_GT / H;
    const int nblkC = (int)(((int64_t)n * G + ROWS1 - 1) / ROWS1), gx1 = nblkC + (n + ROWS1 - 1) / ROWS1;
    const int np_l = L.ec, np_1 = gx1, np_2 = L.uc, npart = np_l + np_1 + np_2;
    size_t bytes = wide_ctl_bytes(B) + sizeof(T) * (nB + 7 * nB * 3 + (size_t)7 * B * W_HP * 2 + (size_t)B * 15 + 2 * (size_t)B * npart +
                                                    wide_part_rows(n, L) * 7 * B * H) + sizeof(unsigned) * (size_t)7 * B + 24 * 256;
    char* base = nullptr;
    ENSURE(h, W_WIDE_B, bytes, base);
    Arena A{base};
    WideCtl c = wide_ctl_carve(A, B);
    WideBwd<T> w{};
    w.lam = A.take<T>(nB); w.kl = A.take<T>(7 * nB); w.x1 = A.take<T>(7 * nB); w.yb2 = A.take<T>(7 * nB);
    w.yb1 = A.take<T>((size_t)7 * B * W_HP); w.x2 = A.take<T>((size_t)7 * B * W_HP);
    w.h = A.take<T>(B); w.th = A.take<T>(7 * (size_t)B); w.hd = A.take<T>(7 * (size_t)B);
    w.part0 = A.take<T>((size_t)B * npart); w.part1 = A.take<T>((size_t)B * npart);
    T* part = A.take<T>(wide_part_rows(n, L) * 7 * B * H);
    unsigned* counters = A.take<unsigned>(7 * B);
    if (!h->wide_counters_zeroed[1] || h->wide_counters_ptr[1] != counters) {
        CK(h, cudaMemsetAsync(counters, 0, sizeof(unsigned) * (size_t)7 * B, st));
        h->wide_counters_zeroed[1] = true; h->wide_counters_ptr[1] = counters;
    }
    w.t0 = t0; w.t1 = t1; w.saveat = d_saveat; w.nsave = nsave; w.abstol = (T)abstol; w.reltol = (T)reltol; w.maxiters = 100000;
    w.rec_t = a.rec_t; w.rec_dt = a.rec_dt; w.rec = a.rec; w.cap = cap; w.nsteps = a.nsteps; w.retcode = a.retcode;
    w.dg = a.dg; w.g = g; w.np = (long long)np; w.npart = npart; w.du0 = d_du0; w.stats = d_bst;
    CK(h, cudaMemset2DAsync(g, sizeof(T) * 2 * np, 0, sizeof(T) * np, (size_t)B, st));        // g[b][0][:] = 0
    CK(h, cudaMemsetAsync(w.part0, 0, sizeof(T) * (size_t)B * npart, st));
    CK(h, cudaMemsetAsync(w.part1, 0, sizeof(T) * (size_t)B * npart, st));
    const dim3 ge(L.ec, (unsigned)B), gr(L.nchunk, L.nbt_red), gp(L.uc, L.nbt_par), gg1(gx1, (unsigned)B), gg2(L.uc, (unsigned)B);
    int64_t launches = 0;
    static const double A_[7][8] = KANODE_TSIT5_A;
    // adjoint rhs of stage slot s: records x_l / ybar_l, kl[s] = -(df/du)^T lambda_s
    const WideLaunch L7 = wide_launch(n, 7 * B, GB);       // all 7 stages of an attempt as (stage, IC) pairs
    // layer-1 forward of EVERY stage of the attempt in one launch: sol(t_s) does not depend on lambda, and the stage times are
    // known when the attempt opens, so the 7 x B (stage, IC) pairs share each unit's weights in one pass
    const float* w1img = nullptr;
    bool l1_tc = false;
    // opt-in (KANODE_WIDE_TC=2): measured on B200 it only ties the CUDA-core kernel (Schrodinger x 32: 181 vs ~175 us per attempt;
    // slower on Allen-Cahn-4096): with N = 16 the 36 small MMAs of a pass keep the tensor pipe 5 % busy and the block waits on
    // their completion before it may overwrite the feature operand (profiles/r01w_tc1_ncu_summary.md)
    if constexpr (sizeof(T) == 4) {
        if (h->wide_tc >= 2 && n % 64 == 0) {
            constexpr int KU = (G + 1 + 3) / 4 * 4, KC = TC1_UC * KU / 4;
            constexpr size_t smem1 = 2 * (2 * (size_t)KC * TC_M * 16 + 2 * (size_t)KC * TC1_N * 16) + 64;
            if (int rc = wide_w1_image<H, G>(h, m, p, &w1img)) return rc;
            constexpr unsigned abit = 64u << (G == 5 ? 0 : 1);
            if (!(h->attr_done & abit)) { CK(h, cudaFuncSetAttribute(wide_l1_fwd_tc_kernel<H, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1)); h->attr_done |= abit; }
            l1_tc = true;
        }
    }
    auto l1_all_stages = [&]() {
        WideIn<T> in{};
        in.rec = w.rec; in.cap = cap; in.ridx = c.ridx; in.th = w.th; in.hd = w.hd; in.brec = (int)B;
        in.xstore = w.x1; in.mask = c.mask7;
        if constexpr (sizeof(T) == 4) {
            if (l1_tc) {
                constexpr int KU = (G + 1 + 3) / 4 * 4, KC = TC1_UC * KU / 4;
                constexpr size_t smem1 = 2 * (2 * (size_t)KC * TC_M * 16 + 2 * (size_t)KC * TC1_N * 16) + 64;
                const int nblk = n / TC1_UC, nmt = (int)((7 * B + TC_M - 1) / TC_M);
                int P1 = (int)(((int64_t)nblk * nmt + 295) / 296); P1 = P1 < 2 ? 2 : P1;        // one block per SM, two waves
                const int nch = (nblk + P1 - 1) / P1;                       // <= 296 partial rows
                wide_l1_fwd_tc_kernel<H, G><<<dim3(nch, nmt), 256, smem1, st>>>(m, w1img, in, 7 * B, P1, part);
                wide_sum_partials_kernel<float, H><<<(unsigned)((7 * B * H + 3) / 4), 128, 0, st>>>(part, nch, 7 * B, w.x2, c.mask7);
                launches += 2;
                return;
            }
        }
        wide_l1_fwd_kernel<T, H, G, 1><<<dim3(L7.nchunk, L7.nbt_red), W_BT, 0, st>>>(m, w1t, in, 7 * B, L7.P, L7.bt_red, part, w.x2, counters);
        ++launches;
    };
    // adjoint rhs of stage slot s: records x_l / ybar_l, kl[s] = -(df/du)^T lambda_s
    auto adj = [&](int s, int ncoef, const T* coef, const int* mask, bool with_l1) {
        if (with_l1) {
            WideIn<T> in{};
            in.rec = w.rec; in.cap = cap; in.ridx = c.ridx + (size_t)s * B; in.th = w.th + (size_t)s * B; in.hd = w.hd + (size_t)s * B;
            in.xstore = w.x1 + (size_t)s * nB; in.mask = mask;
            wide_l1_fwd_kernel<T, H, G, 1><<<gr, W_BT, 0, st>>>(m, w1t, in, B, L.P, L.bt_red, part, w.x2 + (size_t)s * B * W_HP, counters);
            ++launches;
        }
        WideIn<T> il{};
        il.base = w.lam; il.ks = w.kl; il.hs = w.h; il.ncoef = ncoef;
        for (int j = 0; j < ncoef; ++j) il.coef[j] = coef[j];
        il.xstore = w.yb2 + (size_t)s * nB; il.mask = mask;
        wide_l2_reverse<T, H, G>(h, m, p, w.x2 + (size_t)s * B * W_HP, il, B, L, part, w.yb1 + (size_t)s * B * W_HP, counters);
        wide_l1_vjp_kernel<T, H, G><<<gp, W_BT, 0, st>>>(m, w1t, w.x1 + (size_t)s * nB, w.yb1 + (size_t)s * B * W_HP, w.kl + (size_t)s * nB, mask, B, L.bt_par);
        launches += 2;
    };
    WideGp<T> gpa{};
    gpa.x1 = w.x1; gpa.yb1 = w.yb1; gpa.x2 = w.x2; gpa.yb2 = w.yb2; gpa.g = g; gpa.cur = c.cur; gpa.h = w.h; gpa.mask = c.active;
    gpa.abstol = w.abstol; gpa.reltol = w.reltol; gpa.npart = npart;
    h->wide_gp_used = 0;
    auto gp_event = [&]() {
        if ((size_t)h->wide_gp_used == h->wide_gp_ev.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; h->wide_gp_ev.push_back(e); }
        cudaEventRecord(h->wide_gp_ev[h->wide_gp_used++], st);
    };
    auto gpass = [&](int mode, T* dst) {
        const bool timed = mode == 0;
        WideGp<T> q1 = gpa, q2 = gpa;
        if (mode == 0) gp_event();
        if (mode == 3) mode = 0;                                        // inside a graph capture: no timing events
        q1.es_part = dst; q1.off = np_l; q2.es_part = dst; q2.off = np_l + np_1;
        if (mode == 0) { wide_gp1_kernel<T, H, G, 0><<<gg1, W_GT, 0, st>>>(m, q1, B, nblkC); wide_gp2_kernel<T, H, G, 0><<<gg2, W_BT, 0, st>>>(m, q2, B); }
        else if (mode == 1) { wide_gp1_kernel<T, H, G, 1><<<gg1, W_GT, 0, st>>>(m, q1, B, nblkC); wide_gp2_kernel<T, H, G, 1><<<gg2, W_BT, 0, st>>>(m, q2, B); }
        else { wide_gp1_kernel<T, H, G, 2><<<gg1, W_GT, 0, st>>>(m, q1, B, nblkC); wide_gp2_kernel<T, H, G, 2><<<gg2, W_BT, 0, st>>>(m, q2, B); }
        if (timed) gp_event();
        launches += 2;
    };
    wide_bwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, w, n, B, -1);
    wide_bwd_init_kernel<T><<<ge, W_ET, 0, st>>>(c, w, n);
    adj(0, 0, nullptr, c.active, true);
    wide_norm_kernel<T, 1><<<ge, W_ET, 0, st>>>(w.lam, w.kl, nullptr, n, w.abstol, w.reltol, c.active, w.part0, w.part1, npart, 0);
    gpass(1, w.part1);
    wide_bwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, w, n, B, 0);
    const T one[1] = {T(1)};
    adj(1, 1, one, c.active, true);
    wide_norm_kernel<T, 2><<<ge, W_ET, 0, st>>>(w.lam, w.kl, w.kl + nB, n, w.abstol, w.reltol, c.active, w.part0, w.part1, npart, 0);
    gpass(2, w.part1);
    wide_bwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, w, n, B, 1);
    launches += 6;
    int iters = 0;
    const int expect = h->wide_iters[2];
    bool any = t0 < t1;
    bool capturing = false;
    auto attempt = [&]() -> int64_t {
        const int64_t l0 = launches;
        l1_all_stages();
        adj(0, 0, nullptr, c.do_s0, false);                            // only the ICs whose lambda jumped at a save time
        for (int s = 1; s < 7; ++s) {
            T coef[7];
            for (int j = 0; j < s; ++j) coef[j] = (T)A_[s][j];
            adj(s, s, coef, c.active, false);
        }
        wide_err_kernel<T><<<ge, W_ET, 0, st>>>(w.lam, w.yb2 + (size_t)6 * nB, w.kl, n, B, w.h, w.abstol, w.reltol, c.active, w.part0, npart, 0);
        gpass(capturing ? 3 : 0, w.part0);
        wide_bwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, w, n, B, 2);
        wide_bwd_accept_kernel<T><<<ge, W_ET, 0, st>>>(c, w, n, B);
        launches += 3;
        return launches - l0;
    };
    cudaGraphExec_t gexec = nullptr;
    const int gslot = 4 + (sizeof(T) == 8);
    if (h->wide_graph && n <= h->wide_graph_maxn && any) {
        std::vector<char> sig;
        sig_add(sig, m); sig_add(sig, w); sig_add(sig, c); sig_add(sig, L); sig_add(sig, B); sig_add(sig, p); sig_add(sig, w1t); sig_add(sig, g);
        sig_add(sig, part); sig_add(sig, counters); sig_add(sig, npart); sig_add(sig, h->wide_tc); sig_add(sig, h->ws[kanode_handle::W_W2TIMG].p);
        sig_add(sig, w1img); sig_add(sig, l1_tc);
        const int64_t l0 = launches;
        capturing = true;
        const int rc = wide_attempt_graph(h, gslot, sig, attempt, &gexec);
        capturing = false;
        if (rc) return rc;
        launches = l0;
    }
    while (any) {
        if (gexec) { CK(h, cudaGraphLaunch(gexec, st)); launches += h->wide_graphs[gslot].nodes; }
        else attempt();
        ++iters;
        if (iters >= expect) {
            if (int rc = wide_any_active(h, c.active, B, any)) return rc;
            if (!any && iters == expect && iters > 1) --iters;
        }
        if (iters > w.maxiters + 2) break;
    }
    h->wide_iters[2] = iters;
    wide_bwd_finish_kernel<T><<<ge, W_ET, 0, st>>>(c, w, n);
    cudaEventRecord(h->ev[2], st);
    wide_grad_reduce_kernel<T><<<(unsigned)((np + 255) / 256), 256, 0, st>>>(g, c.cur, c.ret, (long long)np, B, d_grad_sum);
    cudaEventRecord(h->ev[3], st);
    h->ev_valid = true;
    h->launches += launches + 2;
    CK(h, cudaGetLastError());
    return 0;
}

}  // namespace kanode
