// kanode_lg.cu — front end of the lane-group adjoint engine for the small-model ensembles (its own translation unit:
// compiles in parallel with kanode_api.cu).  One training step = dense forward Tsit5 solve (thread per trajectory, record
// array-of-structures) + interpolating-adjoint backward solve (lane group per trajectory, kanode_small_lg.cuh) + a
// deterministic fp64 sum of the per-warp gradient partials.
//
// Reference call being replaced: Zygote.gradient(loss, p) at Lotka-Volterra/LV_driver_KANODE.jl:284 with
// loss = mean(abs2, X - predict(p)) (:197-203).
#include "kanode_small_host.h"
#include "kanode_small_lg.cuh"

namespace kanode {

// packed weights in LANE blocks for the lane-group backward kernel (layout: LgSmem::LW): block `lig` holds the UPL hidden
// units of lane `lig`, each in the SmallParams::UW layout
template <class T, class P, int UPL> int upload_packed_lg(kanode_handle* h, const T** out) {
    using SMP = LgSmem<T, P, UPL>;
    constexpr int I = P::I, G = P::G, NQ = P::NQ, LPT = LgGeom<T, P, UPL>::LPT;
    T* d = nullptr;
    const int slot = sizeof(T) == 4 ? 0 : 1;
    if (slot == 0) ENSURE(h, W_WLG32, sizeof(T) * SMP::WLG, d); else ENSURE(h, W_WLG64, sizeof(T) * SMP::WLG, d);
    if (h->wlg_version[slot] != h->params_version) {
        if (int rc = host_params_refresh(h)) return rc;
        std::vector<T> pk((size_t)SMP::WLG, T(0));
        for (int l = 0; l < LPT; ++l)
            for (int u = 0; u < UPL; ++u) {
                const int j = UPL * l + u;
                T* w = pk.data() + (size_t)l * SMP::LW + (size_t)u * P::UW;
                for (int i = 0; i < I; ++i) {                          // w1: index q*I + i (q < G: C1, q = G: W1)
                    for (int g = 0; g < G; ++g) w[g * I + i] = (T)h->params[P::OC1 + (i * G + g) * P::H + j];
                    w[G * I + i] = (T)h->params[P::OW1 + i * P::H + j];
                }
                for (int g = 0; g < G; ++g)
                    for (int o = 0; o < I; ++o) w[NQ + g * I + o] = (T)h->params[P::OC2 + (j * G + g) * I + o];
                for (int o = 0; o < I; ++o) w[NQ + G * I + o] = (T)h->params[P::OW2 + j * I + o];
            }
        CK(h, cudaMemcpyAsync(d, pk.data(), sizeof(T) * pk.size(), cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));                       // pk is a stack-lifetime staging buffer
        h->wlg_version[slot] = h->params_version;
    }
    *out = d;
    return 0;
}

// one thread per hidden unit: the same two images upload_packed / upload_packed_lg build on the host
template <class P, int UPL>
__global__ void __launch_bounds__(64) small_pack_kernel(const float* __restrict__ p, float* __restrict__ wpk, float* __restrict__ wlg) {
    using SMP = LgSmem<float, P, UPL>;
    constexpr int I = P::I, G = P::G, NQ = P::NQ;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P::H) return;
    float* a = wpk + (size_t)j * P::UW;                                          // [C1 (i*G+g) | W1 (i) | C2 (g*I+o) | W2 (o)]
    float* b = wlg + (size_t)(j / UPL) * SMP::LW + (size_t)(j % UPL) * P::UW;    // w1 index q*I + i, then the same w2 block
    for (int i = 0; i < I; ++i) {
        for (int g = 0; g < G; ++g) { const float v = p[P::OC1 + (i * G + g) * P::H + j]; a[i * G + g] = v; b[g * I + i] = v; }
        const float v = p[P::OW1 + i * P::H + j]; a[I * G + i] = v; b[G * I + i] = v;
    }
    for (int g = 0; g < G; ++g)
        for (int o = 0; o < I; ++o) { const float v = p[P::OC2 + (j * G + g) * I + o]; a[NQ + g * I + o] = v; b[NQ + g * I + o] = v; }
    for (int o = 0; o < I; ++o) { const float v = p[P::OW2 + j * I + o]; a[NQ + G * I + o] = v; b[NQ + G * I + o] = v; }
}

int small_pack_dev(kanode_handle* h, const float* d_p, bool* handled) {
    int rc = 0;
    auto run = [&]<class P, int NORM>() -> int {
        constexpr int UPL = 2;
        using SMP = LgSmem<float, P, UPL>;
        float *wpk = nullptr, *wlg = nullptr;
        const bool fresh_pk = h->ws[kanode_handle::W_WPK32].bytes < sizeof(float) * P::WPK, fresh_lg = h->ws[kanode_handle::W_WLG32].bytes < sizeof(float) * SMP::WLG;
        ENSURE(h, W_WPK32, sizeof(float) * P::WPK, wpk);
        ENSURE(h, W_WLG32, sizeof(float) * SMP::WLG, wlg);
        if (fresh_pk) CK(h, cudaMemsetAsync(wpk, 0, sizeof(float) * P::WPK, h->stream));        // pad entries stay zero
        if (fresh_lg) CK(h, cudaMemsetAsync(wlg, 0, sizeof(float) * SMP::WLG, h->stream));
        small_pack_kernel<P, UPL><<<(P::H + 63) / 64, 64, 0, h->stream>>>(d_p, wpk, wlg);
        ++h->launches;
        CK(h, cudaGetLastError());
        h->wpk_version[0] = h->params_version; h->wlg_version[0] = h->params_version;
        return 0;
    };
    *handled = small_dispatch<float>(h, run, rc);
    return rc;
}

// dL/du(t_s) and the loss from stored predictions: dg = 2 (pred - X) / (I nsave), loss += (pred - X)^2 — what the forward
// kernel does in place when the target is already on the device.  Used when a host entry point's target copy is still in
// flight while the forward solve runs (kanode_api.cu: target_late): NaN predictions (unsaved points of a failed solve) give
// zero cotangent and no loss, as in the forward kernel.  Layouts: pred / target / dg [B][nsave][I] (one flat index).
template <class T>
__global__ void __launch_bounds__(256) loss_dg_kernel(const T* __restrict__ pred, const T* __restrict__ target, int64_t count, T scale,
                                                     T* __restrict__ dg, double* __restrict__ loss_sum) {
    double lsum = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const T v = pred[i];
        T d = T(0);
        if (v == v) { const T e = v - target[i]; lsum += (double)e * (double)e; d = scale * e; }
        dg[i] = d;
    }
    for (int off = 16; off > 0; off >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
    if ((threadIdx.x & 31) == 0 && lsum != 0.0) atomicAdd(loss_sum, lsum);
}

#ifndef KANODE_LG_NSLAB
#define KANODE_LG_NSLAB 296        // two slabs per SM
#endif

template <class T>
int small_lg_loss_grad(kanode_handle* h, const T* d_u0, int64_t B, double t0, double t1, const double* d_saveat, int nsave,
                       const T* d_target, double abstol, double reltol, double* d_loss_sum, T* d_grad_sum, T* d_du0,
                       kanode_stats* d_fst, kanode_stats* d_bst, T* d_out_opt, const double* d_rp_fwd, const double* d_rp_bwd,
                       int rp_cap, bool* handled) {
    int rc = 0;
    auto launch = [&]<class P, int NORM, int UPL, int WPB, int MINB>() -> int {
        constexpr int I = P::I;
        P prm; fill_small<T>(h, prm);
        using GM = LgGeom<T, P, UPL>; using SMP = LgSmem<T, P, UPL>; using RL = RecLayout<T, I>;
        constexpr int NSLAB = KANODE_LG_NSLAB;
        const int cap = h->rec_cap;
        const int64_t nwarps = (B + GM::TPW - 1) / GM::TPW;
        const unsigned nblk = (unsigned)((nwarps + WPB - 1) / WPB);
        T *rec = nullptr, *dg = nullptr, *gpart = nullptr; double* slab = nullptr;
        int *nsteps = nullptr, *retc = nullptr;
        ENSURE(h, W_REC, sizeof(T) * (size_t)cap * RL::RS * B, rec);
        ENSURE(h, W_NSTEPS, sizeof(int) * (size_t)B, nsteps);
        ENSURE(h, W_RET, sizeof(int) * (size_t)B, retc);
        ENSURE(h, W_DG, sizeof(T) * (size_t)nsave * I * B, dg);
        ENSURE(h, W_GPART, sizeof(T) * (size_t)nblk * WPB * P::NP, gpart);
        ENSURE(h, W_SLAB, sizeof(double) * (size_t)NSLAB * P::NP, slab);
        SmallFwdArgs<T> a{};
        if (int rcw = upload_packed<T, P>(h, &a.wpk)) return rcw;
        a.u0 = d_u0; a.B = B; a.t0 = t0; a.t1 = t1; a.saveat = d_saveat; a.nsave = nsave;
        a.abstol = (T)abstol; a.reltol = (T)reltol; a.maxiters = 100000; a.out = d_out_opt; a.stats = d_fst;
        a.rec_t = nullptr; a.rec = rec; a.cap = cap; a.nsteps = nsteps; a.retcode = retc;
        a.target = d_target; a.dg = dg; a.loss_sum = d_loss_sum; a.rp_t = d_rp_fwd; a.rp_cap = rp_cap;
        LgBwdArgs<T> bw{};
        if (int rcw = upload_packed_lg<T, P, UPL>(h, &bw.wpk)) return rcw;
        bw.B = B; bw.t0 = t0; bw.t1 = t1; bw.saveat = d_saveat; bw.nsave = nsave;
        bw.abstol = (T)abstol; bw.reltol = (T)reltol; bw.maxiters = h->bwd_maxiters;
        bw.rec = rec; bw.cap = cap; bw.nsteps = nsteps; bw.retcode = retc; bw.dg = dg; bw.gpart = gpart;
        bw.du0 = d_du0; bw.stats = d_bst; bw.attempts = nullptr; bw.rp_t = d_rp_bwd; bw.rp_cap = rp_cap;
        // launch order from the previous call's per-trajectory step margins (same batch size and dtype)
        const int slot = sizeof(T) == 4 ? 0 : 1;
        bool ordered = false, order_forked = false;
        if (h->schedule && !d_rp_bwd && nwarps >= 512 && B < (1ll << 30)) {
            int *mar = nullptr, *ord = nullptr;
            constexpr size_t NCNT = (size_t)32 * LG_OBLK * LG_NBK;         // class counts of the order kernels behind the permutations
            ENSURE(h, W_ATT, sizeof(int) * (size_t)B * 2, mar);
            ENSURE(h, W_ORDER, sizeof(int) * ((size_t)B * 2 + NCNT), ord);
            int* cnt = ord + (size_t)B * 2;
            mar += (size_t)slot * B; ord += (size_t)slot * B;
            if (h->order_B[slot] == B) {
                // the order depends only on the previous call's margins: its two kernels run on the second stream, under the
                // forward solve (fork behind everything queued on the main stream; the adjoint launch joins below)
                // (host entry point with the target copy pending: the second stream belongs to that copy, which is the longer
                // pole — the order kernels and the forward solve both fit under it on the main stream)
                cudaStream_t so = h->stream;
                if (h->aux_stream && !h->target_late) {
                    if (!h->order_ev) CK(h, cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
                    CK(h, cudaEventRecord(h->aux_ev[0], h->stream));
                    CK(h, cudaStreamWaitEvent(h->aux_stream, h->aux_ev[0], 0));
                    so = h->aux_stream;
                }
                lg_order_count_kernel<<<LG_OBLK, 1024, 0, so>>>(mar, (int)B, cnt);
                lg_order_scatter_kernel<<<LG_OBLK, 1024, 0, so>>>(mar, (int)B, cnt, ord);
                if (so != h->stream) { CK(h, cudaEventRecord(h->order_ev, so)); order_forked = true; }
                h->launches += 2;
                bw.order = ord; ordered = true;
            }
            bw.tmargin = mar;
            h->order_B[slot] = B;
        }
        // persistent launch: MINB blocks per SM, warps draw their positions from a ticket counter (kanode_small_lg.cuh)
        int* ticket = nullptr;
        const bool fresh_ticket = h->ws[kanode_handle::W_TICKET].bytes == 0;
        ENSURE(h, W_TICKET, sizeof(int) * 4, ticket);
        if (fresh_ticket) CK(h, cudaMemsetAsync(ticket, 0, sizeof(int) * 4, h->stream));
        ticket += 2 * slot;
        bw.ticket = ticket; bw.nwarps = nwarps;
        const unsigned resident = (unsigned)(MINB * h->sm_count);
        const unsigned grid = (h->lg_persist && nblk > resident) ? resident : nblk;
        const size_t smem = SMP::bytes(WPB);
        auto kern = small_backward_lg_kernel<T, P, NORM, UPL, WPB, MINB>;
        const unsigned abit = sizeof(T) == 4 ? 1u : 2u;                  // once per handle = per device and launch shape
        if (!(h->attr_done & abit)) {
            CK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            h->attr_done |= abit;
        }
        // a host entry point's target copy may still be in flight (kanode_api.cu): solve forward on u0 alone, keep the
        // predictions, and form dL/du(t_s) and the loss once the copy has landed
        const bool late = h->target_late && d_target && !d_rp_fwd;
        if (late) {
            if (!a.out) { T* pred = nullptr; ENSURE(h, W_OUT, sizeof(T) * (size_t)nsave * I * B, pred); a.out = pred; }
            a.target = nullptr;
        }
        if (late) { if (int rcs = start_late_target(h)) return rcs; }    // every small upload of this call is already submitted
        else if (int rcj = join_late_target(h)) return rcj;              // (replay: the forward kernel reads the target itself)
        cudaEventRecord(h->ev[0], h->stream);
        small_forward_kernel<T, P, NORM, true, true><<<blocks_for(B, 64), 64, 0, h->stream>>>(prm, a);
        if (late) {
            if (int rcj = join_late_target(h)) return rcj;
            const int64_t cnt = (int64_t)B * nsave * I;
            const unsigned nb = (unsigned)std::min<int64_t>((cnt + 255) / 256, (int64_t)8 * h->sm_count);
            loss_dg_kernel<T><<<nb, 256, 0, h->stream>>>(a.out, d_target, cnt, T(2) / (T)((double)I * nsave), dg, d_loss_sum);
            ++h->launches;
        }
        if (order_forked) CK(h, cudaStreamWaitEvent(h->stream, h->order_ev, 0));
        cudaEventRecord(h->ev[1], h->stream);
        kern<<<grid, 32 * WPB, smem, h->stream>>>(prm, bw);
        cudaEventRecord(h->ev[2], h->stream);
        reduce_partials_kernel<T><<<dim3(NSLAB, (P::NP + 255) / 256), 256, 0, h->stream>>>(gpart, nwarps, P::NP, slab);   // rows = warp positions handed out
        reduce_slabs_kernel<T><<<(P::NP + 7) / 8, 256, 0, h->stream>>>(slab, NSLAB, P::NP, d_grad_sum, 1.0);   // a warp per column
        cudaEventRecord(h->ev[3], h->stream);
        h->launches += 4;
        h->ev_valid = true;
        CK(h, cudaGetLastError());
        return 0;
    };
    auto run = [&]<class P, int NORM>() -> int {
        // launch shapes (warps per block, blocks per SM the kernel is compiled for): the default comes from B200 measurements
        // (profiles/, DESIGN.md 5: 4 / 6 / 8 warps per SM at 255 registers = 2.68 / 2.22 / 1.93 ms; 9 warps at 224 registers
        // 2.33 ms, 10 at 200 registers 2.74 ms — fewer registers cost more than the extra warps hide);
        // KANODE_LG_SHAPE selects another one for A/B runs
        if constexpr (sizeof(T) == 4) {
            switch (h->lg_shape) {
                case 1: return launch.template operator()<P, NORM, 2, 4, 3>();      // 12 warps/SM at 168 registers (spills): slower on B200
                default: return launch.template operator()<P, NORM, 2, KANODE_LG_WPB, KANODE_LG_MINB>();   // 8 warps/SM, no spills
            }
        } else {
            return launch.template operator()<P, NORM, 1, 4, 2>();
        }
    };
    *handled = small_dispatch<T>(h, run, rc);
    return rc;
}

template int small_lg_loss_grad<float>(kanode_handle*, const float*, int64_t, double, double, const double*, int, const float*, double,
                                       double, double*, float*, float*, kanode_stats*, kanode_stats*, float*, const double*,
                                       const double*, int, bool*);
template int small_lg_loss_grad<double>(kanode_handle*, const double*, int64_t, double, double, const double*, int, const double*, double,
                                        double, double*, double*, double*, kanode_stats*, kanode_stats*, double*, const double*,
                                        const double*, int, bool*);

}  // namespace kanode
