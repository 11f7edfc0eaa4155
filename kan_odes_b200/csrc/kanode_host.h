// kanode_host.h — handle definition and host-side helpers shared by kanode_api.cu and the kernel front-ends.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/kanode.h"

namespace kanode {
// layer table for the generic (block-per-trajectory) kernels: device-friendly POD, passed as a kernel parameter
constexpr int GEN_MAX_G = 32;
struct GenericLayer {
    int I, O, G, norm, basis, use_base;
    int act;                   // base-branch activation of the inputs: 0 = swish (KDense), 1 = identity, 2 = tanh (Dense layers: the
                               // output activation of the layer before acts on this layer's inputs)
    int bias;                  // 1: a virtual input unit I with feature 1 carries the bias (its weights W[I*O + o] = b[o])
    float inv_h;               // Float32 1/h (utils.jl:9)
    int goff;                  // offset of this layer's grid points in GenericModel::grid
    long long offC, offW;      // offsets into the flat parameter vector
    long long rx, ry;          // offsets of x_l (layer input) and ybar_l (output cotangent) in a backward stage record
};
struct GenericModel {
    int n_layers, rhs_kind, n, n_out;   // n_out: output length of one sample (= n except for KANODE_RHS_MAP)
    long long np;
    double lap_scale;          // lap_coef / dx^2
    long long rec_len;         // length of one backward stage record
    GenericLayer L[KANODE_MAX_LAYERS];
    float grid[KANODE_MAX_LAYERS * GEN_MAX_G];   // Float32 LinRange points per layer
};
}  // namespace kanode

inline thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct kanode_handle {
    kanode_desc desc{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    size_t np = 0;
    int n = 0;
    std::vector<double> params;      // host copy (drives the __grid_constant__ parameter blocks)
    bool have_params = false;
    uint64_t params_version = 0;     // bumped by set_params; derived device copies are refreshed lazily
    uint64_t wpk_version[2] = {~0ull, ~0ull};
    uint64_t wlg_version[2] = {~0ull, ~0ull};   // lane-block weight image of the lane-group adjoint kernel
    int bwd_maxiters = 100000;       // KANODE_BWD_MAXIT (timing experiments)
    double reg_act = 0.0, reg_entropy = 0.0;   // kanode_set_regularizer
    // device-resident training (kanode_train_*): Adam hyper-parameters and step count; the fp32 master parameters and the
    // moments live in W_TR_P / W_TR_M / W_TR_V.  params_host_stale: the device copies are ahead of `params`
    bool train_on = false, params_host_stale = false;
    float tr_eta = 0.f, tr_b1 = 0.9f, tr_b2 = 0.999f, tr_eps = 1e-8f;
    int64_t tr_t = 0;
    int n_out = 0;                   // output length of one sample of kanode_rhs (= n except for KANODE_RHS_MAP)
    // kanode_create_multi: this handle routes; one child handle (own stream, own workspace) per device
    std::vector<kanode_handle*> children;
    bool peer_ok = false;            // device children[0] can load from every other child's memory (NVLink / PCIe P2P)
    int last_failed[2] = {0, 0};     // failed forward / adjoint solves of the last host-pointer loss_grad call
    unsigned attr_done = 0;          // per-handle (= per-device) one-time cudaFuncSetAttribute bits
    int rec_cap = 32;
    int64_t order_B[2] = {0, 0};     // batch size the cached launch order (per dtype) was built for; 0 = none
    int schedule = 1;                // 1: launch order of the adjoint warps from the last call's iteration counts (KANODE_SCHEDULE)
    int sm_count = 148;              // multiprocessors of the device (grid of the persistent adjoint kernel)
    int lg_persist = 1;              // 1: persistent adjoint launch, warps draw tickets (KANODE_LG_PERSIST=0: one block per 4 warp positions)
    int lg_shape = 0;                // launch shape of the lane-group adjoint kernel (KANODE_LG_SHAPE; 0 = default)
    int64_t launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // fwd start / bwd start / reduce start / end
    // host entry points: the target (the bulk of a step's host->device bytes) is copied on aux_stream while the forward solve
    // already runs on the u0 that arrived first; aux_ev[0] = fork (aux waits for the main stream), aux_ev[1] = target landed.
    // target_late: the engine must wait for aux_ev[1] before it reads the target (kanode_api.cu: join_late_target)
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t aux_ev[2] = {nullptr, nullptr};
    cudaEvent_t order_ev = nullptr;  // launch-order kernels of the lane-group adjoint (on aux_stream, under the forward solve) done
    bool target_late = false, late_started = false;
    const void* late_src = nullptr; void* late_dst = nullptr; size_t late_bytes = 0;   // the deferred target copy
    int overlap_h2d = 1;             // KANODE_OVERLAP_H2D=0: every copy on the main stream
    bool ev_valid = false;
    std::string err;
    void* stage = nullptr; size_t stage_bytes = 0;   // pinned host staging block for the results of the host entry points
    // peer-memory all-reduce (kanode_peer.cu): this rank's mailbox (own cudaMalloc: exported over CUDA IPC), the peers' mailboxes
    // as mapped here, the call counter (epoch) and a device error word
    void* peer_box = nullptr;
    void* peer_map[KANODE_PEER_MAX_WORLD] = {};
    int peer_rank = -1, peer_world = 0;
    unsigned long long peer_epoch = 0;
    int* peer_err = nullptr;
    // grow-only device workspace, keyed by purpose
    enum { W_U0, W_OUT, W_TARGET, W_STATS_F, W_STATS_B, W_SAVEAT, W_REC_T, W_REC, W_NSTEPS, W_RET, W_DG, W_FAC, W_G,
           W_LOSS, W_GRAD, W_DU0, W_PARAMS, W_PARAMS64, W_WPK32, W_WPK64, W_LAM, W_GEN, W_GEN2, W_LS, W_ATT, W_ORDER, W_WIDE_F, W_WIDE_B, W_W1T32, W_W1T64, W_W2IMG, W_W2TIMG, W_W1IMG, W_WIDE_R, W_WLG32, W_WLG64, W_GPART, W_SLAB, W_RPF, W_RPB, W_COT, W_FAILCNT, W_REG, W_ACT, W_TR_P, W_TR_M, W_TR_V, W_TR_GRAD, W_TR_OUT, W_TR_RAW, W_MULTI_G, W_MULTI_L, W_MULTI_STAGE, W_TICKET, W_COUNT };
    DevBuf ws[W_COUNT];
    kanode::GenericModel gm{};               // layer table for the generic kernels
    // wide (batched lockstep) engine: attempts the last forward-only / dense-forward / backward call needed, counter state
    int wide = 1;                            // 0: force the block-per-trajectory kernels (KANODE_WIDE=0)
    int wide_iters[3] = {0, 0, 0};
    uint64_t wide_w1t_version[2] = {~0ull, ~0ull};
    uint64_t wide_w2img_version = ~0ull, wide_w2timg_version = ~0ull, wide_w1img_version = ~0ull;
    int wide_tc = 1;                         // fp32: 1 = layer-2 contractions on tcgen05, 2 = also the batched layer-1 forward, 0 = CUDA cores (KANODE_WIDE_TC)
    // one lockstep step attempt captured as a CUDA graph (forward-only, dense forward, backward) x (fp32, fp64): replayed
    // per attempt while the kernel arguments (workspace pointers, sizes, tolerances) stay the same
    struct WideGraph { cudaGraphExec_t exec = nullptr; std::vector<char> sig; int nodes = 0; };
    WideGraph wide_graphs[6];
    int wide_graph = 1;                      // KANODE_WIDE_GRAPH=0: launch every kernel directly
    int wide_graph_maxn = 1 << 30;           // KANODE_WIDE_GRAPH_MAXN: states above this launch directly (keeps the per-pass CUDA events: profiling)
    std::vector<cudaEvent_t> wide_gp_ev;     // event pairs around the g passes of the last wide loss_grad call
    int wide_gp_used = 0;
    bool wide_counters_zeroed[2] = {false, false};
    const void* wide_counters_ptr[2] = {nullptr, nullptr};
};

namespace kanode { void peer_release(kanode_handle* h); }   // kanode_peer.cu: closes the mapped mailboxes, frees the own one

inline int fail(kanode_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CK(h, call)                                                                                         \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess)                                                                              \
            return fail(h, e_ == cudaErrorMemoryAllocation ? KANODE_ERR_NOMEM : KANODE_ERR_CUDA, "%s: %s",  \
                        #call, cudaGetErrorString(e_));                                                     \
    } while (0)

// Deferred target copy of the host entry points.  The copy engine serves host->device copies in submission order, so the big
// copy must be SUBMITTED after the small ones the engine still has to make (save times, weight images) or the forward solve
// would wait behind it: the host entry point only registers the copy, the engine starts it on aux_stream right before its
// forward launch (start_late_target) and joins when it first reads the target (join_late_target; an engine that reads the
// target in its forward kernels joins first, which degenerates to a plain copy on the main stream).
inline int start_late_target(kanode_handle* h) {
    if (!h->target_late || h->late_started) return 0;
    CK(h, cudaEventRecord(h->aux_ev[0], h->stream));                   // behind everything already queued on the main stream
    CK(h, cudaStreamWaitEvent(h->aux_stream, h->aux_ev[0], 0));
    CK(h, cudaMemcpyAsync(h->late_dst, h->late_src, h->late_bytes, cudaMemcpyHostToDevice, h->aux_stream));
    CK(h, cudaEventRecord(h->aux_ev[1], h->aux_stream));
    h->late_started = true;
    return 0;
}
inline int join_late_target(kanode_handle* h) {
    if (!h->target_late) return 0;
    h->target_late = false;
    if (h->late_started) { h->late_started = false; CK(h, cudaStreamWaitEvent(h->stream, h->aux_ev[1], 0)); }
    else CK(h, cudaMemcpyAsync(h->late_dst, h->late_src, h->late_bytes, cudaMemcpyHostToDevice, h->stream));
    return 0;
}

inline size_t count_params(const kanode_desc* d) {
    if (!d || d->n_layers < 1 || d->n_layers > KANODE_MAX_LAYERS) return 0;
    size_t np = 0;
    for (int l = 0; l < d->n_layers; ++l) {
        const kanode_layer_desc& s = d->layers[l];
        if (l > 0 && d->layers[l - 1].out_dims != s.in_dims) return 0;
        if (s.kind == KANODE_LAYER_DENSE) {                           // Lux.Dense: weight[out, in] + bias[out]
            if (s.in_dims < 1 || s.out_dims < 1 || s.dense_act < 0 || s.dense_act > 1) return 0;
            if (l == d->n_layers - 1 && s.dense_act != KANODE_ACT_IDENTITY) return 0;
            np += (size_t)(s.in_dims + 1) * s.out_dims;
            continue;
        }
        if (s.kind != KANODE_LAYER_KDENSE) return 0;
        if (s.in_dims < 1 || s.out_dims < 1 || s.grid_len < 2) return 0;
        if (s.normalizer < 0 || s.normalizer > 2 || s.basis < 0 || s.basis > 2) return 0;
        if (!(s.grid_hi > s.grid_lo) || !(s.denominator > 0)) return 0;
        np += (size_t)s.in_dims * s.grid_len * s.out_dims;            // kdense.jl:101
        if (s.use_base_act) np += (size_t)s.in_dims * s.out_dims;     // kdense.jl:103
    }
    if (d->rhs_kind == KANODE_RHS_CHAIN) {
        if (d->layers[0].in_dims != d->n_state || d->layers[d->n_layers - 1].out_dims != d->n_state) return 0;
    } else if (d->rhs_kind == KANODE_RHS_SOURCE_LAPLACIAN) {
        if (d->layers[0].in_dims != 1 || d->layers[d->n_layers - 1].out_dims != 1 || d->n_state < 3 || !(d->dx > 0)) return 0;
    } else if (d->rhs_kind == KANODE_RHS_MAP) {
        if (d->layers[0].in_dims != d->n_state) return 0;
    } else return 0;
    return np;
}

// Julia LinRange{Float32}(lo, hi, G)[g+1]  (kdense.jl:90)
inline float grid_point(const kanode_layer_desc& s, int g) {
    const double t = (double)g / (double)(s.grid_len - 1);
    return (float)((1.0 - t) * (double)s.grid_lo + t * (double)s.grid_hi);
}

inline int ensure(kanode_handle* h, int which, size_t bytes, void** out) {
    DevBuf& b = h->ws[which];
    if (b.bytes < bytes) {
        if (b.p) { CK(h, cudaStreamSynchronize(h->stream)); CK(h, cudaFree(b.p)); b.p = nullptr; b.bytes = 0; }
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) { (void)cudaGetLastError(); want = bytes; e = cudaMalloc(&b.p, want); }
        if (e != cudaSuccess) { b.p = nullptr; return fail(h, KANODE_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
        b.bytes = want;
    }
    *out = b.p;
    return 0;
}
#define ENSURE(h, which, bytes, ptr)                                   \
    do {                                                               \
        void* p_ = nullptr;                                            \
        int rc_ = ensure(h, kanode_handle::which, (bytes), &p_);       \
        if (rc_) return rc_;                                           \
        ptr = reinterpret_cast<decltype(ptr)>(p_);                     \
    } while (0)

