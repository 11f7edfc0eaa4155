// kanode_api.cu — C ABI (include/kanode.h) over the sm_100a kernels.  Plain pointers and sizes only; no torch,
// no C++ exceptions across the boundary, no CPU fallback.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

#include "kanode_host.h"
#include "kanode_small_host.h"
#include "kanode_generic.cuh"
#include "kanode_wide_api.h"

using namespace kanode;

namespace {

int enter(kanode_handle* h) {
    if (!h) return fail(nullptr, KANODE_ERR_INVALID, "null handle");
    CK(h, cudaSetDevice(h->children.empty() ? h->device : h->children[0]->device));
    return 0;
}

int check_saveat(kanode_handle* h, double t0, double t1, const double* saveat, int nsave) {
    if (!(t1 >= t0)) return fail(h, KANODE_ERR_INVALID, "tspan must be increasing");
    if (nsave < 0 || (nsave > 0 && !saveat)) return fail(h, KANODE_ERR_INVALID, "bad saveat");
    for (int s = 0; s < nsave; ++s) {
        if (!(saveat[s] >= t0 && saveat[s] <= t1)) return fail(h, KANODE_ERR_INVALID, "saveat[%d] outside tspan", s);
        if (s > 0 && saveat[s] < saveat[s - 1]) return fail(h, KANODE_ERR_INVALID, "saveat must be ascending");
    }
    return 0;
}

__global__ void __launch_bounds__(256) adam_update_kernel(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ m,
                                                          float* __restrict__ v, size_t n, float eta, float b1, float b2,
                                                          float eps, float c1, float c2, float gs) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = gs * grad[i];
    const float mi = b1 * m[i] + (1.0f - b1) * g;
    const float vi = b2 * v[i] + (1.0f - b2) * g * g;
    m[i] = mi; v[i] = vi;
    p[i] -= eta * (mi * c1) / (sqrtf(vi * c2) + eps);
}

// ---------------------------------------------------------------------------------------------------------
// templated implementations (T = float | double); all pointers are DEVICE pointers here
// ---------------------------------------------------------------------------------------------------------
template <class T> int put_saveat(kanode_handle* h, const double* saveat, int nsave, const double** d_saveat) {
    double* d = nullptr;
    ENSURE(h, W_SAVEAT, sizeof(double) * (size_t)(nsave > 0 ? nsave : 1), d);
    if (nsave > 0) CK(h, cudaMemcpyAsync(d, saveat, sizeof(double) * nsave, cudaMemcpyHostToDevice, h->stream));
    *d_saveat = d;
    return 0;
}

// The lockstep engines index with 32-bit offsets and put the batch on grid.y: larger batches take the block-per-trajectory kernels
inline bool wide_batch_ok(const kanode_handle* h, int64_t B) { return B <= 65535 && (int64_t)B * h->n * 8 < (1ll << 31); }

template <class T> int rhs_dev(kanode_handle* h, const T* d_u, T* d_du, int64_t B) {
    if (!h->have_params) return fail(h, KANODE_ERR_INVALID, "parameters not set");
    if (B <= 0) return 0;
    int rc = 0;
    auto run = [&]<class P, int NORM>() -> int {
        if (int rcr = host_params_refresh(h)) return rcr;              // the weights travel in the kernel parameter block
        P prm; fill_small<T>(h, prm);
        small_rhs_kernel<T, P, NORM><<<blocks_for(B, 128), 128, 0, h->stream>>>(prm, d_u, d_du, B);
        ++h->launches;
        CK(h, cudaGetLastError());
        return 0;
    };
    if (small_dispatch<T>(h, run, rc)) return rc;
    WideKey wk;
    if (h->wide && wide_match(h->desc, wk) && wide_batch_ok(h, B)) return wide_rhs(h, wk, generic_params<T>(h), d_u, d_du, B);
    return generic_rhs<T>(h, d_u, d_du, B);
}

template <class T> int vjp_dev(kanode_handle* h, const T* d_u, const T* d_lam, T* d_ubar, T* d_pbar, int64_t B) {
    if (!h->have_params) return fail(h, KANODE_ERR_INVALID, "parameters not set");
    if (B <= 0) { CK(h, cudaMemsetAsync(d_pbar, 0, sizeof(T) * h->np, h->stream)); return 0; }
    int rc = 0;
    auto run = [&]<class P, int NORM>() -> int {
        if (int rcr = host_params_refresh(h)) return rcr;
        P prm; fill_small<T>(h, prm);
        T *fac = nullptr, *rows = nullptr;
        ENSURE(h, W_FAC, sizeof(T) * (size_t)P::NF * B, fac);
        ENSURE(h, W_G, sizeof(T) * (size_t)P::NP * B, rows);
        small_vjp_kernel<T, P, NORM><<<blocks_for(B, 64), 64, 0, h->stream>>>(prm, d_u, d_lam, d_ubar, fac, rows, B);
        reduce_rows_kernel<T, T><<<P::NP, 256, 0, h->stream>>>(rows, B, d_pbar, 1.0);
        h->launches += 2;
        CK(h, cudaGetLastError());
        return 0;
    };
    if (small_dispatch<T>(h, run, rc)) return rc;
    return generic_vjp<T>(h, d_u, d_lam, d_ubar, d_pbar, B);
}

template <class T>
int solve_dev(kanode_handle* h, const T* d_u0, int64_t B, double t0, double t1, const double* saveat, int nsave,
              double abstol, double reltol, T* d_out, kanode_stats* d_stats) {
    if (!h->have_params) return fail(h, KANODE_ERR_INVALID, "parameters not set");
    if (h->desc.rhs_kind == KANODE_RHS_MAP) return fail(h, KANODE_ERR_UNSUPPORTED, "a KANODE_RHS_MAP handle is a map, not an ODE right-hand side");
    if (int rc = check_saveat(h, t0, t1, saveat, nsave)) return rc;
    if (B <= 0) return 0;
    const double* d_saveat = nullptr;
    if (int rc = put_saveat<T>(h, saveat, nsave, &d_saveat)) return rc;
    int rc = 0;
    auto run = [&]<class P, int NORM>() -> int {
        P prm; fill_small<T>(h, prm);
        SmallFwdArgs<T> a{};
        if (int rcw = upload_packed<T, P>(h, &a.wpk)) return rcw;
        a.u0 = d_u0; a.B = B; a.t0 = t0; a.t1 = t1; a.saveat = d_saveat; a.nsave = nsave;
        a.abstol = (T)abstol; a.reltol = (T)reltol; a.maxiters = 100000; a.out = d_out; a.stats = d_stats;
        small_forward_kernel<T, P, NORM, false><<<blocks_for(B, 64), 64, 0, h->stream>>>(prm, a);
        ++h->launches;
        CK(h, cudaGetLastError());
        return 0;
    };
    if (small_dispatch<T>(h, run, rc)) return rc;
    WideKey wk;
    if (h->wide && wide_match(h->desc, wk) && wide_batch_ok(h, B))
        return wide_solve(h, wk, generic_params<T>(h), d_u0, B, t0, t1, d_saveat, nsave, abstol, reltol, d_out, d_stats);
    int sg = 0;
    if (h->wide && wsrc_match(h->desc, sg) && wide_batch_ok(h, B))
        return wsrc_solve(h, sg, generic_params<T>(h), d_u0, B, t0, t1, d_saveat, nsave, abstol, reltol, d_out, d_stats);
    return generic_solve<T>(h, d_u0, B, t0, t1, d_saveat, nsave, abstol, reltol, d_out, d_stats);
}

template <class T> int reg_apply(kanode_handle* h, double* d_loss, double loss_scale, T* d_grad, double grad_scale);

template <class T>
int loss_grad_dev(kanode_handle* h, const T* d_u0, int64_t B, double t0, double t1, const double* saveat, int nsave,
                  const T* d_target, double abstol, double reltol, double* d_loss_sum, T* d_grad_sum, T* d_du0,
                  kanode_stats* d_fst, kanode_stats* d_bst, T* d_out_opt, const double* d_rp_fwd = nullptr,
                  const double* d_rp_bwd = nullptr, int rp_cap = 0, const T* d_cot = nullptr) {
    if (!h->have_params) return fail(h, KANODE_ERR_INVALID, "parameters not set");
    if (h->desc.rhs_kind == KANODE_RHS_MAP) return fail(h, KANODE_ERR_UNSUPPORTED, "a KANODE_RHS_MAP handle is a map, not an ODE right-hand side");
    if (int rc = check_saveat(h, t0, t1, saveat, nsave)) return rc;
    if (nsave < 1) return fail(h, KANODE_ERR_INVALID, "loss needs at least one save time");
    h->wide_gp_used = 0;
    CK(h, cudaMemsetAsync(d_loss_sum, 0, sizeof(double), h->stream));
    if (B <= 0) { CK(h, cudaMemsetAsync(d_grad_sum, 0, sizeof(T) * h->np, h->stream)); return 0; }
    const double* d_saveat = nullptr;
    if (int rc = put_saveat<T>(h, saveat, nsave, &d_saveat)) return rc;
    if (d_cot) {
        // pullback with caller-supplied cotangents: every engine keeps dL/du(t_s) in W_DG as [B][nsave][n] and, given a null
        // target, leaves it alone (it only zeroes the entries of failed trajectories): pre-fill it and drop the target
        T* dg = nullptr;
        ENSURE(h, W_DG, sizeof(T) * (size_t)nsave * h->n * B, dg);
        CK(h, cudaMemcpyAsync(dg, d_cot, sizeof(T) * (size_t)nsave * h->n * B, cudaMemcpyDeviceToDevice, h->stream));
        d_target = nullptr;
    } else if (!d_target) return fail(h, KANODE_ERR_INVALID, "null target");
    auto engine = [&]() -> int {
        {   // small-model registry: lane-group adjoint engine (kanode_lg.cu); it runs the forward solve BEFORE it joins a late target
            bool handled = false;
            const int rcl = small_lg_loss_grad<T>(h, d_u0, B, t0, t1, d_saveat, nsave, d_target, abstol, reltol, d_loss_sum, d_grad_sum,
                                                  d_du0, d_fst, d_bst, d_out_opt, d_rp_fwd, d_rp_bwd, rp_cap, &handled);
            if (handled) return rcl;
            if (d_rp_fwd || d_rp_bwd) return fail(h, KANODE_ERR_UNSUPPORTED, "dt-replay is implemented by the small-model ensemble kernels only");
        }
        if (int rcj = join_late_target(h)) return rcj;                   // the other engines read the target in their forward kernels
        WideKey wk;
        if (h->wide && wide_match(h->desc, wk) && wide_batch_ok(h, B))
            return wide_loss_grad(h, wk, generic_params<T>(h), d_u0, B, t0, t1, d_saveat, nsave, d_target, abstol, reltol, d_loss_sum,
                                  d_grad_sum, d_du0, d_fst, d_bst, d_out_opt);
        int sg = 0;
        if (h->wide && wsrc_match(h->desc, sg) && wide_batch_ok(h, B))
            return wsrc_loss_grad(h, sg, generic_params<T>(h), d_u0, B, t0, t1, d_saveat, nsave, d_target, abstol, reltol, d_loss_sum,
                                  d_grad_sum, d_du0, d_fst, d_bst, d_out_opt);
        return generic_loss_grad<T>(h, d_u0, B, t0, t1, d_saveat, nsave, d_target, abstol, reltol, d_loss_sum, d_grad_sum,
                                    d_du0, d_fst, d_bst, d_out_opt);
    };
    const int rc_engine = engine();
    if (int rcj = join_late_target(h)) return rcj;                       // whatever path was taken: the main stream is behind the copy
    if (rc_engine) return rc_engine;
    // sparsity regulariser (reg_loss, LV_driver_KANODE.jl:187-201): the sums are un-normalised (caller divides the loss by
    // B*nsave*n and the gradient by B), so reg enters with those factors
    if (!d_cot && (h->reg_act != 0.0 || h->reg_entropy != 0.0))
        return reg_apply<T>(h, d_loss_sum, (double)B * nsave * h->n, d_grad_sum, (double)B);
    return 0;
}

template <class T> int reg_apply(kanode_handle* h, double* d_loss, double loss_scale, T* d_grad, double grad_scale) {
    double* acc = nullptr;
    ENSURE(h, W_REG, 2 * sizeof(double), acc);
    CK(h, cudaMemsetAsync(acc, 0, 2 * sizeof(double), h->stream));
    const T* p = generic_params<T>(h);
    const unsigned nb = (unsigned)std::min<size_t>((h->np + 255) / 256, 1184);
    reg_sums_kernel<T><<<nb, 256, 0, h->stream>>>(p, h->np, acc);
    reg_apply_kernel<T><<<blocks_for((int64_t)h->np, 256), 256, 0, h->stream>>>(p, h->np, acc, h->reg_act, h->reg_entropy, d_loss, loss_scale,
                                                                                 d_grad, grad_scale);
    h->launches += 2;
    CK(h, cudaGetLastError());
    return 0;
}

// failed[0] = forward solves that did not return Success, [1] = of those, dense-record overflows, [2] = failed adjoint solves
__global__ void __launch_bounds__(256) stats_scan_kernel(const kanode_stats* __restrict__ f, const kanode_stats* __restrict__ b, int64_t B,
                                                         int* __restrict__ failed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int ff = 0, fo = 0, fb = 0;
    if (i < B) { ff = f[i].retcode != KANODE_RET_SUCCESS; fo = f[i].retcode == KANODE_RET_RECORD_OVERFLOW; fb = b[i].retcode != KANODE_RET_SUCCESS; }
    ff = __syncthreads_count(ff); fo = __syncthreads_count(fo); fb = __syncthreads_count(fb);
    if (threadIdx.x == 0) {
        if (ff) atomicAdd(&failed[0], ff);
        if (fo) atomicAdd(&failed[1], fo);
        if (fb) atomicAdd(&failed[2], fb);
    }
}

// ---- host-pointer wrappers -----------------------------------------------------------------------------------
template <class T> int set_params_host(kanode_handle* h, const T* p, size_t np) {
    if (int rc = enter(h)) return rc;
    if (!p || np != h->np) return fail(h, KANODE_ERR_INVALID, "expected %zu parameters, got %zu", h->np, np);
    for (size_t i = 0; i < np; ++i) h->params[i] = (double)p[i];
    h->have_params = true; h->params_host_stale = false;
    ++h->params_version;
    return generic_upload_params(h);
}

template <class T> int rhs_host(kanode_handle* h, const T* u, T* du, int64_t B) {
    if (int rc = enter(h)) return rc;
    if (B < 0 || (B > 0 && (!u || !du))) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    if (B == 0) return 0;
    const size_t bytes = sizeof(T) * (size_t)B * h->n, obytes = sizeof(T) * (size_t)B * h->n_out;   // n_out != n only for KANODE_RHS_MAP
    T *d_u = nullptr, *d_du = nullptr;
    ENSURE(h, W_U0, bytes, d_u); ENSURE(h, W_OUT, obytes, d_du);
    CK(h, cudaMemcpyAsync(d_u, u, bytes, cudaMemcpyHostToDevice, h->stream));
    if (int rc = rhs_dev<T>(h, d_u, d_du, B)) return rc;
    CK(h, cudaMemcpyAsync(du, d_du, obytes, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

template <class T> int vjp_host(kanode_handle* h, const T* u, const T* lam, T* ubar, T* pbar, int64_t B) {
    if (int rc = enter(h)) return rc;
    if (B < 0 || !pbar || (B > 0 && (!u || !lam || !ubar))) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    const size_t bytes = sizeof(T) * (size_t)(B > 0 ? B : 1) * h->n, lbytes = sizeof(T) * (size_t)(B > 0 ? B : 1) * h->n_out;
    T *d_u = nullptr, *d_lam = nullptr, *d_ub = nullptr, *d_pb = nullptr;
    ENSURE(h, W_U0, bytes, d_u); ENSURE(h, W_LAM, lbytes, d_lam); ENSURE(h, W_OUT, bytes, d_ub);
    ENSURE(h, W_GRAD, sizeof(T) * h->np, d_pb);
    if (B > 0) {
        CK(h, cudaMemcpyAsync(d_u, u, sizeof(T) * (size_t)B * h->n, cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaMemcpyAsync(d_lam, lam, sizeof(T) * (size_t)B * h->n_out, cudaMemcpyHostToDevice, h->stream));   // cotangent of the output
    }
    if (int rc = vjp_dev<T>(h, d_u, d_lam, d_ub, d_pb, B)) return rc;
    if (B > 0) CK(h, cudaMemcpyAsync(ubar, d_ub, sizeof(T) * (size_t)B * h->n, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(pbar, d_pb, sizeof(T) * h->np, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

template <class T>
int solve_host(kanode_handle* h, const T* u0, int64_t B, double t0, double t1, const double* saveat, int nsave,
               double abstol, double reltol, T* out, kanode_stats* stats) {
    if (int rc = enter(h)) return rc;
    if (B < 0 || (B > 0 && (!u0 || (nsave > 0 && !out)))) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    if (B == 0) return 0;
    T *d_u0 = nullptr, *d_out = nullptr; kanode_stats* d_st = nullptr;
    ENSURE(h, W_U0, sizeof(T) * (size_t)B * h->n, d_u0);
    ENSURE(h, W_OUT, sizeof(T) * (size_t)B * (nsave > 0 ? nsave : 1) * h->n, d_out);
    ENSURE(h, W_STATS_F, sizeof(kanode_stats) * (size_t)B, d_st);
    CK(h, cudaMemcpyAsync(d_u0, u0, sizeof(T) * (size_t)B * h->n, cudaMemcpyHostToDevice, h->stream));
    if (int rc = solve_dev<T>(h, d_u0, B, t0, t1, saveat, nsave, abstol, reltol, d_out, d_st)) return rc;
    if (nsave > 0) CK(h, cudaMemcpyAsync(out, d_out, sizeof(T) * (size_t)B * nsave * h->n, cudaMemcpyDeviceToHost, h->stream));
    if (stats) CK(h, cudaMemcpyAsync(stats, d_st, sizeof(kanode_stats) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

template <class T>
int loss_grad_host(kanode_handle* h, const T* u0, int64_t B, double t0, double t1, const double* saveat, int nsave,
                   const T* target, double abstol, double reltol, T* loss, T* grad, T* du0, kanode_stats* fst,
                   kanode_stats* bst, const double* rp_fwd = nullptr, const double* rp_bwd = nullptr, int rp_cap = 0,
                   T* out = nullptr, const T* cot = nullptr, bool keep_on_device = false) {
    // cot != null: pullback of the solve with the caller's cotangents dL/dpred (kanode_solve_adjoint): no target, no loss,
    // and the gradient is the plain sum over the batch.  keep_on_device: the un-normalised gradient / loss sums stay in
    // W_GRAD / W_LOSS for the multi-device combine (loss and grad are not written).
    if (int rc = enter(h)) return rc;
    if (B <= 0 || !u0 || (!target && !cot) || (!keep_on_device && ((!loss && !cot) || !grad))) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    if ((rp_fwd || rp_bwd) && (!rp_fwd || !rp_bwd || rp_cap < 1)) return fail(h, KANODE_ERR_INVALID, "replay needs both step sequences");
    const size_t nout = (size_t)B * nsave * h->n;
    for (int attempt = 0;; ++attempt) {
        T *d_u0 = nullptr, *d_tg = nullptr, *d_grad = nullptr, *d_du0 = nullptr; double* d_loss = nullptr;
        kanode_stats *d_f = nullptr, *d_b = nullptr;
        ENSURE(h, W_U0, sizeof(T) * (size_t)B * h->n, d_u0);
        ENSURE(h, W_TARGET, sizeof(T) * (nout ? nout : 1), d_tg);
        ENSURE(h, W_GRAD, sizeof(T) * h->np, d_grad);
        ENSURE(h, W_DU0, sizeof(T) * (size_t)B * h->n, d_du0);
        ENSURE(h, W_LOSS, sizeof(double), d_loss);
        ENSURE(h, W_STATS_F, sizeof(kanode_stats) * (size_t)B, d_f);
        ENSURE(h, W_STATS_B, sizeof(kanode_stats) * (size_t)B, d_b);
        CK(h, cudaMemcpyAsync(d_u0, u0, sizeof(T) * (size_t)B * h->n, cudaMemcpyHostToDevice, h->stream));
        T* d_cot = nullptr;
        if (cot) {
            ENSURE(h, W_COT, sizeof(T) * (nout ? nout : 1), d_cot);
            CK(h, cudaMemcpyAsync(d_cot, cot, sizeof(T) * nout, cudaMemcpyHostToDevice, h->stream));
        } else if (h->overlap_h2d && h->aux_stream && (h->overlap_h2d > 1 || sizeof(T) * nout >= (1u << 20))) {   // 2: any size (tests)
            // the target is the bulk of the step's input bytes and the forward solve needs only u0: register the copy, the
            // engine starts it on the second stream behind its own small uploads and joins when it first reads the target
            h->late_src = target; h->late_dst = d_tg; h->late_bytes = sizeof(T) * nout;
            h->target_late = true; h->late_started = false;
        } else CK(h, cudaMemcpyAsync(d_tg, target, sizeof(T) * nout, cudaMemcpyHostToDevice, h->stream));
        double *d_rpf = nullptr, *d_rpb = nullptr; T* d_out = nullptr;
        if (rp_fwd) {
            ENSURE(h, W_RPF, sizeof(double) * (size_t)B * rp_cap, d_rpf);
            ENSURE(h, W_RPB, sizeof(double) * (size_t)B * rp_cap, d_rpb);
            CK(h, cudaMemcpyAsync(d_rpf, rp_fwd, sizeof(double) * (size_t)B * rp_cap, cudaMemcpyHostToDevice, h->stream));
            CK(h, cudaMemcpyAsync(d_rpb, rp_bwd, sizeof(double) * (size_t)B * rp_cap, cudaMemcpyHostToDevice, h->stream));
        }
        if (out) ENSURE(h, W_OUT, sizeof(T) * (nout ? nout : 1), d_out);
        if (int rc = loss_grad_dev<T>(h, d_u0, B, t0, t1, saveat, nsave, d_tg, abstol, reltol, d_loss, d_grad, d_du0,
                                      d_f, d_b, d_out, d_rpf, d_rpb, rp_cap, d_cot)) return rc;
        if (out) CK(h, cudaMemcpyAsync(out, d_out, sizeof(T) * nout, cudaMemcpyDeviceToHost, h->stream));
        // every result goes to ONE pinned staging block with async copies and a single synchronisation; the per-trajectory
        // statistics travel only when the caller asked for them (a 3-counter scan tells whether any solve failed)
        int* d_cnt = nullptr;
        ENSURE(h, W_FAILCNT, 3 * sizeof(int), d_cnt);
        CK(h, cudaMemsetAsync(d_cnt, 0, 3 * sizeof(int), h->stream));
        stats_scan_kernel<<<blocks_for(B, 256), 256, 0, h->stream>>>(d_f, d_b, B, d_cnt);
        ++h->launches;
        const size_t o_f = 0, o_b = o_f + sizeof(kanode_stats) * (size_t)B, o_du = o_b + sizeof(kanode_stats) * (size_t)B,
                     o_g = o_du + sizeof(T) * (size_t)B * h->n, o_l = (o_g + sizeof(T) * h->np + 7) / 8 * 8, o_c = o_l + 8, total = o_c + 16;
        if (h->stage_bytes < total) {
            if (h->stage) { CK(h, cudaStreamSynchronize(h->stream)); cudaFreeHost(h->stage); h->stage = nullptr; h->stage_bytes = 0; }
            CK(h, cudaMallocHost(&h->stage, total + total / 8));
            h->stage_bytes = total + total / 8;
        }
        char* st = static_cast<char*>(h->stage);
        if (fst) CK(h, cudaMemcpyAsync(st + o_f, d_f, sizeof(kanode_stats) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
        if (bst) CK(h, cudaMemcpyAsync(st + o_b, d_b, sizeof(kanode_stats) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
        if (du0) CK(h, cudaMemcpyAsync(st + o_du, d_du0, sizeof(T) * (size_t)B * h->n, cudaMemcpyDeviceToHost, h->stream));
        if (!keep_on_device) {
            CK(h, cudaMemcpyAsync(st + o_g, d_grad, sizeof(T) * h->np, cudaMemcpyDeviceToHost, h->stream));
            CK(h, cudaMemcpyAsync(st + o_l, d_loss, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        }
        CK(h, cudaMemcpyAsync(st + o_c, d_cnt, 3 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        int cnt[3]; std::memcpy(cnt, st + o_c, sizeof cnt);
        if (cnt[1] > 0 && attempt < 6) { h->rec_cap *= 4; continue; }   // grow the dense record and redo the step
        if (fst) std::memcpy(fst, st + o_f, sizeof(kanode_stats) * (size_t)B);
        if (bst) std::memcpy(bst, st + o_b, sizeof(kanode_stats) * (size_t)B);
        if (du0) std::memcpy(du0, st + o_du, sizeof(T) * (size_t)B * h->n);
        if (!keep_on_device) {
            const T* gs = reinterpret_cast<const T*>(st + o_g);
            double lsum = 0; std::memcpy(&lsum, st + o_l, sizeof(double));
            if (loss) *loss = (T)(lsum / ((double)B * nsave * h->n));
            const double gdiv = cot ? 1.0 : (double)B;
            for (size_t i = 0; i < h->np; ++i) grad[i] = (T)((double)gs[i] / gdiv);
        }
        h->last_failed[0] = cnt[0]; h->last_failed[1] = cnt[2];
        // In the reference a failed solve gives a short solution and loss() throws (LV_driver_KANODE.jl:197-203); here the
        // failed trajectories are left out of the sums, every output is still written, and the call reports it.
        if (cnt[0] + cnt[2] > 0)
            return fail(h, KANODE_ERR_SOLVER, "%d forward and %d adjoint solves of %lld did not return Success (see the per-trajectory retcodes); "
                        "loss and gradient leave them out", cnt[0], cnt[2], (long long)B);
        return 0;
    }
}

template <class T>
__global__ void __launch_bounds__(256) pack_sums_kernel(const T* __restrict__ g, const double* __restrict__ loss, double count, size_t np, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) out[i] = (double)g[i];
    else if (i == np) { out[np] = *loss; out[np + 1] = count; }
}
template <class T> int pack_sums(kanode_handle* h, const T* d_grad, const double* d_loss, int64_t count, double* d_packed) {
    if (int rc = enter(h)) return rc;
    if (!d_grad || !d_loss || !d_packed || count < 0) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    pack_sums_kernel<T><<<blocks_for((int64_t)h->np + 1, 256), 256, 0, h->stream>>>(d_grad, d_loss, (double)count, h->np, d_packed);
    ++h->launches;
    CK(h, cudaGetLastError());
    return 0;
}
// Adam on the all-reduced packed sums: g = packed[i] / packed[np + 1] (global trajectory count), all on the device
__global__ void __launch_bounds__(256) adam_packed_kernel(float* __restrict__ p, const double* __restrict__ packed, float* __restrict__ m,
                                                          float* __restrict__ v, size_t n, float eta, float b1, float b2, float eps, float c1, float c2) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = (float)(packed[i] / packed[n + 1]);
    const float mi = b1 * m[i] + (1.0f - b1) * g;
    const float vi = b2 * v[i] + (1.0f - b2) * g * g;
    m[i] = mi; v[i] = vi;
    p[i] -= eta * (mi * c1) / (sqrtf(vi * c2) + eps);
}

// ---- device-resident training ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) f32_to_f64_kernel(const float* __restrict__ a, double* __restrict__ b, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (double)a[i];
}
// acc += sum (a - b)^2
__global__ void __launch_bounds__(256) sq_err_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, double* __restrict__ acc) {
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double d = (double)a[i] - (double)b[i];
        s += d * d;
    }
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(acc, s);
}
__global__ void train_losses_kernel(const double* __restrict__ raw, double s0, double s1, double s2, int have_test, double* __restrict__ out) {
    out[0] = raw[0] * s0; out[1] = raw[1] * s1;
    if (have_test) out[2] = raw[2] * s2;
}

// every device copy of the parameters follows the fp32 master d_p, by kernels on the handle's stream (no host round trip)
int params_follow_dev(kanode_handle* h, const float* d_p) {
    float* pf = nullptr; double* pd = nullptr;
    ENSURE(h, W_PARAMS, sizeof(float) * h->np, pf);
    ENSURE(h, W_PARAMS64, sizeof(double) * h->np, pd);
    if (pf != d_p) CK(h, cudaMemcpyAsync(pf, d_p, sizeof(float) * h->np, cudaMemcpyDeviceToDevice, h->stream));
    f32_to_f64_kernel<<<blocks_for((int64_t)h->np, 256), 256, 0, h->stream>>>(d_p, pd, h->np);
    ++h->launches;
    ++h->params_version;                                               // the wide engines rebuild their images from pf / pd lazily
    h->have_params = true; h->params_host_stale = true;
    bool handled = false;
    if (int rc = small_pack_dev(h, d_p, &handled)) return rc;
    CK(h, cudaGetLastError());
    return 0;
}

int train_apply(kanode_handle* h, const float* d_grad, float grad_scale) {
    float *p = (float*)h->ws[kanode_handle::W_TR_P].p, *m = (float*)h->ws[kanode_handle::W_TR_M].p, *v = (float*)h->ws[kanode_handle::W_TR_V].p;
    ++h->tr_t;
    const float c1 = (float)(1.0 / (1.0 - std::pow((double)h->tr_b1, (double)h->tr_t)));
    const float c2 = (float)(1.0 / (1.0 - std::pow((double)h->tr_b2, (double)h->tr_t)));
    adam_update_kernel<<<blocks_for((int64_t)h->np, 256), 256, 0, h->stream>>>(p, d_grad, m, v, h->np, h->tr_eta, h->tr_b1, h->tr_b2, h->tr_eps,
                                                                                c1, c2, grad_scale);
    ++h->launches;
    return params_follow_dev(h, p);
}

// loss = mean(abs2, target - predict) of a forward-only solve, accumulated un-normalised into *d_acc
int forward_loss_dev(kanode_handle* h, const float* d_u0, int64_t B, double t0, double t1, const double* saveat, int nsave,
                     const float* d_target, double abstol, double reltol, double* d_acc) {
    const size_t nout = (size_t)B * nsave * h->n;
    float* d_out = nullptr;
    ENSURE(h, W_TR_OUT, sizeof(float) * (nout ? nout : 1), d_out);
    if (int rc = solve_dev<float>(h, d_u0, B, t0, t1, saveat, nsave, abstol, reltol, d_out, nullptr)) return rc;
    const unsigned nb = (unsigned)std::min<size_t>((nout + 255) / 256, 2368);
    sq_err_kernel<<<nb, 256, 0, h->stream>>>(d_out, d_target, nout, d_acc);
    ++h->launches;
    CK(h, cudaGetLastError());
    return 0;
}

// ---- several devices behind one handle (kanode_create_multi) ---------------------------------------------------------
constexpr int KANODE_MAX_DEVICES = 16;
template <class T> struct PeerPtrs { const T* g[KANODE_MAX_DEVICES]; const double* l[KANODE_MAX_DEVICES]; int n; };
// out_g[i] = scale * sum_k g_k[i], out_l = sum_k l_k: the partial sums are loaded from the peers' memory, in device order
template <class T>
__global__ void __launch_bounds__(256) peer_sum_kernel(const PeerPtrs<T> pp, size_t np, T* __restrict__ out_g, double* __restrict__ out_l, double scale) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { double l = 0.0; for (int k = 0; k < pp.n; ++k) l += *pp.l[k]; *out_l = l; }
    if (i >= np) return;
    double a = 0.0;
    for (int k = 0; k < pp.n; ++k) a += (double)pp.g[k][i];
    out_g[i] = (T)(a * scale);
}

struct Shard { int k; int64_t b0, bk; };
inline std::vector<Shard> make_shards(const kanode_handle* h, int64_t B) {
    const int nd = (int)h->children.size();
    const int64_t per = (B + nd - 1) / nd;
    std::vector<Shard> s;
    for (int k = 0; k < nd; ++k) {
        const int64_t b0 = std::min<int64_t>(B, (int64_t)k * per), bk = std::min<int64_t>(B, b0 + per) - b0;
        if (bk > 0) s.push_back(Shard{k, b0, bk});
    }
    return s;
}
// one host thread per device runs f(child, shard); a failed solve (KANODE_ERR_SOLVER) does not stop the others
template <class F> int multi_fanout(kanode_handle* h, const std::vector<Shard>& shards, F&& f) {
    std::vector<int> rcs(shards.size(), 0);
    std::vector<std::thread> th;
    for (size_t j = 0; j < shards.size(); ++j) th.emplace_back([&, j] { rcs[j] = f(h->children[shards[j].k], shards[j]); });
    for (auto& t : th) t.join();
    int soft = 0;
    for (size_t j = 0; j < shards.size(); ++j) {
        if (rcs[j] == KANODE_ERR_SOLVER) { soft = KANODE_ERR_SOLVER; h->err = h->children[shards[j].k]->err; continue; }
        if (rcs[j] != 0) return fail(h, rcs[j], "device %d: %s", h->children[shards[j].k]->device, h->children[shards[j].k]->err.c_str());
    }
    return soft;
}

template <class T>
int multi_solve(kanode_handle* h, const T* u0, int64_t B, double t0, double t1, const double* saveat, int nsave, double abstol,
                double reltol, T* out, kanode_stats* stats) {
    if (B <= 0) return 0;
    const size_t n = h->n;
    return multi_fanout(h, make_shards(h, B), [&](kanode_handle* c, const Shard& s) {
        return solve_host<T>(c, u0 + s.b0 * n, s.bk, t0, t1, saveat, nsave, abstol, reltol, out ? out + (size_t)s.b0 * nsave * n : nullptr,
                             stats ? stats + s.b0 : nullptr);
    });
}

template <class T>
int multi_loss_grad(kanode_handle* h, const T* u0, int64_t B, double t0, double t1, const double* saveat, int nsave, const T* target,
                    double abstol, double reltol, T* loss, T* grad, T* du0, kanode_stats* fst, kanode_stats* bst, T* out, const T* cot) {
    if (B <= 0 || !u0 || (!target && !cot) || (!loss && !cot) || !grad) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    const size_t n = h->n, per_traj = (size_t)nsave * n;
    const std::vector<Shard> shards = make_shards(h, B);
    const int soft = multi_fanout(h, shards, [&](kanode_handle* c, const Shard& s) {
        return loss_grad_host<T>(c, u0 + s.b0 * n, s.bk, t0, t1, saveat, nsave, target ? target + s.b0 * per_traj : nullptr, abstol, reltol,
                                 nullptr, nullptr, du0 ? du0 + s.b0 * n : nullptr, fst ? fst + s.b0 : nullptr, bst ? bst + s.b0 : nullptr,
                                 nullptr, nullptr, 0, out ? out + s.b0 * per_traj : nullptr, cot ? cot + s.b0 * per_traj : nullptr, true);
    });
    if (soft != 0 && soft != KANODE_ERR_SOLVER) return soft;
    // the only cross-device step: sum of the per-device gradient / loss sums on the first device
    kanode_handle* c0 = h->children[shards[0].k];
    if (int rc = enter(c0)) return rc;
    T* d_g = nullptr; double* d_l = nullptr;
    ENSURE(c0, W_MULTI_G, sizeof(T) * h->np, d_g); ENSURE(c0, W_MULTI_L, sizeof(double), d_l);
    PeerPtrs<T> pp{}; pp.n = (int)shards.size();
    T* stage = nullptr;
    if (!h->peer_ok && shards.size() > 1) ENSURE(c0, W_MULTI_STAGE, (sizeof(T) * h->np + 8) * shards.size(), stage);
    for (size_t j = 0; j < shards.size(); ++j) {
        kanode_handle* c = h->children[shards[j].k];
        const T* g = (const T*)c->ws[kanode_handle::W_GRAD].p; const double* l = (const double*)c->ws[kanode_handle::W_LOSS].p;
        if (c != c0 && !h->peer_ok) {                                  // no peer loads: stage the partial sums on the first device
            char* base = reinterpret_cast<char*>(stage) + j * (sizeof(T) * h->np + 8);
            CK(c0, cudaMemcpyPeerAsync(base + 8, c0->device, g, c->device, sizeof(T) * h->np, c0->stream));
            CK(c0, cudaMemcpyPeerAsync(base, c0->device, l, c->device, sizeof(double), c0->stream));
            g = reinterpret_cast<const T*>(base + 8); l = reinterpret_cast<const double*>(base);
        }
        pp.g[j] = g; pp.l[j] = l;
    }
    peer_sum_kernel<T><<<blocks_for((int64_t)h->np, 256), 256, 0, c0->stream>>>(pp, h->np, d_g, d_l, cot ? 1.0 : 1.0 / (double)B);
    ++c0->launches;
    CK(c0, cudaGetLastError());
    double lsum = 0.0;
    CK(c0, cudaMemcpyAsync(grad, d_g, sizeof(T) * h->np, cudaMemcpyDeviceToHost, c0->stream));
    CK(c0, cudaMemcpyAsync(&lsum, d_l, sizeof(double), cudaMemcpyDeviceToHost, c0->stream));
    CK(c0, cudaStreamSynchronize(c0->stream));
    if (loss) *loss = (T)(lsum / ((double)B * nsave * n));
    return soft;
}

template <class T> int set_params_any(kanode_handle* h, const T* p, size_t np) {
    if (h && !h->children.empty()) {
        for (kanode_handle* c : h->children) if (int rc = set_params_host<T>(c, p, np)) return fail(h, rc, "%s", c->err.c_str());
        h->have_params = true;
        return 0;
    }
    return set_params_host<T>(h, p, np);
}
template <class T> int edge_activations_host(kanode_handle* h, int layer, const T* x, T* act, int64_t K) {
    if (int rc = enter(h)) return rc;
    if (!h->have_params) return fail(h, KANODE_ERR_INVALID, "parameters not set");
    if (layer < 0 || layer >= h->desc.n_layers || K < 0 || (K > 0 && (!x || !act))) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    if (K == 0) return 0;
    if (int rc = generic_supported(h)) return rc;
    const kanode_layer_desc& L = h->desc.layers[layer];
    const size_t nin = (size_t)K * L.in_dims, nact = nin * L.out_dims;
    T *d_x = nullptr, *d_a = nullptr;
    ENSURE(h, W_U0, sizeof(T) * nin, d_x); ENSURE(h, W_ACT, sizeof(T) * nact, d_a);
    CK(h, cudaMemcpyAsync(d_x, x, sizeof(T) * nin, cudaMemcpyHostToDevice, h->stream));
    edge_activation_kernel<T><<<blocks_for((int64_t)nin, 128), 128, 0, h->stream>>>(h->gm, layer, generic_params<T>(h), d_x, d_a, K);
    ++h->launches;
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(act, d_a, sizeof(T) * nact, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

}  // namespace

// =============================================================================================================
// C ABI
// =============================================================================================================
extern "C" {

const char* kanode_version(void) { return "kanode-b200 0.1.0 (sm_100a)"; }

const char* kanode_last_error(const kanode_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

size_t kanode_param_count(const kanode_desc* desc) { return count_params(desc); }

int kanode_create(const kanode_desc* desc, int device, void* stream, kanode_handle** out) {
    if (!out) return fail(nullptr, KANODE_ERR_INVALID, "null out pointer");
    *out = nullptr;
    const size_t np = count_params(desc);
    if (np == 0) return fail(nullptr, KANODE_ERR_INVALID, "invalid descriptor");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        (void)cudaGetLastError();
        return fail(nullptr, KANODE_ERR_NO_DEVICE, "no CUDA device visible; the KAN-ODE hot path has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, KANODE_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, KANODE_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, KANODE_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    kanode_handle* h = new (std::nothrow) kanode_handle();
    if (!h) return fail(nullptr, KANODE_ERR_NOMEM, "out of host memory");
    h->desc = *desc; h->device = device; h->np = np; h->n = desc->n_state;
    h->n_out = desc->rhs_kind == KANODE_RHS_MAP ? desc->layers[desc->n_layers - 1].out_dims : desc->n_state;
    h->params.assign(np, 0.0);
    h->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    if (const char* e = std::getenv("KANODE_LG_PERSIST")) h->lg_persist = std::atoi(e);
    if (const char* e = std::getenv("KANODE_OVERLAP_H2D")) h->overlap_h2d = std::atoi(e);
    if (const char* e = std::getenv("KANODE_LG_SHAPE")) h->lg_shape = std::atoi(e);
    if (const char* e = std::getenv("KANODE_BWD_MAXIT")) h->bwd_maxiters = std::atoi(e);     // timing experiments only
    if (const char* e = std::getenv("KANODE_SCHEDULE")) h->schedule = std::atoi(e);
    if (const char* e = std::getenv("KANODE_WIDE")) h->wide = std::atoi(e);
    if (const char* e = std::getenv("KANODE_WIDE_TC")) h->wide_tc = std::atoi(e);
    if (const char* e = std::getenv("KANODE_WIDE_GRAPH")) h->wide_graph = std::atoi(e);
    if (const char* e = std::getenv("KANODE_WIDE_GRAPH_MAXN")) h->wide_graph_maxn = std::atoi(e);
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return fail(nullptr, KANODE_ERR_CUDA, "cudaSetDevice failed"); }
    if (stream) { h->stream = (cudaStream_t)stream; h->own_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return fail(nullptr, KANODE_ERR_CUDA, "cudaStreamCreate failed"); }
        h->own_stream = true;
    }
    if (cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->aux_ev[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->aux_ev[1], cudaEventDisableTiming) != cudaSuccess) {
        kanode_destroy(h); return fail(nullptr, KANODE_ERR_CUDA, "aux stream/event creation failed");
    }
    for (auto& e : h->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { kanode_destroy(h); return fail(nullptr, KANODE_ERR_CUDA, "cudaEventCreate failed"); }
    if (int rc = generic_init(h)) { g_create_error = h->err; kanode_destroy(h); return rc; }
    *out = h;
    return 0;
}

int kanode_create_multi(const kanode_desc* desc, const int32_t* devices, int32_t n_devices, kanode_handle** out) {
    if (!out) return fail(nullptr, KANODE_ERR_INVALID, "null out pointer");
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > KANODE_MAX_DEVICES) return fail(nullptr, KANODE_ERR_INVALID, "1..%d devices", KANODE_MAX_DEVICES);
    for (int a = 0; a < n_devices; ++a)
        for (int b = a + 1; b < n_devices; ++b)
            if (devices[a] == devices[b]) return fail(nullptr, KANODE_ERR_INVALID, "device %d listed twice", devices[a]);
    const size_t np = count_params(desc);
    if (np == 0) return fail(nullptr, KANODE_ERR_INVALID, "invalid descriptor");
    kanode_handle* h = new (std::nothrow) kanode_handle();
    if (!h) return fail(nullptr, KANODE_ERR_NOMEM, "out of host memory");
    h->desc = *desc; h->np = np; h->n = desc->n_state;
    h->n_out = desc->rhs_kind == KANODE_RHS_MAP ? desc->layers[desc->n_layers - 1].out_dims : desc->n_state;
    for (int k = 0; k < n_devices; ++k) {
        kanode_handle* c = nullptr;
        if (int rc = kanode_create(desc, devices[k], nullptr, &c)) { kanode_destroy(h); return rc; }   // the error text is already set
        h->children.push_back(c);
    }
    h->device = h->children[0]->device;
    // peer access from the first device to the others: the gradient combine loads their partial sums in place
    h->peer_ok = true;
    if (cudaSetDevice(h->device) != cudaSuccess) { kanode_destroy(h); return fail(nullptr, KANODE_ERR_CUDA, "cudaSetDevice failed"); }
    for (int k = 1; k < n_devices; ++k) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, h->device, devices[k]) != cudaSuccess || !can) { h->peer_ok = false; continue; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(devices[k], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) h->peer_ok = false;
        (void)cudaGetLastError();
    }
    *out = h;
    return 0;
}

int32_t kanode_device_count(const kanode_handle* h) { return h ? (h->children.empty() ? 1 : (int32_t)h->children.size()) : 0; }

int kanode_destroy(kanode_handle* h) {
    if (!h) return 0;
    if (!h->children.empty()) {
        for (kanode_handle* c : h->children) kanode_destroy(c);
        delete h;
        return 0;
    }
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    kanode::peer_release(h);
    for (auto& b : h->ws) if (b.p) cudaFree(b.p);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    for (auto& e : h->aux_ev) if (e) cudaEventDestroy(e);
    if (h->order_ev) cudaEventDestroy(h->order_ev);
    for (auto& e : h->wide_gp_ev) cudaEventDestroy(e);
    if (h->stage) cudaFreeHost(h->stage);
    for (auto& g : h->wide_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (h->aux_stream) { cudaStreamSynchronize(h->aux_stream); cudaStreamDestroy(h->aux_stream); }
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

#define KANODE_SINGLE_ONLY(h, what)                                                                                   \
    if ((h) && !(h)->children.empty()) return fail(h, KANODE_ERR_UNSUPPORTED, what " takes device pointers of one GPU: not available on a multi-device handle")

int kanode_sync(kanode_handle* h) {
    if (h && !h->children.empty()) { for (kanode_handle* c : h->children) if (int rc = kanode_sync(c)) return fail(h, rc, "%s", c->err.c_str()); return 0; }
    if (int rc = enter(h)) return rc;
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int kanode_set_record_capacity(kanode_handle* h, int32_t max_steps) {
    if (!h || max_steps < 1) return fail(h, KANODE_ERR_INVALID, "bad capacity");
    for (kanode_handle* c : h->children) c->rec_cap = max_steps;
    h->rec_cap = max_steps;
    return 0;
}

int64_t kanode_launch_count(const kanode_handle* h) {
    if (!h) return 0;
    int64_t n = h->launches;
    for (const kanode_handle* c : h->children) n += c->launches;
    return n;
}

int kanode_adam_step_dev(kanode_handle* h, float* d_p, const float* d_grad, float* d_m, float* d_v, int64_t t, float eta,
                         float beta1, float beta2, float eps, float grad_scale) {
    KANODE_SINGLE_ONLY(h, "kanode_adam_step_dev");
    if (int rc = enter(h)) return rc;
    if (!d_p || !d_grad || !d_m || !d_v || t < 1) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    const float c1 = (float)(1.0 / (1.0 - std::pow((double)beta1, (double)t)));
    const float c2 = (float)(1.0 / (1.0 - std::pow((double)beta2, (double)t)));
    adam_update_kernel<<<blocks_for((int64_t)h->np, 256), 256, 0, h->stream>>>(d_p, d_grad, d_m, d_v, h->np, eta, beta1, beta2, eps,
                                                                                c1, c2, grad_scale);
    ++h->launches;
    CK(h, cudaGetLastError());
    return 0;
}

int kanode_last_timing(kanode_handle* h, float* ms3) {
    if (h && !h->children.empty()) return kanode_last_timing(h->children[0], ms3);
    if (int rc = enter(h)) return rc;
    if (!ms3 || !h->ev_valid) return fail(h, KANODE_ERR_INVALID, "no timed loss_grad call yet");
    CK(h, cudaEventSynchronize(h->ev[3]));
    for (int i = 0; i < 3; ++i) CK(h, cudaEventElapsedTime(&ms3[i], h->ev[i], h->ev[i + 1]));
    return 0;
}

int kanode_last_gpass_timing(kanode_handle* h, float* ms, int32_t* passes) {
    if (h && !h->children.empty()) return kanode_last_gpass_timing(h->children[0], ms, passes);
    if (int rc = enter(h)) return rc;
    if (!ms || !passes || !h->ev_valid || h->wide_gp_used < 2) return fail(h, KANODE_ERR_INVALID, "no wide loss_grad call yet");
    CK(h, cudaEventSynchronize(h->ev[3]));
    float tot = 0.f;
    for (int i = 0; i + 1 < h->wide_gp_used; i += 2) { float t = 0.f; CK(h, cudaEventElapsedTime(&t, h->wide_gp_ev[i], h->wide_gp_ev[i + 1])); tot += t; }
    *ms = tot; *passes = h->wide_gp_used / 2;
    return 0;
}

int kanode_set_params(kanode_handle* h, const float* p, size_t np) { return set_params_any<float>(h, p, np); }
int kanode_set_params_f64(kanode_handle* h, const double* p, size_t np) { return set_params_any<double>(h, p, np); }
int kanode_set_params_dev(kanode_handle* h, const float* d_p, size_t np) {
    KANODE_SINGLE_ONLY(h, "kanode_set_params_dev");
    if (int rc = enter(h)) return rc;
    if (!d_p || np != h->np) return fail(h, KANODE_ERR_INVALID, "expected %zu parameters, got %zu", h->np, np);
    // device copies follow by kernels; the host copy is fetched lazily by the few entry points that read it
    float* master = nullptr;
    ENSURE(h, W_TR_P, sizeof(float) * np, master);
    if (master != d_p) CK(h, cudaMemcpyAsync(master, d_p, sizeof(float) * np, cudaMemcpyDeviceToDevice, h->stream));
    return params_follow_dev(h, master);
}

static kanode_handle* first_dev(kanode_handle* h) { return h && !h->children.empty() ? h->children[0] : h; }
static int lift(kanode_handle* h, int rc) { if (rc && h && !h->children.empty()) h->err = h->children[0]->err; return rc; }
int kanode_rhs(kanode_handle* h, const float* u, float* du, int64_t batch) { return lift(h, rhs_host<float>(first_dev(h), u, du, batch)); }
int kanode_rhs_f64(kanode_handle* h, const double* u, double* du, int64_t batch) { return lift(h, rhs_host<double>(first_dev(h), u, du, batch)); }
int kanode_rhs_dev(kanode_handle* h, const float* d_u, float* d_du, int64_t batch) {
    KANODE_SINGLE_ONLY(h, "kanode_rhs_dev");
    if (int rc = enter(h)) return rc;
    return rhs_dev<float>(h, d_u, d_du, batch);
}

int kanode_vjp(kanode_handle* h, const float* u, const float* lam, float* ubar, float* pbar, int64_t batch) {
    return lift(h, vjp_host<float>(first_dev(h), u, lam, ubar, pbar, batch));
}
int kanode_vjp_f64(kanode_handle* h, const double* u, const double* lam, double* ubar, double* pbar, int64_t batch) {
    return lift(h, vjp_host<double>(first_dev(h), u, lam, ubar, pbar, batch));
}

int kanode_solve(kanode_handle* h, const float* u0, int64_t batch, double t0, double t1, const double* saveat,
                 int32_t nsave, float abstol, float reltol, float* out, kanode_stats* stats) {
    if (h && !h->children.empty()) return multi_solve<float>(h, u0, batch, t0, t1, saveat, nsave, abstol, reltol, out, stats);
    return solve_host<float>(h, u0, batch, t0, t1, saveat, nsave, abstol, reltol, out, stats);
}
int kanode_solve_f64(kanode_handle* h, const double* u0, int64_t batch, double t0, double t1, const double* saveat,
                     int32_t nsave, double abstol, double reltol, double* out, kanode_stats* stats) {
    if (h && !h->children.empty()) return multi_solve<double>(h, u0, batch, t0, t1, saveat, nsave, abstol, reltol, out, stats);
    return solve_host<double>(h, u0, batch, t0, t1, saveat, nsave, abstol, reltol, out, stats);
}
int kanode_solve_dev(kanode_handle* h, const float* d_u0, int64_t batch, double t0, double t1, const double* saveat,
                     int32_t nsave, float abstol, float reltol, float* d_out, kanode_stats* d_stats) {
    KANODE_SINGLE_ONLY(h, "kanode_solve_dev");
    if (int rc = enter(h)) return rc;
    return solve_dev<float>(h, d_u0, batch, t0, t1, saveat, nsave, abstol, reltol, d_out, d_stats);
}

int kanode_loss_grad(kanode_handle* h, const float* u0, int64_t batch, double t0, double t1, const double* saveat,
                     int32_t nsave, const float* target, float abstol, float reltol, float* loss, float* grad,
                     float* du0, kanode_stats* fwd_stats, kanode_stats* bwd_stats) {
    if (h && !h->children.empty())
        return multi_loss_grad<float>(h, u0, batch, t0, t1, saveat, nsave, target, abstol, reltol, loss, grad, du0, fwd_stats, bwd_stats, nullptr, nullptr);
    return loss_grad_host<float>(h, u0, batch, t0, t1, saveat, nsave, target, abstol, reltol, loss, grad, du0,
                                 fwd_stats, bwd_stats);
}
int kanode_loss_grad_f64(kanode_handle* h, const double* u0, int64_t batch, double t0, double t1, const double* saveat,
                         int32_t nsave, const double* target, double abstol, double reltol, double* loss, double* grad,
                         double* du0, kanode_stats* fwd_stats, kanode_stats* bwd_stats) {
    if (h && !h->children.empty())
        return multi_loss_grad<double>(h, u0, batch, t0, t1, saveat, nsave, target, abstol, reltol, loss, grad, du0, fwd_stats, bwd_stats, nullptr, nullptr);
    return loss_grad_host<double>(h, u0, batch, t0, t1, saveat, nsave, target, abstol, reltol, loss, grad, du0,
                                  fwd_stats, bwd_stats);
}
int kanode_loss_grad_replay(kanode_handle* h, const float* u0, int64_t batch, double t0, double t1, const double* saveat,
                            int32_t nsave, const float* target, float abstol, float reltol, const double* fwd_t,
                            const double* bwd_t, int32_t max_steps, float* loss, float* grad, float* du0, float* out,
                            kanode_stats* fwd_stats, kanode_stats* bwd_stats) {
    KANODE_SINGLE_ONLY(h, "kanode_loss_grad_replay");
    if (!fwd_t || !bwd_t) return fail(h, KANODE_ERR_INVALID, "replay needs both step sequences");
    return loss_grad_host<float>(h, u0, batch, t0, t1, saveat, nsave, target, abstol, reltol, loss, grad, du0,
                                 fwd_stats, bwd_stats, fwd_t, bwd_t, max_steps, out);
}
int kanode_loss_grad_replay_f64(kanode_handle* h, const double* u0, int64_t batch, double t0, double t1, const double* saveat,
                                int32_t nsave, const double* target, double abstol, double reltol, const double* fwd_t,
                                const double* bwd_t, int32_t max_steps, double* loss, double* grad, double* du0, double* out,
                                kanode_stats* fwd_stats, kanode_stats* bwd_stats) {
    KANODE_SINGLE_ONLY(h, "kanode_loss_grad_replay_f64");
    if (!fwd_t || !bwd_t) return fail(h, KANODE_ERR_INVALID, "replay needs both step sequences");
    return loss_grad_host<double>(h, u0, batch, t0, t1, saveat, nsave, target, abstol, reltol, loss, grad, du0,
                                  fwd_stats, bwd_stats, fwd_t, bwd_t, max_steps, out);
}
int kanode_solve_adjoint(kanode_handle* h, const float* u0, int64_t batch, double t0, double t1, const double* saveat, int32_t nsave,
                         float abstol, float reltol, const float* dL_dout, float* out, float* grad, float* du0,
                         kanode_stats* fwd_stats, kanode_stats* bwd_stats) {
    if (!dL_dout) return fail(h, KANODE_ERR_INVALID, "null cotangent");
    if (h && !h->children.empty())
        return multi_loss_grad<float>(h, u0, batch, t0, t1, saveat, nsave, nullptr, abstol, reltol, nullptr, grad, du0, fwd_stats, bwd_stats, out, dL_dout);
    return loss_grad_host<float>(h, u0, batch, t0, t1, saveat, nsave, nullptr, abstol, reltol, nullptr, grad, du0, fwd_stats, bwd_stats,
                                 nullptr, nullptr, 0, out, dL_dout);
}
int kanode_solve_adjoint_f64(kanode_handle* h, const double* u0, int64_t batch, double t0, double t1, const double* saveat,
                             int32_t nsave, double abstol, double reltol, const double* dL_dout, double* out, double* grad,
                             double* du0, kanode_stats* fwd_stats, kanode_stats* bwd_stats) {
    if (!dL_dout) return fail(h, KANODE_ERR_INVALID, "null cotangent");
    if (h && !h->children.empty())
        return multi_loss_grad<double>(h, u0, batch, t0, t1, saveat, nsave, nullptr, abstol, reltol, nullptr, grad, du0, fwd_stats, bwd_stats, out, dL_dout);
    return loss_grad_host<double>(h, u0, batch, t0, t1, saveat, nsave, nullptr, abstol, reltol, nullptr, grad, du0, fwd_stats, bwd_stats,
                                  nullptr, nullptr, 0, out, dL_dout);
}
int kanode_solve_adjoint_dev(kanode_handle* h, const float* d_u0, int64_t batch, double t0, double t1, const double* saveat,
                             int32_t nsave, float abstol, float reltol, const float* d_dL_dout, float* d_out, float* d_grad,
                             float* d_du0, kanode_stats* d_fwd_stats, kanode_stats* d_bwd_stats) {
    KANODE_SINGLE_ONLY(h, "kanode_solve_adjoint_dev");
    if (int rc = enter(h)) return rc;
    if (batch < 0 || !d_dL_dout || !d_grad) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    double* d_loss = nullptr;
    ENSURE(h, W_LOSS, sizeof(double), d_loss);
    return loss_grad_dev<float>(h, d_u0, batch, t0, t1, saveat, nsave, nullptr, abstol, reltol, d_loss, d_grad, d_du0, d_fwd_stats,
                                d_bwd_stats, d_out, nullptr, nullptr, 0, d_dL_dout);
}
int kanode_edge_activations(kanode_handle* h, int32_t layer, const float* x, float* act, int64_t K) {
    return lift(h, edge_activations_host<float>(first_dev(h), layer, x, act, K));
}
int kanode_edge_activations_f64(kanode_handle* h, int32_t layer, const double* x, double* act, int64_t K) {
    return lift(h, edge_activations_host<double>(first_dev(h), layer, x, act, K));
}
int kanode_set_regularizer(kanode_handle* h, double act_reg, double entropy_reg) {
    if (!h || !(act_reg >= 0.0) || !(entropy_reg >= 0.0)) return fail(h, KANODE_ERR_INVALID, "bad regulariser weights");
    h->reg_act = act_reg; h->reg_entropy = entropy_reg;
    for (kanode_handle* c : h->children) { c->reg_act = act_reg; c->reg_entropy = entropy_reg; }
    return 0;
}
int kanode_reg_loss(kanode_handle* h, double act_reg, double entropy_reg, double* loss, float* grad) {
    if (h && !h->children.empty()) return lift(h, kanode_reg_loss(h->children[0], act_reg, entropy_reg, loss, grad));
    if (int rc = enter(h)) return rc;
    if (!h->have_params || !loss) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    double* d_l = nullptr; float* d_g = nullptr;
    ENSURE(h, W_LOSS, sizeof(double), d_l); ENSURE(h, W_GRAD, sizeof(float) * h->np, d_g);
    CK(h, cudaMemsetAsync(d_l, 0, sizeof(double), h->stream));
    CK(h, cudaMemsetAsync(d_g, 0, sizeof(float) * h->np, h->stream));
    const double a0 = h->reg_act, e0 = h->reg_entropy;
    h->reg_act = act_reg; h->reg_entropy = entropy_reg;
    const int rc = reg_apply<float>(h, d_l, 1.0, grad ? d_g : nullptr, 1.0);
    h->reg_act = a0; h->reg_entropy = e0;
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(loss, d_l, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (grad) CK(h, cudaMemcpyAsync(grad, d_g, sizeof(float) * h->np, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}
int kanode_train_begin(kanode_handle* h, float eta, float beta1, float beta2, float eps) {
    KANODE_SINGLE_ONLY(h, "kanode_train_begin");
    if (int rc = enter(h)) return rc;
    if (!h->have_params) return fail(h, KANODE_ERR_INVALID, "parameters not set");
    if (!(eta > 0.f)) return fail(h, KANODE_ERR_INVALID, "bad learning rate");
    if (int rc = host_params_refresh(h)) return rc;
    float *p = nullptr, *m = nullptr, *v = nullptr, *g = nullptr; double* raw = nullptr;
    ENSURE(h, W_TR_P, sizeof(float) * h->np, p); ENSURE(h, W_TR_M, sizeof(float) * h->np, m);
    ENSURE(h, W_TR_V, sizeof(float) * h->np, v); ENSURE(h, W_TR_GRAD, sizeof(float) * h->np, g);
    ENSURE(h, W_TR_RAW, 4 * sizeof(double), raw);
    std::vector<float> tmp(h->np);
    for (size_t i = 0; i < h->np; ++i) tmp[i] = (float)h->params[i];
    CK(h, cudaMemcpyAsync(p, tmp.data(), sizeof(float) * h->np, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemsetAsync(m, 0, sizeof(float) * h->np, h->stream));
    CK(h, cudaMemsetAsync(v, 0, sizeof(float) * h->np, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    h->tr_eta = eta; h->tr_b1 = beta1; h->tr_b2 = beta2; h->tr_eps = eps; h->tr_t = 0; h->train_on = true;
    return 0;
}
int kanode_train_apply_dev(kanode_handle* h, const float* d_grad_sum, float grad_scale) {
    KANODE_SINGLE_ONLY(h, "kanode_train_apply_dev");
    if (int rc = enter(h)) return rc;
    if (!h->train_on || !d_grad_sum) return fail(h, KANODE_ERR_INVALID, "kanode_train_begin first");
    return train_apply(h, d_grad_sum, grad_scale);
}
int kanode_pack_sums_dev(kanode_handle* h, const float* d_grad_sum, const double* d_loss_sum, int64_t count, double* d_packed) {
    KANODE_SINGLE_ONLY(h, "kanode_pack_sums_dev");
    return pack_sums<float>(h, d_grad_sum, d_loss_sum, count, d_packed);
}
int kanode_pack_sums_dev_f64(kanode_handle* h, const double* d_grad_sum, const double* d_loss_sum, int64_t count, double* d_packed) {
    KANODE_SINGLE_ONLY(h, "kanode_pack_sums_dev_f64");
    return pack_sums<double>(h, d_grad_sum, d_loss_sum, count, d_packed);
}
int kanode_train_apply_packed_dev(kanode_handle* h, const double* d_packed) {
    KANODE_SINGLE_ONLY(h, "kanode_train_apply_packed_dev");
    if (int rc = enter(h)) return rc;
    if (!h->train_on || !d_packed) return fail(h, KANODE_ERR_INVALID, "kanode_train_begin first");
    float *p = (float*)h->ws[kanode_handle::W_TR_P].p, *m = (float*)h->ws[kanode_handle::W_TR_M].p, *v = (float*)h->ws[kanode_handle::W_TR_V].p;
    ++h->tr_t;
    const float c1 = (float)(1.0 / (1.0 - std::pow((double)h->tr_b1, (double)h->tr_t)));
    const float c2 = (float)(1.0 / (1.0 - std::pow((double)h->tr_b2, (double)h->tr_t)));
    adam_packed_kernel<<<blocks_for((int64_t)h->np, 256), 256, 0, h->stream>>>(p, d_packed, m, v, h->np, h->tr_eta, h->tr_b1, h->tr_b2, h->tr_eps, c1, c2);
    ++h->launches;
    return params_follow_dev(h, p);
}
int kanode_train_step_dev(kanode_handle* h, const float* d_u0, int64_t batch, double t0, double t1, const double* saveat,
                          int32_t nsave, const float* d_target, float abstol, float reltol, const float* d_u0_test,
                          int64_t batch_test, double t1_test, const double* saveat_test, int32_t nsave_test,
                          const float* d_target_test, double* d_losses) {
    KANODE_SINGLE_ONLY(h, "kanode_train_step_dev");
    if (int rc = enter(h)) return rc;
    if (!h->train_on) return fail(h, KANODE_ERR_INVALID, "kanode_train_begin first");
    if (batch <= 0 || !d_u0 || !d_target || !d_losses) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    const bool test = d_target_test && d_u0_test && batch_test > 0 && nsave_test > 0;
    double* raw = (double*)h->ws[kanode_handle::W_TR_RAW].p;
    float* g = (float*)h->ws[kanode_handle::W_TR_GRAD].p;
    CK(h, cudaMemsetAsync(raw, 0, 4 * sizeof(double), h->stream));
    if (int rc = loss_grad_dev<float>(h, d_u0, batch, t0, t1, saveat, nsave, d_target, abstol, reltol, &raw[0], g, nullptr, nullptr, nullptr,
                                      (float*)nullptr)) return rc;
    if (int rc = train_apply(h, g, 1.0f / (float)batch)) return rc;
    if (int rc = forward_loss_dev(h, d_u0, batch, t0, t1, saveat, nsave, d_target, abstol, reltol, &raw[1])) return rc;
    if (test)
        if (int rc = forward_loss_dev(h, d_u0_test, batch_test, t0, t1_test, saveat_test, nsave_test, d_target_test, abstol, reltol, &raw[2])) return rc;
    const double s = 1.0 / ((double)batch * nsave * h->n), st = test ? 1.0 / ((double)batch_test * nsave_test * h->n) : 0.0;
    train_losses_kernel<<<1, 1, 0, h->stream>>>(raw, s, s, st, test ? 1 : 0, d_losses);
    ++h->launches;
    CK(h, cudaGetLastError());
    return 0;
}
int kanode_train_params(kanode_handle* h, float* p) {
    KANODE_SINGLE_ONLY(h, "kanode_train_params");
    if (int rc = enter(h)) return rc;
    if (!h->train_on || !p) return fail(h, KANODE_ERR_INVALID, "kanode_train_begin first");
    CK(h, cudaMemcpyAsync(p, h->ws[kanode_handle::W_TR_P].p, sizeof(float) * h->np, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < h->np; ++i) h->params[i] = (double)p[i];
    h->params_host_stale = false;
    return 0;
}
int kanode_loss_grad_dev(kanode_handle* h, const float* d_u0, int64_t batch, double t0, double t1, const double* saveat,
                         int32_t nsave, const float* d_target, float abstol, float reltol, double* d_loss_sum,
                         float* d_grad_sum, float* d_du0, kanode_stats* d_fwd_stats, kanode_stats* d_bwd_stats) {
    KANODE_SINGLE_ONLY(h, "kanode_loss_grad_dev");
    if (int rc = enter(h)) return rc;
    if (batch < 0 || !d_loss_sum || !d_grad_sum) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    return loss_grad_dev<float>(h, d_u0, batch, t0, t1, saveat, nsave, d_target, abstol, reltol, d_loss_sum,
                                d_grad_sum, d_du0, d_fwd_stats, d_bwd_stats, (float*)nullptr);
}
int kanode_loss_grad_dev_f64(kanode_handle* h, const double* d_u0, int64_t batch, double t0, double t1,
                             const double* saveat, int32_t nsave, const double* d_target, double abstol, double reltol,
                             double* d_loss_sum, double* d_grad_sum, double* d_du0, kanode_stats* d_fwd_stats,
                             kanode_stats* d_bwd_stats) {
    KANODE_SINGLE_ONLY(h, "kanode_loss_grad_dev_f64");
    if (int rc = enter(h)) return rc;
    if (batch < 0 || !d_loss_sum || !d_grad_sum) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    return loss_grad_dev<double>(h, d_u0, batch, t0, t1, saveat, nsave, d_target, abstol, reltol, d_loss_sum,
                                 d_grad_sum, d_du0, d_fwd_stats, d_bwd_stats, (double*)nullptr);
}

}  // extern "C"
