// kanode_wsrc.cuh — batched lockstep engine for the HIDDEN-SOURCE model (BASELINE config 4a):
//     du/dt = s*D * lap(u) + kan(u)      periodic 3-point Laplacian + pointwise 1 -> 1 KDense (np = G + 1 parameters)
// ("PDE examples/Allen-Cahn_Source.jl":50-54,90-99, Fisher-KPP_Source.jl:55-59,95-104).
//
// Same lockstep scheme, controller, accept / error / norm kernels as the surrogate engine (kanode_wide.cuh); only the
// right-hand side differs: one ELEMENTWISE kernel per stage over [B][n] (stencil + pointwise KAN with its G+1 weights in
// registers, neighbours of the block's first/last node recomputed instead of exchanged), and in the adjoint one elementwise
// kernel that forms y = sol(t_s), lambda_s, dlambda/dt and the block partials of the G+1 parameter-gradient sums.
// dg/dt is so small (11 numbers per IC) that the step-end pass lives in the control kernel.
#pragma once
#include "kanode_wide.cuh"

namespace kanode {

struct SrcModel {
    int n, norm;
    float inv_h;
    float grid[16];
    double lap_scale;      // lap_coef / dx^2
    long long offC, offW;
};

template <class T> __device__ __forceinline__ T wsrc_comb(const WideIn<T>& in, int b, int i, int n, int64_t B) {
    const int64_t e = (int64_t)b * n + i;
    T acc = T(0);
    for (int j = 0; j < in.ncoef; ++j) acc += in.coef[j] * in.ks[(int64_t)j * B * n + e];
    return in.ncoef > 0 ? in.base[e] + in.hs[b] * acc : in.base[e];
}
// k[b][i] = ls * (x[i-1] - 2 x[i] + x[i+1]) + kan(x[i]),  x = uprev + h * sum a_sj k_j
template <class T, int G>
__global__ void __launch_bounds__(W_ET) wsrc_rhs_kernel(const __grid_constant__ SrcModel m, const T* __restrict__ p, const WideIn<T> in, int64_t B,
                                                        T* out, T* xstore) {
    __shared__ T xs[W_ET];
    const int b = blockIdx.y, tid = threadIdx.x, n = m.n, i = blockIdx.x * W_ET + tid;
    if (in.mask && !in.mask[b]) return;
    const bool valid = i < n;
    const T x = valid ? wsrc_comb<T>(in, b, i, n, B) : T(0);
    xs[tid] = x;
    __syncthreads();
    if (!valid) return;
    const T xl = tid > 0 ? xs[tid - 1] : wsrc_comb<T>(in, b, (i + n - 1) % n, n, B);          // periodic corners AC_Source:53-54
    const T xr = (tid + 1 < W_ET && i + 1 < n) ? xs[tid + 1] : wsrc_comb<T>(in, b, (i + 1) % n, n, B);
    T c[G + 1];
    w_features<T, G>(m.norm, (T)m.inv_h, m.grid, x, c);
    T kan = T(0);
#pragma unroll
    for (int g = 0; g < G; ++g) kan += p[m.offC + g] * c[g];
    kan += p[m.offW] * c[G];
    out[(int64_t)b * n + i] = (T)m.lap_scale * (xl - T(2) * x + xr) + kan;                    // AC_Source:92
    if (xstore) xstore[(int64_t)b * n + i] = x;
}

// adjoint stage: kl[b][i] = -( ls * (l[i-1] - 2 l[i] + l[i+1]) + dkan(y_i) * l[i] ),  kg partials [chunk][q] = sum_i l_i * c_q(y_i)
template <class T, int G>
__global__ void __launch_bounds__(W_ET) wsrc_vjp_kernel(const __grid_constant__ SrcModel m, const T* __restrict__ p, const WideIn<T> iny /* dense record */,
                                                        const WideIn<T> inl /* lambda combination */, int64_t B, T* dl, T* lstore, T* kg_part /* [B][nkg][W_HP] of this stage */) {
    __shared__ T ls_[W_ET];
    __shared__ T red[W_ET / 32][W_HP];
    __shared__ T sbw[8];                       // interpolation weights b_1..b_7(theta) and the step of this IC's record row
    __shared__ const T* srow;
    const int b = blockIdx.y, tid = threadIdx.x, n = m.n, i = blockIdx.x * W_ET + tid;
    if (inl.mask && !inl.mask[b]) return;
    if (tid == 0) {
        T bw[7]; interp_weights(iny.th[b], bw);
#pragma unroll
        for (int j = 0; j < 7; ++j) sbw[j] = bw[j];
        sbw[7] = iny.hd[b];
        srow = iny.rec + ((int64_t)b * iny.cap + iny.ridx[b]) * 8 * (int64_t)n;
    }
    const bool valid = i < n;
    const T l = valid ? wsrc_comb<T>(inl, b, i, n, B) : T(0);
    ls_[tid] = l;
    __syncthreads();
    T acc[G + 1];
#pragma unroll
    for (int q = 0; q <= G; ++q) acc[q] = T(0);
    if (valid) {
        const T ll = tid > 0 ? ls_[tid - 1] : wsrc_comb<T>(inl, b, (i + n - 1) % n, n, B);     // the Laplacian is symmetric
        const T lr = (tid + 1 < W_ET && i + 1 < n) ? ls_[tid + 1] : wsrc_comb<T>(inl, b, (i + 1) % n, n, B);
        T yacc = T(0);
#pragma unroll
        for (int j = 0; j < 7; ++j) yacc += sbw[j] * srow[(int64_t)(1 + j) * n + i];
        const T y = srow[i] + sbw[7] * yacc;
        // forward features and their derivatives (utils.jl:15-21, NNlib activation rules)
        const T inv_h = (T)m.inv_h;
        const T xn = normalize_rt(m.norm, y);
        T xnbar = T(0);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const T a = (xn - (T)m.grid[g]) * inv_h;
            const T bb = kexp(-a * a);
            const T db = T(-2) * a * bb;
            acc[g] = l * bb;
            xnbar += db * inv_h * (p[m.offC + g] * l);
        }
        T xb = xnbar * normalize_deriv_rt(m.norm, xn);
        T sw, ds; swish_both(y, sw, ds);
        acc[G] = l * sw;
        xb += (p[m.offW] * l) * ds;
        dl[(int64_t)b * n + i] = -((T)m.lap_scale * (ll - T(2) * l + lr) + xb);
        if (lstore) lstore[(int64_t)b * n + i] = l;
    }
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int q = 0; q <= G; ++q) { const T v = warp_sum(acc[q]); if (lane == 0) red[warp][q] = v; }
    __syncthreads();
    if (tid <= G) {
        T sacc = T(0);
        for (int w = 0; w < W_ET / 32; ++w) sacc += red[w][tid];
        kg_part[(((int64_t)b * gridDim.x) + blockIdx.x) * W_HP + tid] = sacc;
    }
}

// accepted attempt: lambda <- lambda_new (+ jumps); FSAL shift of slot 6 -> slot 0 (dlambda/dt and the kg partials)
template <class T>
__global__ void __launch_bounds__(W_ET) wsrc_bwd_accept_kernel(const WideCtl c, const WideBwd<T> a, const T* lnew, T* kg_part, int n, int64_t B) {
    const int b = blockIdx.y, i = blockIdx.x * W_ET + threadIdx.x;
    if (!c.acc_now[b]) return;
    const bool modified = c.modified[b] != 0;
    if (!modified && blockIdx.x == 0)
        for (int v = threadIdx.x; v < a.nkg * W_HP; v += W_ET)
            kg_part[((int64_t)b * a.nkg) * W_HP + v] = kg_part[(((int64_t)6 * B + b) * a.nkg) * W_HP + v];
    if (i >= n) return;
    const int64_t e = (int64_t)b * n + i;
    T l = lnew[e];
    for (int sp = c.s_hi[b]; sp >= c.s_lo[b]; --sp) l += a.dg[((int64_t)b * a.nsave + sp) * n + i];
    a.lam[e] = l;
    if (!modified) a.kl[e] = a.kl[((int64_t)6 * B + b) * n + i];
}

template <class T>
__global__ void wsrc_grad_reduce_kernel(const T* gsrc, const int* ret, int np, int64_t B, T* out) {
    const int q = threadIdx.x;
    if (q >= np) return;
    double acc = 0.0;
    for (int64_t b = 0; b < B; ++b) if (ret[b] == RET_SUCCESS) acc += (double)gsrc[b * W_HP + q];
    out[q] = (T)acc;
}

// =========================================================================================================
// host side
// =========================================================================================================
inline SrcModel wsrc_model(const kanode_handle* h) {
    const kanode_desc& d = h->desc;
    const kanode_layer_desc& a = d.layers[0];
    SrcModel m{};
    m.n = d.n_state; m.norm = a.normalizer; m.inv_h = 1.0f / a.denominator;
    for (int g = 0; g < a.grid_len; ++g) m.grid[g] = grid_point(a, g);
    m.lap_scale = d.lap_coef / (d.dx * d.dx);
    m.offC = 0; m.offW = a.grid_len;
    return m;
}

template <class T, int G>
int wsrc_forward(kanode_handle* h, const SrcModel& m, const T* p, WideFwd<T> a, int64_t B, bool dense) {
    const int n = m.n;
    const int ec = (n + W_ET - 1) / W_ET;
    a.npart = ec;
    const size_t nB = (size_t)n * B;
    char* base = nullptr;
    ENSURE(h, W_WIDE_F, wide_ctl_bytes(B) + sizeof(T) * (9 * nB + (size_t)B * (1 + 2 * ec)) + 16 * 256, base);
    Arena A{base};
    WideCtl c = wide_ctl_carve(A, B);
    a.uprev = A.take<T>(nB); a.unew = A.take<T>(nB); a.k = A.take<T>(7 * nB); a.h = A.take<T>(B);
    a.part0 = A.take<T>((size_t)B * ec); a.part1 = A.take<T>((size_t)B * ec);
    if (!dense) { a.rec = nullptr; a.rec_t = nullptr; a.rec_dt = nullptr; }
    cudaStream_t st = h->stream;
    const dim3 ge(ec, (unsigned)B);
    int64_t launches = 0;
    auto rhs = [&](int ncoef, const T* coef, int kslot, T* xstore) {
        WideIn<T> in{};
        in.base = a.uprev; in.ks = a.k; in.hs = a.h; in.ncoef = ncoef;
        for (int j = 0; j < ncoef; ++j) in.coef[j] = coef[j];
        in.mask = c.active;
        wsrc_rhs_kernel<T, G><<<ge, W_ET, 0, st>>>(m, p, in, B, a.k + (size_t)kslot * nB, xstore);
        ++launches;
    };
    wide_fwd_init_kernel<T><<<ge, W_ET, 0, st>>>(c, a, n);
    rhs(0, nullptr, 0, nullptr);
    wide_norm_kernel<T, 1><<<ge, W_ET, 0, st>>>(a.uprev, a.k, nullptr, n, a.abstol, a.reltol, c.active, a.part0, a.part1, ec, 0);
    wide_fwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, a, n, 0);
    const T one[1] = {T(1)};
    rhs(1, one, 1, nullptr);
    wide_norm_kernel<T, 2><<<ge, W_ET, 0, st>>>(a.uprev, a.k, a.k + nB, n, a.abstol, a.reltol, c.active, a.part0, a.part1, ec, 0);
    wide_fwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, a, n, 1);
    launches += 5;
    static const double A_[7][8] = KANODE_TSIT5_A;
    auto attempt = [&]() -> int64_t {
        const int64_t l0 = launches;
        for (int s = 1; s < 7; ++s) {
            T coef[7];
            for (int j = 0; j < s; ++j) coef[j] = (T)A_[s][j];
            rhs(s, coef, s, s == 6 ? a.unew : nullptr);
        }
        wide_err_kernel<T><<<ge, W_ET, 0, st>>>(a.uprev, a.unew, a.k, n, B, a.h, a.abstol, a.reltol, c.active, a.part0, ec, 0);
        wide_fwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, a, n, 2);
        wide_fwd_accept_kernel<T><<<ge, W_ET, 0, st>>>(c, a, n, B);
        launches += 3;
        return launches - l0;
    };
    int iters = 0;
    const int slot = dense ? 1 : 0;
    const int expect = h->wide_iters[slot];
    bool any = a.t0 < a.t1;
    cudaGraphExec_t gexec = nullptr;
    const int gslot = slot * 2 + (sizeof(T) == 8);
    if (h->wide_graph && any) {
        std::vector<char> sig;
        sig_add(sig, m); sig_add(sig, a); sig_add(sig, c); sig_add(sig, B); sig_add(sig, p);
        const int64_t l0 = launches;
        if (int rc = wide_attempt_graph(h, gslot, sig, attempt, &gexec)) return rc;
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
        if (iters > a.maxiters + 2) break;
    }
    h->wide_iters[slot] = iters;
    wide_fwd_finish_kernel<T><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(c, a, B);
    h->launches += launches + 1;
    CK(h, cudaGetLastError());
    return 0;
}

template <class T, int G>
int wsrc_solve_t(kanode_handle* h, const T* p, const T* d_u0, int64_t B, double t0, double t1, const double* d_saveat, int nsave,
                 double abstol, double reltol, T* d_out, kanode_stats* d_stats) {
    const SrcModel m = wsrc_model(h);
    WideFwd<T> a{};
    a.u0 = d_u0; a.t0 = t0; a.t1 = t1; a.saveat = d_saveat; a.nsave = nsave; a.abstol = (T)abstol; a.reltol = (T)reltol;
    a.maxiters = 100000; a.out = d_out; a.stats = d_stats;
    return wsrc_forward<T, G>(h, m, p, a, B, false);
}

template <class T, int G>
int wsrc_loss_grad_t(kanode_handle* h, const T* p, const T* d_u0, int64_t B, double t0, double t1, const double* d_saveat, int nsave,
                     const T* d_target, double abstol, double reltol, double* d_loss_sum, T* d_grad_sum, T* d_du0,
                     kanode_stats* d_fst, kanode_stats* d_bst, T* d_out_opt) {
    const SrcModel m = wsrc_model(h);
    const int n = m.n, np = (int)h->np;
    const size_t nB = (size_t)n * B;
    const int cap = h->rec_cap;
    cudaStream_t st = h->stream;
    WideFwd<T> a{};
    a.u0 = d_u0; a.t0 = t0; a.t1 = t1; a.saveat = d_saveat; a.nsave = nsave; a.abstol = (T)abstol; a.reltol = (T)reltol;
    a.maxiters = 100000; a.out = d_out_opt; a.stats = d_fst; a.target = d_target; a.loss_sum = d_loss_sum; a.cap = cap;
    ENSURE(h, W_REC_T, sizeof(double) * (size_t)cap * B, a.rec_t);
    ENSURE(h, W_GEN2, sizeof(T) * (size_t)cap * B, a.rec_dt);
    ENSURE(h, W_REC, sizeof(T) * (size_t)cap * 8 * nB, a.rec);
    ENSURE(h, W_NSTEPS, sizeof(int) * (size_t)B, a.nsteps);
    ENSURE(h, W_RET, sizeof(int) * (size_t)B, a.retcode);
    ENSURE(h, W_DG, sizeof(T) * (size_t)nsave * nB, a.dg);
    cudaEventRecord(h->ev[0], st);
    if (int rc = wsrc_forward<T, G>(h, m, p, a, B, true)) return rc;
    cudaEventRecord(h->ev[1], st);
    // ---- backward ----
    const int ec = (n + W_ET - 1) / W_ET;
    char* base = nullptr;
    ENSURE(h, W_WIDE_B, wide_ctl_bytes(B) + sizeof(T) * (2 * nB + 7 * nB + (size_t)7 * B * ec * W_HP + (size_t)B * (W_HP + 15) + 2 * (size_t)B * ec) + 24 * 256, base);
    Arena A{base};
    WideCtl c = wide_ctl_carve(A, B);
    WideBwd<T> w{};
    w.lam = A.take<T>(nB); w.kl = A.take<T>(7 * nB);
    T* lnew = A.take<T>(nB);
    T* kg_part = A.take<T>((size_t)7 * B * ec * W_HP);
    w.gsrc = A.take<T>((size_t)B * W_HP);
    w.h = A.take<T>(B); w.th = A.take<T>(7 * (size_t)B); w.hd = A.take<T>(7 * (size_t)B);
    w.part0 = A.take<T>((size_t)B * ec); w.part1 = A.take<T>((size_t)B * ec);
    w.t0 = t0; w.t1 = t1; w.saveat = d_saveat; w.nsave = nsave; w.abstol = (T)abstol; w.reltol = (T)reltol; w.maxiters = 100000;
    w.rec_t = a.rec_t; w.rec_dt = a.rec_dt; w.rec = a.rec; w.cap = cap; w.nsteps = a.nsteps; w.retcode = a.retcode;
    w.dg = a.dg; w.g = nullptr; w.np = np; w.npart = ec; w.du0 = d_du0; w.stats = d_bst;
    w.kg_part = kg_part; w.nkg = ec;
    CK(h, cudaMemsetAsync(w.gsrc, 0, sizeof(T) * (size_t)B * W_HP, st));
    CK(h, cudaMemsetAsync(kg_part, 0, sizeof(T) * (size_t)7 * B * ec * W_HP, st));
    const dim3 ge(ec, (unsigned)B);
    int64_t launches = 0;
    static const double A_[7][8] = KANODE_TSIT5_A;
    auto adj = [&](int s, int ncoef, const T* coef, const int* mask) {
        WideIn<T> iy{};
        iy.rec = w.rec; iy.cap = cap; iy.ridx = c.ridx + (size_t)s * B; iy.th = w.th + (size_t)s * B; iy.hd = w.hd + (size_t)s * B;
        WideIn<T> il{};
        il.base = w.lam; il.ks = w.kl; il.hs = w.h; il.ncoef = ncoef;
        for (int j = 0; j < ncoef; ++j) il.coef[j] = coef[j];
        il.mask = mask;
        wsrc_vjp_kernel<T, G><<<ge, W_ET, 0, st>>>(m, p, iy, il, B, w.kl + (size_t)s * nB, s == 6 ? lnew : nullptr, kg_part + (size_t)s * B * ec * W_HP);
        ++launches;
    };
    wide_bwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, w, n, B, -1);
    wide_bwd_init_kernel<T><<<ge, W_ET, 0, st>>>(c, w, n);
    adj(0, 0, nullptr, c.active);
    wide_norm_kernel<T, 1><<<ge, W_ET, 0, st>>>(w.lam, w.kl, nullptr, n, w.abstol, w.reltol, c.active, w.part0, w.part1, ec, 0);
    wide_bwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, w, n, B, 0);
    const T one[1] = {T(1)};
    adj(1, 1, one, c.active);
    wide_norm_kernel<T, 2><<<ge, W_ET, 0, st>>>(w.lam, w.kl, w.kl + nB, n, w.abstol, w.reltol, c.active, w.part0, w.part1, ec, 0);
    wide_bwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, w, n, B, 1);
    launches += 6;
    auto attempt = [&]() -> int64_t {
        const int64_t l0 = launches;
        adj(0, 0, nullptr, c.do_s0);
        for (int s = 1; s < 7; ++s) {
            T coef[7];
            for (int j = 0; j < s; ++j) coef[j] = (T)A_[s][j];
            adj(s, s, coef, c.active);
        }
        wide_err_kernel<T><<<ge, W_ET, 0, st>>>(w.lam, lnew, w.kl, n, B, w.h, w.abstol, w.reltol, c.active, w.part0, ec, 0);
        wide_bwd_ctl_kernel<T><<<(unsigned)B, 128, 0, st>>>(c, w, n, B, 2);
        wsrc_bwd_accept_kernel<T><<<ge, W_ET, 0, st>>>(c, w, lnew, kg_part, n, B);
        launches += 3;
        return launches - l0;
    };
    int iters = 0;
    const int expect = h->wide_iters[2];
    bool any = t0 < t1;
    cudaGraphExec_t gexec = nullptr;
    const int gslot = 4 + (sizeof(T) == 8);
    if (h->wide_graph && any) {
        std::vector<char> sig;
        sig_add(sig, m); sig_add(sig, w); sig_add(sig, c); sig_add(sig, B); sig_add(sig, p); sig_add(sig, lnew); sig_add(sig, kg_part);
        const int64_t l0 = launches;
        if (int rc = wide_attempt_graph(h, gslot, sig, attempt, &gexec)) return rc;
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
    wsrc_grad_reduce_kernel<T><<<1, 32, 0, st>>>(w.gsrc, c.ret, np, B, d_grad_sum);
    cudaEventRecord(h->ev[3], st);
    h->ev_valid = true;
    h->launches += launches + 2;
    CK(h, cudaGetLastError());
    return 0;
}

}  // namespace kanode
