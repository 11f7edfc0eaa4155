// kanode_small_ls.cuh — LOCKSTEP backward pass for large ensembles of small KAN-ODEs.
//
// The monolithic small_backward_kernel keeps a whole adjoint solve in one thread: its 7 stage records need 672 B of
// shared memory and ~250 registers per trajectory, which caps the SM at 8 warps, and a single trajectory is latency
// bound (measured: the kernel takes ~3.0 ms at 2 warps/SM and 3.4 ms at 8 warps/SM — profiles/r01_*).  Here one
// backward STEP ATTEMPT of every trajectory is split into two lean kernels that all trajectories execute in lockstep:
//   ls_step_kernel    thread per trajectory: finishes the previous attempt (error norm, PI controller, accept/reject,
//                     jumps at save times), then runs the 7 fused forward+VJP stage evaluations of the next attempt.
//                     Stage records go to global memory (SoA, coalesced, L2-resident: 672 B x B).
//   ls_gphase_kernel  thread per (trajectory, unit): the step-end pass over the NP gradient components (feature
//                     recompute, rank-1 accumulation, error-estimate terms, g update) — 14 independent work items per
//                     trajectory instead of one long serial loop.
// Per-trajectory step control is unchanged (each trajectory has its own t, dt, accept/reject, tstops); the arithmetic
// is the same as in the monolithic kernel, so fp64 results still reproduce the oracle's step sequence.
#pragma once
#include "kanode_small.cuh"

#ifndef KANODE_LS_BT
#define KANODE_LS_BT 64         // threads per block of the lockstep stage kernel
#endif
#ifndef KANODE_LS_MINB
#define KANODE_LS_MINB 7        // resident blocks per SM it is compiled for (7 x 64 = 448 trajectories/SM: 65,536 in one round)
#endif

namespace kanode {

enum { LS_DONE = 1, LS_ATTEMPT = 2, LS_ACCEPT = 4, LS_MODIFIED = 8 };

template <class T> struct LsState {        // struct of arrays over the B trajectories of this launch (device pointers)
    double *t, *dt, *dtpropose, *qold, *q11;
    int *iter, *sp, *cur, *naccept, *nreject, *nf, *ret, *flags, *ridx;
    T* lam;      // [I][B]
    T* lprev;    // [I][B]
    T* kl;       // [7][I][B]  stage derivatives of lambda
    T* lnew;     // [I][B]
    T* es_l;     // [B]        lambda part of the squared error norm of the current attempt
    T* es_part;  // [NITEM][B] gradient-component parts, one per g-phase work item
    T* rec;      // [7][StageRec::N][B] stage records
    int* active; // number of trajectories still integrating
};

template <class P> struct LsItems {
    static constexpr int OC = (P::H % 5 == 0) ? 5 : (P::H % 4 == 0 ? 4 : (P::H % 2 == 0 ? 2 : 1));
    static constexpr int N1 = P::I * (P::H / OC);     // layer-1 items: (input i, chunk of OC outputs)
    static constexpr int NITEM = P::H + N1;           // layer-2 items: one per hidden unit
};

// per-thread evaluation context: dense-record cache + fused forward/VJP evaluation writing a global stage record
template <class T, class P, int NORM> struct LsEval {
    static constexpr int I = P::I, RS = 1 + 8 * P::I;
    using SR = StageRec<P>;
    const P& prm; const T* wsm; const SmallBwdArgs<T>& a;
    int64_t b, B; int nsteps; T* rec; int nf;
    int ridx; double rt, rt_next; T rdt, ru[P::I], rk[7][P::I];

    __device__ __forceinline__ LsEval(const P& p_, const T* w_, const SmallBwdArgs<T>& a_, int64_t b_, T* rec_, int nsteps_)
        : prm(p_), wsm(w_), a(a_), b(b_), B(a_.B), nsteps(nsteps_), rec(rec_), nf(0) {}
    __device__ __forceinline__ void load_rec(int idx) {
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
    }
    __device__ __forceinline__ void eval_y(double t, T (&y)[P::I]) {
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
    }
    __device__ __forceinline__ void adj_eval(double t, const T (&l)[P::I], T (&dl)[P::I], int slot) {
        T y[P::I], ub[P::I];
        eval_y(t, y);
        T* s = rec + (int64_t)slot * SR::N * B;
        small_vjp_sm<NORM, KANODE_UNROLL_J>(prm, wsm, y, l, ub, s, (int)B, SR::HH, SR::HBAR);
#pragma unroll
        for (int i = 0; i < I; ++i) { s[(int64_t)(SR::Y + i) * B] = y[i]; s[(int64_t)(SR::LAM + i) * B] = l[i]; dl[i] = -ub[i]; }
        ++nf;
    }
};

// visit every gradient component of NS stage records (global, stride B): fn(j, kv[NS]), kv[s] = (df/dp)^T lam at stage s
template <int NS, int NORM, class T, class P, class Fn>
__device__ __forceinline__ void ls_for_each_g(const P& prm, const T* rec, int64_t B, Fn&& fn) {
    constexpr int I = P::I, H = P::H, G = P::G;
    using SR = StageRec<P>;
#pragma unroll 1
    for (int i = 0; i < I; ++i) {
        T c[NS][G + 1];
#pragma unroll
        for (int s = 0; s < NS; ++s) unit_features<NORM>(prm, rec[(int64_t)(s * SR::N + SR::Y + i) * B], c[s]);
#pragma unroll 1
        for (int o = 0; o < H; ++o) {
            T av[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) av[s] = rec[(int64_t)(s * SR::N + SR::HBAR + o) * B];
#pragma unroll
            for (int q = 0; q <= G; ++q) {
                T kv[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) kv[s] = av[s] * c[s][q];
                fn(q < G ? P::OC1 + (i * G + q) * H + o : P::OW1 + i * H + o, kv);
            }
        }
    }
    T al[NS][I];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int o = 0; o < I; ++o) al[s][o] = rec[(int64_t)(s * SR::N + SR::LAM + o) * B];
#pragma unroll 1
    for (int i = 0; i < H; ++i) {
        T c[NS][G + 1];
#pragma unroll
        for (int s = 0; s < NS; ++s) unit_features<NORM>(prm, rec[(int64_t)(s * SR::N + SR::HH + i) * B], c[s]);
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
}

// ---------------------------------------------------------------------------------------------------------
// K0: initialise every trajectory's adjoint solve (jump at T, FSAL evaluation, Hairer initial dt)
// ---------------------------------------------------------------------------------------------------------
template <class T, class P, int NORM>
__global__ void __launch_bounds__(128) ls_init_kernel(const __grid_constant__ P prm, const SmallBwdArgs<T> a, const LsState<T> st) {
    constexpr int I = P::I, NP = P::NP, NZ = I + NP;
    __shared__ __align__(16) T wsm[P::WPK];
    __shared__ uint64_t wbar;
    stage_weights<T, P::WPK>(wsm, &wbar, a.wpk);
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int64_t B = a.B;
    T* gbuf = a.g + b;
#pragma unroll 1
    for (int j = 0; j < NP; ++j) gbuf[(int64_t)j * B] = T(0);
    const int ret = a.retcode[b], nsteps = a.nsteps[b];
    st.cur[b] = 0; st.naccept[b] = 0; st.nreject[b] = 0; st.iter[b] = 0; st.ret[b] = ret;
    if (ret != RET_SUCCESS || nsteps <= 0) {
        st.flags[b] = LS_DONE; st.nf[b] = 0;
        if (a.stats) a.stats[b] = kanode_stats{0, 0, 0, ret};
        if (a.du0) for (int i = 0; i < I; ++i) a.du0[b * I + i] = T(0);
        return;
    }
    LsEval<T, P, NORM> ev(prm, wsm, a, b, st.rec + b, nsteps);
    ev.load_rec(nsteps - 1);
    const double t0 = a.t0, t1 = a.t1, dtmax = fabs(t1 - t0), dtmin0 = fmax(eps_of(t0), eps_of(t1));
    const T abstol = a.abstol, reltol = a.reltol;
    T lam[I], k0[I];
#pragma unroll
    for (int i = 0; i < I; ++i) lam[i] = T(0);
    int sp = a.nsave - 1;
    while (sp >= 0 && a.saveat[sp] == t1) {       // PresetTimeCallback fires at init when t_end is a save time
#pragma unroll
        for (int i = 0; i < I; ++i) lam[i] += a.dg[((int64_t)sp * I + i) * B + b];
        --sp;
    }
    ev.adj_eval(t1, lam, k0, 0);
    double dt;
    {
        T sk[I], s0 = T(0), s1 = T(0);
#pragma unroll
        for (int i = 0; i < I; ++i) {
            sk[i] = abstol + kabs(lam[i]) * reltol;
            const T x0 = lam[i] / sk[i], x1 = k0[i] / sk[i];
            s0 += x0 * x0; s1 += x1 * x1;
        }
        ls_for_each_g<1, NORM>(prm, ev.rec, B, [&](int, const T (&kv)[1]) { const T x = kv[0] / abstol; s1 += x * x; });
        const double d0 = sqrt((double)s0 / NZ), d1 = sqrt((double)s1 / NZ);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
        dt0 = fmin(dt0, dtmax);
        T l1[I], f1[I];
#pragma unroll
        for (int i = 0; i < I; ++i) l1[i] = lam[i] - (T)dt0 * k0[i];
        ev.adj_eval(t1 - dt0, l1, f1, 1);
        ++ev.nf;
        T s2 = T(0);
#pragma unroll
        for (int i = 0; i < I; ++i) { const T x = (f1[i] - k0[i]) / sk[i]; s2 += x * x; }
        ls_for_each_g<2, NORM>(prm, ev.rec, B, [&](int, const T (&kv)[2]) { const T x = (kv[1] - kv[0]) / abstol; s2 += x * x; });
        const double d2 = sqrt((double)s2 / NZ) / dt0, mx = fmax(d1, d2);
        const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        dt = fmax(dtmin0, fmin(fmin(100.0 * dt0, dt1), dtmax));
    }
    st.t[b] = t1; st.dt[b] = dt; st.dtpropose[b] = dt; st.qold[b] = Ctrl::qoldinit; st.q11[b] = 1.0;
    st.sp[b] = sp; st.nf[b] = ev.nf; st.flags[b] = 0; st.ridx[b] = ev.ridx;
#pragma unroll
    for (int i = 0; i < I; ++i) {
        st.lam[(int64_t)i * B + b] = lam[i]; st.lprev[(int64_t)i * B + b] = lam[i];
        st.kl[(int64_t)i * B + b] = k0[i];
#pragma unroll
        for (int j = 1; j < 7; ++j) st.kl[((int64_t)j * I + i) * B + b] = T(0);   // zero-weighted stages must be finite
    }
    atomicAdd(st.active, 1);
}

// ---------------------------------------------------------------------------------------------------------
// K1: finish the previous attempt (loopfooter!), then loopheader! + the stage evaluations of the next attempt
// ---------------------------------------------------------------------------------------------------------
template <class T, class P, int NORM>
__global__ void __launch_bounds__(KANODE_LS_BT, KANODE_LS_MINB) ls_step_kernel(const __grid_constant__ P prm, const SmallBwdArgs<T> a, const LsState<T> st) {
    constexpr int I = P::I, NP = P::NP, NZ = I + NP, NITEM = LsItems<P>::NITEM;
    using SR = StageRec<P>;
    __shared__ __align__(16) T wsm[P::WPK];
    __shared__ uint64_t wbar;
    stage_weights<T, P::WPK>(wsm, &wbar, a.wpk);
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int64_t B = a.B;
    int flags = st.flags[b];
    if (flags & LS_DONE) return;
    const double t0 = a.t0, t1 = a.t1, dtmax = fabs(t1 - t0), dtmin0 = fmax(eps_of(t0), eps_of(t1));
    const T abstol = a.abstol, reltol = a.reltol;
    double t = st.t[b], dt = st.dt[b], dtpropose = st.dtpropose[b], qold = st.qold[b], q11 = st.q11[b];
    int iter = st.iter[b], sp = st.sp[b], ret = RET_SUCCESS;
    T lam[I], lprev[I], kl[7][I];
#pragma unroll
    for (int i = 0; i < I; ++i) { lam[i] = st.lam[(int64_t)i * B + b]; lprev[i] = st.lprev[(int64_t)i * B + b]; }
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int i = 0; i < I; ++i) kl[j][i] = st.kl[((int64_t)j * I + i) * B + b];
    T* rec = st.rec + b;
    bool accept = false, modified = false, finished = false;
    if (flags & LS_ATTEMPT) {
        // ---- loopfooter! of the attempt whose stages ran in the previous launch ----
        T es = st.es_l[b];
#pragma unroll
        for (int k = 0; k < NITEM; ++k) es += st.es_part[(int64_t)k * B + b];
        const double EEst = (double)ksqrt(es / T(NZ));
        if (EEst != EEst) { ret = RET_UNSTABLE; finished = true; }
        else {
            const double tstop = (sp >= 0) ? fmax(a.saveat[sp], t0) : t0;
            const double q = pi_q(EEst, qold, q11);
            accept = EEst <= 1.0;
            if (accept) {
                ++st.naccept[b];
                qold = fmax(EEst, Ctrl::qoldinit);
                const double dtnew = dt / q;
                double tnew = t - dt;
                if (fabs(tnew - tstop) < 100.0 * eps_of(fmax(fabs(t), fabs(tstop)))) tnew = tstop;
                dtpropose = fmax(fmin(dtmax, fabs(dtnew)), fmax(eps_of(tnew), dtmin0));
                t = tnew;
                st.cur[b] ^= 1;
#pragma unroll
                for (int i = 0; i < I; ++i) lam[i] = st.lnew[(int64_t)i * B + b];
                while (sp >= 0 && a.saveat[sp] == t) {                       // jumps at the save times
#pragma unroll
                    for (int i = 0; i < I; ++i) lam[i] += a.dg[((int64_t)sp * I + i) * B + b];
                    --sp; modified = true;
                }
#pragma unroll
                for (int i = 0; i < I; ++i) lprev[i] = lam[i];
                if (!(t > t0)) finished = true;
            } else {
                ++st.nreject[b];
            }
        }
    }
    if (!finished) {
        // ---- loopheader! ----
        if (iter > 0) {
            if (!accept) dt = dt / fmin(1.0 / Ctrl::qmin, q11 / Ctrl::gamma);
            else {
                dt = dtpropose;
                if (!modified) {                                              // FSAL
#pragma unroll
                    for (int i = 0; i < I; ++i) kl[0][i] = kl[6][i];
#pragma unroll
                    for (int f = 0; f < SR::N; ++f) rec[(int64_t)f * B] = rec[(int64_t)(6 * SR::N + f) * B];
                }
            }
        }
        ++iter;
        const double tstop = (sp >= 0) ? fmax(a.saveat[sp], t0) : t0;
        const double dtmin_t = fmax(eps_of(t), dtmin0);
        dt = fmin(fmax(fmin(fabs(dt), dtmax), dtmin_t), t - tstop);
        if (iter > a.maxiters) { ret = RET_MAXITERS; finished = true; }
        else if (!(dt > dtmin_t) && (t - dt > tstop || !accept) && iter > 1) { ret = RET_DTMIN; finished = true; }
        else if (dt != dt) { ret = RET_UNSTABLE; finished = true; }
    }
    if (finished) {
        st.flags[b] = LS_DONE; st.ret[b] = ret; st.t[b] = t; st.sp[b] = sp;
        if (a.du0)
#pragma unroll
            for (int i = 0; i < I; ++i) a.du0[b * I + i] = lam[i];
        if (a.stats) a.stats[b] = kanode_stats{st.naccept[b], st.nreject[b], st.nf[b], ret};
        atomicSub(st.active, 1);
        return;
    }
    // ---- perform_step!: the stage evaluations of this attempt ----
    LsEval<T, P, NORM> ev(prm, wsm, a, b, rec, a.nsteps[b]);
    ev.load_rec(st.ridx[b]);
    const T h = (T)(-dt);
    T lnew[I];
#pragma unroll
    for (int i = 0; i < I; ++i) lnew[i] = lprev[i];
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
        ev.adj_eval(t - tab_c(s) * dt, ls, ks, s);
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
    T es = T(0);
#pragma unroll
    for (int i = 0; i < I; ++i) {
        T ut = T(0);
#pragma unroll
        for (int j = 0; j < 7; ++j) ut += Tab<T>::bt(j) * kl[j][i];
        ut *= h;
        const T sc = abstol + kmax(kabs(lprev[i]), kabs(lnew[i])) * reltol;
        const T r = ut / sc;
        es += r * r;
        if (lnew[i] != lnew[i]) es = lnew[i];                                // propagate NaN into the error norm
    }
    st.t[b] = t; st.dt[b] = dt; st.dtpropose[b] = dtpropose; st.qold[b] = qold; st.q11[b] = q11;
    st.iter[b] = iter; st.sp[b] = sp; st.nf[b] += ev.nf; st.ridx[b] = ev.ridx; st.es_l[b] = es;
    st.flags[b] = LS_ATTEMPT;
#pragma unroll
    for (int i = 0; i < I; ++i) {
        st.lam[(int64_t)i * B + b] = lam[i]; st.lprev[(int64_t)i * B + b] = lprev[i]; st.lnew[(int64_t)i * B + b] = lnew[i];
#pragma unroll
        for (int j = 0; j < 7; ++j) st.kl[((int64_t)j * I + i) * B + b] = kl[j][i];
    }
}

// ---------------------------------------------------------------------------------------------------------
// One work item of the step-end pass over the gradient components: item < H is hidden unit `item` of layer 2
// (its C2 rows and W2 row), item >= H is an (input, output-chunk) tile of layer 1.  rec: stage records with element
// (slot, f) at rec[(slot*SR::N + f)*rs]; gold/gnew: gradient buffers with component j at [j*gs].
// Returns the item's part of the squared error norm.
// ---------------------------------------------------------------------------------------------------------
template <int NORM, class T, class P>
__device__ __forceinline__ T gphase_item(const P& prm, int item, const T* rec, int64_t rs, const T* gold, T* gnew, int64_t gs,
                                         T mh, T abstol, T reltol) {
    constexpr int I = P::I, H = P::H, G = P::G, OC = LsItems<P>::OC;
    using SR = StageRec<P>;
    T es = T(0);
    auto finalize = [&](int j, T g0, T vb, T vt) {
        const T g1 = g0 + mh * vb;
        const T sc = abstol + kmax(kabs(g0), kabs(g1)) * reltol;
        const T r = kdiv(mh * vt, sc);
        es += r * r;
        gnew[(int64_t)j * gs] = g1;
    };
    if (item < H) {                                              // layer 2: hidden unit i, outputs o < I
        const int i = item;
        T g0[G + 1][I], vb[G + 1][I], vt[G + 1][I];
#pragma unroll
        for (int q = 0; q <= G; ++q)
#pragma unroll
            for (int o = 0; o < I; ++o) {
                const int j = q < G ? P::OC2 + (i * G + q) * I + o : P::OW2 + i * I + o;
                g0[q][o] = gold[(int64_t)j * gs]; vb[q][o] = T(0); vt[q][o] = T(0);
            }
#pragma unroll
        for (int s = 0; s < 7; ++s) {
            const T* r = rec + (int64_t)s * SR::N * rs;
            T c[G + 1];
            unit_features<NORM>(prm, r[(int64_t)(SR::HH + i) * rs], c);
            const T wb = Tab<T>::b(s), wt = Tab<T>::bt(s);
#pragma unroll
            for (int o = 0; o < I; ++o) {
                const T l = r[(int64_t)(SR::LAM + o) * rs];
                const T ab = wb * l, at = wt * l;
#pragma unroll
                for (int q = 0; q <= G; ++q) { vb[q][o] += ab * c[q]; vt[q][o] += at * c[q]; }
            }
        }
#pragma unroll
        for (int q = 0; q <= G; ++q)
#pragma unroll
            for (int o = 0; o < I; ++o)
                finalize(q < G ? P::OC2 + (i * G + q) * I + o : P::OW2 + i * I + o, g0[q][o], vb[q][o], vt[q][o]);
    } else {                                                     // layer 1: input i, outputs o0 .. o0+OC-1
        const int io = item - H;
        const int i = io / (H / OC), o0 = (io % (H / OC)) * OC;
        T g0[G + 1][OC], vb[G + 1][OC], vt[G + 1][OC];
#pragma unroll
        for (int q = 0; q <= G; ++q)
#pragma unroll
            for (int oo = 0; oo < OC; ++oo) {
                const int j = (q < G ? P::OC1 + (i * G + q) * H : P::OW1 + i * H) + o0 + oo;
                g0[q][oo] = gold[(int64_t)j * gs]; vb[q][oo] = T(0); vt[q][oo] = T(0);
            }
#pragma unroll
        for (int s = 0; s < 7; ++s) {
            const T* r = rec + (int64_t)s * SR::N * rs;
            T c[G + 1];
            unit_features<NORM>(prm, r[(int64_t)(SR::Y + i) * rs], c);
            const T wb = Tab<T>::b(s), wt = Tab<T>::bt(s);
#pragma unroll
            for (int oo = 0; oo < OC; ++oo) {
                const T hb = r[(int64_t)(SR::HBAR + o0 + oo) * rs];
                const T ab = wb * hb, at = wt * hb;
#pragma unroll
                for (int q = 0; q <= G; ++q) { vb[q][oo] += ab * c[q]; vt[q][oo] += at * c[q]; }
            }
        }
#pragma unroll
        for (int q = 0; q <= G; ++q)
#pragma unroll
            for (int oo = 0; oo < OC; ++oo)
                finalize((q < G ? P::OC1 + (i * G + q) * H : P::OW1 + i * H) + o0 + oo, g0[q][oo], vb[q][oo], vt[q][oo]);
    }
    return es;
}

// K2 of the lockstep engine: blockIdx.y selects the work item, so every warp runs one item type
template <class T, class P, int NORM>
__global__ void __launch_bounds__(128) ls_gphase_kernel(const __grid_constant__ P prm, const SmallBwdArgs<T> a, const LsState<T> st) {
    constexpr int NP = P::NP;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int64_t B = a.B;
    if (!(st.flags[b] & LS_ATTEMPT)) return;
    const int item = blockIdx.y;
    const int cur = st.cur[b];
    const T es = gphase_item<NORM>(prm, item, st.rec + b, B, a.g + b + (int64_t)cur * NP * B, a.g + b + (int64_t)(cur ^ 1) * NP * B, B,
                                   (T)st.dt[b], a.abstol, a.reltol);                 // -h = +dt  (dg/dt = -kv, h = -dt)
    st.es_part[(int64_t)item * B + b] = es;
}

// ---------------------------------------------------------------------------------------------------------
// WARP-PER-TRAJECTORY backward solve for the few trajectories predicted to need many steps.  Their serial chain of
// steps bounds the whole launch, so here one warp owns one trajectory (and, through an exclusive shared-memory
// request, one SM): lane 0 runs the control flow and the 7 fused forward+VJP stage evaluations exactly like the
// thread-per-trajectory kernel; the step-end pass over the NP gradient components — 40 % of a step's instructions —
// is spread over the lanes (one work item per lane).  Same arithmetic, same results.
// ---------------------------------------------------------------------------------------------------------
template <class T, class P, int NORM>
__global__ void __launch_bounds__(384) small_backward_warp_kernel(const __grid_constant__ P prm, const SmallBwdArgs<T> a) {
    constexpr int I = P::I, NP = P::NP, NZ = I + NP, RS = 1 + 8 * I, NITEM = LsItems<P>::NITEM;
    using SR = StageRec<P>;
    static_assert(NITEM <= 32, "one work item per lane");
    // dynamic shared memory: [packed weights | mbarrier | per warp: 7 stage records, -h, current g buffer]
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* wsm = reinterpret_cast<T*>(smem_raw);
    uint64_t* wbar = reinterpret_cast<uint64_t*>(wsm + P::WPK);
    constexpr int PER_WARP = 7 * SR::N + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    T* rec = reinterpret_cast<T*>(wbar + 2) + warp * PER_WARP;
    T& s_mh = rec[7 * SR::N];
    int& s_cur = *reinterpret_cast<int*>(&rec[7 * SR::N + 1]);
    stage_weights<T, P::WPK>(wsm, wbar, a.wpk);
    const int slot_id = blockIdx.x * nwarps + warp;
    if (slot_id >= *a.long_count || slot_id >= a.gidn) return;
    const int64_t b = a.long_list[slot_id];
    const int64_t B = a.B;
    T* gbuf = a.g + b;
    for (int j = lane; j < NP; j += 32) gbuf[(int64_t)j * B] = T(0);
    __syncwarp();
    // ---- everything below mirrors small_backward_kernel; lane 0 holds the solver state ----
    T lam[I], lprev[I], kl[7][I];
#pragma unroll
    for (int i = 0; i < I; ++i) { lam[i] = T(0); lprev[i] = T(0); }
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int i = 0; i < I; ++i) kl[j][i] = T(0);
    int nf = 0, naccept = 0, nreject = 0, ret = a.retcode[b];
    const int nsteps = a.nsteps[b];
    if (ret != RET_SUCCESS || nsteps <= 0) {
        if (lane == 0) {
            if (a.stats) a.stats[b] = kanode_stats{0, 0, 0, ret};
            if (a.du0) for (int i = 0; i < I; ++i) a.du0[b * I + i] = T(0);
            if (a.attempts) a.attempts[b] = 0;
        }
        return;
    }
    int ridx = nsteps - 1;
    double rt = 0, rt_next = 0;
    T rdt = T(1), ru[I], rk[7][I];
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
    auto adj_eval = [&](double t, const T (&l)[I], T (&dl)[I], int slot) {
        while (t < rt && ridx > 0) load_rec(ridx - 1);
        while (t >= rt_next && ridx + 1 < nsteps) load_rec(ridx + 1);
        const T th = (T)((t - rt) / (double)rdt);
        T bw[7]; interp_weights(th, bw);
        T y[I], ub[I];
#pragma unroll
        for (int i = 0; i < I; ++i) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < 7; ++j) acc += bw[j] * rk[j][i];
            y[i] = ru[i] + rdt * acc;
        }
        T* s = rec + slot * SR::N;
        small_vjp_sm<NORM, KANODE_UNROLL_J>(prm, wsm, y, l, ub, s, 1, SR::HH, SR::HBAR);
#pragma unroll
        for (int i = 0; i < I; ++i) { s[SR::Y + i] = y[i]; s[SR::LAM + i] = l[i]; dl[i] = -ub[i]; }
        ++nf;
    };
    const double t0 = a.t0, t1 = a.t1, dtmax = fabs(t1 - t0), dtmin0 = fmax(eps_of(t0), eps_of(t1));
    const T abstol = a.abstol, reltol = a.reltol;
    double t = t1, dt = 0, qold = Ctrl::qoldinit, q11 = 1.0, dtpropose = 0;
    int sp = a.nsave - 1, iter = 0, cur = 0;
    bool accept = false, modified = false;
    if (lane == 0) {
        load_rec(ridx);
        while (sp >= 0 && a.saveat[sp] == t1) {
#pragma unroll
            for (int i = 0; i < I; ++i) lam[i] += a.dg[((int64_t)sp * I + i) * B + b];
            --sp;
        }
#pragma unroll
        for (int i = 0; i < I; ++i) lprev[i] = lam[i];
        adj_eval(t, lam, kl[0], 0);
        T sk[I], s0 = T(0), s1 = T(0);
#pragma unroll
        for (int i = 0; i < I; ++i) {
            sk[i] = abstol + kabs(lam[i]) * reltol;
            const T x0 = lam[i] / sk[i], x1 = kl[0][i] / sk[i];
            s0 += x0 * x0; s1 += x1 * x1;
        }
        ls_for_each_g<1, NORM>(prm, rec, 1, [&](int, const T (&kv)[1]) { const T x = kv[0] / abstol; s1 += x * x; });
        const double d0 = sqrt((double)s0 / NZ), d1 = sqrt((double)s1 / NZ);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
        dt0 = fmin(dt0, dtmax);
        T l1[I], f1[I];
#pragma unroll
        for (int i = 0; i < I; ++i) l1[i] = lam[i] - (T)dt0 * kl[0][i];
        adj_eval(t - dt0, l1, f1, 1);
        ++nf;
        T s2 = T(0);
#pragma unroll
        for (int i = 0; i < I; ++i) { const T x = (f1[i] - kl[0][i]) / sk[i]; s2 += x * x; }
        ls_for_each_g<2, NORM>(prm, rec, 1, [&](int, const T (&kv)[2]) { const T x = (kv[1] - kv[0]) / abstol; s2 += x * x; });
        const double d2 = sqrt((double)s2 / NZ) / dt0, mx = fmax(d1, d2);
        const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        dt = fmax(dtmin0, fmin(fmin(100.0 * dt0, dt1), dtmax));
        dtpropose = dt;
    }
    T es_l = T(0), h = T(0);
    T lnew[I];
#pragma unroll
    for (int i = 0; i < I; ++i) lnew[i] = T(0);
    double tstop = t0;
    for (;;) {
        int go = 0;
        if (lane == 0) {
            go = t > t0;
            if (go) {
                // ---- loopheader! ----
                if (iter > 0) {
                    if (!accept) dt = dt / fmin(1.0 / Ctrl::qmin, q11 / Ctrl::gamma);
                    else {
                        dt = dtpropose;
                        if (!modified) {
#pragma unroll
                            for (int i = 0; i < I; ++i) kl[0][i] = kl[6][i];
#pragma unroll
                            for (int f = 0; f < SR::N; ++f) rec[f] = rec[6 * SR::N + f];
                        }
                    }
                }
                ++iter;
                tstop = (sp >= 0) ? fmax(a.saveat[sp], t0) : t0;
                const double dtmin_t = fmax(eps_of(t), dtmin0);
                dt = fmin(fmax(fmin(fabs(dt), dtmax), dtmin_t), t - tstop);
                if (iter > a.maxiters) { ret = RET_MAXITERS; go = 0; }
                else if (!(dt > dtmin_t) && (t - dt > tstop || !accept) && iter > 1) { ret = RET_DTMIN; go = 0; }
                else if (dt != dt) { ret = RET_UNSTABLE; go = 0; }
            }
            if (go) {
                // ---- stage evaluations ----
                h = (T)(-dt);
#pragma unroll
                for (int i = 0; i < I; ++i) lnew[i] = lprev[i];
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
                es_l = T(0);
#pragma unroll
                for (int i = 0; i < I; ++i) {
                    T ut = T(0);
#pragma unroll
                    for (int j = 0; j < 7; ++j) ut += Tab<T>::bt(j) * kl[j][i];
                    ut *= h;
                    const T sc = abstol + kmax(kabs(lprev[i]), kabs(lnew[i])) * reltol;
                    const T r = ut / sc;
                    es_l += r * r;
                    if (lnew[i] != lnew[i]) es_l = lnew[i];
                }
                s_mh = -h; s_cur = cur;
            }
        }
        go = __shfl_sync(0xffffffffu, go, 0);
        if (!go) break;
        __syncwarp();
        // ---- step-end pass over the gradient components: one work item per lane ----
        T es = T(0);
        if (lane < NITEM) {
            const int c = s_cur;
            es = gphase_item<NORM>(prm, lane, rec, 1, gbuf + (int64_t)c * NP * B, gbuf + (int64_t)(c ^ 1) * NP * B, B, s_mh, abstol, reltol);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) es += __shfl_xor_sync(0xffffffffu, es, off);
        __syncwarp();
        if (lane == 0) {
            // ---- loopfooter! ----
            const double EEst = (double)ksqrt((es_l + es) / T(NZ));
            if (EEst != EEst) { ret = RET_UNSTABLE; t = t0; accept = false; }
            else {
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
                    while (sp >= 0 && a.saveat[sp] == t) {
#pragma unroll
                        for (int i = 0; i < I; ++i) lam[i] += a.dg[((int64_t)sp * I + i) * B + b];
                        --sp; modified = true;
                    }
#pragma unroll
                    for (int i = 0; i < I; ++i) lprev[i] = lam[i];
                } else {
                    ++nreject;
                }
            }
        }
    }
    // result always in buffer 0; zero on failure
    cur = __shfl_sync(0xffffffffu, cur, 0);
    ret = __shfl_sync(0xffffffffu, ret, 0);
    __syncwarp();
    if (ret != RET_SUCCESS) { for (int j = lane; j < NP; j += 32) gbuf[(int64_t)j * B] = T(0); }
    else if (cur == 1) { for (int j = lane; j < NP; j += 32) gbuf[(int64_t)j * B] = gbuf[((int64_t)NP + j) * B]; }
    if (lane == 0) {
        if (a.du0)
#pragma unroll
            for (int i = 0; i < I; ++i) a.du0[b * I + i] = lam[i];
        if (a.stats) a.stats[b] = kanode_stats{naccept, nreject, nf, ret};
        if (a.attempts) {
            a.attempts[b] = naccept + nreject;
            if (a.attempts_sum) atomicAdd(a.attempts_sum, (unsigned long long)(naccept + nreject));
        }
    }
}

// out[j] = sum_b g[cur[b]][j][b] over the trajectories that finished successfully (double accumulation)
template <class T>
__global__ void __launch_bounds__(256) ls_reduce_kernel(const T* g, const int* cur, const int* ret, int NP, int64_t B, T* out) {
    const int j = blockIdx.x;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x)
        if (ret[i] == RET_SUCCESS) acc += (double)g[((int64_t)cur[i] * NP + j) * B + i];
    __shared__ double sh[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
        out[j] = (T)s;
    }
}

}  // namespace kanode
