// kanode_small_lg.cuh — LANE-GROUP backward (interpolating-adjoint) kernel for ensembles of small KAN-ODEs
// ([I,H,I], e.g. Lotka-Volterra [2,10,2] G=5).  Replaces the thread-per-trajectory adjoint kernel of round 1.
//
// Mapping.  A group of LPT = H/UPL lanes owns ONE trajectory (LV fp32: 5 lanes x 2 hidden units, 6 trajectories per
// warp; fp64: 10 lanes x 1 unit).  Everything that scales with the parameter count is distributed over the group and
// never moves:
//   * lane `lig` owns hidden units j = UPL*lig .. UPL*lig+UPL-1 and the 2*UPL*I*(G+1) gradient components that touch
//     them (layer 2: C2[(j,g),o], W2[j,o]; layer 1: C1[(i,g),j], W1[i,j]) — the gradient state g lives in REGISTERS
//     for the whole solve (240 values per trajectory = 48 per lane); nothing per-trajectory is kept in HBM;
//   * the rank-1 factors of dg/dt a stage produces for those components (RBF/SiLU features of the lane's own hidden
//     units, its hidden cotangents) are computed ONCE per stage evaluation by the lane that owns them and parked in a
//     lane-private shared-memory slice until the step-end pass consumes them: no recomputation of activations in the
//     step-end pass, no MUFU work done twice;
//   * what the whole group needs of a stage (dense-output state y(t_s), its 12 input features and their derivatives)
//     depends only on t_s, not on lambda, so all 7 stages are prepared up front, one stage per lane, and broadcast
//     through shared memory;
//   * the sequential part of a Runge-Kutta attempt (lambda stages, error norm of lambda, PI controller, accept/reject,
//     tstops, jumps) is replicated on the lanes of the group from bit-identical inputs, so they agree on every branch;
//     cross-lane sums (hidden -> input cotangent, error norm) are all-gathers by warp shuffle summed in a fixed order.
// The dense forward record is array-of-structures (one 80-byte record per accepted step) so a lane fetches a step with
// five 16-byte loads.  Per-warp gradient sums go to `gpart`; reduce_partials_kernel adds them in a fixed order in fp64.
//
// Reference semantics are those of kanode_small.cuh (same formulas; summation order differs at rounding level):
//   InterpolatingAdjoint backward solve on z=[lambda; g], tstops + jumps at the save times, FSAL re-evaluation after a
//   jump  [EXT SciMLSensitivity 7.69.0], triggered by Zygote.gradient(loss, p) at Lotka-Volterra/LV_driver_KANODE.jl:284
//   Tsit5 stages / error norm / PI controller / Hairer initial dt  [EXT OrdinaryDiffEqTsit5 1.1.0, OrdinaryDiffEqCore 1.9.0]
//   KDense reverse rules  Lotka-Volterra/src/utils.jl:15-21, kdense.jl:109-130
//
// dt-replay: with a.rp_t set, the controller is bypassed and the recorded accepted-step end times of
// another run (the fp64 oracle's) are replayed — separates arithmetic parity from step-size-control parity.
#pragma once
#include "kanode_small.cuh"

#ifndef KANODE_LG_WPB
#define KANODE_LG_WPB 4          // warps per block
#endif
#ifndef KANODE_LG_MINB
#define KANODE_LG_MINB 3         // resident blocks per SM the kernel is compiled for (fp32)
#endif

namespace kanode {

template <class T> struct LgBwdArgs {
    const T* wpk;            // packed weights in lane blocks (LgSmem::LW layout), TMA-staged to shared memory
    int64_t B;
    double t0, t1;
    const double* saveat;    // device, ascending
    int nsave;
    T abstol, reltol;
    int maxiters;
    const T* rec;            // [B][cap][RS] dense forward record (RecLayout)
    int cap;
    const int* nsteps;       // [B] accepted forward steps
    const int* retcode;      // [B] forward return codes
    const T* dg;             // [B][nsave][I]  dL/du(t_s)
    T* gpart;                // [warps][NP]    per-warp gradient sums
    T* du0;                  // [B][I] or null
    kanode_stats* stats;     // [B] or null
    int* attempts;           // [B] or null
    const double* rp_t;      // replay: [B][rp_cap] end times of the accepted backward steps (descending), NaN-padded; or null
    int rp_cap;
};

template <class P, int UPL_> struct LgGeom {
    static constexpr int I = P::I, H = P::H, G = P::G, UPL = UPL_;
    static_assert(H % UPL == 0, "hidden width must split evenly over the lanes of a group");
    static constexpr int LPT = H / UPL;                 // lanes per trajectory
    static_assert(LPT <= 32, "group wider than a warp");
    static constexpr int TPW = 32 / LPT;                // trajectories per warp
    static constexpr int NQ1 = I * (G + 1);             // input features of one stage
    static constexpr int SB = UPL * (G + 1);            // layer-2 factor block of one lane and one stage
    static constexpr int F1S = 2 * NQ1;                 // per stage: features then their input derivatives
    static constexpr int NC2 = I * SB, NC1 = UPL * NQ1; // gradient components per lane: layer 2, layer 1
    static_assert(NC2 + NC1 <= 7 * SB, "g scratch must fit into the consumed factor blocks");
    static constexpr int FACN = 7 * SB + 7 * UPL;       // factors per lane: c2[7][SB] then hb[7][UPL]
};

// shared-memory plan of one block, in units of T (every region a multiple of 16 bytes)
template <class T, class P, int UPL> struct LgSmem {
    using GM = LgGeom<P, UPL>;
    static constexpr int V = 16 / (int)sizeof(T);
    static constexpr int up(int x) { return (x + V - 1) / V * V; }
    // lane stride of the factor slices: padded so that the 16-byte accesses of a quarter warp hit distinct banks
    static constexpr int FACL = (sizeof(T) == 4) ? (up(GM::FACN) % 8 == 4 ? up(GM::FACN) : up(GM::FACN) + 4)
                                                 : (up(GM::FACN) % 4 == 2 ? up(GM::FACN) : up(GM::FACN) + 2);
    static constexpr int F1W = GM::TPW * 7 * GM::F1S;   // per warp: input features of the 7 stages of each trajectory
    static constexpr int LSW = up(GM::TPW * 7 * GM::I); // per warp: stage adjoints lambda_s of each trajectory
    static constexpr int PER_WARP = 32 * FACL + up(F1W) + LSW;
    // packed weights in LANE blocks: lane `lig` of a group reads [UPL][P::UW] at lig*LW; the pad makes the 16-byte
    // accesses of the lanes of a group hit distinct banks (upload_packed_lg builds the same image in global memory)
    static constexpr int LW = UPL * P::UW + V;
    static constexpr int WLG = GM::LPT * LW;
    static constexpr int BAROFF = up(WLG);              // mbarrier (16 bytes) behind the weights
    static constexpr int WOFF = BAROFF + V;
    static constexpr int F1OFF = 32 * FACL, LSOFF = F1OFF + up(F1W);   // offsets inside a warp's slice
    static constexpr size_t bytes(int warps) { return sizeof(T) * (size_t)(WOFF + warps * PER_WARP); }
};

template <class T, class P, int NORM, int UPL, int WPB, int MINB>
__global__ void __launch_bounds__(32 * WPB, MINB) small_backward_lg_kernel(const __grid_constant__ P prm, const LgBwdArgs<T> a) {
    using GM = LgGeom<P, UPL>;
    using SMP = LgSmem<T, P, UPL>;
    using RL = RecLayout<T, P::I>;
    constexpr int I = P::I, G = P::G, NP = P::NP, NZ = I + NP;
    constexpr int LPT = GM::LPT, TPW = GM::TPW, NQ1 = GM::NQ1, SB = GM::SB, F1S = GM::F1S, UW = P::UW;
    constexpr int V = RL::V, RS = RL::RS, FACL = SMP::FACL;
    static_assert(NQ1 % V == 0 && SB % V == 0 && UW % V == 0 && ((G + 1) * I) % V == 0, "vector widths");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* wsm = reinterpret_cast<T*>(smem_raw);
    uint64_t* wbar = reinterpret_cast<uint64_t*>(wsm + SMP::BAROFF);
    stage_weights<T, SMP::WLG>(wsm, wbar, a.wpk);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / LPT, lig = lane - grp * LPT, gbase = grp * LPT;
    T* wbase = wsm + SMP::WOFF + warp * SMP::PER_WARP;
    T* fac = wbase + lane * FACL;                                   // this lane's factors: c2[7][SB] | hb[7][UPL]
    const bool gvalid = grp < TPW;                                  // the 32 - TPW*LPT spare lanes only read
    const int gsl = gvalid ? grp : TPW - 1;
    T* f1g = wbase + SMP::F1OFF + gsl * 7 * F1S;                     // this trajectory's input features [7][F1S]
    T* lsg = wbase + SMP::LSOFF + gsl * 7 * I;   // its stage adjoints [7][I]
    const int j0 = UPL * lig;                                       // first hidden unit of this lane
    const T* wlane = wsm + lig * SMP::LW;                           // its packed weights [UPL][UW]

    const int64_t wg = (int64_t)blockIdx.x * WPB + warp;            // global warp index
    const int64_t b = wg * TPW + grp;
    const bool active = gvalid && b < a.B;
    const int64_t bq = active ? b : 0;                              // clamped: idle groups read trajectory 0, never write

    const int nsteps = active ? a.nsteps[bq] : 0;
    int ret = active ? a.retcode[bq] : RET_SUCCESS;
    bool done = !active || ret != RET_SUCCESS || nsteps <= 0;
    const bool skipped = active && done;                            // failed forward solve: zero gradient, statistics only

    // ---- state replicated on the lanes of a group ----
    T lam[I], lprev[I], kl[7][I];
#pragma unroll
    for (int i = 0; i < I; ++i) { lam[i] = T(0); lprev[i] = T(0); }
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int i = 0; i < I; ++i) kl[j][i] = T(0);                // zero-weighted stages must stay finite
    // ---- gradient state of this lane's components ----
    T g2[I][SB], g1[UPL][NQ1];
#pragma unroll
    for (int o = 0; o < I; ++o)
#pragma unroll
        for (int m = 0; m < SB; ++m) g2[o][m] = T(0);
#pragma unroll
    for (int u = 0; u < UPL; ++u)
#pragma unroll
        for (int m = 0; m < NQ1; ++m) g1[u][m] = T(0);

    int nf = 0, naccept = 0, nreject = 0;
    const double t0 = a.t0, t1 = a.t1, dtmax = fabs(t1 - t0);
    const double dtmin0 = fmax(eps_of(t0), eps_of(t1));
    const T abstol = a.abstol, reltol = a.reltol;
    const T* rbase = a.rec + bq * (int64_t)a.cap * RS;
    const T* dgb = a.dg + bq * (int64_t)a.nsave * I;
    int ridx = nsteps > 0 ? nsteps - 1 : 0;                         // per-lane hint into the dense record

    // ---- P1: y = sol(ts) and the input features of one stage slot, computed by ONE lane of the group ----
    auto prep_slot = [&](int slot, double ts) {
        double rt = rec_get_time(rbase + (int64_t)ridx * RS);
        while (ts < rt && ridx > 0) { --ridx; rt = rec_get_time(rbase + (int64_t)ridx * RS); }
        while (ridx + 1 < nsteps) {                                 // right-continuous at step boundaries
            const double rn = rec_get_time(rbase + (int64_t)(ridx + 1) * RS);
            if (!(ts >= rn)) break;
            ++ridx; rt = rn;
        }
        T r[RS];
        ldv(rbase + (int64_t)ridx * RS, r);
        const T rdt = r[RL::DT];
        T th;
        if constexpr (sizeof(T) == 4) th = (T)(ts - rt) / rdt; else th = (T)((ts - rt) / (double)rdt);
        T bw[7]; interp_weights(th, bw);
        T fd[F1S];                                                  // f[NQ1] then df[NQ1]
#pragma unroll
        for (int i = 0; i < I; ++i) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < 7; ++j) acc += bw[j] * r[RL::K + j * I + i];
            const T y = r[RL::U + i] + rdt * acc;
            const T xn = normalize<NORM>(y);
            const T dn = normalize_deriv<NORM>(xn);
            T rb[G], rdb[G];
            rbf_eval<true>(prm, xn, rb, rdb);
#pragma unroll
            for (int g = 0; g < G; ++g) { fd[i * G + g] = rb[g]; fd[NQ1 + i * G + g] = rdb[g] * dn; }   // utils.jl:18 * d(arg)/d(xn) * norm'
            swish_both(y, fd[I * G + i], fd[NQ1 + I * G + i]);
        }
        stv(f1g + slot * F1S, fd);
    };

    // ---- P2 + S: one adjoint RHS evaluation at stage slot `slot` with stage adjoint ls (all lanes of the group) ----
    //   dl = -(df/du)^T ls;  this lane's factors (features of its hidden units, their cotangents) -> fac, ls -> lsg
    auto stage_eval = [&](int slot, const T (&ls)[I], T (&dl)[I]) {
        T fd[F1S];
        ldv(f1g + slot * F1S, fd);
        T pu[I], c2[SB];
#pragma unroll
        for (int i = 0; i < I; ++i) pu[i] = T(0);
#pragma unroll
        for (int u = 0; u < UPL; ++u) {
            const T* w = wlane + u * UW;
            T w1[NQ1];
            ldv(w, w1);
            T h0 = T(0), h1 = T(0);
#pragma unroll
            for (int q = 0; q + 1 < NQ1; q += 2) kfma2(h0, h1, w1[q], w1[q + 1], fd[q], fd[q + 1]);
            if constexpr (NQ1 % 2 == 1) h0 += w1[NQ1 - 1] * fd[NQ1 - 1];
            const T h = h0 + h1;
            T dh[I];                                                // d h / d y_i
#pragma unroll
            for (int i = 0; i < I; ++i) {
                T acc = w1[I * G + i] * fd[NQ1 + I * G + i];
#pragma unroll
                for (int g = 0; g < G; ++g) acc += w1[i * G + g] * fd[NQ1 + i * G + g];
                dh[i] = acc;
            }
            const T xn = normalize<NORM>(h);
            const T dn = normalize_deriv<NORM>(xn);
            T rb[G], rdb[G];
            rbf_eval<true>(prm, xn, rb, rdb);
            T s, ds; swish_both(h, s, ds);
#pragma unroll
            for (int g = 0; g < G; ++g) c2[u * (G + 1) + g] = rb[g];
            c2[u * (G + 1) + G] = s;
            T w2[(G + 1) * I];
            ldv(w + NQ1, w2);
            T hb = T(0);
#pragma unroll
            for (int o = 0; o < I; ++o) {                           // J2[o] = d f_o / d h_j; hb = sum_o ls[o] * J2[o]
                T acc = T(0);
#pragma unroll
                for (int g = 0; g < G; ++g) acc += w2[g * I + o] * rdb[g];
                const T j2 = acc * dn + w2[G * I + o] * ds;
                hb += ls[o] * j2;
            }
            fac[7 * SB + slot * UPL + u] = hb;
#pragma unroll
            for (int i = 0; i < I; ++i) pu[i] += hb * dh[i];
        }
        stv(fac + slot * SB, c2);
#pragma unroll
        for (int i = 0; i < I; ++i) {                               // all-gather over the group, summed in lane order
            T tot = T(0);
#pragma unroll
            for (int k = 0; k < LPT; ++k) tot += shfl_t(pu[i], gbase + k);
            dl[i] = -tot;
        }
        if (lig == 0 && gvalid) {
#pragma unroll
            for (int i = 0; i < I; ++i) lsg[slot * I + i] = ls[i];
        }
    };
    auto group_sum = [&](T v) {
        T tot = T(0);
#pragma unroll
        for (int k = 0; k < LPT; ++k) tot += shfl_t(v, gbase + k);
        return tot;
    };

    double t = t1;
    int sp = a.nsave - 1;                                           // next preset (save) time, descending
    auto apply_jumps = [&](double tt) {
        bool mod = false;
        while (sp >= 0 && a.saveat[sp] == tt) {
#pragma unroll
            for (int i = 0; i < I; ++i) lam[i] += dgb[sp * I + i];
            --sp; mod = true;
        }
        return mod;
    };
    apply_jumps(t1);                                                // PresetTimeCallback fires at init when t_end is a save time
#pragma unroll
    for (int i = 0; i < I; ++i) lprev[i] = lam[i];

    double dt;                                                      // |dt|; integration runs in -t
    {   // ---- FSAL evaluation + ode_determine_initdt on the augmented state (g(T) = 0: its scale is abstol) ----
        if (lig == 0 && gvalid) prep_slot(0, t1);
        __syncwarp();
        stage_eval(0, lam, kl[0]);
        __syncwarp();
        T sk[I], s0 = T(0), s1 = T(0);
#pragma unroll
        for (int i = 0; i < I; ++i) {
            sk[i] = abstol + kabs(lam[i]) * reltol;
            const T x0 = lam[i] / sk[i], x1 = kl[0][i] / sk[i];
            s0 += x0 * x0; s1 += x1 * x1;
        }
        // this lane's gradient components of stage slot 0 (kv0) and, later, slot 1 (kv1)
        T ga = T(0);
        {
            T c0[SB], f0[NQ1];
            ldv(fac, c0); ldv(f1g, f0);
#pragma unroll
            for (int o = 0; o < I; ++o)
#pragma unroll
                for (int m = 0; m < SB; ++m) { const T x = (lam[o] * c0[m]) / abstol; ga += x * x; }
#pragma unroll
            for (int u = 0; u < UPL; ++u) {
                const T hb0 = fac[7 * SB + u];
#pragma unroll
                for (int m = 0; m < NQ1; ++m) { const T x = (hb0 * f0[m]) / abstol; ga += x * x; }
            }
        }
        s1 += group_sum(ga);
        const double d0 = sqrt((double)s0 / NZ), d1 = sqrt((double)s1 / NZ);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
        dt0 = fmin(dt0, dtmax);
        T l1[I], fv[I];
#pragma unroll
        for (int i = 0; i < I; ++i) l1[i] = lam[i] - (T)dt0 * kl[0][i];
        if (lig == 0 && gvalid) prep_slot(1, t1 - dt0);
        __syncwarp();
        stage_eval(1, l1, fv);
        __syncwarp();
        nf = 3;                                                     // two evaluations + the package's repeated f0
        T s2 = T(0);
#pragma unroll
        for (int i = 0; i < I; ++i) { const T x = (fv[i] - kl[0][i]) / sk[i]; s2 += x * x; }
        T gb = T(0);
        {
            T c0[SB], c1[SB], f0[NQ1], f1[NQ1];
            ldv(fac, c0); ldv(fac + SB, c1); ldv(f1g, f0); ldv(f1g + F1S, f1);
#pragma unroll
            for (int o = 0; o < I; ++o)
#pragma unroll
                for (int m = 0; m < SB; ++m) { const T x = (l1[o] * c1[m] - lam[o] * c0[m]) / abstol; gb += x * x; }
#pragma unroll
            for (int u = 0; u < UPL; ++u) {
                const T hb0 = fac[7 * SB + u], hb1 = fac[7 * SB + UPL + u];
#pragma unroll
                for (int m = 0; m < NQ1; ++m) { const T x = (hb1 * f1[m] - hb0 * f0[m]) / abstol; gb += x * x; }
            }
        }
        s2 += group_sum(gb);
        const double d2 = sqrt((double)s2 / NZ) / dt0;
        const double mx = fmax(d1, d2);
        const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        dt = fmax(dtmin0, fmin(fmin(100.0 * dt0, dt1), dtmax));
    }
    double qold = Ctrl::qoldinit, q11 = 1.0, dtpropose = dt;
    bool accept = false, modified = false;
    int iter = 0;
    if (!(t > t0)) done = true;
    const double* rp = a.rp_t ? a.rp_t + bq * (int64_t)a.rp_cap : nullptr;

    while (__any_sync(0xffffffffu, !done)) {
        // ---- loopheader! ----
        if (iter > 0) {
            if (!accept) dt = dt / fmin(1.0 / Ctrl::qmin, q11 / Ctrl::gamma);
            else dt = dtpropose;
        }
        ++iter;
        const double tstop = (sp >= 0) ? fmax(a.saveat[sp], t0) : t0;
        const double dtmin_t = fmax(eps_of(t), dtmin0);
        dt = fmin(fmax(fmin(fabs(dt), dtmax), dtmin_t), t - tstop);
        double rp_next = t0;
        if (rp) {                                                   // replay: the step ends where the recorded one ended
            rp_next = naccept < a.rp_cap ? rp[naccept] : t0;
            if (!(rp_next < t) || !(rp_next >= t0)) rp_next = tstop;
            dt = t - rp_next;
        }
        if (!done) {
            if (iter > a.maxiters) { ret = RET_MAXITERS; done = true; }
            else if (!rp && !(dt > dtmin_t) && (t - dt > tstop || !accept) && iter > 1) { ret = RET_DTMIN; done = true; }
            else if (dt != dt) { ret = RET_UNSTABLE; done = true; }
        }
        // ---- P1: stage states and input features of the 7 stages, one stage per lane ----
#pragma unroll
        for (int r0 = 0; r0 < 7; r0 += LPT) {
            const int slot = r0 + lig;
            if (slot < 7 && gvalid) prep_slot(slot, t - tab_c(slot) * dt);
        }
        __syncwarp();
        // ---- perform_step! on lambda; after a jump the FSAL stage is re-evaluated (stage 0 of the same loop; without a
        // jump the recomputed stage 0 equals the FSAL value, the evaluation count follows the package) ----
        const T h = (T)(-dt);
        T lnew[I];
#pragma unroll
        for (int i = 0; i < I; ++i) lnew[i] = lprev[i];
#pragma unroll 1
        for (int s = 0; s < 7; ++s) {
            T ls[I], ks[I];
#pragma unroll
            for (int i = 0; i < I; ++i) {
                T acc = T(0);
#pragma unroll
                for (int j = 0; j < 6; ++j) acc += Tab<T>::a(s, j) * kl[j][i];
                ls[i] = lprev[i] + h * acc;
            }
            stage_eval(s, ls, ks);
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
        if (!done) nf += modified ? 7 : 6;
        modified = false;
        __syncwarp();
        // ---- error estimate: lambda part (replicated) ----
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
        // ---- step-end pass over this lane's gradient components: g_new = g + (-h) sum_s b_s kv_s, error term with btilde.
        // kv_s is rank one: (stage adjoint or hidden cotangent of stage s) x (feature of stage s); the features are the
        // ones the stage evaluations parked in shared memory.  g_new goes to the already consumed factor blocks. ----
        T esl = T(0);
        const T mh = -h;
        auto fin = [&](T gold, T vb, T vt) {
            const T gn = gold + vb;
            const T sc = abstol + kmax(kabs(gold), kabs(gn)) * reltol;
            const T r = kdiv(vt, sc);
            esl += r * r;
            return gn;
        };
        {   // layer 2: component (o, m): kv_s = lambda_s[o] * c2_s[m]
            T ab[7][I], at[7][I];
#pragma unroll
            for (int s = 0; s < 7; ++s)
#pragma unroll
                for (int o = 0; o < I; ++o) {
                    const T l = lsg[s * I + o];
                    ab[s][o] = (mh * Tab<T>::b(s)) * l; at[s][o] = (mh * Tab<T>::bt(s)) * l;
                }
#pragma unroll
            for (int m0 = 0; m0 < SB; m0 += V) {
                T c[7][V];
#pragma unroll
                for (int s = 0; s < 7; ++s) ldv(fac + s * SB + m0, c[s]);
                T vb[I][V], vt[I][V];
#pragma unroll
                for (int o = 0; o < I; ++o)
#pragma unroll
                    for (int e = 0; e < V; ++e) { vb[o][e] = T(0); vt[o][e] = T(0); }
#pragma unroll
                for (int s = 0; s < 7; ++s)
#pragma unroll
                    for (int o = 0; o < I; ++o)
#pragma unroll
                        for (int e = 0; e < V; ++e) kfma2b(vb[o][e], vt[o][e], ab[s][o], at[s][o], c[s][e]);
#pragma unroll
                for (int o = 0; o < I; ++o) {
                    T gn[V];
#pragma unroll
                    for (int e = 0; e < V; ++e) gn[e] = fin(g2[o][m0 + e], vb[o][e], vt[o][e]);
                    stv(fac + o * SB + m0, gn);
                }
            }
        }
        {   // layer 1: component (u, m): kv_s = hbar_s[j0+u] * f_s[m]
            T ab[7][UPL], at[7][UPL];
#pragma unroll
            for (int s = 0; s < 7; ++s)
#pragma unroll
                for (int u = 0; u < UPL; ++u) {
                    const T hb = fac[7 * SB + s * UPL + u];
                    ab[s][u] = (mh * Tab<T>::b(s)) * hb; at[s][u] = (mh * Tab<T>::bt(s)) * hb;
                }
#pragma unroll
            for (int m0 = 0; m0 < NQ1; m0 += V) {
                T c[7][V];
#pragma unroll
                for (int s = 0; s < 7; ++s) ldv(f1g + s * F1S + m0, c[s]);
                T vb[UPL][V], vt[UPL][V];
#pragma unroll
                for (int u = 0; u < UPL; ++u)
#pragma unroll
                    for (int e = 0; e < V; ++e) { vb[u][e] = T(0); vt[u][e] = T(0); }
#pragma unroll
                for (int s = 0; s < 7; ++s)
#pragma unroll
                    for (int u = 0; u < UPL; ++u)
#pragma unroll
                        for (int e = 0; e < V; ++e) kfma2b(vb[u][e], vt[u][e], ab[s][u], at[s][u], c[s][e]);
#pragma unroll
                for (int u = 0; u < UPL; ++u) {
                    T gn[V];
#pragma unroll
                    for (int e = 0; e < V; ++e) gn[e] = fin(g1[u][m0 + e], vb[u][e], vt[u][e]);
                    stv(fac + GM::NC2 + u * NQ1 + m0, gn);
                }
            }
        }
        es += group_sum(esl);
        const double EEst = (double)ksqrt(es / T(NZ));
        if (!done && (EEst != EEst || bad)) { ret = RET_UNSTABLE; done = true; }
        // ---- loopfooter!: PI controller ----
        const double q = pi_q(EEst, qold, q11);
        accept = rp ? true : (EEst <= 1.0);
        if (!done) {
            if (accept) {
                ++naccept;
                qold = fmax(EEst, Ctrl::qoldinit);
                const double dtnew = dt / q;
                double tnew = t - dt;
                if (rp) tnew = rp_next;
                else if (fabs(tnew - tstop) < 100.0 * eps_of(fmax(fabs(t), fabs(tstop)))) tnew = tstop;
                dtpropose = fmax(fmin(dtmax, fabs(dtnew)), fmax(eps_of(tnew), dtmin0));
                t = tnew;
#pragma unroll
                for (int o = 0; o < I; ++o) ldv(fac + o * SB, g2[o]);           // commit g_new
#pragma unroll
                for (int u = 0; u < UPL; ++u) ldv(fac + GM::NC2 + u * NQ1, g1[u]);
#pragma unroll
                for (int i = 0; i < I; ++i) lam[i] = lnew[i];
                modified = apply_jumps(t);
#pragma unroll
                for (int i = 0; i < I; ++i) lprev[i] = lam[i];
                if (!(t > t0)) done = true;
            } else {
                ++nreject;
            }
        }
        __syncwarp();                                               // factor / feature slices are rewritten by the next attempt
    }

    // ---- results: per-warp gradient sum (fixed order over the trajectories of the warp), du0, statistics ----
    const bool ok = active && !skipped && ret == RET_SUCCESS;
    T* gp = a.gpart + wg * NP;
#pragma unroll
    for (int o = 0; o < I; ++o)
#pragma unroll
        for (int m = 0; m < SB; ++m) {
            const T v = ok ? g2[o][m] : T(0);
            T tot = T(0);
#pragma unroll
            for (int k = 0; k < TPW; ++k) tot += shfl_t(v, lig + k * LPT);
            if (lane < LPT) {
                const int u = m / (G + 1), q = m % (G + 1), j = j0 + u;
                gp[q < G ? P::OC2 + (j * G + q) * I + o : P::OW2 + j * I + o] = tot;
            }
        }
#pragma unroll
    for (int u = 0; u < UPL; ++u)
#pragma unroll
        for (int m = 0; m < NQ1; ++m) {
            const T v = ok ? g1[u][m] : T(0);
            T tot = T(0);
#pragma unroll
            for (int k = 0; k < TPW; ++k) tot += shfl_t(v, lig + k * LPT);
            if (lane < LPT) gp[P::OC1 + m * P::H + j0 + u] = tot;   // C1 then W1 are contiguous: column m of [H x NQ1]
        }
    if (active && lig == 0) {
        if (skipped) { nf = 0; naccept = 0; nreject = 0; }
        if (a.du0)
#pragma unroll
            for (int i = 0; i < I; ++i) a.du0[b * I + i] = skipped ? T(0) : lam[i];
        if (a.stats) a.stats[b] = kanode_stats{naccept, nreject, nf, ret};
        if (a.attempts) a.attempts[b] = naccept + nreject;
    }
}

// out[c] = scale * sum_r part[r][c]: rows are per-warp gradient sums.  Two deterministic stages: slab sums in fp64
// (grid.x slabs, thread per column, coalesced rows), then the sum over slabs.
template <class T>
__global__ void __launch_bounds__(256) reduce_partials_kernel(const T* __restrict__ part, int64_t rows, int cols, double* __restrict__ slab) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
    double acc = 0.0;
    for (int64_t r = r0; r < r1; ++r) acc += (double)part[r * cols + c];
    slab[(int64_t)blockIdx.x * cols + c] = acc;
}
template <class OutT>
__global__ void __launch_bounds__(256) reduce_slabs_kernel(const double* __restrict__ slab, int nslab, int cols, OutT* __restrict__ out, double scale) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double acc = 0.0;
    for (int s = 0; s < nslab; ++s) acc += slab[(int64_t)s * cols + c];
    out[c] = (OutT)(acc * scale);
}

}  // namespace kanode
