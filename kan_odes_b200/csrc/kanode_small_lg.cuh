// kanode_small_lg.cuh — LANE-GROUP backward (interpolating-adjoint) kernel for ensembles of small KAN-ODEs
// ([I,H,I], e.g. Lotka-Volterra [2,10,2] G=5).  Replaces the thread-per-trajectory adjoint kernel of round 1.
//
// Mapping.  A group of LPT = H/UPL lanes owns ONE trajectory (LV fp32: 5 lanes x 2 hidden units, 6 trajectories per
// warp; fp64: 10 lanes x 1 unit).  Everything that scales with the parameter count is distributed over the group and
// never moves:
//   * lane `lig` owns hidden units j = UPL*lig .. UPL*lig+UPL-1 and the 2*UPL*I*(G+1) gradient components that touch
//     them (layer 2: C2[(j,g),o], W2[j,o]; layer 1: C1[(i,g),j], W1[i,j]) — the gradient state g lives in REGISTERS
//     for the whole solve (240 values per trajectory = 48 per lane); nothing per-trajectory is kept in HBM;
//   * an adjoint RHS evaluation splits into a lambda-INDEPENDENT part A (hidden pre-activations of y(t_s), RBF/SiLU
//     features c2 of the lane's hidden units, the local Jacobian pieces J2 = d f_o / d h_j and dh = d h_j / d y_i) and a
//     tiny lambda-dependent part B (hbar_j = sum_o ls_o J2[o][j], dl = -sum_j hbar_j dh_j, summed over the group by
//     shuffles).  The stage loop is software-pipelined: A of stage s+1 is issued next to the serial chain B of stage s;
//   * Tsit5 has c6 = c7 = 1 and the first stage time of an accepted step is the last one of the step before, so an
//     attempt has only FIVE new stage times: part A (and the input features, one stage time per lane) is evaluated
//     5 times per attempt instead of 7, and the step-end pass runs over 6 feature blocks with merged weights;
//   * the rank-1 factors of dg/dt (features c2 in a lane-private shared-memory slice, hbar, stage adjoints) are parked
//     in shared memory until the step-end pass consumes them: no recomputation of activations, no MUFU work twice;
//   * the two dense-forward records around t live in shared memory (refilled one attempt ahead of use), the next
//     lambda jump (save time, dL/du) is prefetched into registers: no global-memory latency inside an attempt;
//   * the sequential part of a Runge-Kutta attempt (lambda stages, error norm of lambda, PI controller, accept/reject,
//     tstops, jumps) is replicated on the lanes of the group from bit-identical inputs, so they agree on every branch;
//     cross-lane sums (hidden -> input cotangent, error norm) are all-gathers by warp shuffle summed in a fixed order.
// The dense forward record is array-of-structures (one 80-byte record per accepted step).  Per-warp gradient sums go to
// `gpart`; reduce_partials_kernel adds them in a fixed order in fp64.
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

// fp32 only (the fp64 instantiation keeps the oracle's operation order): the 5-lane sums of a group as a fixed tree
// ((a0+a1)+(a2+a3))+a4 instead of a chain, and ONE MUFU.RCP per pair of gradient components in the step-end error norm
// (r0 = vt0 sc1 / (sc0 sc1), r1 = vt1 sc0 / (sc0 sc1); the scales are >= abstol, their product is far from under/overflow).
// Measured together on B200: 248 registers instead of 255 (no spill), backward 1.962 -> 1.930 ms; each alone: 1.953 / 1.968.
#ifndef KANODE_LG_TREESUM
#define KANODE_LG_TREESUM 1
#endif
#ifndef KANODE_LG_RCP1
#define KANODE_LG_RCP1 1
#endif
#ifndef KANODE_LG_PAIRC
#define KANODE_LG_PAIRC 1        // step-end pass: FFMA2 over pairs of components (1) instead of (g_new, error) pairs (0): the
                                 // finish pairs components, so (0) re-pairs ~100 registers per attempt with MOVs; 1.934 -> 1.914 ms
#endif
#ifndef KANODE_LG_STAGE_UNROLL
#define KANODE_LG_STAGE_UNROLL 1 // unroll factor of the pipelined stage loop (5 = fully unrolled)
#endif
#define KANODE_LG_PRAGMA(x) _Pragma(#x)
#define KANODE_LG_UNROLL(n) KANODE_LG_PRAGMA(unroll n)
#ifndef KANODE_LG_WPB
#define KANODE_LG_WPB 4          // warps per block
#endif
#ifndef KANODE_LG_MINB
#define KANODE_LG_MINB 2         // resident blocks per SM the kernel is compiled for (fp32)
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
    // launch order (optional): position -> trajectory (position = warp position * TPW + group).  A trajectory needs many extra
    // attempts when its error-controlled step falls below the spacing of the save times: a clipped remainder follows every
    // proposed step and the solve takes 50-120 attempts instead of ~40.  Which trajectories do so cannot be known in advance
    // and changes with the parameters; what IS stable from one training step to the next is how far a trajectory is from
    // that regime: its step MARGIN = min over attempts of (proposed dt) / (distance to the next save time), > 1 for an
    // unremarkable solve.  Trajectories are started in ascending order of the previous call's margin: the ones at risk first,
    // the certainly harmless ones last, so the serial chain of a long solve does not end up as the tail of the launch.
    // gpart is indexed by the warp POSITION; the order is a stable (deterministic) function of the margins.
    const int* order;        // [B] permutation of the trajectories, or null (identity)
    int* tmargin;            // [B] bit pattern of the (positive) fp32 margin of each trajectory (feeds the next call's order), or null
    // persistent launch: the grid holds only the resident warps (blocks per SM x SMs); every warp draws its next position from
    // ticket[0] until nwarps are handed out, so a warp that ends early does not wait for the slowest warp of its block and no
    // wave of blocks is left half empty.  ticket[1] counts retired warps: the last one zeroes both for the next launch.
    int* ticket;
    int64_t nwarps;          // warp positions (= ceil(B / TPW))
};

template <class T, class P, int UPL_> struct LgGeom {
    static constexpr int I = P::I, H = P::H, G = P::G, UPL = UPL_;
    static_assert(H % UPL == 0, "hidden width must split evenly over the lanes of a group");
    static constexpr int LPT = H / UPL;                 // lanes per trajectory
    static_assert(LPT <= 32, "group wider than a warp");
    static constexpr int TPW = 32 / LPT;                // trajectories per warp
    static constexpr int NB = 6;                        // distinct stage times of a Tsit5 attempt (c6 = c7 = 1)
    static constexpr int NQ1 = I * (G + 1);             // input features of one stage time, index q*I + i (q = G: SiLU)
    static constexpr int SBR = UPL * (G + 1);           // layer-2 factors of one lane and one stage time ...
    static constexpr int VT = 16 / (int)sizeof(T);
    static constexpr int SB = (SBR + VT - 1) / VT * VT; // ... padded to whole 16-byte vectors (pad entries stay zero)
    static constexpr int F1S = 2 * NQ1;                 // per stage time: features then their input derivatives
    static constexpr int NC2 = I * SB, NC1 = UPL * NQ1; // gradient components per lane: layer 2, layer 1
    static_assert((1 + I) * SB + NC1 <= (NB - 1) * SB, "g scratch must fit into the factor blocks 1..NB-2");
    static constexpr int FACN = NB * SB + 7 * UPL;      // factors per lane: c2[NB][SB] then hb[7][UPL]
};

// weight image of one hidden unit (UW values, SmallParams::UW): [w1: q*I + i (q < G: C1[(i,q),j]; q = G: W1[i,j])] then
// [w2: g*I + o (C2[(j,g),o]; g = G: W2[j,o])]

// shared-memory plan of one block, in units of T (every region a multiple of 16 bytes)
template <class T, class P, int UPL> struct LgSmem {
    using GM = LgGeom<T, P, UPL>;
    using RL = RecLayout<T, P::I>;
    static constexpr int V = 16 / (int)sizeof(T);
    static constexpr int up(int x) { return (x + V - 1) / V * V; }
    // lane stride of the factor slices: padded so that the 16-byte accesses of a quarter warp hit distinct banks
    static constexpr int FACL = (sizeof(T) == 4) ? (up(GM::FACN) % 8 == 4 ? up(GM::FACN) : up(GM::FACN) + 4)
                                                 : (up(GM::FACN) % 4 == 2 ? up(GM::FACN) : up(GM::FACN) + 2);
    static constexpr int HBN = up(7 * UPL);             // vector read of the 7 hidden cotangents of a lane
    static_assert(GM::NB * GM::SB + HBN <= FACL, "hbar vector read stays inside the lane slice");
    static constexpr int F1W = GM::TPW * GM::NB * GM::F1S;   // per warp: input features of the stage times of each trajectory
    static constexpr int LSS = up(7 * GM::I);           // per trajectory: stage adjoints lambda_s
    static constexpr int LSW = GM::TPW * LSS;
    static constexpr int RCW = GM::TPW * 2 * RL::RS;    // per trajectory: the two dense-forward records around t
    // per trajectory: two mailbox slots [save time (fp64) | dL/du(t_s) (I)] for the next two jumps (slot = save index & 1)
    static constexpr int MBS = up(RL::OT + GM::I);
    static constexpr int MBW = GM::TPW * 2 * MBS;
    static constexpr int PER_WARP = 32 * FACL + up(F1W) + LSW + RCW + MBW;
    // packed weights in LANE blocks: lane `lig` of a group reads [UPL][P::UW] at lig*LW; the pad makes the 16-byte
    // accesses of the lanes of a group hit distinct banks (upload_packed_lg builds the same image in global memory)
    static constexpr int LW = UPL * P::UW + V;
    static constexpr int WLG = GM::LPT * LW;
    static constexpr int BAROFF = up(WLG);              // mbarrier (16 bytes) behind the weights
    static constexpr int WOFF = BAROFF + V;
    static constexpr int F1OFF = 32 * FACL, LSOFF = F1OFF + up(F1W), RCOFF = LSOFF + LSW, MBOFF = RCOFF + RCW;   // offsets inside a warp's slice
    static constexpr size_t bytes(int warps) { return sizeof(T) * (size_t)(WOFF + warps * PER_WARP); }
};

// error-norm ratio a/b: fp32 = one MUFU.RCP and one FMUL (b >= abstol > 0, far from the denormal range)
__device__ __forceinline__ float kratio(float a, float b) { return a * krcp(b); }
__device__ __forceinline__ double kratio(double a, double b) { return a / b; }
// packed fp32 add (FADD2): d0 = a0 + b0, d1 = a1 + b1
__device__ __forceinline__ void kadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(rd));
}
template <class T> __device__ __forceinline__ void kadd2(T& d0, T& d1, T a0, T a1, T b0, T b1) { d0 = a0 + b0; d1 = a1 + b1; }

// asynchronous global -> shared copies (LDGSTS): no destination register, no scoreboard; completion by wait_all
template <int BYTES> __device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// PI controller of the lane-group kernel.  fp64 follows pi_q (kanode_math.cuh) to the letter: its step sequence equals the
// oracle's.  fp32 evaluates the same formulas with MUFU reciprocals / ex2 and no double-precision division: the fp32 solve
// cannot reproduce fp64 step sequences anyway, a relative change of 1e-7 in dt is far inside that.
//   q11 = EEst^beta1,  q = clamp((q11 / qold^beta2) / gamma, 1/qmax, 1/qmin)   [EXT OrdinaryDiffEqCore 1.9.0, FastPower 1.1.0]
__device__ __forceinline__ float fastlog2f_rcp(float x) {
    const uint32_t bits = __float_as_uint(x);
    const float e = (float)((bits & 0x7F800000u) >> 23);
    float s, fe;
    if (bits & 0x00400000u) { s = __uint_as_float((bits & 0x007FFFFFu) | 0x3f000000u) - 1.0f; fe = e - 126.0f; }
    else                    { s = __uint_as_float((bits & 0x007FFFFFu) | 0x3f800000u) - 1.0f; fe = e - 127.0f; }
    return fe + s * (0.338953f * s + 2.198599f) * krcp(s + 1.523692f);
}
template <class T> struct LgCtrl;
template <> struct LgCtrl<double> {
    using Q = double;                                  // type of qold / q11
    static __device__ __forceinline__ double eest(double es, int nz) { return sqrt(es / (double)nz); }
    static __device__ __forceinline__ double q(double EEst, double qold, double& q11) { return pi_q(EEst, qold, q11); }
    static __device__ __forceinline__ double grow(double dt, double q) { return dt / q; }                         // dt / q
    static __device__ __forceinline__ double shrink(double dt, double q11) { return dt / fmin(1.0 / Ctrl::qmin, q11 / Ctrl::gamma); }
    static __device__ __forceinline__ double qold_next(double EEst) { return fmax(EEst, Ctrl::qoldinit); }
};
template <> struct LgCtrl<float> {
    using Q = float;
    static __device__ __forceinline__ float eest(float es, int nz) {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(es * (1.0f / (float)nz)));
        return r;
    }
    static __device__ __forceinline__ float q(float EEst, float qold, float& q11) {
        if (EEst == 0.0f) return (float)(1.0 / Ctrl::qmax);
        q11 = kex2(Ctrl::beta1 * fastlog2f_rcp(EEst));
        const float qq = q11 * krcp(kex2(Ctrl::beta2 * fastlog2f_rcp(qold)));
        return fmaxf((float)(1.0 / Ctrl::qmax), fminf((float)(1.0 / Ctrl::qmin), qq * (float)(1.0 / Ctrl::gamma)));
    }
    static __device__ __forceinline__ double grow(double dt, float q) { return dt * (double)krcp(q); }
    static __device__ __forceinline__ double shrink(double dt, float q11) {
        return dt * (double)krcp(fminf((float)(1.0 / Ctrl::qmin), q11 * (float)(1.0 / Ctrl::gamma)));
    }
    static __device__ __forceinline__ float qold_next(float EEst) { return fmaxf(EEst, (float)Ctrl::qoldinit); }
};

// Step-end weights of a per-stage quantity x[s*stride + off]: feature blocks 0..NB-2 carry stage k, block NB-1 carries the
// two stages that share the last stage time.  cb = (-h) b_s x_s, ct = (-h) btilde_s x_s.
template <class T, int NB, int N>
__device__ __forceinline__ void lg_merge(const T (&x)[N], int stride, int off, T mh, T (&cb)[NB], T (&ct)[NB]) {
#pragma unroll
    for (int k = 0; k < NB - 1; ++k) { const T v = x[k * stride + off]; cb[k] = (mh * Tab<T>::b(k)) * v; ct[k] = (mh * Tab<T>::bt(k)) * v; }
    const T v5 = x[(NB - 1) * stride + off], v6 = x[NB * stride + off];
    cb[NB - 1] = mh * (Tab<T>::b(NB - 1) * v5 + Tab<T>::b(NB) * v6);
    ct[NB - 1] = mh * (Tab<T>::bt(NB - 1) * v5 + Tab<T>::bt(NB) * v6);
}

template <class T, class P, int NORM, int UPL, int WPB>
__device__ __forceinline__ void small_backward_lg_body(const P& prm, const LgBwdArgs<T>& a) {
    using GM = LgGeom<T, P, UPL>;
    using SMP = LgSmem<T, P, UPL>;
    using RL = RecLayout<T, P::I>;
    constexpr int I = P::I, G = P::G, NP = P::NP, NZ = I + NP;
    constexpr int LPT = GM::LPT, TPW = GM::TPW, NQ1 = GM::NQ1, SB = GM::SB, SBR = GM::SBR, F1S = GM::F1S, UW = P::UW, NB = GM::NB;
    constexpr int V = RL::V, RS = RL::RS, FACL = SMP::FACL, NPIECE = RS / V, LSS = SMP::LSS, HBN = SMP::HBN;
    constexpr int NW2 = (G + 1) * I;
    static_assert(NQ1 % V == 0 && SB % V == 0 && UW % V == 0 && NW2 % V == 0, "vector widths");
    static_assert(NPIECE <= LPT && NQ1 / V <= LPT, "one 16-byte piece of a record / feature block per lane");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* wsm = reinterpret_cast<T*>(smem_raw);
    uint64_t* wbar = reinterpret_cast<uint64_t*>(wsm + SMP::BAROFF);
    stage_weights<T, SMP::WLG>(wsm, wbar, a.wpk);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / LPT, lig = lane - grp * LPT, gbase = grp * LPT;
    T* wbase = wsm + SMP::WOFF + warp * SMP::PER_WARP;
    T* fac = wbase + lane * FACL;                                   // this lane's factors: c2[NB][SB] | hb[7][UPL]
    const bool gvalid = grp < TPW;                                  // the 32 - TPW*LPT spare lanes only read
    const int gsl = gvalid ? grp : TPW - 1;
    T* f1g = wbase + SMP::F1OFF + gsl * NB * F1S;                    // this trajectory's input features [NB][F1S]
    T* lsg = wbase + SMP::LSOFF + gsl * LSS;                         // its stage adjoints [7][I]
    T* rcg = wbase + SMP::RCOFF + gsl * 2 * RS;                      // its record window: record r sits in slot r & 1
    T* mbg = wbase + SMP::MBOFF + gsl * 2 * SMP::MBS;                // its jump mailbox: save index s sits in slot s & 1
    const int j0 = UPL * lig;                                       // first hidden unit of this lane
    const T* wlane = wsm + lig * SMP::LW;                           // its packed weights [UPL][UW]

    for (;;) {
    int tk = 0;
    if (lane == 0) tk = atomicAdd(a.ticket, 1);
    const int wslot = __shfl_sync(0xffffffffu, tk, 0);              // position of this pass of the warp
    if (wslot >= a.nwarps) break;
    const bool active = gvalid && (int64_t)wslot * TPW + grp < a.B;
    auto trajectory = [&]() -> int64_t {                            // the trajectory at this position (re-read at the end: not kept live)
        const int64_t pos = (int64_t)wslot * TPW + grp;
        return (active && a.order) ? (int64_t)a.order[pos] : pos;
    };
    const int64_t bq = active ? trajectory() : 0;                   // clamped: idle groups read trajectory 0, never write

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
    const T inv_abstol = T(1) / abstol;
    auto by_abstol = [&](T x) { if constexpr (sizeof(T) == 4) return x * inv_abstol; else return x / abstol; };
    const T* rbase = a.rec + bq * (int64_t)a.cap * RS;
    const T* dgb = a.dg + bq * (int64_t)a.nsave * I;

    // ---- window on the dense forward record: records ridx (holds t) and ridx-1 in shared memory ----
    const bool haspiece = gvalid && lig < NPIECE;
    int ridx = nsteps > 0 ? nsteps - 1 : 0;
    double rt0 = rec_get_time(rbase + (int64_t)ridx * RS), rt1 = 0.0;
    bool prev_ok = false, pend = false;                             // pend: the lower record is in flight (cp.async)
    auto window_load = [&](int r) {                                 // blocking: this lane's piece of record r
        if (haspiece) { T v[V]; ldv(rbase + (int64_t)r * RS + lig * V, v); stv(rcg + (r & 1) * RS + lig * V, v); }
    };
    window_load(ridx);
    if (ridx > 0) { rt1 = rec_get_time(rbase + (int64_t)(ridx - 1) * RS); window_load(ridx - 1); prev_ok = true; }

    // ---- P1: y = sol(ts) and the input features of one stage time, computed by ONE lane of the group ----
    auto prep_block = [&](int blk, double ts) {
        T r[RS];
        double rt;
        if (ts >= rt0 || ridx == 0) { ldv(rcg + (ridx & 1) * RS, r); rt = rt0; }
        else if (prev_ok && (ts >= rt1 || ridx == 1)) { ldv(rcg + ((ridx - 1) & 1) * RS, r); rt = rt1; }
        else {                                                      // outside the window (rare): search the record in global memory
            int k = ridx - 1;
            double rk = rec_get_time(rbase + (int64_t)k * RS);
            while (ts < rk && k > 0) { --k; rk = rec_get_time(rbase + (int64_t)k * RS); }
            ldv(rbase + (int64_t)k * RS, r); rt = rk;
        }
        const T rdt = r[RL::DT];
        T th;
        if constexpr (sizeof(T) == 4) th = kratio((T)(ts - rt), rdt); else th = (T)((ts - rt) / (double)rdt);
        T bw[7]; interp_weights(th, bw);
        T fd[F1S];                                                  // f[NQ1] then df[NQ1], index q*I + i
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
            for (int g = 0; g < G; ++g) { fd[g * I + i] = rb[g]; fd[NQ1 + g * I + i] = rdb[g] * dn; }   // utils.jl:18 * d(arg)/d(xn) * norm'
            swish_both(y, fd[G * I + i], fd[NQ1 + G * I + i]);
        }
        stv(f1g + blk * F1S, fd);
    };

    // ---- A: lambda-independent part of an adjoint RHS evaluation at stage time `blk`:
    //   features c2 of this lane's hidden units -> fac block `blk`;  J2[u][o] = d f_o / d h_j,  dh[u][i] = d h_j / d y_i ----
    auto stage_a = [&](int blk, T (&J2)[UPL][I], T (&dh)[UPL][I]) {
        T fd[F1S];
        ldv(f1g + blk * F1S, fd);
        T c2[SB];
#pragma unroll
        for (int m = SBR; m < SB; ++m) c2[m] = T(0);
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
            if constexpr (I == 2) {                                 // (dh_0, dh_1) in one packed accumulator
                T d0 = T(0), d1 = T(0);
#pragma unroll
                for (int q = 0; q <= G; ++q) kfma2(d0, d1, w1[2 * q], w1[2 * q + 1], fd[NQ1 + 2 * q], fd[NQ1 + 2 * q + 1]);
                dh[u][0] = d0; dh[u][1] = d1;
            } else {
#pragma unroll
                for (int i = 0; i < I; ++i) {
                    T acc = T(0);
#pragma unroll
                    for (int q = 0; q <= G; ++q) acc += w1[q * I + i] * fd[NQ1 + q * I + i];
                    dh[u][i] = acc;
                }
            }
            const T xn = normalize<NORM>(h);
            const T dn = normalize_deriv<NORM>(xn);
            T rb[G], rdb[G];
            rbf_eval<true>(prm, xn, rb, rdb);
            T s, ds; swish_both(h, s, ds);
#pragma unroll
            for (int g = 0; g < G; ++g) c2[u * (G + 1) + g] = rb[g];
            c2[u * (G + 1) + G] = s;
            T w2[NW2];
            ldv(w + NQ1, w2);
            if constexpr (I == 2) {
                T a0 = T(0), a1 = T(0);
#pragma unroll
                for (int g = 0; g < G; ++g) kfma2b(a0, a1, w2[2 * g], w2[2 * g + 1], rdb[g]);
                J2[u][0] = a0 * dn + w2[2 * G] * ds;
                J2[u][1] = a1 * dn + w2[2 * G + 1] * ds;
            } else {
#pragma unroll
                for (int o = 0; o < I; ++o) {
                    T acc = T(0);
#pragma unroll
                    for (int g = 0; g < G; ++g) acc += w2[g * I + o] * rdb[g];
                    J2[u][o] = acc * dn + w2[G * I + o] * ds;
                }
            }
        }
        stv(fac + blk * SB, c2);
    };

    // ---- B: lambda-dependent part with stage adjoint ls: hbar of this lane's units -> fac, dl = -(df/du)^T ls ----
    auto stage_b = [&](int slot, const T (&ls)[I], const T (&J2)[UPL][I], const T (&dh)[UPL][I], T (&dl)[I]) {
        T pu[I];
#pragma unroll
        for (int i = 0; i < I; ++i) pu[i] = T(0);
#pragma unroll
        for (int u = 0; u < UPL; ++u) {
            T hb = T(0);
#pragma unroll
            for (int o = 0; o < I; ++o) hb += ls[o] * J2[u][o];
            fac[NB * SB + slot * UPL + u] = hb;
#pragma unroll
            for (int i = 0; i < I; ++i) pu[i] += hb * dh[u][i];
        }
#pragma unroll
        for (int i = 0; i < I; ++i) {                               // all-gather over the group, summed in lane order
#if KANODE_LG_TREESUM
            if constexpr (sizeof(T) == 4 && LPT == 5) {
                const T a0 = shfl_t(pu[i], gbase), a1 = shfl_t(pu[i], gbase + 1), a2 = shfl_t(pu[i], gbase + 2), a3 = shfl_t(pu[i], gbase + 3), a4 = shfl_t(pu[i], gbase + 4);
                dl[i] = -(((a0 + a1) + (a2 + a3)) + a4);
                continue;
            }
#endif
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
#if KANODE_LG_TREESUM
        if constexpr (sizeof(T) == 4 && LPT == 5) {
            const T a0 = shfl_t(v, gbase), a1 = shfl_t(v, gbase + 1), a2 = shfl_t(v, gbase + 2), a3 = shfl_t(v, gbase + 3), a4 = shfl_t(v, gbase + 4);
            return ((a0 + a1) + (a2 + a3)) + a4;
        }
#endif
        T tot = T(0);
#pragma unroll
        for (int k = 0; k < LPT; ++k) tot += shfl_t(v, gbase + k);
        return tot;
    };

    double t = t1;
    int sp = a.nsave - 1;                                           // next preset (save) time, descending
    // Jumps: save index s travels as [a.saveat[s] | dL/du(t_s)] through mailbox slot s & 1, fetched asynchronously TWO jumps
    // ahead by the first lane of the group (the save time is needed at the very next loop header, the cotangent at the end of the
    // attempt that reaches t_s).  Equal consecutive save times (two jumps at one instant) switch the trajectory to plain loads.
    constexpr int MBS = SMP::MBS, OT = RL::OT;
    auto mb_fetch = [&](int s) {
        if (s >= 0 && lig == 0 && gvalid) {
            T* slot = mbg + (s & 1) * MBS;
            cp_async<8>(slot, a.saveat + s);
#pragma unroll
            for (int i = 0; i < I; ++i) cp_async<(int)sizeof(T)>(slot + OT + i, dgb + s * I + i);
        }
    };
    bool mb_off = false;
    mb_fetch(sp); mb_fetch(sp - 1);
    cp_async_wait_all();
    __syncwarp();
    double nxt_t = sp >= 0 ? a.saveat[sp] : 0.0;
    T jdg[I];                                                       // cotangent of the next jump, read from the mailbox every attempt
#pragma unroll
    for (int i = 0; i < I; ++i) jdg[i] = sp >= 0 ? mbg[(sp & 1) * MBS + OT + i] : T(0);
    auto apply_jumps = [&](double tt) {
        bool mod = false, first = true;
        while (sp >= 0 && nxt_t == tt) {
            if (first && !mb_off) {
#pragma unroll
                for (int i = 0; i < I; ++i) lam[i] += jdg[i];
                mb_fetch(sp - 2);                                   // refills the slot just consumed
                --sp;
                nxt_t = sp >= 0 ? rec_get_time(mbg + (sp & 1) * MBS) : 0.0;     // landed at least one attempt ago
            } else {
                mb_off = true;
#pragma unroll
                for (int i = 0; i < I; ++i) lam[i] += dgb[sp * I + i];
                --sp;
                nxt_t = sp >= 0 ? a.saveat[sp] : 0.0;
            }
            mod = true; first = false;
        }
        return mod;
    };
    apply_jumps(t1);                                                // PresetTimeCallback fires at init when t_end is a save time
#pragma unroll
    for (int i = 0; i < I; ++i) lprev[i] = lam[i];

    T J0[UPL][I], D0[UPL][I];                                       // part A at the current time t (block 0), carried over attempts
    double dt;                                                      // |dt|; integration runs in -t
    {   // ---- FSAL evaluation + ode_determine_initdt on the augmented state (g(T) = 0: its scale is abstol) ----
        __syncwarp();                                               // record window visible to the group
        if (lig == 0 && gvalid) prep_block(0, t1);
        __syncwarp();
        stage_a(0, J0, D0);
        stage_b(0, lam, J0, D0, kl[0]);
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
                for (int m = 0; m < SB; ++m) { const T x = by_abstol(lam[o] * c0[m]); ga += x * x; }
#pragma unroll
            for (int u = 0; u < UPL; ++u) {
                const T hb0 = fac[NB * SB + u];
#pragma unroll
                for (int m = 0; m < NQ1; ++m) { const T x = by_abstol(hb0 * f0[m]); ga += x * x; }
            }
        }
        s1 += group_sum(ga);
        const double d0 = sqrt((double)s0 / NZ), d1 = sqrt((double)s1 / NZ);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
        dt0 = fmin(dt0, dtmax);
        T l1[I], fv[I];
#pragma unroll
        for (int i = 0; i < I; ++i) l1[i] = lam[i] - (T)dt0 * kl[0][i];
        if (lig == 0 && gvalid) prep_block(1, t1 - dt0);
        __syncwarp();
        T J1[UPL][I], D1[UPL][I];
        stage_a(1, J1, D1);
        stage_b(1, l1, J1, D1, fv);
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
                for (int m = 0; m < SB; ++m) { const T x = by_abstol(l1[o] * c1[m] - lam[o] * c0[m]); gb += x * x; }
#pragma unroll
            for (int u = 0; u < UPL; ++u) {
                const T hb0 = fac[NB * SB + u], hb1 = fac[NB * SB + UPL + u];
#pragma unroll
                for (int m = 0; m < NQ1; ++m) { const T x = by_abstol(hb1 * f1[m] - hb0 * f0[m]); gb += x * x; }
            }
        }
        s2 += group_sum(gb);
        const double d2 = sqrt((double)s2 / NZ) / dt0;
        const double mx = fmax(d1, d2);
        const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        dt = fmax(dtmin0, fmin(fmin(100.0 * dt0, dt1), dtmax));
    }
    using LC = LgCtrl<T>;
    typename LC::Q qold = (typename LC::Q)Ctrl::qoldinit, q11 = (typename LC::Q)1;
    double dtpropose = dt;
    bool accept = false, modified = false, clipped = false;
    float margin = 1e30f;
    int iter = 0;
    if (!(t > t0)) done = true;
    const double* rp = a.rp_t ? a.rp_t + bq * (int64_t)a.rp_cap : nullptr;

    while (__any_sync(0xffffffffu, !done)) {
        // ---- loopheader! ----
        if (iter > 0) {
            if (!accept) dt = LC::shrink(dt, q11);
            else dt = dtpropose;
        }
        ++iter;
        const double tstop = (sp >= 0) ? fmax(nxt_t, t0) : t0;
        // max(eps(t), eps(t0), eps(t1)) = dtmin0: t stays inside [t0, t1] and the spacing of doubles grows with |t|, so the
        // per-attempt eps(t) of the package's formula never wins (same value, ~20 double-precision instructions fewer per attempt)
        const double dtmin_t = dtmin0;
        {   // step margin (launch order of the next call; no effect on this solve): counted once a step has been clipped
            const float want = (float)fmin(fabs(dt), dtmax), room = (float)(t - tstop);
            if (!done && clipped) margin = fminf(margin, want * krcp(room));
            clipped |= want >= room;
        }
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
        // ---- P1: input features at the NEW stage times 1..NB-1 (block 0 = time t is carried over), one stage time per lane ----
#pragma unroll
        for (int r0 = 1; r0 < NB; r0 += LPT) {
            const int blk = r0 + lig;
            if (blk < NB && gvalid) prep_block(blk, t - tab_c(blk) * dt);
        }
        __syncwarp();
        // ---- perform_step! on lambda; after a jump the FSAL stage is re-evaluated (stage 0 of the same loop; without a
        // jump the recomputed stage 0 equals the FSAL value, the evaluation count follows the package).
        // Software pipeline: part A of stage s+1 next to the serial chain of stage s. ----
        const T h = (T)(-dt);
        T lnew[I];
#pragma unroll
        for (int i = 0; i < I; ++i) lnew[i] = lprev[i];
        T Jc[UPL][I], Dc[UPL][I];
#pragma unroll
        for (int u = 0; u < UPL; ++u)
#pragma unroll
            for (int i = 0; i < I; ++i) { Jc[u][i] = J0[u][i]; Dc[u][i] = D0[u][i]; }
        KANODE_LG_UNROLL(KANODE_LG_STAGE_UNROLL)
        for (int s = 0; s < NB - 1; ++s) {
            T Jn[UPL][I], Dn[UPL][I];
            stage_a(s + 1, Jn, Dn);
            T ls[I], ks[I];
#pragma unroll
            for (int i = 0; i < I; ++i) {
                T acc = T(0);
#pragma unroll
                for (int j = 0; j < NB - 2; ++j) acc += Tab<T>::a(s, j) * kl[j][i];
                ls[i] = lprev[i] + h * acc;
            }
            stage_b(s, ls, Jc, Dc, ks);
#pragma unroll
            for (int j = 0; j < NB - 1; ++j)
                if (j == s) {
#pragma unroll
                    for (int i = 0; i < I; ++i) kl[j][i] = ks[i];
                }
#pragma unroll
            for (int u = 0; u < UPL; ++u)
#pragma unroll
                for (int i = 0; i < I; ++i) { Jc[u][i] = Jn[u][i]; Dc[u][i] = Dn[u][i]; }
        }
#pragma unroll
        for (int s = NB - 1; s < 7; ++s) {                          // stages 6 and 7 share the stage time t - dt (block NB-1)
            T ls[I];
#pragma unroll
            for (int i = 0; i < I; ++i) {
                T acc = T(0);
#pragma unroll
                for (int j = 0; j < s; ++j) acc += Tab<T>::a(s, j) * kl[j][i];
                ls[i] = lprev[i] + h * acc;
            }
            stage_b(s, ls, Jc, Dc, kl[s]);
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
            const T r = kratio(ut, sc);
            es += r * r;
            bad |= (lnew[i] != lnew[i]);
        }
        // ---- step-end pass over this lane's gradient components: g_new = g + (-h) sum_s b_s kv_s, error term with btilde.
        // kv_s is rank one: (stage adjoint or hidden cotangent of stage s) x (feature of the stage time of s); the features
        // are the ones part A parked in shared memory; stages 6 and 7 share a feature block (their weights are merged).
        // g_new goes to the already consumed factor blocks 1..NB-2. ----
        T esl = T(0);
        const T mh = -h;
        T esl2 = T(0);
        auto fin2 = [&](T go0, T go1, T vb0, T vb1, T vt0, T vt1, T& gn0, T& gn1) {   // two components per packed instruction
            kadd2(gn0, gn1, go0, go1, vb0, vb1);
            T sc0 = abstol, sc1 = abstol;
            kfma2b(sc0, sc1, kmax(kabs(go0), kabs(gn0)), kmax(kabs(go1), kabs(gn1)), reltol);
            T r0, r1;
#if KANODE_LG_RCP1
            if constexpr (sizeof(T) == 4) { const T inv = krcp(sc0 * sc1); kmul2(r0, r1, vt0 * sc1, vt1 * sc0, inv, inv); } else { r0 = vt0 / sc0; r1 = vt1 / sc1; }
#else
            if constexpr (sizeof(T) == 4) kmul2(r0, r1, vt0, vt1, krcp(sc0), krcp(sc1)); else { r0 = vt0 / sc0; r1 = vt1 / sc1; }
#endif
            kfma2(esl, esl2, r0, r1, r0, r1);
        };
        {   // layer 2: component (o, m): kv_s = lambda_s[o] * c2_s[m]
            T lsv[LSS];
            ldv(lsg, lsv);
            T ab[I][NB], at[I][NB];
#pragma unroll
            for (int o = 0; o < I; ++o) lg_merge<T, NB>(lsv, I, o, mh, ab[o], at[o]);
#pragma unroll
            for (int m0 = 0; m0 < SB; m0 += V) {
                T c[NB][V];
#pragma unroll
                for (int k = 0; k < NB; ++k) ldv(fac + k * SB + m0, c[k]);
                T vb[I][V], vt[I][V];
#pragma unroll
                for (int o = 0; o < I; ++o)
#pragma unroll
                    for (int e = 0; e < V; ++e) { vb[o][e] = T(0); vt[o][e] = T(0); }
#pragma unroll
                for (int k = 0; k < NB; ++k)
#pragma unroll
                    for (int o = 0; o < I; ++o)
#if KANODE_LG_PAIRC
#pragma unroll
                        for (int e = 0; e < V; e += 2) {                // pairs of COMPONENTS (the pairing of the finish below): no re-pairing moves
                            kfma2b(vb[o][e], vb[o][e + 1], c[k][e], c[k][e + 1], ab[o][k]);
                            kfma2b(vt[o][e], vt[o][e + 1], c[k][e], c[k][e + 1], at[o][k]);
                        }
#else
#pragma unroll
                        for (int e = 0; e < V; ++e) kfma2b(vb[o][e], vt[o][e], ab[o][k], at[o][k], c[k][e]);
#endif
#pragma unroll
                for (int o = 0; o < I; ++o) {
                    T gn[V];
#pragma unroll
                    for (int e = 0; e < V; e += 2) fin2(g2[o][m0 + e], g2[o][m0 + e + 1], vb[o][e], vb[o][e + 1], vt[o][e], vt[o][e + 1], gn[e], gn[e + 1]);
                    stv(fac + (1 + o) * SB + m0, gn);
                }
            }
        }
        {   // layer 1: component (u, m): kv_s = hbar_s[j0+u] * f_s[m]
            T hbv[HBN];
            ldv(fac + NB * SB, hbv);
            T ab[UPL][NB], at[UPL][NB];
#pragma unroll
            for (int u = 0; u < UPL; ++u) lg_merge<T, NB>(hbv, UPL, u, mh, ab[u], at[u]);
#pragma unroll
            for (int m0 = 0; m0 < NQ1; m0 += V) {
                T c[NB][V];
#pragma unroll
                for (int k = 0; k < NB; ++k) ldv(f1g + k * F1S + m0, c[k]);
                T vb[UPL][V], vt[UPL][V];
#pragma unroll
                for (int u = 0; u < UPL; ++u)
#pragma unroll
                    for (int e = 0; e < V; ++e) { vb[u][e] = T(0); vt[u][e] = T(0); }
#pragma unroll
                for (int k = 0; k < NB; ++k)
#pragma unroll
                    for (int u = 0; u < UPL; ++u)
#if KANODE_LG_PAIRC
#pragma unroll
                        for (int e = 0; e < V; e += 2) {
                            kfma2b(vb[u][e], vb[u][e + 1], c[k][e], c[k][e + 1], ab[u][k]);
                            kfma2b(vt[u][e], vt[u][e + 1], c[k][e], c[k][e + 1], at[u][k]);
                        }
#else
#pragma unroll
                        for (int e = 0; e < V; ++e) kfma2b(vb[u][e], vt[u][e], ab[u][k], at[u][k], c[k][e]);
#endif
#pragma unroll
                for (int u = 0; u < UPL; ++u) {
                    T gn[V];
#pragma unroll
                    for (int e = 0; e < V; e += 2) fin2(g1[u][m0 + e], g1[u][m0 + e + 1], vb[u][e], vb[u][e + 1], vt[u][e], vt[u][e + 1], gn[e], gn[e + 1]);
                    stv(fac + (1 + I) * SB + u * NQ1 + m0, gn);
                }
            }
        }
        es += group_sum(esl + esl2);
        const typename LC::Q EEst = LC::eest(es, NZ);
        if (!done && (EEst != EEst || bad)) { ret = RET_UNSTABLE; done = true; }
        // ---- loopfooter!: PI controller ----
        const typename LC::Q q = LC::q(EEst, qold, q11);
        accept = rp ? true : (EEst <= (typename LC::Q)1);
        cp_async_wait_all();                                        // copies issued at the end of the previous attempt
        __syncwarp();                                               // ... visible to the group; every lane is past its reads of f1g
        if (pend) { rt1 = rec_get_time(rcg + ((ridx - 1) & 1) * RS); prev_ok = true; pend = false; }   // record window: lower record landed
        if (sp >= 0) {
#pragma unroll
            for (int i = 0; i < I; ++i) jdg[i] = mb_off ? dgb[sp * I + i] : mbg[(sp & 1) * MBS + OT + i];
        }
        if (!done) {
            if (accept) {
                ++naccept;
                qold = LC::qold_next(EEst);
                const double dtnew = LC::grow(dt, q);
                double tnew = t - dt;
                if (rp) tnew = rp_next;
                else if (fabs(tnew - tstop) < 100.0 * eps_of(fmax(fabs(t), fabs(tstop)))) tnew = tstop;
                dtpropose = fmax(fmin(dtmax, fabs(dtnew)), dtmin0);                 // max(eps(tnew), dtmin0) = dtmin0, see loopheader
                t = tnew;
#pragma unroll
                for (int o = 0; o < I; ++o) ldv(fac + (1 + o) * SB, g2[o]);      // commit g_new
#pragma unroll
                for (int u = 0; u < UPL; ++u) ldv(fac + (1 + I) * SB + u * NQ1, g1[u]);
#pragma unroll
                for (int i = 0; i < I; ++i) lam[i] = lnew[i];
                // the last stage time of this step is the first of the next one: block NB-1 -> block 0
                {
                    T c5[SB];
                    ldv(fac + (NB - 1) * SB, c5); stv(fac, c5);
                    if (gvalid && lig < NQ1 / V) { T f5[V]; ldv(f1g + (NB - 1) * F1S + lig * V, f5); stv(f1g + lig * V, f5); }
#pragma unroll
                    for (int u = 0; u < UPL; ++u)
#pragma unroll
                        for (int i = 0; i < I; ++i) { J0[u][i] = Jc[u][i]; D0[u][i] = Dc[u][i]; }
                }
                modified = apply_jumps(t);
#pragma unroll
                for (int i = 0; i < I; ++i) lprev[i] = lam[i];
                if (!(t > t0)) done = true;
            } else {
                ++nreject;
            }
        }
        // ---- record window: slide when t left the upper record ----
        if (t < rt0 && ridx > 0 && !done) {
            if (prev_ok && (t >= rt1 || ridx == 1)) { --ridx; rt0 = rt1; }
            else {                                                  // t jumped over a whole record: search, reload (blocking)
                int k = ridx - 1;
                double rk = rec_get_time(rbase + (int64_t)k * RS);
                while (t < rk && k > 0) { --k; rk = rec_get_time(rbase + (int64_t)k * RS); }
                ridx = k; rt0 = rk;
                window_load(ridx);
            }
            prev_ok = false;
            if (ridx > 0) {                                         // the record below: in flight during the next attempt
                if (haspiece) cp_async<16>(rcg + ((ridx - 1) & 1) * RS + lig * V, rbase + (int64_t)(ridx - 1) * RS + lig * V);
                pend = true;
            }
        }
        __syncwarp();                                               // factor / feature / record slices are rewritten by the next attempt
    }

    // ---- results: per-warp gradient sum (fixed order over the trajectories of the warp), du0, statistics ----
    const bool ok = active && !skipped && ret == RET_SUCCESS;
    T* gp = a.gpart + (int64_t)wslot * NP;
#pragma unroll
    for (int o = 0; o < I; ++o)
#pragma unroll
        for (int m = 0; m < SBR; ++m) {
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
            if (lane < LPT) {
                const int q = m / I, i = m % I, j = j0 + u;          // feature index m = q*I + i
                gp[q < G ? P::OC1 + (i * G + q) * P::H + j : P::OW1 + i * P::H + j] = tot;
            }
        }
    if (active && lig == 0) {
        const int64_t b = trajectory();
        if (skipped) { nf = 0; naccept = 0; nreject = 0; }
        if (a.du0)
#pragma unroll
            for (int i = 0; i < I; ++i) a.du0[b * I + i] = skipped ? T(0) : lam[i];
        if (a.stats) a.stats[b] = kanode_stats{naccept, nreject, nf, ret};
        if (a.attempts) a.attempts[b] = naccept + nreject;
        if (a.tmargin) a.tmargin[b] = __float_as_int(margin);
    }
    __syncwarp();                                                   // the next pass rewrites this warp's shared-memory slices
    }
    if (lane == 0) {                                                // last warp out re-arms the ticket for the next launch
        const int total = (int)(gridDim.x * WPB);
        if (atomicAdd(a.ticket + 1, 1) == total - 1) { a.ticket[0] = 0; a.ticket[1] = 0; __threadfence(); }
    }
}

template <class T, class P, int NORM, int UPL, int WPB, int MINB>
__global__ void __launch_bounds__(32 * WPB, MINB) small_backward_lg_kernel(const __grid_constant__ P prm, const LgBwdArgs<T> a) {
    small_backward_lg_body<T, P, NORM, UPL, WPB>(prm, a);
}
// Launch order from the previous call's per-trajectory step margins (LgBwdArgs::order): a STABLE counting sort into LG_NBK
// margin classes (class 0: margin < 1, then 64 classes per octave), smallest margins first, index order inside a class — a
// deterministic function of the margins.  LG_OBLK blocks of 32 warps; every warp owns a contiguous segment, lanes with equal
// class find each other with match.any (one instruction per element whatever LG_NBK is).  Two launches: lg_order_count_kernel
// fills cnt[warp][class], lg_order_scatter_kernel turns it into first positions and writes the permutation.
constexpr int LG_NBK = 64, LG_OBLK = 8;
__device__ __forceinline__ int lg_margin_class(int bits) {
    const float m = __int_as_float(bits);
    if (!(m >= 1.0f)) return 0;
    const int c = 1 + (int)(__log2f(m) * 64.0f);
    return c < LG_NBK - 1 ? c : LG_NBK - 1;
}
__device__ __forceinline__ int lg_order_segment(int n) { return ((n + 32 * LG_OBLK - 1) / (32 * LG_OBLK) + 31) / 32 * 32; }   // whole warp iterations
__global__ void __launch_bounds__(1024) lg_order_count_kernel(const int* __restrict__ key, int n, int* __restrict__ cnt) {
    __shared__ int c_sm[32][LG_NBK];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < 32 * LG_NBK; i += blockDim.x) (&c_sm[0][0])[i] = 0;
    __syncthreads();
    const int seg = lg_order_segment(n);
    const int64_t lo64 = (int64_t)(blockIdx.x * 32 + w) * seg;
    const int lo = (int)(lo64 < n ? lo64 : n), hi = (int)(lo64 + seg < n ? lo64 + seg : n);
    const unsigned below = (1u << lane) - 1u;
    for (int i0 = lo; i0 < hi; i0 += 32) {
        const int i = i0 + lane;
        const int c = i < hi ? lg_margin_class(key[i]) : LG_NBK + lane;      // idle lanes: classes of their own
        const unsigned m = __match_any_sync(0xffffffffu, c);
        if (i < hi && (m & below) == 0) c_sm[w][c] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    for (int i = tid; i < 32 * LG_NBK; i += blockDim.x) cnt[(size_t)blockIdx.x * 32 * LG_NBK + i] = (&c_sm[0][0])[i];
}
__global__ void __launch_bounds__(1024) lg_order_scatter_kernel(const int* __restrict__ key, int n, const int* __restrict__ cnt, int* __restrict__ order) {
    __shared__ int first[32][LG_NBK];                                // first position of (this block's warp, class)
    __shared__ int tot[LG_NBK];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid < LG_NBK) {                                              // members of class `tid`: all warps; the warps before this block
        int all = 0, before = 0;
        for (int k = 0; k < 32 * LG_OBLK; ++k) { const int v = cnt[(size_t)k * LG_NBK + tid]; if (k < (int)blockIdx.x * 32) before += v; all += v; }
        tot[tid] = all;
        int acc = before;
        for (int k = 0; k < 32; ++k) { first[k][tid] = acc; acc += cnt[((size_t)blockIdx.x * 32 + k) * LG_NBK + tid]; }
    }
    __syncthreads();
    if (tid == 0) { int acc = 0; for (int k = 0; k < LG_NBK; ++k) { const int v = tot[k]; tot[k] = acc; acc += v; } }
    __syncthreads();
    const int seg = lg_order_segment(n);
    const int64_t lo64 = (int64_t)(blockIdx.x * 32 + w) * seg;
    const int lo = (int)(lo64 < n ? lo64 : n), hi = (int)(lo64 + seg < n ? lo64 + seg : n);
    const unsigned below = (1u << lane) - 1u;
    for (int i0 = lo; i0 < hi; i0 += 32) {
        const int i = i0 + lane;
        const int c = i < hi ? lg_margin_class(key[i]) : LG_NBK + lane;
        const unsigned m = __match_any_sync(0xffffffffu, c);
        int base = 0;
        if (i < hi) base = tot[c] + first[w][c];
        __syncwarp();
        if (i < hi) {
            order[base + __popc(m & below)] = i;
            if ((m & below) == 0) first[w][c] += __popc(m);
        }
        __syncwarp();
    }
}

// out[c] = scale * sum_r part[r][c]: rows are per-warp gradient sums.  Two deterministic stages: slab sums in fp64
// (grid.x slabs of consecutive rows — two per SM so the 10 MB of partials stream at HBM speed —, thread per column,
// coalesced rows), then the sum over slabs: a warp per column, lanes stride over the slabs, fixed shuffle tree.
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
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= cols) return;                                           // whole warps leave together
    double acc = 0.0;
    for (int s = lane; s < nslab; s += 32) acc += slab[(int64_t)s * cols + c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[c] = (OutT)(acc * scale);
}

}  // namespace kanode
