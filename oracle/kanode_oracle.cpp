// kanode_oracle.cpp — CPU restatement of the KAN-ODE hot path.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or call this file.  The
// product (kan_odes_b200/, libkanode_b200.so) never links or calls it.
//
// PARITY UNPINNED: the reference (maharshi-coding/KAN-ODEs) ships no tests, no
// golden vectors and no checkpoint for this path, and Julia is
// not available in the build container, so this oracle cannot be checked against
// outputs of the reference itself.  Its trust comes from independent checks in
// tests/test_oracle_*.py (tableau order conditions, VJP vs torch autograd, solve vs
// scipy, whole-solve gradient vs finite differences).
//
// What it restates (file:line into /root/reference; [EXT pkg ver] = un-vendored
// Julia dependency pinned in Lotka-Volterra/Manifest.toml, restated from its
// published algorithm):
//   KDense forward            Lotka-Volterra/src/kdense.jl:109-130
//   rbf/_rbf, rswaf, iqf      Lotka-Volterra/src/utils.jl:8-13, 27-34, 49-54
//   their reverse rules       Lotka-Volterra/src/utils.jl:15-21, 36-43, 56-62
//   grid / denominator        Lotka-Volterra/src/kdense.jl:26-27, 88-92
//   parameter layout          Lotka-Volterra/src/kdense.jl:70-86, LV_driver_KANODE.jl:173-175
//   activations               [EXT NNlib 0.9.24] tanh_fast, softsign, sigmoid_fast, swish
//   NeuralODE rhs             [EXT DiffEqFlux 4.0.0]  dudt(u,p,t) = model(u,p)
//   source-term rhs           "PDE examples/Allen-Cahn_Source.jl":50-54,90-93,
//                             "PDE examples/Fisher-KPP_Source.jl":55-59,95-98
//   Tsit5 step / interpolant  [EXT OrdinaryDiffEqTsit5 1.1.0]
//   loop, PI controller, initdt, tstops  [EXT OrdinaryDiffEqCore 1.9.0, DiffEqBase 6.158.3]
//   fastpower                 [EXT FastPower 1.1.0]
//   InterpolatingAdjoint      [EXT SciMLSensitivity 7.69.0]
//   loss                      Lotka-Volterra/LV_driver_KANODE.jl:197-203
//
// Two arithmetic instantiations: T=double (what the reference drivers effectively
// run) and T=float (mirrors the device path's arithmetic type).
// Time is always double.

#ifdef _OPENMP
#include <omp.h>
#endif
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/kanode.h"

namespace {

// ---------------------------------------------------------------------------------
// scalar functions
// ---------------------------------------------------------------------------------

// [EXT NNlib 0.9.24] tanh_fast(::Float32): rational approximation, |rel err| <= ~5 eps.
inline float tanh_fast_f32(float x) {
    const float x2 = x * x;
    const float n = 1.0f + x2 * (0.1346604f + x2 * (0.0035974074f + x2 * (2.2332108e-5f + x2 * 1.587199e-8f)));
    const float d = 1.0f + x2 * (0.4679937f + x2 * (0.026262015f + x2 * (0.0003453992f + x2 * 8.7767893e-7f)));
    if (x2 < 66.0f) return x * (n / d);
    return x > 0 ? 1.0f : (x < 0 ? -1.0f : x);
}
// tanh_fast(::Float64) is exp-based and agrees with tanh to ~2 eps(Float64); std::tanh stands in.
inline double act_tanh(double x) { return std::tanh(x); }
inline float act_tanh(float x) { return tanh_fast_f32(x); }

template <class T> inline T sigmoid(T x) {
    // [EXT NNlib] sigmoid: t = exp(-|x|); x >= 0 ? 1/(1+t) : t/(1+t)
    const T t = std::exp(-std::fabs(x));
    return x >= 0 ? T(1) / (T(1) + t) : t / (T(1) + t);
}

template <class T> inline T normalize(int kind, T x) {
    switch (kind) {
        case KANODE_NORM_TANH: return act_tanh(x);
        case KANODE_NORM_SOFTSIGN: return x / (T(1) + std::fabs(x));
        default: return sigmoid(x);
    }
}
// derivative expressed through the primal output (NNlib / ChainRules scalar rules)
template <class T> inline T normalize_deriv(int kind, T omega) {
    switch (kind) {
        case KANODE_NORM_TANH: return T(1) - omega * omega;
        case KANODE_NORM_SOFTSIGN: { const T a = T(1) - std::fabs(omega); return a * a; }
        default: return omega * (T(1) - omega);
    }
}

// basis value and d(basis)/d(arg) as the reference's reverse rules compute it
template <class T> inline void basis_eval(int kind, T a, T& y, T& dy) {
    switch (kind) {
        case KANODE_BASIS_RBF: y = std::exp(-a * a); dy = T(-2) * a * y; break;              // utils.jl:13,18
        case KANODE_BASIS_RSWAF: { const T tx = std::tanh(a); y = T(1) - tx * tx; dy = T(-2) * tx * y; } break; // :33,40
        default: y = T(1) / (T(1) + a * a); dy = T(-2) * a * y; break;                       // utils.jl:54,59 (verbatim)
    }
}

// [EXT FastPower 1.1.0] fastlog2 / fastpower, always Float32 inside
inline float fastlog2(float x) {
    uint32_t bits; std::memcpy(&bits, &x, 4);
    const float e = (float)((bits & 0x7F800000u) >> 23);
    float s, fe;
    if (bits & 0x00400000u) {
        const uint32_t b2 = (bits & 0x007FFFFFu) | 0x3f000000u;
        std::memcpy(&s, &b2, 4); fe = e - 126.0f; s = s - 1.0f;
    } else {
        const uint32_t b2 = (bits & 0x007FFFFFu) | 0x3f800000u;
        std::memcpy(&s, &b2, 4); fe = e - 127.0f; s = s - 1.0f;
    }
    return fe + s * (0.338953f * s + 2.198599f) / (s + 1.523692f);
}
inline float fastpower(double x, double y) {
    if (x == 0) return 0.0f;
    return std::exp2((float)y * fastlog2((float)x));
}

// ---------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------
template <class T> struct Layer {
    int I, O, G, norm, basis, use_base;
    int kind = 0, dense_act = 0;   // kind 1: Lux.Dense, y = act(W x + b) (LV_driver_MLP.jl:61); params [vec(W[O,I]); b[O]] at offW, offB
    size_t offB = 0;
    std::vector<T> grid;  // Float32 LinRange values promoted to T (kdense.jl:90)
    T inv_h;              // Float32 (1/h) promoted to T (utils.jl:9)
    size_t offC, offW;
};

template <class T> struct Model {
    std::vector<Layer<T>> L;
    int rhs_kind = 0, n = 0;
    size_t np = 0;
    T lap_scale = 0;  // lap_coef / dx^2
    int maxw = 0;     // widest layer interface
};

size_t param_count(const kanode_desc* d) {
    if (!d || d->n_layers < 1 || d->n_layers > KANODE_MAX_LAYERS) return 0;
    size_t np = 0;
    for (int l = 0; l < d->n_layers; ++l) {
        const auto& s = d->layers[l];
        if (s.kind == KANODE_LAYER_DENSE) {
            if (s.in_dims < 1 || s.out_dims < 1) return 0;
            np += (size_t)(s.in_dims + 1) * s.out_dims;               // weight[out, in] + bias[out]
            continue;
        }
        if (s.in_dims < 1 || s.out_dims < 1 || s.grid_len < 2) return 0;
        np += (size_t)s.in_dims * s.grid_len * s.out_dims;            // kdense.jl:101
        if (s.use_base_act) np += (size_t)s.in_dims * s.out_dims;     // kdense.jl:103
    }
    return np;
}

template <class T> bool build_model(const kanode_desc* d, Model<T>& m) {
    if (param_count(d) == 0) return false;
    m.rhs_kind = d->rhs_kind; m.n = d->n_state;
    size_t off = 0;
    for (int l = 0; l < d->n_layers; ++l) {
        const auto& s = d->layers[l];
        Layer<T> L;
        L.I = s.in_dims; L.O = s.out_dims; L.G = s.grid_len;
        L.norm = s.normalizer; L.basis = s.basis; L.use_base = s.use_base_act;
        if (s.kind == KANODE_LAYER_DENSE) {
            L.kind = 1; L.dense_act = s.dense_act; L.G = 0; L.use_base = 0; L.inv_h = T(1);
            L.offC = off; L.offW = off; off += (size_t)L.O * L.I; L.offB = off; off += (size_t)L.O;
            m.maxw = std::max(m.maxw, std::max(L.I, L.O));
            if (l > 0 && d->layers[l - 1].out_dims != s.in_dims) return false;
            m.L.push_back(std::move(L));
            continue;
        }
        L.grid.resize(L.G);
        for (int g = 0; g < L.G; ++g) {
            // Julia LinRange{Float32}: lerpi(j,d,a,b) = T((1-t)*a + t*b), t = j/d in Float64
            const double t = (double)g / (double)(L.G - 1);
            L.grid[g] = (T)(float)((1.0 - t) * (double)s.grid_lo + t * (double)s.grid_hi);
        }
        L.inv_h = (T)(1.0f / s.denominator);
        L.offC = off; off += (size_t)L.O * L.G * L.I;
        L.offW = off; if (L.use_base) off += (size_t)L.O * L.I;
        m.maxw = std::max(m.maxw, std::max(L.I, L.O));
        if (l > 0 && d->layers[l - 1].out_dims != s.in_dims) return false;
        m.L.push_back(std::move(L));
    }
    m.np = off;
    if (m.rhs_kind == KANODE_RHS_CHAIN) {
        if (m.L.front().I != m.n || m.L.back().O != m.n) return false;
    } else if (m.rhs_kind == KANODE_RHS_SOURCE_LAPLACIAN) {
        if (m.L.front().I != 1 || m.L.back().O != 1 || m.n < 3) return false;
        m.lap_scale = (T)(d->lap_coef / (d->dx * d->dx));
    } else if (m.rhs_kind == KANODE_RHS_MAP) {                      // a chain used as a map: only chain_forward / activations apply
        if (m.L.front().I != m.n) return false;
    } else return false;
    return true;
}

// per-layer saved intermediates for the reverse pass
template <class T> struct Tape {
    std::vector<std::vector<T>> x, xn, b, db, sw, dsw, z;  // per layer (z: pre-activations of Dense layers)
    void init(const Model<T>& m) {
        const size_t nl = m.L.size();
        x.resize(nl); xn.resize(nl); b.resize(nl); db.resize(nl); sw.resize(nl); dsw.resize(nl); z.resize(nl);
        for (size_t l = 0; l < nl; ++l) {
            const auto& L = m.L[l];
            x[l].resize(L.I); xn[l].resize(L.I); sw[l].resize(L.I); dsw[l].resize(L.I); z[l].resize(L.O);
            b[l].resize((size_t)L.I * L.G); db[l].resize((size_t)L.I * L.G);
        }
    }
};

// y = chain(x): kdense.jl:109-130 applied layer by layer (Lux.Chain)
template <class T>
void chain_forward(const Model<T>& m, const T* p, const T* xin, T* yout, Tape<T>& tp) {
    std::vector<T> cur(xin, xin + m.L.front().I), nxt;
    for (size_t l = 0; l < m.L.size(); ++l) {
        const auto& L = m.L[l];
        nxt.assign(L.O, T(0));
        const T* C = p + L.offC; const T* W = p + L.offW;
        if (L.kind == 1) {                                             // Lux.Dense: y = act(W x + b)  (LV_driver_MLP.jl:61)
            for (int i = 0; i < L.I; ++i) tp.x[l][i] = cur[i];
            for (int o = 0; o < L.O; ++o) {
                T zz = p[L.offB + o];
                for (int i = 0; i < L.I; ++i) zz += W[(size_t)i * L.O + o] * cur[i];
                tp.z[l][o] = zz;
                nxt[o] = L.dense_act == KANODE_ACT_TANH ? std::tanh(zz) : zz;
            }
            cur.swap(nxt);
            continue;
        }
        for (int i = 0; i < L.I; ++i) {
            const T xi = cur[i];
            tp.x[l][i] = xi;
            const T xn = normalize(L.norm, xi);                        // kdense.jl:116
            tp.xn[l][i] = xn;
            for (int g = 0; g < L.G; ++g) {
                const T a = (xn - L.grid[g]) * L.inv_h;                // utils.jl:9
                T y, dy; basis_eval(L.basis, a, y, dy);                // utils.jl:13
                tp.b[l][(size_t)i * L.G + g] = y; tp.db[l][(size_t)i * L.G + g] = dy;
                const T* col = C + ((size_t)i * L.G + g) * L.O;        // column (i,g) of C[O, G*I]
                for (int o = 0; o < L.O; ++o) nxt[o] += col[o] * y;    // kdense.jl:120
            }
            if (L.use_base) {
                const T sg = sigmoid(xi);
                const T s = xi * sg;                                   // swish on RAW x, kdense.jl:123
                tp.sw[l][i] = s; tp.dsw[l][i] = s + sg * (T(1) - s);
                const T* col = W + (size_t)i * L.O;
                for (int o = 0; o < L.O; ++o) nxt[o] += col[o] * s;
            }
        }
        cur.swap(nxt);
    }
    std::copy(cur.begin(), cur.end(), yout);
}

// reverse of chain_forward using the saved tape: xbar = J_x^T ybar ; pbar += J_p^T ybar
template <class T>
void chain_reverse(const Model<T>& m, const T* p, const T* ybar, T* xbar, T* pbar, const Tape<T>& tp) {
    std::vector<T> cur(ybar, ybar + m.L.back().O), nxt;
    for (int l = (int)m.L.size() - 1; l >= 0; --l) {
        const auto& L = m.L[l];
        nxt.assign(L.I, T(0));
        const T* C = p + L.offC; const T* W = p + L.offW;
        T* Cb = pbar ? pbar + L.offC : nullptr; T* Wb = pbar ? pbar + L.offW : nullptr;
        if (L.kind == 1) {                                             // Dense: zbar = ybar .* act'(z); Wbar += zbar x'; bbar += zbar; xbar = W' zbar
            for (int o = 0; o < L.O; ++o) {
                const T th = L.dense_act == KANODE_ACT_TANH ? std::tanh(tp.z[l][o]) : T(0);
                const T zb = cur[o] * (L.dense_act == KANODE_ACT_TANH ? (T(1) - th * th) : T(1));
                if (pbar) { pbar[L.offB + o] += zb; for (int i = 0; i < L.I; ++i) Wb[(size_t)i * L.O + o] += zb * tp.x[l][i]; }
                for (int i = 0; i < L.I; ++i) nxt[i] += W[(size_t)i * L.O + o] * zb;
            }
            cur.swap(nxt);
            continue;
        }
        for (int i = 0; i < L.I; ++i) {
            T xnbar = 0;
            for (int g = 0; g < L.G; ++g) {
                const size_t ig = (size_t)i * L.G + g;
                const T* col = C + ig * L.O;
                T bbar = 0;
                for (int o = 0; o < L.O; ++o) bbar += col[o] * cur[o];
                if (Cb) { T* cb = Cb + ig * L.O; const T y = tp.b[l][ig]; for (int o = 0; o < L.O; ++o) cb[o] += cur[o] * y; }
                xnbar += tp.db[l][ig] * bbar * L.inv_h;                // utils.jl:18 then d(arg)/d(xn) = 1/h
            }
            T xb = xnbar * normalize_deriv(L.norm, tp.xn[l][i]);
            if (L.use_base) {
                const T* col = W + (size_t)i * L.O;
                T sbar = 0;
                for (int o = 0; o < L.O; ++o) sbar += col[o] * cur[o];
                if (Wb) { T* wb = Wb + (size_t)i * L.O; const T s = tp.sw[l][i]; for (int o = 0; o < L.O; ++o) wb[o] += cur[o] * s; }
                xb += sbar * tp.dsw[l][i];
            }
            nxt[i] = xb;
        }
        cur.swap(nxt);
    }
    std::copy(cur.begin(), cur.end(), xbar);
}

// du = f(u)  (autonomous).  Chain: du = chain(u).  Source: du = s*D*lap*u + chain.(u)
template <class T> struct RhsWork { Tape<T> tp; std::vector<Tape<T>> node_tp; };

template <class T>
void rhs_eval(const Model<T>& m, const T* p, const T* u, T* du, Tape<T>& tp) {
    if (m.rhs_kind == KANODE_RHS_CHAIN) { chain_forward(m, p, u, du, tp); return; }
    const int n = m.n;
    for (int j = 0; j < n; ++j) {
        const T um = u[(j + n - 1) % n], up = u[(j + 1) % n];          // periodic corners AC_Source:53-54
        T k; chain_forward(m, p, u + j, &k, tp);
        du[j] = m.lap_scale * (um - T(2) * u[j] + up) + k;             // AC_Source:92
    }
}

// ubar = (df/du)^T lam ; pbar += (df/dp)^T lam   (recomputes the forward, like Zygote.pullback)
template <class T>
void rhs_vjp(const Model<T>& m, const T* p, const T* u, const T* lam, T* ubar, T* pbar, Tape<T>& tp) {
    if (m.rhs_kind == KANODE_RHS_CHAIN) {
        std::vector<T> y(m.n);
        chain_forward(m, p, u, y.data(), tp);
        chain_reverse(m, p, lam, ubar, pbar, tp);
        return;
    }
    const int n = m.n;
    for (int j = 0; j < n; ++j) {
        T k, xb; chain_forward(m, p, u + j, &k, tp);
        chain_reverse(m, p, lam + j, &xb, pbar, tp);
        const T lm = lam[(j + n - 1) % n], lp = lam[(j + 1) % n];      // lap is symmetric
        ubar[j] = m.lap_scale * (lm - T(2) * lam[j] + lp) + xb;
    }
}

// ---------------------------------------------------------------------------------
// Tsit5  [EXT OrdinaryDiffEqTsit5 1.1.0]
// ---------------------------------------------------------------------------------
namespace tab {
constexpr double c1 = 0.161, c2 = 0.327, c3 = 0.9, c4 = 0.9800255409045097;
constexpr double a21 = 0.161;
constexpr double a31 = -0.008480655492356989, a32 = 0.335480655492357;
constexpr double a41 = 2.8971530571054935, a42 = -6.359448489975075, a43 = 4.3622954328695815;
constexpr double a51 = 5.325864828439257, a52 = -11.748883564062828, a53 = 7.4955393428898365, a54 = -0.09249506636175525;
constexpr double a61 = 5.86145544294642, a62 = -12.92096931784711, a63 = 8.159367898576159, a64 = -0.071584973281401, a65 = -0.028269050394068383;
constexpr double a71 = 0.09646076681806523, a72 = 0.01, a73 = 0.4798896504144996, a74 = 1.379008574103742, a75 = -3.290069515436081, a76 = 2.324710524099774;
constexpr double bt1 = -0.00178001105222577714, bt2 = -0.0008164344596567469, bt3 = 0.007880878010261995,
                 bt4 = -0.1447110071732629, bt5 = 0.5823571654525552, bt6 = -0.45808210592918697, bt7 = 0.015151515151515152;
// dense output b_i(theta)
constexpr double r011 = 1.0, r012 = -2.763706197274826, r013 = 2.9132554618219126, r014 = -1.0530884977290216;
constexpr double r22 = 0.13169999999999998, r23 = -0.2234, r24 = 0.1017;
constexpr double r32 = 3.9302962368947516, r33 = -5.941033872131505, r34 = 2.490627285651253;
constexpr double r42 = -12.411077166933676, r43 = 30.33818863028232, r44 = -16.548102889244902;
constexpr double r52 = 37.50931341651104, r53 = -88.1789048947664, r54 = 47.37952196281928;
constexpr double r62 = -27.896526289197286, r63 = 65.09189467479366, r64 = -34.87065786149661;
constexpr double r72 = 1.5, r73 = -4.0, r74 = 2.5;
}  // namespace tab

template <class T> inline void interp_weights(T th, T b[7]) {
    using namespace tab;
    const T th2 = th * th;
    b[0] = th * (T(r011) + th * (T(r012) + th * (T(r013) + th * T(r014))));
    b[1] = th2 * (T(r22) + th * (T(r23) + th * T(r24)));
    b[2] = th2 * (T(r32) + th * (T(r33) + th * T(r34)));
    b[3] = th2 * (T(r42) + th * (T(r43) + th * T(r44)));
    b[4] = th2 * (T(r52) + th * (T(r53) + th * T(r54)));
    b[5] = th2 * (T(r62) + th * (T(r63) + th * T(r64)));
    b[6] = th2 * (T(r72) + th * (T(r73) + th * T(r74)));
}

// one accepted forward step: everything the interpolant needs
template <class T> struct StepRec { double t; T dt; std::vector<T> u, k; /* k: 7*N */ };

template <class T> struct Dense {
    int N = 0;
    std::vector<StepRec<T>> steps;
    double t_end = 0;
    // u(t): left-continuous choice when `left`, right-continuous otherwise
    void eval(double t, T* out, bool left) const {
        // steps are increasing in time (forward solve)
        size_t lo = 0, hi = steps.size();
        while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (steps[mid].t <= t) lo = mid; else hi = mid; }
        size_t k = lo;
        if (left && k > 0 && steps[k].t == t) --k;
        const auto& s = steps[k];
        const T th = (T)((t - s.t) / (double)s.dt);
        T b[7]; interp_weights(th, b);
        for (int i = 0; i < N; ++i) {
            T acc = 0;
            for (int j = 0; j < 7; ++j) acc += b[j] * s.k[(size_t)j * N + i];
            out[i] = s.u[i] + s.dt * acc;
        }
    }
};

template <class T> inline T rms_scaled(const T* num, const T* a, const T* b, T abstol, T reltol, int N) {
    // DiffEqBase calculate_residuals + ODE_DEFAULT_NORM
    T s = 0;
    for (int i = 0; i < N; ++i) {
        const T sc = abstol + std::max(std::fabs(a[i]), std::fabs(b[i])) * reltol;
        const T r = num[i] / sc; s += r * r;
    }
    return std::sqrt(s / (T)N);
}

struct SolveOpts {
    double abstol = 1e-6, reltol = 1e-3;
    int maxiters = 100000;
    // controller [EXT OrdinaryDiffEqCore 1.9.0 defaults for Tsit5]
    double beta1 = 7.0 / 50.0, beta2 = 2.0 / 25.0, gamma = 0.9, qmin = 0.2, qmax = 10.0, qoldinit = 1e-4;
};

// Generic adaptive Tsit5 over z in R^N.
//   f(t, z, dz)                     autonomous here, t passed for completeness
//   preset: times (in integration order) at which `affect(idx, z)` fires after the step that lands on
//           them (or at init when equal to t0); they are tstops.  affect returns true if z was modified.
//   on_accept(rec-like)             called with (tprev, dt, uprev, k[7]) for every accepted step
template <class T, class F, class Affect, class OnAccept>
int tsit5_solve(int N, T* u, double t0, double t1, const SolveOpts& o, F&& f,
                const std::vector<double>& preset, Affect&& affect, OnAccept&& on_accept, kanode_stats& st) {
    using namespace tab;
    const double tdir = t1 >= t0 ? 1.0 : -1.0;
    const T abstol = (T)o.abstol, reltol = (T)o.reltol;
    st.naccept = st.nreject = st.nf = 0; st.retcode = KANODE_RET_SUCCESS;
    if (t0 == t1) return 0;

    // tstops: preset times strictly after t0 (in direction), plus the end time
    std::vector<double> tstops;
    size_t next_preset = 0;
    for (double tp : preset) if (tdir * tp > tdir * t0 && tdir * tp < tdir * t1) tstops.push_back(tp);
    tstops.push_back(t1);
    size_t ts_i = 0;

    double t = t0;
    // callbacks initialise first (PresetTimeCallback fires at t0 if t0 is a preset time)
    while (next_preset < preset.size() && preset[next_preset] == t0) { affect(next_preset, u); ++next_preset; }

    std::vector<T> k(7 * (size_t)N), uprev(u, u + N), tmp(N), utilde(N), unew(N), sk(N), f1(N);
    T* k1 = k.data(); T* k2 = k1 + N; T* k3 = k2 + N; T* k4 = k3 + N; T* k5 = k4 + N; T* k6 = k5 + N; T* k7 = k6 + N;
    f(t, u, k1); st.nf += 1;  // fsalfirst

    const double dtmax = std::fabs(t1 - t0);
    auto eps_of = [](double x) { return std::nextafter(std::fabs(x), INFINITY) - std::fabs(x); };
    const double dtmin0 = std::max(eps_of(t0), eps_of(t1));

    // ---- initial dt: Hairer heuristic, ode_determine_initdt (out-of-place) ----
    double dt;
    {
        const double smalldt = 1e-6;
        for (int i = 0; i < N; ++i) sk[i] = abstol + std::fabs(u[i]) * reltol;
        T s0 = 0, s1 = 0;
        for (int i = 0; i < N; ++i) { const T a = u[i] / sk[i]; s0 += a * a; const T b = k1[i] / sk[i]; s1 += b * b; }
        const double d0 = std::sqrt((double)s0 / N), d1 = std::sqrt((double)s1 / N);
        st.nf += 1;  // f0 (the package evaluates it again; counted like the package does)
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? smalldt : (d0 / d1) / 100.0;
        dt0 = std::min(dt0, dtmax);
        const double dt0s = tdir * dt0;
        for (int i = 0; i < N; ++i) tmp[i] = u[i] + (T)dt0s * k1[i];
        f(t + dt0s, tmp.data(), f1.data()); st.nf += 1;
        T s2 = 0;
        for (int i = 0; i < N; ++i) { const T a = (f1[i] - k1[i]) / sk[i]; s2 += a * a; }
        const double d2 = std::sqrt((double)s2 / N) / dt0;
        const double mx = std::max(d1, d2);
        double dt1v = (mx <= 1e-15) ? std::max(smalldt, dt0 * 1e-3) : std::pow(10.0, -(2.0 + std::log10(mx)) / 5.0);
        dt = tdir * std::max(dtmin0, std::min(std::min(100.0 * dt0, dt1v), dtmax));
    }

    double qold = o.qoldinit, q11 = 1.0, dtpropose = dt;
    bool accept = false, u_modified = false;
    int iter = 0;
    T EEst = 1;

    while (ts_i < tstops.size()) {
        while (tdir * t < tdir * tstops[ts_i]) {
            // ---- loopheader! ----
            if (iter > 0) {
                if (!accept) {
                    dt = dt / std::min(1.0 / o.qmin, q11 / o.gamma);            // step_reject_controller!
                } else {
                    std::copy(u, u + N, uprev.begin());                         // apply_step!
                    dt = dtpropose;
                    if (u_modified) { f(t, u, k1); st.nf += 1; u_modified = false; }  // reeval FSAL
                    else std::copy(k7, k7 + N, k1);
                }
            }
            ++iter;
            const double dtmin_t = std::max(eps_of(t), dtmin0);
            {   // fix_dt_at_bounds! + modify_dt_for_tstops!
                double a = std::min(std::fabs(dt), dtmax); a = std::max(a, dtmin_t);
                a = std::min(a, std::fabs(tstops[ts_i] - t));
                dt = tdir * a;
            }
            // ---- check_error! ----
            if (iter > o.maxiters) { st.retcode = KANODE_RET_MAXITERS; return 0; }
            if (!(std::fabs(dt) > dtmin_t) && (tdir * (t + dt) < tdir * tstops[ts_i] || !accept) && iter > 1) {
                st.retcode = KANODE_RET_DT_LESS_THAN_MIN; return 0;
            }
            if (std::isnan(dt)) { st.retcode = KANODE_RET_UNSTABLE; return 0; }
            for (int i = 0; i < N; ++i) if (std::isnan(u[i])) { st.retcode = KANODE_RET_UNSTABLE; return 0; }

            // ---- perform_step! (Tsit5ConstantCache) ----
            const T h = (T)dt;
            for (int i = 0; i < N; ++i) tmp[i] = uprev[i] + h * (T(a21) * k1[i]);
            f(t + c1 * dt, tmp.data(), k2);
            for (int i = 0; i < N; ++i) tmp[i] = uprev[i] + h * (T(a31) * k1[i] + T(a32) * k2[i]);
            f(t + c2 * dt, tmp.data(), k3);
            for (int i = 0; i < N; ++i) tmp[i] = uprev[i] + h * (T(a41) * k1[i] + T(a42) * k2[i] + T(a43) * k3[i]);
            f(t + c3 * dt, tmp.data(), k4);
            for (int i = 0; i < N; ++i) tmp[i] = uprev[i] + h * (T(a51) * k1[i] + T(a52) * k2[i] + T(a53) * k3[i] + T(a54) * k4[i]);
            f(t + c4 * dt, tmp.data(), k5);
            for (int i = 0; i < N; ++i) tmp[i] = uprev[i] + h * (T(a61) * k1[i] + T(a62) * k2[i] + T(a63) * k3[i] + T(a64) * k4[i] + T(a65) * k5[i]);
            f(t + dt, tmp.data(), k6);
            for (int i = 0; i < N; ++i) unew[i] = uprev[i] + h * (T(a71) * k1[i] + T(a72) * k2[i] + T(a73) * k3[i] + T(a74) * k4[i] + T(a75) * k5[i] + T(a76) * k6[i]);
            f(t + dt, unew.data(), k7);
            st.nf += 6;
            for (int i = 0; i < N; ++i)
                utilde[i] = h * (T(bt1) * k1[i] + T(bt2) * k2[i] + T(bt3) * k3[i] + T(bt4) * k4[i] + T(bt5) * k5[i] + T(bt6) * k6[i] + T(bt7) * k7[i]);
            EEst = rms_scaled(utilde.data(), uprev.data(), unew.data(), abstol, reltol, N);

            // ---- loopfooter!: PI controller ----
            double q;
            if (EEst == 0) q = 1.0 / o.qmax;
            else {
                q11 = fastpower((double)EEst, o.beta1);
                q = q11 / fastpower(qold, o.beta2);
                q = std::max(1.0 / o.qmax, std::min(1.0 / o.qmin, q / o.gamma));
            }
            accept = (EEst <= 1);   // NaN -> reject
            if (std::isnan((double)EEst)) { st.retcode = KANODE_RET_UNSTABLE; return 0; }
            if (accept) {
                st.naccept += 1;
                // step_accept_controller! (qsteady_min = qsteady_max = 1)
                if (q == 1.0) q = 1.0;
                qold = std::max((double)EEst, o.qoldinit);
                const double dtnew = dt / q;
                const double tprev = t;
                double ttmp = t + dt;
                {   // fixed_t_for_floatingpoint_error!
                    const double tstop = tstops[ts_i];
                    const double mx = std::max(std::fabs(t), std::fabs(tstop));
                    if (std::fabs(ttmp - tstop) < 100.0 * eps_of(mx)) ttmp = tstop;
                }
                t = ttmp;
                {   // calc_dt_propose!
                    double a = std::min(dtmax, std::fabs(dtnew)); a = std::max(a, std::max(eps_of(t), dtmin0));
                    dtpropose = tdir * a;
                }
                on_accept(tprev, h, uprev.data(), k.data());
                std::copy(unew.begin(), unew.end(), u);
                // handle_callbacks!: preset-time affects that coincide with the new t
                while (next_preset < preset.size() && preset[next_preset] == t) {
                    if (affect(next_preset, u)) u_modified = true;
                    ++next_preset;
                }
            } else {
                st.nreject += 1;
            }
        }
        // handle_tstop!
        while (ts_i < tstops.size() && tstops[ts_i] == t) ++ts_i;
        if (ts_i < tstops.size() && tdir * tstops[ts_i] < tdir * t) ++ts_i;  // defensive
    }
    return 0;
}

// ---------------------------------------------------------------------------------
// forward solve with saveat  (NeuralODE call, LV_driver_KANODE.jl:180-184)
// ---------------------------------------------------------------------------------
template <class T>
void forward_dense(const Model<T>& m, const T* p, const T* u0, double t0, double t1, const SolveOpts& o,
                   Dense<T>& dense, T* u_end, kanode_stats& st) {
    const int N = m.n;
    Tape<T> tp; tp.init(m);
    std::vector<T> u(u0, u0 + N);
    dense.N = N; dense.steps.clear(); dense.t_end = t1;
    auto f = [&](double, const T* z, T* dz) { rhs_eval(m, p, z, dz, tp); };
    auto affect = [](size_t, T*) { return false; };
    auto on_accept = [&](double tprev, T dt, const T* uprev, const T* k) {
        StepRec<T> r; r.t = tprev; r.dt = dt; r.u.assign(uprev, uprev + N); r.k.assign(k, k + 7 * (size_t)N);
        dense.steps.push_back(std::move(r));
    };
    tsit5_solve<T>(N, u.data(), t0, t1, o, f, {}, affect, on_accept, st);
    if (u_end) std::copy(u.begin(), u.end(), u_end);
    if (dense.steps.empty()) {  // t0 == t1 or immediate failure: constant record so eval() works
        StepRec<T> r; r.t = t0; r.dt = T(1); r.u.assign(u0, u0 + N); r.k.assign(7 * (size_t)N, T(0));
        dense.steps.push_back(std::move(r));
    }
}

template <class T>
int solve_one(const Model<T>& m, const T* p, const T* u0, double t0, double t1, const double* saveat, int nsave,
              const SolveOpts& o, T* out /* [nsave][n] */, kanode_stats& st, Dense<T>* keep = nullptr) {
    Dense<T> local; Dense<T>& d = keep ? *keep : local;
    forward_dense(m, p, u0, t0, t1, o, d, (T*)nullptr, st);
    for (int s = 0; s < nsave; ++s) d.eval(saveat[s], out + (size_t)s * m.n, /*left=*/true);
    return 0;
}

// ---------------------------------------------------------------------------------
// loss + interpolating adjoint  [EXT SciMLSensitivity 7.69.0]
// ---------------------------------------------------------------------------------
template <class T>
int loss_grad_one(const Model<T>& m, const T* p, const T* u0, double t0, double t1, const double* saveat, int nsave,
                  const T* target, const SolveOpts& o, double& loss_sum, T* grad /* [np], overwritten */,
                  T* du0 /* [n] or null */, kanode_stats& fst, kanode_stats& bst, T* out_opt,
                  double* fwd_t = nullptr, double* bwd_t = nullptr, int step_cap = 0,
                  const T* cot = nullptr /* caller-supplied dL/dpred [nsave][n]: replaces the MSE cotangent (target unused) */) {
    const int n = m.n; const size_t np = m.np; const int N = n + (int)np;
    Dense<T> dense;
    std::vector<T> out((size_t)nsave * n);
    solve_one(m, p, u0, t0, t1, saveat, nsave, o, out.data(), fst, &dense);
    if (out_opt) std::copy(out.begin(), out.end(), out_opt);
    if (fwd_t) for (int s = 0; s < step_cap; ++s)       // end time of each accepted forward step (dt-replay tooling)
        fwd_t[s] = s < (int)dense.steps.size() ? (s + 1 < (int)dense.steps.size() ? dense.steps[s + 1].t : dense.t_end) : NAN;
    if (bwd_t) std::fill(bwd_t, bwd_t + step_cap, (double)NAN);
    // loss = mean(abs2, X - pred)  => dL/dpred = 2 (pred - X) / (n*nsave)
    std::vector<T> dg((size_t)nsave * n);
    double ls = 0;
    const T scale = T(2) / (T)((double)n * nsave);
    if (cot) std::copy(cot, cot + out.size(), dg.begin());      // any loss(pred): Zygote hands its cotangent to the adjoint
    else for (size_t i = 0; i < out.size(); ++i) { const T e = out[i] - target[i]; ls += (double)e * (double)e; dg[i] = scale * e; }
    loss_sum = ls;
    std::fill(grad, grad + np, T(0));
    if (fst.retcode != KANODE_RET_SUCCESS) { bst = kanode_stats{0, 0, 0, fst.retcode}; if (du0) std::fill(du0, du0 + n, T(0)); return 0; }

    // backward problem on z = [lambda; g], z(T) = 0, integrated T -> t0 with tstops at the save times
    std::vector<int> order(nsave);
    for (int s = 0; s < nsave; ++s) order[s] = s;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return saveat[a] > saveat[b]; });
    std::vector<double> preset; for (int s : order) preset.push_back(saveat[s]);

    Tape<T> tp; tp.init(m);
    std::vector<T> z(N, T(0)), y(n), ub(n);
    auto f = [&](double t, const T* zz, T* dz) {
        dense.eval(t, y.data(), /*left=*/false);                 // y = sol(t)
        std::fill(dz + n, dz + N, T(0));
        rhs_vjp(m, p, y.data(), zz, ub.data(), dz + n, tp);      // Zygote.pullback(...)(lambda)
        for (int i = 0; i < n; ++i) dz[i] = -ub[i];              // dlambda = -(df/du)^T lambda
        for (int i = n; i < N; ++i) dz[i] = -dz[i];              // dg      = -(df/dp)^T lambda
    };
    auto affect = [&](size_t idx, T* zz) {
        const int s = order[idx];
        for (int i = 0; i < n; ++i) zz[i] += dg[(size_t)s * n + i];   // lambda += dL/du(t_s)
        return true;
    };
    // accepted adjoint steps: step k starts at tprev_k, so its end time is the start of step k+1 (t0 for the last one)
    std::vector<double> bstart;
    auto on_accept = [&](double tprev, T, const T*, const T*) { bstart.push_back(tprev); };
    tsit5_solve<T>(N, z.data(), t1, t0, o, f, preset, affect, on_accept, bst);
    if (bwd_t) for (int s = 0; s < step_cap && s < (int)bstart.size(); ++s) bwd_t[s] = s + 1 < (int)bstart.size() ? bstart[s + 1] : t0;
    if (du0) std::copy(z.begin(), z.begin() + n, du0);
    std::copy(z.begin() + n, z.end(), grad);
    return 0;
}

template <class T> struct Api {
    static int rhs(const kanode_desc* d, const T* p, const T* u, T* du, int64_t batch) {
        Model<T> m; if (!build_model(d, m)) return KANODE_ERR_INVALID;
#pragma omp parallel
        {
            Tape<T> tp; tp.init(m);
#pragma omp for schedule(static)
            for (int64_t b = 0; b < batch; ++b) rhs_eval(m, p, u + b * m.n, du + b * m.n, tp);
        }
        return 0;
    }
    static int vjp(const kanode_desc* d, const T* p, const T* u, const T* lam, T* ubar, T* pbar, int64_t batch) {
        Model<T> m; if (!build_model(d, m)) return KANODE_ERR_INVALID;
        Tape<T> tp; tp.init(m);
        std::fill(pbar, pbar + m.np, T(0));
        for (int64_t b = 0; b < batch; ++b) rhs_vjp(m, p, u + b * m.n, lam + b * m.n, ubar + b * m.n, pbar, tp);
        return 0;
    }
    static int solve(const kanode_desc* d, const T* p, const T* u0, int64_t batch, double t0, double t1,
                     const double* saveat, int nsave, double abstol, double reltol, T* out, kanode_stats* stats,
                     double* step_t /* [batch][cap] or null */, int cap) {
        Model<T> m; if (!build_model(d, m)) return KANODE_ERR_INVALID;
        SolveOpts o; o.abstol = abstol; o.reltol = reltol;
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t b = 0; b < batch; ++b) {
            kanode_stats st; Dense<T> dense;
            solve_one(m, p, u0 + b * m.n, t0, t1, saveat, nsave, o, out + (size_t)b * nsave * m.n, st, &dense);
            if (stats) stats[b] = st;
            if (step_t) for (int s = 0; s < cap; ++s)
                step_t[(size_t)b * cap + s] = s < (int)dense.steps.size() ? dense.steps[s].t + (double)dense.steps[s].dt : NAN;
        }
        return 0;
    }
    static int loss_grad(const kanode_desc* d, const T* p, const T* u0, int64_t batch, double t0, double t1,
                         const double* saveat, int nsave, const T* target, double abstol, double reltol,
                         double* loss, T* grad, T* du0, kanode_stats* fst, kanode_stats* bst, T* out_opt,
                         double* fwd_t = nullptr, double* bwd_t = nullptr, int step_cap = 0) {
        Model<T> m; if (!build_model(d, m)) return KANODE_ERR_INVALID;
        SolveOpts o; o.abstol = abstol; o.reltol = reltol;
        std::vector<double> gsum(m.np, 0.0); double lsum = 0;
#pragma omp parallel
        {
            std::vector<T> g(m.np); std::vector<double> gl(m.np, 0.0); double ll = 0;
#pragma omp for schedule(dynamic, 16)
            for (int64_t b = 0; b < batch; ++b) {
                kanode_stats f, bk; double ls = 0;
                loss_grad_one(m, p, u0 + b * m.n, t0, t1, saveat, nsave, target + (size_t)b * nsave * m.n, o, ls,
                              g.data(), du0 ? du0 + b * m.n : nullptr, f, bk,
                              out_opt ? out_opt + (size_t)b * nsave * m.n : nullptr,
                              fwd_t ? fwd_t + (size_t)b * step_cap : nullptr, bwd_t ? bwd_t + (size_t)b * step_cap : nullptr, step_cap);
                ll += ls; for (size_t i = 0; i < m.np; ++i) gl[i] += (double)g[i];
                if (fst) fst[b] = f;
                if (bst) bst[b] = bk;
            }
#pragma omp critical
            { lsum += ll; for (size_t i = 0; i < m.np; ++i) gsum[i] += gl[i]; }
        }
        *loss = lsum / ((double)batch * nsave * m.n);
        for (size_t i = 0; i < m.np; ++i) grad[i] = (T)(gsum[i] / (double)batch);
        return 0;
    }
    // pullback of the solve: grad = sum_b (d pred_b / d p)^T cot_b, du0[b] = (d pred_b / d u0_b)^T cot_b  (no 1/batch: the
    // cotangent carries every scale), what Zygote.gradient(loss, p) does for an arbitrary loss(pred)
    // (LV_driver_KANODE.jl:197-203 with the optional reg term, Burgers_Surrogate.jl:105-107 with a transposed target)
    static int adjoint(const kanode_desc* d, const T* p, const T* u0, int64_t batch, double t0, double t1,
                       const double* saveat, int nsave, const T* cot, double abstol, double reltol,
                       T* grad, T* du0, kanode_stats* fst, kanode_stats* bst, T* out_opt) {
        Model<T> m; if (!build_model(d, m)) return KANODE_ERR_INVALID;
        SolveOpts o; o.abstol = abstol; o.reltol = reltol;
        std::vector<double> gsum(m.np, 0.0);
#pragma omp parallel
        {
            std::vector<T> g(m.np); std::vector<double> gl(m.np, 0.0);
#pragma omp for schedule(dynamic, 16)
            for (int64_t b = 0; b < batch; ++b) {
                kanode_stats f, bk; double ls = 0;
                loss_grad_one(m, p, u0 + b * m.n, t0, t1, saveat, nsave, (const T*)nullptr, o, ls, g.data(),
                              du0 ? du0 + b * m.n : nullptr, f, bk, out_opt ? out_opt + (size_t)b * nsave * m.n : nullptr,
                              nullptr, nullptr, 0, cot + (size_t)b * nsave * m.n);
                for (size_t i = 0; i < m.np; ++i) gl[i] += (double)g[i];
                if (fst) fst[b] = f;
                if (bst) bst[b] = bk;
            }
#pragma omp critical
            { for (size_t i = 0; i < m.np; ++i) gsum[i] += gl[i]; }
        }
        for (size_t i = 0; i < m.np; ++i) grad[i] = (T)gsum[i];
        return 0;
    }
    // a chain evaluated as a map x[I_0] -> y[O_last] (direct layer call, kdense.jl:109-130)
    static int map(const kanode_desc* d, const T* p, const T* x, T* y, int64_t batch) {
        Model<T> m; if (!build_model(d, m)) return KANODE_ERR_INVALID;
        Tape<T> tp; tp.init(m);
        const int I = m.L.front().I, O = m.L.back().O;
        for (int64_t b = 0; b < batch; ++b) chain_forward(m, p, x + b * I, y + b * O, tp);
        return 0;
    }
    // per-edge activations of layer l at its inputs x[K][I]: act[k][i][o] = sum_g C[o,(i,g)] basis_g(norm(x_i)) + W[o,i] swish(x_i)
    // (Activation_getter.jl:22-31,44-54; their sum over i is the layer output, :33-36)
    static int edge_activations(const kanode_desc* d, int l, const T* p, const T* x, T* act, int64_t K) {
        Model<T> m; if (!build_model(d, m)) return KANODE_ERR_INVALID;
        if (l < 0 || l >= (int)m.L.size()) return KANODE_ERR_INVALID;
        const auto& L = m.L[l];
        const T* C = p + L.offC; const T* W = p + L.offW;
        for (int64_t k = 0; k < K; ++k)
            for (int i = 0; i < L.I; ++i) {
                const T xi = x[k * L.I + i];
                const T xn = normalize(L.norm, xi);
                T* a = act + ((size_t)k * L.I + i) * L.O;
                std::fill(a, a + L.O, T(0));
                for (int g = 0; g < L.G; ++g) {
                    T y, dy; basis_eval(L.basis, (xn - L.grid[g]) * L.inv_h, y, dy);
                    const T* col = C + ((size_t)i * L.G + g) * L.O;
                    for (int o = 0; o < L.O; ++o) a[o] += col[o] * y;
                }
                if (L.use_base) { const T s = xi * sigmoid(xi); const T* col = W + (size_t)i * L.O; for (int o = 0; o < L.O; ++o) a[o] += col[o] * s; }
            }
        return 0;
    }
    // reg_loss(p, act_reg, entropy_reg) (LV_driver_KANODE.jl:187-194) and its gradient (what Zygote differentiates at :199-201)
    static int reg_loss(const T* p, size_t np, double act_reg, double entropy_reg, double* loss, T* grad) {
        double S = 0; for (size_t i = 0; i < np; ++i) S += std::fabs((double)p[i]);
        double E = 0;
        for (size_t i = 0; i < np; ++i) { const double e = std::fabs((double)p[i]) / S; if (e > 0) E -= e * std::log(e); }
        *loss = S * act_reg + E * entropy_reg;
        if (grad) for (size_t i = 0; i < np; ++i) {
            const double a = std::fabs((double)p[i]), sg = p[i] > 0 ? 1.0 : (p[i] < 0 ? -1.0 : 0.0);
            const double dE = a > 0 ? -(std::log(a / S) + E) / S : 0.0;          // d(-sum e log e)/d|p_i|
            grad[i] = (T)(sg * (act_reg + entropy_reg * dE));
        }
        return 0;
    }
};

}  // namespace

extern "C" {

size_t kanode_oracle_param_count(const kanode_desc* d) { return param_count(d); }

// OpenMP thread count of the batched entry points (a launcher such as torchrun exports OMP_NUM_THREADS=1; the bench sets
// the count it reports explicitly).  n <= 0 leaves it unchanged.  Returns the count in effect.
int kanode_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n; return 1;
#endif
}

#define ORACLE_DEFINE(SUF, T)                                                                                        \
    int kanode_oracle_rhs_##SUF(const kanode_desc* d, const T* p, const T* u, T* du, int64_t batch) {                \
        return Api<T>::rhs(d, p, u, du, batch);                                                                      \
    }                                                                                                                \
    int kanode_oracle_vjp_##SUF(const kanode_desc* d, const T* p, const T* u, const T* lam, T* ubar, T* pbar,        \
                                int64_t batch) {                                                                     \
        return Api<T>::vjp(d, p, u, lam, ubar, pbar, batch);                                                         \
    }                                                                                                                \
    int kanode_oracle_solve_##SUF(const kanode_desc* d, const T* p, const T* u0, int64_t batch, double t0, double t1, \
                                  const double* saveat, int nsave, double abstol, double reltol, T* out,             \
                                  kanode_stats* stats, double* step_t, int cap) {                                    \
        return Api<T>::solve(d, p, u0, batch, t0, t1, saveat, nsave, abstol, reltol, out, stats, step_t, cap);       \
    }                                                                                                                \
    int kanode_oracle_loss_grad_##SUF(const kanode_desc* d, const T* p, const T* u0, int64_t batch, double t0,       \
                                      double t1, const double* saveat, int nsave, const T* target, double abstol,    \
                                      double reltol, double* loss, T* grad, T* du0, kanode_stats* fst,               \
                                      kanode_stats* bst, T* out_opt) {                                               \
        return Api<T>::loss_grad(d, p, u0, batch, t0, t1, saveat, nsave, target, abstol, reltol, loss, grad, du0,    \
                                 fst, bst, out_opt);                                                                 \
    }                                                                                                                \
    /* same, also returning the END times of the accepted forward / adjoint steps ([batch][step_cap], NaN padded) */  \
    int kanode_oracle_loss_grad_steps_##SUF(const kanode_desc* d, const T* p, const T* u0, int64_t batch, double t0, \
                                            double t1, const double* saveat, int nsave, const T* target,             \
                                            double abstol, double reltol, double* loss, T* grad, T* du0,             \
                                            kanode_stats* fst, kanode_stats* bst, T* out_opt, double* fwd_t,         \
                                            double* bwd_t, int step_cap) {                                           \
        return Api<T>::loss_grad(d, p, u0, batch, t0, t1, saveat, nsave, target, abstol, reltol, loss, grad, du0,    \
                                 fst, bst, out_opt, fwd_t, bwd_t, step_cap);                                         \
    }

ORACLE_DEFINE(f64, double)
ORACLE_DEFINE(f32, float)

#define ORACLE_DEFINE2(SUF, T)                                                                                       \
    int kanode_oracle_adjoint_##SUF(const kanode_desc* d, const T* p, const T* u0, int64_t batch, double t0, double t1, \
                                    const double* saveat, int nsave, const T* cot, double abstol, double reltol,     \
                                    T* grad, T* du0, kanode_stats* fst, kanode_stats* bst, T* out_opt) {            \
        return Api<T>::adjoint(d, p, u0, batch, t0, t1, saveat, nsave, cot, abstol, reltol, grad, du0, fst, bst, out_opt); \
    }                                                                                                                \
    int kanode_oracle_map_##SUF(const kanode_desc* d, const T* p, const T* x, T* y, int64_t batch) {                 \
        return Api<T>::map(d, p, x, y, batch);                                                                       \
    }                                                                                                                \
    int kanode_oracle_edge_activations_##SUF(const kanode_desc* d, int l, const T* p, const T* x, T* act, int64_t K) { \
        return Api<T>::edge_activations(d, l, p, x, act, K);                                                         \
    }                                                                                                                \
    int kanode_oracle_reg_loss_##SUF(const T* p, size_t np, double act_reg, double entropy_reg, double* loss, T* grad) { \
        return Api<T>::reg_loss(p, np, act_reg, entropy_reg, loss, grad);                                            \
    }
ORACLE_DEFINE2(f64, double)
ORACLE_DEFINE2(f32, float)

// exposed for the controller / tableau tests
float kanode_oracle_fastpower(double x, double y) { return fastpower(x, y); }
float kanode_oracle_tanh_fast_f32(float x) { return tanh_fast_f32(x); }
void kanode_oracle_interp_weights(double th, double* b7) { interp_weights<double>(th, b7); }
void kanode_oracle_tableau(double* c /*6*/, double* a /*7x7 row-major, row s = stage s+1*/, double* bt /*7*/) {
    using namespace tab;
    const double cc[6] = {c1, c2, c3, c4, 1.0, 1.0};
    std::copy(cc, cc + 6, c);
    std::fill(a, a + 49, 0.0);
    const double rows[6][6] = {{a21, 0, 0, 0, 0, 0}, {a31, a32, 0, 0, 0, 0}, {a41, a42, a43, 0, 0, 0},
                               {a51, a52, a53, a54, 0, 0}, {a61, a62, a63, a64, a65, 0}, {a71, a72, a73, a74, a75, a76}};
    for (int s = 0; s < 6; ++s) for (int j = 0; j < 6; ++j) a[(s + 1) * 7 + j] = rows[s][j];
    const double b[7] = {bt1, bt2, bt3, bt4, bt5, bt6, bt7};
    std::copy(b, b + 7, bt);
}

}  // extern "C"
