// kanode_small_host.h — host-side helpers of the small-model ensemble kernels shared by kanode_api.cu and kanode_lg.cu:
// the compile-time registry of [I,H,I] chains, the __grid_constant__ parameter block and the packed weight images.
#pragma once
#include <vector>

#include "kanode_host.h"
#include "kanode_small.cuh"

namespace kanode {

inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

// ---------------------------------------------------------------------------------------------------------
// small-model registry: [I,H,I] chains with compile-time shapes (thread-per-trajectory kernels)
// ---------------------------------------------------------------------------------------------------------
struct SmallKey { int I, H, G, norm; };
inline bool small_match(const kanode_desc& d, SmallKey& k) {
    if (d.rhs_kind != KANODE_RHS_CHAIN || d.n_layers != 2) return false;
    const kanode_layer_desc &a = d.layers[0], &b = d.layers[1];
    if (a.kind != KANODE_LAYER_KDENSE || b.kind != KANODE_LAYER_KDENSE) return false;
    if (a.basis != KANODE_BASIS_RBF || b.basis != KANODE_BASIS_RBF || !a.use_base_act || !b.use_base_act) return false;
    if (a.grid_len != b.grid_len || a.normalizer != b.normalizer || a.grid_lo != b.grid_lo || a.grid_hi != b.grid_hi ||
        a.denominator != b.denominator) return false;
    k = SmallKey{a.in_dims, a.out_dims, a.grid_len, a.normalizer};
    return true;
}

template <class T, class P> void fill_small(const kanode_handle* h, P& p) {
    for (int i = 0; i < P::NP; ++i) p.w[i] = (T)h->params[i];
    const kanode_layer_desc& s = h->desc.layers[0];
    const double inv_h = (double)(1.0f / s.denominator);               // Float32 1/h (utils.jl:9)
    const double sc = KRbfScale<T>::value;
    p.hs = (T)(inv_h * sc);
    for (int g = 0; g < P::G; ++g) p.gs[g] = (T)((double)grid_point(s, g) * inv_h * sc);
    p.dk = (T)(-2.0 * inv_h / sc);
}

// Device-resident training leaves the host copy of the parameters behind (kanode_train_*): bring it back before anything
// that reads h->params (blocks on the stream).
inline int host_params_refresh(kanode_handle* h) {
    if (!h->params_host_stale) return 0;
    std::vector<float> tmp(h->np);
    CK(h, cudaMemcpyAsync(tmp.data(), h->ws[kanode_handle::W_TR_P].p, sizeof(float) * h->np, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < h->np; ++i) h->params[i] = (double)tmp[i];
    h->params_host_stale = false;
    return 0;
}

// packed per-hidden-unit weights for the shared-memory kernels (layout: SmallParams::UW), uploaded when the parameters changed
template <class T, class P> int upload_packed(kanode_handle* h, const T** out) {
    T* d = nullptr;
    const int slot = sizeof(T) == 4 ? 0 : 1;
    if (slot == 0) ENSURE(h, W_WPK32, sizeof(T) * P::WPK, d); else ENSURE(h, W_WPK64, sizeof(T) * P::WPK, d);
    if (h->wpk_version[slot] != h->params_version) {
        if (int rc = host_params_refresh(h)) return rc;
        std::vector<T> pk((size_t)P::WPK, T(0));
        constexpr int I = P::I, H = P::H, G = P::G, NQ = P::NQ;
        for (int j = 0; j < H; ++j) {
            T* w = pk.data() + (size_t)j * P::UW;
            for (int i = 0; i < I; ++i) {
                for (int g = 0; g < G; ++g) w[i * G + g] = (T)h->params[P::OC1 + (i * G + g) * H + j];
                w[I * G + i] = (T)h->params[P::OW1 + i * H + j];
            }
            for (int g = 0; g < G; ++g)
                for (int o = 0; o < I; ++o) w[NQ + g * I + o] = (T)h->params[P::OC2 + (j * G + g) * I + o];
            for (int o = 0; o < I; ++o) w[NQ + G * I + o] = (T)h->params[P::OW2 + j * I + o];
        }
        CK(h, cudaMemcpyAsync(d, pk.data(), sizeof(T) * pk.size(), cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));                       // pk is a stack-lifetime staging buffer
        h->wpk_version[slot] = h->params_version;
    }
    *out = d;
    return 0;
}

// Visitor: calls fn.template operator()<P, NORM>() for the instantiation matching the descriptor.
#define KANODE_SMALL_CASES(X) X(2, 10, 5, NORM_TANH)

template <class T, class Fn> bool small_dispatch(const kanode_handle* h, Fn&& fn, int& rc) {
    SmallKey k;
    if (!small_match(h->desc, k)) return false;
#define X(I_, H_, G_, N_)                                                         \
    if (k.I == I_ && k.H == H_ && k.G == G_ && k.norm == N_) {                     \
        rc = fn.template operator()<SmallParams<T, I_, H_, G_>, N_>();             \
        return true;                                                               \
    }
    KANODE_SMALL_CASES(X)
#undef X
    return false;
}


// Lane-group adjoint engine (kanode_lg.cu): dense forward solve + interpolating-adjoint backward solve + gradient sum for
// the chains of the small registry.  *handled = false when the descriptor is not in the registry.
template <class T>
int small_lg_loss_grad(kanode_handle* h, const T* d_u0, int64_t B, double t0, double t1, const double* d_saveat, int nsave,
                       const T* d_target, double abstol, double reltol, double* d_loss_sum, T* d_grad_sum, T* d_du0,
                       kanode_stats* d_fst, kanode_stats* d_bst, T* d_out_opt, const double* d_rp_fwd, const double* d_rp_bwd,
                       int rp_cap, bool* handled);

// Device-side refresh of the small-model weight images (forward kernels' packed per-unit weights, lane-block image of the
// adjoint kernel) from fp32 device parameters: no host copy, no synchronisation.  *handled = false outside the registry.
int small_pack_dev(kanode_handle* h, const float* d_p, bool* handled);

}  // namespace kanode
