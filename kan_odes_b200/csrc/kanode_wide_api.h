// kanode_wide_api.h — entry points of the wide lockstep engine (kanode_wide.cuh, compiled in kanode_wide.cu)
#pragma once
#include "kanode_host.h"

namespace kanode {

struct WideKey { int H, G; };
inline bool wide_match(const kanode_desc& d, WideKey& k) {
    if (d.rhs_kind != KANODE_RHS_CHAIN || d.n_layers != 2) return false;
    const kanode_layer_desc &a = d.layers[0], &b = d.layers[1];
    if (a.kind != KANODE_LAYER_KDENSE || b.kind != KANODE_LAYER_KDENSE) return false;
    if (a.basis != KANODE_BASIS_RBF || b.basis != KANODE_BASIS_RBF || !a.use_base_act || !b.use_base_act) return false;
    if (a.grid_len != b.grid_len || a.grid_len > 16) return false;
    if (a.in_dims != b.out_dims || a.in_dims < 16) return false;     // narrow states belong to the thread-per-trajectory kernels
    k = WideKey{a.out_dims, a.grid_len};
    return (k.H == 10) && (k.G == 5 || k.G == 10);
}


// hidden-source model (periodic Laplacian + pointwise 1 -> 1 KDense) handled by the lockstep engine of kanode_wsrc.cuh
inline bool wsrc_match(const kanode_desc& d, int& G) {
    if (d.rhs_kind != KANODE_RHS_SOURCE_LAPLACIAN || d.n_layers != 1) return false;
    const kanode_layer_desc& a = d.layers[0];
    if (a.kind != KANODE_LAYER_KDENSE || a.in_dims != 1 || a.out_dims != 1 || a.basis != KANODE_BASIS_RBF || !a.use_base_act) return false;
    G = a.grid_len;
    return G == 5 || G == 10;
}

#define KANODE_WIDE_DECL(T)                                                                                                      \
    int wide_rhs(kanode_handle* h, WideKey k, const T* p, const T* d_u, T* d_du, int64_t B);                                    \
    int wsrc_solve(kanode_handle* h, int G, const T* p, const T* d_u0, int64_t B, double t0, double t1, const double* d_saveat,  \
                   int nsave, double abstol, double reltol, T* d_out, kanode_stats* d_stats);                                    \
    int wsrc_loss_grad(kanode_handle* h, int G, const T* p, const T* d_u0, int64_t B, double t0, double t1,                      \
                       const double* d_saveat, int nsave, const T* d_target, double abstol, double reltol, double* d_loss_sum,   \
                       T* d_grad_sum, T* d_du0, kanode_stats* d_fst, kanode_stats* d_bst, T* d_out_opt);                         \
    int wide_solve(kanode_handle* h, WideKey k, const T* p, const T* d_u0, int64_t B, double t0, double t1, const double* d_saveat, \
                   int nsave, double abstol, double reltol, T* d_out, kanode_stats* d_stats);                                    \
    int wide_loss_grad(kanode_handle* h, WideKey k, const T* p, const T* d_u0, int64_t B, double t0, double t1,                  \
                       const double* d_saveat, int nsave, const T* d_target, double abstol, double reltol, double* d_loss_sum,   \
                       T* d_grad_sum, T* d_du0, kanode_stats* d_fst, kanode_stats* d_bst, T* d_out_opt);
KANODE_WIDE_DECL(float)
KANODE_WIDE_DECL(double)
#undef KANODE_WIDE_DECL

}  // namespace kanode
