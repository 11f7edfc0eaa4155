// kanode_wide.cu — instantiations of the wide lockstep engine (its own translation unit: compiles in parallel with kanode_api.cu)
#include "kanode_wide.cuh"
#include "kanode_wsrc.cuh"

namespace kanode {

#define KANODE_WIDE_DEF(T)                                                                                                       \
    int wsrc_solve(kanode_handle* h, int G, const T* p, const T* d_u0, int64_t B, double t0, double t1, const double* d_saveat,  \
                   int nsave, double abstol, double reltol, T* d_out, kanode_stats* d_stats) {                                   \
        if (G == 5) return wsrc_solve_t<T, 5>(h, p, d_u0, B, t0, t1, d_saveat, nsave, abstol, reltol, d_out, d_stats);           \
        return wsrc_solve_t<T, 10>(h, p, d_u0, B, t0, t1, d_saveat, nsave, abstol, reltol, d_out, d_stats);                      \
    }                                                                                                                            \
    int wsrc_loss_grad(kanode_handle* h, int G, const T* p, const T* d_u0, int64_t B, double t0, double t1,                      \
                       const double* d_saveat, int nsave, const T* d_target, double abstol, double reltol, double* d_loss_sum,   \
                       T* d_grad_sum, T* d_du0, kanode_stats* d_fst, kanode_stats* d_bst, T* d_out_opt) {                        \
        if (G == 5)                                                                                                              \
            return wsrc_loss_grad_t<T, 5>(h, p, d_u0, B, t0, t1, d_saveat, nsave, d_target, abstol, reltol, d_loss_sum,          \
                                          d_grad_sum, d_du0, d_fst, d_bst, d_out_opt);                                           \
        return wsrc_loss_grad_t<T, 10>(h, p, d_u0, B, t0, t1, d_saveat, nsave, d_target, abstol, reltol, d_loss_sum,             \
                                       d_grad_sum, d_du0, d_fst, d_bst, d_out_opt);                                              \
    }                                                                                                                            \
    int wide_rhs(kanode_handle* h, WideKey k, const T* p, const T* d_u, T* d_du, int64_t B) {                                   \
        if (k.G == 5) return wide_rhs_t<T, 10, 5>(h, p, d_u, d_du, B);                                                           \
        return wide_rhs_t<T, 10, 10>(h, p, d_u, d_du, B);                                                                        \
    }                                                                                                                            \
    int wide_solve(kanode_handle* h, WideKey k, const T* p, const T* d_u0, int64_t B, double t0, double t1, const double* d_saveat, \
                   int nsave, double abstol, double reltol, T* d_out, kanode_stats* d_stats) {                                   \
        if (k.G == 5) return wide_solve_t<T, 10, 5>(h, p, d_u0, B, t0, t1, d_saveat, nsave, abstol, reltol, d_out, d_stats);     \
        return wide_solve_t<T, 10, 10>(h, p, d_u0, B, t0, t1, d_saveat, nsave, abstol, reltol, d_out, d_stats);                  \
    }                                                                                                                            \
    int wide_loss_grad(kanode_handle* h, WideKey k, const T* p, const T* d_u0, int64_t B, double t0, double t1,                  \
                       const double* d_saveat, int nsave, const T* d_target, double abstol, double reltol, double* d_loss_sum,   \
                       T* d_grad_sum, T* d_du0, kanode_stats* d_fst, kanode_stats* d_bst, T* d_out_opt) {                        \
        if (k.G == 5)                                                                                                            \
            return wide_loss_grad_t<T, 10, 5>(h, p, d_u0, B, t0, t1, d_saveat, nsave, d_target, abstol, reltol, d_loss_sum,      \
                                              d_grad_sum, d_du0, d_fst, d_bst, d_out_opt);                                       \
        return wide_loss_grad_t<T, 10, 10>(h, p, d_u0, B, t0, t1, d_saveat, nsave, d_target, abstol, reltol, d_loss_sum,         \
                                           d_grad_sum, d_du0, d_fst, d_bst, d_out_opt);                                          \
    }
KANODE_WIDE_DEF(float)
KANODE_WIDE_DEF(double)
#undef KANODE_WIDE_DEF

}  // namespace kanode
