// kanode_generic.cuh — (stub, filled in next) block-per-trajectory kernels for arbitrary KDense chains.
#pragma once
#include "kanode_host.h"
#include "kanode_math.cuh"
namespace kanode {
inline int generic_init(kanode_handle*) { return 0; }
inline int generic_upload_params(kanode_handle*) { return 0; }
template <class T> int generic_rhs(kanode_handle* h, const T*, T*, int64_t) { return fail(h, KANODE_ERR_UNSUPPORTED, "generic path not built"); }
template <class T> int generic_vjp(kanode_handle* h, const T*, const T*, T*, T*, int64_t) { return fail(h, KANODE_ERR_UNSUPPORTED, "generic path not built"); }
template <class T> int generic_solve(kanode_handle* h, const T*, int64_t, double, double, const double*, int, double, double, T*, kanode_stats*) { return fail(h, KANODE_ERR_UNSUPPORTED, "generic path not built"); }
template <class T> int generic_loss_grad(kanode_handle* h, const T*, int64_t, double, double, const double*, int, const T*, double, double, double*, T*, T*, kanode_stats*, kanode_stats*, T*) { return fail(h, KANODE_ERR_UNSUPPORTED, "generic path not built"); }
}
