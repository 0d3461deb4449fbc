// rbk_sum.cu -- af_sum_by_key / af_product_by_key (+ _nan) instantiations, promoted like af_sum
// (/root/reference/src/api/c/reduce.cpp:230-304).
#include "reduce_dispatch.cuh"
#include "rbk_kernels.cuh"
namespace afb {
bool rbk_family_sum(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const void *keys, dim_t n, const View &vals, int dim,
                    bool cn, double nv, const unsigned *tile_base, void *keys_out, void *vals_out, dim_t n_reduced) {
#define X(ci, Ti, co, To)                                                                                          \
    if (ti == ci && to == co) {                                                                                    \
        if (op == OP_ADD) run_rbk<OP_ADD, Ti, To>(c, keys, n, vals, dim, cn, nv, tile_base, keys_out, vals_out, n_reduced); \
        else run_rbk<OP_MUL, Ti, To>(c, keys, n, vals, dim, cn, nv, tile_base, keys_out, vals_out, n_reduced);     \
        return true;                                                                                               \
    }
    AFB_SUM_PAIRS(X)
#undef X
    return false;
}
}  // namespace afb
