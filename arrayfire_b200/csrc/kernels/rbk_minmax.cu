// rbk_minmax.cu -- af_min_by_key / af_max_by_key instantiations: T -> T for all 14 dtypes
// (/root/reference/src/api/c/reduce.cpp:156-228).  Elements of a run are folded in position order, so the complex
// magnitude rule (ties keep the earlier element) needs no index tracking here.
#include "reduce_dispatch.cuh"
#include "rbk_kernels.cuh"
namespace afb {
bool rbk_family_minmax(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const void *keys, dim_t n, const View &vals, int dim,
                       bool cn, double nv, const unsigned *tile_base, void *keys_out, void *vals_out, dim_t n_reduced) {
    if (ti != to) return false;
#define X(code, T)                                                                                               \
    if (ti == code) {                                                                                            \
        if (op == OP_MIN) run_rbk<OP_MIN, T, T>(c, keys, n, vals, dim, cn, nv, tile_base, keys_out, vals_out, n_reduced); \
        else run_rbk<OP_MAX, T, T>(c, keys, n, vals, dim, cn, nv, tile_base, keys_out, vals_out, n_reduced);     \
        return true;                                                                                             \
    }
    AFB_ALL_TYPES(X)
#undef X
    return false;
}
}  // namespace afb
