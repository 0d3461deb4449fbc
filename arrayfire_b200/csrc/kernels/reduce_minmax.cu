// reduce_minmax.cu -- af_min / af_max (+_all) instantiations: T -> T for all 14 dtypes.
#include "reduce_dispatch.cuh"
namespace afb {
bool reduce_family_minmax(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, int dim, bool all,
                          bool cn, double nv, void *out) {
    if (ti != to) return false;
#define X(code, T)                                                                        \
    if (ti == code) {                                                                     \
        if (op == OP_MIN) reduce_one<OP_MIN, T, T>(c, in, dim, all, cn, nv, out);         \
        else reduce_one<OP_MAX, T, T>(c, in, dim, all, cn, nv, out);                      \
        return true;                                                                      \
    }
    AFB_ALL_TYPES(X)
#undef X
    return false;
}
}  // namespace afb
