// reduce_sum.cu -- af_sum (+_nan, _all) instantiations with the reference's output promotion.
#include "reduce_dispatch.cuh"
namespace afb {
bool reduce_family_sum(const LaunchCtx &c, Op, af_dtype ti, af_dtype to, const View &in, int dim, bool all, bool cn,
                       double nv, void *out, const HostRet *ret, const ShardLink *link) {
#define X(ci, Ti, co, To)                                                \
    if (ti == ci && to == co) {                                          \
        reduce_one<OP_ADD, Ti, To>(c, in, dim, all, cn, nv, out, ret, link);        \
        return true;                                                     \
    }
    AFB_SUM_PAIRS(X)
#undef X
    return false;
}
}  // namespace afb
