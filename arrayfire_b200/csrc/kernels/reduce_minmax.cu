// reduce_minmax.cu -- af_min / af_max (+_all) instantiations: T -> T for all 14 dtypes.
#include "reduce_dispatch.cuh"
namespace afb {
// complex: magnitude order with first-occurrence ties (FirstPolicy)
template<int OP, typename T>
static void reduce_first_occurrence(const LaunchCtx &c, const View &in, int dim, bool all, void *out, const HostRet *ret, const ShardLink *link) {
    typedef FirstPolicy<OP, T> P;
    typename P::Params prm;
    prm.change_nan = 0;
    prm.nanval = 0;
    typename P::Out o;
    o.val = (T *)out;
    if (all) run_all<P>(c, in, prm, o, ret, link);
    else run_reduce<P>(c, in, dim, prm, o);
}
bool reduce_family_minmax(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, int dim, bool all,
                          bool cn, double nv, void *out, const HostRet *ret, const ShardLink *link) {
    if (ti != to) return false;
    if (ti == c32 || ti == c64) {
        if (ti == c32) {
            if (op == OP_MIN) reduce_first_occurrence<OP_MIN, cfloat>(c, in, dim, all, out, ret, link);
            else reduce_first_occurrence<OP_MAX, cfloat>(c, in, dim, all, out, ret, link);
        } else {
            if (op == OP_MIN) reduce_first_occurrence<OP_MIN, cdouble>(c, in, dim, all, out, ret, link);
            else reduce_first_occurrence<OP_MAX, cdouble>(c, in, dim, all, out, ret, link);
        }
        return true;
    }
#define X(code, T)                                                                        \
    if (ti == code) {                                                                     \
        if (op == OP_MIN) reduce_one<OP_MIN, T, T>(c, in, dim, all, cn, nv, out, ret, link);         \
        else reduce_one<OP_MAX, T, T>(c, in, dim, all, cn, nv, out, ret, link);                      \
        return true;                                                                      \
    }
    AFB_REAL_TYPES(X)
#undef X
    return false;
}
}  // namespace afb
