// ireduce.cu -- af_imin / af_imax (+_all): (value, u32 index) reductions for all 14 dtypes.
#include "reduce_host.cuh"
#include "reduce_dispatch.cuh"
namespace afb {
template<int OP, typename T>
static void ireduce_one(const LaunchCtx &c, const View &in, int dim, bool all, void *out_val, unsigned *out_loc) {
    typedef ArgPolicy<OP, T> P;
    typename P::Params prm;
    typename P::Out o;
    o.val = (T *)out_val;
    o.loc = out_loc;
    if (all) run_all<P>(c, in, prm, o);
    else run_reduce<P>(c, in, dim, prm, o);
}
static bool route(const LaunchCtx &c, Op op, af_dtype t, const View &in, int dim, bool all, void *v, unsigned *l) {
    if (op != OP_MIN && op != OP_MAX) return false;
#define X(code, T)                                                              \
    if (t == code) {                                                            \
        if (op == OP_MIN) ireduce_one<OP_MIN, T>(c, in, dim, all, v, l);        \
        else ireduce_one<OP_MAX, T>(c, in, dim, all, v, l);                     \
        return true;                                                            \
    }
    AFB_ALL_TYPES(X)
#undef X
    return false;
}
bool launch_ireduce(const LaunchCtx &c, Op op, af_dtype t, const View &in, int dim, void *v, unsigned *l) {
    return route(c, op, t, in, dim, false, v, l);
}
bool launch_ireduce_all(const LaunchCtx &c, Op op, af_dtype t, const View &in, void *v, unsigned *l) {
    return route(c, op, t, in, 0, true, v, l);
}
}  // namespace afb
