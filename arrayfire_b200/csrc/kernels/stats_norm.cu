// stats_norm.cu -- the reductions af_norm is made of (/root/reference/src/api/c/norm.cpp:45-117): sum or max of |x|,
// |x|^2 or |x|^p, whole array or along a dimension, with the element-wise part fused into the load.
#include "stats_kernels.cuh"
namespace afb {

#define AFB_NORM_TYPES(X) \
    X(f32, float, float) X(f64, double, double) X(c32, cfloat, float) X(c64, cdouble, double) X(f16, __half, float)

template<int OP>
static bool norm_family(const LaunchCtx &c, af_dtype ti, const View &in, int dim, int pre, double p, void *out,
                        const HostRet *ret) {
#define X(code, Ti, R)                                        \
    if (ti == code) {                                         \
        typedef NormPolicy<OP, Ti, R> P;                      \
        typename P::Params prm;                               \
        prm.pre = pre;                                        \
        prm.p = p;                                            \
        typename P::Out o;                                    \
        o.val = (R *)out;                                     \
        if (dim < 0) run_all<P>(c, in, prm, o, ret);          \
        else run_reduce<P>(c, in, dim, prm, o);               \
        return true;                                          \
    }
    AFB_NORM_TYPES(X)
#undef X
    return false;
}

bool launch_norm_reduce(const LaunchCtx &c, Op op, af_dtype ti, const View &in, int dim, int pre, double p, void *out,
                        const HostRet *ret) {
    if (op == OP_ADD) return norm_family<OP_ADD>(c, ti, in, dim, pre, p, out, ret);
    if (op == OP_MAX) return norm_family<OP_MAX>(c, ti, in, dim, pre, p, out, ret);
    return false;
}

}  // namespace afb
