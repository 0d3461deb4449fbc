// stats_var.cu -- af_var / af_stdev / af_meanvar (+ _all) instantiations: the (Ti -> To) pairs of
// /root/reference/src/api/c/var.cpp:208-262 and stdev.cpp:101-160; ONE pass over the input (Welford / pairwise merge).
#include "stats_kernels.cuh"
namespace afb {

#define AFB_STAT_TYPES(X)                                                                                       \
    X(f32, float, float, float) X(f64, double, double, double) X(c32, cfloat, cfloat, cfloat)                   \
    X(c64, cdouble, cdouble, cdouble) X(s32, int, float, float) X(u32, uint, float, float)                     \
    X(s64, intl, double, double) X(u64, uintl, double, double) X(s16, short, float, float)                     \
    X(u16, ushort, float, float) X(s8, schar, float, float) X(u8, uchar, float, float) X(b8, b8_t, float, float) \
    X(f16, __half, float, __half)

bool launch_var(const LaunchCtx &c, af_dtype ti, const View &in, int dim, void *out_var, void *out_mean, double norm,
                bool take_sqrt) {
#define X(code, Ti, Tc, To)                                   \
    if (ti == code) {                                         \
        typedef VarPolicy<Ti, Tc, To> P;                      \
        typename P::Out o;                                    \
        o.val = (To *)out_var;                                \
        o.mean = (To *)out_mean;                              \
        o.norm = norm;                                        \
        o.take_sqrt = take_sqrt ? 1 : 0;                      \
        run_reduce<P>(c, in, dim, typename P::Params(), o);   \
        return true;                                          \
    }
    AFB_STAT_TYPES(X)
#undef X
    return false;
}

bool launch_var_all(const LaunchCtx &c, af_dtype ti, const View &in, void *out, double norm, bool take_sqrt,
                    const HostRet *ret) {
#define X(code, Ti, Tc, To)                                   \
    if (ti == code) {                                         \
        typedef VarPolicy<Ti, Tc, Tc> P;                      \
        typename P::Out o;                                    \
        o.val = (Tc *)out;                                    \
        o.mean = nullptr;                                     \
        o.norm = norm;                                        \
        o.take_sqrt = take_sqrt ? 1 : 0;                      \
        run_all<P>(c, in, typename P::Params(), o, ret);      \
        return true;                                          \
    }
    AFB_STAT_TYPES(X)
#undef X
    return false;
}

}  // namespace afb
