// stats_mean.cu -- af_mean / af_mean_all instantiations: (Ti -> To) pairs of /root/reference/src/api/c/mean.cpp:71-86,155-180
// (integers up to 32 bits and b8 average in f32, 64-bit integers in f64, f16 computes in f32).
#include "stats_kernels.cuh"
namespace afb {

#define AFB_STAT_TYPES(X)                                                                                       \
    X(f32, float, float, float) X(f64, double, double, double) X(c32, cfloat, cfloat, cfloat)                   \
    X(c64, cdouble, cdouble, cdouble) X(s32, int, float, float) X(u32, uint, float, float)                     \
    X(s64, intl, double, double) X(u64, uintl, double, double) X(s16, short, float, float)                     \
    X(u16, ushort, float, float) X(s8, schar, float, float) X(u8, uchar, float, float) X(b8, b8_t, float, float) \
    X(f16, __half, float, __half)

bool launch_mean(const LaunchCtx &c, af_dtype ti, const View &in, int dim, void *out) {
#define X(code, Ti, Tc, To)                                   \
    if (ti == code) {                                         \
        typedef MeanPolicy<Ti, Tc, To> P;                     \
        typename P::Out o;                                    \
        o.val = (To *)out;                                    \
        run_reduce<P>(c, in, dim, typename P::Params(), o);   \
        return true;                                          \
    }
    AFB_STAT_TYPES(X)
#undef X
    return false;
}

// the scalar form keeps f16 means in f32 (mean.cpp:168-170)
bool launch_mean_all(const LaunchCtx &c, af_dtype ti, const View &in, void *out, const HostRet *ret) {
#define X(code, Ti, Tc, To)                                   \
    if (ti == code) {                                         \
        typedef MeanPolicy<Ti, Tc, Tc> P;                     \
        typename P::Out o;                                    \
        o.val = (Tc *)out;                                    \
        run_all<P>(c, in, typename P::Params(), o, ret);      \
        return true;                                          \
    }
    AFB_STAT_TYPES(X)
#undef X
    return false;
}

}  // namespace afb
