// reduce_dispatch.cuh -- (op, Ti, To) -> kernel-instantiation tables for the value reductions.
// The triples are exactly those the reference instantiates
// (/root/reference/src/backend/cpu/reduce.cpp:138-258); anything else returns false
// and the API layer reports AF_ERR_TYPE.  Split over several .cu files so nvcc compiles them in parallel.
#pragma once

#include "reduce_host.cuh"

namespace afb {

template<int OP, typename Ti, typename To>
inline void reduce_one(const LaunchCtx &c, const View &in, int dim, bool all, bool change_nan, double nanval, void *out,
                       const HostRet *ret = nullptr, const ShardLink *link = nullptr) {
    typedef ValPolicy<OP, Ti, To> P;
    typename P::Params prm;
    prm.change_nan = change_nan ? 1 : 0;
    prm.nanval = nanval;
    typename P::Out o;
    o.val = (To *)out;
    if (all) run_all<P>(c, in, prm, o, ret, link);
    else run_reduce<P>(c, in, dim, prm, o);
}

#define AFB_ALL_TYPES(X)                                                                              \
    X(f32, float) X(f64, double) X(c32, cfloat) X(c64, cdouble) X(s32, int) X(u32, uint) X(s64, intl) \
    X(u64, uintl) X(s16, short) X(u16, ushort) X(s8, schar) X(u8, uchar) X(b8, b8_t) X(f16, __half)

#define AFB_REAL_TYPES(X)                                                                             \
    X(f32, float) X(f64, double) X(s32, int) X(u32, uint) X(s64, intl) X(u64, uintl) X(s16, short)    \
    X(u16, ushort) X(s8, schar) X(u8, uchar) X(b8, b8_t) X(f16, __half)

#define AFB_SUM_PAIRS(X)                                                                              \
    X(f32, float, f32, float) X(f64, double, f64, double) X(c32, cfloat, c32, cfloat)                 \
    X(c64, cdouble, c64, cdouble) X(s32, int, s32, int) X(u32, uint, u32, uint) X(s64, intl, s64, intl) \
    X(u64, uintl, u64, uintl) X(s16, short, s32, int) X(u16, ushort, u32, uint) X(s8, schar, s32, int) \
    X(u8, uchar, u32, uint) X(f16, __half, f32, float)

// each family lives in its own translation unit
bool reduce_family_minmax(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, int dim, bool all,
                          bool change_nan, double nanval, void *out, const HostRet *ret, const ShardLink *link);
bool reduce_family_sum(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, int dim, bool all,
                       bool change_nan, double nanval, void *out, const HostRet *ret, const ShardLink *link);
bool reduce_family_prod(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, int dim, bool all,
                        bool change_nan, double nanval, void *out, const HostRet *ret, const ShardLink *link);
bool reduce_family_logic(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, int dim, bool all,
                         bool change_nan, double nanval, void *out, const HostRet *ret, const ShardLink *link);

}  // namespace afb
