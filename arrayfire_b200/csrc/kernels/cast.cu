// cast.cu -- af_cast between any two of the 14 dtypes (dense result of a strided view).
//
// Needed around the hot path because the reference's C++ wrapper converts through it (`array::as()`,
// /root/reference/src/api/cpp/array.cpp:1135-1139 -> af_cast, src/api/c/cast.cpp:34-84) and because the statistics
// rows built on the reductions promote their inputs (src/api/c/mean.cpp:65-107).  Semantics follow
// /root/reference/src/backend/cuda/cast.hpp:32-86: a C-style conversion per element, real -> complex fills a zero
// imaginary part, complex -> real is rejected by the C layer (cast.cpp:43-45), anything -> b8 is `x != 0`.
#include "scan_host.cuh"

namespace afb {

// one conversion rule, resolved at compile time (real source types only reach the real branch)
template<typename To, typename Ti> AFB_DI To cast_one(Ti x) {
    if constexpr (same_t<To, Ti>::value && !same_t<To, b8_t>::value) {
        return x;
    } else if constexpr (is_cplx<Ti>::value) {            // complex -> complex (complex -> real never gets here)
        typedef decltype(To().x) R;
        return To{(R)x.x, (R)x.y};
    } else if constexpr (same_t<To, b8_t>::value) {        // anything -> b8: x != 0
        return b8_t{(signed char)nonzero<Ti>(x)};
    } else if constexpr (is_cplx<To>::value) {             // real -> complex: zero imaginary part
        typedef decltype(To().x) R;
        return To{cast_one<R, Ti>(x), (R)0};
    } else if constexpr (same_t<Ti, __half>::value) {      // half -> real, through float
        return (To)__half2float(x);
    } else if constexpr (same_t<Ti, b8_t>::value) {        // b8 -> real: the stored byte
        return cast_one<To, signed char>(x.v);
    } else if constexpr (same_t<To, __half>::value) {      // real -> half, through float
        return __float2half_rn((float)x);
    } else {
        return (To)x;
    }
}

template<typename Ti, typename To>
__global__ void __launch_bounds__(256) cast_any_kernel(const Ti *__restrict__ in, To *__restrict__ out, dim_t d0, dim_t d1,
                                                        dim_t d2, dim_t n, dim_t s1, dim_t s2, dim_t s3, int linear) {
    for (dim_t i = (dim_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (dim_t)gridDim.x * blockDim.x) {
        dim_t src = i;
        if (!linear) {
            const dim_t x = i % d0, r = i / d0;
            const dim_t y = r % d1, r2 = r / d1;
            const dim_t z = r2 % d2, w = r2 / d2;
            src = x + y * s1 + z * s2 + w * s3;
        }
        out[i] = cast_one<To, Ti>(ldg_stream<Ti>(in + src));
    }
}

template<typename Ti, typename To> static void cast_any(const LaunchCtx &c, const View &in, void *out) {
    const dim_t n = in.dims[0] * in.dims[1] * in.dims[2] * in.dims[3];
    if (n == 0) return;
    dim_t blocks = (n + 255) / 256;
    const dim_t cap = (dim_t)c.sms * 16;
    if (blocks > cap) blocks = cap;
    cast_any_kernel<Ti, To><<<(unsigned)blocks, 256, 0, c.stream>>>((const Ti *)in.ptr, (To *)out, in.dims[0], in.dims[1],
                                                                     in.dims[2], n, in.strides[1], in.strides[2],
                                                                     in.strides[3], view_is_linear(in) ? 1 : 0);
    note_launches(1);
    post_launch_check("cast_any_kernel");
}

#define AFB_ALL_TYPES(X)                                                                                     \
    X(f32, float) X(f64, double) X(c32, cfloat) X(c64, cdouble) X(s32, int) X(u32, uint) X(s64, intl) X(u64, uintl) \
    X(s16, short) X(u16, ushort) X(s8, schar) X(u8, uchar) X(b8, b8_t) X(f16, __half)
#define AFB_REAL_TYPES(X)                                                                                    \
    X(f32, float) X(f64, double) X(s32, int) X(u32, uint) X(s64, intl) X(u64, uintl)                          \
    X(s16, short) X(u16, ushort) X(s8, schar) X(u8, uchar) X(b8, b8_t) X(f16, __half)

template<typename Ti> static bool cast_from_real(const LaunchCtx &c, af_dtype to, const View &in, void *out) {
#define X(code, To)                     \
    if (to == code) {                   \
        cast_any<Ti, To>(c, in, out);   \
        return true;                    \
    }
    AFB_ALL_TYPES(X)
#undef X
    return false;
}
template<typename Ti> static bool cast_from_cplx(const LaunchCtx &c, af_dtype to, const View &in, void *out) {
    if (to == c32) { cast_any<Ti, cfloat>(c, in, out); return true; }
    if (to == c64) { cast_any<Ti, cdouble>(c, in, out); return true; }
    return false;   // complex -> real: rejected (cast.cpp:43-45)
}

bool launch_cast_any(const LaunchCtx &c, af_dtype ti, af_dtype to, const View &in, void *out) {
#define X(code, Ti) \
    if (ti == code) return cast_from_real<Ti>(c, to, in, out);
    AFB_REAL_TYPES(X)
#undef X
    if (ti == c32) return cast_from_cplx<cfloat>(c, to, in, out);
    if (ti == c64) return cast_from_cplx<cdouble>(c, to, in, out);
    return false;
}

}  // namespace afb
