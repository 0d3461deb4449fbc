// device_ops.cuh -- element types, op rules and warp/block primitives shared by every kernel.
//
// Op semantics follow the reference's functors (behaviour, not code):
//   /root/reference/src/backend/common/Binary.hpp:29-139     init() / combine per op
//   /root/reference/src/backend/common/Transform.hpp:30-62   per-element pre-op (NaN -> identity for min/max,
//                                                            x != 0 for count/any/all)
//   /root/reference/src/backend/cpu/math.hpp:94-125          identities (+-inf for floats, type max/lowest for ints)
//   /root/reference/src/backend/cpu/kernel/ireduce.hpp:21-63 (value,index) ordering: compare as double,
//                                                            min keeps the LAST equal position, max the FIRST
// CUDA-backend-only quirks (device maxval<T>() = 1<<(bits-1), half math in __half) are deliberately NOT
// reproduced: the parity oracle is the CPU backend.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../common.hpp"

namespace afb {

// ---------------------------------------------------------------- element types
struct cfloat { float x, y; };
struct cdouble { double x, y; };
typedef signed char schar;
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef long long intl;
typedef unsigned long long uintl;
struct b8_t { signed char v; };  // reference b8 is (signed) `char`; a distinct type keeps dispatch unambiguous

template<typename T> struct is_cplx { static constexpr bool value = false; };
template<> struct is_cplx<cfloat> { static constexpr bool value = true; };
template<> struct is_cplx<cdouble> { static constexpr bool value = true; };
template<typename T> struct is_fp { static constexpr bool value = false; };
template<> struct is_fp<float> { static constexpr bool value = true; };
template<> struct is_fp<double> { static constexpr bool value = true; };
template<typename A, typename B> struct same_t { static constexpr bool value = false; };
template<typename A> struct same_t<A, A> { static constexpr bool value = true; };

template<af_dtype> struct type_of;
template<> struct type_of<f32> { typedef float T; };
template<> struct type_of<f64> { typedef double T; };
template<> struct type_of<c32> { typedef cfloat T; };
template<> struct type_of<c64> { typedef cdouble T; };
template<> struct type_of<s32> { typedef int T; };
template<> struct type_of<u32> { typedef uint T; };
template<> struct type_of<s64> { typedef intl T; };
template<> struct type_of<u64> { typedef uintl T; };
template<> struct type_of<s16> { typedef short T; };
template<> struct type_of<u16> { typedef ushort T; };
template<> struct type_of<s8> { typedef schar T; };
template<> struct type_of<u8> { typedef uchar T; };
template<> struct type_of<b8> { typedef b8_t T; };
template<> struct type_of<f16> { typedef __half T; };

// compute type: half accumulates in float (cpu/types.hpp:47-62)
template<typename T> struct compute_of { typedef T type; };
template<> struct compute_of<__half> { typedef float type; };

#define AFB_DI __device__ __forceinline__

// ---------------------------------------------------------------- conversions
template<typename To, typename Ti> struct Conv {
    AFB_DI static To go(Ti x) { return (To)x; }
};
template<typename T> struct Conv<T, T> { AFB_DI static T go(T x) { return x; } };
template<> struct Conv<float, __half> { AFB_DI static float go(__half x) { return __half2float(x); } };
template<> struct Conv<__half, float> { AFB_DI static __half go(float x) { return __float2half_rn(x); } };
template<> struct Conv<double, __half> { AFB_DI static double go(__half x) { return (double)__half2float(x); } };
template<typename To> struct Conv<To, b8_t> { AFB_DI static To go(b8_t x) { return (To)x.v; } };
template<> struct Conv<b8_t, b8_t> { AFB_DI static b8_t go(b8_t x) { return x; } };
template<typename To, typename Ti> AFB_DI To conv(Ti x) { return Conv<To, Ti>::go(x); }

// ---------------------------------------------------------------- predicates
template<typename T> AFB_DI bool is_nan(T) { return false; }
template<> AFB_DI bool is_nan<float>(float v) { return v != v; }
template<> AFB_DI bool is_nan<double>(double v) { return v != v; }
template<> AFB_DI bool is_nan<__half>(__half v) { return __hisnan(v); }
template<> AFB_DI bool is_nan<cfloat>(cfloat v) { return v.x != v.x || v.y != v.y; }
template<> AFB_DI bool is_nan<cdouble>(cdouble v) { return v.x != v.x || v.y != v.y; }

template<typename T> AFB_DI bool nonzero(T v) { return v != (T)0; }
template<> AFB_DI bool nonzero<__half>(__half v) { return __half2float(v) != 0.0f; }  // NaN != 0 -> true
template<> AFB_DI bool nonzero<cfloat>(cfloat v) { return v.x != 0.0f || v.y != 0.0f; }
template<> AFB_DI bool nonzero<cdouble>(cdouble v) { return v.x != 0.0 || v.y != 0.0; }
template<> AFB_DI bool nonzero<b8_t>(b8_t v) { return v.v != 0; }

// |z| exactly as the host computes it: hypot, un-fused
AFB_DI float mag(cfloat v) { return hypotf(v.x, v.y); }
AFB_DI double mag(cdouble v) { return hypot(v.x, v.y); }

// the `double` key ireduce orders by (kernel/ireduce.hpp:21-27)
template<typename T> AFB_DI double order_key(T v) { return (double)v; }
template<> AFB_DI double order_key<__half>(__half v) { return (double)__half2float(v); }
template<> AFB_DI double order_key<b8_t>(b8_t v) { return (double)(v.v > 0); }
template<> AFB_DI double order_key<cfloat>(cfloat v) { return (double)mag(v); }
template<> AFB_DI double order_key<cdouble>(cdouble v) { return mag(v); }

// ---------------------------------------------------------------- identities
template<typename T> struct Lim;
template<> struct Lim<float> { AFB_DI static float hi() { return __int_as_float(0x7f800000); } AFB_DI static float lo() { return __int_as_float(0xff800000); } };
template<> struct Lim<double> { AFB_DI static double hi() { return __longlong_as_double(0x7ff0000000000000LL); } AFB_DI static double lo() { return __longlong_as_double(0xfff0000000000000LL); } };
template<> struct Lim<__half> { AFB_DI static __half hi() { return __ushort_as_half(0x7c00); } AFB_DI static __half lo() { return __ushort_as_half(0xfc00); } };
template<> struct Lim<int> { AFB_DI static int hi() { return 2147483647; } AFB_DI static int lo() { return -2147483647 - 1; } };
template<> struct Lim<uint> { AFB_DI static uint hi() { return 0xffffffffu; } AFB_DI static uint lo() { return 0u; } };
template<> struct Lim<intl> { AFB_DI static intl hi() { return 0x7fffffffffffffffLL; } AFB_DI static intl lo() { return -0x7fffffffffffffffLL - 1; } };
template<> struct Lim<uintl> { AFB_DI static uintl hi() { return 0xffffffffffffffffULL; } AFB_DI static uintl lo() { return 0ULL; } };
template<> struct Lim<short> { AFB_DI static short hi() { return 32767; } AFB_DI static short lo() { return -32768; } };
template<> struct Lim<ushort> { AFB_DI static ushort hi() { return 65535; } AFB_DI static ushort lo() { return 0; } };
template<> struct Lim<schar> { AFB_DI static schar hi() { return 127; } AFB_DI static schar lo() { return -128; } };
template<> struct Lim<uchar> { AFB_DI static uchar hi() { return 255; } AFB_DI static uchar lo() { return 0; } };
template<> struct Lim<b8_t> { AFB_DI static b8_t hi() { return b8_t{1}; } AFB_DI static b8_t lo() { return b8_t{0}; } };  // Binary<char,min/max>::init
template<> struct Lim<cfloat> { AFB_DI static cfloat hi() { return cfloat{Lim<float>::hi(), 0.f}; } AFB_DI static cfloat lo() { return cfloat{0.f, 0.f}; } };
template<> struct Lim<cdouble> { AFB_DI static cdouble hi() { return cdouble{Lim<double>::hi(), 0.}; } AFB_DI static cdouble lo() { return cdouble{0., 0.}; } };

// ---------------------------------------------------------------- op rules
// Rule<OP,Ti,T>: T is the accumulation type.  join(a, b): `a` covers EARLIER positions than `b`
// (matters only for the sign of a zero / which of two equal-magnitude complex values survives).
template<int OP, typename Ti, typename T> struct Rule;

template<typename T> AFB_DI T add_(T a, T b) { return a + b; }
template<> AFB_DI cfloat add_<cfloat>(cfloat a, cfloat b) { return cfloat{a.x + b.x, a.y + b.y}; }
template<> AFB_DI cdouble add_<cdouble>(cdouble a, cdouble b) { return cdouble{a.x + b.x, a.y + b.y}; }
template<typename T> AFB_DI T mul_(T a, T b) { return a * b; }
// (a+bi)(c+di) with every product and sum individually rounded (no FMA contraction), as the host does
template<> AFB_DI cfloat mul_<cfloat>(cfloat a, cfloat b) {
    return cfloat{__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x))};
}
template<> AFB_DI cdouble mul_<cdouble>(cdouble a, cdouble b) {
    return cdouble{__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x))};
}
template<typename T> AFB_DI T one_() { return (T)1; }
template<> AFB_DI cfloat one_<cfloat>() { return cfloat{1.f, 0.f}; }
template<> AFB_DI cdouble one_<cdouble>() { return cdouble{1., 0.}; }
template<typename T> AFB_DI T zero_() { return (T)0; }
template<> AFB_DI cfloat zero_<cfloat>() { return cfloat{0.f, 0.f}; }
template<> AFB_DI cdouble zero_<cdouble>() { return cdouble{0., 0.}; }
template<> AFB_DI b8_t zero_<b8_t>() { return b8_t{0}; }
template<> AFB_DI b8_t one_<b8_t>() { return b8_t{1}; }

// static_cast<T>(x) of Transform<Ti,To,op> (complex <- complex only)
template<typename T, typename Ti> struct Widen { AFB_DI static T go(Ti x) { return conv<T, Ti>(x); } };

template<typename Ti, typename T> struct Rule<OP_ADD, Ti, T> {
    AFB_DI static T init() { return zero_<T>(); }
    AFB_DI static T pre(Ti x) { return Widen<T, Ti>::go(x); }
    AFB_DI static T join(T a, T b) { return add_<T>(a, b); }
};
template<typename Ti, typename T> struct Rule<OP_MUL, Ti, T> {
    AFB_DI static T init() { return one_<T>(); }
    AFB_DI static T pre(Ti x) { return Widen<T, Ti>::go(x); }
    AFB_DI static T join(T a, T b) { return mul_<T>(a, b); }
};
template<typename Ti, typename T> struct Rule<OP_NOTZERO, Ti, T> {
    AFB_DI static T init() { return zero_<T>(); }
    AFB_DI static T pre(Ti x) { return (T)nonzero<Ti>(x); }
    AFB_DI static T join(T a, T b) { return a + b; }
};
template<typename Ti> struct Rule<OP_OR, Ti, b8_t> {
    AFB_DI static b8_t init() { return b8_t{0}; }
    AFB_DI static b8_t pre(Ti x) { return b8_t{(signed char)nonzero<Ti>(x)}; }
    AFB_DI static b8_t join(b8_t a, b8_t b) { return b8_t{(signed char)(a.v || b.v)}; }
};
template<typename Ti> struct Rule<OP_AND, Ti, b8_t> {
    AFB_DI static b8_t init() { return b8_t{1}; }
    AFB_DI static b8_t pre(Ti x) { return b8_t{(signed char)nonzero<Ti>(x)}; }
    AFB_DI static b8_t join(b8_t a, b8_t b) { return b8_t{(signed char)(a.v && b.v)}; }
};

// min/max: the reference evaluates op(x_i, acc) serially = std::min(x_i, acc) = (acc < x_i) ? acc : x_i, i.e. on a
// tie the LATER element survives.  join(a = earlier, b = later) keeps that rule so scans are bit-identical
// even for +0/-0; complex compares magnitudes: |later| < |earlier| ? later : earlier  (ties keep the EARLIER one).
template<typename T> AFB_DI T min_(T a, T b) { return (a < b) ? a : b; }
template<typename T> AFB_DI T max_(T a, T b) { return (b < a) ? a : b; }
template<> AFB_DI b8_t min_<b8_t>(b8_t a, b8_t b) { return b8_t{(signed char)((a.v > 0) && (b.v > 0))}; }
template<> AFB_DI b8_t max_<b8_t>(b8_t a, b8_t b) { return b8_t{(signed char)((a.v > 0) || (b.v > 0))}; }
template<> AFB_DI cfloat min_<cfloat>(cfloat a, cfloat b) { return mag(b) < mag(a) ? b : a; }
template<> AFB_DI cdouble min_<cdouble>(cdouble a, cdouble b) { return mag(b) < mag(a) ? b : a; }
template<> AFB_DI cfloat max_<cfloat>(cfloat a, cfloat b) { return mag(b) > mag(a) ? b : a; }
template<> AFB_DI cdouble max_<cdouble>(cdouble a, cdouble b) { return mag(b) > mag(a) ? b : a; }

template<typename Ti, typename T> struct Rule<OP_MIN, Ti, T> {
    AFB_DI static T init() { return Lim<T>::hi(); }
    AFB_DI static T pre(Ti x) { return is_nan<Ti>(x) ? init() : conv<T, Ti>(x); }
    AFB_DI static T join(T a, T b) { return min_<T>(a, b); }
};
template<typename Ti, typename T> struct Rule<OP_MAX, Ti, T> {
    AFB_DI static T init() { return Lim<T>::lo(); }
    AFB_DI static T pre(Ti x) { return is_nan<Ti>(x) ? init() : conv<T, Ti>(x); }
    AFB_DI static T join(T a, T b) { return max_<T>(a, b); }
};

// `in_val = IS_NAN(in_val) ? nanval : in_val` of cpu/kernel/reduce.hpp:57.  For real types the
// conditional's type is double, so every element round-trips through double (lossy for 64-bit ints > 2^53).
template<typename T> AFB_DI T nan_swap(T v, double nanval) { return (T)(is_nan<T>(v) ? nanval : (double)v); }
template<> AFB_DI cfloat nan_swap<cfloat>(cfloat v, double nanval) { return is_nan(v) ? cfloat{(float)nanval, 0.f} : v; }
template<> AFB_DI cdouble nan_swap<cdouble>(cdouble v, double nanval) { return is_nan(v) ? cdouble{nanval, 0.} : v; }
template<> AFB_DI b8_t nan_swap<b8_t>(b8_t v, double) { return b8_t{(signed char)(double)v.v}; }

// ---------------------------------------------------------------- (value, index) accumulator for ireduce
template<typename T> struct Arg {
    T val;
    uint idx;
};

// ---------------------------------------------------------------- shuffles of arbitrary PODs
template<typename T> AFB_DI T shfl_down(T v, int delta) {
    constexpr int W = (sizeof(T) + 3) / 4;
    union { T t; uint32_t w[W]; } u;
    u.w[W - 1] = 0;
    u.t = v;
#pragma unroll
    for (int i = 0; i < W; i++) u.w[i] = __shfl_down_sync(0xffffffffu, u.w[i], delta);
    return u.t;
}
template<typename T> AFB_DI T shfl_up(T v, int delta) {
    constexpr int W = (sizeof(T) + 3) / 4;
    union { T t; uint32_t w[W]; } u;
    u.w[W - 1] = 0;
    u.t = v;
#pragma unroll
    for (int i = 0; i < W; i++) u.w[i] = __shfl_up_sync(0xffffffffu, u.w[i], delta);
    return u.t;
}
template<typename T> AFB_DI T shfl_idx(T v, int lane) {
    constexpr int W = (sizeof(T) + 3) / 4;
    union { T t; uint32_t w[W]; } u;
    u.w[W - 1] = 0;
    u.t = v;
#pragma unroll
    for (int i = 0; i < W; i++) u.w[i] = __shfl_sync(0xffffffffu, u.w[i], lane);
    return u.t;
}

// ---------------------------------------------------------------- streaming loads
// 128-bit read-only loads that do not allocate in L1 (data is touched once).
struct alignas(16) Vec16 { uint32_t w[4]; };
AFB_DI Vec16 ldg_stream16(const void *p) {
    Vec16 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "l"(p));
    return r;
}
AFB_DI void stg_stream16(void *p, const Vec16 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]) : "memory");
}
template<typename T> AFB_DI T ldg_stream(const T *p) { return __ldg(p); }
template<> AFB_DI cfloat ldg_stream<cfloat>(const cfloat *p) { float2 v = __ldg((const float2 *)p); return cfloat{v.x, v.y}; }
template<> AFB_DI cdouble ldg_stream<cdouble>(const cdouble *p) { double2 v = __ldg((const double2 *)p); return cdouble{v.x, v.y}; }
template<> AFB_DI b8_t ldg_stream<b8_t>(const b8_t *p) { return b8_t{__ldg((const signed char *)p)}; }
template<> AFB_DI __half ldg_stream<__half>(const __half *p) { return __ushort_as_half(__ldg((const unsigned short *)p)); }

// elements per 16-byte vector
template<typename T> struct VecN { static constexpr int value = 16 / (int)sizeof(T); };

// A look-back poll that has not seen its predecessor after this many tries (several seconds) is a protocol
// bug, not a slow neighbour: trap, so that the launch fails loudly instead of hanging the device.
constexpr unsigned kSpinLimit = 1u << 25;

// acquire / release flags for tile descriptors and tickets
AFB_DI uint ld_acquire(const uint *p) {
    uint v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
AFB_DI void st_release(uint *p, uint v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
AFB_DI uint ld_relaxed(const uint *p) {
    uint v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// system-scope release / acquire: flags read by the host (pinned memory) or by peer GPUs (NVLink-mapped memory)
AFB_DI void st_release_sys(unsigned *p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
AFB_DI unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// ONE thread: the 10 payload words of a host result -> the pinned wire record (common.hpp: HostRet), flag in line: five
// 16-byte stores, no fence
AFB_DI void host_ret_send(void *wire, unsigned seq, const unsigned *w) {
#pragma unroll
    for (int j = 0; j < 5; j++)
        asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"((char *)wire + 16 * j), "r"(w[2 * j]), "r"(seq),
                     "r"(w[2 * j + 1]), "r"(seq)
                     : "memory");
}
AFB_DI unsigned long long ld_relaxed64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
AFB_DI void st_relaxed64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace afb
