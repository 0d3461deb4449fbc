// oracle.cpp -- CPU restatement of the reference's reduce / ireduce / scan /
// scan_by_key / where semantics.
//
// TEST INFRASTRUCTURE ONLY (see oracle_abi.h).  This is the checker the CUDA
// path is compared against; it is never linked into, loaded by or called from
// libafb200.so.  Parity status: PINNED -- validated (tests/test_oracle_vs_ref.py)
// against the reference's own CPU kernel templates compiled from /root/reference
// (oracle/_ref/libafref.so, built by oracle/Makefile) and against the inline
// known-answer vectors of the reference's tests (tests/golden/).
//
// What is restated, and from where (paths relative to /root/reference):
//   op semantics       src/backend/common/Binary.hpp:29-139      (init + combine)
//                      src/backend/common/Transform.hpp:30-62    (per-element pre-op)
//                      src/backend/cpu/math.hpp:32-125, math.cpp:30-40
//   reduce / reduce_all src/backend/cpu/kernel/reduce.hpp:40-63, 162-200
//   ireduce(_all)      src/backend/cpu/kernel/ireduce.hpp:21-104, src/backend/cpu/ireduce.cpp:59-90
//   scan               src/backend/cpu/kernel/scan.hpp:37-72
//   scan_by_key        src/backend/cpu/kernel/scan_by_key.hpp:43-86
//   where              src/backend/cpu/where.cpp:27-63
//   (op,Ti,To) triples src/backend/cpu/{reduce,ireduce,scan,scan_by_key,where}.cpp instantiation lists
//
// The structure here is deliberately NOT the reference's recursive
// reduce_dim<D> template: every op walks "lines" (the elements along `dim`
// for one fixed choice of the other three coordinates) produced by one
// generic line enumerator, serially, in the reference's accumulation order
// acc = combine(transform(x_i), acc).

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <limits>
#include <type_traits>

#define ORC_PREFIX orc
#include "oracle_abi.h"

namespace {

using cfloat  = std::complex<float>;
using cdouble = std::complex<double>;
using b8_t    = char;  // reference: b8 is plain (signed on x86) `char`
using half_t  = _Float16;
typedef orc_dim_t dim_t;

template<typename T> struct is_cplx : std::false_type {};
template<> struct is_cplx<cfloat> : std::true_type {};
template<> struct is_cplx<cdouble> : std::true_type {};

// compute type inside a reduction: half accumulates in float (common/kernel_type.hpp, cpu/types.hpp:47-62)
template<typename T> struct compute_of { using type = T; };
template<> struct compute_of<half_t> { using type = float; };
template<typename T> using compute_t = typename compute_of<T>::type;

template<typename T> inline bool not_self_equal(const T &v) { return !(v == v); }  // IS_NAN macro

// ---- identities (cpu/math.hpp:94-125) ------------------------------------
template<typename T> inline T hi_identity() {  // maxval<T>
    if constexpr (std::is_floating_point<T>::value || std::is_same<T, half_t>::value)
        return (T)std::numeric_limits<float>::infinity();
    else
        return std::numeric_limits<T>::max();
}
template<typename T> inline T lo_identity() {  // minval<T>
    if constexpr (std::is_floating_point<T>::value || std::is_same<T, half_t>::value)
        return (T)(-std::numeric_limits<float>::infinity());
    else
        return std::numeric_limits<T>::lowest();
}

template<typename T> inline double magnitude(const T &v) { return (double)v; }       // kernel/ireduce.hpp:21-27
inline double magnitude(const b8_t &v) { return (double)(v > 0); }
inline double magnitude(const cfloat &v) { return (double)std::abs(v); }
inline double magnitude(const cdouble &v) { return (double)std::abs(v); }

// ---- Binary<T,op> + Transform<Ti,T,op> -----------------------------------
template<int OP, typename Ti, typename T> struct Rule;

template<typename Ti, typename T> struct Rule<ORC_OP_ADD, Ti, T> {
    static T init() { return T(0); }
    static T pre(Ti x) { return static_cast<T>(x); }
    static T join(T x, T acc) { return x + acc; }
};
template<typename Ti, typename T> struct Rule<ORC_OP_MUL, Ti, T> {
    static T init() { return T(1); }
    static T pre(Ti x) { return static_cast<T>(x); }
    static T join(T x, T acc) { return x * acc; }
};
template<typename Ti, typename T> struct Rule<ORC_OP_NOTZERO, Ti, T> {
    static T init() { return T(0); }
    static T pre(Ti x) { return (T)(x != Ti(0)); }
    static T join(T x, T acc) { return x + acc; }
};
template<typename Ti, typename T> struct Rule<ORC_OP_OR, Ti, T> {
    static T init() { return T(0); }
    static T pre(Ti x) { return (T)(x != Ti(0)); }
    static T join(T x, T acc) { return x || acc; }
};
template<typename Ti, typename T> struct Rule<ORC_OP_AND, Ti, T> {
    static T init() { return T(1); }
    static T pre(Ti x) { return (T)(x != Ti(0)); }
    static T join(T x, T acc) { return x && acc; }
};
template<typename Ti, typename T> struct Rule<ORC_OP_MIN, Ti, T> {
    static T init() {
        if constexpr (std::is_same<T, b8_t>::value) return 1;
        else if constexpr (is_cplx<T>::value) return T(hi_identity<typename T::value_type>());
        else return hi_identity<T>();
    }
    static T pre(Ti x) { return not_self_equal(x) ? init() : T(x); }
    static T join(T x, T acc) {
        if constexpr (std::is_same<T, b8_t>::value) return (T)std::min(x > 0, acc > 0);
        else if constexpr (is_cplx<T>::value) return std::abs(x) < std::abs(acc) ? x : acc;
        else return std::min(x, acc);
    }
};
template<typename Ti, typename T> struct Rule<ORC_OP_MAX, Ti, T> {
    static T init() {
        if constexpr (std::is_same<T, b8_t>::value) return 0;
        else if constexpr (is_cplx<T>::value) return T(typename T::value_type(0));
        else return lo_identity<T>();
    }
    static T pre(Ti x) { return not_self_equal(x) ? init() : T(x); }
    static T join(T x, T acc) {
        if constexpr (std::is_same<T, b8_t>::value) return (T)std::max(x > 0, acc > 0);
        else if constexpr (is_cplx<T>::value) return std::abs(x) > std::abs(acc) ? x : acc;
        else return std::max(x, acc);
    }
};

// `in_val = IS_NAN(in_val) ? nanval : in_val` (kernel/reduce.hpp:57): for real types the
// conditional's common type is double, so EVERY element takes a round trip through double
// when change_nan is set (lossy for 64-bit integers beyond 2^53); complex stays complex.
template<typename T> inline T nan_swap(T v, double nanval) {
    if constexpr (is_cplx<T>::value) return not_self_equal(v) ? T((typename T::value_type)nanval) : v;
    else return (T)(not_self_equal(v) ? nanval : (double)v);
}

// ---- line enumeration ----------------------------------------------------
// Calls fn(in_base, out_base) once per line along `dim`, visiting the other
// coordinates with the lowest dimension fastest.
struct Shape {
    dim_t d[4], s[4];
};
inline void dense_strides(const dim_t d[4], dim_t s[4]) {
    s[0] = 1;
    for (int i = 1; i < 4; i++) s[i] = s[i - 1] * d[i - 1];
}
template<typename F> inline void for_each_line(const Shape &in, const dim_t ostr[4], int dim, F fn) {
    dim_t n[4] = {in.d[0], in.d[1], in.d[2], in.d[3]};
    n[dim] = 1;
    for (dim_t l = 0; l < n[3]; l++)
        for (dim_t k = 0; k < n[2]; k++)
            for (dim_t j = 0; j < n[1]; j++)
                for (dim_t i = 0; i < n[0]; i++)
                    fn(i * in.s[0] + j * in.s[1] + k * in.s[2] + l * in.s[3],
                       i * ostr[0] + j * ostr[1] + k * ostr[2] + l * ostr[3]);
}
inline Shape mk(const dim_t *dims, const dim_t *strides) {
    Shape sh;
    for (int i = 0; i < 4; i++) { sh.d[i] = dims[i]; sh.s[i] = strides[i]; }
    return sh;
}
inline bool empty(const dim_t *d) { return d[0] * d[1] * d[2] * d[3] == 0; }

// ---- reduce --------------------------------------------------------------
template<int OP, typename Ti, typename To>
void reduce_lines(const void *vin, const dim_t *dims, const dim_t *strides, int dim, bool change_nan,
                  double nanval, void *vout) {
    using Tc = compute_t<To>;
    using R  = Rule<OP, Ti, Tc>;
    const Ti *in = (const Ti *)vin;
    To *out      = (To *)vout;
    if (empty(dims)) return;
    Shape sh = mk(dims, strides);
    dim_t od[4] = {dims[0], dims[1], dims[2], dims[3]}, os[4];
    od[dim] = 1;
    dense_strides(od, os);
    const dim_t len = dims[dim], step = strides[dim];
    for_each_line(sh, os, dim, [&](dim_t ib, dim_t ob) {
        Tc acc = R::init();
        for (dim_t i = 0; i < len; i++) {
            Tc v = R::pre(in[ib + i * step]);
            if (change_nan) v = nan_swap(v, nanval);
            acc = R::join(v, acc);
        }
        out[ob] = (To)acc;
    });
}

template<int OP, typename Ti, typename To>
void reduce_everything(const void *vin, const dim_t *dims, const dim_t *strides, bool change_nan,
                       double nanval, void *vout) {
    using Tc = compute_t<To>;
    using R  = Rule<OP, Ti, Tc>;
    const Ti *in = (const Ti *)vin;
    Tc acc       = R::init();
    for (dim_t l = 0; l < dims[3]; l++)
        for (dim_t k = 0; k < dims[2]; k++)
            for (dim_t j = 0; j < dims[1]; j++) {
                const Ti *p = in + j * strides[1] + k * strides[2] + l * strides[3];
                for (dim_t i = 0; i < dims[0]; i++) {
                    Tc v = R::pre(p[i]);
                    if (change_nan) v = nan_swap(v, nanval);
                    acc = R::join(v, acc);
                }
            }
    *(To *)vout = (To)acc;
}

// ---- ireduce ---------------------------------------------------------------
// running (value, index) with the reference's asymmetric tie rules:
//   min keeps the LAST position among equals (idx >  best), kernel/ireduce.hpp:38-45
//   max keeps the FIRST position among equals (idx <= best), kernel/ireduce.hpp:57-62
template<int OP, typename T> struct Tracker {
    T val;
    uint32_t idx;
    Tracker(T first, uint32_t i) : val(first), idx(i) {
        bool isnan;
        if constexpr (is_cplx<T>::value) isnan = std::isnan(first.real()) || std::isnan(first.imag());
        else if constexpr (std::is_floating_point<T>::value) isnan = std::isnan(first);
        else if constexpr (std::is_same<T, half_t>::value) isnan = std::isnan((float)first);
        else isnan = false;
        if (isnan) val = Rule<OP, T, T>::init();
    }
    void offer(T v, uint32_t i) {
        const double a = magnitude(v), b = magnitude(val);
        bool take;
        if (OP == ORC_OP_MIN) take = (a < b) || (a == b && i > idx);
        else take = (a > b) || (a == b && i <= idx);
        if (take) { val = v; idx = i; }
    }
};

template<int OP, typename T>
void ireduce_lines(const void *vin, const dim_t *dims, const dim_t *strides, int dim, void *vout,
                   uint32_t *loc) {
    const T *in = (const T *)vin;
    T *out      = (T *)vout;
    if (empty(dims)) return;
    Shape sh = mk(dims, strides);
    dim_t od[4] = {dims[0], dims[1], dims[2], dims[3]}, os[4];
    od[dim] = 1;
    dense_strides(od, os);
    const dim_t len = dims[dim], step = strides[dim];
    for_each_line(sh, os, dim, [&](dim_t ib, dim_t ob) {
        Tracker<OP, T> t(in[ib], 0);
        for (dim_t i = 0; i < len; i++) t.offer(in[ib + i * step], (uint32_t)i);
        out[ob] = t.val;
        loc[ob] = t.idx;
    });
}

template<int OP, typename T>
void ireduce_everything(const void *vin, const dim_t *dims, const dim_t *strides, void *vout,
                        uint32_t *loc) {
    const T *in = (const T *)vin;
    Tracker<OP, T> t(in[0], 0);
    uint32_t pos = 0;
    for (dim_t l = 0; l < dims[3]; l++)
        for (dim_t k = 0; k < dims[2]; k++)
            for (dim_t j = 0; j < dims[1]; j++) {
                const T *p = in + j * strides[1] + k * strides[2] + l * strides[3];
                for (dim_t i = 0; i < dims[0]; i++) t.offer(p[i], pos++);
            }
    *(T *)vout = t.val;
    *loc       = t.idx;
}

template<int OP, typename T>
void ireduce_ragged_lines(const void *vin, const dim_t *dims, const dim_t *strides, int dim, const uint32_t *rlen,
                          void *vout, uint32_t *loc) {
    const T *in = (const T *)vin;
    T *out      = (T *)vout;
    if (empty(dims)) return;
    Shape sh = mk(dims, strides);
    dim_t od[4] = {dims[0], dims[1], dims[2], dims[3]}, os[4];
    od[dim] = 1;
    dense_strides(od, os);
    const dim_t len = dims[dim], step = strides[dim];
    for_each_line(sh, os, dim, [&](dim_t ib, dim_t ob) {
        Tracker<OP, T> t(in[ib], 0);   // element 0 seeds the result even when the ragged length is 0
        const dim_t lim = std::min<dim_t>(len, (dim_t)rlen[ob]);
        for (dim_t i = 0; i < lim; i++) t.offer(in[ib + i * step], (uint32_t)i);
        out[ob] = t.val;
        loc[ob] = t.idx;
    });
}

// ---- reduce by key -----------------------------------------------------------
// One pass over the keys finds the runs of equal consecutive keys (kernel/reduce.hpp:64-86); every line of `vals`
// along `dim` is then folded run by run: the first element of a run is taken as is, the others as
// acc = combine(transform(x_i), acc) (kernel/reduce.hpp:131-158).
template<typename Tk> long long count_runs(const void *vkeys, dim_t n, void *vkeys_out) {
    const Tk *keys = (const Tk *)vkeys;
    Tk *out        = (Tk *)vkeys_out;
    long long runs = 0;
    for (dim_t i = 0; i < n; i++) {
        if (i == 0 || keys[i] != keys[i - 1]) {
            if (out) out[runs] = keys[i];
            runs++;
        }
    }
    return runs;
}

template<int OP, typename Ti, typename To, typename Tk>
void reduce_by_key_lines(const void *vkeys, const void *vin, const dim_t *dims, const dim_t *strides, int dim,
                         bool change_nan, double nanval, long long runs, void *vout) {
    using Tc = compute_t<To>;
    using R  = Rule<OP, Ti, Tc>;
    const Tk *keys = (const Tk *)vkeys;
    const Ti *in   = (const Ti *)vin;
    To *out        = (To *)vout;
    if (empty(dims)) return;
    Shape sh = mk(dims, strides);
    dim_t od[4] = {dims[0], dims[1], dims[2], dims[3]}, os[4];
    od[dim] = runs;
    dense_strides(od, os);
    const dim_t len = dims[dim], step = strides[dim], ostep = os[dim];
    for_each_line(sh, os, dim, [&](dim_t ib, dim_t ob) {
        dim_t seg = 0;
        Tc acc    = R::init();
        for (dim_t i = 0; i < len; i++) {
            Tc v = R::pre(in[ib + i * step]);
            if (change_nan) v = nan_swap(v, nanval);
            if (i == 0) {
                acc = R::join(v, acc);   // the first run starts from init() (reduce.hpp:137,146-148)
            } else if (keys[i] != keys[i - 1]) {
                out[ob + seg * ostep] = (To)acc;
                seg++;
                acc = v;
            } else {
                acc = R::join(v, acc);
            }
        }
        out[ob + seg * ostep] = (To)acc;
    });
}

// ---- scan ------------------------------------------------------------------
template<int OP, typename Ti, typename To>
void scan_lines(const void *vin, const dim_t *dims, const dim_t *strides, int dim, bool inclusive,
                void *vout) {
    using R      = Rule<OP, Ti, To>;
    const Ti *in = (const Ti *)vin;
    To *out      = (To *)vout;
    if (empty(dims)) return;
    Shape sh = mk(dims, strides);
    dim_t os[4];
    dense_strides(dims, os);
    const dim_t len = dims[dim], istep = strides[dim], ostep = os[dim];
    for_each_line(sh, os, dim, [&](dim_t ib, dim_t ob) {
        To acc = R::init();
        if (!inclusive) out[ob] = R::init();
        for (dim_t i = 0; i < len; i++) {
            acc = R::join(R::pre(in[ib + i * istep]), acc);
            if (inclusive) out[ob + i * ostep] = acc;
            else if (i + 1 < len) out[ob + (i + 1) * ostep] = acc;  // last input is dropped
        }
    });
}

// ---- scan_by_key -------------------------------------------------------------
// A segment is a maximal run of equal consecutive keys along `dim`; every line
// starts a fresh segment (kernel/scan_by_key.hpp:63-84).
template<int OP, typename T, typename Tk>
void scan_by_key_lines(const void *vkey, const dim_t *kstrides, const void *vin, const dim_t *dims,
                       const dim_t *strides, int dim, bool inclusive, void *vout) {
    using R       = Rule<OP, T, T>;
    const T *in   = (const T *)vin;
    const Tk *key = (const Tk *)vkey;
    T *out        = (T *)vout;
    if (empty(dims)) return;
    dim_t os[4];
    dense_strides(dims, os);
    const dim_t len = dims[dim], istep = strides[dim], kstep = kstrides[dim], ostep = os[dim];
    dim_t n[4] = {dims[0], dims[1], dims[2], dims[3]};
    n[dim] = 1;
    for (dim_t l = 0; l < n[3]; l++)
        for (dim_t k = 0; k < n[2]; k++)
            for (dim_t j = 0; j < n[1]; j++)
                for (dim_t i0 = 0; i0 < n[0]; i0++) {
                    const dim_t ib = i0 * strides[0] + j * strides[1] + k * strides[2] + l * strides[3];
                    const dim_t kb = i0 * kstrides[0] + j * kstrides[1] + k * kstrides[2] + l * kstrides[3];
                    const dim_t ob = i0 * os[0] + j * os[1] + k * os[2] + l * os[3];
                    T acc = R::init();
                    for (dim_t i = 0; i < len; i++) {
                        const bool head = (i == 0) || (key[kb + i * kstep] != key[kb + (i - 1) * kstep]);
                        const T x = R::pre(in[ib + i * istep]);
                        if (inclusive) {
                            // reference: a new segment restarts from the bare element (no join with init);
                            // the very first element of a line IS joined with init (scan_by_key.hpp:74-80)
                            acc = (head && i > 0) ? x : R::join(x, acc);
                            out[ob + i * ostep] = acc;
                        } else {
                            // exclusive: heads get the identity, others the running value *before* x_i
                            if (head) acc = R::init();
                            out[ob + i * ostep] = acc;
                            acc = R::join(x, acc);
                        }
                    }
                }
}

// ---- where -------------------------------------------------------------------
template<typename T>
long long where_hits(const void *vin, const dim_t *dims, const dim_t *strides, uint32_t *out) {
    const T *in = (const T *)vin;
    const T zero = T(0);
    long long count = 0;
    uint32_t pos    = 0;
    for (dim_t l = 0; l < dims[3]; l++)
        for (dim_t k = 0; k < dims[2]; k++)
            for (dim_t j = 0; j < dims[1]; j++) {
                const T *p = in + j * strides[1] + k * strides[2] + l * strides[3];
                for (dim_t i = 0; i < dims[0]; i++, pos++)
                    if (p[i] != zero) out[count++] = pos;
            }
    return count;
}

// ---- dispatch ----------------------------------------------------------------
#define FOR_ALL_TYPES(X)                                                                           \
    X(ORC_F32, float) X(ORC_F64, double) X(ORC_C32, cfloat) X(ORC_C64, cdouble) X(ORC_S32, int)    \
    X(ORC_U32, unsigned) X(ORC_S64, long long) X(ORC_U64, unsigned long long) X(ORC_S16, short)    \
    X(ORC_U16, unsigned short) X(ORC_S8, signed char) X(ORC_U8, unsigned char) X(ORC_B8, b8_t)     \
    X(ORC_F16, half_t)
#define FOR_NO_HALF_TYPES(X)                                                                       \
    X(ORC_F32, float) X(ORC_F64, double) X(ORC_C32, cfloat) X(ORC_C64, cdouble) X(ORC_S32, int)    \
    X(ORC_U32, unsigned) X(ORC_S64, long long) X(ORC_U64, unsigned long long) X(ORC_S16, short)    \
    X(ORC_U16, unsigned short) X(ORC_S8, signed char) X(ORC_U8, unsigned char) X(ORC_B8, b8_t)

// (Ti -> To) pairs for add/mul, Appendix A / cpu/reduce.cpp:170-210
#define FOR_SUM_PAIRS(X)                                                                           \
    X(ORC_F32, float, ORC_F32, float) X(ORC_F64, double, ORC_F64, double)                          \
    X(ORC_C32, cfloat, ORC_C32, cfloat) X(ORC_C64, cdouble, ORC_C64, cdouble)                      \
    X(ORC_S32, int, ORC_S32, int) X(ORC_U32, unsigned, ORC_U32, unsigned)                          \
    X(ORC_S64, long long, ORC_S64, long long) X(ORC_U64, unsigned long long, ORC_U64, unsigned long long) \
    X(ORC_S16, short, ORC_S32, int) X(ORC_U16, unsigned short, ORC_U32, unsigned)                  \
    X(ORC_S8, signed char, ORC_S32, int) X(ORC_U8, unsigned char, ORC_U32, unsigned)               \
    X(ORC_F16, half_t, ORC_F32, float)
// scan pairs, cpu/scan.cpp:76-96 (b8 both -> s32 and -> u32 exist; the C API reaches u32)
#define FOR_SCAN_PAIRS(X)                                                                          \
    X(ORC_F32, float, ORC_F32, float) X(ORC_F64, double, ORC_F64, double)                          \
    X(ORC_C32, cfloat, ORC_C32, cfloat) X(ORC_C64, cdouble, ORC_C64, cdouble)                      \
    X(ORC_S32, int, ORC_S32, int) X(ORC_U32, unsigned, ORC_U32, unsigned)                          \
    X(ORC_S64, long long, ORC_S64, long long) X(ORC_U64, unsigned long long, ORC_U64, unsigned long long) \
    X(ORC_S16, short, ORC_S32, int) X(ORC_U16, unsigned short, ORC_U32, unsigned)                  \
    X(ORC_S8, signed char, ORC_S32, int) X(ORC_U8, unsigned char, ORC_U32, unsigned)               \
    X(ORC_B8, b8_t, ORC_U32, unsigned) X(ORC_B8, b8_t, ORC_S32, int)
#define FOR_SBK_VALS(X)                                                                            \
    X(ORC_F32, float) X(ORC_F64, double) X(ORC_C32, cfloat) X(ORC_C64, cdouble) X(ORC_S32, int)    \
    X(ORC_U32, unsigned) X(ORC_S64, long long) X(ORC_U64, unsigned long long)

}  // namespace

extern "C" {

int orc_is_reference(void) { return 0; }

int orc_reduce(int op, int ti, int to, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4],
               int dim, int change_nan, double nanval, void *out) {
#define CALL(OP, Ti, To) { reduce_lines<OP, Ti, To>(in, dims, strides, dim, change_nan != 0, nanval, out); return 0; }
    if (op == ORC_OP_MIN || op == ORC_OP_MAX) {
#define X(code, T) if (ti == code && to == code) { if (op == ORC_OP_MIN) CALL(ORC_OP_MIN, T, T) else CALL(ORC_OP_MAX, T, T) }
        FOR_ALL_TYPES(X)
#undef X
    } else if (op == ORC_OP_ADD || op == ORC_OP_MUL) {
#define X(ci, Ti, co, To) if (ti == ci && to == co) { if (op == ORC_OP_ADD) CALL(ORC_OP_ADD, Ti, To) else CALL(ORC_OP_MUL, Ti, To) }
        FOR_SUM_PAIRS(X)
#undef X
    } else if (op == ORC_OP_NOTZERO) {
#define X(code, T) if (ti == code && to == ORC_U32) CALL(ORC_OP_NOTZERO, T, unsigned)
        FOR_ALL_TYPES(X)
#undef X
    } else if (op == ORC_OP_AND || op == ORC_OP_OR) {
#define X(code, T) if (ti == code && to == ORC_B8) { if (op == ORC_OP_AND) CALL(ORC_OP_AND, T, b8_t) else CALL(ORC_OP_OR, T, b8_t) }
        FOR_ALL_TYPES(X)
#undef X
    }
#undef CALL
    return -1;
}

int orc_reduce_all(int op, int ti, int to, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4],
                   int change_nan, double nanval, void *out) {
#define CALL(OP, Ti, To) { reduce_everything<OP, Ti, To>(in, dims, strides, change_nan != 0, nanval, out); return 0; }
    if (op == ORC_OP_MIN || op == ORC_OP_MAX) {
#define X(code, T) if (ti == code && to == code) { if (op == ORC_OP_MIN) CALL(ORC_OP_MIN, T, T) else CALL(ORC_OP_MAX, T, T) }
        FOR_ALL_TYPES(X)
#undef X
    } else if (op == ORC_OP_ADD || op == ORC_OP_MUL) {
#define X(ci, Ti, co, To) if (ti == ci && to == co) { if (op == ORC_OP_ADD) CALL(ORC_OP_ADD, Ti, To) else CALL(ORC_OP_MUL, Ti, To) }
        FOR_SUM_PAIRS(X)
#undef X
    } else if (op == ORC_OP_NOTZERO) {
#define X(code, T) if (ti == code && to == ORC_U32) CALL(ORC_OP_NOTZERO, T, unsigned)
        FOR_ALL_TYPES(X)
#undef X
    } else if (op == ORC_OP_AND || op == ORC_OP_OR) {
#define X(code, T) if (ti == code && to == ORC_B8) { if (op == ORC_OP_AND) CALL(ORC_OP_AND, T, b8_t) else CALL(ORC_OP_OR, T, b8_t) }
        FOR_ALL_TYPES(X)
#undef X
    }
#undef CALL
    return -1;
}

int orc_ireduce(int op, int t, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], int dim,
                void *out_val, uint32_t *out_loc) {
    if (op != ORC_OP_MIN && op != ORC_OP_MAX) return -1;
#define X(code, T)                                                                     \
    if (t == code) {                                                                   \
        if (op == ORC_OP_MIN) ireduce_lines<ORC_OP_MIN, T>(in, dims, strides, dim, out_val, out_loc); \
        else ireduce_lines<ORC_OP_MAX, T>(in, dims, strides, dim, out_val, out_loc);   \
        return 0;                                                                      \
    }
    FOR_ALL_TYPES(X)
#undef X
    return -1;
}

int orc_ireduce_all(int op, int t, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4],
                    void *out_val, uint32_t *out_loc) {
    if (op != ORC_OP_MIN && op != ORC_OP_MAX) return -1;
    if (empty(dims)) return -1;
#define X(code, T)                                                                     \
    if (t == code) {                                                                   \
        if (op == ORC_OP_MIN) ireduce_everything<ORC_OP_MIN, T>(in, dims, strides, out_val, out_loc); \
        else ireduce_everything<ORC_OP_MAX, T>(in, dims, strides, out_val, out_loc);   \
        return 0;                                                                      \
    }
    FOR_ALL_TYPES(X)
#undef X
    return -1;
}

int orc_scan(int op, int ti, int to, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], int dim,
             int inclusive, void *out) {
#define X(ci, Ti, co, To)                                                                         \
    if (ti == ci && to == co) {                                                                   \
        switch (op) {                                                                             \
            case ORC_OP_ADD: scan_lines<ORC_OP_ADD, Ti, To>(in, dims, strides, dim, inclusive != 0, out); return 0; \
            case ORC_OP_MUL: scan_lines<ORC_OP_MUL, Ti, To>(in, dims, strides, dim, inclusive != 0, out); return 0; \
            case ORC_OP_MIN: scan_lines<ORC_OP_MIN, Ti, To>(in, dims, strides, dim, inclusive != 0, out); return 0; \
            case ORC_OP_MAX: scan_lines<ORC_OP_MAX, Ti, To>(in, dims, strides, dim, inclusive != 0, out); return 0; \
            default: break;                                                                       \
        }                                                                                         \
    }
    FOR_SCAN_PAIRS(X)
#undef X
    if (op == ORC_OP_NOTZERO && ti == ORC_B8 && to == ORC_U32) {  // af_accum on b8, api/c/scan.cpp:148
        scan_lines<ORC_OP_NOTZERO, b8_t, unsigned>(in, dims, strides, dim, inclusive != 0, out);
        return 0;
    }
    return -1;
}

}  // extern "C"

template<int OP, typename T>
static int sbk_key(int tk, const void *key, const orc_dim_t *ks, const void *in, const orc_dim_t *dims,
                   const orc_dim_t *strides, int dim, bool inclusive, void *out) {
    switch (tk) {
        case ORC_S32: scan_by_key_lines<OP, T, int>(key, ks, in, dims, strides, dim, inclusive, out); return 0;
        case ORC_U32: scan_by_key_lines<OP, T, unsigned>(key, ks, in, dims, strides, dim, inclusive, out); return 0;
        case ORC_S64: scan_by_key_lines<OP, T, long long>(key, ks, in, dims, strides, dim, inclusive, out); return 0;
        case ORC_U64: scan_by_key_lines<OP, T, unsigned long long>(key, ks, in, dims, strides, dim, inclusive, out); return 0;
        default: return -1;
    }
}

extern "C" {

int orc_scan_by_key(int op, int t, int tk, const void *key, const orc_dim_t kstrides[4], const void *in,
                    const orc_dim_t dims[4], const orc_dim_t strides[4], int dim, int inclusive, void *out) {
#define X(code, T)                                                                                   \
    if (t == code) {                                                                                 \
        switch (op) {                                                                                \
            case ORC_OP_ADD: return sbk_key<ORC_OP_ADD, T>(tk, key, kstrides, in, dims, strides, dim, inclusive != 0, out); \
            case ORC_OP_MUL: return sbk_key<ORC_OP_MUL, T>(tk, key, kstrides, in, dims, strides, dim, inclusive != 0, out); \
            case ORC_OP_MIN: return sbk_key<ORC_OP_MIN, T>(tk, key, kstrides, in, dims, strides, dim, inclusive != 0, out); \
            case ORC_OP_MAX: return sbk_key<ORC_OP_MAX, T>(tk, key, kstrides, in, dims, strides, dim, inclusive != 0, out); \
            default: return -1;                                                                      \
        }                                                                                            \
    }
    FOR_SBK_VALS(X)
#undef X
    return -1;
}

long long orc_where(int t, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], uint32_t *out) {
#define X(code, T) if (t == code) return where_hits<T>(in, dims, strides, out);
    FOR_NO_HALF_TYPES(X)
#undef X
    return -1;
}

}  // extern "C"

template<int OP, typename Ti, typename To>
static long long rbk_key(int tk, const void *keys, const void *vals, const orc_dim_t *dims, const orc_dim_t *strides, int dim,
                         bool cn, double nv, void *keys_out, void *vals_out) {
    const orc_dim_t n = dims[dim];
    long long runs;
    if (tk == ORC_S32) runs = count_runs<int>(keys, n, keys_out);
    else if (tk == ORC_U32) runs = count_runs<unsigned>(keys, n, keys_out);
    else return -1;
    if (vals_out) {
        if (tk == ORC_S32) reduce_by_key_lines<OP, Ti, To, int>(keys, vals, dims, strides, dim, cn, nv, runs, vals_out);
        else reduce_by_key_lines<OP, Ti, To, unsigned>(keys, vals, dims, strides, dim, cn, nv, runs, vals_out);
    }
    return runs;
}

extern "C" {

long long orc_reduce_by_key(int op, int ti, int to, int tk, const void *keys, const void *vals, const orc_dim_t dims[4],
                            const orc_dim_t strides[4], int dim, int change_nan, double nanval, void *keys_out,
                            void *vals_out) {
#define CALL(OP, Ti, To) return rbk_key<OP, Ti, To>(tk, keys, vals, dims, strides, dim, change_nan != 0, nanval, keys_out, vals_out);
    if (op == ORC_OP_MIN || op == ORC_OP_MAX) {
#define X(code, T) if (ti == code && to == code) { if (op == ORC_OP_MIN) CALL(ORC_OP_MIN, T, T) else CALL(ORC_OP_MAX, T, T) }
        FOR_ALL_TYPES(X)
#undef X
    } else if (op == ORC_OP_ADD || op == ORC_OP_MUL) {
#define X(ci, Ti, co, To) if (ti == ci && to == co) { if (op == ORC_OP_ADD) CALL(ORC_OP_ADD, Ti, To) else CALL(ORC_OP_MUL, Ti, To) }
        FOR_SUM_PAIRS(X)
#undef X
    } else if (op == ORC_OP_NOTZERO) {
#define X(code, T) if (ti == code && to == ORC_U32) CALL(ORC_OP_NOTZERO, T, unsigned)
        FOR_ALL_TYPES(X)
#undef X
    } else if (op == ORC_OP_AND || op == ORC_OP_OR) {
#define X(code, T) if (ti == code && to == ORC_B8) { if (op == ORC_OP_AND) CALL(ORC_OP_AND, T, b8_t) else CALL(ORC_OP_OR, T, b8_t) }
        FOR_ALL_TYPES(X)
#undef X
    }
#undef CALL
    return -1;
}

int orc_ireduce_ragged(int op, int t, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], int dim,
                       const uint32_t *rlen, void *out_val, uint32_t *out_loc) {
    if (op != ORC_OP_MIN && op != ORC_OP_MAX) return -1;
#define X(code, T)                                                                                          \
    if (t == code) {                                                                                        \
        if (op == ORC_OP_MIN) ireduce_ragged_lines<ORC_OP_MIN, T>(in, dims, strides, dim, rlen, out_val, out_loc); \
        else ireduce_ragged_lines<ORC_OP_MAX, T>(in, dims, strides, dim, rlen, out_val, out_loc);           \
        return 0;                                                                                           \
    }
    FOR_ALL_TYPES(X)
#undef X
    return -1;
}

}  // extern "C"

// ==== statistics =======================================================================
namespace {

// mean: src/backend/cpu/kernel/mean.hpp:18-127 -- MeanOp keeps a running mean and a running count; one serial loop per line
template<typename Tc, typename Tw> struct RunningMean {
    Tc mean;
    Tw count;
    RunningMean() : mean(Tc(0)), count(Tw(0)) {}
    void push(Tc x, Tw w) {                       // mean.hpp:28-38
        if (w != Tw(0) || count != Tw(0)) {
            Tw rs = count, ns = w;
            count += w;
            rs = rs / count;
            ns = ns / count;
            mean = (rs * mean) + (ns * x);
        }
    }
};
template<typename T> struct base_of { using type = T; };
template<> struct base_of<cfloat> { using type = float; };
template<> struct base_of<cdouble> { using type = double; };

// float -> half as the CPU backend stores it: common/half.hpp:258-420 (table-driven conversion, round to nearest with TIES
// AWAY from zero -- HALF_ROUND_TIES_TO_EVEN is off), not the IEEE ties-to-even of a plain cast
inline half_t half_ties_away(float v) {
    half_t h = (half_t)v;
    const float hf = (float)h;
    if (!(v == v) || hf == v || std::isinf(hf)) return h;
    uint16_t bits;
    std::memcpy(&bits, &h, 2);
    const bool toward_larger = std::fabs(v) > std::fabs(hf);
    uint16_t nb = (uint16_t)(toward_larger ? bits + 1 : bits - 1);
    half_t n;
    std::memcpy(&n, &nb, 2);
    const float nf = (float)n;
    if (std::fabs(v - hf) == std::fabs(nf - v) && std::fabs(nf) > std::fabs(hf)) return n;   // exact tie: the larger magnitude
    return h;
}
template<typename To, typename Tc> inline To stat_store(Tc v) { return (To)v; }
template<> inline half_t stat_store<half_t, float>(float v) { return half_ties_away(v); }

// Ti: stored input type, Tc: compute type (float / double / complex), To: stored output type
template<typename Ti, typename Tc, typename To>
void mean_lines(const void *vin, const dim_t *dims, const dim_t *strides, int dim, void *vout) {
    using Tw = typename base_of<Tc>::type;
    const Ti *in = (const Ti *)vin;
    To *out = (To *)vout;
    if (empty(dims)) return;
    Shape sh = mk(dims, strides);
    dim_t od[4] = {dims[0], dims[1], dims[2], dims[3]}, os[4];
    od[dim] = 1;
    dense_strides(od, os);
    const dim_t len = dims[dim], step = strides[dim];
    for_each_line(sh, os, dim, [&](dim_t ib, dim_t ob) {
        RunningMean<Tc, Tw> op;
        for (dim_t i = 0; i < len; i++) op.push(Tc(in[ib + i * step]), Tw(1));
        out[ob] = stat_store<To, Tc>(op.mean);
    });
}
template<typename Ti, typename Tc>
Tc mean_everything(const void *vin, const dim_t *dims, const dim_t *strides) {     // cpu/mean.cpp:113-141
    using Tw = typename base_of<Tc>::type;
    const Ti *in = (const Ti *)vin;
    RunningMean<Tc, Tw> op;
    for (dim_t l = 0; l < dims[3]; l++)
        for (dim_t k = 0; k < dims[2]; k++)
            for (dim_t j = 0; j < dims[1]; j++) {
                const Ti *p = in + j * strides[1] + k * strides[2] + l * strides[3];
                for (dim_t i = 0; i < dims[0]; i++) op.push(Tc(p[i]), Tw(1));
            }
    return op.mean;
}
// var / stdev along a dimension: api/c/var.cpp:104-136 (meanvar) and stdev.cpp:61-96: cast, mean along dim, (x - mean)^2
// summed serially (reduce<af_add_t>), times 1 / denom [, sqrt].  `denom` is computed by the caller (the two entry points
// treat the DEFAULT bias differently).
template<typename Ti, typename Tc, typename To>
void var_lines(const void *vin, const dim_t *dims, const dim_t *strides, int dim, double denom, int take_sqrt, void *vout,
               void *vmean) {
    using Tw = typename base_of<Tc>::type;
    const Ti *in = (const Ti *)vin;
    To *out = (To *)vout, *mout = (To *)vmean;
    if (empty(dims)) return;
    Shape sh = mk(dims, strides);
    dim_t od[4] = {dims[0], dims[1], dims[2], dims[3]}, os[4];
    od[dim] = 1;
    dense_strides(od, os);
    const dim_t len = dims[dim], step = strides[dim];
    for_each_line(sh, os, dim, [&](dim_t ib, dim_t ob) {
        RunningMean<Tc, Tw> op;
        for (dim_t i = 0; i < len; i++) op.push(Tc((To)in[ib + i * step]), Tw(1));   // cast<outType>(in) first
        const Tc m = (Tc)stat_store<To, Tc>(op.mean);
        Tc acc = Tc(0);
        for (dim_t i = 0; i < len; i++) {
            const Tc d = Tc((To)in[ib + i * step]) - m;
            acc = d * d + acc;
        }
        Tc v = acc * Tc((Tw)(1.0 / denom));
        if (take_sqrt) v = std::sqrt(v);
        out[ob] = stat_store<To, Tc>(v);
        if (mout) mout[ob] = stat_store<To, Tc>(op.mean);
    });
}
template<typename Ti, typename Tc>
Tc var_everything(const void *vin, const dim_t *dims, const dim_t *strides, double denom, int take_sqrt) {   // var.cpp:52-75
    const Ti *in = (const Ti *)vin;
    const Tc m = mean_everything<Ti, Tc>(vin, dims, strides);
    Tc acc = Tc(0);
    for (dim_t l = 0; l < dims[3]; l++)
        for (dim_t k = 0; k < dims[2]; k++)
            for (dim_t j = 0; j < dims[1]; j++) {
                const Ti *p = in + j * strides[1] + k * strides[2] + l * strides[3];
                for (dim_t i = 0; i < dims[0]; i++) {
                    const Tc d = Tc(p[i]) - m;
                    acc = d * d + acc;
                }
            }
    using Tw = typename base_of<Tc>::type;
    Tc v = acc / Tc((Tw)denom);
    if (take_sqrt) v = std::sqrt(v);
    return v;
}

#define FOR_STAT_TYPES(X)                                                                                           \
    X(ORC_F32, float, float, float) X(ORC_F64, double, double, double) X(ORC_C32, cfloat, cfloat, cfloat)           \
    X(ORC_C64, cdouble, cdouble, cdouble) X(ORC_S32, int, float, float) X(ORC_U32, unsigned, float, float)          \
    X(ORC_S64, long long, double, double) X(ORC_U64, unsigned long long, double, double)                            \
    X(ORC_S16, short, float, float) X(ORC_U16, unsigned short, float, float) X(ORC_S8, signed char, float, float)   \
    X(ORC_U8, unsigned char, float, float) X(ORC_B8, b8_t, float, float) X(ORC_F16, half_t, float, half_t)

}  // namespace

extern "C" {

/* af_mean along dim (api/c/mean.cpp:65-93): out dtype f32 / f64 / c32 / c64 / f16 per the table above */
int ORC_FN(mean)(int ti, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], int dim, void *out) {
#define X(code, Ti, Tc, To) if (ti == code) { mean_lines<Ti, Tc, To>(in, dims, strides, dim, out); return 0; }
    FOR_STAT_TYPES(X)
#undef X
    return -1;
}
/* af_mean_all (mean.cpp:151-187): out = 2 doubles (real, imag) */
int ORC_FN(mean_all)(int ti, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], double *out) {
    out[0] = out[1] = 0;
#define X(code, Ti, Tc, To) if (ti == code) { const std::complex<double> v(mean_everything<Ti, Tc>(in, dims, strides)); out[0] = v.real(); out[1] = v.imag(); return 0; }
    FOR_STAT_TYPES(X)
#undef X
    return -1;
}
/* af_var_v2 / af_stdev_v2 / af_meanvar along dim; mean_out may be NULL */
int ORC_FN(var)(int ti, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], int dim, double denom,
                int take_sqrt, void *out, void *mean_out) {
#define X(code, Ti, Tc, To) if (ti == code) { var_lines<Ti, Tc, To>(in, dims, strides, dim, denom, take_sqrt, out, mean_out); return 0; }
    FOR_STAT_TYPES(X)
#undef X
    return -1;
}
/* af_var_all_v2 / af_stdev_all_v2: out = 2 doubles */
int ORC_FN(var_all)(int ti, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], double denom, int take_sqrt,
                    double *out) {
    out[0] = out[1] = 0;
#define X(code, Ti, Tc, To) if (ti == code) { const std::complex<double> v(var_everything<Ti, Tc>(in, dims, strides, denom, take_sqrt)); out[0] = v.real(); out[1] = v.imag(); return 0; }
    FOR_STAT_TYPES(X)
#undef X
    return -1;
}

}  // extern "C"
