// ref_driver.cpp -- thin extern "C" dispatcher over the REFERENCE'S OWN CPU
// kernel templates, compiled in place from /root/reference by oracle/Makefile
// into oracle/_ref/libafref.so (git-ignored, travels to the GPU box).
//
// TEST INFRASTRUCTURE ONLY (see oracle_abi.h).  No reference source is copied
// into this repository: the headers are #included from where they lie
// (-I/root/reference/src/backend/cpu etc.), and this file only instantiates
//   kernel::reduce_dim / reduce_all      src/backend/cpu/kernel/reduce.hpp:20-63,162-200
//   kernel::ireduce_dim / MinMaxOp       src/backend/cpu/kernel/ireduce.hpp:29-104
//   kernel::scan_dim                     src/backend/cpu/kernel/scan.hpp:19-72
//   kernel::scan_dim_by_key              src/backend/cpu/kernel/scan_by_key.hpp:19-86
//   kernel::mean_dim / MeanOp            src/backend/cpu/kernel/mean.hpp:18-127
// for the (op, Ti, To) triples the reference's cpu/*.cpp files instantiate.
// The host-side wrappers that pick the recursion depth (cpu/reduce.cpp:48-66:
// reduce_funcs[in.ndims()-1], cpu/scan.cpp:24-70) and the whole-array ireduce
// loop (cpu/ireduce.cpp:59-90, which needs Array<T>) are re-expressed here.
// `where` (cpu/where.cpp, needs Array/memAlloc) cannot be compiled standalone
// and is a restatement in this library as well.

#include <Param.hpp>
#include <kernel/ireduce.hpp>
#include <kernel/mean.hpp>
#include <kernel/reduce.hpp>
#include <kernel/scan.hpp>
#include <kernel/scan_by_key.hpp>
#include <math.hpp>

#include <cstdint>
#include <cstring>
#include <vector>

#define ORC_PREFIX ref
#include "oracle_abi.h"

using af::dim4;
using namespace arrayfire::cpu;
using arrayfire::common::half;

namespace {

dim4 D(const orc_dim_t *d) { return dim4(d[0], d[1], d[2], d[3]); }
dim4 dense(const dim4 &d) {
    dim4 s(1, 1, 1, 1);
    for (int i = 1; i < 4; i++) s[i] = s[i - 1] * d[i - 1];
    return s;
}
int ndims_of(const dim4 &d) { return (int)d.ndims(); }

template<af_op_t op, typename Ti, typename To>
int do_reduce(const void *in, const orc_dim_t *dims, const orc_dim_t *strides, int dim, bool cn, double nv,
              void *out) {
    dim4 idims = D(dims), odims = D(dims);
    if (idims.elements() == 0) return 0;
    odims[dim] = 1;
    Param<To> o((To *)out, odims, dense(odims));
    CParam<Ti> i((const Ti *)in, idims, D(strides));
    switch (ndims_of(idims)) {
        case 1: kernel::reduce_dim<op, Ti, To, 1>()(o, 0, i, 0, dim, cn, nv); break;
        case 2: kernel::reduce_dim<op, Ti, To, 2>()(o, 0, i, 0, dim, cn, nv); break;
        case 3: kernel::reduce_dim<op, Ti, To, 3>()(o, 0, i, 0, dim, cn, nv); break;
        default: kernel::reduce_dim<op, Ti, To, 4>()(o, 0, i, 0, dim, cn, nv); break;
    }
    return 0;
}

template<af_op_t op, typename Ti, typename To>
int do_reduce_all(const void *in, const orc_dim_t *dims, const orc_dim_t *strides, bool cn, double nv, void *out) {
    dim4 one(1, 1, 1, 1);
    Param<To> o((To *)out, one, one);
    CParam<Ti> i((const Ti *)in, D(dims), D(strides));
    kernel::reduce_all<op, Ti, To>()(o, i, cn, nv);
    return 0;
}

template<af_op_t op, typename T>
int do_ireduce(const void *in, const orc_dim_t *dims, const orc_dim_t *strides, int dim, void *out, uint32_t *loc) {
    dim4 idims = D(dims), odims = D(dims);
    if (idims.elements() == 0) return 0;
    odims[dim] = 1;
    Param<T> o((T *)out, odims, dense(odims));
    Param<uint> l((uint *)loc, odims, dense(odims));
    CParam<T> i((const T *)in, idims, D(strides));
    CParam<uint> rlen(nullptr, dim4(0), dim4(0));
    switch (ndims_of(idims)) {
        case 1: kernel::ireduce_dim<op, T, 1>()(o, l, 0, i, 0, dim, rlen); break;
        case 2: kernel::ireduce_dim<op, T, 2>()(o, l, 0, i, 0, dim, rlen); break;
        case 3: kernel::ireduce_dim<op, T, 3>()(o, l, 0, i, 0, dim, rlen); break;
        default: kernel::ireduce_dim<op, T, 4>()(o, l, 0, i, 0, dim, rlen); break;
    }
    return 0;
}

template<af_op_t op, typename T>
int do_ireduce_all(const void *vin, const orc_dim_t *dims, const orc_dim_t *strides, void *out, uint32_t *loc) {
    const T *in = (const T *)vin;
    kernel::MinMaxOp<op, T> Op(in[0], 0);
    orc_dim_t pos = 0;
    for (orc_dim_t l = 0; l < dims[3]; l++)
        for (orc_dim_t k = 0; k < dims[2]; k++)
            for (orc_dim_t j = 0; j < dims[1]; j++)
                for (orc_dim_t i = 0; i < dims[0]; i++)
                    Op(in[i + j * strides[1] + k * strides[2] + l * strides[3]], pos++);
    *(T *)out = Op.m_val;
    *loc      = Op.m_idx;
    return 0;
}

template<af_op_t op, typename Ti, typename To, bool incl>
void do_scan_n(Param<To> o, CParam<Ti> i, int nd, int dim) {
    switch (nd) {
        case 1: kernel::scan_dim<op, Ti, To, 1, incl>()(o, 0, i, 0, dim); break;
        case 2: kernel::scan_dim<op, Ti, To, 2, incl>()(o, 0, i, 0, dim); break;
        case 3: kernel::scan_dim<op, Ti, To, 3, incl>()(o, 0, i, 0, dim); break;
        default: kernel::scan_dim<op, Ti, To, 4, incl>()(o, 0, i, 0, dim); break;
    }
}

template<af_op_t op, typename Ti, typename To>
int do_scan(const void *in, const orc_dim_t *dims, const orc_dim_t *strides, int dim, bool incl, void *out) {
    dim4 idims = D(dims);
    if (idims.elements() == 0) return 0;
    Param<To> o((To *)out, idims, dense(idims));
    CParam<Ti> i((const Ti *)in, idims, D(strides));
    if (incl) do_scan_n<op, Ti, To, true>(o, i, ndims_of(idims), dim);
    else do_scan_n<op, Ti, To, false>(o, i, ndims_of(idims), dim);
    return 0;
}

template<af_op_t op, typename T, typename Tk>
int do_sbk(const void *key, const orc_dim_t *ks, const void *in, const orc_dim_t *dims, const orc_dim_t *strides,
           int dim, bool incl, void *out) {
    dim4 idims = D(dims);
    if (idims.elements() == 0) return 0;
    Param<T> o((T *)out, idims, dense(idims));
    CParam<T> i((const T *)in, idims, D(strides));
    CParam<Tk> k((const Tk *)key, idims, D(ks));
    switch (ndims_of(idims)) {
        case 1: { kernel::scan_dim_by_key<op, T, Tk, T, 1> f(incl); f(o, 0, k, 0, i, 0, dim); } break;
        case 2: { kernel::scan_dim_by_key<op, T, Tk, T, 2> f(incl); f(o, 0, k, 0, i, 0, dim); } break;
        case 3: { kernel::scan_dim_by_key<op, T, Tk, T, 3> f(incl); f(o, 0, k, 0, i, 0, dim); } break;
        default: { kernel::scan_dim_by_key<op, T, Tk, T, 4> f(incl); f(o, 0, k, 0, i, 0, dim); } break;
    }
    return 0;
}

template<af_op_t op, typename T>
int do_sbk_key(int tk, const void *key, const orc_dim_t *ks, const void *in, const orc_dim_t *dims,
               const orc_dim_t *strides, int dim, bool incl, void *out) {
    switch (tk) {
        case ORC_S32: return do_sbk<op, T, int>(key, ks, in, dims, strides, dim, incl, out);
        case ORC_U32: return do_sbk<op, T, uint>(key, ks, in, dims, strides, dim, incl, out);
        case ORC_S64: return do_sbk<op, T, intl>(key, ks, in, dims, strides, dim, incl, out);
        case ORC_U64: return do_sbk<op, T, uintl>(key, ks, in, dims, strides, dim, incl, out);
        default: return -1;
    }
}

template<typename T>
long long do_where(const void *vin, const orc_dim_t *dims, const orc_dim_t *strides, uint32_t *out) {
    const T *in  = (const T *)vin;
    const T zero = scalar<T>(0);
    long long n  = 0;
    uint32_t pos = 0;
    for (orc_dim_t l = 0; l < dims[3]; l++)
        for (orc_dim_t k = 0; k < dims[2]; k++)
            for (orc_dim_t j = 0; j < dims[1]; j++)
                for (orc_dim_t i = 0; i < dims[0]; i++, pos++)
                    if (in[i + j * strides[1] + k * strides[2] + l * strides[3]] != zero) out[n++] = pos;
    return n;
}

#define ALL_TYPES(X)                                                                               \
    X(ORC_F32, float) X(ORC_F64, double) X(ORC_C32, cfloat) X(ORC_C64, cdouble) X(ORC_S32, int)    \
    X(ORC_U32, uint) X(ORC_S64, intl) X(ORC_U64, uintl) X(ORC_S16, short) X(ORC_U16, ushort)       \
    X(ORC_S8, schar) X(ORC_U8, uchar) X(ORC_B8, char) X(ORC_F16, half)
#define NO_HALF_TYPES(X)                                                                           \
    X(ORC_F32, float) X(ORC_F64, double) X(ORC_C32, cfloat) X(ORC_C64, cdouble) X(ORC_S32, int)    \
    X(ORC_U32, uint) X(ORC_S64, intl) X(ORC_U64, uintl) X(ORC_S16, short) X(ORC_U16, ushort)       \
    X(ORC_S8, schar) X(ORC_U8, uchar) X(ORC_B8, char)
#define SUM_PAIRS(X)                                                                               \
    X(ORC_F32, float, ORC_F32, float) X(ORC_F64, double, ORC_F64, double)                          \
    X(ORC_C32, cfloat, ORC_C32, cfloat) X(ORC_C64, cdouble, ORC_C64, cdouble)                      \
    X(ORC_S32, int, ORC_S32, int) X(ORC_U32, uint, ORC_U32, uint) X(ORC_S64, intl, ORC_S64, intl)  \
    X(ORC_U64, uintl, ORC_U64, uintl) X(ORC_S16, short, ORC_S32, int) X(ORC_U16, ushort, ORC_U32, uint) \
    X(ORC_S8, schar, ORC_S32, int) X(ORC_U8, uchar, ORC_U32, uint) X(ORC_F16, half, ORC_F32, float)
#define SCAN_PAIRS(X)                                                                              \
    X(ORC_F32, float, ORC_F32, float) X(ORC_F64, double, ORC_F64, double)                          \
    X(ORC_C32, cfloat, ORC_C32, cfloat) X(ORC_C64, cdouble, ORC_C64, cdouble)                      \
    X(ORC_S32, int, ORC_S32, int) X(ORC_U32, uint, ORC_U32, uint) X(ORC_S64, intl, ORC_S64, intl)  \
    X(ORC_U64, uintl, ORC_U64, uintl) X(ORC_S16, short, ORC_S32, int) X(ORC_U16, ushort, ORC_U32, uint) \
    X(ORC_S8, schar, ORC_S32, int) X(ORC_U8, uchar, ORC_U32, uint) X(ORC_B8, char, ORC_U32, uint)  \
    X(ORC_B8, char, ORC_S32, int)
#define SBK_VALS(X)                                                                                \
    X(ORC_F32, float) X(ORC_F64, double) X(ORC_C32, cfloat) X(ORC_C64, cdouble) X(ORC_S32, int)    \
    X(ORC_U32, uint) X(ORC_S64, intl) X(ORC_U64, uintl)

}  // namespace

extern "C" {

int ref_is_reference(void) { return 1; }

int ref_reduce(int op, int ti, int to, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4],
               int dim, int cn, double nv, void *out) {
    if (op == ORC_OP_MIN || op == ORC_OP_MAX) {
#define X(c, T) if (ti == c && to == c) return op == ORC_OP_MIN ? do_reduce<af_min_t, T, T>(in, dims, strides, dim, cn, nv, out) : do_reduce<af_max_t, T, T>(in, dims, strides, dim, cn, nv, out);
        ALL_TYPES(X)
#undef X
    } else if (op == ORC_OP_ADD || op == ORC_OP_MUL) {
#define X(ci, Ti, co, To) if (ti == ci && to == co) return op == ORC_OP_ADD ? do_reduce<af_add_t, Ti, To>(in, dims, strides, dim, cn, nv, out) : do_reduce<af_mul_t, Ti, To>(in, dims, strides, dim, cn, nv, out);
        SUM_PAIRS(X)
#undef X
    } else if (op == ORC_OP_NOTZERO) {
#define X(c, T) if (ti == c && to == ORC_U32) return do_reduce<af_notzero_t, T, uint>(in, dims, strides, dim, cn, nv, out);
        ALL_TYPES(X)
#undef X
    } else if (op == ORC_OP_AND || op == ORC_OP_OR) {
#define X(c, T) if (ti == c && to == ORC_B8) return op == ORC_OP_AND ? do_reduce<af_and_t, T, char>(in, dims, strides, dim, cn, nv, out) : do_reduce<af_or_t, T, char>(in, dims, strides, dim, cn, nv, out);
        ALL_TYPES(X)
#undef X
    }
    return -1;
}

int ref_reduce_all(int op, int ti, int to, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4],
                   int cn, double nv, void *out) {
    if (op == ORC_OP_MIN || op == ORC_OP_MAX) {
#define X(c, T) if (ti == c && to == c) return op == ORC_OP_MIN ? do_reduce_all<af_min_t, T, T>(in, dims, strides, cn, nv, out) : do_reduce_all<af_max_t, T, T>(in, dims, strides, cn, nv, out);
        ALL_TYPES(X)
#undef X
    } else if (op == ORC_OP_ADD || op == ORC_OP_MUL) {
#define X(ci, Ti, co, To) if (ti == ci && to == co) return op == ORC_OP_ADD ? do_reduce_all<af_add_t, Ti, To>(in, dims, strides, cn, nv, out) : do_reduce_all<af_mul_t, Ti, To>(in, dims, strides, cn, nv, out);
        SUM_PAIRS(X)
#undef X
    } else if (op == ORC_OP_NOTZERO) {
#define X(c, T) if (ti == c && to == ORC_U32) return do_reduce_all<af_notzero_t, T, uint>(in, dims, strides, cn, nv, out);
        ALL_TYPES(X)
#undef X
    } else if (op == ORC_OP_AND || op == ORC_OP_OR) {
#define X(c, T) if (ti == c && to == ORC_B8) return op == ORC_OP_AND ? do_reduce_all<af_and_t, T, char>(in, dims, strides, cn, nv, out) : do_reduce_all<af_or_t, T, char>(in, dims, strides, cn, nv, out);
        ALL_TYPES(X)
#undef X
    }
    return -1;
}

int ref_ireduce(int op, int t, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], int dim,
                void *out_val, uint32_t *out_loc) {
    if (op != ORC_OP_MIN && op != ORC_OP_MAX) return -1;
#define X(c, T) if (t == c) return op == ORC_OP_MIN ? do_ireduce<af_min_t, T>(in, dims, strides, dim, out_val, out_loc) : do_ireduce<af_max_t, T>(in, dims, strides, dim, out_val, out_loc);
    ALL_TYPES(X)
#undef X
    return -1;
}

int ref_ireduce_all(int op, int t, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4],
                    void *out_val, uint32_t *out_loc) {
    if (op != ORC_OP_MIN && op != ORC_OP_MAX) return -1;
    if (dims[0] * dims[1] * dims[2] * dims[3] == 0) return -1;
#define X(c, T) if (t == c) return op == ORC_OP_MIN ? do_ireduce_all<af_min_t, T>(in, dims, strides, out_val, out_loc) : do_ireduce_all<af_max_t, T>(in, dims, strides, out_val, out_loc);
    ALL_TYPES(X)
#undef X
    return -1;
}

int ref_scan(int op, int ti, int to, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], int dim,
             int incl, void *out) {
#define X(ci, Ti, co, To)                                                                          \
    if (ti == ci && to == co) {                                                                    \
        switch (op) {                                                                              \
            case ORC_OP_ADD: return do_scan<af_add_t, Ti, To>(in, dims, strides, dim, incl, out);  \
            case ORC_OP_MUL: return do_scan<af_mul_t, Ti, To>(in, dims, strides, dim, incl, out);  \
            case ORC_OP_MIN: return do_scan<af_min_t, Ti, To>(in, dims, strides, dim, incl, out);  \
            case ORC_OP_MAX: return do_scan<af_max_t, Ti, To>(in, dims, strides, dim, incl, out);  \
            default: break;                                                                        \
        }                                                                                          \
    }
    SCAN_PAIRS(X)
#undef X
    if (op == ORC_OP_NOTZERO && ti == ORC_B8 && to == ORC_U32)
        return do_scan<af_notzero_t, char, uint>(in, dims, strides, dim, incl, out);
    return -1;
}

int ref_scan_by_key(int op, int t, int tk, const void *key, const orc_dim_t ks[4], const void *in,
                    const orc_dim_t dims[4], const orc_dim_t strides[4], int dim, int incl, void *out) {
#define X(c, T)                                                                                     \
    if (t == c) {                                                                                   \
        switch (op) {                                                                               \
            case ORC_OP_ADD: return do_sbk_key<af_add_t, T>(tk, key, ks, in, dims, strides, dim, incl, out); \
            case ORC_OP_MUL: return do_sbk_key<af_mul_t, T>(tk, key, ks, in, dims, strides, dim, incl, out); \
            case ORC_OP_MIN: return do_sbk_key<af_min_t, T>(tk, key, ks, in, dims, strides, dim, incl, out); \
            case ORC_OP_MAX: return do_sbk_key<af_max_t, T>(tk, key, ks, in, dims, strides, dim, incl, out); \
            default: return -1;                                                                     \
        }                                                                                           \
    }
    SBK_VALS(X)
#undef X
    return -1;
}

long long ref_where(int t, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], uint32_t *out) {
#define X(c, T) if (t == c) return do_where<T>(in, dims, strides, out);
    NO_HALF_TYPES(X)
#undef X
    return -1;
}

// ---- reduce_by_key: the reference's own two steps (cpu/reduce.cpp:73-108): kernel::n_reduced_keys, then
// kernel::reduce_dim_by_key<op,Ti,Tk,To,ndims>
}  // extern "C"

namespace {
template<af_op_t op, typename Ti, typename To, typename Tk>
long long do_rbk(const void *keys, const void *vals, const orc_dim_t *dims, const orc_dim_t *strides, int dim, bool cn,
                 double nv, void *keys_out, void *vals_out) {
    const orc_dim_t n = dims[dim];
    dim4 kd(n, 1, 1, 1);
    std::vector<Tk> full((size_t)n);
    int n_reduced = 0;
    kernel::n_reduced_keys<Tk>(Param<Tk>(full.data(), kd, dense(kd)), &n_reduced, CParam<Tk>((const Tk *)keys, kd, dense(kd)));
    if (keys_out) memcpy(keys_out, full.data(), sizeof(Tk) * (size_t)n_reduced);
    if (!vals_out) return n_reduced;
    dim4 idims = D(dims), odims = D(dims);
    odims[dim] = n_reduced;
    Param<To> o((To *)vals_out, odims, dense(odims));
    CParam<Ti> v((const Ti *)vals, idims, D(strides));
    CParam<Tk> k((const Tk *)keys, kd, dense(kd));
    switch (ndims_of(idims)) {
        case 1: kernel::reduce_dim_by_key<op, Ti, Tk, To, 1>()(o, 0, k, v, 0, &n_reduced, dim, cn, nv); break;
        case 2: kernel::reduce_dim_by_key<op, Ti, Tk, To, 2>()(o, 0, k, v, 0, &n_reduced, dim, cn, nv); break;
        case 3: kernel::reduce_dim_by_key<op, Ti, Tk, To, 3>()(o, 0, k, v, 0, &n_reduced, dim, cn, nv); break;
        default: kernel::reduce_dim_by_key<op, Ti, Tk, To, 4>()(o, 0, k, v, 0, &n_reduced, dim, cn, nv); break;
    }
    return n_reduced;
}
template<af_op_t op, typename Ti, typename To>
long long do_rbk_key(int tk, const void *keys, const void *vals, const orc_dim_t *dims, const orc_dim_t *strides, int dim,
                     bool cn, double nv, void *keys_out, void *vals_out) {
    if (tk == ORC_S32) return do_rbk<op, Ti, To, int>(keys, vals, dims, strides, dim, cn, nv, keys_out, vals_out);
    if (tk == ORC_U32) return do_rbk<op, Ti, To, uint>(keys, vals, dims, strides, dim, cn, nv, keys_out, vals_out);
    return -1;
}
template<af_op_t op, typename T>
int do_ragged(const void *in, const orc_dim_t *dims, const orc_dim_t *strides, int dim, const uint32_t *rlen, void *out,
              uint32_t *loc) {
    dim4 idims = D(dims), odims = D(dims);
    if (idims.elements() == 0) return 0;
    odims[dim] = 1;
    Param<T> o((T *)out, odims, dense(odims));
    Param<uint> l((uint *)loc, odims, dense(odims));
    CParam<T> i((const T *)in, idims, D(strides));
    CParam<uint> r((const uint *)rlen, odims, dense(odims));
    switch (ndims_of(idims)) {
        case 1: kernel::ireduce_dim<op, T, 1>()(o, l, 0, i, 0, dim, r); break;
        case 2: kernel::ireduce_dim<op, T, 2>()(o, l, 0, i, 0, dim, r); break;
        case 3: kernel::ireduce_dim<op, T, 3>()(o, l, 0, i, 0, dim, r); break;
        default: kernel::ireduce_dim<op, T, 4>()(o, l, 0, i, 0, dim, r); break;
    }
    return 0;
}
}  // namespace

extern "C" {

long long ref_reduce_by_key(int op, int ti, int to, int tk, const void *keys, const void *vals, const orc_dim_t dims[4],
                            const orc_dim_t strides[4], int dim, int cn, double nv, void *keys_out, void *vals_out) {
#define ARGS tk, keys, vals, dims, strides, dim, cn != 0, nv, keys_out, vals_out
    if (op == ORC_OP_MIN || op == ORC_OP_MAX) {
#define X(c, T) if (ti == c && to == c) return op == ORC_OP_MIN ? do_rbk_key<af_min_t, T, T>(ARGS) : do_rbk_key<af_max_t, T, T>(ARGS);
        ALL_TYPES(X)
#undef X
    } else if (op == ORC_OP_ADD || op == ORC_OP_MUL) {
#define X(ci, Ti, co, To) if (ti == ci && to == co) return op == ORC_OP_ADD ? do_rbk_key<af_add_t, Ti, To>(ARGS) : do_rbk_key<af_mul_t, Ti, To>(ARGS);
        SUM_PAIRS(X)
#undef X
    } else if (op == ORC_OP_NOTZERO) {
#define X(c, T) if (ti == c && to == ORC_U32) return do_rbk_key<af_notzero_t, T, uint>(ARGS);
        ALL_TYPES(X)
#undef X
    } else if (op == ORC_OP_AND || op == ORC_OP_OR) {
#define X(c, T) if (ti == c && to == ORC_B8) return op == ORC_OP_AND ? do_rbk_key<af_and_t, T, char>(ARGS) : do_rbk_key<af_or_t, T, char>(ARGS);
        ALL_TYPES(X)
#undef X
    }
#undef ARGS
    return -1;
}

int ref_ireduce_ragged(int op, int t, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], int dim,
                       const uint32_t *rlen, void *out_val, uint32_t *out_loc) {
    if (op != ORC_OP_MIN && op != ORC_OP_MAX) return -1;
#define X(c, T) if (t == c) return op == ORC_OP_MIN ? do_ragged<af_min_t, T>(in, dims, strides, dim, rlen, out_val, out_loc) : do_ragged<af_max_t, T>(in, dims, strides, dim, rlen, out_val, out_loc);
    ALL_TYPES(X)
#undef X
    return -1;
}

}  // extern "C"

// ==== statistics: the reference's own running-mean kernel (kernel/mean.hpp); var / stdev are composed at the C-API level
// from JIT element-wise nodes there and are pinned through the restatement's use of this mean ===============================
namespace {
template<typename Ti, typename Tw, typename To>
int do_mean(const void *in, const orc_dim_t *dims, const orc_dim_t *strides, int dim, void *out) {
    dim4 idims = D(dims), odims = D(dims);
    if (idims.elements() == 0) return 0;
    odims[dim] = 1;
    Param<To> o((To *)out, odims, dense(odims));
    CParam<Ti> i((const Ti *)in, idims, D(strides));
    switch (ndims_of(idims)) {   // cpu/mean.cpp:40-45: mean_funcs[in.ndims() - 1]
        case 1: kernel::mean_dim<Ti, Tw, To, 1>()(o, 0, i, 0, dim); break;
        case 2: kernel::mean_dim<Ti, Tw, To, 2>()(o, 0, i, 0, dim); break;
        case 3: kernel::mean_dim<Ti, Tw, To, 3>()(o, 0, i, 0, dim); break;
        default: kernel::mean_dim<Ti, Tw, To, 4>()(o, 0, i, 0, dim); break;
    }
    return 0;
}
// cpu/mean.cpp:113-141 (needs Array<T>): the same MeanOp over the whole view
template<typename Ti, typename Tw, typename To>
To do_mean_all(const void *vin, const orc_dim_t *dims, const orc_dim_t *strides) {
    const Ti *in = (const Ti *)vin;
    kernel::MeanOp<compute_t<Ti>, compute_t<To>, compute_t<Tw>> Op(0, 0);
    for (orc_dim_t l = 0; l < dims[3]; l++)
        for (orc_dim_t k = 0; k < dims[2]; k++)
            for (orc_dim_t j = 0; j < dims[1]; j++)
                for (orc_dim_t i = 0; i < dims[0]; i++)
                    Op(compute_t<Ti>(in[i + j * strides[1] + k * strides[2] + l * strides[3]]), 1);
    return To(Op.runningMean);
}
}  // namespace

// the instantiation list of cpu/mean.cpp:143-158
#define FOR_MEAN_TYPES(X)                                                                               \
    X(ORC_F32, float, float, float) X(ORC_F64, double, double, double) X(ORC_C32, cfloat, float, cfloat)   \
    X(ORC_C64, cdouble, double, cdouble) X(ORC_S32, int, float, float) X(ORC_U32, uint, float, float)      \
    X(ORC_S64, intl, double, double) X(ORC_U64, uintl, double, double) X(ORC_S16, short, float, float)     \
    X(ORC_U16, ushort, float, float) X(ORC_S8, schar, float, float) X(ORC_U8, uchar, float, float)         \
    X(ORC_B8, char, float, float) X(ORC_F16, half, float, half)

extern "C" {
int ref_mean(int ti, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], int dim, void *out) {
#define X(code, Ti, Tw, To) if (ti == code) return do_mean<Ti, Tw, To>(in, dims, strides, dim, out);
    FOR_MEAN_TYPES(X)
#undef X
    return -1;
}
int ref_mean_all(int ti, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], double *out) {
    out[0] = out[1] = 0;
    if (ti == ORC_C32) { cfloat v = do_mean_all<cfloat, float, cfloat>(in, dims, strides); out[0] = v.real(); out[1] = v.imag(); return 0; }
    if (ti == ORC_C64) { cdouble v = do_mean_all<cdouble, double, cdouble>(in, dims, strides); out[0] = v.real(); out[1] = v.imag(); return 0; }
    if (ti == ORC_F16) { out[0] = (double)do_mean_all<half, float, float>(in, dims, strides); return 0; }
#define X(code, Ti, Tw, To) if (ti == code) { out[0] = (double)do_mean_all<Ti, Tw, To>(in, dims, strides); return 0; }
    X(ORC_F32, float, float, float) X(ORC_F64, double, double, double) X(ORC_S32, int, float, float)
    X(ORC_U32, uint, float, float) X(ORC_S64, intl, double, double) X(ORC_U64, uintl, double, double)
    X(ORC_S16, short, float, float) X(ORC_U16, ushort, float, float) X(ORC_S8, schar, float, float)
    X(ORC_U8, uchar, float, float) X(ORC_B8, char, float, float)
#undef X
    return -1;
}
}  // extern "C"
