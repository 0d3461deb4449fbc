// api_extra.cpp -- the array-handling entry points the reference's C++ wrapper (src/api/cpp/array.cpp, data.cpp,
// device.cpp, index.cpp) calls around the reduce / scan / where path, beyond the lifecycle set of api_array.cpp:
// reshaping views, type conversion, sequence indexing through af_index_gen, type predicates, the 64-bit / complex
// constants.  With them (and the NOT_SUPPORTED stubs of api_stubs.cpp for everything outside the path) the UNMODIFIED
// wrapper sources link against libafb200.so with no undefined symbol (tests/test_cpp_wrapper.py builds and runs that).
//
// Reference twins (argument checks, ownership and error codes restated, not copied):
//   af_moddims / af_flat          /root/reference/src/api/c/moddims.cpp:42-112
//   af_cast                       /root/reference/src/api/c/cast.cpp:34-84
//   af_index_gen                  /root/reference/src/api/c/index.cpp:185-300 (sequence indexers only here)
//   af_constant_complex/long/ulong /root/reference/src/api/c/data.cpp:66-118
//   af_is_*                        /root/reference/src/api/c/array.cpp:379-475, src/backend/common/ArrayInfo.cpp:108-165
//   af_eval_multiple, manual eval  /root/reference/src/api/c/array.cpp:258-330
//   af_device_info                 /root/reference/src/api/c/device.cpp:173-186
//   af_get_allocated_bytes         /root/reference/src/api/c/internal.cpp
#include <cstdlib>

#include "common.hpp"

using namespace afb;

namespace afb {
Array *make_dense_copy(const Array &a);
bool launch_cast_any(const LaunchCtx &c, af_dtype ti, af_dtype to, const View &in, void *out);
}  // namespace afb

// size_of(af_dtype): C++ linkage helper of the reference's C layer (src/api/c/type_util.hpp) that the wrapper's
// array.cpp / device.cpp call directly; a backend library that hosts the wrapper has to export it.
__attribute__((visibility("default"))) size_t size_of(af_dtype type) { return dtype_size(type); }

namespace {

void dims_from(unsigned ndims, const dim_t *dims, dim_t out[4]) {
    out[0] = out[1] = out[2] = out[3] = 1;
    for (unsigned i = 0; i < ndims && i < 4; i++) out[i] = dims[i];
}

Array *reshaped(const Array &in, const dim_t nd[4]) {
    // a linear array is re-labelled in place (shares the buffer); a strided view is materialised first
    if (in.is_linear()) {
        dim_t st[4];
        dense_strides(nd, st);
        return new_view(in, nd, in.info.offset, st);
    }
    Array *c = make_dense_copy(in);
    for (int i = 0; i < 4; i++) c->info.dims[i] = nd[i];
    dense_strides(c->info.dims, c->info.strides);
    return c;
}

bool type_is(af_dtype t, std::initializer_list<af_dtype> set) {
    for (af_dtype s : set)
        if (s == t) return true;
    return false;
}

template<typename F> af_err predicate(bool *result, const af_array arr, F f) {
    AFB_API_BEGIN
    AFB_ARG(0, result != nullptr);
    *result = f(*from_handle(arr));
    AFB_API_END
}

bool g_manual_eval = false;

}  // namespace

extern "C" {

af_err af_moddims(af_array *out, const af_array in_h, const unsigned ndims, const dim_t *const dims) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    if (ndims == 0) {  // moddims.cpp:45-48
        *out = retain(*in);
        return AF_SUCCESS;
    }
    AFB_ARG(3, dims != nullptr);
    AFB_ARG(2, ndims <= 4);
    dim_t nd[4];
    dims_from(ndims, dims, nd);
    if (nd[0] * nd[1] * nd[2] * nd[3] != in->elements())
        fail(AF_ERR_SIZE, "af_moddims: the new dimensions hold %lld elements, the array %lld",
             (long long)(nd[0] * nd[1] * nd[2] * nd[3]), (long long)in->elements());
    check_same_device(*in);
    *out = reshaped(*in, nd);
    AFB_API_END
}

// af_tile (/root/reference/src/api/c/tile.cpp:41-90): the array repeated x / y / z / w times along dims 0..3 -- here because the
// weighted mean broadcasts its weights with it (api/c/mean.cpp:112-124)
af_err af_tile(af_array *out, const af_array in_h, const unsigned x, const unsigned y, const unsigned z, const unsigned w) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    if (in->ndims() == 0) {                                  // tile.cpp:46
        *out = retain(*in);
        return AF_SUCCESS;
    }
    if (in->elements() <= 0) fail(AF_ERR_SIZE, "af_tile: the input holds no element");          // DIM_ASSERT(1, ...)
    if ((dim_t)x * y * z * w <= 0) fail(AF_ERR_SIZE, "af_tile: every repeat count must be at least 1");   // DIM_ASSERT(2, ...)
    check_same_device(*in);
    const dim_t od[4] = {in->info.dims[0] * x, in->info.dims[1] * y, in->info.dims[2] * z, in->info.dims[3] * w};
    Array *res = new_array(in->info.type, od);
    try {
        LaunchCtx c = ctx_for_active_device();
        launch_tile(c, in->info.type, view_of(*in), od, res->data());
    } catch (...) {
        delete res;
        throw;
    }
    *out = res;
    AFB_API_END
}

af_err af_flat(af_array *out, const af_array in_h) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    if (in->ndims() <= 1 && in->info.dims[0] == in->elements()) {  // already a column vector (moddims.cpp:95-98)
        *out = retain(*in);
        return AF_SUCCESS;
    }
    check_same_device(*in);
    const dim_t nd[4] = {in->elements(), 1, 1, 1};
    *out = reshaped(*in, nd);
    AFB_API_END
}

af_err af_cast(af_array *out, const af_array in_h, const af_dtype type) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    AFB_ARG(2, dtype_valid((int)type));
    Array *in = from_handle(in_h);
    const af_dtype ti = in->info.type;
    if (ti == type) {  // cast.cpp:39-41
        *out = retain(*in);
        return AF_SUCCESS;
    }
    if ((ti == c32 || ti == c64) && !(type == c32 || type == c64))
        fail(AF_ERR_TYPE, "af_cast: casting is not allowed from complex (%d) to real (%d) types", (int)ti, (int)type);
    check_same_device(*in);
    Array *res = new_array(type, in->info.dims);
    try {
        if (in->elements() > 0) {
            LaunchCtx c = ctx_for_active_device();
            if (!launch_cast_any(c, ti, type, view_of(*in), res->data())) AFB_TYPE_ERROR(1, ti);
        }
    } catch (...) {
        delete res;
        throw;
    }
    *out = res;
    AFB_API_END
}

// Sequence indexers are served by af_index (views sharing the parent buffer); array indexers (gather) are outside
// the path: AF_ERR_NOT_SUPPORTED.
af_err af_index_gen(af_array *out, const af_array in, const dim_t ndims, const af_index_t *indices) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    AFB_ARG(2, ndims > 0 && ndims <= 4);
    AFB_ARG(3, indices != nullptr);
    Array *src = from_handle(in);
    if (src->ndims() <= 0) {  // index.cpp:226-229: indexing an empty array gives an empty array
        const dim_t d[4] = {0, 1, 1, 1};
        *out = new_array(src->info.type, d);
        return AF_SUCCESS;
    }
    if (ndims == 1 && src->ndims() != 1) {  // linear indexing: flatten first (index.cpp:231-237)
        af_array flat = nullptr;
        af_err e = af_flat(&flat, in);
        if (e != AF_SUCCESS) return e;
        e = af_index_gen(out, flat, ndims, indices);
        af_release_array(flat);
        return e;
    }
    af_seq seqs[4] = {{1, 1, 0}, {1, 1, 0}, {1, 1, 0}, {1, 1, 0}};
    for (dim_t i = 0; i < ndims; i++) {
        if (!indices[i].isSeq)
            fail(AF_ERR_NOT_SUPPORTED, "af_index_gen: array indexers (gather) are not provided by libafb200; "
                                       "only af_seq / span indexing is on the reduce-scan-where path");
        seqs[i] = indices[i].idx.seq;
    }
    const unsigned used = (unsigned)ndims;
    return af_index(out, in, used, seqs);
    AFB_API_END
}

af_seq af_make_seq(double begin, double end, double step) { return af_seq{begin, end, step}; }

// ------------------------------------------------------------------ constants
af_err af_constant_complex(af_array *arr, const double real, const double imag, const unsigned ndims,
                           const dim_t *const dims, const af_dtype type) {
    AFB_API_BEGIN
    AFB_ARG(0, arr != nullptr);
    AFB_ARG(3, ndims >= 1 && ndims <= 4);
    AFB_ARG(4, dims != nullptr);
    if (!(type == c32 || type == c64)) AFB_TYPE_ERROR(5, type);
    dim_t d[4];
    dims_from(ndims, dims, d);
    for (int i = 0; i < 4; i++) AFB_ARG(4, d[i] >= 0);
    Array *a = new_array(type, d);
    try {
        LaunchCtx c = ctx_for_active_device();
        launch_fill(c, type, a->data(), a->elements(), real, imag);
    } catch (...) {
        delete a;
        throw;
    }
    *arr = a;
    AFB_API_END
}

// 64-bit constants must not travel through double (values above 2^53): upload one element and replicate it
static af_err constant_64(af_array *arr, const void *val, af_dtype type, const unsigned ndims, const dim_t *const dims) {
    AFB_API_BEGIN
    AFB_ARG(0, arr != nullptr);
    AFB_ARG(2, ndims >= 1 && ndims <= 4);
    AFB_ARG(3, dims != nullptr);
    dim_t d[4];
    dims_from(ndims, dims, d);
    for (int i = 0; i < 4; i++) AFB_ARG(3, d[i] >= 0);
    Array *a = new_array(type, d);
    try {
        const dim_t n = a->elements();
        if (n > 0) {
            LaunchCtx c = ctx_for_active_device();
            // a double pair fill writes the same 8-byte pattern: reinterpret the bits
            double bits;
            memcpy(&bits, val, 8);
            launch_fill(c, f64, a->data(), n, bits, 0.0);
        }
    } catch (...) {
        delete a;
        throw;
    }
    *arr = a;
    AFB_API_END
}
af_err af_constant_long(af_array *arr, const long long val, const unsigned ndims, const dim_t *const dims) {
    return constant_64(arr, &val, s64, ndims, dims);
}
af_err af_constant_ulong(af_array *arr, const unsigned long long val, const unsigned ndims, const dim_t *const dims) {
    return constant_64(arr, &val, u64, ndims, dims);
}

// ------------------------------------------------------------------ predicates
af_err af_is_row(bool *r, const af_array a) {
    return predicate(r, a, [](const Array &x) { return x.info.dims[0] == 1 && x.info.dims[1] > 1 && x.info.dims[2] == 1 && x.info.dims[3] == 1; });
}
af_err af_is_column(bool *r, const af_array a) {
    return predicate(r, a, [](const Array &x) { return x.info.dims[0] > 1 && x.info.dims[1] == 1 && x.info.dims[2] == 1 && x.info.dims[3] == 1; });
}
af_err af_is_complex(bool *r, const af_array a) { return predicate(r, a, [](const Array &x) { return type_is(x.info.type, {c32, c64}); }); }
af_err af_is_real(bool *r, const af_array a) { return predicate(r, a, [](const Array &x) { return !type_is(x.info.type, {c32, c64}); }); }
af_err af_is_double(bool *r, const af_array a) { return predicate(r, a, [](const Array &x) { return type_is(x.info.type, {f64, c64}); }); }
af_err af_is_single(bool *r, const af_array a) { return predicate(r, a, [](const Array &x) { return type_is(x.info.type, {f32, c32}); }); }
af_err af_is_half(bool *r, const af_array a) { return predicate(r, a, [](const Array &x) { return x.info.type == f16; }); }
af_err af_is_realfloating(bool *r, const af_array a) { return predicate(r, a, [](const Array &x) { return type_is(x.info.type, {f32, f64, f16}); }); }
af_err af_is_floating(bool *r, const af_array a) { return predicate(r, a, [](const Array &x) { return type_is(x.info.type, {f32, f64, f16, c32, c64}); }); }
af_err af_is_integer(bool *r, const af_array a) { return predicate(r, a, [](const Array &x) { return type_is(x.info.type, {s32, u32, s64, u64, s16, u16, s8, u8}); }); }
af_err af_is_bool(bool *r, const af_array a) { return predicate(r, a, [](const Array &x) { return x.info.type == b8; }); }
af_err af_is_sparse(bool *r, const af_array a) { return predicate(r, a, [](const Array &) { return false; }); }

af_err af_get_allocated_bytes(size_t *bytes, const af_array arr) {
    AFB_API_BEGIN
    AFB_ARG(0, bytes != nullptr);
    Array *a = from_handle(arr);
    *bytes = a->buf ? a->buf->bytes : 0;
    AFB_API_END
}

// ------------------------------------------------------------------ evaluation (nothing is lazy here)
af_err af_eval_multiple(const int num, af_array *arrays) {
    AFB_API_BEGIN
    AFB_ARG(1, arrays != nullptr || num == 0);
    for (int i = 0; i < num; i++) from_handle(arrays[i]);
    AFB_API_END
}
af_err af_set_manual_eval_flag(bool flag) {
    g_manual_eval = flag;
    return AF_SUCCESS;
}
af_err af_get_manual_eval_flag(bool *flag) {
    AFB_API_BEGIN
    AFB_ARG(0, flag != nullptr);
    *flag = g_manual_eval;
    AFB_API_END
}

af_err af_device_info(char *d_name, char *d_platform, char *d_toolkit, char *d_compute) {
    AFB_API_BEGIN
    cudaDeviceProp p;
    AFB_CUDA(cudaGetDeviceProperties(&p, active_device()));
    int rt = 0;
    cudaRuntimeGetVersion(&rt);
    // the reference documents 64-byte caller buffers (include/af/device.h:318-324)
    if (d_name) snprintf(d_name, 64, "%.63s", p.name);
    if (d_platform) snprintf(d_platform, 10, "CUDA");
    if (d_toolkit) snprintf(d_toolkit, 64, "v%d.%d", rt / 1000, (rt % 1000) / 10);
    if (d_compute) snprintf(d_compute, 10, "%d.%d", p.major, p.minor);
    AFB_API_END
}

}  // extern "C"
