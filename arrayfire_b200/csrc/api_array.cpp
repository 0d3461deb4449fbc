// api_array.cpp -- array lifecycle, device, memory and error entry points of libafb200:
// the minimal set the reference's C++ wrapper needs around the reduce / scan / where path.
//
// Each function mirrors the argument checks, ownership rules and error codes of its reference twin:
//   array lifecycle   /root/reference/src/api/c/array.cpp:39-475
//   af_index (seqs)   /root/reference/src/api/c/index.cpp:62-124, src/backend/cuda/Array.cpp:386-417 (createSubArray),
//                     src/backend/common/ArrayInfo.cpp:184-219 (toDims/toOffset/toStride), common/dim4.cpp:109-144
//   af_constant       /root/reference/src/api/c/data.cpp:45
//   device / memory   /root/reference/src/api/c/device.cpp:57-405, memory.cpp:54-370
//   errors            /root/reference/src/api/c/error.cpp:19-45
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.hpp"

using namespace afb;

namespace afb {
Array *lazy_constant(af_dtype type, const dim_t dims[4], double val);
}

namespace {

void dims_from(unsigned ndims, const dim_t *dims, dim_t out[4]) {
    out[0] = 0;
    out[1] = out[2] = out[3] = 1;
    for (unsigned i = 0; i < ndims && i < 4; i++) out[i] = dims[i];
}

// dense copy of any view into a fresh array
Array *dense_copy(const Array &a) {
    Array *r = new_array(a.info.type, a.info.dims);
    try {
        const dim_t n = a.elements();
        if (n > 0) {
            LaunchCtx c = ctx_for_active_device();
            if (a.is_linear()) {
                AFB_CUDA(cudaMemcpyAsync(r->data(), a.data(), (size_t)n * dtype_size(a.info.type),
                                         cudaMemcpyDeviceToDevice, c.stream));
            } else {
                launch_copy_strided(c, a.info.type, view_of(a), r->data());
            }
        }
    } catch (...) {
        delete r;
        throw;
    }
    return r;
}

}  // namespace

namespace afb {
Array *make_dense_copy(const Array &a) { return dense_copy(a); }
}  // namespace afb

extern "C" {

// ------------------------------------------------------------------ creation / destruction
af_err af_create_array(af_array *arr, const void *const data, const unsigned ndims, const dim_t *const dims,
                       const af_dtype type) {
    AFB_API_BEGIN
    AFB_ARG(0, arr != nullptr);
    AFB_ARG(2, ndims >= 1 && ndims <= 4);
    AFB_ARG(3, dims != nullptr);
    AFB_ARG(4, dtype_valid((int)type));
    dim_t d[4];
    dims_from(ndims, dims, d);
    for (int i = 0; i < 4; i++) AFB_ARG(3, d[i] >= 1);  // array.cpp: DIM_ASSERT(dims[i] > 0)
    Array *a = new_array(type, d);
    try {
        const size_t bytes = (size_t)a->elements() * dtype_size(type);
        if (bytes) {
            AFB_ARG(1, data != nullptr);
            upload_sync(a->device(), a->data(), data, bytes);  // the host buffer may be reused as soon as we return
        }
    } catch (...) {
        delete a;
        throw;
    }
    *arr = a;
    AFB_API_END
}

af_err af_create_handle(af_array *arr, const unsigned ndims, const dim_t *const dims, const af_dtype type) {
    AFB_API_BEGIN
    AFB_ARG(0, arr != nullptr);
    AFB_ARG(1, ndims <= 4);
    if (ndims > 0) AFB_ARG(2, dims != nullptr);
    AFB_ARG(3, dtype_valid((int)type));
    dim_t d[4];
    dims_from(ndims, dims, d);  // ndims == 0 -> dim4(0): an empty array
    *arr = new_array(type, d);
    AFB_API_END
}

af_err af_device_array(af_array *arr, void *data, const unsigned ndims, const dim_t *const dims, const af_dtype type) {
    AFB_API_BEGIN
    AFB_ARG(0, arr != nullptr);
    AFB_ARG(1, data != nullptr);
    AFB_ARG(2, ndims >= 1 && ndims <= 4);
    AFB_ARG(3, dims != nullptr);
    AFB_ARG(4, dtype_valid((int)type));
    dim_t d[4];
    dims_from(ndims, dims, d);
    for (int i = 0; i < 4; i++) AFB_ARG(3, d[i] >= 1);
    Array *a = new Array();
    a->info.devId = (unsigned)active_device() | ((unsigned)AF_BACKEND_CUDA << 8);
    a->info.type = type;
    for (int i = 0; i < 4; i++) a->info.dims[i] = d[i];
    a->info.offset = 0;
    dense_strides(d, a->info.strides);
    a->info.is_sparse = false;
    try {
        a->buf = adopt_device_ptr(data, (size_t)a->elements() * dtype_size(type));
    } catch (...) {
        delete a;
        throw;
    }
    *arr = a;
    AFB_API_END
}

af_err af_copy_array(af_array *arr, const af_array in) {
    AFB_API_BEGIN
    AFB_ARG(0, arr != nullptr);
    Array *a = from_handle(in);
    check_same_device(*a);
    *arr = dense_copy(*a);
    AFB_API_END
}

af_err af_write_array(af_array arr, const void *data, const size_t bytes, af_source src) {
    AFB_API_BEGIN
    Array *a = from_handle(arr);
    check_same_device(*a);
    AFB_ARG(1, data != nullptr || bytes == 0);
    AFB_ARG(2, bytes <= (size_t)a->elements() * dtype_size(a->info.type));
    if (!a->is_linear()) fail(AF_ERR_NOT_SUPPORTED, "af_write_array: destination must be a linear array");
    cudaStream_t s = stream_of(a->device());
    AFB_CUDA(cudaMemcpyAsync(a->data(), data, bytes, src == afHost ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    if (src == afHost) AFB_CUDA(cudaStreamSynchronize(s));
    AFB_API_END
}

af_err af_get_data_ptr(void *data, const af_array arr) {
    AFB_API_BEGIN
    Array *a = from_handle(arr);
    const dim_t n = a->elements();
    if (n == 0) return AF_SUCCESS;
    AFB_ARG(0, data != nullptr);
    check_same_device(*a);
    const size_t bytes = (size_t)n * dtype_size(a->info.type);
    if (a->is_linear()) {
        download_sync(a->device(), data, a->data(), bytes);
    } else {
        Array *tmp = dense_copy(*a);
        try {
            download_sync(a->device(), data, tmp->data(), bytes);
        } catch (...) {
            delete tmp;
            throw;
        }
        delete tmp;
    }
    AFB_API_END
}

af_err af_release_array(af_array arr) {
    AFB_API_BEGIN
    if (arr == nullptr) return AF_SUCCESS;  // array.cpp:231
    Array *a = peek_handle(arr);
    a->magic = 0;
    delete a;
    AFB_API_END
}

af_err af_retain_array(af_array *out, const af_array in) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    *out = retain(*peek_handle(in));
    AFB_API_END
}

af_err af_get_data_ref_count(int *use_count, const af_array in) {
    AFB_API_BEGIN
    AFB_ARG(0, use_count != nullptr);
    *use_count = (int)from_handle(in)->buf.use_count();
    AFB_API_END
}

af_err af_eval(af_array in) {
    AFB_API_BEGIN
    from_handle(in);  // nothing is lazy here: every array is already materialised
    AFB_API_END
}

// ------------------------------------------------------------------ queries
af_err af_get_elements(dim_t *elems, const af_array arr) {
    AFB_API_BEGIN
    AFB_ARG(0, elems != nullptr);
    *elems = peek_handle(arr)->elements();
    AFB_API_END
}
af_err af_get_type(af_dtype *type, const af_array arr) {
    AFB_API_BEGIN
    AFB_ARG(0, type != nullptr);
    *type = peek_handle(arr)->info.type;
    AFB_API_END
}
af_err af_get_dims(dim_t *d0, dim_t *d1, dim_t *d2, dim_t *d3, const af_array arr) {
    AFB_API_BEGIN
    Array *a = peek_handle(arr);
    if (d0) *d0 = a->info.dims[0];
    if (d1) *d1 = a->info.dims[1];
    if (d2) *d2 = a->info.dims[2];
    if (d3) *d3 = a->info.dims[3];
    AFB_API_END
}
af_err af_get_strides(dim_t *s0, dim_t *s1, dim_t *s2, dim_t *s3, const af_array arr) {
    AFB_API_BEGIN
    Array *a = from_handle(arr);
    if (s0) *s0 = a->info.strides[0];
    if (s1) *s1 = a->info.strides[1];
    if (s2) *s2 = a->info.strides[2];
    if (s3) *s3 = a->info.strides[3];
    AFB_API_END
}
af_err af_get_offset(dim_t *offset, const af_array arr) {
    AFB_API_BEGIN
    AFB_ARG(0, offset != nullptr);
    *offset = from_handle(arr)->info.offset;
    AFB_API_END
}
af_err af_is_linear(bool *result, const af_array arr) {
    AFB_API_BEGIN
    AFB_ARG(0, result != nullptr);
    *result = from_handle(arr)->is_linear();
    AFB_API_END
}
af_err af_get_numdims(unsigned *result, const af_array arr) {
    AFB_API_BEGIN
    AFB_ARG(0, result != nullptr);
    *result = (unsigned)peek_handle(arr)->ndims();
    AFB_API_END
}
af_err af_is_empty(bool *result, const af_array arr) {
    AFB_API_BEGIN
    AFB_ARG(0, result != nullptr);
    *result = peek_handle(arr)->elements() == 0;
    AFB_API_END
}
af_err af_is_scalar(bool *result, const af_array arr) {
    AFB_API_BEGIN
    AFB_ARG(0, result != nullptr);
    *result = peek_handle(arr)->elements() == 1;
    AFB_API_END
}
af_err af_is_vector(bool *result, const af_array arr) {
    AFB_API_BEGIN
    AFB_ARG(0, result != nullptr);
    Array *a = peek_handle(arr);
    int big = 0;
    for (int i = 0; i < 4; i++) big += a->info.dims[i] > 1;
    *result = a->elements() > 0 && big == 1;
    AFB_API_END
}
af_err af_get_size_of(size_t *size, af_dtype type) {
    AFB_API_BEGIN
    AFB_ARG(0, size != nullptr);
    AFB_ARG(1, dtype_valid((int)type));
    *size = dtype_size(type);
    AFB_API_END
}

af_err af_get_scalar(void *output_value, const af_array arr) {
    AFB_API_BEGIN
    AFB_ARG(0, output_value != nullptr);
    Array *a = from_handle(arr);
    AFB_ARG(1, a->elements() >= 1);
    check_same_device(*a);
    cudaStream_t s = stream_of(a->device());
    AFB_CUDA(cudaMemcpyAsync(output_value, a->data(), dtype_size(a->info.type), cudaMemcpyDeviceToHost, s));
    AFB_CUDA(cudaStreamSynchronize(s));
    AFB_API_END
}

af_err af_constant(af_array *arr, const double val, const unsigned ndims, const dim_t *const dims, const af_dtype type) {
    AFB_API_BEGIN
    AFB_ARG(0, arr != nullptr);
    AFB_ARG(2, ndims >= 1 && ndims <= 4);
    AFB_ARG(3, dims != nullptr);
    AFB_ARG(4, dtype_valid((int)type));
    dim_t d[4];
    dims_from(ndims, dims, d);
    for (int i = 0; i < 4; i++) AFB_ARG(3, d[i] >= 0);
    if ((type == f32 || type == f64 || type == b8 || type == s32 || type == u32 || type == s16 || type == u16 || type == s8 || type == u8) &&
        d[0] * d[1] * d[2] * d[3] > 0) {
        // a ScalarNode, as in the reference (data.cpp:45-90): no memory until a consumer needs the values
        *arr = lazy_constant(type, d, val);
        return AF_SUCCESS;
    }
    Array *a = new_array(type, d);
    try {
        LaunchCtx c = ctx_for_active_device();
        launch_fill(c, type, a->data(), a->elements(), val, 0.0);
    } catch (...) {
        delete a;
        throw;
    }
    *arr = a;
    AFB_API_END
}

// ------------------------------------------------------------------ af_index with af_seq ranges
namespace {
bool seq_is_span(const af_seq &s) { return s.step == 0 && s.begin == 1 && s.end == 1; }
dim_t seq_elements(const af_seq &s) {
    if (s.step > DBL_MIN) return (dim_t)((s.end - s.begin) / std::fabs(s.step)) + 1;
    if (s.step < -DBL_MIN) return (dim_t)((s.begin - s.end) / std::fabs(s.step)) + 1;
    return (dim_t)0x7fffffffffffffffLL;
}
}  // namespace

af_err af_index(af_array *out, const af_array in, const unsigned ndims, const af_seq *const index) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    AFB_ARG(2, ndims > 0 && ndims <= 4);
    AFB_ARG(3, index != nullptr);
    Array *src = from_handle(in);
    check_same_device(*src);

    // canonical (non-negative) sequences, index.cpp:62-67,95-106
    af_seq seqs[4];
    for (unsigned i = 0; i < ndims; i++) {
        const dim_t len = src->info.dims[i];
        seqs[i] = index[i];
        if (!seq_is_span(index[i])) {
            seqs[i].begin = std::signbit(index[i].begin) ? (len + index[i].begin) : index[i].begin;
            seqs[i].end = std::signbit(index[i].end) ? (len + index[i].end) : index[i].end;
            AFB_ARG(3, seqs[i].begin >= 0. && seqs[i].end >= 0.);
            if (std::signbit(seqs[i].step)) AFB_ARG(3, seqs[i].begin >= seqs[i].end);
            else AFB_ARG(3, seqs[i].begin <= seqs[i].end);
        }
    }

    // a 1-D index into a multi-dimensional array addresses the flattened array (index.cpp:76-78);
    // non-linear parents are materialised first (Array.cpp:393-396)
    Array *parent = src;
    Array *owned = nullptr;
    const bool flat = (ndims == 1 && src->ndims() != 1);
    if (!src->is_linear()) {
        owned = dense_copy(*src);
        parent = owned;
    }
    try {
        dim_t pdims[4], pstrides[4];
        for (int i = 0; i < 4; i++) {
            pdims[i] = parent->info.dims[i];
            pstrides[i] = parent->info.strides[i];
        }
        if (flat) {
            pdims[0] = parent->elements();
            pdims[1] = pdims[2] = pdims[3] = 1;
            dense_strides(pdims, pstrides);
        }
        dim_t dims[4] = {pdims[0], pdims[1], pdims[2], pdims[3]};
        dim_t strides[4] = {pstrides[0], pstrides[1], pstrides[2], pstrides[3]};
        dim_t offset = parent->info.offset;
        for (unsigned i = 0; i < ndims; i++) {
            const af_seq &s = seqs[i];
            if (seq_is_span(s)) continue;
            AFB_ARG(3, s.begin < (double)pdims[i] && s.end < (double)pdims[i]);
            dims[i] = seq_elements(s);
            if (dims[i] > pdims[i]) fail(AF_ERR_SIZE, "af_index: size mismatch between input and output");
            offset += (dim_t)s.begin * pstrides[i];
            if (s.step != 0) strides[i] = pstrides[i] * (dim_t)s.step;
        }
        for (unsigned i = ndims; i < 4 && !flat; i++) dims[i] = pdims[i];
        Array *view = new_view(*parent, dims, offset, strides);
        // the path's kernels (like the reference's) require strides[0] == 1 and non-negative strides
        if (strides[0] != 1 || strides[1] < 0 || strides[2] < 0 || strides[3] < 0) {
            // generic gather copy handles any stride sign
            Array *dense = nullptr;
            try {
                dense = new_array(view->info.type, view->info.dims);
                LaunchCtx c = ctx_for_active_device();
                launch_copy_strided(c, view->info.type, view_of(*view), dense->data());
            } catch (...) {
                delete dense;
                delete view;
                throw;
            }
            delete view;
            view = dense;
        }
        *out = view;
    } catch (...) {
        delete owned;
        throw;
    }
    delete owned;
    AFB_API_END
}

// ------------------------------------------------------------------ device management
af_err af_init(void) {
    AFB_API_BEGIN
    if (device_count() <= 0) fail(AF_ERR_DRIVER, "no CUDA device available: libafb200 has no CPU fallback");
    active_device();
    AFB_API_END
}

af_err af_info_string(char **str, const bool) {
    AFB_API_BEGIN
    AFB_ARG(0, str != nullptr);
    std::string s = "afb200 (B200-native reduce/scan/where backend for ArrayFire 3.10 API)\n";
    const int n = device_count();
    for (int d = 0; d < n; d++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) != cudaSuccess) continue;
        char line[512];
        snprintf(line, sizeof(line), "%c[%d] %s, %zu MB, CUDA Compute %d.%d, %d SMs\n", d == active_device() ? '*' : '-', d,
                 p.name, (size_t)(p.totalGlobalMem >> 20), p.major, p.minor, p.multiProcessorCount);
        s += line;
    }
    char *m = (char *)malloc(s.size() + 1);
    if (!m) fail(AF_ERR_NO_MEM, "host allocation failed");
    memcpy(m, s.c_str(), s.size() + 1);
    *str = m;
    AFB_API_END
}

af_err af_info(void) {
    char *s = nullptr;
    af_err e = af_info_string(&s, false);
    if (e == AF_SUCCESS && s) {
        fputs(s, stdout);
        free(s);
    }
    return e;
}

af_err af_get_device_count(int *n) {
    AFB_API_BEGIN
    AFB_ARG(0, n != nullptr);
    *n = device_count();
    AFB_API_END
}
af_err af_set_device(const int device) {
    AFB_API_BEGIN
    AFB_ARG(0, device >= 0 && device < device_count());
    set_active_device(device);
    AFB_API_END
}
af_err af_get_device(int *device) {
    AFB_API_BEGIN
    AFB_ARG(0, device != nullptr);
    *device = active_device();
    AFB_API_END
}
af_err af_sync(const int device) {
    AFB_API_BEGIN
    const int d = device < 0 ? active_device() : device;
    AFB_ARG(0, d < device_count());
    AFB_CUDA(cudaStreamSynchronize(stream_of(d)));
    AFB_API_END
}
af_err af_get_dbl_support(bool *available, const int) {
    AFB_API_BEGIN
    AFB_ARG(0, available != nullptr);
    *available = true;
    AFB_API_END
}
af_err af_get_half_support(bool *available, const int) {
    AFB_API_BEGIN
    AFB_ARG(0, available != nullptr);
    *available = true;
    AFB_API_END
}

// ------------------------------------------------------------------ memory
af_err af_alloc_device(void **ptr, const dim_t bytes) {
    AFB_API_BEGIN
    AFB_ARG(0, ptr != nullptr);
    AFB_ARG(1, bytes >= 0);
    *ptr = pool_alloc_raw((size_t)bytes);
    AFB_API_END
}
af_err af_alloc_device_v2(void **ptr, const dim_t bytes) { return af_alloc_device(ptr, bytes); }
af_err af_free_device(void *ptr) {
    AFB_API_BEGIN
    pool_free_raw(ptr);
    AFB_API_END
}
af_err af_free_device_v2(void *ptr) { return af_free_device(ptr); }
af_err af_alloc_pinned(void **ptr, const dim_t bytes) {
    AFB_API_BEGIN
    AFB_ARG(0, ptr != nullptr);
    AFB_ARG(1, bytes >= 0);
    active_device();
    AFB_CUDA(cudaMallocHost(ptr, (size_t)(bytes > 0 ? bytes : 1)));
    AFB_API_END
}
af_err af_free_pinned(void *ptr) {
    AFB_API_BEGIN
    if (ptr) AFB_CUDA(cudaFreeHost(ptr));
    AFB_API_END
}
af_err af_alloc_host(void **ptr, const dim_t bytes) {
    AFB_API_BEGIN
    AFB_ARG(0, ptr != nullptr);
    *ptr = malloc((size_t)(bytes > 0 ? bytes : 1));
    if (!*ptr) fail(AF_ERR_NO_MEM, "host allocation failed");
    AFB_API_END
}
af_err af_free_host(void *ptr) {
    free(ptr);
    return AF_SUCCESS;
}
af_err af_device_mem_info(size_t *alloc_bytes, size_t *alloc_buffers, size_t *lock_bytes, size_t *lock_buffers) {
    AFB_API_BEGIN
    pool_stats(active_device(), alloc_bytes, alloc_buffers, lock_bytes, lock_buffers);
    AFB_API_END
}
af_err af_print_mem_info(const char *msg, const int device_id) {
    AFB_API_BEGIN
    const int d = device_id < 0 ? active_device() : device_id;
    size_t ab, an, lb, ln;
    pool_stats(d, &ab, &an, &lb, &ln);
    printf("%s\n  device %d: %zu buffers / %zu bytes allocated, %zu buffers / %zu bytes in use\n", msg ? msg : "", d, an, ab,
           ln, lb);
    AFB_API_END
}
af_err af_device_gc(void) {
    AFB_API_BEGIN
    pool_gc(active_device());
    AFB_API_END
}
af_err af_set_mem_step_size(const size_t step_bytes) {
    pool_set_step(step_bytes);
    return AF_SUCCESS;
}
af_err af_get_mem_step_size(size_t *step_bytes) {
    AFB_API_BEGIN
    AFB_ARG(0, step_bytes != nullptr);
    *step_bytes = pool_get_step();
    AFB_API_END
}
af_err af_lock_array(const af_array arr) {
    AFB_API_BEGIN
    from_handle(arr)->buf->user_locked = true;
    AFB_API_END
}
af_err af_unlock_array(const af_array arr) {
    AFB_API_BEGIN
    from_handle(arr)->buf->user_locked = false;
    AFB_API_END
}
af_err af_lock_device_ptr(const af_array arr) { return af_lock_array(arr); }
af_err af_unlock_device_ptr(const af_array arr) { return af_unlock_array(arr); }
af_err af_is_locked_array(bool *res, const af_array arr) {
    AFB_API_BEGIN
    AFB_ARG(0, res != nullptr);
    *res = from_handle(arr)->buf->user_locked;
    AFB_API_END
}
af_err af_get_device_ptr(void **ptr, const af_array arr) {
    AFB_API_BEGIN
    AFB_ARG(0, ptr != nullptr);
    Array *a = from_handle(arr);
    // the reference hands out the buffer of a linear owner and locks it (cuda/Array.hpp:128-133)
    if (!a->is_linear()) fail(AF_ERR_NOT_SUPPORTED, "af_get_device_ptr: array must be linear (af_copy_array it first)");
    a->buf->user_locked = true;
    *ptr = a->elements() ? a->data() : nullptr;
    AFB_API_END
}

// ------------------------------------------------------------------ errors / version / backend
void af_get_last_error(char **msg, dim_t *len) {
    std::string s = take_last_error();
    if (s.size() > 1024) s.resize(1024);
    if (len) *len = (dim_t)s.size();
    if (msg) {
        if (s.empty()) {
            *msg = nullptr;
        } else {
            *msg = (char *)malloc(s.size() + 1);
            if (*msg) memcpy(*msg, s.c_str(), s.size() + 1);
        }
    }
}

const char *af_err_to_string(const af_err err) {
    switch (err) {
        case AF_SUCCESS: return "Success";
        case AF_ERR_NO_MEM: return "Device out of memory";
        case AF_ERR_DRIVER: return "Driver not available or incompatible";
        case AF_ERR_RUNTIME: return "Runtime error ";
        case AF_ERR_INVALID_ARRAY: return "Invalid array";
        case AF_ERR_ARG: return "Invalid input argument";
        case AF_ERR_SIZE: return "Invalid input size";
        case AF_ERR_TYPE: return "Function does not support this data type";
        case AF_ERR_DIFF_TYPE: return "Input types are not the same";
        case AF_ERR_BATCH: return "Invalid batch configuration";
        case AF_ERR_DEVICE: return "Input does not belong to the current device.";
        case AF_ERR_NOT_SUPPORTED: return "Function not supported";
        case AF_ERR_NOT_CONFIGURED: return "Function not configured to build";
        case AF_ERR_NONFREE: return "Function unavailable. ArrayFire compiled without Non-Free algorithms support";
        case AF_ERR_NO_DBL: return "Double precision not supported for this device";
        case AF_ERR_NO_GFX: return "Graphics functionality unavailable. ArrayFire compiled without Graphics support";
        case AF_ERR_NO_HALF: return "Half precision floats not supported for this device";
        case AF_ERR_LOAD_LIB: return "Failed to load dynamic library. ";
        case AF_ERR_LOAD_SYM: return "Failed to load symbol";
        case AF_ERR_ARR_BKND_MISMATCH: return "There was a mismatch between an array and the current backend";
        case AF_ERR_INTERNAL: return "Internal error";
        case AF_ERR_UNKNOWN: return "Unknown error";
    }
    return "Unknown error";
}

af_err af_get_version(int *major, int *minor, int *patch) {
    if (major) *major = 3;
    if (minor) *minor = 10;
    if (patch) *patch = 0;
    return AF_SUCCESS;
}
af_err af_get_backend_id(af_backend *backend, const af_array in) {
    AFB_API_BEGIN
    AFB_ARG(0, backend != nullptr);
    from_handle(in);
    *backend = AF_BACKEND_CUDA;
    AFB_API_END
}
af_err af_get_device_id(int *device, const af_array in) {
    AFB_API_BEGIN
    AFB_ARG(0, device != nullptr);
    *device = from_handle(in)->device();
    AFB_API_END
}
af_err af_get_active_backend(af_backend *backend) {
    if (backend) *backend = AF_BACKEND_CUDA;
    return AF_SUCCESS;
}
af_err af_get_backend_count(unsigned *num_backends) {
    if (num_backends) *num_backends = 1;
    return AF_SUCCESS;
}
af_err af_get_available_backends(int *backends) {
    if (backends) *backends = AF_BACKEND_CUDA;
    return AF_SUCCESS;
}
af_err af_set_backend(const af_backend bknd) {
    if (bknd == AF_BACKEND_CUDA || bknd == AF_BACKEND_DEFAULT) return AF_SUCCESS;
    return store_error(AF_ERR_ARG, "af_set_backend: libafb200 only provides the CUDA backend");
}
af_err afcu_get_stream(void **stream, int id) {
    AFB_API_BEGIN
    AFB_ARG(0, stream != nullptr);
    *stream = (void *)stream_of(id);
    AFB_API_END
}
af_err afcu_get_native_id(int *nativeid, int id) {
    AFB_API_BEGIN
    AFB_ARG(0, nativeid != nullptr);
    AFB_ARG(1, id >= 0 && id < device_count());
    *nativeid = id;
    AFB_API_END
}
af_err afcu_set_native_id(int nativeid) { return af_set_device(nativeid); }

}  // extern "C"
