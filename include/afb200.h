/*
 * afb200.h -- C-ABI of libafb200.so, the B200-native backend for ArrayFire's
 * reduce / scan / scan_by_key / where hot path.
 *
 * Every af_* symbol below has the SAME name, argument list, ownership and error
 * behaviour as the reference symbol it replaces, so the reference's C++ wrapper
 * (src/api/cpp/{reduce,scan,where,array,device}.cpp) or any C caller can bind to
 * this library for that path.  Citations are relative to /root/reference.
 * Types mirror include/af/defines.h (af_err :71-208, af_dtype :210-233,
 * af_array :243, dim_t :50, af_backend :417-421, af_binary_op :433-438) and
 * include/af/seq.h:20-30 (af_seq); a translation unit may include EITHER this header
 * or <arrayfire.h>, not both.
 *
 * The afx_* symbols at the end are this library's small extension for sharded
 * (one-process-per-GPU) whole-array ops: the per-shard partials that are combined
 * across ranks with NCCL allgather/allreduce by the host layer (arrayfire_b200/sharded.py).
 */
#ifndef AFB200_H
#define AFB200_H

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFAPI __attribute__((visibility("default")))

typedef long long dim_t;
typedef void *af_array;

typedef enum {
    AF_SUCCESS = 0,
    AF_ERR_NO_MEM = 101,
    AF_ERR_DRIVER = 102,
    AF_ERR_RUNTIME = 103,
    AF_ERR_INVALID_ARRAY = 201,
    AF_ERR_ARG = 202,
    AF_ERR_SIZE = 203,
    AF_ERR_TYPE = 204,
    AF_ERR_DIFF_TYPE = 205,
    AF_ERR_BATCH = 207,
    AF_ERR_DEVICE = 208,
    AF_ERR_NOT_SUPPORTED = 301,
    AF_ERR_NOT_CONFIGURED = 302,
    AF_ERR_NONFREE = 303,
    AF_ERR_NO_DBL = 401,
    AF_ERR_NO_GFX = 402,
    AF_ERR_NO_HALF = 403,
    AF_ERR_LOAD_LIB = 501,
    AF_ERR_LOAD_SYM = 502,
    AF_ERR_ARR_BKND_MISMATCH = 503,
    AF_ERR_INTERNAL = 998,
    AF_ERR_UNKNOWN = 999
} af_err;

typedef enum { f32, c32, f64, c64, b8, s32, u32, u8, s64, u64, s16, u16, f16, s8 } af_dtype;
typedef enum { afDevice, afHost } af_source;
typedef enum {
    AF_BACKEND_DEFAULT = 0, AF_BACKEND_CPU = 1, AF_BACKEND_CUDA = 2, AF_BACKEND_OPENCL = 4, AF_BACKEND_ONEAPI = 8
} af_backend;
typedef enum { AF_BINARY_ADD = 0, AF_BINARY_MUL = 1, AF_BINARY_MIN = 2, AF_BINARY_MAX = 3 } af_binary_op;
/* include/af/defines.h:525-529 */
typedef enum { AF_VARIANCE_DEFAULT = 0, AF_VARIANCE_SAMPLE = 1, AF_VARIANCE_POPULATION = 2 } af_var_bias;
/* include/af/defines.h:366-377 */
typedef enum {
    AF_NORM_VECTOR_1, AF_NORM_VECTOR_INF, AF_NORM_VECTOR_2, AF_NORM_VECTOR_P, AF_NORM_MATRIX_1, AF_NORM_MATRIX_INF,
    AF_NORM_MATRIX_2, AF_NORM_MATRIX_L_PQ, AF_NORM_EUCLID = AF_NORM_VECTOR_2
} af_norm_type;
typedef struct af_seq { double begin, end, step; } af_seq;
/* include/af/index.h:19-31 */
typedef struct af_index_t {
    union {
        af_array arr;
        af_seq seq;
    } idx;
    bool isSeq;
    bool isBatch;
} af_index_t;

/* ===================== hot path: reductions along a dimension ==========================
 * include/af/algorithm.h:708-776,1095-1140; src/api/c/reduce.cpp:461-497 */
AFAPI af_err af_sum(af_array *out, const af_array in, const int dim);                         /* reduce.cpp:469 */
AFAPI af_err af_sum_nan(af_array *out, const af_array in, const int dim, const double nanval); /* reduce.cpp:477 */
AFAPI af_err af_product(af_array *out, const af_array in, const int dim);                     /* reduce.cpp:473 */
AFAPI af_err af_product_nan(af_array *out, const af_array in, const int dim, const double nanval); /* reduce.cpp:482 */
AFAPI af_err af_min(af_array *out, const af_array in, const int dim);                         /* reduce.cpp:461 */
AFAPI af_err af_max(af_array *out, const af_array in, const int dim);                         /* reduce.cpp:465 */
AFAPI af_err af_count(af_array *out, const af_array in, const int dim);                       /* reduce.cpp:487 */
AFAPI af_err af_all_true(af_array *out, const af_array in, const int dim);                    /* reduce.cpp:491 */
AFAPI af_err af_any_true(af_array *out, const af_array in, const int dim);                    /* reduce.cpp:495 */

/* ===================== hot path: whole-array reductions ================================
 * src/api/c/reduce.cpp:862-916,1150-1168 */
AFAPI af_err af_sum_all(double *real, double *imag, const af_array in);
AFAPI af_err af_sum_nan_all(double *real, double *imag, const af_array in, const double nanval);
AFAPI af_err af_product_all(double *real, double *imag, const af_array in);
AFAPI af_err af_product_nan_all(double *real, double *imag, const af_array in, const double nanval);
AFAPI af_err af_min_all(double *real, double *imag, const af_array in);
AFAPI af_err af_max_all(double *real, double *imag, const af_array in);
AFAPI af_err af_count_all(double *real, double *imag, const af_array in);
AFAPI af_err af_all_true_all(double *real, double *imag, const af_array in);
AFAPI af_err af_any_true_all(double *real, double *imag, const af_array in);
AFAPI af_err af_sum_all_array(af_array *out, const af_array in);
AFAPI af_err af_sum_nan_all_array(af_array *out, const af_array in, const double nanval);
AFAPI af_err af_product_all_array(af_array *out, const af_array in);
AFAPI af_err af_product_nan_all_array(af_array *out, const af_array in, const double nanval);
AFAPI af_err af_min_all_array(af_array *out, const af_array in);
AFAPI af_err af_max_all_array(af_array *out, const af_array in);
AFAPI af_err af_count_all_array(af_array *out, const af_array in);
AFAPI af_err af_all_true_all_array(af_array *out, const af_array in);
AFAPI af_err af_any_true_all_array(af_array *out, const af_array in);

/* ===================== hot path: index-returning reductions ============================
 * src/api/c/reduce.cpp:994-1000 (ireduce_common :949-992), :1140-1148 (ireduce_all_common :1080-1138) */
AFAPI af_err af_imin(af_array *out, af_array *idx, const af_array in, const int dim);
AFAPI af_err af_imax(af_array *out, af_array *idx, const af_array in, const int dim);
AFAPI af_err af_imin_all(double *real, double *imag, unsigned *idx, const af_array in);
AFAPI af_err af_imax_all(double *real, double *imag, unsigned *idx, const af_array in);

/* ===================== next row: reductions by key, ragged max ========
 * src/api/c/reduce.cpp:500-554, 1076-1078; include/af/algorithm.h:777-1093.  `keys` is a vector of s32 / u32 with
 * vals.dims[dim] elements; runs of equal consecutive keys are reduced; keys_out gets one key per run. */
AFAPI af_err af_sum_by_key(af_array *keys_out, af_array *vals_out, const af_array keys, const af_array vals, const int dim);
AFAPI af_err af_sum_by_key_nan(af_array *keys_out, af_array *vals_out, const af_array keys, const af_array vals,
                               const int dim, const double nanval);
AFAPI af_err af_product_by_key(af_array *keys_out, af_array *vals_out, const af_array keys, const af_array vals, const int dim);
AFAPI af_err af_product_by_key_nan(af_array *keys_out, af_array *vals_out, const af_array keys, const af_array vals,
                                   const int dim, const double nanval);
AFAPI af_err af_min_by_key(af_array *keys_out, af_array *vals_out, const af_array keys, const af_array vals, const int dim);
AFAPI af_err af_max_by_key(af_array *keys_out, af_array *vals_out, const af_array keys, const af_array vals, const int dim);
AFAPI af_err af_count_by_key(af_array *keys_out, af_array *vals_out, const af_array keys, const af_array vals, const int dim);
AFAPI af_err af_all_true_by_key(af_array *keys_out, af_array *vals_out, const af_array keys, const af_array vals, const int dim);
AFAPI af_err af_any_true_by_key(af_array *keys_out, af_array *vals_out, const af_array keys, const af_array vals, const int dim);
/* (value, index) maximum of every line along dim over its first ragged_len[line] elements (u32 lengths) */
AFAPI af_err af_max_ragged(af_array *val, af_array *idx, const af_array in, const af_array ragged_len, const int dim);

/* ===================== hot path: scans =================================================
 * src/api/c/scan.cpp:118-157 (af_accum), :159-223 (af_scan), :225-283 (af_scan_by_key) */
AFAPI af_err af_accum(af_array *out, const af_array in, const int dim);
AFAPI af_err af_scan(af_array *out, const af_array in, const int dim, af_binary_op op, bool inclusive_scan);
AFAPI af_err af_scan_by_key(af_array *out, const af_array key, const af_array in, const int dim, af_binary_op op,
                            bool inclusive_scan);

/* ===================== hot path: stream compaction =====================================
 * src/api/c/where.cpp:34-65 */
AFAPI af_err af_where(af_array *idx, const af_array in);

/* ===================== array lifecycle around the path =================================
 * src/api/c/array.cpp:39-475, data.cpp:45, index.cpp:84-124 */
AFAPI af_err af_create_array(af_array *arr, const void *const data, const unsigned ndims, const dim_t *const dims,
                             const af_dtype type);                                           /* array.cpp:39 */
AFAPI af_err af_create_handle(af_array *arr, const unsigned ndims, const dim_t *const dims, const af_dtype type);
AFAPI af_err af_copy_array(af_array *arr, const af_array in);
AFAPI af_err af_write_array(af_array arr, const void *data, const size_t bytes, af_source src);
AFAPI af_err af_get_data_ptr(void *data, const af_array arr);
AFAPI af_err af_release_array(af_array arr);
AFAPI af_err af_retain_array(af_array *out, const af_array in);
AFAPI af_err af_get_data_ref_count(int *use_count, const af_array in);
AFAPI af_err af_eval(af_array in);
AFAPI af_err af_get_elements(dim_t *elems, const af_array arr);
AFAPI af_err af_get_type(af_dtype *type, const af_array arr);
AFAPI af_err af_get_dims(dim_t *d0, dim_t *d1, dim_t *d2, dim_t *d3, const af_array arr);
AFAPI af_err af_get_numdims(unsigned *result, const af_array arr);
AFAPI af_err af_is_empty(bool *result, const af_array arr);
AFAPI af_err af_is_scalar(bool *result, const af_array arr);
AFAPI af_err af_is_vector(bool *result, const af_array arr);
AFAPI af_err af_get_scalar(void *output_value, const af_array arr);
AFAPI af_err af_get_strides(dim_t *s0, dim_t *s1, dim_t *s2, dim_t *s3, const af_array arr);   /* internal.cpp */
AFAPI af_err af_get_offset(dim_t *offset, const af_array arr);                               /* internal.cpp */
AFAPI af_err af_is_linear(bool *result, const af_array arr);                                 /* internal.cpp */
AFAPI af_err af_constant(af_array *arr, const double val, const unsigned ndims, const dim_t *const dims,
                         const af_dtype type);                                               /* data.cpp:45 */
AFAPI af_err af_index(af_array *out, const af_array in, const unsigned ndims, const af_seq *const index); /* index.cpp:84 */
AFAPI af_err af_get_size_of(size_t *size, af_dtype type);

/* ===================== array handling the C++ wrapper needs around the path ============
 * (src/api/cpp/array.cpp, data.cpp, index.cpp call these; api_extra.cpp) */
AFAPI af_err af_moddims(af_array *out, const af_array in, const unsigned ndims, const dim_t *const dims); /* moddims.cpp:42 */
AFAPI af_err af_flat(af_array *out, const af_array in);                                        /* moddims.cpp:84 */
AFAPI af_err af_tile(af_array *out, const af_array in, const unsigned x, const unsigned y, const unsigned z,
                     const unsigned w);                                                    /* tile.cpp:82 */
AFAPI af_err af_cast(af_array *out, const af_array in, const af_dtype type);                   /* cast.cpp:34 */
/* af_seq / span indexers only (views); array indexers return AF_ERR_NOT_SUPPORTED.  index.cpp:217 */
AFAPI af_err af_index_gen(af_array *out, const af_array in, const dim_t ndims, const af_index_t *indices);
AFAPI af_seq af_make_seq(double begin, double end, double step);                               /* seq.h:242 */
AFAPI af_err af_constant_complex(af_array *arr, const double real, const double imag, const unsigned ndims,
                                 const dim_t *const dims, const af_dtype type);               /* data.cpp:66 */
AFAPI af_err af_constant_long(af_array *arr, const long long val, const unsigned ndims, const dim_t *const dims);
AFAPI af_err af_constant_ulong(af_array *arr, const unsigned long long val, const unsigned ndims, const dim_t *const dims);
AFAPI af_err af_is_row(bool *result, const af_array arr);                                      /* array.cpp:379-475 */
AFAPI af_err af_is_column(bool *result, const af_array arr);
AFAPI af_err af_is_complex(bool *result, const af_array arr);
AFAPI af_err af_is_real(bool *result, const af_array arr);
AFAPI af_err af_is_double(bool *result, const af_array arr);
AFAPI af_err af_is_single(bool *result, const af_array arr);
AFAPI af_err af_is_half(bool *result, const af_array arr);
AFAPI af_err af_is_realfloating(bool *result, const af_array arr);
AFAPI af_err af_is_floating(bool *result, const af_array arr);
AFAPI af_err af_is_integer(bool *result, const af_array arr);
AFAPI af_err af_is_bool(bool *result, const af_array arr);
AFAPI af_err af_is_sparse(bool *result, const af_array arr);
AFAPI af_err af_get_allocated_bytes(size_t *bytes, const af_array arr);                       /* internal.cpp */
AFAPI af_err af_eval_multiple(const int num, af_array *arrays);                                /* array.cpp:258 */
AFAPI af_err af_set_manual_eval_flag(bool flag);
AFAPI af_err af_get_manual_eval_flag(bool *flag);
AFAPI af_err af_device_info(char *d_name, char *d_platform, char *d_toolkit, char *d_compute); /* device.cpp:173 */

/* ===================== device / memory / error =========================================
 * src/api/c/device.cpp:57-405, memory.cpp:54-370, error.cpp:19-45, include/af/cuda.h:50-75 */
AFAPI af_err af_init(void);
AFAPI af_err af_info(void);
AFAPI af_err af_info_string(char **str, const bool verbose);
AFAPI af_err af_get_device_count(int *num_of_devices);
AFAPI af_err af_set_device(const int device);
AFAPI af_err af_get_device(int *device);
AFAPI af_err af_sync(const int device);
AFAPI af_err af_get_dbl_support(bool *available, const int device);
AFAPI af_err af_get_half_support(bool *available, const int device);
AFAPI af_err af_alloc_device(void **ptr, const dim_t bytes);
AFAPI af_err af_free_device(void *ptr);
AFAPI af_err af_alloc_device_v2(void **ptr, const dim_t bytes);
AFAPI af_err af_free_device_v2(void *ptr);
AFAPI af_err af_alloc_pinned(void **ptr, const dim_t bytes);
AFAPI af_err af_free_pinned(void *ptr);
AFAPI af_err af_alloc_host(void **ptr, const dim_t bytes);
AFAPI af_err af_free_host(void *ptr);
AFAPI af_err af_device_array(af_array *arr, void *data, const unsigned ndims, const dim_t *const dims,
                             const af_dtype type);
AFAPI af_err af_device_mem_info(size_t *alloc_bytes, size_t *alloc_buffers, size_t *lock_bytes, size_t *lock_buffers);
AFAPI af_err af_print_mem_info(const char *msg, const int device_id);
AFAPI af_err af_device_gc(void);
AFAPI af_err af_set_mem_step_size(const size_t step_bytes);
AFAPI af_err af_get_mem_step_size(size_t *step_bytes);
AFAPI af_err af_lock_array(const af_array arr);
AFAPI af_err af_unlock_array(const af_array arr);
AFAPI af_err af_lock_device_ptr(const af_array arr);
AFAPI af_err af_unlock_device_ptr(const af_array arr);
AFAPI af_err af_is_locked_array(bool *res, const af_array arr);
AFAPI af_err af_get_device_ptr(void **ptr, const af_array arr);
AFAPI void af_get_last_error(char **msg, dim_t *len);                                        /* error.cpp:19-37 */
AFAPI const char *af_err_to_string(const af_err err);                                        /* error.cpp:39-45 */
AFAPI af_err af_get_version(int *major, int *minor, int *patch);
AFAPI af_err af_get_backend_id(af_backend *backend, const af_array in);
AFAPI af_err af_get_device_id(int *device, const af_array in);
AFAPI af_err af_get_active_backend(af_backend *backend);
AFAPI af_err af_get_backend_count(unsigned *num_backends);
AFAPI af_err af_get_available_backends(int *backends);
AFAPI af_err af_set_backend(const af_backend bknd);
AFAPI af_err afcu_get_stream(void **stream /* cudaStream_t* */, int id);                     /* cuda.h:50 */
AFAPI af_err afcu_get_native_id(int *nativeid, int id);                                      /* cuda.h:63 */
AFAPI af_err afcu_set_native_id(int nativeid);                                               /* cuda.h:75 */

/* ===================== statistics on the reduction path =====================
 * include/af/statistics.h:348-659, include/af/lapack.h:568.  One fused streaming pass per input (kernels/stats_*.cu);
 * floating-point results agree with the reference within 1e-5 (f32) / 1e-12 (f64) relative, not bit for bit.
 * The weighted forms (real inputs; weights of the input's dims) are compositions of those passes; weighted statistics of
 * complex / f16 inputs return AF_ERR_NOT_SUPPORTED. */
AFAPI af_err af_mean(af_array *out, const af_array in, const dim_t dim);                                  /* mean.cpp:65 */
AFAPI af_err af_mean_all(double *real, double *imag, const af_array in);                                  /* mean.cpp:151 */
AFAPI af_err af_mean_weighted(af_array *out, const af_array in, const af_array weights, const dim_t dim);
AFAPI af_err af_mean_all_weighted(double *real, double *imag, const af_array in, const af_array weights);
AFAPI af_err af_var(af_array *out, const af_array in, const bool isbiased, const dim_t dim);              /* var.cpp:187 */
AFAPI af_err af_var_v2(af_array *out, const af_array in, const af_var_bias bias, const dim_t dim);        /* var.cpp:194 */
AFAPI af_err af_var_weighted(af_array *out, const af_array in, const af_array weights, const dim_t dim); /* var.cpp:263 */
AFAPI af_err af_var_all(double *realVal, double *imagVal, const af_array in, const bool isbiased);        /* var.cpp:336 */
AFAPI af_err af_var_all_v2(double *realVal, double *imagVal, const af_array in, const af_var_bias bias);  /* var.cpp:343 */
AFAPI af_err af_var_all_weighted(double *realVal, double *imagVal, const af_array in, const af_array weights);
AFAPI af_err af_meanvar(af_array *mean, af_array *var, const af_array in, const af_array weights, const af_var_bias bias,
                        const dim_t dim);                                                                 /* var.cpp:418 */
AFAPI af_err af_stdev(af_array *out, const af_array in, const dim_t dim);                                 /* stdev.cpp:130 */
AFAPI af_err af_stdev_v2(af_array *out, const af_array in, const af_var_bias bias, const dim_t dim);      /* stdev.cpp:134 */
AFAPI af_err af_stdev_all(double *real, double *imag, const af_array in);                                 /* stdev.cpp:97 */
AFAPI af_err af_stdev_all_v2(double *real, double *imag, const af_array in, const af_var_bias bias);      /* stdev.cpp:101 */
AFAPI af_err af_cov(af_array *out, const af_array X, const af_array Y, const bool isbiased);              /* covariance.cpp:65 */
AFAPI af_err af_cov_v2(af_array *out, const af_array X, const af_array Y, const af_var_bias bias);        /* covariance.cpp:72 */
AFAPI af_err af_corrcoef(double *realVal, double *imagVal, const af_array X, const af_array Y);           /* corrcoef.cpp:63 */
AFAPI af_err af_norm(double *out, const af_array in, const af_norm_type type, const double p, const double q); /* norm.cpp:119 */

/* ===================== callers next to the scan / compaction path =====================
 * include/af/image.h (af_sat), include/af/algorithm.h (af_diff1, af_diff2, af_set_unique).  af_sat is two single-pass scans
 * (dim 0 by the pipelined kernel, dim 1 by the tiled look-back kernel); af_set_unique takes SORTED input only. */
AFAPI af_err af_sat(af_array *out, const af_array in);                                                    /* sat.cpp:32 */
AFAPI af_err af_diff1(af_array *out, const af_array in, const int dim);                                   /* diff.cpp:36 */
AFAPI af_err af_diff2(af_array *out, const af_array in, const int dim);                                   /* diff.cpp:76 */
AFAPI af_err af_set_unique(af_array *out, const af_array in, const bool is_sorted);                       /* set.cpp:32 */

/* ===================== lazily evaluated element-wise expressions =====================
 * include/af/arith.h.  These return UNEVALUATED arrays (the reference's JIT nodes, src/api/c/binary.cpp:40-120); a
 * whole-array reduction (af_sum_all, af_count_all, af_min_all, af_max_all, af_any_true_all, af_all_true_all) that consumes
 * one interprets the expression in its load path -- sum<float>(sqrt(x*x + y*y) < 1) of examples/benchmarks/pi.cpp streams x
 * and y once and writes nothing.  Any other consumer evaluates the expression into an array first.  f32 / f64 / b8 operands
 * of identical dims (`batch` is ignored); s32 / u32 / s16 / u16 / s8 / u8 operands in comparisons and logical operations
 * (compared in the reference's implicit() type, exactly); everything else returns AF_ERR_NOT_SUPPORTED.  af_constant of
 * these types is lazy as well (no memory until needed). */
AFAPI af_err af_add(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_sub(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_mul(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_div(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_minof(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_maxof(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_pow(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_lt(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_gt(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_le(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_ge(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_eq(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_neq(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_and(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_or(af_array *out, const af_array lhs, const af_array rhs, const bool batch);
AFAPI af_err af_not(af_array *out, const af_array in);
AFAPI af_err af_sqrt(af_array *out, const af_array in);
AFAPI af_err af_abs(af_array *out, const af_array in);

/* ===================== afx_*: sharded whole-array extension ============================
 * New functionality (the reference has no multi-GPU path).  One process
 * per GPU holds the shard [rank*N/G, (rank+1)*N/G) of a flattened array; each call below
 * computes the shard-local piece ON THE DEVICE, leaving its O(1) partial in `partial`
 * (device memory, so the host layer can hand it to NCCL without a sync). */

/* partial <- whole-shard reduction with the promoted accumulation type of af_<op>_all;
 * op: 0 sum, 1 product, 2 min, 3 max, 4 count, 5 all_true, 6 any_true.  *partial_type gets its dtype. */
AFAPI af_err afx_reduce_all_partial(void *partial, af_dtype *partial_type, const af_array in, int op,
                                    bool change_nan, double nanval);
/* (value, shard-local linear index) of imin (op 2) / imax (op 3) over the shard */
AFAPI af_err afx_ireduce_all_partial(void *partial_val, unsigned *partial_idx, const af_array in, int op);
/* scan along dim 0 of a 1-D shard whose running value BEFORE the shard is *carry (device pointer to one
 * element of the output dtype, or NULL for the identity).  With `key` non-NULL it is a segmented scan and
 * the carry applies only while the key equals *carry_key (the last key of the previous shard). */
AFAPI af_err afx_scan_shard(af_array *out, const af_array key, const af_array in, af_binary_op op,
                            bool inclusive_scan, const void *carry, const void *carry_key);
/* Segmented scan of a 1-D shard WITHOUT carry-in, plus its 32-byte summary in device memory `summary`:
 *   [0,16)  reduction of the shard's last open segment (one element of the output dtype)
 *   [16,20) u32, 1 if the shard contains a segment head (0: the whole shard is one run)
 *   [24,32) u64, position of the first segment head (~0 if none)
 * The host layer all-gathers (summary, first key, last key) and repairs the first segment with the call below. */
AFAPI af_err afx_scan_shard_local(af_array *out, void *summary, const af_array key, const af_array in,
                                  af_binary_op op, bool inclusive_scan);
/* out[i] = op(*carry, out[i]) for i < count, in place; carry is a HOST pointer to one element of out's dtype */
AFAPI af_err afx_scan_shard_fixup(af_array out, const void *carry, dim_t count, af_binary_op op);
/* af_where over a shard with every index offset by `base` (= rank*N/G); returns the hit count too */
AFAPI af_err afx_where_shard(af_array *idx, dim_t *count, const af_array in, unsigned base);
/* The same in two asynchronous halves, so that a multi-GPU caller pays ONE host synchronisation for the kernel
 * and the exchange of the counts together: _launch enqueues the compaction into a worst-case sized u32 array and
 * leaves the hit count (u64) in device memory at *count_dev (a per-launch slot of a 64-entry ring: valid until 63 further
 * count-returning calls have been issued on this device);
 * _finish trims the array to `count` hits once the caller has read the count back. */
AFAPI af_err afx_where_shard_launch(af_array *idx, void **count_dev, const af_array in, unsigned base);
AFAPI af_err afx_where_shard_finish(af_array *idx, dim_t count);
/* All-gather of one 64-byte record per rank WITHOUT a collective library: `peer_bufs[r]` is rank r's symmetric buffer
 * (>= 1088 bytes, zero-initialised, peer-mapped into this process -- e.g. torch.distributed._symmetric_memory), `payload`
 * this rank's record in device memory.  One CTA stores the record into every peer's buffer, flags it, waits for all
 * peers' flags and copies the world x 64 gathered bytes to `out` (host); the call returns when they have arrived.
 * `epoch` must be 1, 2, 3, ... on successive calls, identically on every rank (world <= 8). */
AFAPI af_err afx_peer_exchange(void *out, const void *payload, const unsigned long long *peer_bufs, int rank, int world,
                               unsigned epoch);
/* ---- the same ops with the cross-GPU combine FUSED into the kernels (kernels/peer_comm.cuh): the thread that holds the
 * shard's partial stores it into every rank's symmetric buffer over NVLink peer memory, waits for the others and folds the
 * world's partials in rank order -- no second launch, no collective-library call, no host round trip before the combine.
 * `peer_bufs[r]` = device address (in THIS process) of rank r's symmetric buffer: >= 2048 bytes, zero-initialised, mapped
 * into every rank (torch.distributed._symmetric_memory, or cuMemCreate + IPC handles).  Every rank must issue the same
 * afx_sharded_* calls in the same order (they are collectives); world <= 8. */
AFAPI af_err afx_comm_create(void **comm, int rank, int world, const unsigned long long *peer_bufs);
AFAPI af_err afx_comm_release(void *comm);
/* af_<op>_all of the GLOBAL array (op codes of afx_reduce_all_partial); identical result on every rank */
AFAPI af_err afx_sharded_reduce_all(double *real, double *imag, void *comm, const af_array shard, int op, bool change_nan,
                                    double nanval);
/* af_imin_all (op 2) / af_imax_all (op 3) of the global array; *idx is the GLOBAL index, shard_base this shard's offset */
AFAPI af_err afx_sharded_ireduce_all(double *real, double *imag, unsigned *idx, void *comm, const af_array shard, int op,
                                     unsigned shard_base);
/* this rank's slice of af_scan (key == NULL) / af_scan_by_key of the global 1-D array; fully asynchronous (two launches) */
AFAPI af_err afx_sharded_scan(af_array *out, void *comm, const af_array key, const af_array shard, af_binary_op op,
                              bool inclusive_scan);
/* this rank's slice of af_where of the global array (global indices): the slice starts at *offset of *total hits */
AFAPI af_err afx_sharded_where(af_array *idx, dim_t *offset, dim_t *total, void *comm, const af_array shard,
                               unsigned shard_base);
/* ---- issue / collect: the whole-array ops whose result is O(1) host data (a scalar, an index, a hit count), split in
 * two so that a caller can enqueue MANY ops back to back and collect the results afterwards -- the GPU never idles through
 * a host round trip between two ops, and (sharded) ranks that arrive at different times overlap their skew with work.
 * `comm` may be NULL: the op then runs on the calling device alone (af_<op>_all / af_imin_all / af_where semantics, with
 * `shard_base` added to the indices).  The kernel's last thread publishes the result into a pinned host record
 * (csrc/common.hpp HostRet); afx_pending_* spins on that record's flag -- no cudaMemcpy, no stream synchronisation.
 * At most 63 ops may be pending per device.  Every pending handle must be consumed by exactly one afx_pending_* call. */
typedef void *afx_pending;
AFAPI af_err afx_reduce_all_async(afx_pending *pending, void *comm, const af_array shard, int op, bool change_nan, double nanval);
AFAPI af_err afx_ireduce_all_async(afx_pending *pending, void *comm, const af_array shard, int op, unsigned shard_base);
AFAPI af_err afx_where_async(afx_pending *pending, void *comm, const af_array shard, unsigned shard_base);
/* wait for and decode a reduction (idx may be NULL unless the op was an index reduction); consumes the handle */
AFAPI af_err afx_pending_scalar(afx_pending pending, double *real, double *imag, unsigned *idx);
/* wait for a where: the indices trimmed to the hit count (exact-size copy when sparse), slice offset, global count */
AFAPI af_err afx_pending_where(afx_pending pending, af_array *idx, dim_t *offset, dim_t *total);
/* drop a pending op without collecting its result */
AFAPI af_err afx_pending_release(afx_pending pending);
/* number of CUDA kernels this library launched on the calling thread's device since the last reset */
AFAPI af_err afx_launch_count(unsigned long long *count, bool reset);

#ifdef __cplusplus
}
#endif
#endif /* AFB200_H */
