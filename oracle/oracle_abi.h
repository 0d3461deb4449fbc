/*
 * oracle_abi.h -- C interface shared by the two CPU checkers under oracle/.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load these libraries.  libafb200.so never does.
 *
 * Two libraries export this same interface under different prefixes:
 *   orc_*  oracle/oracle.cpp      -- a fresh restatement of the reference's CPU
 *                                    backend semantics for the reduce / scan /
 *                                    where path (travels to the GPU box)
 *   ref_*  oracle/ref_driver.cpp  -- the reference's OWN cpu kernel templates
 *                                    (src/backend/cpu/kernel/{reduce,ireduce,
 *                                    scan,scan_by_key}.hpp) compiled in place
 *                                    from /root/reference into oracle/_ref/
 *
 * Conventions (match the reference's Param/CParam, src/backend/cpu/Param.hpp:23-118):
 *   - arrays are column-major views: dims[4], strides[4] in ELEMENTS; the
 *     caller has already applied the view offset to the pointer
 *   - outputs are freshly allocated dense arrays (strides = calcStrides(odims),
 *     src/backend/common/ArrayInfo.cpp:23-33)
 *   - dtype codes are af_dtype values (include/af/defines.h:210-233)
 *   - op codes are the ORC_OP_* below (ADD..MAX == af_binary_op,
 *     include/af/defines.h:433-438; the rest name af_op_t members of
 *     src/api/c/optypes.hpp)
 *   - every function returns 0 on success, -1 for a (op, Ti, To) triple the
 *     reference does not instantiate (Appendix A of SURVEY.md)
 */
#ifndef ORACLE_ABI_H
#define ORACLE_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_F32 = 0, ORC_C32 = 1, ORC_F64 = 2, ORC_C64 = 3, ORC_B8 = 4, ORC_S32 = 5, ORC_U32 = 6,
    ORC_U8 = 7, ORC_S64 = 8, ORC_U64 = 9, ORC_S16 = 10, ORC_U16 = 11, ORC_F16 = 12, ORC_S8 = 13
};

enum {
    ORC_OP_ADD = 0,     /* af_add_t     */
    ORC_OP_MUL = 1,     /* af_mul_t     */
    ORC_OP_MIN = 2,     /* af_min_t     */
    ORC_OP_MAX = 3,     /* af_max_t     */
    ORC_OP_NOTZERO = 4, /* af_notzero_t */
    ORC_OP_AND = 5,     /* af_and_t     */
    ORC_OP_OR = 6       /* af_or_t      */
};

#ifndef ORC_PREFIX
#define ORC_PREFIX orc
#endif
#define ORC_CAT2(a, b) a##_##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define ORC_FN(name) ORC_CAT(ORC_PREFIX, name)

typedef long long orc_dim_t;

/* detail::reduce<op,Ti,To>(in, dim, change_nan, nanval)  (cpu/reduce.cpp:48-66) */
int ORC_FN(reduce)(int op, int ti, int to, const void *in, const orc_dim_t dims[4],
                   const orc_dim_t strides[4], int dim, int change_nan, double nanval, void *out);

/* detail::reduce_all<op,Ti,To>(in, change_nan, nanval) -> 1 element of To (cpu/reduce.cpp:114-124) */
int ORC_FN(reduce_all)(int op, int ti, int to, const void *in, const orc_dim_t dims[4],
                       const orc_dim_t strides[4], int change_nan, double nanval, void *out);

/* detail::ireduce<op,T>(out, loc, in, dim)  (cpu/ireduce.cpp:31-43); op is MIN or MAX */
int ORC_FN(ireduce)(int op, int t, const void *in, const orc_dim_t dims[4],
                    const orc_dim_t strides[4], int dim, void *out_val, uint32_t *out_loc);

/* detail::ireduce_all<op,T>(&loc, in) (cpu/ireduce.cpp:59-90) */
int ORC_FN(ireduce_all)(int op, int t, const void *in, const orc_dim_t dims[4],
                        const orc_dim_t strides[4], void *out_val, uint32_t *out_loc);

/* detail::scan<op,Ti,To>(in, dim, inclusive) (cpu/scan.cpp:24-70) */
int ORC_FN(scan)(int op, int ti, int to, const void *in, const orc_dim_t dims[4],
                 const orc_dim_t strides[4], int dim, int inclusive, void *out);

/* detail::scan<op,Ti,Tk,To>(key, in, dim, inclusive), Ti == To (cpu/scan_by_key.cpp:22-40) */
int ORC_FN(scan_by_key)(int op, int t, int tk, const void *key, const orc_dim_t kstrides[4],
                        const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], int dim,
                        int inclusive, void *out);

/* detail::where<T>(in) (cpu/where.cpp:27-63): writes the hit indices, returns their count
 * (or -1 for an unsupported dtype).  `out` must hold dims.elements() entries.
 * The ref_* library restates this one loop too (where.cpp needs Array/memAlloc). */
long long ORC_FN(where)(int t, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4],
                        uint32_t *out);

/* detail::reduce_by_key<op,Ti,Tk,To>(keys_out, vals_out, keys, vals, dim, change_nan, nanval)
 * (cpu/reduce.cpp:73-108, cpu/kernel/reduce.hpp:64-160).  `keys`: n dense keys of dtype tk (s32 / u32), n == dims[dim].
 * Returns the number of runs (or -1 for an unsupported triple).  With vals_out == NULL only the count is computed;
 * otherwise keys_out holds one key per run and vals_out is the dense array with dims[dim] = runs. */
long long ORC_FN(reduce_by_key)(int op, int ti, int to, int tk, const void *keys, const void *vals,
                                const orc_dim_t dims[4], const orc_dim_t strides[4], int dim, int change_nan,
                                double nanval, void *keys_out, void *vals_out);

/* detail::rreduce<op,T>(out, loc, in, dim, rlen) (cpu/ireduce.cpp:47-57, kernel/ireduce.hpp:86-104): ireduce of every line
 * over its first min(len, rlen[line]) elements; rlen is dense over the output dims */
int ORC_FN(ireduce_ragged)(int op, int t, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4],
                           int dim, const uint32_t *rlen, void *out_val, uint32_t *out_loc);

/* ---- statistics.  mean: cpu/kernel/mean.hpp (running mean) -- exported by BOTH libraries;
 * var / var_all are composed at the reference's C-API level (api/c/var.cpp, stdev.cpp) and exist in the restatement only.
 * Output dtype of the dim-wise forms: f32 for integers <= 32 bits and b8, f64 for f64 / s64 / u64, c32, c64, f16 -> f16.
 * The *_all forms write 2 doubles (real, imag). */
int ORC_FN(mean)(int ti, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], int dim, void *out);
int ORC_FN(mean_all)(int ti, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], double *out);
int ORC_FN(var)(int ti, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], int dim, double denom,
                int take_sqrt, void *out, void *mean_out);
int ORC_FN(var_all)(int ti, const void *in, const orc_dim_t dims[4], const orc_dim_t strides[4], double denom, int take_sqrt,
                    double *out);

/* 1 if this library was built from the reference's own headers, 0 for the restatement */
int ORC_FN(is_reference)(void);

#ifdef __cplusplus
}
#endif
#endif
