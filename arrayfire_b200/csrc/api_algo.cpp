// api_algo.cpp -- the hot-path C entry points: af_sum ... af_where and the afx_* shard helpers.
//
// Argument checks, early-outs, dtype promotion and error codes restate the reference's C-API layer:
//   /root/reference/src/api/c/reduce.cpp:82-121 (reduce_type), 190-228 (reduce_common), 305-378 (reduce_promote),
//       571-860 (whole-array forms), 918-1000 (ireduce), 1080-1168 (ireduce_all, *_nan_all)
//   /root/reference/src/api/c/scan.cpp:118-283, where.cpp:34-65
// Everything below the dispatch is this library's own CUDA path (kernels/*.cu); there is NO CPU fallback:
// with no usable device every entry point fails with AF_ERR_DRIVER.
#include <cstdlib>

#include <atomic>
#include <memory>

#include "common.hpp"

using namespace afb;

namespace afb {
Array *make_dense_copy(const Array &a);
bool expr_reduce_all(const Array &a, int mode, af_dtype *result_type, HostRet *ret_out);
}

namespace {

struct OpLock {
    std::lock_guard<std::mutex> g;
    explicit OpLock(int device) : g(op_mutex(device)) {}
};

struct Holder {  // deletes a half-built result on exceptions
    Array *a = nullptr;
    ~Holder() { delete a; }
    Array *release() { Array *r = a; a = nullptr; return r; }
};

// promotion of af_sum / af_product (reduce.cpp:322-371)
bool sum_promote(af_dtype ti, Op op, Op *eff, af_dtype *to) {
    *eff = op;
    switch (ti) {
        case f32: case f64: case c32: case c64: case u32: case s32: case u64: case s64: *to = ti; return true;
        case u16: case u8: *to = u32; return true;
        case s16: case s8: *to = s32; return true;
        case f16: *to = f32; return true;
        case b8:
            if (op == OP_MUL) { *eff = OP_AND; *to = b8; }
            else { *eff = OP_NOTZERO; *to = u32; }
            return true;
    }
    return false;
}

enum Family { FAM_PROMOTE, FAM_COMMON, FAM_TYPE };

void resolve(Family fam, Op op, af_dtype ti, Op *eff, af_dtype *to) {
    *eff = op;
    switch (fam) {
        case FAM_PROMOTE:
            if (!sum_promote(ti, op, eff, to)) AFB_TYPE_ERROR(1, ti);
            break;
        case FAM_COMMON: *to = ti; break;
        case FAM_TYPE: *to = (op == OP_NOTZERO) ? u32 : b8; break;
    }
}

// ---- along a dimension -------------------------------------------------------------------------
af_err reduce_dimwise(af_array *out, const af_array in_h, int dim, Family fam, Op op, bool change_nan, double nanval) {
    AFB_API_BEGIN
    AFB_ARG(2, dim >= 0);
    AFB_ARG(2, dim < 4);
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    if (dim >= in->ndims()) {  // also covers empty inputs; no promotion (reduce.cpp:314-317)
        *out = retain(*in);
        return AF_SUCCESS;
    }
    check_same_device(*in);
    Op eff;
    af_dtype to;
    resolve(fam, op, in->info.type, &eff, &to);
    dim_t od[4] = {in->info.dims[0], in->info.dims[1], in->info.dims[2], in->info.dims[3]};
    od[dim] = 1;
    Holder res;
    res.a = new_array(to, od);
    {
        OpLock lk(in->device());
        LaunchCtx c = ctx_for_active_device();
        if (!launch_reduce(c, eff, in->info.type, to, view_of(*in), dim, change_nan, nanval, res.a->data()))
            AFB_TYPE_ERROR(1, in->info.type);
    }
    *out = res.release();
    AFB_API_END
}

// ---- whole array ---------------------------------------------------------------------------------
// Runs the reduction into `dst` (device, one element of *to; null: the device slot paired with the host record).
// Non-linear views with short columns are materialised first so the flattening kernel streams long spans.
// With `ret` the kernel also publishes the value to a pinned host record (no D2H copy / stream synchronisation).
void reduce_all_into(const Array &in, Op eff, af_dtype to, bool change_nan, double nanval, void *dst, HostRet *ret = nullptr,
                     const ShardLink *link = nullptr) {
    const Array *src = &in;
    Holder tmp;
    if (!in.is_linear() && in.info.dims[0] < 4096) {
        tmp.a = make_dense_copy(in);
        src = tmp.a;
    }
    OpLock lk(in.device());
    LaunchCtx c = ctx_for_active_device();
    if (ret) {
        *ret = c.scratch->next_host_ret();
        if (!dst) dst = ret->dev;
    }
    if (!launch_reduce_all(c, eff, in.info.type, to, view_of(*src), change_nan, nanval, dst, ret, link))
        AFB_TYPE_ERROR(1, in.info.type);
}

double identity_of(Op op) { return (op == OP_MUL || op == OP_AND) ? 1.0 : 0.0; }

// one element of dtype t (host bytes) -> the (real, imag) doubles of the scalar API
void decode_scalar(af_dtype t, const void *raw, double *real, double *imag) {
    switch (t) {
        case f32: *real = *(const float *)raw; break;
        case f64: *real = *(const double *)raw; break;
        case c32: *real = ((const float *)raw)[0]; if (imag) *imag = ((const float *)raw)[1]; break;
        case c64: *real = ((const double *)raw)[0]; if (imag) *imag = ((const double *)raw)[1]; break;
        case s32: *real = *(const int *)raw; break;
        case u32: *real = *(const unsigned *)raw; break;
        case s64: *real = (double)*(const long long *)raw; break;
        case u64: *real = (double)*(const unsigned long long *)raw; break;
        case s16: *real = *(const short *)raw; break;
        case u16: *real = *(const unsigned short *)raw; break;
        case s8: case b8: *real = *(const signed char *)raw; break;
        case u8: *real = *(const unsigned char *)raw; break;
        case f16: {
            // IEEE binary16 -> double
            unsigned short h = *(const unsigned short *)raw;
            int sign = h >> 15, e = (h >> 10) & 31, m = h & 1023;
            double v;
            if (e == 0) v = std::ldexp((double)m, -24);
            else if (e == 31) v = m ? NAN : INFINITY;
            else v = std::ldexp((double)(m | 1024), e - 25);
            *real = sign ? -v : v;
        } break;
    }
}

af_err reduce_all_scalar(double *real, double *imag, const af_array in_h, Family fam, Op op, bool change_nan,
                         double nanval, bool need_ndims) {
    AFB_API_BEGIN
    // an unevaluated element-wise expression (api_expr.cpp) is reduced FUSED: the interpreter runs in the reduction's load
    // path, the operands are streamed once and no temporary is written
    {
        Array *lazy = peek_handle(in_h);
        if (lazy->expr && real != nullptr && lazy->elements() > 0 && !change_nan &&
            (lazy->info.type == f32 || lazy->info.type == f64 || lazy->info.type == b8) &&   // a lazy INTEGER constant takes the evaluated path
            (op == OP_ADD || op == OP_NOTZERO || op == OP_MIN || op == OP_MAX || op == OP_OR || op == OP_AND)) {
            check_same_device(*lazy);
            *real = 0;
            if (imag) *imag = 0;
            // sum of a b8 expression counts (reduce.cpp:358-366), like the evaluated path
            int mode = op == OP_ADD ? (lazy->info.type == b8 ? EXR_COUNT : EXR_SUM)
                     : op == OP_NOTZERO ? EXR_COUNT : op == OP_MIN ? EXR_MIN : op == OP_MAX ? EXR_MAX : op == OP_OR ? EXR_ANY : EXR_ALL;
            if ((op == OP_MIN || op == OP_MAX) && lazy->info.type == b8) mode = op == OP_MIN ? EXR_ALL : EXR_ANY;   // b8 min = and, max = or
            af_dtype rt;
            HostRet ret;
            if (expr_reduce_all(*lazy, mode, &rt, &ret)) {
                LaunchCtx c = ctx_for_active_device();
                wait_host_ret(c, ret);
                decode_scalar(rt, ret.host, real, imag);
                return AF_SUCCESS;
            }
        }
    }
    Array *in = from_handle(in_h);
    if (need_ndims) AFB_ARG(2, in->ndims() > 0);  // reduce.cpp:646
    AFB_ARG(0, real != nullptr);
    *real = 0;
    if (imag) *imag = 0;
    Op eff;
    af_dtype to;
    resolve(fam, op, in->info.type, &eff, &to);
    const bool cplx = (in->info.type == c32 || in->info.type == c64);
    if (cplx && fam != FAM_TYPE) AFB_ARG(1, imag != nullptr);
    if (cplx && fam == FAM_PROMOTE) change_nan = false;  // the scalar API drops nanval for complex (reduce.cpp:762-773)
    if (in->elements() == 0) {
        *real = identity_of(eff);
        return AF_SUCCESS;
    }
    check_same_device(*in);
    // the last CTA of the reduction writes the scalar straight into a pinned host record and raises its flag: the host
    // spins on the flag instead of paying cudaMemcpyAsync + cudaStreamSynchronize (the reference's getScalar,
    // src/backend/cuda/copy.cpp:178-184)
    HostRet ret;
    reduce_all_into(*in, eff, to, change_nan, nanval, nullptr, &ret);
    LaunchCtx c = ctx_for_active_device();
    wait_host_ret(c, ret);
    decode_scalar(to, ret.host, real, imag);
    AFB_API_END
}

af_err reduce_all_array(af_array *out, const af_array in_h, Family fam, Op op, bool change_nan, double nanval,
                        bool need_ndims) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    if (need_ndims) AFB_ARG(2, in->ndims() > 0);
    Op eff;
    af_dtype to;
    resolve(fam, op, in->info.type, &eff, &to);
    check_same_device(*in);
    Holder res;
    dim_t one[4] = {1, 1, 1, 1};
    res.a = new_array(to, one);
    if (in->elements() == 0) {
        LaunchCtx c = ctx_for_active_device();
        launch_fill(c, to, res.a->data(), 1, identity_of(eff), 0.0);
    } else {
        reduce_all_into(*in, eff, to, change_nan, nanval, res.a->data());
    }
    *out = res.release();
    AFB_API_END
}

// ---- ireduce -------------------------------------------------------------------------------------
af_err ireduce_dimwise(af_array *val, af_array *idx, const af_array in_h, int dim, Op op) {
    AFB_API_BEGIN
    AFB_ARG(3, dim >= 0);
    AFB_ARG(3, dim < 4);
    AFB_ARG(0, val != nullptr);
    AFB_ARG(1, idx != nullptr);
    Array *in = from_handle(in_h);
    AFB_ARG(2, in->ndims() > 0);
    check_same_device(*in);
    if (dim >= in->ndims()) {  // values = input, indices = 0 (reduce.cpp:959-963)
        Holder loc;
        loc.a = new_array(u32, in->info.dims);
        LaunchCtx c = ctx_for_active_device();
        launch_fill(c, u32, loc.a->data(), loc.a->elements(), 0.0, 0.0);
        *val = retain(*in);
        *idx = loc.release();
        return AF_SUCCESS;
    }
    dim_t od[4] = {in->info.dims[0], in->info.dims[1], in->info.dims[2], in->info.dims[3]};
    od[dim] = 1;
    Holder res, loc;
    res.a = new_array(in->info.type, od);
    loc.a = new_array(u32, od);
    {
        OpLock lk(in->device());
        LaunchCtx c = ctx_for_active_device();
        if (!launch_ireduce(c, op, in->info.type, view_of(*in), dim, res.a->data(), (unsigned *)loc.a->data()))
            AFB_TYPE_ERROR(1, in->info.type);
    }
    *val = res.release();
    *idx = loc.release();
    AFB_API_END
}

void ireduce_all_into(const Array &in, Op op, void *dval, unsigned *dloc, HostRet *ret = nullptr,
                      const ShardLink *link = nullptr) {
    const Array *src = &in;
    Holder tmp;
    if (!in.is_linear() && in.info.dims[0] < 4096) {
        tmp.a = make_dense_copy(in);
        src = tmp.a;
    }
    OpLock lk(in.device());
    LaunchCtx c = ctx_for_active_device();
    if (ret) {
        *ret = c.scratch->next_host_ret();
        if (!dval) {   // device copy goes to the slot paired with the record: value at +0, index at +32
            dval = ret->dev;
            dloc = (unsigned *)((char *)ret->dev + 32);
        }
    }
    if (!launch_ireduce_all(c, op, in.info.type, view_of(*src), dval, dloc, ret, link)) AFB_TYPE_ERROR(1, in.info.type);
}

af_err ireduce_all_scalar(double *real, double *imag, unsigned *loc, const af_array in_h, Op op) {
    AFB_API_BEGIN
    Array *in = from_handle(in_h);
    AFB_ARG(3, in->ndims() > 0);
    AFB_ARG(0, real != nullptr);
    AFB_ARG(2, loc != nullptr);
    *real = 0;
    if (imag) *imag = 0;
    const af_dtype t = in->info.type;
    if (t == f16) AFB_TYPE_ERROR(1, t);  // no f16 case in the reference (reduce.cpp:1095-1131)
    if (t == c32 || t == c64) AFB_ARG(1, imag != nullptr);
    check_same_device(*in);
    HostRet ret;
    ireduce_all_into(*in, op, nullptr, nullptr, &ret);
    LaunchCtx c = ctx_for_active_device();
    wait_host_ret(c, ret);
    decode_scalar(t, ret.host, real, imag);
    *loc = *(const unsigned *)((const char *)ret.host + 32);
    AFB_API_END
}

// ---- scans -----------------------------------------------------------------------------------------
bool scan_promote(af_dtype ti, af_dtype *to) {
    switch (ti) {
        case f32: case f64: case c32: case c64: case u32: case s32: case u64: case s64: *to = ti; return true;
        case u16: case u8: case b8: *to = u32; return true;
        case s16: case s8: *to = s32; return true;
        default: return false;
    }
}

af_err scan_common(af_array *out, const af_array in_h, int dim, int binop, bool inclusive, bool accum) {
    AFB_API_BEGIN
    AFB_ARG(2, dim >= 0);
    AFB_ARG(2, dim < 4);
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    if (dim >= in->ndims()) {
        *out = retain(*in);
        return AF_SUCCESS;
    }
    check_same_device(*in);
    af_dtype to;
    if (!scan_promote(in->info.type, &to)) AFB_TYPE_ERROR(1, in->info.type);
    Op op;
    if (accum) {
        op = (in->info.type == b8) ? OP_NOTZERO : OP_ADD;  // scan.cpp:146-148
    } else {
        if (binop < 0 || binop > 3) fail(AF_ERR_ARG, "Incorrect binary operation enum for argument number 2");
        op = (Op)binop;
    }
    Holder res;
    res.a = new_array(to, in->info.dims);
    {
        OpLock lk(in->device());
        LaunchCtx c = ctx_for_active_device();
        if (!launch_scan(c, op, in->info.type, to, view_of(*in), dim, inclusive, res.a->data(), nullptr))
            AFB_TYPE_ERROR(1, in->info.type);
    }
    *out = res.release();
    AFB_API_END
}

bool sbk_promote(af_dtype ti, af_dtype *to) {
    switch (ti) {
        case f32: case f64: case c32: case c64: case s32: case u32: case s64: case u64: *to = ti; return true;
        case s16: case s8: *to = s32; return true;
        case u16: case u8: case b8: *to = u32; return true;
        default: return false;
    }
}

// shared by af_scan_by_key and afx_scan_shard
Array *scan_by_key_impl(const Array &key, const Array &in, int dim, int binop, bool inclusive, const void *carry,
                        const void *carry_key, void *summary = nullptr) {
    for (int i = 0; i < 4; i++) AFB_ARG(2, in.info.dims[i] == key.info.dims[i]);
    af_dtype to;
    if (!sbk_promote(in.info.type, &to)) AFB_TYPE_ERROR(1, in.info.type);
    if (binop < 0 || binop > 3) fail(AF_ERR_ARG, "Incorrect binary operation enum for argument number 3");
    const af_dtype tk = key.info.type;
    if (!(tk == s32 || tk == u32 || tk == s64 || tk == u64)) AFB_TYPE_ERROR(1, tk);
    OpLock lk(in.device());
    LaunchCtx c = ctx_for_active_device();
    const Array *vals = &in;
    Holder cast;
    if (to != in.info.type) {  // castArray<Ti>(in), scan.cpp:45-58
        cast.a = new_array(to, in.info.dims);
        if (!launch_cast(c, in.info.type, to, view_of(in), cast.a->data())) AFB_TYPE_ERROR(1, in.info.type);
        vals = cast.a;
    }
    Holder res;
    res.a = new_array(to, in.info.dims);
    if (!launch_scan_by_key(c, (Op)binop, to, tk, view_of(key), view_of(*vals), dim, inclusive, res.a->data(), carry,
                            carry_key, summary))
        AFB_TYPE_ERROR(1, in.info.type);
    return res.release();
}

// ---- where -------------------------------------------------------------------------------------------
// enqueue the compaction into a worst-case sized output; the u64 hit count lands at *d_count_out (device)
Array *where_launch(const Array &in, unsigned base, unsigned long long **d_count_out) {
    const af_dtype t = in.info.type;
    if (t == f16) AFB_TYPE_ERROR(1, t);
    const dim_t n = in.elements();
    if (n >= 0xffffffffLL) fail(AF_ERR_SIZE, "af_where: indices are 32-bit, the input has too many elements");
    const Array *src = &in;
    Holder tmp;
    if (!in.is_linear()) {
        tmp.a = make_dense_copy(in);
        src = tmp.a;
    }
    Holder big;
    dim_t cap[4] = {n, 1, 1, 1};
    big.a = new_array(u32, cap);
    OpLock lk(in.device());
    LaunchCtx c = ctx_for_active_device();
    // every launch owns one slot of the 64-entry ring: a later where() on this device cannot overwrite the count before
    // the caller's gather has consumed it (at most 63 launches may be outstanding)
    HostRet ret = c.scratch->next_host_ret();
    unsigned long long *d_count = (unsigned long long *)ret.dev;
    if (!launch_where(c, t, view_of(*src), (unsigned *)big.a->data(), d_count, base, nullptr)) AFB_TYPE_ERROR(1, t);
    *d_count_out = d_count;
    return big.release();
}

// trim a worst-case sized where() output to its `h_count` hits (exact-size copy when the result is sparse)
Array *where_finish(Array *big_raw, dim_t h_count) {
    Holder big;
    big.a = big_raw;
    const dim_t n = big.a->info.dims[0];
    const char *e = getenv("AFB200_WHERE_SHRINK_DIV");
    const dim_t div = e ? atoll(e) : 8;
    if (h_count == 0 || (div > 0 && h_count < n / div)) {
        dim_t d[4] = {h_count, 1, 1, 1};
        Holder exact;
        exact.a = new_array(u32, d);
        if (h_count) {
            cudaStream_t s = stream_of(big.a->device());
            AFB_CUDA(cudaMemcpyAsync(exact.a->data(), big.a->data(), (size_t)h_count * 4, cudaMemcpyDeviceToDevice, s));
        }
        return exact.release();
    }
    big.a->info.dims[0] = h_count;
    dense_strides(big.a->info.dims, big.a->info.strides);
    return big.release();
}

// enqueue the compaction into a worst-case sized output (straight from the pool: the kernel writes indices to their final
// position); the kernel's last tile publishes {count, offset of the shard's slice, global count} through `ret`
Array *where_begin(const Array &in, unsigned base, const ShardLink *link, HostRet *ret) {
    const af_dtype t = in.info.type;
    if (t == f16) AFB_TYPE_ERROR(1, t);
    const dim_t n = in.elements();
    if (n >= 0xffffffffLL) fail(AF_ERR_SIZE, "af_where: indices are 32-bit, the input has too many elements");
    const Array *src = &in;
    Holder tmp;
    if (!in.is_linear()) {
        tmp.a = make_dense_copy(in);
        src = tmp.a;
    }
    Holder big;
    dim_t cap[4] = {n, 1, 1, 1};
    big.a = new_array(u32, cap);
    OpLock lk(in.device());
    LaunchCtx c = ctx_for_active_device();
    *ret = c.scratch->next_host_ret();
    if (!launch_where(c, t, view_of(*src), (unsigned *)big.a->data(), (unsigned long long *)ret->dev, base, ret, link))
        AFB_TYPE_ERROR(1, t);
    return big.release();
}

// the count comes back once, at the end (the reference syncs mid-pipeline to size the output)
Array *where_impl(const Array &in, unsigned base, dim_t *count_out, const ShardLink *link = nullptr, dim_t *offset_out = nullptr,
                  dim_t *total_out = nullptr) {
    HostRet ret;
    Holder big;
    big.a = where_begin(in, base, link, &ret);
    LaunchCtx c = ctx_for_active_device();
    wait_host_ret(c, ret);
    const unsigned long long *rec = (const unsigned long long *)ret.host;
    const dim_t h_count = (dim_t)rec[0];
    if (offset_out) *offset_out = (dim_t)rec[1];
    if (total_out) *total_out = (dim_t)rec[2];
    if (count_out) *count_out = h_count;
    return where_finish(big.release(), h_count);   // few hits: an exact-size buffer instead of the worst-case allocation
}

}  // namespace

extern "C" {

af_err af_sum(af_array *out, const af_array in, const int dim) { return reduce_dimwise(out, in, dim, FAM_PROMOTE, OP_ADD, false, 0); }
af_err af_sum_nan(af_array *out, const af_array in, const int dim, const double nanval) { return reduce_dimwise(out, in, dim, FAM_PROMOTE, OP_ADD, true, nanval); }
af_err af_product(af_array *out, const af_array in, const int dim) { return reduce_dimwise(out, in, dim, FAM_PROMOTE, OP_MUL, false, 0); }
af_err af_product_nan(af_array *out, const af_array in, const int dim, const double nanval) { return reduce_dimwise(out, in, dim, FAM_PROMOTE, OP_MUL, true, nanval); }
af_err af_min(af_array *out, const af_array in, const int dim) { return reduce_dimwise(out, in, dim, FAM_COMMON, OP_MIN, false, 0); }
af_err af_max(af_array *out, const af_array in, const int dim) { return reduce_dimwise(out, in, dim, FAM_COMMON, OP_MAX, false, 0); }
af_err af_count(af_array *out, const af_array in, const int dim) { return reduce_dimwise(out, in, dim, FAM_TYPE, OP_NOTZERO, false, 0); }
af_err af_all_true(af_array *out, const af_array in, const int dim) { return reduce_dimwise(out, in, dim, FAM_TYPE, OP_AND, false, 0); }
af_err af_any_true(af_array *out, const af_array in, const int dim) { return reduce_dimwise(out, in, dim, FAM_TYPE, OP_OR, false, 0); }

af_err af_sum_all(double *r, double *i, const af_array in) { return reduce_all_scalar(r, i, in, FAM_PROMOTE, OP_ADD, false, 0, false); }
af_err af_sum_nan_all(double *r, double *i, const af_array in, const double nanval) { return reduce_all_scalar(r, i, in, FAM_PROMOTE, OP_ADD, true, nanval, false); }
af_err af_product_all(double *r, double *i, const af_array in) { return reduce_all_scalar(r, i, in, FAM_PROMOTE, OP_MUL, false, 0, false); }
af_err af_product_nan_all(double *r, double *i, const af_array in, const double nanval) { return reduce_all_scalar(r, i, in, FAM_PROMOTE, OP_MUL, true, nanval, false); }
af_err af_min_all(double *r, double *i, const af_array in) { return reduce_all_scalar(r, i, in, FAM_COMMON, OP_MIN, false, 0, true); }
af_err af_max_all(double *r, double *i, const af_array in) { return reduce_all_scalar(r, i, in, FAM_COMMON, OP_MAX, false, 0, true); }
af_err af_count_all(double *r, double *i, const af_array in) { return reduce_all_scalar(r, i, in, FAM_TYPE, OP_NOTZERO, false, 0, false); }
af_err af_all_true_all(double *r, double *i, const af_array in) { return reduce_all_scalar(r, i, in, FAM_TYPE, OP_AND, false, 0, false); }
af_err af_any_true_all(double *r, double *i, const af_array in) { return reduce_all_scalar(r, i, in, FAM_TYPE, OP_OR, false, 0, false); }

af_err af_sum_all_array(af_array *out, const af_array in) { return reduce_all_array(out, in, FAM_PROMOTE, OP_ADD, false, 0, false); }
af_err af_sum_nan_all_array(af_array *out, const af_array in, const double nanval) { return reduce_all_array(out, in, FAM_PROMOTE, OP_ADD, true, nanval, false); }
af_err af_product_all_array(af_array *out, const af_array in) { return reduce_all_array(out, in, FAM_PROMOTE, OP_MUL, false, 0, false); }
af_err af_product_nan_all_array(af_array *out, const af_array in, const double nanval) { return reduce_all_array(out, in, FAM_PROMOTE, OP_MUL, true, nanval, false); }
af_err af_min_all_array(af_array *out, const af_array in) { return reduce_all_array(out, in, FAM_COMMON, OP_MIN, false, 0, true); }
af_err af_max_all_array(af_array *out, const af_array in) { return reduce_all_array(out, in, FAM_COMMON, OP_MAX, false, 0, true); }
af_err af_count_all_array(af_array *out, const af_array in) { return reduce_all_array(out, in, FAM_TYPE, OP_NOTZERO, false, 0, false); }
af_err af_all_true_all_array(af_array *out, const af_array in) { return reduce_all_array(out, in, FAM_TYPE, OP_AND, false, 0, false); }
af_err af_any_true_all_array(af_array *out, const af_array in) { return reduce_all_array(out, in, FAM_TYPE, OP_OR, false, 0, false); }

af_err af_imin(af_array *out, af_array *idx, const af_array in, const int dim) { return ireduce_dimwise(out, idx, in, dim, OP_MIN); }
af_err af_imax(af_array *out, af_array *idx, const af_array in, const int dim) { return ireduce_dimwise(out, idx, in, dim, OP_MAX); }
af_err af_imin_all(double *r, double *i, unsigned *idx, const af_array in) { return ireduce_all_scalar(r, i, idx, in, OP_MIN); }
af_err af_imax_all(double *r, double *i, unsigned *idx, const af_array in) { return ireduce_all_scalar(r, i, idx, in, OP_MAX); }

af_err af_accum(af_array *out, const af_array in, const int dim) { return scan_common(out, in, dim, AF_BINARY_ADD, true, true); }
af_err af_scan(af_array *out, const af_array in, const int dim, af_binary_op op, bool inclusive_scan) {
    return scan_common(out, in, dim, (int)op, inclusive_scan, false);
}

af_err af_scan_by_key(af_array *out, const af_array key_h, const af_array in_h, const int dim, af_binary_op op,
                      bool inclusive_scan) {
    AFB_API_BEGIN
    AFB_ARG(2, dim >= 0);
    AFB_ARG(2, dim < 4);
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    Array *key = from_handle(key_h);
    if (dim >= in->ndims()) {
        *out = retain(*in);
        return AF_SUCCESS;
    }
    check_same_device(*in);
    check_same_device(*key);
    *out = scan_by_key_impl(*key, *in, dim, (int)op, inclusive_scan, nullptr, nullptr);
    AFB_API_END
}

af_err af_where(af_array *idx, const af_array in_h) {
    AFB_API_BEGIN
    AFB_ARG(0, idx != nullptr);
    Array *in = from_handle(in_h);
    if (in->ndims() == 0) {  // where.cpp:39-41
        dim_t d[4] = {0, 1, 1, 1};
        *idx = new_array(u32, d);
        return AF_SUCCESS;
    }
    check_same_device(*in);
    *idx = where_impl(*in, 0u, nullptr);
    AFB_API_END
}

// ===================================================================================================
// afx_*: shard-local pieces of the sharded whole-array ops (see include/afb200.h)
// ===================================================================================================
af_err afx_reduce_all_partial(void *partial, af_dtype *partial_type, const af_array in_h, int op, bool change_nan,
                              double nanval) {
    AFB_API_BEGIN
    AFB_ARG(0, partial != nullptr);
    AFB_ARG(3, op >= 0 && op <= 6);
    Array *in = from_handle(in_h);
    AFB_ARG(2, in->elements() > 0);
    check_same_device(*in);
    const Family fam = (op == OP_ADD || op == OP_MUL) ? FAM_PROMOTE : (op == OP_MIN || op == OP_MAX) ? FAM_COMMON : FAM_TYPE;
    Op eff;
    af_dtype to;
    resolve(fam, (Op)op, in->info.type, &eff, &to);
    if (partial_type) *partial_type = to;
    reduce_all_into(*in, eff, to, change_nan, nanval, partial);
    AFB_API_END
}

af_err afx_ireduce_all_partial(void *partial_val, unsigned *partial_idx, const af_array in_h, int op) {
    AFB_API_BEGIN
    AFB_ARG(0, partial_val != nullptr);
    AFB_ARG(1, partial_idx != nullptr);
    AFB_ARG(3, op == OP_MIN || op == OP_MAX);
    Array *in = from_handle(in_h);
    AFB_ARG(2, in->ndims() > 0);
    check_same_device(*in);
    ireduce_all_into(*in, (Op)op, partial_val, partial_idx);
    AFB_API_END
}

af_err afx_scan_shard(af_array *out, const af_array key_h, const af_array in_h, af_binary_op op, bool inclusive_scan,
                      const void *carry, const void *carry_key) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    AFB_ARG(2, in->elements() > 0 && in->elements() == in->info.dims[0]);  // a 1-D shard
    check_same_device(*in);
    if (key_h) {
        Array *key = from_handle(key_h);
        check_same_device(*key);
        *out = scan_by_key_impl(*key, *in, 0, (int)op, inclusive_scan, carry_key ? carry : nullptr, carry_key);
    } else {
        af_dtype to;
        if (!scan_promote(in->info.type, &to)) AFB_TYPE_ERROR(1, in->info.type);
        AFB_ARG(3, (int)op >= 0 && (int)op <= 3);
        Holder res;
        res.a = new_array(to, in->info.dims);
        {
            OpLock lk(in->device());
            LaunchCtx c = ctx_for_active_device();
            if (!launch_scan(c, (Op)op, in->info.type, to, view_of(*in), 0, inclusive_scan, res.a->data(), carry))
                AFB_TYPE_ERROR(1, in->info.type);
        }
        *out = res.release();
    }
    AFB_API_END
}

af_err afx_scan_shard_local(af_array *out, void *summary, const af_array key_h, const af_array in_h, af_binary_op op,
                            bool inclusive_scan) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    AFB_ARG(1, summary != nullptr);
    AFB_ARG(2, key_h != nullptr);
    Array *in = from_handle(in_h);
    Array *key = from_handle(key_h);
    AFB_ARG(3, in->elements() > 0 && in->elements() == in->info.dims[0]);  // a 1-D shard
    check_same_device(*in);
    check_same_device(*key);
    *out = scan_by_key_impl(*key, *in, 0, (int)op, inclusive_scan, nullptr, nullptr, summary);
    AFB_API_END
}

af_err afx_scan_shard_fixup(af_array out_h, const void *carry, dim_t count, af_binary_op op) {
    AFB_API_BEGIN
    AFB_ARG(1, carry != nullptr);
    AFB_ARG(3, (int)op >= 0 && (int)op <= 3);
    Array *out = from_handle(out_h);
    AFB_ARG(2, count >= 0 && count <= out->elements());
    AFB_ARG(0, out->is_linear());
    check_same_device(*out);
    OpLock lk(out->device());
    LaunchCtx c = ctx_for_active_device();
    if (!launch_scan_fixup(c, (Op)op, out->info.type, out->data(), count, carry)) AFB_TYPE_ERROR(0, out->info.type);
    AFB_API_END
}

af_err afx_where_shard(af_array *idx, dim_t *count, const af_array in_h, unsigned base) {
    AFB_API_BEGIN
    AFB_ARG(0, idx != nullptr);
    Array *in = from_handle(in_h);
    AFB_ARG(2, in->elements() > 0);
    check_same_device(*in);
    *idx = where_impl(*in, base, count);
    AFB_API_END
}

af_err afx_where_shard_launch(af_array *idx, void **count_dev, const af_array in_h, unsigned base) {
    AFB_API_BEGIN
    AFB_ARG(0, idx != nullptr);
    AFB_ARG(1, count_dev != nullptr);
    Array *in = from_handle(in_h);
    AFB_ARG(2, in->elements() > 0);
    check_same_device(*in);
    unsigned long long *d = nullptr;
    *idx = where_launch(*in, base, &d);
    *count_dev = d;
    AFB_API_END
}

af_err afx_where_shard_finish(af_array *idx, dim_t count) {
    AFB_API_BEGIN
    AFB_ARG(0, idx != nullptr && *idx != nullptr);
    Array *big = from_handle(*idx);
    AFB_ARG(1, count >= 0 && count <= big->info.dims[0]);
    check_same_device(*big);
    *idx = where_finish(big, count);
    AFB_API_END
}

af_err afx_peer_exchange(void *out, const void *payload, const unsigned long long *peer_bufs, int rank, int world,
                         unsigned epoch) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    AFB_ARG(1, payload != nullptr);
    AFB_ARG(2, peer_bufs != nullptr);
    AFB_ARG(4, world >= 1 && world <= 8 && rank >= 0 && rank < world);
    AFB_ARG(5, epoch != 0);
    const int dev = active_device();
    OpLock lk(dev);
    LaunchCtx c = ctx_for_active_device();
    // pinned landing zone: [0, 512) gathered records, [2048, 2052) arrival flag
    unsigned char *hp = (unsigned char *)c.scratch->host_pinned;
    volatile unsigned *flag = (volatile unsigned *)(hp + 2048);
    *flag = 0;   // the kernel raises it to `epoch`; never trust what an earlier exchange (or cudaMallocHost) left there
    launch_peer_exchange(c, payload, peer_bufs, rank, world, epoch, hp, (unsigned *)(hp + 2048));
    // spin on the host flag instead of a stream synchronisation: the kernel raises it after its last store
    unsigned long long spins = 0;
    while (*flag != epoch) {
        if ((++spins & 0xfffff) == 0) {
            cudaError_t e = cudaStreamQuery(c.stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) AFB_CUDA(e);
            if (e == cudaSuccess && *flag != epoch) fail(AF_ERR_INTERNAL, "afx_peer_exchange: the exchange kernel ended without raising its flag");
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    memcpy(out, hp, (size_t)world * 64);
    AFB_API_END
}

// ===================================================================================================
// afx_sharded_*: whole-array ops over an array sharded across GPUs, combine fused into the kernels
// (kernels/peer_comm.cuh).  Every rank calls the same op in the same order with its own shard.
// ===================================================================================================
}  // extern "C"

namespace {
constexpr unsigned kCommMagic = 0xAFC0111Eu;
struct Comm {
    unsigned magic = kCommMagic;
    ShardLink link;
    std::mutex mu;
};
Comm *comm_of(void *h) {
    Comm *c = reinterpret_cast<Comm *>(h);
    if (!c || c->magic != kCommMagic) fail(AF_ERR_ARG, "invalid afx_comm handle");
    return c;
}
// the link of the next collective op of this communicator (epochs advance in lockstep on every rank)
ShardLink next_link(Comm *c, unsigned idx_base, int mode) {
    std::lock_guard<std::mutex> lk(c->mu);
    if (++c->link.epoch == 0) c->link.epoch = 2;   // never 0, and the parity keeps alternating across the wrap
    ShardLink l = c->link;
    l.idx_base = idx_base;
    l.mode = mode;
    return l;
}

// an op that has been enqueued and whose O(1) host result has not been collected yet (afx_*_async)
constexpr unsigned kPendingMagic = 0xAFB2F07Eu;
enum PendingKind { PEND_SCALAR = 0, PEND_INDEXED = 1, PEND_WHERE = 2 };
struct Pending {
    unsigned magic = kPendingMagic;
    int kind = PEND_SCALAR;
    int device = 0;
    af_dtype type = f32;     // dtype of the value in the record
    HostRet ret{};
    unsigned idx_add = 0;
    Array *big = nullptr;    // where: the worst-case sized output
    ~Pending() {
        magic = 0;
        delete big;
    }
};
LaunchCtx ctx_for_device(int d) { return LaunchCtx{d, stream_of(d), sm_count(d), &scratch_of(d)}; }
Pending *pending_of(void *h) {
    Pending *p = reinterpret_cast<Pending *>(h);
    if (!p || p->magic != kPendingMagic) fail(AF_ERR_ARG, "invalid afx_pending handle");
    return p;
}
}  // namespace

extern "C" {

af_err afx_comm_create(void **comm, int rank, int world, const unsigned long long *peer_bufs) {
    AFB_API_BEGIN
    AFB_ARG(0, comm != nullptr);
    AFB_ARG(2, world >= 1 && world <= 8);
    AFB_ARG(1, rank >= 0 && rank < world);
    AFB_ARG(3, peer_bufs != nullptr || world == 1);
    Comm *c = new Comm();
    for (int r = 0; r < 8; r++) c->link.bufs[r] = (r < world && peer_bufs) ? peer_bufs[r] : 0ULL;
    c->link.rank = rank;
    c->link.world = world;
    c->link.epoch = 0;
    c->link.idx_base = 0;
    c->link.mode = 0;
    *comm = c;
    AFB_API_END
}

af_err afx_comm_release(void *comm) {
    AFB_API_BEGIN
    if (!comm) return AF_SUCCESS;
    Comm *c = comm_of(comm);
    c->magic = 0;
    delete c;
    AFB_API_END
}

af_err afx_reduce_all_async(afx_pending *pending, void *comm, const af_array in_h, int op, bool change_nan, double nanval) {
    AFB_API_BEGIN
    AFB_ARG(0, pending != nullptr);
    AFB_ARG(3, op >= 0 && op <= 6);
    Comm *cm = comm ? comm_of(comm) : nullptr;
    Array *in = from_handle(in_h);
    AFB_ARG(2, in->elements() > 0);
    check_same_device(*in);
    const Family fam = (op == OP_ADD || op == OP_MUL) ? FAM_PROMOTE : (op == OP_MIN || op == OP_MAX) ? FAM_COMMON : FAM_TYPE;
    Op eff;
    af_dtype to;
    resolve(fam, (Op)op, in->info.type, &eff, &to);
    std::unique_ptr<Pending> p(new Pending());
    p->kind = PEND_SCALAR;
    p->device = in->device();
    p->type = to;
    ShardLink link;
    if (cm) link = next_link(cm, 0u, 0);
    reduce_all_into(*in, eff, to, change_nan, nanval, nullptr, &p->ret, cm ? &link : nullptr);
    *pending = p.release();
    AFB_API_END
}

af_err afx_ireduce_all_async(afx_pending *pending, void *comm, const af_array in_h, int op, unsigned shard_base) {
    AFB_API_BEGIN
    AFB_ARG(0, pending != nullptr);
    AFB_ARG(3, op == OP_MIN || op == OP_MAX);
    Comm *cm = comm ? comm_of(comm) : nullptr;
    Array *in = from_handle(in_h);
    AFB_ARG(2, in->ndims() > 0);
    const af_dtype t = in->info.type;
    if (t == f16) AFB_TYPE_ERROR(2, t);
    check_same_device(*in);
    std::unique_ptr<Pending> p(new Pending());
    p->kind = PEND_INDEXED;
    p->device = in->device();
    p->type = t;
    ShardLink link;
    if (cm) link = next_link(cm, shard_base, 0);
    ireduce_all_into(*in, (Op)op, nullptr, nullptr, &p->ret, cm ? &link : nullptr);
    p->idx_add = cm ? 0u : shard_base;   // a lone shard: the kernel did not add the base
    *pending = p.release();
    AFB_API_END
}

af_err afx_where_async(afx_pending *pending, void *comm, const af_array in_h, unsigned shard_base) {
    AFB_API_BEGIN
    AFB_ARG(0, pending != nullptr);
    Comm *cm = comm ? comm_of(comm) : nullptr;
    Array *in = from_handle(in_h);
    AFB_ARG(2, in->elements() > 0);
    check_same_device(*in);
    std::unique_ptr<Pending> p(new Pending());
    p->kind = PEND_WHERE;
    p->device = in->device();
    p->type = u32;
    ShardLink link;
    if (cm) link = next_link(cm, 0u, 0);
    p->big = where_begin(*in, shard_base, cm ? &link : nullptr, &p->ret);
    *pending = p.release();
    AFB_API_END
}

af_err afx_pending_scalar(afx_pending pending, double *real, double *imag, unsigned *idx) {
    AFB_API_BEGIN
    AFB_ARG(1, real != nullptr);
    std::unique_ptr<Pending> p(pending_of(pending));   // consumed whatever happens below
    if (p->kind == PEND_WHERE) fail(AF_ERR_ARG, "afx_pending_scalar: the pending op is a where (use afx_pending_where)");
    if (p->kind == PEND_INDEXED) AFB_ARG(3, idx != nullptr);
    if ((p->type == c32 || p->type == c64)) AFB_ARG(2, imag != nullptr);
    *real = 0;
    if (imag) *imag = 0;
    LaunchCtx c = ctx_for_device(p->device);
    wait_host_ret(c, p->ret);
    decode_scalar(p->type, p->ret.host, real, imag);
    if (idx) *idx = p->kind == PEND_INDEXED ? *(const unsigned *)((const char *)p->ret.host + 32) + p->idx_add : 0u;
    AFB_API_END
}

af_err afx_pending_where(afx_pending pending, af_array *idx, dim_t *offset, dim_t *total) {
    AFB_API_BEGIN
    AFB_ARG(1, idx != nullptr);
    std::unique_ptr<Pending> p(pending_of(pending));
    if (p->kind != PEND_WHERE) fail(AF_ERR_ARG, "afx_pending_where: the pending op is not a where");
    LaunchCtx c = ctx_for_device(p->device);
    wait_host_ret(c, p->ret);
    const unsigned long long *rec = (const unsigned long long *)p->ret.host;
    if (offset) *offset = (dim_t)rec[1];
    if (total) *total = (dim_t)rec[2];
    Array *big = p->big;
    p->big = nullptr;
    *idx = where_finish(big, (dim_t)rec[0]);
    AFB_API_END
}

af_err afx_pending_release(afx_pending pending) {
    AFB_API_BEGIN
    if (!pending) return AF_SUCCESS;
    delete pending_of(pending);   // the kernel may still be running: its output buffer goes back to the stream-ordered pool
    AFB_API_END
}

// the blocking forms: issue + collect
af_err afx_sharded_reduce_all(double *real, double *imag, void *comm, const af_array in_h, int op, bool change_nan,
                              double nanval) {
    AFB_API_BEGIN
    AFB_ARG(0, real != nullptr);
    comm_of(comm);
    afx_pending p = nullptr;
    af_err e = afx_reduce_all_async(&p, comm, in_h, op, change_nan, nanval);
    if (e != AF_SUCCESS) return e;
    return afx_pending_scalar(p, real, imag, nullptr);
    AFB_API_END
}

af_err afx_sharded_ireduce_all(double *real, double *imag, unsigned *idx, void *comm, const af_array in_h, int op,
                               unsigned shard_base) {
    AFB_API_BEGIN
    AFB_ARG(0, real != nullptr);
    AFB_ARG(2, idx != nullptr);
    comm_of(comm);
    afx_pending p = nullptr;
    af_err e = afx_ireduce_all_async(&p, comm, in_h, op, shard_base);
    if (e != AF_SUCCESS) return e;
    return afx_pending_scalar(p, real, imag, idx);
    AFB_API_END
}

af_err afx_sharded_scan(af_array *out, void *comm, const af_array key_h, const af_array in_h, af_binary_op op,
                        bool inclusive_scan) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    AFB_ARG(4, (int)op >= 0 && (int)op <= 3);
    Comm *cm = comm_of(comm);
    Array *in = from_handle(in_h);
    AFB_ARG(3, in->elements() > 0 && in->elements() == in->info.dims[0]);   // a 1-D shard
    check_same_device(*in);
    if (key_h) {
        // ---- segmented: carry-less local scan (+ summary), then ONE kernel exchanges the summaries and repairs the
        // shard's first segment -- two launches, no host synchronisation
        Array *key = from_handle(key_h);
        check_same_device(*key);
        AFB_ARG(2, key->is_linear());
        const ShardLink link = next_link(cm, 0u, 0);
        HostRet slot;
        {
            OpLock lk(in->device());
            LaunchCtx c = ctx_for_active_device();
            slot = c.scratch->next_host_ret();   // only its 64-byte device half is used: the shard summary
        }
        Holder res;
        res.a = scan_by_key_impl(*key, *in, 0, (int)op, inclusive_scan, nullptr, nullptr, slot.dev);
        {
            OpLock lk(in->device());
            LaunchCtx c = ctx_for_active_device();
            if (!launch_sbk_exchange_fixup(c, (Op)op, res.a->info.type, key->info.type, res.a->data(), in->elements(), key->data(),
                                           slot.dev, link))
                AFB_TYPE_ERROR(3, in->info.type);
        }
        *out = res.release();
        return AF_SUCCESS;
    }
    // ---- plain: the shard totals are exchanged inside the reduction kernel, which leaves the fold of the LOWER ranks --
    // this shard's carry-in -- in device memory; the scan kernel's walkers read it from there.  Two launches, no host
    // synchronisation; the second read of the shard comes from L2 when the shard fits (126 MB).
    af_dtype to;
    if (!scan_promote(in->info.type, &to)) AFB_TYPE_ERROR(3, in->info.type);
    const af_dtype ti = in->info.type;
    const bool promoted = ti != to;
    if (ti == b8 || (promoted && ((int)op == OP_MIN || (int)op == OP_MAX)))
        fail(AF_ERR_NOT_SUPPORTED, "afx_sharded_scan: no sharded form for this (type, op) pair (b8, or min / max of a promoted type)");
    const ShardLink link = next_link(cm, 0u, 1);
    Holder res;
    res.a = new_array(to, in->info.dims);
    {
        OpLock lk(in->device());
        LaunchCtx c = ctx_for_active_device();
        HostRet slot = c.scratch->next_host_ret();
        if (!launch_reduce_all(c, (Op)op, ti, to, view_of(*in), false, 0.0, slot.dev, nullptr, &link)) AFB_TYPE_ERROR(3, ti);
        if (!launch_scan(c, (Op)op, ti, to, view_of(*in), 0, inclusive_scan, res.a->data(), slot.dev)) AFB_TYPE_ERROR(3, ti);
    }
    *out = res.release();
    AFB_API_END
}

af_err afx_sharded_where(af_array *idx, dim_t *offset, dim_t *total, void *comm, const af_array in_h, unsigned shard_base) {
    AFB_API_BEGIN
    AFB_ARG(0, idx != nullptr);
    comm_of(comm);
    afx_pending p = nullptr;
    af_err e = afx_where_async(&p, comm, in_h, shard_base);
    if (e != AF_SUCCESS) return e;
    return afx_pending_where(p, idx, offset, total);
    AFB_API_END
}

af_err afx_launch_count(unsigned long long *count, bool reset) {
    AFB_API_BEGIN
    AFB_ARG(0, count != nullptr);
    *count = launches_total(reset);
    AFB_API_END
}

}  // extern "C"
