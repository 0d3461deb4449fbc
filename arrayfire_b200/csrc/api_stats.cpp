// api_stats.cpp -- SURVEY.md section 8(f)-2: the statistics built on the reduction path.
//
// Entry points, argument checks, dtype tables and the (quirky) bias rules restate the reference's C layer:
//   /root/reference/src/api/c/mean.cpp:65-187      af_mean, af_mean_all            (weighted forms: not supported yet)
//   /root/reference/src/api/c/var.cpp:52-75,104-136,187-262,336-372,418-470   af_var(_v2), af_var_all(_v2), af_meanvar
//   /root/reference/src/api/c/stdev.cpp:44-160     af_stdev(_v2), af_stdev_all(_v2) (no complex case, as in the reference)
//   /root/reference/src/api/c/norm.cpp:45-151      af_norm
//   /root/reference/src/api/c/corrcoef.cpp:39-101  af_corrcoef;  covariance.cpp:41-109  af_cov(_v2)
// Every function is ONE streaming pass per input through this library's reduction kernels (kernels/stats_*.cu) -- the
// reference composes them from element-wise temporaries and separate reductions.  No CPU fallback.
#include <cmath>

#include "common.hpp"

using namespace afb;

namespace afb {
Array *make_dense_copy(const Array &a);
}

namespace {

struct OpLock {
    std::lock_guard<std::mutex> g;
    explicit OpLock(int device) : g(op_mutex(device)) {}
};
struct Holder {
    Array *a = nullptr;
    ~Holder() { delete a; }
    Array *release() { Array *r = a; a = nullptr; return r; }
};

// output dtype of mean / var / stdev along a dimension (mean.cpp:71-86)
bool stat_out_type(af_dtype ti, af_dtype *to) {
    switch (ti) {
        case f32: case s32: case u32: case s16: case u16: case s8: case u8: case b8: *to = f32; return true;
        case f64: case s64: case u64: *to = f64; return true;
        case c32: *to = c32; return true;
        case c64: *to = c64; return true;
        case f16: *to = f16; return true;
    }
    return false;
}
// dtype of the whole-array scalar as the kernels produce it (f16 computes in f32)
af_dtype stat_scalar_type(af_dtype ti) {
    af_dtype to = f32;
    stat_out_type(ti, &to);
    return to == f16 ? f32 : to;
}
void scalar_out(af_dtype t, const void *raw, double *real, double *imag) {
    switch (t) {
        case f32: *real = *(const float *)raw; break;
        case f64: *real = *(const double *)raw; break;
        case c32: *real = ((const float *)raw)[0]; if (imag) *imag = ((const float *)raw)[1]; break;
        case c64: *real = ((const double *)raw)[0]; if (imag) *imag = ((const double *)raw)[1]; break;
        default: break;
    }
}
// whole-array kernels stream linear views (or views with long columns); anything else is packed first
const Array *linear_src(const Array &in, Holder &tmp) {
    if (!in.is_linear() && in.info.dims[0] < 4096) {
        tmp.a = make_dense_copy(in);
        return tmp.a;
    }
    return &in;
}

enum { BIAS_DEFAULT = 0, BIAS_SAMPLE = 1, BIAS_POPULATION = 2 };

af_err moments_dim(af_array *var_out, af_array *mean_out, bool want_mean, const af_array in_h, int bias, dim_t dim,
                   bool take_sqrt, bool stdev_rule, int dim_arg_index) {
    AFB_API_BEGIN
    AFB_ARG(dim_arg_index, dim >= 0 && dim <= 3);
    AFB_ARG(want_mean ? 1 : 0, var_out != nullptr);
    if (want_mean) AFB_ARG(0, mean_out != nullptr);
    else mean_out = nullptr;
    Array *in = from_handle(in_h);
    const af_dtype ti = in->info.type;
    af_dtype to;
    if (!stat_out_type(ti, &to)) AFB_TYPE_ERROR(1, ti);
    if (take_sqrt && (ti == c32 || ti == c64 || ti == f16)) AFB_TYPE_ERROR(1, ti);   // stdev.cpp:139-160
    check_same_device(*in);
    dim_t od[4] = {in->info.dims[0], in->info.dims[1], in->info.dims[2], in->info.dims[3]};
    const dim_t n = od[dim];
    od[dim] = 1;
    // var.cpp:118-121: 1 / (population ? N : N - 1) -- the DEFAULT bias divides by N - 1 on this path;
    // stdev.cpp:80-81: N - (bias == SAMPLE)
    const double denom = stdev_rule ? (double)(n - (bias == BIAS_SAMPLE ? 1 : 0)) : (double)(bias == BIAS_POPULATION ? n : n - 1);
    Holder v, m;
    v.a = new_array(to, od);
    if (mean_out) m.a = new_array(to, od);
    if (in->elements() > 0) {
        OpLock lk(in->device());
        LaunchCtx c = ctx_for_active_device();
        if (!launch_var(c, ti, view_of(*in), (int)dim, v.a->data(), mean_out ? m.a->data() : nullptr, 1.0 / denom, take_sqrt))
            AFB_TYPE_ERROR(1, ti);
    }
    *var_out = v.release();
    if (mean_out) *mean_out = m.release();
    AFB_API_END
}

af_err moments_all(double *real, double *imag, const af_array in_h, int bias, bool take_sqrt) {
    AFB_API_BEGIN
    AFB_ARG(0, real != nullptr);
    Array *in = from_handle(in_h);
    const af_dtype ti = in->info.type;
    af_dtype to;
    if (!stat_out_type(ti, &to)) AFB_TYPE_ERROR(2, ti);
    if (take_sqrt && (ti == c32 || ti == c64 || ti == f16)) AFB_TYPE_ERROR(2, ti);   // stdev.cpp:101-125
    if (ti == c32 || ti == c64) AFB_ARG(1, imag != nullptr);
    *real = 0;
    if (imag) *imag = 0;
    AFB_ARG(2, in->elements() > 0);
    check_same_device(*in);
    const double denom = (double)(in->elements() - (bias == BIAS_SAMPLE ? 1 : 0));   // var.cpp:71-73, stdev.cpp:55-57
    Holder tmp;
    const Array *src = linear_src(*in, tmp);
    HostRet ret;
    LaunchCtx c;
    {
        OpLock lk(in->device());
        c = ctx_for_active_device();
        ret = c.scratch->next_host_ret();
        if (!launch_var_all(c, ti, view_of(*src), ret.dev, 1.0 / denom, take_sqrt, &ret)) AFB_TYPE_ERROR(2, ti);
    }
    wait_host_ret(c, ret);
    scalar_out(stat_scalar_type(ti), ret.host, real, imag);
    AFB_API_END
}

// whole-array reduction of the norm family into the device slot of a host record; returns the scalar
double norm_all(const Array &in, Op op, int pre, double p) {
    Holder tmp;
    const Array *src = linear_src(in, tmp);
    HostRet ret;
    LaunchCtx c;
    {
        OpLock lk(in.device());
        c = ctx_for_active_device();
        ret = c.scratch->next_host_ret();
        if (!launch_norm_reduce(c, op, in.info.type, view_of(*src), -1, pre, p, ret.dev, &ret)) AFB_TYPE_ERROR(1, in.info.type);
    }
    wait_host_ret(c, ret);
    const af_dtype t = in.info.type;
    return (t == f64 || t == c64) ? *(const double *)ret.host : (double)*(const float *)ret.host;
}

// sums of |x| (or |x|^p) along dim 0 / 1 as a real array
Array *norm_dim(const Array &in, int dim, int pre, double p) {
    const af_dtype t = in.info.type;
    const af_dtype rt = (t == f64 || t == c64) ? f64 : f32;
    dim_t od[4] = {in.info.dims[0], in.info.dims[1], in.info.dims[2], in.info.dims[3]};
    od[dim] = 1;
    Holder r;
    r.a = new_array(rt, od);
    OpLock lk(in.device());
    LaunchCtx c = ctx_for_active_device();
    if (!launch_norm_reduce(c, OP_ADD, t, view_of(in), dim, pre, p, r.a->data(), nullptr)) AFB_TYPE_ERROR(1, t);
    return r.release();
}

}  // namespace

extern "C" {

af_err af_mean(af_array *out, const af_array in_h, const dim_t dim) {
    AFB_API_BEGIN
    AFB_ARG(2, dim >= 0 && dim <= 3);
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    const af_dtype ti = in->info.type;
    af_dtype to;
    if (!stat_out_type(ti, &to)) AFB_TYPE_ERROR(1, ti);
    check_same_device(*in);
    dim_t od[4] = {in->info.dims[0], in->info.dims[1], in->info.dims[2], in->info.dims[3]};
    od[dim] = 1;
    Holder r;
    r.a = new_array(to, od);
    if (in->elements() > 0) {
        OpLock lk(in->device());
        LaunchCtx c = ctx_for_active_device();
        if (!launch_mean(c, ti, view_of(*in), (int)dim, r.a->data())) AFB_TYPE_ERROR(1, ti);
    }
    *out = r.release();
    AFB_API_END
}

af_err af_mean_all(double *real, double *imag, const af_array in_h) {
    AFB_API_BEGIN
    AFB_ARG(0, real != nullptr);
    Array *in = from_handle(in_h);
    const af_dtype ti = in->info.type;
    af_dtype to;
    if (!stat_out_type(ti, &to)) AFB_TYPE_ERROR(2, ti);
    if (ti == c32 || ti == c64) AFB_ARG(1, imag != nullptr);
    *real = 0;
    if (imag) *imag = 0;
    AFB_ARG(2, in->elements() > 0);
    check_same_device(*in);
    Holder tmp;
    const Array *src = linear_src(*in, tmp);
    HostRet ret;
    LaunchCtx c;
    {
        OpLock lk(in->device());
        c = ctx_for_active_device();
        ret = c.scratch->next_host_ret();
        if (!launch_mean_all(c, ti, view_of(*src), ret.dev, &ret)) AFB_TYPE_ERROR(2, ti);
    }
    wait_host_ret(c, ret);
    scalar_out(stat_scalar_type(ti), ret.host, real, imag);
    AFB_API_END
}

af_err af_var_v2(af_array *out, const af_array in, const af_var_bias bias, const dim_t dim) {
    return moments_dim(out, nullptr, false, in, (int)bias, dim, false, false, 3);
}
af_err af_var(af_array *out, const af_array in, const bool isbiased, const dim_t dim) {
    return af_var_v2(out, in, isbiased ? AF_VARIANCE_SAMPLE : AF_VARIANCE_POPULATION, dim);   // var.cpp:187-192
}
static af_err meanvar_weighted(af_array *mean, af_array *var, const af_array in, const af_array weights, int bias, dim_t dim);

af_err af_meanvar(af_array *mean, af_array *var, const af_array in, const af_array weights, const af_var_bias bias,
                  const dim_t dim) {
    if (weights != nullptr) {   // an EMPTY weights array means "unweighted" (var.cpp:113: weights.isEmpty()); af::array() is one
        bool weighted = false;
        {
            AFB_API_BEGIN
            Array *w = from_handle(weights);
            AFB_ARG(3, w->info.type == f32 || w->info.type == f64);                              // var.cpp:424-426
            weighted = w->elements() > 0;
            }
            catch (...) { return ::afb::current_exception_to_err(); }
        }
        if (weighted) {
            {
                AFB_API_BEGIN
                AFB_ARG(5, dim >= 0 && dim <= 3);
                AFB_ARG(0, mean != nullptr);
                AFB_ARG(1, var != nullptr);
                }
                catch (...) { return ::afb::current_exception_to_err(); }
            }
            return meanvar_weighted(mean, var, in, weights, (int)bias, dim);
        }
    }
    return moments_dim(var, mean, true, in, (int)bias, dim, false, false, 5);
}
af_err af_var_all_v2(double *real, double *imag, const af_array in, const af_var_bias bias) {
    return moments_all(real, imag, in, (int)bias, false);
}
af_err af_var_all(double *real, double *imag, const af_array in, const bool isbiased) {
    return af_var_all_v2(real, imag, in, isbiased ? AF_VARIANCE_SAMPLE : AF_VARIANCE_POPULATION);
}
af_err af_stdev_v2(af_array *out, const af_array in, const af_var_bias bias, const dim_t dim) {
    return moments_dim(out, nullptr, false, in, (int)bias, dim, true, true, 2);
}
af_err af_stdev(af_array *out, const af_array in, const dim_t dim) { return af_stdev_v2(out, in, AF_VARIANCE_POPULATION, dim); }
af_err af_stdev_all_v2(double *real, double *imag, const af_array in, const af_var_bias bias) {
    return moments_all(real, imag, in, (int)bias, true);
}
af_err af_stdev_all(double *real, double *imag, const af_array in) { return af_stdev_all_v2(real, imag, in, AF_VARIANCE_POPULATION); }

af_err af_norm(double *out, const af_array in_h, const af_norm_type type, const double p, const double q) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    if (in->ndims() > 2) fail(AF_ERR_BATCH, "af_norm: can not be used in batch mode");    // norm.cpp:123-125
    const af_dtype t = in->info.type;
    AFB_ARG(1, t == f32 || t == f64 || t == c32 || t == c64 || t == f16);                  // isFloating()
    *out = 0;
    if (in->ndims() == 0) return AF_SUCCESS;
    check_same_device(*in);
    const af_dtype rt = (t == f64 || t == c64) ? f64 : f32;
    auto vector_norm = [&](double pp) -> double {
        if (pp == 1) return norm_all(*in, OP_ADD, 0, 0.0);
        if (std::isinf(pp)) return norm_all(*in, OP_MAX, 0, 0.0);
        if (pp == 2) return std::sqrt(norm_all(*in, OP_ADD, 1, 0.0));
        return std::pow(norm_all(*in, OP_ADD, 2, pp), 1.0 / pp);
    };
    auto max_of = [&](Array *sums_raw) -> double {        // reduce_all<af_max_t> of a real array
        Holder sums;
        sums.a = sums_raw;
        HostRet ret;
        LaunchCtx c;
        {
            OpLock lk(in->device());
            c = ctx_for_active_device();
            ret = c.scratch->next_host_ret();
            if (!launch_reduce_all(c, OP_MAX, rt, rt, view_of(*sums.a), false, 0.0, ret.dev, &ret)) AFB_TYPE_ERROR(1, rt);
        }
        wait_host_ret(c, ret);
        return rt == f64 ? *(const double *)ret.host : (double)*(const float *)ret.host;
    };
    switch (type) {
        case AF_NORM_VECTOR_2: *out = vector_norm(2); break;
        case AF_NORM_VECTOR_1: *out = vector_norm(1); break;
        case AF_NORM_VECTOR_INF: *out = vector_norm(INFINITY); break;
        case AF_NORM_VECTOR_P: *out = vector_norm(p); break;
        case AF_NORM_MATRIX_1: *out = max_of(norm_dim(*in, 0, 0, 0.0)); break;
        case AF_NORM_MATRIX_INF: *out = max_of(norm_dim(*in, 1, 0, 0.0)); break;
        case AF_NORM_MATRIX_L_PQ: {
            // norm.cpp:80-105: column p-norms, then the q-norm of those
            Holder cols;
            cols.a = p == 1 ? norm_dim(*in, 0, 0, 0.0) : norm_dim(*in, 0, 2, p);
            const double inner_root = p == 1 ? 1.0 : 1.0 / p;
            // sum over columns of (colsum^(1/p))^q, fused into ONE reduction: |x|^(q/p) of the column sums
            if (q == 1 && p == 1) *out = norm_all(*cols.a, OP_ADD, 0, 0.0);
            else if (q == 1) *out = norm_all(*cols.a, OP_ADD, 2, inner_root);
            else *out = std::pow(norm_all(*cols.a, OP_ADD, 2, inner_root * q), 1.0 / q);
        } break;
        default: fail(AF_ERR_NOT_SUPPORTED, "af_norm: this type of norm is not supported");   // incl. AF_NORM_MATRIX_2
    }
    AFB_API_END
}

af_err af_corrcoef(double *real, double *imag, const af_array x_h, const af_array y_h) {
    (void)imag;
    AFB_API_BEGIN
    AFB_ARG(0, real != nullptr);
    Array *x = from_handle(x_h), *y = from_handle(y_h);
    const af_dtype t = x->info.type;
    AFB_ARG(2, t == y->info.type);
    AFB_ARG(2, x->ndims() == y->ndims());
    for (int i = 0; i < x->ndims(); i++) AFB_ARG(2, x->info.dims[i] == y->info.dims[i]);
    if (t == c32 || t == c64 || t == f16) AFB_TYPE_ERROR(1, t);      // corrcoef.cpp:84-97
    AFB_ARG(1, x->elements() > 0);
    check_same_device(*x);
    check_same_device(*y);
    Holder tx, ty;
    const Array *sx = x, *sy = y;
    if (!x->is_linear()) { tx.a = make_dense_copy(*x); sx = tx.a; }
    if (!y->is_linear()) { ty.a = make_dense_copy(*y); sy = ty.a; }
    const bool dbl = (t == f64 || t == s64 || t == u64);
    Holder sums;
    dim_t five[4] = {5, 1, 1, 1};
    sums.a = new_array(dbl ? f64 : f32, five);
    double s[5];
    {
        OpLock lk(x->device());
        LaunchCtx c = ctx_for_active_device();
        if (!launch_corr_sums(c, t, sx->data(), sy->data(), x->elements(), sums.a->data())) AFB_TYPE_ERROR(1, t);
        if (dbl) {
            AFB_CUDA(cudaMemcpyAsync(s, sums.a->data(), 5 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
            AFB_CUDA(cudaStreamSynchronize(c.stream));
        } else {
            float f[5];
            AFB_CUDA(cudaMemcpyAsync(f, sums.a->data(), 5 * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
            AFB_CUDA(cudaStreamSynchronize(c.stream));
            for (int k = 0; k < 5; k++) s[k] = f[k];
        }
    }
    const double n = (double)x->elements();
    // corrcoef.cpp:55-58 (evaluated in the compute type there; the final formula in double here)
    *real = (n * s[4] - s[0] * s[1]) / (std::sqrt(n * s[2] - s[0] * s[0]) * std::sqrt(n * s[3] - s[1] * s[1]));
    if (!dbl) *real = (double)(float)*real;
    AFB_API_END
}

af_err af_cov_v2(af_array *out, const af_array x_h, const af_array y_h, const af_var_bias bias) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    Array *x = from_handle(x_h), *y = from_handle(y_h);
    const af_dtype t = x->info.type;
    AFB_ARG(1, x->ndims() <= 2);
    AFB_ARG(2, x->ndims() == y->ndims());
    AFB_ARG(2, x->info.dims[0] == y->info.dims[0]);
    AFB_ARG(2, x->info.dims[1] == y->info.dims[1]);
    AFB_ARG(2, t == y->info.type);
    if (t == c32 || t == c64 || t == f16 || t == b8) AFB_TYPE_ERROR(1, t);    // covariance.cpp:92-104
    AFB_ARG(1, x->elements() > 0);
    check_same_device(*x);
    check_same_device(*y);
    Holder tx, ty;
    const Array *sx = x, *sy = y;
    if (!x->is_linear()) { tx.a = make_dense_copy(*x); sx = tx.a; }
    if (!y->is_linear()) { ty.a = make_dense_copy(*y); sy = ty.a; }
    const bool dbl = (t == f64 || t == s64 || t == u64);
    const af_dtype tc = dbl ? f64 : f32;
    const dim_t n = x->info.dims[0], ncols = x->info.dims[1];
    const double denom = (double)(bias == AF_VARIANCE_SAMPLE ? n - 1 : n);    // covariance.cpp:49
    dim_t od[4] = {1, ncols, 1, 1}, two[4] = {2, 1, 1, 1};
    Holder res, means;
    res.a = new_array(tc, od);
    means.a = new_array(tc, two);
    {
        OpLock lk(x->device());
        LaunchCtx c = ctx_for_active_device();
        const size_t esz = dbl ? 8 : 4;
        // the means of the WHOLE arrays (covariance.cpp:51-54), left on the device
        if (!launch_mean_all(c, t, view_of(*sx), means.a->data(), nullptr)) AFB_TYPE_ERROR(1, t);
        if (!launch_mean_all(c, t, view_of(*sy), (char *)means.a->data() + esz, nullptr)) AFB_TYPE_ERROR(1, t);
        if (!launch_cov_cols(c, t, sx->data(), sy->data(), n, ncols, means.a->data(), (char *)means.a->data() + esz,
                             1.0 / denom, res.a->data()))
            AFB_TYPE_ERROR(1, t);
    }
    *out = res.release();
    AFB_API_END
}
af_err af_cov(af_array *out, const af_array x, const af_array y, const bool isbiased) {
    return af_cov_v2(out, x, y, isbiased ? AF_VARIANCE_SAMPLE : AF_VARIANCE_POPULATION);
}

// ---- weighted forms, composed from this library's own lazy element-wise expressions (api_expr.cpp) and reductions:
//   mean_w     = sum(w * x) / sum(w)                      (mean.cpp:97-149, cpu/kernel/mean.hpp:41-82 computes a running form)
//   var_all_w  = sum(w * (x - mean_w)^2) / sum(w)         (var.cpp:77-102) -- the whole-array forms run FUSED: x and w are
//                streamed once per reduction and no temporary is written
// Real inputs only (integers are cast to the f32 / f64 the reference computes in); the dim-wise weighted variance needs a
// broadcast of the means along `dim`, which this library has no kernel for yet.
namespace {
struct H {           // releases af_array temporaries
    af_array h = nullptr;
    ~H() { if (h) af_release_array(h); }
};
#define AFB_TRY(call)                        \
    do {                                     \
        af_err e_ = (call);                  \
        if (e_ != AF_SUCCESS) return e_;     \
    } while (0)

// x as the compute type (f32, or f64 for f64 / s64 / u64) and w as the same type; complex / f16 are not handled.
// `wt` receives the weights tiled up to the dims of the input when `may_tile` (af_mean_weighted, mean.cpp:112-124: a
// weights dim of 1 is repeated); the other weighted forms of the reference do not tile, nor do they here.
af_err weighted_operands(H &wt, const af_array in, const af_array weights, int arg_w, bool may_tile = false) {
    unsigned rep[4] = {1, 1, 1, 1};
    bool tile = false;
    {
        AFB_API_BEGIN
        Array *a = from_handle(in), *ww = from_handle(weights);
        AFB_ARG(arg_w, ww->info.type == f32 || ww->info.type == f64);            // mean.cpp:105-110
        const af_dtype t = a->info.type;
        if (t == c32 || t == c64 || t == f16) fail(AF_ERR_NOT_SUPPORTED, "weighted statistics: complex and f16 inputs are not supported by libafb200");
        for (int i = 0; i < 4; i++) {
            if (a->info.dims[i] == ww->info.dims[i]) continue;
            if (!may_tile) fail(AF_ERR_NOT_SUPPORTED, "weighted statistics: the weights must have the dims of the input (no broadcast)");
            AFB_ARG(arg_w, ww->info.dims[i] == 1);                               // mean.cpp:118
            rep[i] = (unsigned)a->info.dims[i];
            tile = true;
        }
        }
        catch (...) { return ::afb::current_exception_to_err(); }
    }
    if (tile) return af_tile(&wt.h, weights, rep[0], rep[1], rep[2], rep[3]);
    return af_retain_array(&wt.h, weights);
}
af_err weighted_cast(H &x, H &w, const af_array in, const af_array weights) {
    af_dtype t;
    AFB_TRY(af_get_type(&t, in));
    const af_dtype tc = (t == f64 || t == s64 || t == u64) ? f64 : f32;
    AFB_TRY(af_cast(&x.h, in, tc));
    AFB_TRY(af_cast(&w.h, weights, tc));
    return AF_SUCCESS;
}
}  // namespace

af_err af_mean_weighted(af_array *out, const af_array in, const af_array weights, const dim_t dim) {
    {
        AFB_API_BEGIN
        AFB_ARG(3, dim >= 0 && dim <= 3);
        AFB_ARG(0, out != nullptr);
        }
        catch (...) { return ::afb::current_exception_to_err(); }
    }
    H x, w, wt, xw, num, den, q;
    AFB_TRY(weighted_operands(wt, in, weights, 2, true));
    AFB_TRY(weighted_cast(x, w, in, wt.h));
    AFB_TRY(af_mul(&xw.h, x.h, w.h, false));
    AFB_TRY(af_sum(&num.h, xw.h, (int)dim));
    AFB_TRY(af_sum(&den.h, w.h, (int)dim));
    AFB_TRY(af_div(&q.h, num.h, den.h, false));
    AFB_TRY(af_eval(q.h));
    *out = q.h;
    q.h = nullptr;
    return AF_SUCCESS;
}

af_err af_mean_all_weighted(double *real, double *imag, const af_array in, const af_array weights) {
    {
        AFB_API_BEGIN
        AFB_ARG(0, real != nullptr);
        }
        catch (...) { return ::afb::current_exception_to_err(); }
    }
    H x, w, wt, xw;
    AFB_TRY(weighted_operands(wt, in, weights, 3));
    AFB_TRY(weighted_cast(x, w, in, weights));
    AFB_TRY(af_mul(&xw.h, x.h, w.h, false));
    double num = 0, den = 0, im = 0;
    AFB_TRY(af_sum_all(&num, &im, xw.h));          // fused: one pass over x and w
    AFB_TRY(af_sum_all(&den, &im, w.h));
    *real = num / den;
    if (imag) *imag = 0;
    return AF_SUCCESS;
}

af_err af_var_all_weighted(double *real, double *imag, const af_array in, const af_array weights) {
    double mean = 0, im = 0;
    AFB_TRY(af_mean_all_weighted(&mean, &im, in, weights));
    H x, w, m, d, d2, wd2;
    AFB_TRY(weighted_cast(x, w, in, weights));
    af_dtype tc;
    AFB_TRY(af_get_type(&tc, x.h));
    dim_t dims[4];
    AFB_TRY(af_get_dims(&dims[0], &dims[1], &dims[2], &dims[3], x.h));
    AFB_TRY(af_constant(&m.h, mean, 4, dims, tc));
    AFB_TRY(af_sub(&d.h, x.h, m.h, false));
    AFB_TRY(af_mul(&d2.h, d.h, d.h, false));
    AFB_TRY(af_mul(&wd2.h, d2.h, w.h, false));
    double num = 0, den = 0;
    AFB_TRY(af_sum_all(&num, &im, wd2.h));         // fused: sum(w * (x - mean)^2) in one pass
    AFB_TRY(af_sum_all(&den, &im, w.h));
    *real = num / den;
    if (imag) *imag = 0;
    return AF_SUCCESS;
}

af_err af_var_weighted(af_array *out, const af_array in, const af_array weights, const dim_t dim) {
    {
        AFB_API_BEGIN
        AFB_ARG(3, dim >= 0 && dim <= 3);                                        // var.cpp:267
        AFB_ARG(0, out != nullptr);
        }
        catch (...) { return ::afb::current_exception_to_err(); }
    }
    H m;
    return meanvar_weighted(&m.h, out, in, weights, (int)AF_VARIANCE_POPULATION, dim);      // var.cpp:283-325
}

}  // extern "C"


// The dim-wise weighted mean / variance pair of the reference (var.cpp:104-142): mean = sum(w x) / sum(w) and
//     var = sum_dim((x - mean)^2) / (sum(w) - (bias == SAMPLE))
// -- the weights enter the mean and the normaliser only, the squared deviations are NOT weighted (var.cpp:134-139).  No
// broadcast of the means is needed: with m the plain mean and M2 = sum((x - m)^2) of ONE Welford pass over x (launch_var),
//     sum((x - mean)^2) = M2 + n (m - mean)^2,
// so x is read twice (the moments, the weighted sum) and w twice (its sum, the weighted sum); everything after that runs on
// the reduced arrays.
static af_err meanvar_weighted(af_array *mean, af_array *var, const af_array in, const af_array weights, int bias, dim_t dim) {
    H x, w, wt, wm, m, v, wsum, d, d2, t, nn, t2, den, q, one;
    AFB_TRY(weighted_operands(wt, in, weights, 2));
    AFB_TRY(af_mean_weighted(&wm.h, in, weights, dim));
    AFB_TRY(moments_dim(&v.h, &m.h, true, in, BIAS_POPULATION, dim, false, false, 5));
    AFB_TRY(weighted_cast(x, w, in, weights));
    AFB_TRY(af_sum(&wsum.h, w.h, (int)dim));
    af_dtype tc;
    AFB_TRY(af_get_type(&tc, wm.h));
    dim_t id[4], od[4];
    AFB_TRY(af_get_dims(&id[0], &id[1], &id[2], &id[3], in));
    AFB_TRY(af_get_dims(&od[0], &od[1], &od[2], &od[3], wm.h));
    AFB_TRY(af_sub(&d.h, m.h, wm.h, false));
    AFB_TRY(af_mul(&d2.h, d.h, d.h, false));
    AFB_TRY(af_add(&t.h, v.h, d2.h, false));
    AFB_TRY(af_constant(&nn.h, (double)id[dim], 4, od, tc));
    AFB_TRY(af_mul(&t2.h, t.h, nn.h, false));
    if (bias == BIAS_SAMPLE) {                                                   // var.cpp:127-129
        AFB_TRY(af_constant(&one.h, 1.0, 4, od, tc));
        AFB_TRY(af_sub(&den.h, wsum.h, one.h, false));
    } else {
        AFB_TRY(af_retain_array(&den.h, wsum.h));
    }
    AFB_TRY(af_div(&q.h, t2.h, den.h, false));
    AFB_TRY(af_eval(q.h));
    *mean = wm.h;
    *var = q.h;
    wm.h = q.h = nullptr;
    return AF_SUCCESS;
}
