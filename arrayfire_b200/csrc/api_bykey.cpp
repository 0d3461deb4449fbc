// api_bykey.cpp -- SURVEY.md section 8(f)-1: the af_*_by_key reduction family and af_max_ragged.
//
// Argument checks, promotion tables and error codes restate the reference's C layer:
//   /root/reference/src/api/c/reduce.cpp:46-80 (reduce_by_key / reduce_key: s32 / u32 keys only),
//       122-154 (reduce_by_key_type: count / all / any), 156-228 (reduce_by_key_common: min / max),
//       230-304 (reduce_promote_by_key: sum / product, b8 -> count), 500-554 (entry points),
//       1002-1078 (rreduce_common / af_max_ragged)
// Results follow the CPU backend (src/backend/cpu/reduce.cpp:73-108, kernel/reduce.hpp:64-160, kernel/ireduce.hpp:86-104).
// The kernels are this library's own (kernels/rbk_kernels.cuh, kernels/ireduce.cu); there is no CPU fallback.
#include "common.hpp"

using namespace afb;

namespace afb {
Array *make_dense_copy(const Array &a);
}

namespace {

struct Holder {
    Array *a = nullptr;
    ~Holder() { delete a; }
    Array *release() { Array *r = a; a = nullptr; return r; }
};

enum Family { FAM_PROMOTE, FAM_COMMON, FAM_TYPE };

bool is_vector(const Array &a) {  // ArrayInfo::isVector, common/ArrayInfo.cpp:122-130
    int singular = 0, non_singular = 0;
    for (int i = 0; i < 4; i++) {
        non_singular += (a.info.dims[i] != 0 && a.info.dims[i] != 1);
        singular += (a.info.dims[i] == 1);
    }
    return singular == 3 && non_singular == 1;
}

af_err reduce_by_key(af_array *keys_out, af_array *vals_out, const af_array keys_h, const af_array vals_h, int dim,
                     Family fam, Op op, bool change_nan, double nanval) {
    AFB_API_BEGIN
    AFB_ARG(4, dim >= 0);
    AFB_ARG(4, dim < 4);
    AFB_ARG(0, keys_out != nullptr);
    AFB_ARG(1, vals_out != nullptr);
    Array *keys = from_handle(keys_h);
    Array *vals = from_handle(vals_h);
    AFB_ARG(2, is_vector(*keys));
    // count / all / any compare against the element count, the other families against dims[0] (reduce.cpp:135,172,246)
    const dim_t klen = fam == FAM_TYPE ? keys->elements() : keys->info.dims[0];
    AFB_ARG(2, vals->info.dims[dim] == klen);
    const af_dtype tk = keys->info.type, ti = vals->info.type;
    if (!(tk == s32 || tk == u32)) AFB_TYPE_ERROR(2, tk);
    Op eff = op;
    af_dtype to = ti;
    switch (fam) {
        case FAM_PROMOTE:
            switch (ti) {
                case f32: case f64: case c32: case c64: case u32: case s32: case u64: case s64: to = ti; break;
                case u16: case u8: to = u32; break;
                case s16: case s8: to = s32; break;
                case f16: to = f32; break;
                case b8: eff = OP_NOTZERO; to = u32; break;   // both sum and product of b8 count non-zeros (reduce.cpp:292-295)
            }
            break;
        case FAM_COMMON: to = ti; break;
        case FAM_TYPE: to = (op == OP_NOTZERO) ? u32 : b8; break;
    }
    check_same_device(*keys);
    check_same_device(*vals);
    const dim_t n = klen;
    // the kernels read the keys as a dense vector
    const Array *kd = keys;
    Holder ktmp;
    if (!(keys->info.dims[0] == n && keys->info.strides[0] == 1)) {
        ktmp.a = make_dense_copy(*keys);
        kd = ktmp.a;
    }
    const dim_t tile = rbk_tile_elems();
    const dim_t ntiles = (n + tile - 1) / tile;
    Holder bases;
    const dim_t bd[4] = {rbk_scratch_elems(ntiles), 1, 1, 1};
    bases.a = new_array(u32, bd);
    // ---- pass 1: number of runs (the host needs it to size the outputs)
    HostRet ret;
    LaunchCtx c;
    {
        std::lock_guard<std::mutex> lk(op_mutex(vals->device()));
        c = ctx_for_active_device();
        ret = c.scratch->next_host_ret();
        launch_rbk_count(c, kd->data(), n, (unsigned *)bases.a->data(), ret);
    }
    wait_host_ret(c, ret);
    const dim_t n_reduced = (dim_t) * (const unsigned long long *)ret.host;
    // ---- pass 2
    dim_t kod[4] = {keys->info.dims[0], keys->info.dims[1], keys->info.dims[2], keys->info.dims[3]};
    kod[0] = n_reduced;
    if (keys->info.dims[0] != n) { kod[0] = n_reduced; kod[1] = kod[2] = kod[3] = 1; }
    dim_t vod[4] = {vals->info.dims[0], vals->info.dims[1], vals->info.dims[2], vals->info.dims[3]};
    vod[dim] = n_reduced;
    Holder ko, vo;
    ko.a = new_array(tk, kod);
    vo.a = new_array(to, vod);
    {
        std::lock_guard<std::mutex> lk(op_mutex(vals->device()));
        c = ctx_for_active_device();
        if (!launch_rbk_reduce(c, eff, ti, to, kd->data(), n, view_of(*vals), dim, change_nan, nanval,
                               (const unsigned *)bases.a->data(), ko.a->data(), vo.a->data(), n_reduced))
            AFB_TYPE_ERROR(3, ti);
    }
    *keys_out = ko.release();
    *vals_out = vo.release();
    AFB_API_END
}

}  // namespace

extern "C" {

af_err af_sum_by_key(af_array *ko, af_array *vo, const af_array k, const af_array v, const int dim) { return reduce_by_key(ko, vo, k, v, dim, FAM_PROMOTE, OP_ADD, false, 0); }
af_err af_sum_by_key_nan(af_array *ko, af_array *vo, const af_array k, const af_array v, const int dim, const double nanval) { return reduce_by_key(ko, vo, k, v, dim, FAM_PROMOTE, OP_ADD, true, nanval); }
af_err af_product_by_key(af_array *ko, af_array *vo, const af_array k, const af_array v, const int dim) { return reduce_by_key(ko, vo, k, v, dim, FAM_PROMOTE, OP_MUL, false, 0); }
af_err af_product_by_key_nan(af_array *ko, af_array *vo, const af_array k, const af_array v, const int dim, const double nanval) { return reduce_by_key(ko, vo, k, v, dim, FAM_PROMOTE, OP_MUL, true, nanval); }
af_err af_min_by_key(af_array *ko, af_array *vo, const af_array k, const af_array v, const int dim) { return reduce_by_key(ko, vo, k, v, dim, FAM_COMMON, OP_MIN, false, 0); }
af_err af_max_by_key(af_array *ko, af_array *vo, const af_array k, const af_array v, const int dim) { return reduce_by_key(ko, vo, k, v, dim, FAM_COMMON, OP_MAX, false, 0); }
af_err af_count_by_key(af_array *ko, af_array *vo, const af_array k, const af_array v, const int dim) { return reduce_by_key(ko, vo, k, v, dim, FAM_TYPE, OP_NOTZERO, false, 0); }
af_err af_all_true_by_key(af_array *ko, af_array *vo, const af_array k, const af_array v, const int dim) { return reduce_by_key(ko, vo, k, v, dim, FAM_TYPE, OP_AND, false, 0); }
af_err af_any_true_by_key(af_array *ko, af_array *vo, const af_array k, const af_array v, const int dim) { return reduce_by_key(ko, vo, k, v, dim, FAM_TYPE, OP_OR, false, 0); }

af_err af_max_ragged(af_array *val, af_array *idx, const af_array in_h, const af_array ragged_len, const int dim) {
    AFB_API_BEGIN
    AFB_ARG(3, dim >= 0);
    AFB_ARG(3, dim < 4);
    AFB_ARG(0, val != nullptr);
    AFB_ARG(1, idx != nullptr);
    Array *in = from_handle(in_h);
    AFB_ARG(2, in->ndims() > 0);
    check_same_device(*in);
    if (dim >= in->ndims()) {  // values = input, indices = 0 (reduce.cpp:1013-1017)
        Holder loc;
        loc.a = new_array(u32, in->info.dims);
        LaunchCtx c = ctx_for_active_device();
        launch_fill(c, u32, loc.a->data(), loc.a->elements(), 0.0, 0.0);
        *val = retain(*in);
        *idx = loc.release();
        return AF_SUCCESS;
    }
    Array *rl = from_handle(ragged_len);
    dim_t od[4] = {in->info.dims[0], in->info.dims[1], in->info.dims[2], in->info.dims[3]};
    od[dim] = 1;
    for (int i = 0; i < 4; i++) AFB_ARG(4, rl->info.dims[i] == od[i]);
    if (rl->info.type != u32) AFB_TYPE_ERROR(4, rl->info.type);
    check_same_device(*rl);
    const Array *rd = rl;
    Holder rtmp;
    if (!rl->is_linear()) {
        rtmp.a = make_dense_copy(*rl);
        rd = rtmp.a;
    }
    Holder res, loc;
    res.a = new_array(in->info.type, od);
    loc.a = new_array(u32, od);
    {
        std::lock_guard<std::mutex> lk(op_mutex(in->device()));
        LaunchCtx c = ctx_for_active_device();
        if (!launch_ireduce_ragged(c, OP_MAX, in->info.type, view_of(*in), dim, (const unsigned *)rd->data(), res.a->data(),
                                   (unsigned *)loc.a->data()))
            AFB_TYPE_ERROR(2, in->info.type);
    }
    *val = res.release();
    *idx = loc.release();
    AFB_API_END
}

}  // extern "C"
