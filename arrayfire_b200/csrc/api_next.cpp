// api_next.cpp -- SURVEY.md section 8(f)-3: the callers that sit right next to the scan / compaction path.
//   af_sat         /root/reference/src/api/c/sat.cpp:32-61, imgproc_common.hpp:25-30   (cast, scan dim 0, scan dim 1)
//   af_diff1/2     /root/reference/src/api/c/diff.cpp:36-115, cpu/kernel/diff.hpp:18-81
//   af_set_unique  /root/reference/src/api/c/set.cpp:32-64, cpu/set.cpp:32-51 -- SORTED input only (the adjacent-differs
//                  mask + the where compaction + a gather); unsorted input needs a sort, which is outside this path
// All device work is this library's own kernels; no CPU fallback.
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
struct HandleGuard {   // releases an af_array produced by a public entry point
    af_array h = nullptr;
    ~HandleGuard() { if (h) af_release_array(h); }
};

af_err diff_common(af_array *out, const af_array in_h, int dim, int order) {
    AFB_API_BEGIN
    AFB_ARG(2, dim >= 0 && dim < 4);
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    const af_dtype t = in->info.type;
    if (in->info.dims[dim] < order + 1) return af_create_handle(out, 0, nullptr, t);   // diff.cpp:45-47
    if (t == f16) AFB_TYPE_ERROR(1, t);
    check_same_device(*in);
    dim_t od[4] = {in->info.dims[0], in->info.dims[1], in->info.dims[2], in->info.dims[3]};
    od[dim] -= order;
    Holder r;
    r.a = new_array(t, od);
    if (r.a->elements() > 0) {
        OpLock lk(in->device());
        LaunchCtx c = ctx_for_active_device();
        if (!launch_diff(c, t, view_of(*in), dim, order, r.a->data())) AFB_TYPE_ERROR(1, t);
    }
    *out = r.release();
    AFB_API_END
}
}  // namespace

extern "C" {

af_err af_diff1(af_array *out, const af_array in, const int dim) { return diff_common(out, in, dim, 1); }
af_err af_diff2(af_array *out, const af_array in, const int dim) { return diff_common(out, in, dim, 2); }

af_err af_sat(af_array *out, const af_array in_h) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    AFB_ARG(1, in->ndims() >= 2);
    const af_dtype t = in->info.type;
    if (t == c32 || t == c64 || t == f16) AFB_TYPE_ERROR(1, t);      // sat.cpp:44-57
    // b8 sums as `int` (sat<int, char>); every other narrow type takes the scan's own promotion (s8/s16 -> s32, u8/u16 -> u32)
    HandleGuard casted, rows;
    af_array src = in_h;
    if (t == b8) {
        af_err e = af_cast(&casted.h, in_h, s32);
        if (e != AF_SUCCESS) return e;
        src = casted.h;
    }
    af_err e = af_scan(&rows.h, src, 0, AF_BINARY_ADD, true);
    if (e != AF_SUCCESS) return e;
    return af_scan(out, rows.h, 1, AF_BINARY_ADD, true);
    AFB_API_END
}

af_err af_set_unique(af_array *out, const af_array in_h, const bool is_sorted) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    Array *in = from_handle(in_h);
    if (in->elements() == 0 || in->elements() == 1) return af_retain_array(out, in_h);   // set.cpp:36-38
    int non_singular = 0;
    for (int i = 0; i < 4; i++) non_singular += in->info.dims[i] != 1;
    AFB_ARG(1, non_singular == 1);                                                      // isVector()
    const af_dtype t = in->info.type;
    if (t == c32 || t == c64 || t == f16) AFB_TYPE_ERROR(1, t);
    if (!is_sorted) fail(AF_ERR_NOT_SUPPORTED, "af_set_unique: unsorted input needs af_sort, which libafb200 does not provide");
    check_same_device(*in);
    Holder dense;
    const Array *src = in;
    if (!in->is_linear()) { dense.a = make_dense_copy(*in); src = dense.a; }
    const dim_t n = in->elements();
    dim_t nd[4] = {n, 1, 1, 1};
    Holder mask;
    mask.a = new_array(u8, nd);
    {
        OpLock lk(in->device());
        LaunchCtx c = ctx_for_active_device();
        if (!launch_differs(c, t, src->data(), n, (unsigned char *)mask.a->data())) AFB_TYPE_ERROR(1, t);
    }
    HandleGuard idx;
    af_err e = af_where(&idx.h, (af_array)mask.a);
    if (e != AF_SUCCESS) return e;
    Array *ix = from_handle(idx.h);
    dim_t rd[4] = {ix->elements(), 1, 1, 1};
    Holder res;
    res.a = new_array(t, rd);
    {
        OpLock lk(in->device());
        LaunchCtx c = ctx_for_active_device();
        if (!launch_gather(c, t, src->data(), (const unsigned *)ix->data(), rd[0], res.a->data())) AFB_TYPE_ERROR(1, t);
    }
    *out = res.release();
    AFB_API_END
}

}  // extern "C"
