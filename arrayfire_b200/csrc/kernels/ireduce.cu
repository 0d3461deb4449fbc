// ireduce.cu -- af_imin / af_imax (+_all): (value, u32 index) reductions for all 14 dtypes.
#include "reduce_host.cuh"
#include "reduce_dispatch.cuh"
namespace afb {
template<int OP, typename T>
static void ireduce_one(const LaunchCtx &c, const View &in, int dim, bool all, void *out_val, unsigned *out_loc,
                        const HostRet *ret, const ShardLink *link) {
    typedef ArgPolicy<OP, T> P;
    typename P::Params prm;
    prm.seed = (link && link->world > 1 && link->idx_base != 0) ? 0 : 1;
    typename P::Out o;
    o.val = (T *)out_val;
    o.loc = out_loc;
    if (all) run_all<P>(c, in, prm, o, ret, link);
    else run_reduce<P>(c, in, dim, prm, o);
}
// ---- af_max_ragged (/root/reference/src/api/c/reduce.cpp:1002-1078, cpu/kernel/ireduce.hpp:86-104): the (value, index)
// reduction of line j along `dim` over its first min(len, rlen[j]) elements only; element 0 always seeds the result
// (even for rlen[j] == 0).  One warp per line: a side function of the path, not a bandwidth kernel.
template<int OP, typename T>
__global__ void __launch_bounds__(256) ragged_kernel(const T *__restrict__ in, dim_t len, dim_t sd, dim_t nlines, dim_t l0,
                                                     dim_t l1, dim_t s0, dim_t s1, dim_t s2, const uint *__restrict__ rlen,
                                                     T *out_val, uint *out_loc) {
    typedef ArgPolicy<OP, T> P;
    typedef typename P::Acc Acc;
    const int lane = threadIdx.x & 31;
    const dim_t warps = (dim_t)gridDim.x * (blockDim.x >> 5);
    for (dim_t line = (dim_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); line < nlines; line += warps) {
        // line id -> coordinates over the three non-reduced dims (extents l0, l1, rest), input strides s0, s1, s2
        const dim_t i0 = line % l0, r = line / l0;
        const dim_t i1 = r % l1, i2 = r / l1;
        const T *base = in + i0 * s0 + i1 * s1 + i2 * s2;
        dim_t lim = (dim_t)rlen[line];
        if (lim > len) lim = len;
        typename P::Params prm;
        Acc acc = P::init();
        if (lane == 0) acc = P::load(prm, ldg_stream<T>(base), 0u);          // the seed
        for (dim_t i = lane; i < lim; i += 32) acc = P::join(acc, P::load(prm, ldg_stream<T>(base + i * sd), (uint)i));
        acc = warp_reduce<P>(acc);
        if (lane == 0) {
            typename P::Out o;
            o.val = out_val;
            o.loc = out_loc;
            P::store(o, (size_t)line, acc);
        }
    }
}

template<int OP, typename T>
static void ragged_one(const LaunchCtx &c, const View &in, int dim, const unsigned *rlen, void *out_val, unsigned *out_loc) {
    // output (and rlen) are dense over the non-reduced dims in order; a line id enumerates them the same way
    dim_t ext[3], str[3];
    int m = 0;
    for (int i = 0; i < 4; i++)
        if (i != dim) { ext[m] = in.dims[i]; str[m] = in.strides[i]; m++; }
    const dim_t nlines = ext[0] * ext[1] * ext[2];
    if (nlines == 0) return;
    dim_t blocks = (nlines + 7) / 8;
    const dim_t cap = (dim_t)c.sms * 16;
    if (blocks > cap) blocks = cap;
    ragged_kernel<OP, T><<<(unsigned)blocks, 256, 0, c.stream>>>((const T *)in.ptr, in.dims[dim], in.strides[dim], nlines, ext[0] > 0 ? ext[0] : 1,
                                                                  ext[1] > 0 ? ext[1] : 1, str[0], str[1], str[2], rlen, (T *)out_val, out_loc);
    note_launches(1);
    post_launch_check("ragged_kernel");
}

bool launch_ireduce_ragged(const LaunchCtx &c, Op op, af_dtype t, const View &in, int dim, const unsigned *rlen,
                           void *out_val, unsigned *out_loc) {
    if (op != OP_MAX && op != OP_MIN) return false;
#define X(code, T)                                                                  \
    if (t == code) {                                                                \
        if (op == OP_MAX) ragged_one<OP_MAX, T>(c, in, dim, rlen, out_val, out_loc); \
        else ragged_one<OP_MIN, T>(c, in, dim, rlen, out_val, out_loc);             \
        return true;                                                                \
    }
    AFB_ALL_TYPES(X)
#undef X
    return false;
}

static bool route(const LaunchCtx &c, Op op, af_dtype t, const View &in, int dim, bool all, void *v, unsigned *l,
                  const HostRet *ret, const ShardLink *link) {
    if (op != OP_MIN && op != OP_MAX) return false;
#define X(code, T)                                                              \
    if (t == code) {                                                            \
        if (op == OP_MIN) ireduce_one<OP_MIN, T>(c, in, dim, all, v, l, ret, link);        \
        else ireduce_one<OP_MAX, T>(c, in, dim, all, v, l, ret, link);                     \
        return true;                                                            \
    }
    AFB_ALL_TYPES(X)
#undef X
    return false;
}
bool launch_ireduce(const LaunchCtx &c, Op op, af_dtype t, const View &in, int dim, void *v, unsigned *l) {
    return route(c, op, t, in, dim, false, v, l, nullptr, nullptr);
}
bool launch_ireduce_all(const LaunchCtx &c, Op op, af_dtype t, const View &in, void *v, unsigned *l, const HostRet *ret,
                        const ShardLink *link) {
    return route(c, op, t, in, 0, true, v, l, ret, link);
}
}  // namespace afb
