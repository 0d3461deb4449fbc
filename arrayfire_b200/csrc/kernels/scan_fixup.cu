// scan_fixup.cu -- cross-shard fix-up of a sharded scan_by_key (new functionality).
//
// A shard is scanned locally without any carry-in (lookback_kernel, summary mode).  After the per-shard
// summaries have been all-gathered, the only elements whose value depends on other shards are those of the
// shard's FIRST segment, when that segment continues the previous shard's last run:
//     out[i] = op(carry, out[i])   for i < first_head
// `carry` is the reduction of the run's elements that live on lower ranks.  The prefix is a mean run
// length (about 512 elements at BASELINE config C4), so this kernel is launch-latency sized; a run that
// covers a whole shard degenerates to one extra read+write pass over that shard.
#include "scan_host.cuh"

namespace afb {

template<int OP, typename T>
__global__ void __launch_bounds__(256) scan_fixup_kernel(T *out, dim_t count, const T carry) {
    typedef Rule<OP, T, T> R;
    for (dim_t i = (dim_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (dim_t)gridDim.x * blockDim.x)
        out[i] = R::join(carry, out[i]);
}

template<int OP, typename T>
static void fixup(const LaunchCtx &c, void *out, dim_t count, const void *carry_host) {
    T carry;
    memcpy(&carry, carry_host, sizeof(T));
    dim_t blocks = (count + 255) / 256;
    const dim_t cap = (dim_t)c.sms * 16;
    if (blocks > cap) blocks = cap;
    scan_fixup_kernel<OP, T><<<(unsigned)blocks, 256, 0, c.stream>>>((T *)out, count, carry);
    note_launches(1);
    post_launch_check("scan_fixup_kernel");
}

template<int OP>
static bool fixup_family(const LaunchCtx &c, af_dtype t, void *out, dim_t count, const void *carry_host) {
#define X(code, T)                                  \
    if (t == code) {                                \
        fixup<OP, T>(c, out, count, carry_host);    \
        return true;                                \
    }
    AFB_SBK_VALS(X)
#undef X
    return false;
}

bool launch_scan_fixup(const LaunchCtx &c, Op op, af_dtype t, void *out, dim_t count, const void *carry_host) {
    if (count <= 0) return true;
    switch (op) {
        case OP_ADD: return fixup_family<OP_ADD>(c, t, out, count, carry_host);
        case OP_MUL: return fixup_family<OP_MUL>(c, t, out, count, carry_host);
        case OP_MIN: return fixup_family<OP_MIN>(c, t, out, count, carry_host);
        case OP_MAX: return fixup_family<OP_MAX>(c, t, out, count, carry_host);
        default: return false;
    }
}

}  // namespace afb
