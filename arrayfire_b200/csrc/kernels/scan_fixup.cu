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
#include "peer_comm.cuh"

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

// ------------------------------------------------------------------------------------------------
// sbk_exchange_fixup_kernel: the cross-GPU step of a sharded scan_by_key in ONE launch, no host round trip.
// Thread 0 of CTA 0 builds this shard's record -- the reduction of its last open segment, "the shard contains a segment
// head", its first and last key -- exchanges it with the other ranks over peer memory (peer_comm.cuh) and derives the
// carry that enters this shard: the tail of rank-1 when the key continues across the boundary, extended further down
// while the lower shards are single segments of the same key.  The result (carry, element count = position of the
// shard's first head) goes to a generation-tagged device slot; every CTA -- the grid is at most one CTA per SM, so all
// are co-resident with CTA 0 -- waits for it and applies out[i] = op(carry, out[i]) to its share of the first segment.
// ------------------------------------------------------------------------------------------------
template<typename T> struct alignas(16) SbkCarrySlot {
    T carry;
    unsigned long long count;
    uint ready;     // epoch when valid
};

template<int OP, typename T, typename Tk>
__global__ void __launch_bounds__(256) sbk_exchange_fixup_kernel(T *out, dim_t n, const Tk *__restrict__ keys,
                                                                  const unsigned char *__restrict__ summary,
                                                                  SbkCarrySlot<T> *slot, const PeerComm comm) {
    typedef Rule<OP, T, T> R;
    __shared__ T sm_carry;
    __shared__ unsigned long long sm_count;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            PeerRecord rec, all_recs[kPeerMax];
#pragma unroll
            for (int j = 0; j < 16; j++) rec.w[j] = 0u;
            const T tail = reinterpret_cast<const Slot<T> *>(summary)->v;
            const uint has_head = *reinterpret_cast<const uint *>(summary + 16);
            rec_put(rec, tail, 0);
            rec_put(rec, has_head, 16);
            rec_put(rec, (unsigned long long)keys[0], 32);
            rec_put(rec, (unsigned long long)keys[n - 1], 40);
            peer_exchange_serial<12>(comm, rec, all_recs, comm.rank + 1);   // only the LOWER ranks' records are needed
            // carry-in of this rank: the tail of rank-1 when the key continues across the boundary, extended by that
            // shard's own carry while it is a single run of the same key (a run spanning several shards)
            bool have = false;
            T carry = R::init();
            for (int r = 1; r <= comm.rank; r++) {
                const PeerRecord &prev = all_recs[r - 1], &cur = all_recs[r];
                if (rec_get<unsigned long long>(prev, 40) != rec_get<unsigned long long>(cur, 32)) {
                    have = false;     // the key changes at this boundary: nothing is carried across it
                    continue;
                }
                T t = rec_get<T>(prev, 0);
                if (have && rec_get<uint>(prev, 16) == 0u) t = R::join(carry, t);
                carry = t;
                have = true;
            }
            const unsigned long long first_head = *reinterpret_cast<const unsigned long long *>(summary + 24);
            unsigned long long count = have ? (first_head < (unsigned long long)n ? first_head : (unsigned long long)n) : 0ULL;
            if (comm.rank == 0) count = 0;
            slot->carry = carry;
            slot->count = count;
            __threadfence();
            st_release(&slot->ready, comm.epoch);
        }
        unsigned long long spins = 0;
        while (ld_acquire(&slot->ready) != comm.epoch) {
            if (++spins > (1ull << 31)) __trap();
        }
        sm_carry = slot->carry;
        sm_count = slot->count;
    }
    __syncthreads();
    const T carry = sm_carry;
    const dim_t count = (dim_t)sm_count;
    for (dim_t i = (dim_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (dim_t)gridDim.x * blockDim.x)
        out[i] = R::join(carry, out[i]);
}

template<int OP, typename T, typename Tk>
static void sbk_exchange(const LaunchCtx &c, void *out, dim_t n, const void *keys, const void *summary, const ShardLink &link) {
    static void *slots[64] = {nullptr};   // one 64-byte slot per device, allocated once
    const int dv = c.device & 63;
    if (!slots[dv]) {
        AFB_CUDA(cudaMalloc(&slots[dv], 64));
        AFB_CUDA(cudaMemsetAsync(slots[dv], 0, 64, c.stream));
    }
    static_assert(sizeof(SbkCarrySlot<T>) <= 64, "carry slot");
    const PeerComm pc = peer_comm_of(&link);
    int grid = c.sms;                     // <= one CTA per SM: every CTA is resident together with CTA 0
    sbk_exchange_fixup_kernel<OP, T, Tk><<<grid, 256, 0, c.stream>>>((T *)out, n, (const Tk *)keys, (const unsigned char *)summary,
                                                                       (SbkCarrySlot<T> *)slots[dv], pc);
    note_launches(1);
    post_launch_check("sbk_exchange_fixup_kernel");
}

template<int OP>
static bool sbk_exchange_family(const LaunchCtx &c, af_dtype t, af_dtype tk, void *out, dim_t n, const void *keys,
                                const void *summary, const ShardLink &link) {
    const bool k32 = (tk == s32 || tk == u32), k64 = (tk == s64 || tk == u64);
    if (!k32 && !k64) return false;
#define X(code, T)                                                                  \
    if (t == code) {                                                                \
        if (k32) sbk_exchange<OP, T, uint>(c, out, n, keys, summary, link);         \
        else sbk_exchange<OP, T, uintl>(c, out, n, keys, summary, link);            \
        return true;                                                                \
    }
    AFB_SBK_VALS(X)
#undef X
    return false;
}

bool launch_sbk_exchange_fixup(const LaunchCtx &c, Op op, af_dtype t, af_dtype tk, void *out, dim_t n, const void *keys,
                               const void *summary, const ShardLink &link) {
    if (n <= 0 || link.world < 2) return true;
    switch (op) {
        case OP_ADD: return sbk_exchange_family<OP_ADD>(c, t, tk, out, n, keys, summary, link);
        case OP_MUL: return sbk_exchange_family<OP_MUL>(c, t, tk, out, n, keys, summary, link);
        case OP_MIN: return sbk_exchange_family<OP_MIN>(c, t, tk, out, n, keys, summary, link);
        case OP_MAX: return sbk_exchange_family<OP_MAX>(c, t, tk, out, n, keys, summary, link);
        default: return false;
    }
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
