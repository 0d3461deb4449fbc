// peer.cu -- the cross-GPU exchange of the O(1) per-shard partials, fused into ONE tiny kernel over peer memory.
//
// New functionality (the reference is single-device).  Every sharded whole-array op ends with "each
// rank contributes 64 bytes, every rank needs all of them, the result goes to the host".  Done with NCCL that is a
// collective launch (~10 us of CPU, ~20 us of GPU) + a read-back + a stream synchronisation per op: ~90 us, more than
// the kernels themselves once 2^30 elements are split over 8 GPUs.  Here the ranks share a symmetric buffer (peer-mapped
// through NVLink / NVSwitch, set up once by the host layer) and one CTA per rank
//   1. stores its 64-byte payload straight into slot [rank] of EVERY peer's buffer (16-byte peer stores),
//   2. raises flag [rank] = epoch in every peer's buffer (release, system scope),
//   3. waits until all `world` flags in its OWN buffer carry this epoch (acquire, system scope),
//   4. copies the gathered world x 64 bytes to pinned host memory and raises a host flag the caller spins on.
// Slots and flags are double-buffered by epoch parity: a rank can only reach epoch e+2 after every peer has
// published e+1, i.e. after every peer has consumed e.
#include "device_ops.cuh"

namespace afb {

constexpr int kPeerMax = 8;
constexpr int kPeerPayload = 64;
constexpr int kPeerFlagOff = 2 * kPeerMax * kPeerPayload;   // flags follow the 2 x 8 slots

struct PeerArgs {
    unsigned long long bufs[kPeerMax];   // device addresses of every rank's symmetric buffer (index = rank)
    int rank, world;
    unsigned epoch;
    const uint4 *payload;                // this rank's 64-byte record (device)
    uint4 *host_out;                     // pinned host: world x 64 bytes
    unsigned *host_flag;                 // pinned host
};

__global__ void __launch_bounds__(64) peer_exchange_kernel(const PeerArgs a) {
    const int t = threadIdx.x;
    const int par = (int)(a.epoch & 1u);
    if (t < a.world * 4) {
        const int peer = t >> 2, v = t & 3;
        uint4 val = a.payload[v];
        uint4 *dst = reinterpret_cast<uint4 *>(a.bufs[peer] + (size_t)((par * kPeerMax + a.rank) * kPeerPayload)) + v;
        asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
    }
    __threadfence_system();
    __syncthreads();
    if (t < a.world) {
        unsigned *f = reinterpret_cast<unsigned *>(a.bufs[t] + kPeerFlagOff) + par * kPeerMax + a.rank;
        st_release_sys(f, a.epoch);
        const unsigned *mine = reinterpret_cast<const unsigned *>(a.bufs[a.rank] + kPeerFlagOff) + par * kPeerMax + t;
        unsigned spins = 0;
        while (ld_acquire_sys(mine) != a.epoch) {
            if (++spins > (1u << 28)) __trap();   // minutes: ranks may legitimately arrive seconds apart; never is a protocol error
        }
    }
    __syncthreads();
    if (t < a.world * 4) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.bufs[a.rank] + (size_t)((par * kPeerMax + (t >> 2)) * kPeerPayload)) + (t & 3);
        uint4 val;
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "l"(src) : "memory");
        a.host_out[t] = val;
    }
    __threadfence_system();
    __syncthreads();
    if (t == 0) st_release_sys(a.host_flag, a.epoch);
}

void launch_peer_exchange(const LaunchCtx &c, const void *payload, const unsigned long long *bufs, int rank, int world,
                          unsigned epoch, void *host_out, unsigned *host_flag) {
    PeerArgs a;
    for (int i = 0; i < kPeerMax; i++) a.bufs[i] = i < world ? bufs[i] : 0ULL;
    a.rank = rank;
    a.world = world;
    a.epoch = epoch;
    a.payload = (const uint4 *)payload;
    a.host_out = (uint4 *)host_out;
    a.host_flag = host_flag;
    peer_exchange_kernel<<<1, 64, 0, c.stream>>>(a);
    note_launches(1);
    post_launch_check("peer_exchange_kernel");
}

}  // namespace afb
