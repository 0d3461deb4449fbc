// peer_comm.cuh -- the cross-GPU combine of the sharded whole-array ops, done INSIDE the kernel that produced the shard's
// partial (the reference is single-device).
//
// Round 1 ended every sharded op with a second launch (peer_exchange_kernel) + a host spin + a fold in Python: a fixed
// ~50 us per op, more than the kernels themselves once 2^30 elements are split over 8 GPUs (3.65x at 8 GPUs).  Now the
// LAST thread of the producing kernel (the last-arriver CTA of span_kernel, the walker of the last tile of the pipelined
// scan / compaction kernels) does the exchange itself over NVLink / NVSwitch peer memory:
//   1. sends its record (<= 64 bytes) into slot [rank] of EVERY rank's symmetric buffer,
//   2. collects the records of the other ranks from its OWN buffer as they arrive,
//   3. folds them in RANK ORDER (deterministic, identical on every rank, equal to the serial answer) and hands the result
//      to the device output and / or the pinned host record of the call.
// One launch, no collective-library call, no host round trip between partial and combine.
//
// Wire format ("LL": flag in line, no fences): a record travels as 16-byte lines {data, epoch, data, epoch}, each written
// with ONE 16-byte store.  The receiver polls a line until both flag words carry the epoch of the op -- the data words of a
// line that shows the right flags are valid (8-byte store atomicity is all that is assumed, as in NCCL's LL protocol), so
// neither side needs a system-scope fence (round 2, first cut: st.release.sys per peer = one MEMBAR.SYS per peer, 16 us at
// 2 GPUs and 35 us at 8 per op).  Slots are double-buffered by epoch parity: a rank reaches epoch e+2 only after every peer
// has published e+1, i.e. after the kernel that consumed e has finished.  The buffers start zeroed and epochs are never 0.
// The buffers are peer-mapped once by the host layer (torch.distributed._symmetric_memory or cuMem + IPC handles).
#pragma once

#include "device_ops.cuh"

namespace afb {

constexpr int kPeerMax = 8;
constexpr int kPeerRecord = 64;                                    // payload bytes per rank and parity
constexpr int kPeerSlotBytes = 2 * kPeerRecord;                    // on the wire: every 8 payload bytes ride with 8 flag bytes
constexpr int kPeerBufBytes = 2 * kPeerMax * kPeerSlotBytes;       // 2048

struct PeerComm {
    unsigned long long bufs[kPeerMax];   // device addresses of every rank's symmetric buffer (index = rank)
    int rank, world;                     // world == 0: no exchange (single-GPU launch)
    unsigned epoch;
};

// host side: kernel argument from the API layer's ShardLink (world = 0: no exchange)
inline PeerComm peer_comm_of(const ShardLink *link) {
    PeerComm pc;
    pc.world = 0;
    pc.rank = 0;
    pc.epoch = 0;
    for (int r = 0; r < kPeerMax; r++) pc.bufs[r] = 0ULL;
    if (link && link->world > 1) {
        for (int r = 0; r < link->world && r < kPeerMax; r++) pc.bufs[r] = link->bufs[r];
        pc.rank = link->rank;
        pc.world = link->world;
        pc.epoch = link->epoch;
    }
    return pc;
}

struct alignas(16) PeerRecord {
    uint w[16];
};

AFB_DI char *peer_slot(const PeerComm &c, int owner, int writer) {
    const int par = (int)(c.epoch & 1u);
    return reinterpret_cast<char *>(c.bufs[owner]) + (size_t)((par * kPeerMax + writer) * kPeerSlotBytes);
}

// ONE thread: the first NW words of this rank's record -> rank `dst`'s buffer (fire and forget)
template<int NW> AFB_DI void peer_send(const PeerComm &c, int dst, const PeerRecord &rec) {
    static_assert(NW >= 1 && NW <= 16, "a peer record carries at most 64 bytes");
    char *slot = peer_slot(c, dst, c.rank);
#pragma unroll
    for (int j = 0; j < (NW + 1) / 2; j++) {
        const uint d0 = rec.w[2 * j], d1 = (2 * j + 1 < 16) ? rec.w[2 * j + 1] : 0u;
        asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(slot + 16 * j), "r"(d0), "r"(c.epoch), "r"(d1), "r"(c.epoch)
                     : "memory");
    }
}

// ONE thread: wait for rank `src`'s record of this epoch in OUR buffer.  Ranks may legitimately arrive seconds apart
// (host-side skew); never arriving is a protocol error -> trap instead of hanging the device.
template<int NW> AFB_DI PeerRecord peer_recv(const PeerComm &c, int src) {
    const char *slot = peer_slot(c, c.rank, src);
    PeerRecord rec;
#pragma unroll
    for (int j = 0; j < 8; j++) rec.w[2 * j] = rec.w[2 * j + 1] = 0u;
#pragma unroll
    for (int j = 0; j < (NW + 1) / 2; j++) {
        uint d0, f0, d1, f1;
        unsigned long long spins = 0;
        while (true) {
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d0), "=r"(f0), "=r"(d1), "=r"(f1) : "l"(slot + 16 * j) : "memory");
            if (f0 == c.epoch && f1 == c.epoch) break;
            if (++spins > (1ull << 31)) __trap();
        }
        rec.w[2 * j] = d0;
        rec.w[2 * j + 1] = d1;
    }
    return rec;
}

// ONE thread does the whole exchange: sends to every rank (itself included), then collects ranks [0, upto) into out[]
template<int NW> AFB_DI void peer_exchange_serial(const PeerComm &c, const PeerRecord &mine, PeerRecord *out, int upto) {
    for (int p = 0; p < c.world; p++) peer_send<NW>(c, p, mine);
    for (int r = 0; r < upto; r++) out[r] = peer_recv<NW>(c, r);
}

// The first WARP of a CTA does the exchange (all 32 lanes call; `mine` is read from lane 0): lane r sends to rank r and
// collects rank r's record -- the stores and the polls of all ranks overlap.  The records land in `sm` (shared memory,
// kPeerMax entries); valid for the whole warp on return.
template<int NW> AFB_DI void peer_exchange_warp(const PeerComm &c, const PeerRecord &mine, PeerRecord *sm) {
    const int lane = (int)(threadIdx.x & 31u);
    PeerRecord rec;
#pragma unroll
    for (int j = 0; j < 16; j++) rec.w[j] = j < NW ? __shfl_sync(0xffffffffu, mine.w[j], 0) : 0u;
    if (lane < c.world) {
        peer_send<NW>(c, lane, rec);
        sm[lane] = peer_recv<NW>(c, lane);
    }
    __syncwarp();
}

// typed views of a record: a POD of at most 48 bytes at offset 0, 16 more bytes of op-specific extras at offset 48
template<typename T> AFB_DI void rec_put(PeerRecord &rec, const T &v, int byte_off = 0) {
    static_assert(sizeof(T) <= 48, "a peer record carries at most 48 bytes of payload (+ 16 of op-specific extras)");
    static_assert(sizeof(T) % 4 == 0 || sizeof(T) < 4, "record fields are whole words (or one sub-word scalar)");
    union V { T t; uint w[(sizeof(T) + 3) / 4]; AFB_DI V() {} } u;
#pragma unroll
    for (int i = 0; i < (int)((sizeof(T) + 3) / 4); i++) u.w[i] = 0u;
    u.t = v;
#pragma unroll
    for (int i = 0; i < (int)((sizeof(T) + 3) / 4); i++) rec.w[byte_off / 4 + i] = u.w[i];
}
template<typename T> AFB_DI T rec_get(const PeerRecord &rec, int byte_off = 0) {
    union V { T t; uint w[(sizeof(T) + 3) / 4]; AFB_DI V() {} } v;
#pragma unroll
    for (int i = 0; i < (int)((sizeof(T) + 3) / 4); i++) v.w[i] = rec.w[byte_off / 4 + i];
    return v.t;
}

}  // namespace afb
