// peer_comm.cuh -- the cross-GPU combine of the sharded whole-array ops, done INSIDE the kernel that produced the shard's
// partial (the reference is single-device).
//
// Round 1 ended every sharded op with a second launch (peer_exchange_kernel) + a host spin + a fold in Python: a fixed
// ~50 us per op, more than the kernels themselves once 2^30 elements are split over 8 GPUs (3.65x at 8 GPUs).  Now the
// LAST thread of the producing kernel (the last-arriver CTA of span_kernel, the walker of the last tile of the pipelined
// scan / compaction kernels) does the exchange itself over NVLink / NVSwitch peer memory:
//   1. stores its 64-byte record into slot [rank] of EVERY rank's symmetric buffer (four 16-byte peer stores),
//   2. raises flag [rank] = epoch there (release, system scope -- orders the record before the flag),
//   3. waits until the `world` flags in its OWN buffer carry the epoch (acquire, system scope),
//   4. folds the records in RANK ORDER (deterministic, identical on every rank, equal to the serial answer) and hands the
//      result to the device output and / or the pinned host record of the call.
// One launch, no collective-library call, no host round trip between partial and combine.  Slots and flags are
// double-buffered by epoch parity: a rank reaches epoch e+2 only after every peer has published e+1, i.e. consumed e.
// The buffers are peer-mapped once by the host layer (torch.distributed._symmetric_memory or cuMem + IPC handles).
#pragma once

#include "device_ops.cuh"

namespace afb {

constexpr int kPeerMax = 8;
constexpr int kPeerRecord = 64;                                  // bytes per rank and parity
constexpr int kPeerFlagOff = 2 * kPeerMax * kPeerRecord;         // flags follow the 2 x 8 records
constexpr int kPeerBufBytes = kPeerFlagOff + 2 * kPeerMax * 4;   // 1088

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
    uint4 q[4];
};

AFB_DI void st_peer16(void *dst, const uint4 &v) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
AFB_DI uint4 ld_peer16(const void *src) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
    return v;
}

// ONE thread: publish this rank's record to every rank (itself included) and raise the flags
AFB_DI void peer_publish(const PeerComm &c, const PeerRecord &rec) {
    const int par = (int)(c.epoch & 1u);
    for (int p = 0; p < c.world; p++) {
        char *slot = reinterpret_cast<char *>(c.bufs[p]) + (size_t)((par * kPeerMax + c.rank) * kPeerRecord);
#pragma unroll
        for (int j = 0; j < 4; j++) st_peer16(slot + 16 * j, rec.q[j]);
    }
    for (int p = 0; p < c.world; p++) {
        unsigned *f = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(c.bufs[p]) + kPeerFlagOff) + par * kPeerMax + c.rank;
        st_release_sys(f, c.epoch);   // release: this thread's record stores above are visible before the flag
    }
}

// ONE thread: wait until ranks [0, upto) have published this epoch into OUR buffer.  Ranks may legitimately arrive
// seconds apart (host-side skew); never arriving is a protocol error -> trap instead of hanging the device.
AFB_DI void peer_wait(const PeerComm &c, int upto) {
    const int par = (int)(c.epoch & 1u);
    const unsigned *flags = reinterpret_cast<const unsigned *>(reinterpret_cast<const char *>(c.bufs[c.rank]) + kPeerFlagOff) + par * kPeerMax;
    for (int r = 0; r < upto; r++) {
        unsigned long long spins = 0;
        while (ld_acquire_sys(flags + r) != c.epoch) {
            if (++spins > (1ull << 31)) __trap();
        }
    }
}

// record of rank r in OUR buffer (valid after peer_wait covered r)
AFB_DI PeerRecord peer_read(const PeerComm &c, int r) {
    const int par = (int)(c.epoch & 1u);
    const char *slot = reinterpret_cast<const char *>(c.bufs[c.rank]) + (size_t)((par * kPeerMax + r) * kPeerRecord);
    PeerRecord rec;
#pragma unroll
    for (int j = 0; j < 4; j++) rec.q[j] = ld_peer16(slot + 16 * j);
    return rec;
}

// typed views of a record: a POD of at most 48 bytes at offset 0, 16 more bytes of op-specific extras at offset 48
template<typename T> AFB_DI void rec_put(PeerRecord &rec, const T &v, int byte_off = 0) {
    static_assert(sizeof(T) <= 48, "a peer record carries at most 48 bytes of payload");
    union U { PeerRecord r; char b[kPeerRecord]; AFB_DI U() {} } u;
    u.r = rec;
    const char *s = reinterpret_cast<const char *>(&v);
    for (int i = 0; i < (int)sizeof(T); i++) u.b[byte_off + i] = s[i];
    rec = u.r;
}
template<typename T> AFB_DI T rec_get(const PeerRecord &rec, int byte_off = 0) {
    union U { PeerRecord r; char b[kPeerRecord]; AFB_DI U() {} } u;
    u.r = rec;
    union V { T t; char b[sizeof(T)]; AFB_DI V() {} } v;
    for (int i = 0; i < (int)sizeof(T); i++) v.b[i] = u.b[byte_off + i];
    return v.t;
}

}  // namespace afb
