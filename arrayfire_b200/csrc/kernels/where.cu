// where.cu -- stream compaction behind af_where: indices of the non-zero elements, ascending.
//
// Replaces the reference's 5-step pipeline (/root/reference/src/backend/cuda/kernel/where.hpp:25-103:
// scan_first into a full-size uint temporary, scan of block sums, D2H read of the total + sync, allocation,
// scatter kernel re-reading input and temporary: ~10 B/element of extra traffic for a u8 mask) with ONE
// kernel: each CTA takes a tile by ticket, counts its hits in registers, resolves its output offset by a
// decoupled look-back over single-word (generation | state | count) descriptors and writes its indices
// straight to their final position.  Traffic = N*sizeof(T) read + 4*count written.
// Semantics: /root/reference/src/backend/cpu/where.cpp:27-63 (x != 0, NaN counts, view-order linear index).
#include "scan_host.cuh"
#include "peer_comm.cuh"

namespace afb {

constexpr int kWhereSlots = 4;

template<typename T>
struct WhereArgs {
    const T *in;
    dim_t n;
    uint *out;
    unsigned long long *total;     // device: number of hits
    unsigned long long *descs;     // one packed word per tile
    uint *tile_counter;
    dim_t ntiles;
    uint epoch;
    uint base;                     // added to every index (sharded where)
    int vec_ok;
    void *host_wire;                  // optional pinned wire record of the count (common.hpp HostRet)
    uint host_seq;
    PeerComm comm;                    // world > 1: `in` is one shard, the counts are exchanged in the kernel
};

// the total goes to device memory and, when asked for, to the pinned host record (one thread calls this).  For a shard
// of a multi-GPU array the thread first exchanges the count with the other ranks (peer_comm.cuh): the record then also
// carries the offset of this shard's slice in the global output and the global count.
template<typename A> AFB_DI void where_publish_total(const A &a, unsigned long long total) {
    unsigned long long offset = 0, all = total;
    if (a.comm.world > 1) {
        PeerRecord rec, all_recs[kPeerMax];
#pragma unroll
        for (int j = 0; j < 16; j++) rec.w[j] = 0u;
        rec_put(rec, total);
        peer_exchange_serial<2>(a.comm, rec, all_recs, a.comm.world);
        all = 0;
        for (int r = 0; r < a.comm.world; r++) {
            const unsigned long long c = rec_get<unsigned long long>(all_recs[r]);
            if (r < a.comm.rank) offset += c;
            all += c;
        }
    }
    a.total[0] = total;
    a.total[1] = offset;
    a.total[2] = all;
    if (a.host_wire) {
        uint w[10];
        w[0] = (uint)total, w[1] = (uint)(total >> 32);
        w[2] = (uint)offset, w[3] = (uint)(offset >> 32);
        w[4] = (uint)all, w[5] = (uint)(all >> 32);
        w[6] = w[7] = w[8] = w[9] = 0u;
        host_ret_send(a.host_wire, a.host_seq, w);
    }
}

template<typename T>
__global__ void __launch_bounds__(kScanThreads) where_kernel(const WhereArgs<T> a) {
    constexpr int E = VecN<T>::value;
    constexpr int U = kWhereSlots;
    constexpr dim_t TILE = (dim_t)kScanThreads * U * E;
    __shared__ uint sm_warp[kScanWarps];
    __shared__ uint sm_prefix;
    __shared__ uint sm_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const uint t = atomicAdd(a.tile_counter, 1u);
        sm_tile = t;
        if (t == (uint)(a.ntiles - 1)) *a.tile_counter = 0u;
    }
    __syncthreads();
    const dim_t tile = sm_tile;
    const dim_t warp_lo = tile * TILE + (dim_t)warp * (U * 32 * E);

    uint mask[U], cnt[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const dim_t p0 = warp_lo + ((dim_t)u * 32 + lane) * E;
        uint m = 0;
        if (a.vec_ok && p0 + E <= a.n) {
            const Vec16 v = ldg_stream16(a.in + p0);
            const T *e = reinterpret_cast<const T *>(&v);
#pragma unroll
            for (int k = 0; k < E; k++) m |= (nonzero<T>(e[k]) ? 1u : 0u) << k;
        } else {
#pragma unroll
            for (int k = 0; k < E; k++)
                if (p0 + k < a.n) m |= (nonzero<T>(ldg_stream<T>(a.in + p0 + k)) ? 1u : 0u) << k;
        }
        mask[u] = m;
        cnt[u] = __popc(m);
    }
    // warp: exclusive offsets per slot, slots chained in order
    uint slot_excl[U], warp_total = 0;
#pragma unroll
    for (int u = 0; u < U; u++) {
        uint inc = cnt[u];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        slot_excl[u] = warp_total + inc - cnt[u];
        warp_total += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) sm_warp[warp] = warp_total;
    __syncthreads();
    uint warp_excl = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kScanWarps; w++) {
        const uint t = sm_warp[w];
        if (w < warp) warp_excl += t;
        tile_total += t;
    }
    // look-back on packed descriptors (the layout every packed user of Scratch::wtiles shares):
    // [63:32] count, [31:3] generation, [1:0] status
    const uint tag = a.epoch << 3;
    if (warp == 0) {
        uint excl = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed64(&a.descs[0], ((unsigned long long)tile_total << 32) | tag | TILE_PREFIX);
        } else {
            if (lane == 0) st_relaxed64(&a.descs[tile], ((unsigned long long)tile_total << 32) | tag | TILE_AGGREGATE);
            dim_t look = tile - 1;
            while (true) {
                const dim_t mine = look - lane;
                const bool valid = mine >= 0;
                unsigned long long d = 0;
                if (valid) {
                    { uint spins = 0; do { d = ld_relaxed64(&a.descs[mine]); if (++spins > kSpinLimit) __trap(); } while ((((uint)d) & ~7u) != tag || (d & 3u) == TILE_INVALID); }
                }
                const bool is_prefix = valid && (d & 3u) == TILE_PREFIX;
                const uint pmask = __ballot_sync(0xffffffffu, is_prefix);
                const int stop = pmask ? (__ffs(pmask) - 1) : 31;
                uint v = (valid && lane <= stop) ? (uint)(d >> 32) : 0u;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
                excl += __shfl_sync(0xffffffffu, v, 0);
                if (pmask) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed64(&a.descs[tile], ((unsigned long long)(excl + tile_total) << 32) | tag | TILE_PREFIX);
        }
        if (lane == 0) {
            sm_prefix = excl;
            if (tile == a.ntiles - 1) where_publish_total(a, (unsigned long long)excl + tile_total);
        }
    }
    __syncthreads();
    const uint before_warp = sm_prefix + warp_excl;
#pragma unroll
    for (int u = 0; u < U; u++) {
        uint m = mask[u];
        if (!m) continue;
        const dim_t p0 = warp_lo + ((dim_t)u * 32 + lane) * E;
        uint *o = a.out + before_warp + slot_excl[u];
        const uint first = a.base + (uint)p0;
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            *o++ = first + k;
        }
    }
}

template<typename T>
static void run_where_simple(const LaunchCtx &c, const View &in, uint *out, unsigned long long *d_count, uint base,
                             const HostRet *ret, const ShardLink *link) {
    WhereArgs<T> a;
    a.in = (const T *)in.ptr;
    a.n = in.dims[0] * in.dims[1] * in.dims[2] * in.dims[3];
    a.out = out;
    a.total = d_count;
    constexpr dim_t TILE = (dim_t)kScanThreads * kWhereSlots * VecN<T>::value;
    a.ntiles = (a.n + TILE - 1) / TILE;
    if (a.ntiles >= 0xffffffffLL) fail(AF_ERR_SIZE, "where: too many tiles");
    a.descs = (unsigned long long *)c.scratch->need_wtiles((size_t)a.ntiles * 8, c.stream);
    a.epoch = c.scratch->next_wepoch(c.stream);
    a.tile_counter = c.scratch->tile_counter;
    a.base = base;
    a.vec_ok = aligned_to(in.ptr, 16) ? 1 : 0;
    a.host_wire = ret ? ret->wire : nullptr;
    a.host_seq = ret ? ret->seq : 0u;
    a.comm = peer_comm_of(link);
    where_kernel<T><<<(unsigned)a.ntiles, kScanThreads, 0, c.stream>>>(a);
    note_launches(1);
    post_launch_check("where_kernel");
}

// ------------------------------------------------------------------------------------------------
// chain_where_kernel: persistent, TMA-fed compaction (same pipeline as chain_scan_kernel)
// ------------------------------------------------------------------------------------------------
// Per tile: masks of non-zero elements -> counts -> CTA offsets -> look-back on packed descriptors ->
// every warp compacts one 512-byte slot at a time into its private shared-memory strip and flushes the
// strip with fully coalesced 128-byte stores (the one-tile-per-CTA kernel above scattered 4-byte stores
// straight from registers: 8x write amplification in L2 sectors at 50 % density).
constexpr int kCW_U = 8;   // 16-byte vectors per thread per tile

template<typename T>
struct ChainWhereArgs {
    const T *in;
    dim_t n;
    uint *out;
    unsigned long long *total;
    unsigned long long *descs;     // packed: [63:32] count, [31:3] generation, [1:0] status
    dim_t ntiles;
    uint gen;
    uint base;
    uint skip;                     // pipe_where_kernel: leading elements of `in` that are not part of the array (unaligned views)
    unsigned long long *ticket;
    unsigned long long ticket_base;
    unsigned long long *debug;     // optional per-CTA cycle counters of pipe_where_kernel (AFB200_PIPE_DEBUG=1)
    void *host_wire;                  // optional pinned wire record of the count (common.hpp HostRet)
    uint host_seq;
    PeerComm comm;                    // world > 1: `in` is one shard, the counts are exchanged in the kernel
};

// bit k of the result: element k of the 16-byte vector is non-zero
template<typename T> AFB_DI uint nonzero_mask(const Vec16 &v) {
    constexpr int E = VecN<T>::value;
    if constexpr (sizeof(T) == 1) {
        // 5 instructions per 4 bytes (r01: ~9): bit 7 of every byte <- (byte != 0) [and, add, or-and]; one multiply
        // gathers the four flags into the product's top nibble (0x204081 = 2^21 + 2^14 + 2^7 + 1 moves bits 7 / 15 / 23 /
        // 31 to 28 / 29 / 30 / 31, the cross terms land on distinct lower bits, so nothing carries); a funnel shift
        // appends the nibble to the running mask
        uint m = 0;
#pragma unroll
        for (int w = 3; w >= 0; w--) {
            const uint x = v.w[w];
            const uint nz = (x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;
            m = __funnelshift_l(nz * 0x00204081u, m, 4);
        }
        return m;
    } else {
        const T *e = reinterpret_cast<const T *>(&v);
        uint m = 0;
#pragma unroll
        for (int k = 0; k < E; k++) m |= (nonzero<T>(e[k]) ? 1u : 0u) << k;
        return m;
    }
}

template<typename T>
__global__ void __launch_bounds__(kChainThreads, 2) chain_where_kernel(const ChainWhereArgs<T> a) {
    constexpr int E = VecN<T>::value;
    constexpr int U = kCW_U;
    constexpr int TILE = kChainThreads * U * E;          // elements; 32 KB of input
    constexpr size_t STAGE = (size_t)TILE * sizeof(T);
    constexpr int STRIP = 32 * E;                        // indices one warp can produce from one slot

    extern __shared__ __align__(128) unsigned char chain_smem[];   // derived pointers stay in the shared address space
    unsigned char *stage_base = chain_smem;
    uint *strips = reinterpret_cast<uint *>(stage_base + STAGE * kChainStages);   // [8][STRIP]
    __shared__ uint64_t full_bar[kChainStages];
    __shared__ long long sm_tile_id[kChainStages];
    __shared__ int sm_staged[kChainStages];
    __shared__ uint sm_warp[2][8];
    __shared__ uint sm_prefix[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kChainStages; s++) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    long long next_ticket = 0;
    bool drawing = true;
    auto publish = [&](int s, long long tile) {
        sm_tile_id[s] = tile;
        if (tile >= a.ntiles) {
            sm_staged[s] = 0;
            mbar_arrive(&full_bar[s]);
            return;
        }
        const dim_t tile_lo = tile * TILE;
        const bool whole = tile_lo + TILE <= a.n;
        sm_staged[s] = whole ? 1 : 0;
        if (whole) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&full_bar[s], (uint32_t)STAGE);
            bulk_g2s(stage_base + (size_t)s * STAGE, a.in + tile_lo, (uint32_t)STAGE, &full_bar[s]);
        } else {
            mbar_arrive(&full_bar[s]);
        }
    };
    auto draw = [&]() -> long long { return (long long)(atomicAdd(a.ticket, 1ULL) - a.ticket_base); };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kChainStages; s++) {
            if (drawing) {
                const long long t = draw();
                publish(s, t);
                if (t >= a.ntiles) drawing = false;
            }
        }
        if (drawing) next_ticket = draw();
    }

    const uint tag = a.gen << 3;
    uint *strip = strips + warp * STRIP;
    for (uint it = 0;; it++) {
        const int s = it % kChainStages;
        mbar_wait(&full_bar[s], (it / kChainStages) & 1);
        const long long tile = sm_tile_id[s];
        if (tile >= a.ntiles) break;
        const bool staged = sm_staged[s] != 0;
        const dim_t tile_lo = tile * TILE;
        const T *sin = reinterpret_cast<const T *>(stage_base + (size_t)s * STAGE);
        const int warp_off = warp * (U * 32 * E);

        // ---- masks and counts
        uint mask[U];
        uint pc[U / 2];   // two 16-bit counts per word: slot 2j in the low half, 2j+1 in the high half
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int q = warp_off + (u * 32 + lane) * E;
            uint m = 0;
            if (staged) {
                m = nonzero_mask<T>(*reinterpret_cast<const Vec16 *>(sin + q));
            } else {
                const dim_t p0 = tile_lo + q;
#pragma unroll
                for (int k = 0; k < E; k++)
                    if (p0 + k < a.n) m |= (nonzero<T>(ldg_stream<T>(a.in + p0 + k)) ? 1u : 0u) << k;
            }
            mask[u] = m;
        }
#pragma unroll
        for (int j = 0; j < U / 2; j++) pc[j] = __popc(mask[2 * j]) | (__popc(mask[2 * j + 1]) << 16);
        // ---- warp: inclusive scans of the packed counts (a slot holds at most 32*16 = 512 hits: no carry
        //      between the halves)
        uint inc[U / 2];
#pragma unroll
        for (int j = 0; j < U / 2; j++) {
            uint x = pc[j];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint o = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += o;
            }
            inc[j] = x;
        }
        uint slot_excl[U], slot_cnt[U], warp_total = 0;   // lane-exclusive offset inside the slot; hits of the slot
#pragma unroll
        for (int j = 0; j < U / 2; j++) {
            const uint tot = __shfl_sync(0xffffffffu, inc[j], 31);
            const uint ex = inc[j] - pc[j];
            slot_excl[2 * j] = ex & 0xffffu;
            slot_excl[2 * j + 1] = ex >> 16;
            slot_cnt[2 * j] = tot & 0xffffu;
            slot_cnt[2 * j + 1] = tot >> 16;
            warp_total += (tot & 0xffffu) + (tot >> 16);
        }
        const int pb = it & 1;
        if (lane == 0) sm_warp[pb][warp] = warp_total;
        __syncthreads();
        if (tid == 0 && drawing) {   // every warp holds its masks in registers: refill this stage
            publish(s, next_ticket);
            if (next_ticket >= a.ntiles) drawing = false;
            else next_ticket = draw();
        }
        uint warp_excl = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const uint t = sm_warp[pb][w];
            if (w < warp) warp_excl += t;
            tile_total += t;
        }
        // ---- look-back (warp 0)
        if (warp == 0) {
            uint excl = 0;
            if (tile == 0) {
                if (lane == 0) st_relaxed64(&a.descs[0], ((unsigned long long)tile_total << 32) | tag | TILE_PREFIX);
            } else {
                if (lane == 0) st_relaxed64(&a.descs[tile], ((unsigned long long)tile_total << 32) | tag | TILE_AGGREGATE);
                long long look = tile - 1;
                while (true) {
                    const long long mine = look - lane;
                    const bool valid = mine >= 0;
                    unsigned long long d = 0;
                    if (valid) {
                        { uint spins = 0; do { d = ld_relaxed64(&a.descs[mine]); if (++spins > kSpinLimit) __trap(); } while ((((uint)d) & ~7u) != tag || (d & 3u) == TILE_INVALID); }
                    }
                    const bool is_prefix = valid && (d & 3u) == TILE_PREFIX;
                    const uint pmask = __ballot_sync(0xffffffffu, is_prefix);
                    const int stop = pmask ? (__ffs(pmask) - 1) : 31;
                    uint v = (valid && lane <= stop) ? (uint)(d >> 32) : 0u;
                    v = __reduce_add_sync(0xffffffffu, v);
                    excl += v;
                    if (pmask) break;
                    look -= 32;
                }
                if (lane == 0) st_relaxed64(&a.descs[tile], ((unsigned long long)(excl + tile_total) << 32) | tag | TILE_PREFIX);
            }
            if (lane == 0) {
                sm_prefix[pb] = excl;
                if (tile == a.ntiles - 1) where_publish_total(a, (unsigned long long)excl + tile_total);
            }
        }
        __syncthreads();
        // ---- compact slot by slot through the warp's strip, flush coalesced
        uint gpos = sm_prefix[pb] + warp_excl;   // output position of the warp's next slot
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint cnt = slot_cnt[u];
            if (cnt == 0) continue;              // warp-uniform
            uint m = mask[u];
            uint o = slot_excl[u];
            const uint first = a.base + (uint)(tile_lo + warp_off + (u * 32 + lane) * E);
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                strip[o++] = first + k;
            }
            __syncwarp();
            for (uint i = lane; i < cnt; i += 32) a.out[gpos + i] = strip[i];
            __syncwarp();
            gpos += cnt;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pipe_where_kernel: the software-pipelined compaction (same machinery as pipe_scan_kernel, pipe_kernels.cuh)
// ------------------------------------------------------------------------------------------------
// One CTA per SM: 16 compute warps + publisher + walker + DMA warp.  A tile is kPW_Sub sub-tiles of 32 KB, each
// staged by one bulk-async copy.  Pass 1 of tile i turns every 16-byte vector into a bit mask (kept in shared
// memory: the stage itself is released for refill immediately) and counts per (sub-tile, warp); the publisher
// makes the tile count visible to the other CTAs; the walker keeps the running count of ALL earlier tiles
// (walk_fold: coalesced reads of published counts, no prefix chain); pass 2 of tile i-1 -- executed after pass 1
// of tile i, so the walker's L2 round trip (2000-4000 cycles under load) is hidden -- compacts 512 bytes of input
// at a time through a per-warp strip and flushes it with coalesced stores.  Tiles are large (128 KB of input)
// because one walker step per tile must fit into the tile period even for a sparse mask.
#ifndef AFB_PW_WARPS
#define AFB_PW_WARPS 16
#define AFB_PW_SUB 4
#define AFB_PW_STAGES 4
#endif
constexpr int kPW_Warps = AFB_PW_WARPS;
constexpr int kPW_U = 4;        // 512-byte slots per warp per sub-tile
constexpr int kPW_Sub = AFB_PW_SUB;      // sub-tiles (bulk copies) per tile
constexpr int kPW_Stages = AFB_PW_STAGES;
constexpr int kPW_Ring = 4;     // per-tile barrier / scratch ring
constexpr int kPW_Threads = 32 * (kPW_Warps + 3);

template<typename T> struct PipeWhereLayout {
    static constexpr int E = VecN<T>::value;
    static constexpr int NVEC = kPW_Warps * kPW_U * 32;          // 16-byte vectors per sub-tile
    static constexpr int SUBTILE = NVEC * E;                     // elements per sub-tile
    static constexpr int TILE = SUBTILE * kPW_Sub;               // elements per tile
    static constexpr size_t stage_bytes = (size_t)NVEC * 16;     // 32 KB
    static constexpr int STRIP = 32 * E + 4;                     // + room for the output alignment shift
    static constexpr size_t strip_bytes = (size_t)kPW_Warps * STRIP * sizeof(uint);
    static constexpr size_t mask_bytes = (size_t)2 * kPW_Sub * NVEC * sizeof(unsigned short);
    static constexpr size_t total = stage_bytes * kPW_Stages + strip_bytes + mask_bytes + 128;
};

template<typename T>
__global__ void __launch_bounds__(kPW_Threads, 1) pipe_where_kernel(const ChainWhereArgs<T> a) {
    typedef PipeWhereLayout<T> Lay;
    typedef Rule<OP_ADD, uint, uint> R;
    typedef SOps<R, uint, false> S;
    typedef PipeCodec<S, uint, false> Codec;
    constexpr int E = Lay::E;
    constexpr int TILE = Lay::TILE, SUBTILE = Lay::SUBTILE;
    constexpr int U = kPW_U, NW = kPW_Warps, ST = kPW_Stages, SUB = kPW_Sub, RING = kPW_Ring;
    constexpr int NE = 64;           // (sub-tile, warp) count entries per tile, two per publisher / walker lane; unused ones stay 0
    static_assert(SUB * NW <= NE, "publisher / walker handle two entries per lane");

    extern __shared__ __align__(128) unsigned char pipe_smem[];   // see pipe_scan_kernel: keeps the accesses LDS/STS
    unsigned char *stage_base = pipe_smem;
    uint *strips = reinterpret_cast<uint *>(stage_base + Lay::stage_bytes * ST);
    unsigned short *masks = reinterpret_cast<unsigned short *>(stage_base + Lay::stage_bytes * ST + Lay::strip_bytes);
    __shared__ uint64_t full_bar[ST], empty_bar[ST];
    __shared__ uint64_t p1_bar[RING], pub_bar[RING], lb_bar[RING];
    __shared__ long long sm_stage_tile[ST];     // tile of the sub-tile staged here (or an end marker)
    __shared__ int sm_stage_valid[ST];          // valid elements of that sub-tile; SUBTILE: staged by a bulk copy
    __shared__ long long sm_tile_of[RING];      // tile handled in this ring slot (written by pass 1)
    __shared__ uint sm_wtot[RING][NE];          // hits per (sub-tile, warp)
    __shared__ uint sm_wexc[RING][NE];          // exclusive offsets inside the tile
    __shared__ uint sm_total[RING];
    __shared__ uint sm_wp[RING][NE];            // absolute output offsets

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (SUB * NW < NE)
        for (int e = tid; e < RING * NE; e += kPW_Threads) (&sm_wtot[0][0])[e] = 0u;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < ST; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NW);
        }
#pragma unroll
        for (int b = 0; b < RING; b++) {
            mbar_init(&p1_bar[b], NW);
            mbar_init(&pub_bar[b], 1);
            mbar_init(&lb_bar[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint tag = a.gen << 3;

    if (warp == NW + 2) {
        // =========================================================================== DMA warp (one thread)
        if (lane != 0) return;
        // raw counter values; ticket_base is subtracted at the point of use so that nothing waits on the atomic itself
        auto draw = [&]() -> unsigned long long { return atomicAdd(a.ticket, 1ULL); };
        long long m = 0;                       // sub-tiles published so far; sub-tile m lives in stage m % ST
        unsigned long long tk0 = draw(), tk1 = draw();  // tickets are drawn two tiles ahead of their use
        while (true) {
            const long long tile = (long long)(tk0 - a.ticket_base);
            tk0 = tk1;
            const bool live = tile < a.ntiles;
            if (live) tk1 = draw();
            for (int sub = 0; sub < (live ? SUB : 1); sub++, m++) {
                const int s = (int)(m % ST);
                if (m >= ST) mbar_wait(&empty_bar[s], (uint32_t)(((m / ST) - 1) & 1));   // pass 1 is done with the stage
                sm_stage_tile[s] = tile;
                if (!live) {
                    sm_stage_valid[s] = 0;
                    mbar_arrive(&full_bar[s]);
                    break;
                }
                const dim_t lo = tile * TILE + (dim_t)sub * SUBTILE;
                const dim_t left = a.n - lo;
                const int valid = left >= SUBTILE ? SUBTILE : (left > 0 ? (int)left : 0);
                sm_stage_valid[s] = valid;
                if (valid == SUBTILE) {
                    fence_proxy_async();
                    mbar_arrive_expect_tx(&full_bar[s], (uint32_t)Lay::stage_bytes);
                    bulk_g2s(stage_base + (size_t)s * Lay::stage_bytes, a.in + lo, (uint32_t)Lay::stage_bytes, &full_bar[s]);
                } else {
                    mbar_arrive(&full_bar[s]);   // ragged / empty sub-tile: filled (zero-padded) by the compute warps
                }
            }
            if (!live) break;
        }
        return;
    }

    if (warp == NW) {
        // =========================================================================== publisher warp
        for (long long j = 0;; j++) {
            const int slot = (int)(j % RING);
            mbar_wait(&p1_bar[slot], (uint32_t)((j / RING) & 1));
            const long long tile = sm_tile_of[slot];
            if (tile < a.ntiles) {
                const uint v0 = sm_wtot[slot][2 * lane], v1 = sm_wtot[slot][2 * lane + 1];
                uint inc = v0 + v1;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const uint o = __shfl_up_sync(0xffffffffu, inc, dd);
                    if (lane >= dd) inc += o;
                }
                const uint total = __shfl_sync(0xffffffffu, inc, 31);
                if (lane == 0) {
                    unsigned long long w[1];
                    Codec::pack(total, tag, TILE_AGGREGATE, w);
                    Codec::store(a.descs, tile, w);
                }
                sm_wexc[slot][2 * lane] = inc - v0 - v1;
                sm_wexc[slot][2 * lane + 1] = inc - v1;
                if (lane == 0) sm_total[slot] = total;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pub_bar[slot]);
            if (tile >= a.ntiles) break;
        }
        return;
    }

    if (warp == NW + 1) {
        // =========================================================================== walker warp
        long long pos = 0;      // R = hits in the tiles [0, pos)
        uint R = 0u;
        for (long long j = 0;; j++) {
            const int slot = (int)(j % RING);
            mbar_wait(&pub_bar[slot], (uint32_t)((j / RING) & 1));
            const long long tile = sm_tile_of[slot];
            if (tile >= a.ntiles) break;
            R = walk_fold<S, Codec>(a.descs, pos, tile, tag, lane, R);
            sm_wp[slot][2 * lane] = R + sm_wexc[slot][2 * lane];
            sm_wp[slot][2 * lane + 1] = R + sm_wexc[slot][2 * lane + 1];
            const uint total = sm_total[slot];
            if (lane == 0 && tile == a.ntiles - 1) where_publish_total(a, (unsigned long long)R + total);
            __syncwarp();
            if (lane == 0) mbar_arrive(&lb_bar[slot]);
            R += total;
            pos = tile + 1;
        }
        return;
    }

    // =============================================================================== compute warps
    uint *strip = strips + warp * Lay::STRIP;
    long long m = 0;            // sub-tile counter (matches the DMA warp's)
    long long tile_prev = 0;
    bool ended = false;
    unsigned long long cw[4] = {0, 0, 0, 0};   // wait full, pass 1, wait walker, pass 2
    const long long cw_start = clock64();
    for (long long i = 0;; i++) {
        const int b = (int)(i % RING);
        // ---------------------------------------------------------------- pass 1 of tile i: masks + counts
        long long tile = 0;
        unsigned short *mk = masks + (size_t)(i & 1) * (SUB * Lay::NVEC);
        uint cnts[SUB];
#pragma unroll 1
        for (int sub = 0; sub < SUB; sub++) {
            const int s = (int)(m % ST);
            long long k0 = clock64();
            mbar_wait(&full_bar[s], (uint32_t)((m / ST) & 1));
            cw[0] += clock64() - k0;
            k0 = clock64();
            if (sub == 0) {
                tile = sm_stage_tile[s];
                if (tile >= a.ntiles) {
                    ended = true;
                    break;
                }
            }
            unsigned char *stg = stage_base + (size_t)s * Lay::stage_bytes;
            const int valid = sm_stage_valid[s];
            if (valid != SUBTILE) {
                T *dst = reinterpret_cast<T *>(stg);
                const dim_t lo = tile * TILE + (dim_t)sub * SUBTILE;
                for (int e = warp * 32 + lane; e < SUBTILE; e += NW * 32) {
                    if (e < valid) dst[e] = ldg_stream<T>(a.in + lo + e);
                    else memset(&dst[e], 0, sizeof(T));
                }
                named_bar_sync(1, NW * 32);
            }
            uint cnt = 0;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int q = (warp * U + u) * 32 + lane;
                const Vec16 v = *reinterpret_cast<const Vec16 *>(stg + (size_t)q * 16);
                uint mm = nonzero_mask<T>(v);
                if (q == 0 && sub == 0 && tile == 0) mm &= ~((1u << a.skip) - 1u);   // unaligned view: the elements in front of it
                mk[sub * Lay::NVEC + q] = (unsigned short)mm;
                cnt += __popc(mm);
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            cnts[sub] = cnt;
            if (lane == 0) {
                sm_wtot[b][sub * NW + warp] = cnt;
                mbar_arrive(&empty_bar[s]);   // the stage may be refilled (the reduce above ordered every lane's reads)
            }
            cw[1] += clock64() - k0;
            m++;
        }
        if (lane == 0 && warp == 0) sm_tile_of[b] = ended ? a.ntiles : tile;
        __syncwarp();
        if (lane == 0) mbar_arrive(&p1_bar[b]);
        if (i >= 1) {
            // ---------------------------------------------------------------- pass 2 of tile i-1: compaction
            const int bp = (int)((i - 1) % RING);
            long long k1 = clock64();
            mbar_wait(&lb_bar[bp], (uint32_t)(((i - 1) / RING) & 1));
            cw[2] += clock64() - k1;
            k1 = clock64();
            const unsigned short *mkp = masks + (size_t)((i - 1) & 1) * (SUB * Lay::NVEC);
            // flush `total` entries of the warp's strip (filled from index `shift` = gpos & 3 on, so that its 16-byte quads
            // line up with `out`); the strip holds offsets relative to `add`.  Whole quads: one predicate, one 16-byte
            // shared-memory load and one 16-byte store each, addressed by immediates off two per-lane bases; the ragged head
            // (shift != 0) and tail (lim & 3) element-wise outside the unrolled loop.  (r02a tested every quad against both
            // ends inside the loop: ~75 instructions per flush of 256 indices, as many as the expansion itself.)
            // REL: strip entries are offsets relative to `add` (E < 16: saves the per-element add of the expansion; with the 16
            // elements per vector of 1-byte types the registers that takes cost more than the adds -- measured)
            constexpr bool REL = E < 16;
            auto flush = [&](uint gpos, uint shift, uint total, uint add) {
                if (!REL) add = 0u;
                if (total <= 32u) {                  // one coalesced scalar store
                    if (lane < total) a.out[gpos + lane] = strip[shift + lane] + add;
                    return;
                }
                uint *out_al = a.out + (gpos - shift);
                const uint lim = shift + total;
                constexpr int kQ = (32 * E + 4 + 127) / 128;
                const uint4 *sq = reinterpret_cast<const uint4 *>(strip) + lane;
                uint4 *oq = reinterpret_cast<uint4 *>(out_al) + lane;
                const int lim_l = (int)lim - 4 * lane;       // quad lane + 32 j is whole iff 128 j + 4 <= lim_l
                uint4 v[kQ];
                bool whole[kQ];
#pragma unroll
                for (int j = 0; j < kQ; j++) {
                    whole[j] = 128 * j + 4 <= lim_l;
                    if (j == 0) whole[j] = whole[j] && (lane > 0 || shift == 0u);
                    if (whole[j]) v[j] = sq[32 * j];
                }
                asm volatile("" ::: "memory");       // every shared-memory load is issued before the first store
#pragma unroll
                for (int j = 0; j < kQ; j++) {
                    if (whole[j]) {
                        if (REL) v[j].x += add, v[j].y += add, v[j].z += add, v[j].w += add;
                        oq[32 * j] = v[j];
                    }
                }
                if (shift != 0u && lane >= shift && lane < 4u) out_al[lane] = strip[lane] + add;     // total > 32: lim > 4
                const uint tail = lim & ~3u;
                if (tail + lane < lim) out_al[tail + lane] = strip[tail + lane] + add;
            };
            // strip entries are offsets inside the warp's slot / group: lane * E + k sits in registers for the whole
            // kernel (opaque to the compiler, so that it keeps them instead of re-deriving one per element)
            uint rk[E];
#pragma unroll
            for (int k = 0; k < E; k++) {
                rk[k] = (uint)(lane * E + k);
                if (REL) asm volatile("" : "+r"(rk[k]));
            }
#pragma unroll 1
            for (int sub = 0; sub < SUB; sub++) {
                uint gpos = sm_wp[bp][sub * NW + warp];
                const uint gtotal = sm_wtot[bp][sub * NW + warp];   // hits of this warp's U slots (pass 1)
                if (gtotal == 0u) continue;                         // warp-uniform
                const dim_t sub_lo = tile_prev * TILE + (dim_t)sub * SUBTILE;
                if (gtotal <= 32u * E) {
                    // ---- sparse group (the warp's U x 512 input bytes hold at most one strip of hits): a lane owns U
                    // CONSECUTIVE vectors (one 8-byte load of their masks), so its hits are contiguous in the output and
                    // ONE warp scan + one flush serve all U slots (r01: one scan + flush per slot; 1 % mask 311 M -> ~110 M
                    // warp instructions per 2^30 bytes)
                    static_assert(U == 4, "the sparse path reads the masks of 4 vectors as one 8-byte word");
                    const int q0 = warp * U * 32 + lane * U;
                    const uint2 mw = *reinterpret_cast<const uint2 *>(mkp + sub * Lay::NVEC + q0);
                    const uint c = __popc(mw.x) + __popc(mw.y);
                    uint inc = c;
#pragma unroll
                    for (int dd = 1; dd < 32; dd <<= 1) {
                        const uint o = __shfl_up_sync(0xffffffffu, inc, dd);
                        if (lane >= dd) inc += o;
                    }
                    const uint shift = gpos & 3u;
                    uint *p = strip + (shift + inc - c);
                    const uint gfirst = a.base + (uint)(sub_lo + (dim_t)(warp * U * 32) * E);
                    const uint first = (uint)(lane * U * E) + (REL ? 0u : gfirst);   // REL: relative to the group's first element
                    uint m = mw.x;
                    while (m) {                      // vectors 0 (bits 0..15) and 1 (bits 16..31)
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        *p++ = first + (uint)((b >> 4) * E + (b & 15));
                    }
                    m = mw.y;
                    while (m) {                      // vectors 2 and 3
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        *p++ = first + (uint)(2 * E + (b >> 4) * E + (b & 15));
                    }
                    __syncwarp();
                    flush(gpos, shift, gtotal, gfirst);
                    __syncwarp();
                    continue;
                }
                // ---- dense group: slot by slot, a lane owns ONE vector per slot.  The U masks are fetched and their U warp
                // scans run interleaved up front (independent shuffle chains), so only expansion + flush stay serial per slot
                uint mm[U], cc[U], inc[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    mm[u] = mkp[sub * Lay::NVEC + (warp * U + u) * 32 + lane];
                    cc[u] = __popc(mm[u]);
                    inc[u] = cc[u];
                }
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const uint o = __shfl_up_sync(0xffffffffu, inc[u], dd);
                        if (lane >= dd) inc[u] += o;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint total = __shfl_sync(0xffffffffu, inc[u], 31);
                    if (total == 0u) continue;        // warp-uniform
                    const uint shift = gpos & 3u;
                    // branch-free expansion through running pointers (one per half vector: two short dependency chains):
                    // per element one predicated store and one predicated pointer bump (r01 recomputed the strip address
                    // from the offset for every element)
                    const uint m = mm[u];
                    const uint sfirst = a.base + (uint)(sub_lo + (dim_t)((warp * U + u) * 32) * E);
                    const uint voff = REL ? 0u : sfirst;
                    uint *p0 = strip + (shift + inc[u] - cc[u]);
                    constexpr int H = E >= 2 ? E / 2 : 1;
                    uint *p1 = p0 + __popc(m & ((1u << H) - 1u));
#pragma unroll
                    for (int k = 0; k < H; k++) {
                        const bool hit = (m >> k) & 1u;
                        if (hit) *p0 = rk[k] + voff;
                        if (hit) p0++;
                    }
#pragma unroll
                    for (int k = H; k < E; k++) {
                        const bool hit = (m >> k) & 1u;
                        if (hit) *p1 = rk[k] + voff;
                        if (hit) p1++;
                    }
                    __syncwarp();
                    flush(gpos, shift, total, sfirst);
                    __syncwarp();
                    gpos += total;
                }
            }
            cw[3] += clock64() - k1;
        }
        if (ended) break;
        tile_prev = tile;
    }
    if (a.debug && warp == 0 && lane == 0) {
        for (int q = 0; q < 4; q++) a.debug[blockIdx.x * 8 + q] = cw[q];
        a.debug[blockIdx.x * 8 + 4] = clock64() - cw_start;
    }
}

template<typename T>
static void run_where(const LaunchCtx &c, const View &in, uint *out, unsigned long long *d_count, uint base,
                      const HostRet *ret, const ShardLink *link) {
    const dim_t n = in.dims[0] * in.dims[1] * in.dims[2] * in.dims[3];
    constexpr dim_t CT = (dim_t)kChainThreads * kCW_U * VecN<T>::value;
    // an input that is not 16-byte aligned (a view with an offset) is read from the aligned address below it: the few
    // elements in front belong to the same allocation (allocations are 256-byte aligned), their mask bits are cleared in
    // the kernel and the indices are shifted back (r02a sent such views to the 1.2 TB/s first-generation kernel)
    const uint skip = (uint)((reinterpret_cast<uintptr_t>(in.ptr) & 15) / sizeof(T));
    const bool elem_aligned = (reinterpret_cast<uintptr_t>(in.ptr) % sizeof(T)) == 0;
    if (elem_aligned && aligned_to(out, 16) && n >= 2 * (dim_t)PipeWhereLayout<T>::TILE && !env_int("AFB200_NO_PIPE", 0)) {
        typedef PipeWhereLayout<T> PL;
        auto kern = pipe_where_kernel<T>;
        static bool configured[64] = {false};
        const int dv = c.device & 63;
        if (!configured[dv]) {
            AFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PL::total));
            configured[dv] = true;
        }
        ChainWhereArgs<T> a;
        a.debug = nullptr;
        a.in = (const T *)in.ptr - skip;
        a.n = n + skip;
        a.skip = skip;
        a.out = out;
        a.total = d_count;
        a.ntiles = (a.n + PL::TILE - 1) / PL::TILE;
        a.descs = (unsigned long long *)c.scratch->need_wtiles((size_t)a.ntiles * 8, c.stream);
        a.gen = c.scratch->next_wepoch(c.stream);
        a.base = base - skip;
        a.host_wire = ret ? ret->wire : nullptr;
        a.host_seq = ret ? ret->seq : 0u;
        a.comm = peer_comm_of(link);
        dim_t grid = (dim_t)c.sms;
        if (grid > a.ntiles) grid = a.ntiles;
        a.ticket = c.scratch->chain_ticket;
        a.ticket_base = c.scratch->chain_base;
        c.scratch->chain_base += (unsigned long long)a.ntiles + 2ULL * (unsigned long long)grid;  // two tickets in flight per CTA
        static unsigned long long *dbg_buf = nullptr;
        const bool dbg = env_int("AFB200_PIPE_DEBUG", 0) != 0;
        if (dbg && !dbg_buf) AFB_CUDA(cudaMalloc(&dbg_buf, 1024 * 8 * 8));
        a.debug = dbg ? dbg_buf : nullptr;
        kern<<<(unsigned)grid, kPW_Threads, PL::total, c.stream>>>(a);
        note_launches(1);
        post_launch_check("pipe_where_kernel");
        if (dbg) {
            static unsigned long long h[1024 * 8];
            AFB_CUDA(cudaStreamSynchronize(c.stream));
            AFB_CUDA(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
            const char *names[5] = {"cmp_wait_full", "cmp_p1", "cmp_wait_walker", "cmp_p2", "total"};
            fprintf(stderr, "[pipe_where debug] grid=%d ntiles=%lld\n", (int)grid, (long long)a.ntiles);
            for (int q = 0; q < 5; q++) {
                double sum = 0;
                for (int b = 0; b < (int)grid; b++) sum += (double)h[b * 8 + q];
                fprintf(stderr, "   %-16s mean %.0f\n", names[q], sum / (double)grid);
            }
        }
        return;
    }
    if (aligned_to(in.ptr, 16) && n >= CT && !env_int("AFB200_NO_CHAIN", 0)) {
        constexpr size_t smem = (size_t)CT * sizeof(T) * kChainStages + 8 * 32 * VecN<T>::value * sizeof(uint) + 128;
        auto kern = chain_where_kernel<T>;
        static int resident[64] = {0};
        const int dv = c.device & 63;
        if (!resident[dv]) {
            AFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int nb = 0;
            AFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kChainThreads, smem));
            resident[dv] = nb > 0 ? nb : 1;
        }
        ChainWhereArgs<T> a;
        a.debug = nullptr;
        a.in = (const T *)in.ptr;
        a.n = n;
        a.out = out;
        a.total = d_count;
        a.ntiles = (n + CT - 1) / CT;
        a.descs = (unsigned long long *)c.scratch->need_wtiles((size_t)a.ntiles * 8, c.stream);
        a.gen = c.scratch->next_wepoch(c.stream);
        a.base = base;
        a.skip = 0u;
        a.host_wire = ret ? ret->wire : nullptr;
        a.host_seq = ret ? ret->seq : 0u;
        a.comm = peer_comm_of(link);
        int ctas = env_int("AFB200_CHAIN_CTAS_PER_SM", resident[dv]);
        if (ctas > resident[dv]) ctas = resident[dv];
        dim_t grid = (dim_t)c.sms * ctas;
        if (grid > a.ntiles) grid = a.ntiles;
        a.ticket = c.scratch->chain_ticket;
        a.ticket_base = c.scratch->chain_base;
        c.scratch->chain_base += (unsigned long long)a.ntiles + (unsigned long long)grid;
        kern<<<(unsigned)grid, kChainThreads, smem, c.stream>>>(a);
        note_launches(1);
        post_launch_check("chain_where_kernel");
        return;
    }
    run_where_simple<T>(c, in, out, d_count, base, ret, link);
}

bool launch_where(const LaunchCtx &c, af_dtype t, const View &in, unsigned *out, unsigned long long *d_count,
                  unsigned base, const HostRet *ret, const ShardLink *link) {
    switch (t) {
        case f32: run_where<float>(c, in, out, d_count, base, ret, link); return true;
        case f64: run_where<double>(c, in, out, d_count, base, ret, link); return true;
        case c32: run_where<cfloat>(c, in, out, d_count, base, ret, link); return true;
        case c64: run_where<cdouble>(c, in, out, d_count, base, ret, link); return true;
        case s32: run_where<int>(c, in, out, d_count, base, ret, link); return true;
        case u32: run_where<uint>(c, in, out, d_count, base, ret, link); return true;
        case s64: run_where<intl>(c, in, out, d_count, base, ret, link); return true;
        case u64: run_where<uintl>(c, in, out, d_count, base, ret, link); return true;
        case s16: run_where<short>(c, in, out, d_count, base, ret, link); return true;
        case u16: run_where<ushort>(c, in, out, d_count, base, ret, link); return true;
        case s8: run_where<schar>(c, in, out, d_count, base, ret, link); return true;
        case u8: run_where<uchar>(c, in, out, d_count, base, ret, link); return true;
        case b8: run_where<b8_t>(c, in, out, d_count, base, ret, link); return true;
        default: return false;
    }
}

}  // namespace afb
