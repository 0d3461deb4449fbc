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
};

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
            if (tile == a.ntiles - 1) *a.total = (unsigned long long)excl + tile_total;
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
static void run_where_simple(const LaunchCtx &c, const View &in, uint *out, unsigned long long *d_count, uint base) {
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
    unsigned long long *ticket;
    unsigned long long ticket_base;
};

// bit k of the result: element k of the 16-byte vector is non-zero
template<typename T> AFB_DI uint nonzero_mask(const Vec16 &v) {
    constexpr int E = VecN<T>::value;
    if constexpr (sizeof(T) == 1) {
        uint m = 0;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const uint x = v.w[w];
            // bit 7 of every byte <- (byte != 0), then gather the four bits into the top nibble
            const uint nz = ((x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u) >> 7;
            m |= ((nz * 0x10204080u) >> 28) << (4 * w);
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

    extern __shared__ unsigned char smem_raw[];
    unsigned char *stage_base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
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
                if (tile == a.ntiles - 1) *a.total = (unsigned long long)excl + tile_total;
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

template<typename T>
static void run_where(const LaunchCtx &c, const View &in, uint *out, unsigned long long *d_count, uint base) {
    const dim_t n = in.dims[0] * in.dims[1] * in.dims[2] * in.dims[3];
    constexpr dim_t CT = (dim_t)kChainThreads * kCW_U * VecN<T>::value;
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
        a.in = (const T *)in.ptr;
        a.n = n;
        a.out = out;
        a.total = d_count;
        a.ntiles = (n + CT - 1) / CT;
        a.descs = (unsigned long long *)c.scratch->need_wtiles((size_t)a.ntiles * 8, c.stream);
        a.gen = c.scratch->next_wepoch(c.stream);
        a.base = base;
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
    run_where_simple<T>(c, in, out, d_count, base);
}

bool launch_where(const LaunchCtx &c, af_dtype t, const View &in, unsigned *out, unsigned long long *d_count,
                  unsigned base) {
    switch (t) {
        case f32: run_where<float>(c, in, out, d_count, base); return true;
        case f64: run_where<double>(c, in, out, d_count, base); return true;
        case c32: run_where<cfloat>(c, in, out, d_count, base); return true;
        case c64: run_where<cdouble>(c, in, out, d_count, base); return true;
        case s32: run_where<int>(c, in, out, d_count, base); return true;
        case u32: run_where<uint>(c, in, out, d_count, base); return true;
        case s64: run_where<intl>(c, in, out, d_count, base); return true;
        case u64: run_where<uintl>(c, in, out, d_count, base); return true;
        case s16: run_where<short>(c, in, out, d_count, base); return true;
        case u16: run_where<ushort>(c, in, out, d_count, base); return true;
        case s8: run_where<schar>(c, in, out, d_count, base); return true;
        case u8: run_where<uchar>(c, in, out, d_count, base); return true;
        case b8: run_where<b8_t>(c, in, out, d_count, base); return true;
        default: return false;
    }
}

}  // namespace afb
