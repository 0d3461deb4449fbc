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
    // look-back on packed descriptors: [63:34] generation, [33:32] state, [31:0] count
    const unsigned long long tag = (unsigned long long)a.epoch << 34;
    if (warp == 0) {
        uint excl = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed64(&a.descs[0], tag | ((unsigned long long)TILE_PREFIX << 32) | tile_total);
        } else {
            if (lane == 0) st_relaxed64(&a.descs[tile], tag | ((unsigned long long)TILE_AGGREGATE << 32) | tile_total);
            dim_t look = tile - 1;
            while (true) {
                const dim_t mine = look - lane;
                const bool valid = mine >= 0;
                unsigned long long d = 0;
                if (valid) {
                    do { d = ld_relaxed64(&a.descs[mine]); } while ((d >> 34) != a.epoch || ((d >> 32) & 3u) == TILE_INVALID);
                }
                const bool is_prefix = valid && ((d >> 32) & 3u) == TILE_PREFIX;
                const uint pmask = __ballot_sync(0xffffffffu, is_prefix);
                const int stop = pmask ? (__ffs(pmask) - 1) : 31;
                uint v = (valid && lane <= stop) ? (uint)d : 0u;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
                excl += __shfl_sync(0xffffffffu, v, 0);
                if (pmask) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed64(&a.descs[tile], tag | ((unsigned long long)TILE_PREFIX << 32) | (excl + tile_total));
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
static void run_where(const LaunchCtx &c, const View &in, uint *out, unsigned long long *d_count, uint base) {
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
