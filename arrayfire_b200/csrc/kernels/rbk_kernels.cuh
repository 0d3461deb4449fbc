// rbk_kernels.cuh -- reductions BY KEY (af_sum_by_key, af_min_by_key, ... SURVEY.md section 8(f)-1).
//
// Semantics (/root/reference/src/backend/cpu/kernel/reduce.hpp:64-160, src/backend/cpu/reduce.cpp:73-108): `keys` is a
// vector of n 32-bit keys, runs of EQUAL CONSECUTIVE keys are segments (keys need not be sorted); every line of `vals`
// along `dim` (n elements) is reduced per segment; keys_out holds one key per run, vals_out has dims[dim] = #runs.
//
// What it replaces: the reference CUDA backend loops on the HOST -- per round a block reduce, a cub::DeviceScan, a
// compaction and a D2H read of a "needs another round" flag (src/backend/cuda/reduce_impl.hpp:88-180,
// kernel/reduce_by_key.hpp:33-621).  Here:
//   pass 1  rbk_count_kernel   one streaming pass over the KEYS: runs per 2048-key tile AND the run-head bit of every key
//                              (one bit per key, n / 8 bytes: pass 2 never reads the keys again); the last CTA to arrive
//                              (ticket) turns the counts into exclusive tile bases and publishes the total to a pinned host
//                              record (the host needs it to size the outputs -- the one unavoidable host round trip).
//   pass 2  rbk_tile_kernel    dim 0: one CTA per (tile, column): head bits (1 byte per thread) + blocked value loads, a
//                              segmented scan in registers / shuffles,
//                              every run TAIL writes its reduction straight to out[tile_base + local run index]; heads
//                              write keys_out.  Only the run that enters a tile from the left is incomplete: its partial is
//                              written anyway and the tile's (value, has-head) aggregate goes to a side array ...
//   pass 3  rbk_carry_*        ... which two small launches scan (segmented, one thread per tile) to add the carried-in
//                              part to the first output of each tile.  Deterministic: no atomics on values, fixed order.
//           rbk_dim_kernel     dim > 0: lanes along the contiguous dim 0, every thread walks its line serially.
// Traffic (dim 0): keys once + values once + n / 4 bytes of head bits + outputs (r02a read the keys twice and compared
// them again in pass 2: 74 instructions per element, 1.13 ms for 2^28 s32 keys + values); the reference's loop re-reads
// keys and values every round.
#pragma once

#include "scan_kernels.cuh"

namespace afb {

constexpr int kRbkThreads = 256;
constexpr int kRbkEPT = 8;                               // consecutive elements per thread
constexpr int kRbkTile = kRbkThreads * kRbkEPT;          // 2048 keys per tile

// ------------------------------------------------------------------------------------------------
// pass 1: runs per tile, exclusive bases, total
// ------------------------------------------------------------------------------------------------
struct RbkCountArgs {
    const uint *keys;
    dim_t n, ntiles;
    uint *tile_base;                 // [ntiles + 1]
    unsigned char *bits;             // [ntiles * kRbkTile / 8 + 1]: bit p = key p starts a run; bit n = end-of-line marker
    uint *ticket;                    // self-resetting arrival counter
    void *host_wire;                 // pinned wire record (common.hpp HostRet)
    uint host_seq;
};

// the head bits live behind the tile bases in the same allocation (16-byte aligned)
__host__ __device__ inline size_t rbk_bits_word_offset(dim_t ntiles) { return ((size_t)ntiles + 1 + 3) & ~(size_t)3; }
inline size_t rbk_scratch_words(dim_t ntiles) { return rbk_bits_word_offset(ntiles) + (size_t)ntiles * (kRbkTile / 32) + 4; }

// heads among the EPT keys starting at p0 (bit k: key p0+k starts a run); keys past n never do
AFB_DI uint rbk_heads(const uint *__restrict__ keys, dim_t n, dim_t p0, uint (&k)[kRbkEPT]) {
    if (p0 >= n) return 0u;
    const int nv = (n - p0) >= kRbkEPT ? kRbkEPT : (int)(n - p0);
    if (nv == kRbkEPT && ((reinterpret_cast<uintptr_t>(keys + p0) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < kRbkEPT / 4; j++) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(keys + p0) + j);
            k[4 * j] = q.x; k[4 * j + 1] = q.y; k[4 * j + 2] = q.z; k[4 * j + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kRbkEPT; j++) k[j] = (j < nv) ? __ldg(keys + p0 + j) : 0u;
    }
    uint hb = 0u;
    const bool first = (p0 == 0) || (__ldg(keys + p0 - 1) != k[0]);
    hb |= first ? 1u : 0u;
#pragma unroll
    for (int j = 1; j < kRbkEPT; j++) hb |= (k[j] != k[j - 1] ? 1u : 0u) << j;
    return nv == kRbkEPT ? hb : (hb & ((1u << nv) - 1u));
}

static __global__ void __launch_bounds__(kRbkThreads) rbk_count_kernel(const RbkCountArgs a) {
    __shared__ uint sm[kRbkThreads / 32];
    __shared__ int sm_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // a WARP owns a tile: its 8 sub-chunks of 256 keys are independent loads (16 x 16 bytes in flight per lane), one warp
    // reduction per tile, no CTA barrier in the streaming loop (first cut: a CTA per tile with two barriers and a serial
    // sum by thread 0 per 8 KB -- 2.5 TB/s).  The key in front of a lane's chunk comes from the neighbouring lane by
    // shuffle (r02a: one more dependent scalar load per chunk).
    const dim_t nwarps = (dim_t)gridDim.x * (kRbkThreads / 32);
    const bool al = (reinterpret_cast<uintptr_t>(a.keys) & 15) == 0;
    for (dim_t t = (dim_t)blockIdx.x * (kRbkThreads / 32) + warp; t < a.ntiles; t += nwarps) {
        uint c = 0;
        const dim_t base = t * kRbkTile;
        unsigned char *bt = a.bits + t * (kRbkTile / 8);
        if (al && base + kRbkTile <= a.n) {
            uint last = t > 0 ? __ldg(a.keys + base - 1) : 0u;     // key in front of the sub-chunk's lane 0
#pragma unroll
            for (int sub = 0; sub < kRbkThreads / 32; sub++) {
                const uint4 *p = reinterpret_cast<const uint4 *>(a.keys + base + (dim_t)(sub * 32 + lane) * kRbkEPT);
                const uint4 q0 = __ldg(p), q1 = __ldg(p + 1);
                const uint up = __shfl_up_sync(0xffffffffu, q1.w, 1);
                const uint prevk = lane ? up : last;
                last = __shfl_sync(0xffffffffu, q1.w, 31);
                uint hb = (q0.x != prevk ? 1u : 0u) | (q0.y != q0.x ? 2u : 0u) | (q0.z != q0.y ? 4u : 0u) | (q0.w != q0.z ? 8u : 0u) |
                          (q1.x != q0.w ? 16u : 0u) | (q1.y != q1.x ? 32u : 0u) | (q1.z != q1.y ? 64u : 0u) | (q1.w != q1.z ? 128u : 0u);
                if (t == 0 && sub == 0 && lane == 0) hb |= 1u;      // position 0 always starts a run
                c += __popc(hb);
                bt[sub * 32 + lane] = (unsigned char)hb;
            }
        } else {
#pragma unroll 1
            for (int sub = 0; sub < kRbkThreads / 32; sub++) {
                uint k[kRbkEPT];
                const dim_t p0 = base + (dim_t)(sub * 32 + lane) * kRbkEPT;
                uint hb = rbk_heads(a.keys, a.n, p0, k);
                c += __popc(hb);
                if (p0 <= a.n && a.n < p0 + kRbkEPT) hb |= 1u << (int)(a.n - p0);   // end-of-line marker (not a run)
                bt[sub * 32 + lane] = (unsigned char)hb;
            }
        }
        if (t == a.ntiles - 1 && lane == 0) a.bits[a.ntiles * (kRbkTile / 8)] = (a.n == a.ntiles * kRbkTile) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) a.tile_base[t] = c;
    }
    // ---- the last CTA to arrive scans the counts in place
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const uint arrived = atomicAdd(a.ticket, 1u);
        sm_last = (arrived == gridDim.x - 1);
    }
    __syncthreads();
    if (!sm_last) return;
    __threadfence();
    // every thread owns a contiguous share, a multiple of 4 entries, read and rewritten as 16-byte vectors with 8 loads in
    // flight (first cut: one dependent 4-byte L2 access per entry and thread -- the whole GPU idled ~150 us behind it)
    dim_t per = (a.ntiles + kRbkThreads - 1) / kRbkThreads;
    per = (per + 3) / 4 * 4;
    const dim_t lo = min((dim_t)tid * per, a.ntiles), hi = min(lo + per, a.ntiles);
    const dim_t nvec = (hi - lo) / 4;
    const uint4 *vb = reinterpret_cast<const uint4 *>(a.tile_base + lo);   // lo is a multiple of 4, the array is 16-byte aligned
    uint sum = 0;
    {
        dim_t v = 0;
        for (; v + 8 <= nvec; v += 8) {
            uint4 q[8];
#pragma unroll
            for (int u = 0; u < 8; u++) q[u] = __ldcg(vb + v + u);
#pragma unroll
            for (int u = 0; u < 8; u++) sum += q[u].x + q[u].y + q[u].z + q[u].w;
        }
        for (; v < nvec; v++) {
            const uint4 q = __ldcg(vb + v);
            sum += q.x + q.y + q.z + q.w;
        }
        for (dim_t i = lo + nvec * 4; i < hi; i++) sum += __ldcg(a.tile_base + i);
    }
    uint inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads();
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    uint before = inc - sum;
#pragma unroll
    for (int w = 0; w < kRbkThreads / 32; w++)
        if (w < warp) before += sm[w];
    uint run = before;
    {
        uint4 *wb = reinterpret_cast<uint4 *>(a.tile_base + lo);
        dim_t v = 0;
        for (; v + 8 <= nvec; v += 8) {
            uint4 q[8];
#pragma unroll
            for (int u = 0; u < 8; u++) q[u] = __ldcg(vb + v + u);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                uint4 o;
                o.x = run; run += q[u].x;
                o.y = run; run += q[u].y;
                o.z = run; run += q[u].z;
                o.w = run; run += q[u].w;
                wb[v + u] = o;
            }
        }
        for (; v < nvec; v++) {
            const uint4 q = __ldcg(vb + v);
            uint4 o;
            o.x = run; run += q.x;
            o.y = run; run += q.y;
            o.z = run; run += q.z;
            o.w = run; run += q.w;
            wb[v] = o;
        }
        for (dim_t i = lo + nvec * 4; i < hi; i++) {
            const uint c = __ldcg(a.tile_base + i);
            a.tile_base[i] = run;
            run += c;
        }
    }
    if (tid == kRbkThreads - 1) {
        a.tile_base[a.ntiles] = run;   // per >= 1 element for the low threads; the last thread's `run` is the total
        *a.ticket = 0u;
        uint w[10];
#pragma unroll
        for (int j = 0; j < 10; j++) w[j] = 0u;
        w[0] = (uint)(unsigned long long)run, w[1] = (uint)((unsigned long long)run >> 32);
        host_ret_send(a.host_wire, a.host_seq, w);
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2 (dim 0): per (tile, column) segmented reduction
// ------------------------------------------------------------------------------------------------
template<typename Ti, typename To, typename Acc>
struct RbkArgs {
    const uint *keys;
    dim_t n, ntiles;
    const Ti *vals;
    ColGeom g;                      // columns of `vals` (g.len == n)
    const uint *tile_base;
    const unsigned char *bits;      // run-head bits written by pass 1
    uint *keys_out;                 // may be null
    To *vals_out;                   // column c starts at c * n_reduced
    dim_t n_reduced;
    Slot<Seg<Acc>> *tile_agg;       // [ncols * ntiles]
    ValParams prm;
};

template<int OP, typename Ti, typename To>
__global__ void __launch_bounds__(kRbkThreads) rbk_tile_kernel(const RbkArgs<Ti, To, typename compute_of<To>::type> a) {
    typedef typename compute_of<To>::type Acc;
    typedef Rule<OP, Ti, Acc> R;
    typedef SOps<R, Acc, true> S;
    typedef typename S::State State;
    constexpr int EPT = kRbkEPT, NW = kRbkThreads / 32;
    __shared__ State sm_w[NW];
    __shared__ uint sm_h[NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const dim_t nwork = a.ntiles * a.g.ncols;
    static_assert(EPT == 8, "one byte of head bits per thread");
    // (tile, column) of the work item, advanced without 64-bit divisions (~100 instructions each, per thread and item).
    // Measured and rejected: fetching the next item's values one iteration ahead -- 58 registers instead of 32 halve the
    // resident CTAs: 0.66 -> 0.84 ms for 2^28 keys + values; capped at 40 registers (6 CTAs per SM) 0.77 ms; CTAs of 128
    // threads (1024-key tiles) 0.78 ms; the fetch moved behind the serial scan (32 registers kept) 0.71 ms.  The kernel
    // wants 2048 resident threads per SM and no extra loads in its dependency chain.
    dim_t t = (dim_t)blockIdx.x % a.ntiles, c = (dim_t)blockIdx.x / a.ntiles;
    for (dim_t w = blockIdx.x; w < nwork; w += gridDim.x) {
        const dim_t p0 = t * kRbkTile + (dim_t)tid * EPT;
        const dim_t bi = t * (kRbkTile / 8) + tid;
        const uint rawb = (uint)__ldg(a.bits + bi) | ((uint)__ldg(a.bits + bi + 1) << 8);
        const int nv = p0 >= a.n ? 0 : ((a.n - p0) >= EPT ? EPT : (int)(a.n - p0));
        const uint vmask = (1u << nv) - 1u;
        const uint hb = rawb & vmask;            // element j starts a run
        const uint tb = (rawb >> 1) & vmask;     // element j ends one: its successor starts a run, or the line ends
        const Ti *col = a.vals + col_offset(a.g, c);
        Acc v[EPT];
        {
            Ti raw[EPT];
            constexpr size_t kAlign = sizeof(Ti) * EPT >= 16 ? 16 : sizeof(Ti) * EPT;
            if (nv == EPT && (reinterpret_cast<uintptr_t>(col + p0) % kAlign) == 0) {
                load_elems<Ti, EPT>(col + p0, raw);
#pragma unroll
                for (int j = 0; j < EPT; j++) v[j] = R::pre(raw[j]);
            } else {
#pragma unroll
                for (int j = 0; j < EPT; j++) v[j] = j < nv ? R::pre(ldg_stream<Ti>(col + p0 + j)) : R::init();
            }
        }
        if (a.prm.change_nan) {
#pragma unroll
            for (int j = 0; j < EPT; j++) v[j] = nan_swap<Acc>(v[j], a.prm.nanval);
        }
        // the first run of a line starts from init() (reduce.hpp:137,146-148: b8 min / max normalise through it); a later
        // run's first element is taken as is (the reference seeds the accumulator with it, reduce.hpp:152-156)
        if (p0 == 0 && nv > 0) v[0] = R::join(R::init(), v[0]);
        Acc x[EPT];
        Acc run = v[0];
        x[0] = run;
#pragma unroll
        for (int j = 1; j < EPT; j++) {
            run = ((hb >> j) & 1u) ? v[j] : R::join(run, v[j]);
            x[j] = run;
        }
        // segmented scan over the threads: a thread with a head cuts the chain, so lane L folds the lanes from the nearest
        // head lane at or below it -- known from ONE ballot -- and only values travel through the shuffles (r02a shuffled
        // (value, flag) pairs and every thread folded the 8 warp aggregates serially: 45 instructions per element)
        const uint hm = __ballot_sync(0xffffffffu, hb != 0u);
        const uint hle = hm & (0xffffffffu >> (31 - lane));
        const int start = hle ? 31 - __clz(hle) : 0;
        Acc incv = run;
        uint hinc = __popc(hb);
        const uint hc = hinc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const Acc o = shfl_up(incv, d);
            const uint ho = __shfl_up_sync(0xffffffffu, hinc, d);
            if (lane - d >= start) incv = R::join(o, incv);
            if (lane >= d) hinc += ho;
        }
        State exc = State{shfl_up(incv, 1), (hm & ((1u << lane) - 1u)) != 0u ? 1 : 0};
        if (lane == 0) exc = S::identity();
        __syncthreads();   // protects sm_* against the previous work item
        if (lane == 31) {
            sm_w[warp] = State{incv, hm != 0u ? 1 : 0};
            sm_h[warp] = hinc;
        }
        __syncthreads();
        // the NW warp aggregates, scanned the same way by every warp (lanes 0 .. NW-1)
        const State wv = lane < NW ? sm_w[lane] : S::identity();
        uint whinc = lane < NW ? sm_h[lane] : 0u;
        const uint wm = __ballot_sync(0xffffffffu, wv.flag != 0);
        const uint wle = wm & (0xffffffffu >> (31 - lane));
        const int wstart = wle ? 31 - __clz(wle) : 0;
        Acc wincv = wv.val;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
            const Acc o = shfl_up(wincv, d);
            const uint ho = __shfl_up_sync(0xffffffffu, whinc, d);
            if (lane - d >= wstart) wincv = R::join(o, wincv);
            if (lane >= d) whinc += ho;
        }
        const int wsrc = warp > 0 ? warp - 1 : 0;
        const Acc bval = shfl_idx(wincv, wsrc);
        const uint bh = __shfl_sync(0xffffffffu, whinc, wsrc);
        State before = S::identity();
        uint hbefore = hinc - hc;
        if (warp > 0) {
            before = State{bval, (wm & ((1u << warp) - 1u)) != 0u ? 1 : 0};
            hbefore += bh;
        }
        const State prefix = S::join(before, exc);   // the tile's elements in front of this thread
        if (hb | tb) {                               // most threads of long runs hold neither a head nor a tail
            const uint base = __ldg(a.tile_base + t) + hbefore - 1u;   // (- 1: the run that entered the tile from the left)
            To *out = a.vals_out + c * a.n_reduced;
#pragma unroll
            for (int j = 0; j < EPT; j++) {
                if (((hb | tb) >> j) & 1u) {
                    const uint upto = hb & ((2u << j) - 1u);           // heads among this thread's elements 0..j
                    const uint seg = base + __popc(upto);
                    if (((hb >> j) & 1u) && c == 0 && a.keys_out) a.keys_out[seg] = __ldg(a.keys + p0 + j);
                    if ((tb >> j) & 1u) out[seg] = conv<To, Acc>(upto == 0u ? R::join(prefix.val, x[j]) : x[j]);
                }
            }
        }
        if (tid == kRbkThreads - 1) a.tile_agg[c * a.ntiles + t].v = S::join(prefix, State{run, hb != 0u ? 1 : 0});
        t += gridDim.x;
        while (t >= a.ntiles) {
            t -= a.ntiles;
            c++;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pass 3 (dim 0): add what a run accumulated in earlier tiles to the tile that holds its tail.
// A segmented scan over the per-tile aggregates, ONE THREAD PER TILE (coalesced), in two small launches:
//   rbk_carry_local_kernel   every CTA scans its 256 aggregates (warp shuffles + shared memory), leaves each tile's
//                            CTA-local exclusive state and the CTA's own aggregate;
//   rbk_carry_apply_kernel   every CTA folds the aggregates of the CTAs before it (one warp, in order), completes its
//                            tiles' carries and applies them.
// (First cut: one CTA walked all tiles with a dependent-load chain per tile -- 1.03 ms for the 131072 tiles of 2^28 keys,
// more than the two streaming passes together.)
// ------------------------------------------------------------------------------------------------
template<int OP, typename To>
__global__ void __launch_bounds__(kRbkThreads) rbk_carry_local_kernel(dim_t ntiles,
                                                                       const Slot<Seg<typename compute_of<To>::type>> *__restrict__ tile_agg,
                                                                       Slot<Seg<typename compute_of<To>::type>> *__restrict__ loc_exc,
                                                                       Slot<Seg<typename compute_of<To>::type>> *__restrict__ blk_agg) {
    typedef typename compute_of<To>::type Acc;
    typedef Rule<OP, To, Acc> R;
    typedef SOps<R, Acc, true> S;
    typedef typename S::State State;
    constexpr int NW = kRbkThreads / 32;
    __shared__ State sm_w[NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const dim_t c = blockIdx.y, nblk = gridDim.x;
    const dim_t t = (dim_t)blockIdx.x * kRbkThreads + tid;
    const State mine = t < ntiles ? tile_agg[c * ntiles + t].v : S::identity();
    State inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const State o = shfl_up(inc, d);
        if (lane >= d) inc = S::join(o, inc);
    }
    State exc = shfl_up(inc, 1);
    if (lane == 0) exc = S::identity();
    if (lane == 31) sm_w[warp] = inc;
    __syncthreads();
    State carry = S::identity();
#pragma unroll
    for (int q = 0; q < NW; q++)
        if (q < warp) carry = S::join(carry, sm_w[q]);
    if (t < ntiles) loc_exc[c * ntiles + t].v = S::join(carry, exc);
    if (tid == kRbkThreads - 1) blk_agg[c * nblk + blockIdx.x].v = S::join(carry, inc);
}

template<int OP, typename To>
__global__ void __launch_bounds__(kRbkThreads) rbk_carry_apply_kernel(const uint *__restrict__ keys, dim_t ntiles,
                                                                       const uint *__restrict__ tile_base,
                                                                       const Slot<Seg<typename compute_of<To>::type>> *__restrict__ loc_exc,
                                                                       const Slot<Seg<typename compute_of<To>::type>> *__restrict__ blk_agg,
                                                                       To *vals_out, dim_t n_reduced) {
    typedef typename compute_of<To>::type Acc;
    typedef Rule<OP, To, Acc> R;
    typedef SOps<R, Acc, true> S;
    typedef typename S::State State;
    __shared__ State sm_before;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const dim_t c = blockIdx.y, nblk = gridDim.x;
    if (warp == 0) {
        // state of everything before this CTA's first tile: the aggregates of the CTAs before it, folded in order
        // (each lane takes a contiguous share, then an ordered shuffle scan)
        const dim_t nb = blockIdx.x;
        const dim_t per = (nb + 31) / 32;
        const dim_t lo = min((dim_t)lane * per, nb), hi = min(lo + per, nb);
        State tot = S::identity();
        for (dim_t q = lo; q < hi; q++) tot = S::join(tot, blk_agg[c * nblk + q].v);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const State o = shfl_up(tot, d);
            if (lane >= d) tot = S::join(o, tot);
        }
        if (lane == 31) sm_before = tot;
    }
    __syncthreads();
    const dim_t t = (dim_t)blockIdx.x * kRbkThreads + tid;
    if (t == 0 || t >= ntiles) return;
    const bool starts_run = (__ldg(reinterpret_cast<const unsigned char *>(tile_base + rbk_bits_word_offset(ntiles)) + t * (kRbkTile / 8)) & 1u) != 0u;
    const uint b = __ldg(tile_base + t), heads = __ldg(tile_base + t + 1) - b;
    // the run entering from the left ends in this tile iff the tile has a head or is the last one
    if (!starts_run && (heads > 0u || t == ntiles - 1)) {
        const State carry = S::join(sm_before, loc_exc[c * ntiles + t].v);
        To *o = vals_out + c * n_reduced + (b - 1u);
        *o = conv<To, Acc>(R::join(carry.val, conv<Acc, To>(*o)));
    }
}

// keys_out for the dim > 0 path: one key per run, positions from the tile bases
static __global__ void __launch_bounds__(kRbkThreads) rbk_keys_out_kernel(const uint *__restrict__ keys, dim_t n, dim_t ntiles,
                                                                    const uint *__restrict__ tile_base, uint *keys_out) {
    __shared__ uint sm_h[kRbkThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (dim_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        uint k[kRbkEPT];
        const uint hb = rbk_heads(keys, n, t * kRbkTile + (dim_t)tid * kRbkEPT, k);
        const uint hc = __popc(hb);
        uint hinc = hc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint o = __shfl_up_sync(0xffffffffu, hinc, d);
            if (lane >= d) hinc += o;
        }
        __syncthreads();
        if (lane == 31) sm_h[warp] = hinc;
        __syncthreads();
        uint pos = __ldg(tile_base + t) + hinc - hc;
#pragma unroll
        for (int q = 0; q < kRbkThreads / 32; q++)
            if (q < warp) pos += sm_h[q];
#pragma unroll
        for (int j = 0; j < kRbkEPT; j++)
            if ((hb >> j) & 1u) keys_out[pos++] = k[j];
    }
}

// ------------------------------------------------------------------------------------------------
// dim > 0: one thread per line, lanes along the contiguous dim 0
// ------------------------------------------------------------------------------------------------
template<typename Ti, typename To>
struct RbkDimArgs {
    const uint *keys;
    const Ti *vals;
    To *vals_out;
    DimGeom g;          // len = n, sd = input stride of the reduced dim; oa / ob = output strides of the batch dims
    dim_t od;           // output stride of the reduced dim
    ValParams prm;
};

template<int OP, typename Ti, typename To>
__global__ void __launch_bounds__(256) rbk_dim_kernel(const RbkDimArgs<Ti, To> a) {
    typedef typename compute_of<To>::type Acc;
    typedef Rule<OP, Ti, Acc> R;
    const dim_t x0 = (dim_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x0 >= a.g.d0) return;
    for (dim_t batch = blockIdx.y; batch < a.g.na * a.g.nb; batch += gridDim.y) {
        const dim_t ia = batch % a.g.na, ib = batch / a.g.na;
        const Ti *in = a.vals + ia * a.g.sa + ib * a.g.sb + x0;
        To *out = a.vals_out + ia * a.g.oa + ib * a.g.ob + x0;
        Acc acc = R::init();
        uint prev = 0;
        dim_t seg = 0;
        for (dim_t i = 0; i < a.g.len; i++) {
            const uint key = __ldg(a.keys + i);
            Acc v = R::pre(ldg_stream<Ti>(in + i * a.g.sd));
            if (a.prm.change_nan) v = nan_swap<Acc>(v, a.prm.nanval);
            if (i == 0) {
                acc = R::join(R::init(), v);   // the first run starts from init() (reduce.hpp:137,146-148)
            } else if (key != prev) {
                out[seg * a.od] = conv<To, Acc>(acc);
                seg++;
                acc = v;
            } else {
                acc = R::join(acc, v);
            }
            prev = key;
        }
        if (a.g.len > 0) out[seg * a.od] = conv<To, Acc>(acc);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template<int OP, typename Ti, typename To>
void run_rbk(const LaunchCtx &c, const void *keys, dim_t n, const View &vals, int dim, bool change_nan, double nanval,
             const unsigned *tile_base, void *keys_out, void *vals_out, dim_t n_reduced) {
    typedef typename compute_of<To>::type Acc;
    const dim_t ntiles = (n + kRbkTile - 1) / kRbkTile;
    ValParams prm;
    prm.change_nan = change_nan ? 1 : 0;
    prm.nanval = nanval;
    if (dim == 0) {
        RbkArgs<Ti, To, Acc> a;
        a.keys = (const uint *)keys;
        a.n = n;
        a.ntiles = ntiles;
        a.vals = (const Ti *)vals.ptr;
        a.g = col_geom(vals);
        a.tile_base = tile_base;
        a.bits = reinterpret_cast<const unsigned char *>(tile_base + rbk_bits_word_offset(ntiles));
        a.keys_out = (uint *)keys_out;
        a.vals_out = (To *)vals_out;
        a.n_reduced = n_reduced;
        a.prm = prm;
        if (a.g.ncols == 0) return;
        // one partials block: [tile aggregates | CTA-local exclusive states | per-CTA aggregates of the carry scan]
        const dim_t nblk = (ntiles + kRbkThreads - 1) / kRbkThreads;
        const size_t per = (size_t)ntiles * a.g.ncols;
        Slot<Seg<Acc>> *scr = (Slot<Seg<Acc>> *)c.scratch->need_partials((2 * per + (size_t)nblk * a.g.ncols) * sizeof(Slot<Seg<Acc>>));
        a.tile_agg = scr;
        Slot<Seg<Acc>> *loc_exc = scr + per, *blk_agg = scr + 2 * per;
        dim_t grid = ntiles * a.g.ncols;
        const dim_t cap = (dim_t)c.sms * 32;
        if (grid > cap) grid = cap;
        rbk_tile_kernel<OP, Ti, To><<<(unsigned)grid, kRbkThreads, 0, c.stream>>>(a);
        post_launch_check("rbk_tile_kernel");
        if (a.g.ncols > 65535) fail(AF_ERR_SIZE, "reduce by key: too many columns (%lld)", (long long)a.g.ncols);
        if (ntiles > 1) {
            const dim3 cgrid((unsigned)nblk, (unsigned)a.g.ncols, 1);
            rbk_carry_local_kernel<OP, To><<<cgrid, kRbkThreads, 0, c.stream>>>(ntiles, a.tile_agg, loc_exc, blk_agg);
            post_launch_check("rbk_carry_local_kernel");
            rbk_carry_apply_kernel<OP, To><<<cgrid, kRbkThreads, 0, c.stream>>>((const uint *)keys, ntiles, tile_base, loc_exc,
                                                                                 blk_agg, (To *)vals_out, n_reduced);
            post_launch_check("rbk_carry_apply_kernel");
            note_launches(3);
        } else {
            note_launches(1);
        }
        return;
    }
    if (keys_out) {
        dim_t grid = ntiles;
        const dim_t cap = (dim_t)c.sms * 16;
        if (grid > cap) grid = cap;
        rbk_keys_out_kernel<<<(unsigned)grid, kRbkThreads, 0, c.stream>>>((const uint *)keys, n, ntiles, tile_base, (uint *)keys_out);
        post_launch_check("rbk_keys_out_kernel");
        note_launches(1);
    }
    RbkDimArgs<Ti, To> a;
    a.keys = (const uint *)keys;
    a.vals = (const Ti *)vals.ptr;
    a.vals_out = (To *)vals_out;
    DimGeom &g = a.g;
    g.d0 = vals.dims[0];
    g.len = n;
    g.sd = vals.strides[dim];
    int others[2], m = 0;
    for (int i = 1; i < 4; i++)
        if (i != dim) others[m++] = i;
    g.na = vals.dims[others[0]];
    g.sa = vals.strides[others[0]];
    g.nb = vals.dims[others[1]];
    g.sb = vals.strides[others[1]];
    dim_t od[4] = {vals.dims[0], vals.dims[1], vals.dims[2], vals.dims[3]}, os[4];
    od[dim] = n_reduced;
    dense_strides(od, os);
    g.oa = os[others[0]];
    g.ob = os[others[1]];
    a.od = os[dim];
    a.prm = prm;
    if (g.d0 == 0 || g.na * g.nb == 0) return;
    dim_t gy = g.na * g.nb;
    if (gy > 65535) gy = 65535;
    const int threads = g.d0 >= 256 ? 256 : (g.d0 >= 128 ? 128 : (g.d0 >= 64 ? 64 : 32));
    const dim_t gx = (g.d0 + threads - 1) / threads;
    rbk_dim_kernel<OP, Ti, To><<<dim3((unsigned)gx, (unsigned)gy), threads, 0, c.stream>>>(a);
    post_launch_check("rbk_dim_kernel");
    note_launches(1);
}

// families, one translation unit each (rbk_sum.cu, rbk_minmax.cu, rbk_logic.cu)
bool rbk_family_sum(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const void *keys, dim_t n, const View &vals, int dim,
                    bool cn, double nv, const unsigned *tile_base, void *keys_out, void *vals_out, dim_t n_reduced);
bool rbk_family_minmax(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const void *keys, dim_t n, const View &vals, int dim,
                       bool cn, double nv, const unsigned *tile_base, void *keys_out, void *vals_out, dim_t n_reduced);
bool rbk_family_logic(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const void *keys, dim_t n, const View &vals, int dim,
                      bool cn, double nv, const unsigned *tile_base, void *keys_out, void *vals_out, dim_t n_reduced);

}  // namespace afb
