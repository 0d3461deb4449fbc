// scan_kernels.cuh -- inclusive / exclusive scans and segmented (by-key) scans.
//
// What they replace in the reference (behaviour studied, design is new):
//   scan_first + scan_first_bcast, scan_dim + scan_dim_bcast            (3 launches, ~4N words of traffic)
//       /root/reference/src/backend/cuda/kernel/scan_first.cuh:19-135, scan_dim.cuh:19-169
//   scanbykey_first_{nonfinal,final,bcast}, scanbykey_dim_*              (3 launches + 3 temporaries)
//       /root/reference/src/backend/cuda/kernel/scan_first_by_key.cuh:24-314, scan_dim_by_key.cuh:25-400
// Semantics (what every element must equal) follow the CPU backend:
//       /root/reference/src/backend/cpu/kernel/scan.hpp:37-72, scan_by_key.hpp:43-86
//
// B200 design:
//   * lookback_kernel -- dim-0 scans of long columns in ONE pass (2N words of traffic): every CTA takes a
//     tile by ticket, scans it in registers, publishes its aggregate and resolves its prefix with a
//     decoupled look-back over the predecessors' descriptors (acquire/release flags, generation-tagged so
//     the descriptor array is never cleared between launches).  Data never goes through shared memory:
//     each lane keeps the 16-byte vectors it loaded (warp-striped, perfectly coalesced) and the scan is
//     composed of in-vector scans + one warp shuffle-scan per vector slot.  A segmented scan carries a
//     (value, head-seen) pair through the same machinery.
//   * warp_scan_cols_kernel -- many short columns: one warp per column.
//   * strip_scan_kernel -- scans along dim>0: lanes run along the contiguous dim 0, a CTA walks down the
//     scanned axis in strips, carrying the running value.
#pragma once

#include "reduce_kernels.cuh"

namespace afb {

// ------------------------------------------------------------------------------------------------
// scan state: plain value, or (value, "a segment head has been seen") for by-key scans
// ------------------------------------------------------------------------------------------------
template<typename T> struct Seg {
    T val;
    int flag;
};

template<typename R, typename T, bool KEYED> struct SOps;
template<typename R, typename T> struct SOps<R, T, false> {
    typedef T State;
    AFB_DI static State identity() { return R::init(); }
    AFB_DI static State lift(T v, bool) { return v; }
    AFB_DI static State join(State a, State b) { return R::join(a, b); }  // a: earlier positions
    AFB_DI static T value(State s) { return s; }
};
template<typename R, typename T> struct SOps<R, T, true> {
    typedef Seg<T> State;
    AFB_DI static State identity() { return State{R::init(), 0}; }
    AFB_DI static State lift(T v, bool head) { return State{v, head ? 1 : 0}; }
    AFB_DI static State join(State a, State b) { return State{b.flag ? b.val : R::join(a.val, b.val), a.flag | b.flag}; }
    AFB_DI static T value(State s) { return s.val; }
};

// elements per thread-vector: the wider of input / output uses 16-byte accesses
template<typename Ti, typename To> struct ScanVec {
    static constexpr int wide = sizeof(Ti) > sizeof(To) ? (int)sizeof(Ti) : (int)sizeof(To);
    static constexpr int value = 16 / wide > 0 ? 16 / wide : 1;
};

template<typename T, int E> AFB_DI void store_pack(T *p, const Pack<T, E> &v) {
    constexpr int B = (int)sizeof(T) * E;
    if constexpr (E == 1) {
        *p = v.e[0];
    } else if constexpr (B == 16) {
        stg_stream16(p, *reinterpret_cast<const Vec16 *>(&v));
    } else if constexpr (B == 8) {
        *reinterpret_cast<uint2 *>(p) = *reinterpret_cast<const uint2 *>(&v);
    } else {
        *reinterpret_cast<uint32_t *>(p) = *reinterpret_cast<const uint32_t *>(&v);
    }
}

// E consecutive elements with the widest legal accesses (<= 16 bytes each)
template<typename T, int E> AFB_DI void load_elems(const T *p, T (&out)[E]) {
    constexpr int B = (int)sizeof(T) * E;
    if constexpr (B > 16) {
        constexpr int PER = 16 / (int)sizeof(T);
#pragma unroll
        for (int i = 0; i < E / PER; i++) {
            const Pack<T, PER> t = load_pack<T, PER>(p + i * PER);
#pragma unroll
            for (int k = 0; k < PER; k++) out[i * PER + k] = t.e[k];
        }
    } else {
        const Pack<T, E> t = load_pack<T, E>(p);
#pragma unroll
        for (int k = 0; k < E; k++) out[k] = t.e[k];
    }
}

constexpr int kScanWarps = 8;
constexpr int kScanThreads = 32 * kScanWarps;
constexpr int kScanSlots = 4;  // vectors per thread

enum { TILE_INVALID = 0, TILE_AGGREGATE = 1, TILE_PREFIX = 2 };

// One tile descriptor: a generation-tagged status word + the two published values.
template<typename State> struct alignas(16) TileDesc {
    uint status;  // (epoch << 2) | state
    uint pad[3];
    Slot<State> aggregate;
    Slot<State> inclusive;
};

template<typename Ti, typename To, typename Tk>
struct LookbackArgs {
    const Ti *in;
    const Tk *key;      // null when not keyed
    To *out;
    ColGeom g;          // input columns
    dim_t ks1, ks2, ks3;  // key strides of dims 1..3
    dim_t tiles_per_col;
    dim_t ntiles;
    int inclusive;
    int vec_ok;         // 16-byte accesses are legal for in / key / out
    uint epoch;
    void *tiles;        // TileDesc<State>[ntiles]
    uint *tile_counter;
    const To *carry;    // optional running value before position 0 (sharded scans; single column only)
    const Tk *carry_key;  // keyed: the carry applies while key == *carry_key
    // optional shard summary (single column only), device memory, 32 bytes:
    //   [0,16) reduction of the last open segment (whole shard for a plain scan), carry excluded
    //   [16,20) u32: a segment head occurs in the shard   [24,32) u64: position of the first head (~0 if none)
    void *summary;
};

template<int OP, typename Ti, typename To, bool KEYED, typename Tk>
__global__ void __launch_bounds__(kScanThreads) lookback_kernel(const LookbackArgs<Ti, To, Tk> a) {
    typedef Rule<OP, Ti, To> R;
    typedef SOps<R, To, KEYED> S;
    typedef typename S::State State;
    typedef TileDesc<State> Desc;
    constexpr int E = ScanVec<Ti, To>::value;
    constexpr int U = kScanSlots;
    constexpr int TILE = kScanThreads * U * E;

    __shared__ State sm_warp[kScanWarps];
    __shared__ uint sm_head[kScanWarps];
    __shared__ State sm_tile_prefix;
    __shared__ uint sm_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- take a tile by ticket: a tile only ever waits for tiles with smaller tickets, which are
    //      already resident or finished -> forward progress without relying on block scheduling order
    if (tid == 0) {
        const uint t = atomicAdd(a.tile_counter, 1u);
        sm_tile = t;
        if (t == (uint)(a.ntiles - 1)) *a.tile_counter = 0u;  // everyone has drawn: reset for the next launch
    }
    __syncthreads();
    const dim_t tile = sm_tile;
    const dim_t col = tile / a.tiles_per_col, tcol = tile % a.tiles_per_col;
    const dim_t tile_lo = tcol * TILE;
    const Ti *in = a.in + col_offset(a.g, col);
    To *out = a.out + col * a.g.len;  // output is dense
    const Tk *key = nullptr;
    if (KEYED) {
        ColGeom kg = a.g;
        kg.s1 = a.ks1; kg.s2 = a.ks2; kg.s3 = a.ks3;
        key = a.key + col_offset(kg, col);
    }

    // ---- load (warp-striped vectors) and scan inside each vector
    State x[U][E];       // inclusive scan inside the vector
    State vtot[U];
    uint heads[U];       // bit k: element k of the vector starts a segment (by-key scans)
    const dim_t warp_lo = tile_lo + (dim_t)warp * (U * 32 * E);
#pragma unroll
    for (int u = 0; u < U; u++) {
        const dim_t p0 = warp_lo + ((dim_t)u * 32 + lane) * E;
        struct { Ti e[E]; } v;
        struct { Tk e[E]; } kv;
        Tk kprev = Tk();
        const bool full = a.vec_ok && (p0 + E <= a.g.len);
        if (full) {
            load_elems<Ti, E>(in + p0, v.e);
            if (KEYED) load_elems<Tk, E>(key + p0, kv.e);
        } else {
#pragma unroll
            for (int k = 0; k < E; k++) {
                if (p0 + k < a.g.len) {
                    v.e[k] = ldg_stream<Ti>(in + p0 + k);
                    if (KEYED) kv.e[k] = __ldg(key + p0 + k);
                }
            }
        }
        if (KEYED) {
            // key just before this vector: the neighbour lane's last key, or memory at the seams
            Tk last = kv.e[E - 1];
            Tk from_lane = shfl_up(last, 1);
            if (lane > 0) kprev = from_lane;
            else if (p0 > 0 && p0 < a.g.len) kprev = __ldg(key + p0 - 1);
        }
        State run = S::identity();
        heads[u] = 0u;
#pragma unroll
        for (int k = 0; k < E; k++) {
            State s = S::identity();
            if (p0 + k < a.g.len) {
                bool head = false;
                if (KEYED) {
                    // position 0 of a line has nothing before it, so it needs no head mark (and, like the
                    // reference, its value IS joined with the identity) -- unless a carry from a previous
                    // shard must be cut off because the key changed across the shard boundary
                    if (p0 + k == 0) head = a.carry_key ? (kv.e[0] != *a.carry_key) : false;
                    else head = kv.e[k] != ((k == 0) ? kprev : kv.e[k - 1]);
                    if (head) heads[u] |= 1u << k;
                }
                s = S::lift(R::pre(v.e[k]), head);
            }
            run = (k == 0) ? s : S::join(run, s);
            x[u][k] = run;
        }
        vtot[u] = run;
    }

    // ---- warp level: one shuffle-scan per vector slot, slots chained in order
    State slot_excl[U];  // everything in this warp's region before vector (u, lane)
    State warp_total = S::identity();
#pragma unroll
    for (int u = 0; u < U; u++) {
        State inc = vtot[u];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            State o = shfl_up(inc, d);
            if (lane >= d) inc = S::join(o, inc);
        }
        State exc = shfl_up(inc, 1);
        if (lane == 0) exc = S::identity();
        slot_excl[u] = S::join(warp_total, exc);
        warp_total = S::join(warp_total, shfl_idx(inc, 31));
    }

    // ---- CTA level
    if (lane == 0) sm_warp[warp] = warp_total;
    if (KEYED && a.summary) {  // first segment head of the tile, as an offset from tile_lo
        uint first = 0xffffffffu;
#pragma unroll
        for (int u = U - 1; u >= 0; u--)
            if (heads[u]) first = (uint)(warp * (U * 32 * E) + (u * 32 + lane) * E + (__ffs(heads[u]) - 1));
        first = __reduce_min_sync(0xffffffffu, first);
        if (lane == 0) sm_head[warp] = first;
    }
    __syncthreads();
    State warp_excl = S::identity(), tile_total = S::identity();
#pragma unroll
    for (int w = 0; w < kScanWarps; w++) {
        const State t = sm_warp[w];
        if (w < warp) warp_excl = S::join(warp_excl, t);
        tile_total = S::join(tile_total, t);
    }

    // ---- decoupled look-back (warp 0)
    Desc *descs = reinterpret_cast<Desc *>(a.tiles);
    const uint tag = a.epoch << 2;
    if (warp == 0) {
        State excl = S::identity();
        if (tcol == 0) {
            if (a.carry) excl = S::lift(*a.carry, false);
            if (lane == 0) {
                descs[tile].inclusive.v = S::join(excl, tile_total);
                __threadfence();
                st_release(&descs[tile].status, tag | TILE_PREFIX);
            }
        } else {
            if (lane == 0) {
                descs[tile].aggregate.v = tile_total;
                __threadfence();
                st_release(&descs[tile].status, tag | TILE_AGGREGATE);
            }
            dim_t look = tile - 1;          // nearest predecessor handled by lane 0
            const dim_t first = tile - tcol;  // first tile of this column
            bool done = false;
            while (!done) {
                const dim_t mine = look - lane;
                const bool valid = mine >= first;
                uint st = 0;
                if (valid) {
                    { uint spins = 0; do { st = ld_acquire(&descs[mine].status); if (++spins > kSpinLimit) __trap(); } while ((st & ~3u) != tag || (st & 3u) == TILE_INVALID); }
                }
                const bool is_prefix = valid && (st & 3u) == TILE_PREFIX;
                const uint pmask = __ballot_sync(0xffffffffu, is_prefix);
                const uint vmask = __ballot_sync(0xffffffffu, valid);
                // lanes up to (and including) the nearest PREFIX take part; the rest contribute the identity
                int stop = pmask ? (__ffs(pmask) - 1) : 31;
                State v = S::identity();
                if (valid && lane <= stop) v = load_cg(is_prefix ? &descs[mine].inclusive : &descs[mine].aggregate);
                // ordered fold: lane+d holds an EARLIER tile than lane
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    State o = shfl_down(v, d);
                    if (lane + d < 32) v = S::join(o, v);
                }
                v = shfl_idx(v, 0);
                excl = S::join(v, excl);
                if (pmask) done = true;
                else if (vmask != 0xffffffffu) {
                    // ran past the first tile of the column without meeting a PREFIX: cannot happen (the first
                    // tile always publishes PREFIX), kept as a guard against spinning forever
                    done = true;
                } else look -= 32;
            }
            if (lane == 0) {
                descs[tile].inclusive.v = S::join(excl, tile_total);
                __threadfence();
                st_release(&descs[tile].status, tag | TILE_PREFIX);
            }
        }
        if (lane == 0) sm_tile_prefix = excl;
        if (a.summary && lane == 0) {
            if (KEYED) {
                uint first = 0xffffffffu;
#pragma unroll
                for (int w = 0; w < kScanWarps; w++) first = min(first, sm_head[w]);
                if (first != 0xffffffffu)
                    atomicMin(reinterpret_cast<unsigned long long *>((char *)a.summary + 24),
                              (unsigned long long)(tile_lo + first));
            }
            if (tcol == a.tiles_per_col - 1) {
                // the state at the end of the shard (summary mode is launched without a carry-in)
                const State fin = S::join(excl, tile_total);
                *reinterpret_cast<Slot<To> *>(a.summary) = Slot<To>{S::value(fin)};
                uint flag = 0;
                if constexpr (KEYED) flag = fin.flag ? 1u : 0u;
                *reinterpret_cast<uint *>((char *)a.summary + 16) = flag;
            }
        }
    }
    __syncthreads();
    const State before_warp = S::join(sm_tile_prefix, warp_excl);

    // ---- write
#pragma unroll
    for (int u = 0; u < U; u++) {
        const dim_t p0 = warp_lo + ((dim_t)u * 32 + lane) * E;
        if (p0 >= a.g.len) continue;
        const State before_vec = S::join(before_warp, slot_excl[u]);
        Pack<To, E> o;
#pragma unroll
        for (int k = 0; k < E; k++) {
            State r;
            if (a.inclusive) r = S::join(before_vec, x[u][k]);
            else r = (k == 0) ? before_vec : S::join(before_vec, x[u][k - 1]);
            To val = S::value(r);
            // exclusive segmented scan: a segment head restarts from the identity
            if (KEYED && !a.inclusive && ((heads[u] >> k) & 1u)) val = R::init();
            o.e[k] = val;
        }
        if (a.vec_ok && p0 + E <= a.g.len) store_pack<To, E>(out + p0, o);
        else {
#pragma unroll
            for (int k = 0; k < E; k++)
                if (p0 + k < a.g.len) out[p0 + k] = o.e[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// warp_scan_cols_kernel: many short columns along dim 0, one warp per column
// ------------------------------------------------------------------------------------------------
template<typename Ti, typename To, typename Tk>
struct WarpScanArgs {
    const Ti *in;
    const Tk *key;
    To *out;
    ColGeom g;
    dim_t ks1, ks2, ks3;
    int inclusive;
};

template<int OP, typename Ti, typename To, bool KEYED, typename Tk>
__global__ void __launch_bounds__(kScanThreads) warp_scan_cols_kernel(const WarpScanArgs<Ti, To, Tk> a) {
    typedef Rule<OP, Ti, To> R;
    typedef SOps<R, To, KEYED> S;
    typedef typename S::State State;
    const int lane = threadIdx.x & 31;
    const dim_t warps_total = (dim_t)gridDim.x * kScanWarps;
    ColGeom kg = a.g;
    kg.s1 = a.ks1; kg.s2 = a.ks2; kg.s3 = a.ks3;
    for (dim_t col = (dim_t)blockIdx.x * kScanWarps + (threadIdx.x >> 5); col < a.g.ncols; col += warps_total) {
        const Ti *in = a.in + col_offset(a.g, col);
        const Tk *key = KEYED ? a.key + col_offset(kg, col) : nullptr;
        To *out = a.out + col * a.g.len;
        State carry = S::identity();
        for (dim_t base = 0; base < a.g.len; base += 32) {
            const dim_t i = base + lane;
            const bool valid = i < a.g.len;
            State s = S::identity();
            bool head = false;
            if (valid) {
                if (KEYED) head = (i > 0) && (__ldg(key + i) != __ldg(key + i - 1));
                s = S::lift(R::pre(ldg_stream<Ti>(in + i)), head);
            }
            State inc = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                State o = shfl_up(inc, d);
                if (lane >= d) inc = S::join(o, inc);
            }
            State exc = shfl_up(inc, 1);
            if (lane == 0) exc = S::identity();
            if (valid) {
                To val;
                if (a.inclusive) val = S::value(S::join(carry, inc));
                else val = (KEYED && head) ? R::init() : S::value(S::join(carry, exc));
                out[i] = val;
            }
            carry = S::join(carry, shfl_idx(inc, 31));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// strip_scan_kernel: scan along dim > 0
// ------------------------------------------------------------------------------------------------
template<typename Ti, typename To, typename Tk>
struct StripScanArgs {
    const Ti *in;
    const Tk *key;
    To *out;
    dim_t d0;
    dim_t len, sd, ksd, osd;          // scanned extent; its input / key / output strides
    dim_t na, sa, ksa, osa;           // batch dim a
    dim_t nb, sb, ksb, osb;           // batch dim b
    int inclusive;
};

constexpr int kStripRows = 8;   // rows per thread per strip
constexpr int kStripY = 8;      // threads along the scanned axis

template<int OP, typename Ti, typename To, bool KEYED, typename Tk>
__global__ void __launch_bounds__(32 * kStripY) strip_scan_kernel(const StripScanArgs<Ti, To, Tk> a) {
    typedef Rule<OP, Ti, To> R;
    typedef SOps<R, To, KEYED> S;
    typedef typename S::State State;
    __shared__ State sm[kStripY][33];
    const int lane = threadIdx.x, wy = threadIdx.y;
    const dim_t x0 = (dim_t)blockIdx.x * 32 + lane;
    const bool live = x0 < a.d0;
    for (dim_t batch = blockIdx.y; batch < a.na * a.nb; batch += gridDim.y) {
        const dim_t ia = batch % a.na, ib = batch / a.na;
        const Ti *in = a.in + ia * a.sa + ib * a.sb + x0;
        const Tk *key = KEYED ? a.key + ia * a.ksa + ib * a.ksb + x0 : nullptr;
        To *out = a.out + ia * a.osa + ib * a.osb + x0;
        State carry = S::identity();
        for (dim_t r0 = 0; r0 < a.len; r0 += kStripY * kStripRows) {
            const dim_t rs = r0 + (dim_t)wy * kStripRows;
            Ti v[kStripRows];
            Tk kv[kStripRows + 1];
            if (live) {
#pragma unroll
                for (int j = 0; j < kStripRows; j++)
                    if (rs + j < a.len) v[j] = ldg_stream<Ti>(in + (rs + j) * a.sd);
                if (KEYED) {
                    if (rs > 0 && rs < a.len) kv[0] = __ldg(key + (rs - 1) * a.ksd);
#pragma unroll
                    for (int j = 0; j < kStripRows; j++)
                        if (rs + j < a.len) kv[j + 1] = __ldg(key + (rs + j) * a.ksd);
                }
            }
            State x[kStripRows];
            uint heads = 0u;
            State run = S::identity();
#pragma unroll
            for (int j = 0; j < kStripRows; j++) {
                State s = S::identity();
                if (live && rs + j < a.len) {
                    bool head = false;
                    if (KEYED) head = (rs + j > 0) && (kv[j + 1] != kv[j]);
                    if (head) heads |= 1u << j;
                    s = S::lift(R::pre(v[j]), head);
                }
                run = (j == 0) ? s : S::join(run, s);
                x[j] = run;
            }
            __syncthreads();
            sm[wy][lane] = run;
            __syncthreads();
            State before = carry, total = carry;
#pragma unroll
            for (int y = 0; y < kStripY; y++) {
                const State t = sm[y][lane];
                if (y < wy) before = S::join(before, t);
                total = S::join(total, t);
            }
            carry = total;
            if (live) {
#pragma unroll
                for (int j = 0; j < kStripRows; j++) {
                    if (rs + j >= a.len) break;
                    To val;
                    if (a.inclusive) val = S::value(S::join(before, x[j]));
                    else if (KEYED && ((heads >> j) & 1u)) val = R::init();
                    else val = S::value(j == 0 ? before : S::join(before, x[j - 1]));
                    out[(rs + j) * a.osd] = val;
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// dim_scan_kernel: scan along dim > 0 with the scanned axis SPLIT into tiles (4-byte outputs, not keyed)
// ------------------------------------------------------------------------------------------------
// strip_scan_kernel above gives one CTA a 32-lane slice of dim 0 and lets it walk the whole scanned axis: an 8192 x 8192
// matrix keeps only 256 CTAs of scalar loads busy, and a 32 x 2^24 one a single CTA.  Here a tile is kDsRows rows of the
// scanned axis by 32 x VEC elements of dim 0 (VEC = 4: every access is a 16-byte vector, 32 KB per tile); tiles are dealt
// by ticket in split-major order, so every predecessor of a tile (same dim-0 slice, earlier rows) is already running or
// done, and the carry travels down the scanned axis by decoupled look-back: ONE pass, 2N words of traffic for any shape.
// A descriptor is one 64-bit word per dim-0 element {value : 32 | generation : 29 | state : 3}, written and read with
// single 8-byte accesses -- value and flag can never be seen apart, so no fences.  (The reference: scan_dim + scan_dim_bcast,
// 3 launches, /root/reference/src/backend/cuda/kernel/scan_dim.cuh:19-169.)
constexpr int kDsRowsPerThread = 8;
constexpr int kDsY = 8;
constexpr int kDsRows = kDsRowsPerThread * kDsY;   // rows of the scanned axis per tile

template<typename Ti, typename To>
struct DimScanArgs {
    const Ti *in;
    To *out;
    dim_t d0;
    dim_t len, sd, osd;               // scanned extent; its input / output strides
    dim_t na, sa, osa;                // batch dim a
    dim_t nb, sb, osb;                // batch dim b
    dim_t nxb;                        // dim-0 slices of 32 x VEC elements
    dim_t nchains;                    // nxb * na * nb
    dim_t ntiles;                     // nchains * splits
    int inclusive;
    uint gen;                         // generation of this launch's descriptors
    unsigned long long *descs;        // [ntiles][32 * VEC]
    unsigned long long *ticket;       // monotone counter
    unsigned long long ticket_base;
};

template<int OP, typename Ti, typename To, int VEC>
__global__ void __launch_bounds__(32 * kDsY) dim_scan_kernel(const DimScanArgs<Ti, To> a) {
    static_assert(sizeof(To) == 4, "packed 64-bit descriptors carry a 4-byte value");
    typedef Rule<OP, Ti, To> R;
    constexpr int RPT = kDsRowsPerThread;
    constexpr int W = 32 * VEC;
    __shared__ To sm[kDsY][W + 1];
    __shared__ To sm_carry[W];
    __shared__ long long sm_tile;
    const int lane = threadIdx.x, wy = threadIdx.y;
    const uint tag = a.gen << 3;
    auto bits_of = [](To v) -> uint { uint b; memcpy(&b, &v, 4); return b; };
    auto val_of = [](uint b) -> To { To v; memcpy(&v, &b, 4); return v; };

    while (true) {
        __syncthreads();
        if (lane == 0 && wy == 0) sm_tile = (long long)(atomicAdd(a.ticket, 1ULL) - a.ticket_base);
        __syncthreads();
        const long long tile = sm_tile;
        if (tile >= a.ntiles) return;
        const dim_t split = tile / a.nchains, chain = tile % a.nchains;
        const dim_t xb = chain % a.nxb, batch = chain / a.nxb;
        const dim_t ia = batch % a.na, ib = batch / a.na;
        const dim_t x0 = (xb * 32 + lane) * VEC;
        const bool live = x0 < a.d0;       // d0 is a multiple of VEC when VEC > 1
        const Ti *in = a.in + ia * a.sa + ib * a.sb + x0;
        To *out = a.out + ia * a.osa + ib * a.osb + x0;
        const dim_t rs = split * kDsRows + (dim_t)wy * RPT;

        // ---- load the thread's RPT rows x VEC elements, scan down the rows
        To x[RPT][VEC];
        if (live) {
            Pack<Ti, VEC> v[RPT];
#pragma unroll
            for (int j = 0; j < RPT; j++)
                if (rs + j < a.len) v[j] = load_pack<Ti, VEC>(in + (rs + j) * a.sd);
#pragma unroll
            for (int j = 0; j < RPT; j++)
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    const To e = (rs + j < a.len) ? R::pre(v[j].e[k]) : R::init();
                    x[j][k] = (j == 0) ? e : R::join(x[j - 1][k], e);
                }
        } else {
#pragma unroll
            for (int j = 0; j < RPT; j++)
#pragma unroll
                for (int k = 0; k < VEC; k++) x[j][k] = R::init();
        }
#pragma unroll
        for (int k = 0; k < VEC; k++) sm[wy][lane * VEC + k] = x[RPT - 1][k];
        __syncthreads();

        // ---- the first warp owns the tile's descriptors: aggregate, look-back, inclusive prefix
        if (wy == 0) {
            unsigned long long *mine = a.descs + (size_t)tile * W;
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                const int xi = lane + 32 * k;             // descriptor words are lane-contiguous: coalesced
                To agg = sm[0][xi];
#pragma unroll
                for (int y = 1; y < kDsY; y++) agg = R::join(agg, sm[y][xi]);
                To carry = R::init();
                if (split > 0) {
                    st_relaxed64(mine + xi, ((unsigned long long)bits_of(agg) << 32) | tag | TILE_AGGREGATE);
                    To acc = R::init();
                    bool first = true;
                    for (dim_t p = split - 1;; p--) {
                        const unsigned long long *d = a.descs + (size_t)(p * a.nchains + chain) * W + xi;
                        unsigned long long w;
                        unsigned long long spins = 0;
                        while (true) {
                            w = ld_relaxed64(d);
                            if (((uint)w & ~7u) == tag && ((uint)w & 7u) != TILE_INVALID) break;
                            if (++spins > (1ull << 28)) __trap();
                        }
                        const To v = val_of((uint)(w >> 32));
                        acc = first ? v : R::join(v, acc);      // v covers EARLIER rows
                        first = false;
                        if (((uint)w & 7u) == TILE_PREFIX) break;
                    }
                    carry = acc;
                    agg = R::join(carry, agg);
                }
                st_relaxed64(mine + xi, ((unsigned long long)bits_of(agg) << 32) | tag | TILE_PREFIX);
                sm_carry[xi] = carry;
            }
        }
        __syncthreads();

        // ---- carry + the rows of the warps above, then the stores
        if (live) {
            To before[VEC];
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                const int xi = lane * VEC + k;
                To b = (split > 0) ? sm_carry[xi] : R::init();     // R::init() joins as the identity, as in the serial loop
                for (int y = 0; y < kDsY; y++)
                    if (y < wy) b = R::join(b, sm[y][xi]);
                before[k] = b;
            }
#pragma unroll
            for (int j = 0; j < RPT; j++) {
                if (rs + j >= a.len) break;
                Pack<To, VEC> o;
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    if (a.inclusive) o.e[k] = R::join(before[k], x[j][k]);
                    else o.e[k] = (j == 0) ? before[k] : R::join(before[k], x[j - 1][k]);
                }
                store_pack<To, VEC>(out + (rs + j) * a.osd, o);
            }
        }
    }
}

}  // namespace afb
