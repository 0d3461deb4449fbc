// pipe_kernels.cuh -- the software-pipelined, TMA-in / TMA-out single-pass scan of libafb200
// (4- and 8-byte values, by key or not, and the promoting scans of 1- / 2-byte inputs).
//
// Why it exists (profiles/r01_summary.md): the first persistent chained scan (chain_kernels.cuh) kept every warp of
// a CTA on the CTA barrier while warp 0 walked the look-back chain -- ncu showed 65-70 % of the stall samples on
// `barrier`, 10-13 us per tile, 15-28 % of DRAM peak -- and spent 22-60 thread-instructions per element on
// per-slot shuffle scans.  On a B200 whose HBM is saturated every synchronous L2 round trip (descriptor poll, ticket
// atomic) costs 1700-4000 cycles, so the design rule here is: NO compute warp ever waits on one.
//
//   * ONE CTA per SM, warp-specialised: 12-16 compute warps + a publisher warp + two walker warps + a one-thread DMA
//     warp, talking through mbarriers only.  Tiles are drawn by monotone ticket (a tile only ever waits for tiles with
//     smaller tickets, whose CTAs are already running), staged into a ring of shared-memory stages by bulk-async copies
//     (TMA, mbarrier complete_tx) three tiles ahead, transformed IN PLACE in shared memory and written back by
//     bulk-async stores: no thread issues a global load/store on the fast path.
//   * blocked register layout: a thread owns EPT consecutive elements (EPT*size bytes is an ODD number of 16-byte
//     vectors, so the 128-bit shared-memory accesses of a quarter-warp are conflict-free): one serial scan per
//     thread + ONE shuffle scan per warp per tile (~6-14 instructions per element instead of 22-60).
//   * two passes, D = 3 tiles apart: pass 1 of tile i (local scan in place, warp totals) is followed by pass 2 of tile
//     i-3 (add the value carried into the warp's chunk), so the tile's prefix has three tile periods to arrive.
//   * publisher: the moment the warp totals of a tile exist, its AGGREGATE goes out as packed 64-bit word(s)
//     (value | generation | head flag | status; relaxed store, no fences, never cleared between launches).
//   * walkers instead of a chain of prefixes: every CTA keeps its own running fold over ALL published aggregates in
//     ticket order (walk_fold: coalesced 256-byte descriptor reads, 8 in flight).  The fold of the foreign tiles
//     between two own tiles depends on nothing this CTA computes, so the two walkers overlap it; only the cheap chain
//     R_j = R_{j-1} (+) total_{j-1} (+) F_j is handed from walker to walker through shared memory.  Fixed fold order:
//     bitwise reproducible floating-point scans.
//
// Semantics: /root/reference/src/backend/cpu/kernel/scan.hpp:37-72, scan_by_key.hpp:43-86.
#pragma once

#include <type_traits>

#include "chain_kernels.cuh"

namespace afb {

// ------------------------------------------------------------------------------------------------
// more bulk-copy / mbarrier PTX
// ------------------------------------------------------------------------------------------------
AFB_DI void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes) : "memory");
}
AFB_DI void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template<int N> AFB_DI void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
AFB_DI void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
AFB_DI void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

constexpr int kPipeWindow = 8;   // 32-descriptor chunks per walker poll (window = 256 tiles)

// ------------------------------------------------------------------------------------------------
// tile descriptors of the pipelined kernels: W = sizeof(value) / 4 packed 64-bit words per tile, each word
//   [63:32] 32 bits of the value   [31:3] generation   [2] segment-head flag   [1:0] status
// Every word carries the generation tag, so a reader that sees the tag of THIS launch in all W words has a
// consistent value (a descriptor is written exactly once per launch): relaxed accesses, no fences.
// ------------------------------------------------------------------------------------------------
template<typename S, typename To, bool KEYED> struct PipeCodec {
    typedef typename S::State State;
    static constexpr int W = (int)sizeof(To) / 4;
    static_assert(sizeof(To) % 4 == 0 && W >= 1 && W <= 2, "packed descriptors hold 4- or 8-byte values");
    AFB_DI static void pack(State s, uint32_t tag, uint32_t status, unsigned long long (&w)[W]) {
        uint32_t flag = 0;
        if constexpr (KEYED) flag = s.flag ? 4u : 0u;
        union { To t; uint32_t u[W]; } x;
        x.t = S::value(s);
#pragma unroll
        for (int j = 0; j < W; j++) w[j] = ((unsigned long long)x.u[j] << 32) | tag | flag | status;
    }
    AFB_DI static State unpack(const unsigned long long (&w)[W]) {
        union { To t; uint32_t u[W]; } x;
#pragma unroll
        for (int j = 0; j < W; j++) x.u[j] = (uint32_t)(w[j] >> 32);
        if constexpr (KEYED) return State{x.t, (int)((w[0] >> 2) & 1u)};
        else return x.t;
    }
    AFB_DI static bool valid(const unsigned long long (&w)[W], uint32_t tag) {
        bool ok = true;
#pragma unroll
        for (int j = 0; j < W; j++) ok = ok && ((((uint32_t)w[j]) & ~7u) == tag) && ((((uint32_t)w[j]) & 3u) != TILE_INVALID);
        return ok;
    }
    AFB_DI static void load(const unsigned long long *pd, long long tile, unsigned long long (&w)[W]) {
        if constexpr (W == 1) {
            w[0] = ld_relaxed64(pd + tile);
        } else {
            asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(pd + 2 * tile) : "memory");
        }
    }
    AFB_DI static void store(unsigned long long *pd, long long tile, const unsigned long long (&w)[W]) {
        if constexpr (W == 1) {
            st_relaxed64(pd + tile, w[0]);
        } else {
            asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(pd + 2 * tile), "l"(w[0]), "l"(w[1]) : "memory");
        }
    }
};

// ------------------------------------------------------------------------------------------------
// walker: R <- R (+) fold of the published aggregates of tiles [pos, end), executed by one full warp.
// Every CTA keeps its own running fold over ALL tiles of the launch (in ticket order) instead of chasing a
// chain of published prefixes: a descriptor only ever carries a tile AGGREGATE, every warp-wide load is one
// coalesced 256-byte request, up to 8 requests are in flight per poll, and the fold order is fixed (bitwise
// reproducible results for floating-point scans).  With a saturated memory system a synchronous L2 round trip
// costs 2000-4000 cycles (measured); the prefix chain of the classic decoupled look-back made the distance to
// the nearest resolved prefix grow with that latency (2-3 x 148 tiles), the walker never waits for anybody's
// prefix, only for aggregates that were published a tile period earlier.
// ------------------------------------------------------------------------------------------------
template<typename S, typename Codec>
AFB_DI typename S::State walk_fold(const unsigned long long *pd, long long pos, long long end, uint tag, int lane,
                                   typename S::State R, unsigned long long *npolls = nullptr, int nwin = kPipeWindow) {
    typedef typename S::State State;
    constexpr int W = Codec::W;
    while (pos < end) {
        unsigned long long d[kPipeWindow][W];
        uint need = (1u << nwin) - 1u;   // chunks that still have unpublished entries (warp-uniform)
        const long long wend = (pos + 32LL * nwin < end) ? pos + 32LL * nwin : end;
        uint spins = 0;
#pragma unroll
        for (int q = 0; q < kPipeWindow; q++)
#pragma unroll
            for (int j = 0; j < W; j++) d[q][j] = 0ULL;
        while (true) {
#pragma unroll
            for (int q = 0; q < kPipeWindow; q++) {
                const long long t = pos + 32 * q + lane;
                if (((need >> q) & 1u) && t < wend) Codec::load(pd, t, d[q]);
            }
            uint still = 0u;
#pragma unroll
            for (int q = 0; q < kPipeWindow; q++) {
                const long long t = pos + 32 * q + lane;
                const bool ok = (t >= wend) || Codec::valid(d[q], tag);
                if (!__all_sync(0xffffffffu, ok)) still |= 1u << q;
            }
            if (npolls) (*npolls)++;
            need = still;
            if (!need) break;
            if (++spins > kSpinLimit) __trap();
            __nanosleep(100);   // the missing aggregates are at least a publish + L2 round trip away
        }
#pragma unroll
        for (int q = 0; q < kPipeWindow; q++) {
            if (pos + 32 * q < wend) {                        // warp-uniform
                const long long t = pos + 32 * q + lane;
                State v = (t < wend) ? Codec::unpack(d[q]) : S::identity();
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {         // lane + dd holds a LATER tile than lane
                    State o = shfl_down(v, dd);
                    if (lane + dd < 32) v = S::join(v, o);
                }
                R = S::join(R, shfl_idx(v, 0));
            }
        }
        pos = wend;
    }
    return R;
}

// ------------------------------------------------------------------------------------------------
// pipe_scan_kernel
// ------------------------------------------------------------------------------------------------
// Warp roles of the (single) CTA of an SM:
//   warps 0..NW-1      compute: pass 1 of tile i (local scan in place, warp totals), then pass 2 of tile i-D
//                      (add the value carried into the warp's chunk); mbarriers only, never a CTA-wide barrier
//   warp NW            publisher: folds the warp totals of tile j and publishes the tile aggregate at once
//   warps NW+1, NW+2   walkers: walker k resolves this CTA's tiles j = k (mod 2): fold of the foreign aggregates since
//                      the previous own tile (one loaded L2 round trip, 3000+ cycles), chained with the running fold
//                      handed over by the other walker; writes the per-warp carries pass 2 waits for
//   warp NW+3          DMA (one thread): draws tile tickets two ahead, keeps the bulk loads STAGES-D-1 tiles ahead of
//                      pass 1 and writes finished tiles back with bulk stores
// Measured alternatives (profiles/r01_summary.md): one or three walkers, D = 2, a classic PREFIX chain, 128-bit
// descriptor loads, delayed polls -- all slower or neutral.
constexpr int kPipeWarps = 12;                      // compute warps (default; a template parameter of the kernel)
constexpr int kPipeWalkers = 2;                     // walker warps; walker k resolves the tiles j = k (mod kPipeWalkers)
constexpr int kPipeLB = 1 + kPipeWalkers;           // publisher warp + walker warps
constexpr int pipe_threads(int nw) { return 32 * (nw + kPipeLB + 1); }   // + publisher + walkers + DMA warp
constexpr int kPipeLag = 3;                         // default D: pass 2 runs D tiles behind pass 1 (template parameter)
constexpr int kPipeRing = 8;                        // per-tile barrier / scratch ring (> D + 1)

template<typename Ti, typename To, typename Tk>
struct PipeScanArgs {
    const Ti *in;
    const Tk *key;
    To *out;
    ColGeom g;
    dim_t ks1, ks2, ks3;
    dim_t tiles_per_col;
    dim_t ntiles;
    int inclusive;
    uint gen;
    unsigned long long *descs;
    unsigned long long *ticket;
    unsigned long long ticket_base;
    const To *carry;
    const Tk *carry_key;
    void *summary;
    unsigned long long *debug;   // optional per-CTA cycle counters (AFB200_PIPE_DEBUG=1), 16 words per CTA
    int ahead;                   // bulk loads in flight ahead of pass 1 (<= STAGES - D - 1)
    int window;                  // 32-descriptor chunks per walker poll (1..kPipeWindow)
};

template<bool KEYED, int EPT, int STAGES, int NWARPS = kPipeWarps, int DLAG = kPipeLag, int ES = 4, int KS = 4> struct PipeLayout {
    static constexpr int TILE = NWARPS * 32 * EPT;
    static constexpr int KPAD = 16 / KS;                            // key elements kept in front of the tile (16 bytes)
    static constexpr int KSTAGES = KEYED ? STAGES - DLAG : 0;           // keys are dead after pass 1
    static constexpr size_t val_bytes = (size_t)TILE * ES;
    static constexpr size_t key_stage_bytes = ((size_t)TILE * KS + 16 + 127) / 128 * 128;
    static constexpr size_t total = val_bytes * STAGES + key_stage_bytes * KSTAGES + 128;
    static_assert(val_bytes % 128 == 0, "value stages must stay 128-byte aligned");
};

template<int OP, typename Ti, typename To, bool KEYED, typename Tk, int EPT, int STAGES, int NWARPS, int DLAG>
__global__ void __launch_bounds__(pipe_threads(NWARPS), 1) pipe_scan_kernel(const PipeScanArgs<Ti, To, Tk> a) {
    constexpr int ES = (int)sizeof(To), KS = (int)sizeof(Tk);
    constexpr int EV = 16 / ES;                 // elements per 16-byte vector
    // NARROW: promoting scans (u8/s8/b8/s16/u16 -> 32 bit, never by key).  The input tile is staged in the ring the keys
    // use otherwise (dead after pass 1); pass 1 converts and writes the 32-bit results into the value stage.
    constexpr bool NARROW = sizeof(Ti) < sizeof(To);
    constexpr int IS = (int)sizeof(Ti);
    static_assert((sizeof(Ti) == sizeof(To) || (NARROW && !KEYED && ES == 4 && (EPT * IS) % 4 == 0)) && (ES == 4 || ES == 8) &&
                  (KS == 4 || KS == 8), "pipe scan: 4- or 8-byte results, in place or promoted from 1- / 2-byte inputs");
    static_assert(EPT % EV == 0 && ((EPT / EV) & 1) == 1 && EPT <= 32, "EPT*ES bytes must be an odd number of 16-byte vectors");
    constexpr bool KVEC = (EPT * KS) % 16 == 0; // the thread's keys can be read as 16-byte vectors
    constexpr int KEV = 16 / KS;
    typedef Rule<OP, Ti, To> R;
    typedef SOps<R, To, KEYED> S;
    typedef typename S::State State;
    typedef PipeCodec<S, To, KEYED> Codec;
    typedef PipeLayout<KEYED || NARROW, EPT, STAGES, NWARPS, DLAG, (int)sizeof(To), NARROW ? (int)sizeof(Ti) : (int)sizeof(Tk)> Lay;
    constexpr int TILE = Lay::TILE;
    constexpr int NV = EPT / EV;
    constexpr int WCHUNK = 32 * EPT;
    constexpr int KPAD = Lay::KPAD;
    constexpr int NW = NWARPS, D = DLAG, RING = kPipeRing;
    static_assert(D >= 1 && D + 2 <= kPipeRing && D < STAGES - 1, "lag out of range");
    constexpr int KST = Lay::KSTAGES > 0 ? Lay::KSTAGES : 1;

    // (a pointer that provably derives from the __shared__ array keeps its address space: the round trip through
    //  uintptr_t used for manual alignment made every stage access a generic LD.E/ST.E with 64-bit address math)
    extern __shared__ __align__(128) unsigned char pipe_smem[];
    unsigned char *val_base = pipe_smem;
    unsigned char *key_base = val_base + Lay::val_bytes * STAGES;
    __shared__ uint64_t full_bar[STAGES];
    __shared__ uint64_t p1_bar[RING], lb_bar[RING], p2_bar[RING];
    __shared__ long long sm_tile_id[STAGES];
    __shared__ int sm_valid[STAGES];             // valid elements of the tile; TILE: staged / stored by bulk copies
    __shared__ int sm_first[STAGES];             // 1: the tile is the first of its column
    __shared__ State sm_wtot[RING][NWARPS];
    __shared__ uint sm_wfh[RING][NWARPS];            // first segment head inside the warp's chunk (WCHUNK: none)
    __shared__ To sm_wp[RING][NWARPS];               // value carried into each warp's chunk
    __shared__ State sm_R[RING];                 // running fold handed from walker to walker
    __shared__ uint64_t r_bar[RING];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&full_bar[s], 1);
#pragma unroll
        for (int b = 0; b < RING; b++) {
            mbar_init(&p1_bar[b], NW);
            mbar_init(&lb_bar[b], 1);
            mbar_init(&p2_bar[b], NW);
            mbar_init(&r_bar[b], 1);
        }
        mbar_arrive(&r_bar[0]);   // tile 0 has no predecessor: its hand-over phase is completed here
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    ColGeom kg = a.g;
    kg.s1 = a.ks1; kg.s2 = a.ks2; kg.s3 = a.ks3;
    const uint tag = a.gen << 3;

    if (warp == NW + kPipeLB) {
        // =========================================================================== DMA warp (one thread)
        if (lane != 0) return;
        // returns the RAW counter value: subtracting ticket_base right here would make the thread wait for the atomic's
        // round trip (~2500 cycles under load, measured) on every draw instead of two tiles later, when the ticket is used
        auto draw = [&]() -> unsigned long long { return atomicAdd(a.ticket, 1ULL); };
        long long pub = 0;        // tiles published so far; tile n uses value stage n % STAGES, key stage n % KST
        long long n_live = -1;    // number of real tiles of this CTA, known once the end ticket has been drawn
        auto publish = [&](long long n, long long tile) {
            const int s = (int)(n % STAGES);
            sm_tile_id[s] = tile;
            if (tile >= a.ntiles) {
                sm_valid[s] = 0;
                mbar_arrive(&full_bar[s]);
                return;
            }
            const dim_t col = tile / a.tiles_per_col, tcol = tile % a.tiles_per_col;
            const dim_t tile_lo = tcol * TILE;
            const dim_t left = a.g.len - tile_lo;
            const int valid = left >= TILE ? TILE : (int)left;
            sm_valid[s] = valid;
            sm_first[s] = tcol == 0 ? 1 : 0;
            if (valid == TILE) {
                fence_proxy_async();
                if constexpr (NARROW) {
                    // the narrow input goes to the side ring; the value stage is written by pass 1
                    mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(TILE * IS));
                    bulk_g2s(key_base + (size_t)(n % KST) * Lay::key_stage_bytes, a.in + col_offset(a.g, col) + tile_lo,
                             (uint32_t)(TILE * IS), &full_bar[s]);
                    return;
                }
                const uint32_t kb = KEYED ? (uint32_t)(TILE * KS + (tile_lo > 0 ? 16 : 0)) : 0u;
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)Lay::val_bytes + kb);
                bulk_g2s(val_base + (size_t)s * Lay::val_bytes, a.in + col_offset(a.g, col) + tile_lo, (uint32_t)Lay::val_bytes, &full_bar[s]);
                if (KEYED) {
                    const Tk *ksrc = a.key + col_offset(kg, col) + tile_lo;
                    unsigned char *kdst = key_base + (size_t)(n % KST) * Lay::key_stage_bytes;
                    if (tile_lo > 0) bulk_g2s(kdst, ksrc - KPAD, kb, &full_bar[s]);   // key[tile_lo-1] lands in the pad
                    else bulk_g2s(kdst + 16, ksrc, kb, &full_bar[s]);
                }
            } else {
                mbar_arrive(&full_bar[s]);   // ragged last tile of a column: the compute warps copy it themselves
            }
        };
        // Tickets are drawn two tiles ahead of their use (the atomic's round trip is ~1 us under load) and at most
        // two per tile period, never a whole ring at once: a CTA that drew STAGES consecutive tickets would own
        // STAGES consecutive tiles, and since every tile waits for its predecessor's aggregate the CTAs would start
        // one after the other.
        unsigned long long tk0 = draw(), tk1 = draw();
        bool drawing = true;
        unsigned long long dma_dbg[4] = {0, 0, 0, 0};
        const long long dma_start = clock64();
        auto publish_next = [&]() {
            const long long t = (long long)(tk0 - a.ticket_base);
            tk0 = tk1;
            // keys: stage pub % KST is free once pass 1 of tile pub - KST has read it (pass 1 arrives on p1_bar after
            // its key loads; tiles are processed in order, so this wait can never block the tile it waits for)
            const long long w0 = clock64();
            if ((KEYED || NARROW) && pub >= KST) mbar_wait(&p1_bar[(pub - KST) % RING], (uint32_t)(((pub - KST) / RING) & 1));
            dma_dbg[2] += clock64() - w0;
            publish(pub, t);
            pub++;
            if (t >= a.ntiles) { drawing = false; n_live = pub - 1; }
            else tk1 = draw();
        };
        publish_next();
        if (drawing) publish_next();
        for (long long j = 0;; j++) {
            if (j >= D) {
                const long long n = j - D;            // pass 2 of tile n is the next to finish
                if (!drawing && n >= n_live) break;
                long long w1 = clock64();
                mbar_wait(&p2_bar[n % RING], (uint32_t)((n / RING) & 1));
                dma_dbg[0] += clock64() - w1;
                const int sp = (int)(n % STAGES);
                if (sm_valid[sp] == TILE) {
                    const long long tp = sm_tile_id[sp];
                    const dim_t col = tp / a.tiles_per_col, tcol = tp % a.tiles_per_col;
                    bulk_s2g(a.out + col * a.g.len + tcol * TILE, val_base + (size_t)sp * Lay::val_bytes, (uint32_t)Lay::val_bytes);
                }
                bulk_commit();
                w1 = clock64();
                bulk_wait_read<1>();                  // stores of tiles <= n-1 have finished reading shared memory
                dma_dbg[1] += clock64() - w1;
            }
            // tile m may be published once the store of tile m - STAGES has drained: m - STAGES <= j - D - 1
            const long long lim = j + a.ahead;
            for (int burst = 0; drawing && burst < 2 && pub <= lim; burst++) publish_next();
        }
        bulk_wait_all();
        if (a.debug) {
            dma_dbg[3] = clock64() - dma_start;
            for (int q = 0; q < 4; q++) a.debug[blockIdx.x * 16 + 4 + q] = dma_dbg[q];
        }
        return;
    }

    if (warp == NW) {
        // =========================================================================== publisher warp
        // publishes the aggregate of every tile the moment its warp totals exist (the walkers of all other CTAs
        // are waiting for it); never blocks on anything but this CTA's own pass 1
        for (long long j = 0;; j++) {
            const int s = (int)(j % STAGES), slot = (int)(j % RING);
            mbar_wait(&full_bar[s], (uint32_t)((j / STAGES) & 1));
            const long long tile = sm_tile_id[s];
            if (tile >= a.ntiles) break;
            mbar_wait(&p1_bar[slot], (uint32_t)((j / RING) & 1));
            State st = (lane < NW) ? sm_wtot[slot][lane] : S::identity();
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {     // ordered reduction: lane + dd holds a LATER warp
                State o = shfl_down(st, dd);
                if (lane + dd < 32) st = S::join(st, o);
            }
            if (lane == 0) {
                unsigned long long w[Codec::W];
                Codec::pack(st, tag, TILE_AGGREGATE, w);
                Codec::store(a.descs, tile, w);
            }
        }
        return;
    }

    if (warp > NW) {
        // =========================================================================== walker warps
        // Walker k resolves this CTA's tiles j = k (mod kPipeWalkers).  The expensive part -- F_j, the fold of the
        // foreign aggregates published between this CTA's tiles j-1 and j (one loaded L2 round trip, 3000+ cycles,
        // more than a tile period) -- does not depend on anything this CTA computes, so the walkers overlap it;
        // only the cheap chain R_j = R_{j-1} (+) total_{j-1} (+) F_j is handed from walker to walker through
        // shared memory (sm_R / r_bar).
        const int wk = warp - NW - 1;
        unsigned long long dbg[4] = {0, 0, 0, 0};
        const long long t_start = clock64();
        long long prev = -1;                    // this CTA's previous tile
        for (long long j = 0;; j++) {
            const int s = (int)(j % STAGES), slot = (int)(j % RING);
            mbar_wait(&full_bar[s], (uint32_t)((j / STAGES) & 1));   // every walker follows the tile stream
            const long long tile = sm_tile_id[s];
            if (tile >= a.ntiles) break;
            if ((int)(j % kPipeWalkers) != wk) {
                prev = tile;
                continue;
            }
            // ---- F: foreign tiles between the previous own tile and this one, not reaching back over a column start
            const long long c0 = clock64();
            const long long cstart = tile - tile % a.tiles_per_col;
            const bool chained = prev >= cstart;                    // the previous own tile sits in the same column
            const long long from = chained ? prev + 1 : cstart;
            State F = walk_fold<S, Codec>(a.descs, from, tile, tag, lane, S::identity(), a.debug ? &dbg[2] : nullptr, a.window);
            dbg[1] += clock64() - c0;
            // ---- R: everything in this column up to and including the previous own tile
            State excl = F;
            if (chained) {
                mbar_wait(&r_bar[slot], (uint32_t)((j / RING) & 1));
                excl = S::join(sm_R[slot], F);
            } else if (cstart == 0 && a.carry) {
                excl = S::join(S::lift(*a.carry, false), F);   // sharded scans: the carry sits in front of tile 0
            }
            // ---- own tile: per-warp carries for pass 2
            mbar_wait(&p1_bar[slot], (uint32_t)((j / RING) & 1));
            State st = (lane < NW) ? sm_wtot[slot][lane] : S::identity();
#pragma unroll
            for (int dd = 1; dd < NW; dd <<= 1) {
                State o = shfl_up(st, dd);
                if (lane >= dd) st = S::join(o, st);
            }
            const State tile_total = shfl_idx(st, NW - 1);
            State wexc = shfl_up(st, 1);
            if (lane == 0) wexc = S::identity();
            const dim_t tcol = tile % a.tiles_per_col;
            if (a.summary) {
                if (KEYED) {
                    uint fh = (lane < NW) ? sm_wfh[slot][lane] : (uint)WCHUNK;
                    uint hp = fh < (uint)WCHUNK ? (uint)(lane * WCHUNK) + fh : 0xffffffffu;
                    hp = __reduce_min_sync(0xffffffffu, hp);
                    if (lane == 0 && hp != 0xffffffffu)
                        atomicMin(reinterpret_cast<unsigned long long *>((char *)a.summary + 24),
                                  (unsigned long long)(tcol * TILE + hp));
                }
                if (lane == 0 && tcol == a.tiles_per_col - 1) {
                    const State fin = S::join(excl, tile_total);
                    *reinterpret_cast<Slot<To> *>(a.summary) = Slot<To>{S::value(fin)};
                    uint flag = 0;
                    if constexpr (KEYED) flag = fin.flag ? 1u : 0u;
                    *reinterpret_cast<uint *>((char *)a.summary + 16) = flag;
                }
            }
            if (lane < NW) sm_wp[slot][lane] = S::value(S::join(excl, wexc));
            // hand the running fold to the walker of this CTA's next tile
            if (lane == 0) sm_R[(slot + 1) % RING] = S::join(excl, tile_total);
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&lb_bar[slot]);
                mbar_arrive(&r_bar[(slot + 1) % RING]);
            }
            prev = tile;
            dbg[3]++;
        }
        if (a.debug && lane == 0 && wk == 0) {
            dbg[0] = clock64() - t_start;
            for (int q = 0; q < 4; q++) a.debug[blockIdx.x * 16 + q] = dbg[q];
        }
        return;
    }

    // =============================================================================== compute warps
    // (ring positions are kept incrementally in 32-bit registers: 64-bit `%` and `/` by run-time values cost ~100
    //  instructions each and used to dominate the per-tile instruction count of these warps)
    const int t = warp * 32 + lane;          // 0 .. NW*32-1
    const int base = t * EPT;                // first element of this thread inside the tile
    uint fh[D + 1];                          // first head of the warp's chunk in tiles i .. i-D (registers: constant indices)
#pragma unroll
    for (int q = 0; q <= D; q++) fh[q] = WCHUNK;
    unsigned long long cw_full = 0, cw_lb = 0, cw_p1 = 0, cw_p2 = 0;
    bool ended = false;
    uint n_live = 0;                         // valid once `ended`
    int s1 = 0, ks1 = 0;                     // value / key stage of pass 1's tile
    uint ph1 = 0;                            // phase parity of full_bar[s1]
    int s2 = 0;                              // value stage of pass 2's tile
    for (uint i = 0;; i++) {
        long long k0 = clock64();
        if (!ended) {
            const int s = s1, b = (int)(i & (RING - 1));
            mbar_wait(&full_bar[s], ph1);
            cw_full += clock64() - k0;
            k0 = clock64();
            const long long tile = sm_tile_id[s];
            if (tile >= a.ntiles) {
                ended = true;
                n_live = i;
            } else {
                // ---------------------------------------------------------------- pass 1 of tile i
                const int valid = sm_valid[s];
                auto pass1 = [&](auto rag_c) {
                constexpr bool RAG = decltype(rag_c)::value;
                const bool col_start = sm_first[s] != 0;     // first tile of its column
                To *sv = reinterpret_cast<To *>(val_base + (size_t)s * Lay::val_bytes);
                Tk *sk = reinterpret_cast<Tk *>(key_base + (size_t)ks1 * Lay::key_stage_bytes) + KPAD;
                if (valid != TILE) {
                    // ragged last tile of a column: plain cooperative copy, padded to a full tile with the op's identity
                    // (and the last key repeated), so that the one fast path below serves every tile: elements past the
                    // end cannot influence stored outputs, and identities keep the tile total exact
                    const dim_t col = tile / a.tiles_per_col, tcol = tile % a.tiles_per_col;
                    const dim_t tile_lo = tcol * TILE;
                    const Ti *gin = a.in + col_offset(a.g, col) + tile_lo;
                    if constexpr (NARROW) {
                        for (int e = t; e < valid; e += NW * 32) sv[e] = R::pre(ldg_stream<Ti>(gin + e));   // converted here, RAG below
                    } else {
                        for (int e = t; e < valid; e += NW * 32) reinterpret_cast<Ti *>(sv)[e] = ldg_stream<Ti>(gin + e);
                    }
                    for (int e = valid + t; e < TILE; e += NW * 32) sv[e] = R::init();   // in place: Ti and To are the same type
                    if (KEYED) {
                        const Tk *gk = a.key + col_offset(kg, col) + tile_lo;
                        for (int e = t; e < valid; e += NW * 32) sk[e] = __ldg(gk + e);
                        const Tk last = __ldg(gk + valid - 1);
                        for (int e = valid + t; e < TILE; e += NW * 32) sk[e] = last;
                        if (t == 0 && tile_lo > 0) sk[-1] = __ldg(gk - 1);
                    }
                    named_bar_sync(1, NW * 32);
                }
                To xin[EPT];   // the thread's elements after the op's pre-transform
                if constexpr (NARROW && !RAG) {
                    constexpr int NWORD = EPT * IS / 4;
                    const uint32_t *si = reinterpret_cast<const uint32_t *>(key_base + (size_t)ks1 * Lay::key_stage_bytes) + (size_t)t * NWORD;
                    uint32_t wds[NWORD];
#pragma unroll
                    for (int j = 0; j < NWORD; j++) wds[j] = si[j];
#pragma unroll
                    for (int k = 0; k < EPT; k++) xin[k] = R::pre(reinterpret_cast<const Ti *>(wds)[k]);
                } else if constexpr (NARROW && RAG) {
#pragma unroll
                    for (int j = 0; j < NV; j++) {
                        const Vec16 q = *reinterpret_cast<const Vec16 *>(sv + base + EV * j);
#pragma unroll
                        for (int k = 0; k < EV; k++) xin[EV * j + k] = reinterpret_cast<const To *>(&q)[k];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NV; j++) {
                        const Vec16 q = *reinterpret_cast<const Vec16 *>(reinterpret_cast<const Ti *>(sv) + base + EV * j);
#pragma unroll
                        for (int k = 0; k < EV; k++) xin[EV * j + k] = R::pre(reinterpret_cast<const Ti *>(&q)[k]);
                    }
                }
                uint hb = 0u;   // bit k: element k of this thread starts a segment
                if (KEYED) {
                    Tk kk[EPT];
                    if constexpr (KVEC) {
#pragma unroll
                        for (int j = 0; j < EPT / KEV; j++) {
                            // keys sit 16 bytes into their buffer: still 16-byte aligned
                            const Vec16 q = *reinterpret_cast<const Vec16 *>(sk + base + KEV * j);
#pragma unroll
                            for (int k = 0; k < KEV; k++) kk[KEV * j + k] = reinterpret_cast<const Tk *>(&q)[k];
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < EPT; k++) kk[k] = sk[base + k];
                    }
                    Tk kprev;
                    if (base > 0 || !col_start) kprev = sk[base - 1];
                    else kprev = a.carry_key ? *a.carry_key : kk[0];   // position 0 of a column: a head only against the carry key
#pragma unroll
                    for (int k = 0; k < EPT; k++) {
                        const bool head = kk[k] != (k == 0 ? kprev : kk[k - 1]);
                        hb |= (head ? 1u : 0u) << k;
                    }
                    if (RAG) {
                        const int nv = valid - base;   // heads past the end do not exist
                        hb = nv <= 0 ? 0u : (nv >= 32 ? hb : (hb & ((1u << nv) - 1u)));
                    }
                }
                // ---- serial scan inside the thread
                To x[EPT];
                To run = R::init();
#pragma unroll
                for (int k = 0; k < EPT; k++) {
                    const To e = xin[k];
                    if (k == 0) run = e;
                    else if (KEYED) run = ((hb >> k) & 1u) ? e : R::join(run, e);
                    else run = R::join(run, e);
                    x[k] = run;
                }
                // ---- one shuffle scan per warp
                State inc;
                if constexpr (KEYED) inc = State{run, hb != 0u ? 1 : 0};
                else inc = run;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    State o = shfl_up(inc, dd);
                    if (lane >= dd) inc = S::join(o, inc);
                }
                State exc = shfl_up(inc, 1);
                if (lane == 0) exc = S::identity();
                if (lane == 31) sm_wtot[b][warp] = inc;
                const To ev = S::value(exc);
                if (KEYED) {
                    uint pos = hb ? (uint)(lane * EPT + (__ffs(hb) - 1)) : (uint)WCHUNK;
                    pos = __reduce_min_sync(0xffffffffu, pos);
                    fh[0] = pos;
                    if (lane == 0) sm_wfh[b][warp] = pos;
                }
                // the warp total is all the publisher / walker need: let them work while the outputs are formed
                __syncwarp();
                if (lane == 0) mbar_arrive(&p1_bar[b]);
                // ---- local results (everything except the value carried into the warp's chunk), in place
                To o[EPT];
#pragma unroll
                for (int k = 0; k < EPT; k++) {
                    // open(k): no segment head among elements 0..k of this thread
                    if (a.inclusive) {
                        const bool open = !KEYED || (hb & ((2u << k) - 1u)) == 0u;
                        o[k] = open ? R::join(ev, x[k]) : x[k];
                    } else {
                        if (k == 0) o[k] = ev;
                        else {
                            const bool open = !KEYED || (hb & ((1u << k) - 1u)) == 0u;
                            o[k] = open ? R::join(ev, x[k - 1]) : x[k - 1];
                        }
                        if (KEYED && ((hb >> k) & 1u)) o[k] = R::init();
                    }
                }
#pragma unroll
                for (int j = 0; j < NV; j++) {
                    Vec16 q;
#pragma unroll
                    for (int k = 0; k < EV; k++) reinterpret_cast<To *>(&q)[k] = o[EV * j + k];
                    *reinterpret_cast<Vec16 *>(sv + base + EV * j) = q;
                }
                };
                if (NARROW && valid != TILE) pass1(std::true_type{});   // ragged promoted tile: converted by the copy above
                else pass1(std::false_type{});
                cw_p1 += clock64() - k0;
                if (++s1 == STAGES) { s1 = 0; ph1 ^= 1u; }
                if (++ks1 == KST) ks1 = 0;
            }
        }
        if (i >= (uint)D) {
            // ---------------------------------------------------------------- pass 2 of tile n = i - D
            const uint n = i - D;
            if (ended && n >= n_live) break;
            const int sp = s2, bp = (int)(n & (RING - 1));
            k0 = clock64();
            mbar_wait(&lb_bar[bp], (n / RING) & 1u);
            cw_lb += clock64() - k0;
            k0 = clock64();
            const To wp = sm_wp[bp][warp];
            const int valid = sm_valid[sp];
            To *sv = reinterpret_cast<To *>(val_base + (size_t)sp * Lay::val_bytes);
            // elements of this thread that are still "open" towards the warp's left edge
            int nadd = EPT;
            if (KEYED) {
                const int d = (int)fh[D] - lane * EPT;
                nadd = d < 0 ? 0 : (d > EPT ? EPT : d);
            }
            if (valid == TILE) {
                if (nadd > 0) {
#pragma unroll
                    for (int j = 0; j < NV; j++) {
                        Vec16 q = *reinterpret_cast<const Vec16 *>(sv + base + EV * j);
#pragma unroll
                        for (int k = 0; k < EV; k++) {
                            To &e = reinterpret_cast<To *>(&q)[k];
                            if (!KEYED || EV * j + k < nadd) e = R::join(wp, e);
                        }
                        *reinterpret_cast<Vec16 *>(sv + base + EV * j) = q;
                    }
                }
                fence_proxy_async();   // make the generic-proxy writes (both passes) visible to the bulk store
            } else {
                const long long tp = sm_tile_id[sp];
                const dim_t col = tp / a.tiles_per_col, tcol = tp % a.tiles_per_col;
                To *gout = a.out + col * a.g.len + tcol * TILE;
#pragma unroll
                for (int k = 0; k < EPT; k++) {
                    if (base + k < valid) {
                        To e = sv[base + k];
                        if (k < nadd) e = R::join(wp, e);
                        gout[base + k] = e;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&p2_bar[bp]);
            cw_p2 += clock64() - k0;
            if (++s2 == STAGES) s2 = 0;
        } else if (ended && n_live == 0) {
            break;
        }
#pragma unroll
        for (int q = D; q >= 1; q--) fh[q] = fh[q - 1];
    }
    if (a.debug && t == 0) {
        a.debug[blockIdx.x * 16 + 8] = cw_full;
        a.debug[blockIdx.x * 16 + 9] = cw_lb;
        a.debug[blockIdx.x * 16 + 10] = cw_p1;
        a.debug[blockIdx.x * 16 + 11] = cw_p2;
    }
}

}  // namespace afb
