// chain_kernels.cuh -- the persistent, TMA-fed single-pass ("chained") scan of libafb200.
//
// Replaces, for aligned long columns, the one-tile-per-CTA look-back kernel of scan_kernels.cuh, whose
// ncu capture on B200 showed 48-64 % of all stall samples on the CTA barrier and 15-19 % of DRAM peak
// (profiles/r01_first_pass.md): the ticket atomic, the loads and the look-back of a tile were strictly
// serial inside a CTA, so HBM idled whenever a tile waited on its predecessors.
//
// Design (B200):
//   * persistent CTAs (grid = SMs x resident CTAs) of 8 warps; thread 0 doubles as the producer:
//   * the producer draws tile tickets (64-bit monotone counter: a tile only ever waits for tiles with
//     SMALLER tickets, which are held by CTAs that are already running -> no co-residency assumption) and
//     stages each tile with ONE bulk-async copy (cp.async.bulk / TMA, mbarrier complete_tx) per operand into
//     a ring of shared-memory stages, running a full ring ahead of the consumers: HBM stays busy while a
//     tile's look-back is in flight;
//   * consumers pull 16-byte vectors from shared memory (conflict-free), scan in registers + warp shuffles,
//     resolve the tile prefix by decoupled look-back and write 16-byte coalesced streaming stores;
//   * descriptors of 4-byte states are ONE packed 64-bit word (value | generation | flag | status), read and
//     written with relaxed 64-bit accesses: no fences on the look-back path.  Wider states keep the
//     (status, aggregate, inclusive) record with release/acquire.
//   * tiles carry >= 48 KB of traffic so that (tiles per microsecond) x (L2 round trip) stays within one or two
//     32-wide look-back windows.
// Semantics: /root/reference/src/backend/cpu/kernel/scan.hpp:37-72, scan_by_key.hpp:43-86.
#pragma once

#include "scan_kernels.cuh"

namespace afb {

// ------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy PTX
// ------------------------------------------------------------------------------------------------
AFB_DI uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
AFB_DI void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
AFB_DI void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
AFB_DI void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: a waiting warp is parked by the hardware until the phase completes (or the
// hint expires) instead of burning issue slots in a spin loop -- the service warps of the pipelined kernels spend
// most of their life waiting and were executing ~1/3 of all instructions of pipe_where_kernel (ncu, r01c).
AFB_DI bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x100000u) : "memory");
    return ok != 0;
}
AFB_DI unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
AFB_DI void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    // a phase that does not complete within 20 s of wall time is a protocol bug (a bulk copy that never lands, a
    // role that died): fail loudly instead of hanging the device.  (Every failed try parks the warp for up to ~1 ms.)
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 40000000000LL) __trap();   // ~20 s at 1.97 GHz
    }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completes on `bar`
AFB_DI void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
AFB_DI void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kChainConsumers = 256;           // 8 warps; thread 0 also drives the bulk-copy ring
constexpr int kChainThreads = kChainConsumers;
constexpr int kChainStages = 3;

// ------------------------------------------------------------------------------------------------
// packed descriptors: [63:32] value bits, [31:3] generation, [2] segment-head flag, [1:0] status
// ------------------------------------------------------------------------------------------------
template<typename T> AFB_DI uint32_t bits_of(T v) {
    static_assert(sizeof(T) == 4, "packed descriptors hold 4-byte values");
    union { T t; uint32_t u; } x;
    x.t = v;
    return x.u;
}
template<typename T> AFB_DI T from_bits(uint32_t u) {
    union { T t; uint32_t u; } x;
    x.u = u;
    return x.t;
}

template<typename S, typename To, bool KEYED> struct DescCodec {
    typedef typename S::State State;
    AFB_DI static unsigned long long pack(State s, uint32_t tag /* generation << 3 */, uint32_t status) {
        uint32_t flag = 0;
        if constexpr (KEYED) flag = s.flag ? 4u : 0u;
        return ((unsigned long long)bits_of<To>(S::value(s)) << 32) | tag | flag | status;
    }
    AFB_DI static State unpack(unsigned long long d) {
        const To v = from_bits<To>((uint32_t)(d >> 32));
        if constexpr (KEYED) return State{v, (int)((d >> 2) & 1u)};
        else return v;
    }
};

template<typename Ti, typename To, typename Tk>
struct ChainScanArgs {
    const Ti *in;
    const Tk *key;
    To *out;
    ColGeom g;
    dim_t ks1, ks2, ks3;
    dim_t tiles_per_col;
    dim_t ntiles;
    int inclusive;
    uint gen;                            // generation tag of this launch
    void *descs;                         // packed: u64[ntiles]; generic: TileDesc<State>[ntiles]
    unsigned long long *ticket;          // monotone 64-bit counter, never reset
    unsigned long long ticket_base;      // its value when this launch starts
    const To *carry;
    const Tk *carry_key;
    void *summary;
};

template<typename Ti, typename Tk, bool KEYED, int TILE> struct ChainSmem {
    static constexpr size_t in_bytes = (size_t)TILE * sizeof(Ti);
    static constexpr size_t key_bytes = KEYED ? (size_t)TILE * sizeof(Tk) : 0;
    static constexpr size_t stage_bytes = in_bytes + key_bytes;
    static constexpr size_t total = stage_bytes * kChainStages + 128;  // + alignment slack
};

template<int OP, typename Ti, typename To, bool KEYED, typename Tk, int U>
__global__ void __launch_bounds__(kChainThreads, 2) chain_scan_kernel(const ChainScanArgs<Ti, To, Tk> a) {
    typedef Rule<OP, Ti, To> R;
    typedef SOps<R, To, KEYED> S;
    typedef typename S::State State;
    constexpr int E = ScanVec<Ti, To>::value;
    constexpr int TILE = kChainConsumers * U * E;
    constexpr bool PACKED = sizeof(To) == 4;
    typedef ChainSmem<Ti, Tk, KEYED, TILE> Lay;
    typedef DescCodec<S, To, KEYED> Codec;
    typedef TileDesc<State> Desc;

    extern __shared__ __align__(128) unsigned char chain_smem[];   // derived pointers stay in the shared address space
    unsigned char *stage_base = chain_smem;
    __shared__ uint64_t full_bar[kChainStages];
    __shared__ long long sm_tile_id[kChainStages];
    __shared__ int sm_staged[kChainStages];     // 1: the tile's operands are in shared memory
    __shared__ State sm_warp[2][8];     // double-buffered by tile parity: two CTA barriers per tile suffice
    __shared__ uint sm_head[2][8];
    __shared__ State sm_tile_prefix[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kChainStages; s++) {
            mbar_init(&full_bar[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    ColGeom kg = a.g;
    kg.s1 = a.ks1; kg.s2 = a.ks2; kg.s3 = a.ks3;

    // ---- thread 0 doubles as the producer: it draws tile tickets one ring ahead and stages each tile with one
    //      bulk-async copy per operand.  The ticket for the NEXT refill is always drawn one iteration early so
    //      that the atomic's round trip overlaps the current tile.
    long long next_ticket = 0;
    bool drawing = true;  // false once this CTA has drawn a ticket past the end (exactly one per CTA)
    auto publish = [&](int s, long long tile) {
        sm_tile_id[s] = tile;
        if (tile >= a.ntiles) {
            sm_staged[s] = 0;
            mbar_arrive(&full_bar[s]);
            return;
        }
        const dim_t col = tile / a.tiles_per_col, tcol = tile % a.tiles_per_col;
        const dim_t tile_lo = tcol * TILE;
        const bool whole = tile_lo + TILE <= a.g.len;
        sm_staged[s] = whole ? 1 : 0;
        if (whole) {
            unsigned char *dst = stage_base + (size_t)s * Lay::stage_bytes;
            fence_proxy_async();  // the consumers' generic reads of this stage come before the async writes
            mbar_arrive_expect_tx(&full_bar[s], (uint32_t)Lay::stage_bytes);
            bulk_g2s(dst, a.in + col_offset(a.g, col) + tile_lo, (uint32_t)Lay::in_bytes, &full_bar[s]);
            if (KEYED) bulk_g2s(dst + Lay::in_bytes, a.key + col_offset(kg, col) + tile_lo, (uint32_t)Lay::key_bytes, &full_bar[s]);
        } else {
            mbar_arrive(&full_bar[s]);  // ragged last tile of a column: consumers load it straight from global
        }
    };
    // raw counter values: ticket_base is subtracted where the ticket is used, so nothing waits on the atomic's round trip
    auto draw = [&]() -> unsigned long long { return atomicAdd(a.ticket, 1ULL); };
    // Tickets are drawn one tile period apart (two per iteration while the ring fills), never a whole ring at once: a CTA
    // owning kChainStages consecutive tiles makes the CTAs start one after the other (see pipe_kernels.cuh).
    unsigned long long raw_next = 0;
    long long pub = 0;    // tiles published by this CTA; tile n lives in stage n % kChainStages
    auto publish_next = [&]() {
        const long long t = (long long)(raw_next - a.ticket_base);
        publish((int)(pub % kChainStages), t);
        pub++;
        if (t >= a.ntiles) drawing = false;
        else raw_next = draw();
    };
    if (tid == 0) {
        raw_next = draw();
        publish_next();
        if (drawing) publish_next();
    }

    // ================================================================== tile loop
    const uint tag = a.gen << 3;
    for (uint it = 0;; it++) {
        const int s = it % kChainStages;
        mbar_wait(&full_bar[s], (it / kChainStages) & 1);
        const long long tile = sm_tile_id[s];
        if (tile >= a.ntiles) break;
        const bool staged = sm_staged[s] != 0;
        const dim_t col = tile / a.tiles_per_col, tcol = tile % a.tiles_per_col;
        const dim_t tile_lo = tcol * TILE;
        const Ti *in = a.in + col_offset(a.g, col);
        To *out = a.out + col * a.g.len;
        const Tk *key = KEYED ? a.key + col_offset(kg, col) : nullptr;
        const Ti *sin = reinterpret_cast<const Ti *>(stage_base + (size_t)s * Lay::stage_bytes);
        const Tk *skey = reinterpret_cast<const Tk *>(stage_base + (size_t)s * Lay::stage_bytes + Lay::in_bytes);

        // ---- operands -> registers
        Pack<Ti, E> v[U];
        Pack<Tk, E> kv[U];
        Tk kfirst = Tk();  // key just before the tile
        const int warp_off = warp * (U * 32 * E);
        if (KEYED && tid == 0 && tile_lo > 0) kfirst = __ldg(key + tile_lo - 1);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int q = warp_off + (u * 32 + lane) * E;  // offset inside the tile
            if (staged) {
                v[u] = *reinterpret_cast<const Pack<Ti, E> *>(sin + q);
                if (KEYED) kv[u] = *reinterpret_cast<const Pack<Tk, E> *>(skey + q);
            } else {
#pragma unroll
                for (int k = 0; k < E; k++) {
                    if (tile_lo + q + k < a.g.len) {
                        v[u].e[k] = ldg_stream<Ti>(in + tile_lo + q + k);
                        if (KEYED) kv[u].e[k] = __ldg(key + tile_lo + q + k);
                    }
                }
            }
        }
        // key preceding each vector: previous lane's last key, the previous slot's lane 31, or shared / global memory
        Tk kprev[U];
        if (KEYED) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int q = warp_off + (u * 32 + lane) * E;
                const Tk from_lane = shfl_up(kv[u].e[E - 1], 1);
                const Tk from_slot = (u > 0) ? shfl_idx(kv[u > 0 ? u - 1 : 0].e[E - 1], 31) : Tk();  // all lanes take part
                if (lane > 0) kprev[u] = from_lane;
                else if (u > 0) kprev[u] = from_slot;
                else if (q == 0) kprev[u] = kfirst;
                else if (staged) kprev[u] = skey[q - 1];
                else kprev[u] = (tile_lo + q < a.g.len) ? __ldg(key + tile_lo + q - 1) : Tk();
            }
        }
        // ---- scan inside each vector
        State x[U][E];
        State vtot[U];
        uint heads[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const dim_t p0 = tile_lo + warp_off + (u * 32 + lane) * E;
            State run = S::identity();
            heads[u] = 0u;
#pragma unroll
            for (int k = 0; k < E; k++) {
                State st = S::identity();
                if (staged || p0 + k < a.g.len) {
                    bool head = false;
                    if (KEYED) {
                        if (p0 + k == 0) head = a.carry_key ? (kv[u].e[0] != *a.carry_key) : false;
                        else head = kv[u].e[k] != ((k == 0) ? kprev[u] : kv[u].e[k - 1]);
                        if (head) heads[u] |= 1u << k;
                    }
                    st = S::lift(R::pre(v[u].e[k]), head);
                }
                run = (k == 0) ? st : S::join(run, st);
                x[u][k] = run;
            }
            vtot[u] = run;
        }

        // ---- warp level
        State slot_excl[U];
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
        const int pb = it & 1;
        if (lane == 0) sm_warp[pb][warp] = warp_total;
        if (KEYED && a.summary) {
            uint first = 0xffffffffu;
#pragma unroll
            for (int u = U - 1; u >= 0; u--)
                if (heads[u]) first = (uint)(warp_off + (u * 32 + lane) * E + (__ffs(heads[u]) - 1));
            first = __reduce_min_sync(0xffffffffu, first);
            if (lane == 0) sm_head[pb][warp] = first;
        }
        __syncthreads();
        // every warp has its operands in registers: refill this stage with the tile drawn one iteration ago
        if (tid == 0) {
            // the stages of the tiles <= it are free: tile `pub` may go out while pub - kChainStages <= it
            for (int burst = 0; drawing && burst < 2 && pub <= (long long)it + kChainStages; burst++) publish_next();
        }
        State warp_excl = S::identity(), tile_total = S::identity();
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const State t = sm_warp[pb][w];
            if (w < warp) warp_excl = S::join(warp_excl, t);
            tile_total = S::join(tile_total, t);
        }

        // ---- decoupled look-back (warp 0)
        if (warp == 0) {
            State excl = S::identity();
            unsigned long long *pd = reinterpret_cast<unsigned long long *>(a.descs);
            Desc *gd = reinterpret_cast<Desc *>(a.descs);
            if (tcol == 0) {
                if (a.carry) excl = S::lift(*a.carry, false);
                if (lane == 0) {
                    const State incl = S::join(excl, tile_total);
                    if constexpr (PACKED) {
                        st_relaxed64(&pd[tile], Codec::pack(incl, tag, TILE_PREFIX));
                    } else {
                        gd[tile].inclusive.v = incl;
                        __threadfence();
                        st_release(&gd[tile].status, (a.gen << 2) | TILE_PREFIX);
                    }
                }
            } else {
                if (lane == 0) {
                    if constexpr (PACKED) {
                        st_relaxed64(&pd[tile], Codec::pack(tile_total, tag, TILE_AGGREGATE));
                    } else {
                        gd[tile].aggregate.v = tile_total;
                        __threadfence();
                        st_release(&gd[tile].status, (a.gen << 2) | TILE_AGGREGATE);
                    }
                }
                long long look = tile - 1;
                const long long first = tile - tcol;
                while (true) {
                    const long long mine = look - lane;
                    const bool valid = mine >= first;
                    State pv = S::identity();
                    bool is_prefix = false;
                    if constexpr (PACKED) {
                        unsigned long long d = 0;
                        if (valid) {
                            { uint spins = 0; do { d = ld_relaxed64(&pd[mine]); if (++spins > kSpinLimit) __trap(); } while ((((uint32_t)d) & ~7u) != tag || (d & 3u) == TILE_INVALID); }
                            pv = Codec::unpack(d);
                        }
                        is_prefix = valid && (d & 3u) == TILE_PREFIX;
                    } else {
                        uint st = 0;
                        if (valid) {
                            { uint spins = 0; do { st = ld_acquire(&gd[mine].status); if (++spins > kSpinLimit) __trap(); } while ((st & ~3u) != (a.gen << 2) || (st & 3u) == TILE_INVALID); }
                        }
                        is_prefix = valid && (st & 3u) == TILE_PREFIX;
                    }
                    const uint pmask = __ballot_sync(0xffffffffu, is_prefix);
                    const int stop = pmask ? (__ffs(pmask) - 1) : 31;
                    if constexpr (!PACKED) {
                        if (valid && lane <= stop) pv = load_cg(is_prefix ? &gd[mine].inclusive : &gd[mine].aggregate);
                    }
                    if (!(valid && lane <= stop)) pv = S::identity();
                    // ordered fold: lane+d holds an EARLIER tile than lane
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        State o = shfl_down(pv, d);
                        if (lane + d < 32) pv = S::join(o, pv);
                    }
                    pv = shfl_idx(pv, 0);
                    excl = S::join(pv, excl);
                    if (pmask) break;
                    look -= 32;
                }
                if (lane == 0) {
                    const State incl = S::join(excl, tile_total);
                    if constexpr (PACKED) {
                        st_relaxed64(&pd[tile], Codec::pack(incl, tag, TILE_PREFIX));
                    } else {
                        gd[tile].inclusive.v = incl;
                        __threadfence();
                        st_release(&gd[tile].status, (a.gen << 2) | TILE_PREFIX);
                    }
                }
            }
            if (lane == 0) {
                sm_tile_prefix[pb] = excl;
                if (a.summary) {
                    if (KEYED) {
                        uint first = 0xffffffffu;
#pragma unroll
                        for (int w = 0; w < 8; w++) first = min(first, sm_head[pb][w]);
                        if (first != 0xffffffffu)
                            atomicMin(reinterpret_cast<unsigned long long *>((char *)a.summary + 24),
                                      (unsigned long long)(tile_lo + first));
                    }
                    if (tcol == a.tiles_per_col - 1) {
                        const State fin = S::join(excl, tile_total);
                        *reinterpret_cast<Slot<To> *>(a.summary) = Slot<To>{S::value(fin)};
                        uint flag = 0;
                        if constexpr (KEYED) flag = fin.flag ? 1u : 0u;
                        *reinterpret_cast<uint *>((char *)a.summary + 16) = flag;
                    }
                }
            }
        }
        __syncthreads();
        const State before_warp = S::join(sm_tile_prefix[pb], warp_excl);

        // ---- write
#pragma unroll
        for (int u = 0; u < U; u++) {
            const dim_t p0 = tile_lo + warp_off + (u * 32 + lane) * E;
            if (!staged && p0 >= a.g.len) continue;
            const State before_vec = S::join(before_warp, slot_excl[u]);
            Pack<To, E> o;
#pragma unroll
            for (int k = 0; k < E; k++) {
                State r;
                if (a.inclusive) r = S::join(before_vec, x[u][k]);
                else r = (k == 0) ? before_vec : S::join(before_vec, x[u][k - 1]);
                To val = S::value(r);
                if (KEYED && !a.inclusive && ((heads[u] >> k) & 1u)) val = R::init();
                o.e[k] = val;
            }
            if (staged || p0 + E <= a.g.len) store_pack<To, E>(out + p0, o);
            else {
#pragma unroll
                for (int k = 0; k < E; k++)
                    if (p0 + k < a.g.len) out[p0 + k] = o.e[k];
            }
        }
    }
}

}  // namespace afb
