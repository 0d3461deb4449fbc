// reduce_kernels.cuh -- the three reduction kernels of libafb200, generic over a Policy
// (plain value reductions and (value,index) reductions share them).
//
// What they replace in the reference (behaviour studied, design is new):
//   reduce_first_kernel / reduce_all_kernel / reduce_dim_kernel
//       /root/reference/src/backend/cuda/kernel/reduce.hpp:33-111,190-382  (scalar loads, smem tree, 2nd launch)
//   ireduceFirst / ireduceDim + host-side finish
//       /root/reference/src/backend/cuda/kernel/ireduce.cuh:19-252, kernel/ireduce.hpp:163-258
//
// B200 design:
//   * span_kernel    -- dim-0 and whole-array reductions of LONG columns.  Each CTA streams a contiguous
//                       chunk with 128-bit L1-bypassing loads (UNROLL independent loads in flight per
//                       thread), reduces with warp shuffles, then
//                         - CTAs of one thread-block CLUSTER combine through distributed shared memory,
//                         - cluster leaders publish one partial each and the LAST arriver (self-resetting
//                           ticket) folds the partials in a fixed order -> single launch, deterministic.
//   * warp_cols_kernel -- dim-0 reductions of SHORT columns: a group of 1..32 lanes per column.
//   * dim_kernel     -- reductions along dim>0: lanes (x VEC-wide vectors) run along the contiguous dim 0 so
//                       every warp request is a full 128..512-byte line; the reduced axis is split over the
//                       warps of a CTA and over the CTAs of a cluster, combined through (distributed) smem.
#pragma once

#include <cooperative_groups.h>

#include "device_ops.cuh"
#include "peer_comm.cuh"

namespace afb {
namespace cg = cooperative_groups;

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// ------------------------------------------------------------------------------------------------
// policies
// ------------------------------------------------------------------------------------------------
struct ValParams {
    int change_nan;
    double nanval;
};

template<int OP, typename Ti, typename To> struct ValPolicy {
    static constexpr bool kSplitAcc = false;   // one accumulator: more registers cost sums occupancy (measured)
    typedef Ti In;
    typedef typename compute_of<To>::type Acc;
    typedef Rule<OP, Ti, Acc> R;
    typedef ValParams Params;
    struct Out { To *val; };
    __host__ __device__ static Out host_out(const Out &, void *rec) { return Out{(To *)rec}; }   // `out`, re-pointed at a host record
    AFB_DI static Acc init() { return R::init(); }
    AFB_DI static Acc load(const Params &p, In x, uint) {
        Acc v = R::pre(x);
        if (p.change_nan) v = nan_swap<Acc>(v, p.nanval);
        return v;
    }
    AFB_DI static Acc join(Acc a, Acc b) { return R::join(a, b); }
    AFB_DI static Acc globalize(Acc a, uint) { return a; }   // shard partial -> record exchanged between ranks
    AFB_DI static void store(const Out &o, size_t i, Acc a) { o.val[i] = conv<To, Acc>(a); }
    // fold one 16-byte vector of inputs whose first element has index idx0
    AFB_DI static Acc fold16(const Params &p, Acc acc, const Vec16 &v, uint idx0) {
        constexpr int V = VecN<In>::value;
        if constexpr (OP == OP_NOTZERO && sizeof(In) == 1 && same_t<Acc, uint>::value) {
            // 4 bytes at a time: per-byte "!= 0" mask, then a population count
#pragma unroll
            for (int w = 0; w < 4; w++) acc += __popc(__vcmpne4(v.w[w], 0u)) >> 3;
            return acc;
        } else {
            const In *e = reinterpret_cast<const In *>(&v);
#pragma unroll
            for (int k = 0; k < V; k++) acc = join(acc, load(p, e[k], idx0 + k));
            return acc;
        }
    }
};

// Types whose own ordering coincides with the order of the `double` key the reference compares
// (cpu/kernel/ireduce.hpp:21-27): the conversion to double is exact and monotonic for them.  64-bit integers
// (lossy conversion), b8 (compares v > 0) and complex (magnitudes) stay on the generic path.
template<typename T> struct NativeOrder { static constexpr bool value = false; };
#define AFB_NATIVE_ORDER(T)                                                        \
    template<> struct NativeOrder<T> {                                              \
        static constexpr bool value = true;                                         \
        typedef T Cmp;                                                              \
        AFB_DI static T cmp(T v) { return v; }                                      \
    };
AFB_NATIVE_ORDER(float) AFB_NATIVE_ORDER(double) AFB_NATIVE_ORDER(int) AFB_NATIVE_ORDER(uint)
AFB_NATIVE_ORDER(short) AFB_NATIVE_ORDER(ushort) AFB_NATIVE_ORDER(schar) AFB_NATIVE_ORDER(uchar)
#undef AFB_NATIVE_ORDER
template<> struct NativeOrder<__half> {
    static constexpr bool value = true;
    typedef float Cmp;
    AFB_DI static float cmp(__half v) { return __half2float(v); }
};

struct ArgParams {
    int seed = 1;   // 0: this is not the first shard of a sharded array -- its element 0 is not the reduction's seed
};

template<int OP, typename T> struct ArgPolicy {
    static constexpr bool kSplitAcc = NativeOrder<T>::value;   // break the compare/select chain of the streaming path
    typedef T In;
    typedef Arg<T> Acc;
    typedef ArgParams Params;
    struct Out { T *val; uint *loc; };
    __host__ __device__ static Out host_out(const Out &, void *rec) { return Out{(T *)rec, (uint *)((char *)rec + 32)}; }   // value at +0, index at +32
    // Internally element i carries rank i+1, so that "no candidate yet" (rank 0 for min, UINT_MAX for max)
    // loses every tie against a real element -- distinct values can share an ordering key (b8 compares
    // `v > 0`, 64-bit integers compare as double, complex compares magnitudes).
    AFB_DI static Acc init() {
        if (OP == OP_MIN) return Acc{Lim<T>::hi(), 0u};
        return Acc{Lim<T>::lo(), 0xffffffffu};
    }
    AFB_DI static Acc load(const Params &p, In x, uint idx) {
        if (is_nan<T>(x)) {
            // NaN candidates never win; only the seed (position 0 of the WHOLE array) turns into the identity at index 0
            if (idx == 0 && p.seed) return Acc{OP == OP_MIN ? Lim<T>::hi() : Lim<T>::lo(), 1u};
            return init();
        }
        return Acc{x, idx + 1u};
    }
    // sharded arrays: ranks become global by adding the shard's offset; "no candidate" stays what it is
    AFB_DI static Acc globalize(Acc a, uint base) {
        if (a.idx != (OP == OP_MIN ? 0u : 0xffffffffu)) a.idx += base;
        return a;
    }
    AFB_DI static Acc join(Acc a, Acc b) {
        const double ka = order_key<T>(a.val), kb = order_key<T>(b.val);
        bool take_b;
        if (OP == OP_MIN) take_b = (kb < ka) || (kb == ka && b.idx > a.idx);   // min keeps the LAST equal position
        else take_b = (kb > ka) || (kb == ka && b.idx < a.idx);                // max keeps the FIRST
        return take_b ? b : a;
    }
    AFB_DI static void store(const Out &o, size_t i, Acc a) {
        o.val[i] = a.val;
        uint idx = a.idx - 1u;
        if constexpr (NativeOrder<T>::value) {
            // the streaming fast path below never records an element that merely EQUALS the identity (max) and
            // leaves "no candidate" when nothing was recorded.  In both cases the reference's answer is 0:
            // a result equal to the identity means every element is the identity or NaN, and element 0 -- the
            // seed, turned into the identity if NaN -- wins (max: first occurrence; min: the non-strict
            // compare below has already recorded the last one, if any).
            if (OP == OP_MAX && !(order_key<T>(a.val) > order_key<T>(Lim<T>::lo()))) idx = 0u;
            if (OP == OP_MIN && a.idx == 0u) idx = 0u;
        }
        o.loc[i] = idx;
    }
    AFB_DI static Acc fold16(const Params &p, Acc acc, const Vec16 &v, uint idx0) {
        constexpr int V = VecN<In>::value;
        const In *e = reinterpret_cast<const In *>(&v);
        if constexpr (NativeOrder<T>::value) {
            // Elements arrive in increasing index order and the type's own compare orders exactly like the
            // reference's (double) key, so one compare per element suffices: strict '>' keeps the FIRST maximum,
            // '<=' keeps the LAST minimum; NaN fails both.  (2 instead of ~12 instructions per element: the
            // generic path made imax compute-bound, sm__throughput 81 % in profiles/r01a_first_pass.md.)
            typedef typename NativeOrder<T>::Cmp C;
            C best = NativeOrder<T>::cmp(acc.val);
            T bval = acc.val;
            uint bidx = acc.idx;
#pragma unroll
            for (int k = 0; k < V; k++) {
                const C x = NativeOrder<T>::cmp(e[k]);
                const bool take = (OP == OP_MAX) ? (x > best) : (x <= best);
                best = take ? x : best;
                bval = take ? e[k] : bval;
                bidx = take ? idx0 + k + 1u : bidx;
            }
            return Acc{bval, bidx};
        } else {
#pragma unroll
            for (int k = 0; k < V; k++) acc = join(acc, load(p, e[k], idx0 + k));
            return acc;
        }
    }
};

// Complex min / max: the reference keeps the running value on a magnitude tie
// (`abs(x) < abs(acc) ? x : acc`, cpu/math.cpp:30-40), i.e. the FIRST element of smallest/largest magnitude --
// and the identity itself if nothing beats it.  A parallel reduction reproduces that by carrying the position:
// ties go to the smaller position, the identity sits at position 0 and element i at position i+1.
template<int OP, typename T> struct FirstPolicy {
    static constexpr bool kSplitAcc = false;
    typedef T In;
    typedef Arg<T> Acc;
    typedef ValParams Params;
    struct Out { T *val; };
    __host__ __device__ static Out host_out(const Out &, void *rec) { return Out{(T *)rec}; }
    AFB_DI static T ident() { return OP == OP_MIN ? Lim<T>::hi() : Lim<T>::lo(); }
    AFB_DI static Acc init() { return Acc{ident(), 0u}; }
    AFB_DI static Acc load(const Params &, In x, uint idx) { return Acc{is_nan<T>(x) ? ident() : x, idx + 1u}; }
    AFB_DI static Acc join(Acc a, Acc b) {
        const auto ka = mag(a.val), kb = mag(b.val);
        bool take_b;
        if (OP == OP_MIN) take_b = (kb < ka) || (kb == ka && b.idx < a.idx);
        else take_b = (kb > ka) || (kb == ka && b.idx < a.idx);
        return take_b ? b : a;
    }
    AFB_DI static Acc globalize(Acc a, uint base) {
        if (a.idx != 0u) a.idx += base;
        return a;
    }
    AFB_DI static void store(const Out &o, size_t i, Acc a) { o.val[i] = a.val; }
    AFB_DI static Acc fold16(const Params &p, Acc acc, const Vec16 &v, uint idx0) {
        constexpr int V = VecN<In>::value;
        const In *e = reinterpret_cast<const In *>(&v);
#pragma unroll
        for (int k = 0; k < V; k++) acc = join(acc, load(p, e[k], idx0 + k));
        return acc;
    }
};

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
// a partial written to global memory: at least word-aligned, a whole number of words
template<typename T> struct alignas(alignof(T) < 4 ? 4 : alignof(T)) Slot {
    T v;
};
// L2-coherent load of a slot written by another SM
template<typename T> AFB_DI T load_cg(const Slot<T> *p) {
    constexpr int W = sizeof(Slot<T>) / 4;
    union U { Slot<T> t; uint32_t w[W]; AFB_DI U() {} } u;
    const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
#pragma unroll
    for (int i = 0; i < W; i++) u.w[i] = __ldcg(q + i);
    return u.t.v;
}

template<typename P> AFB_DI typename P::Acc warp_reduce(typename P::Acc v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = P::join(v, shfl_down(v, d));
    return v;
}

// result valid in thread 0; `sm` holds >= 32 Acc
template<typename P> AFB_DI typename P::Acc block_reduce(typename P::Acc v, typename P::Acc *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_reduce<P>(v);
    __syncthreads();  // protect `sm` against a previous use
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? sm[lane] : P::init();
        v = warp_reduce<P>(v);
    }
    return v;
}

// Every thread of the CTA folds its share of base[lo, hi).  idx_bias + i is the index reported for base[i].
template<typename P>
AFB_DI typename P::Acc reduce_span(const typename P::Params &prm, const typename P::In *__restrict__ base, dim_t lo,
                                   dim_t hi, uint idx_bias, int tid, int nthr) {
    typedef typename P::In In;
    typedef typename P::Acc Acc;
    constexpr int V = VecN<In>::value;
    Acc acc = P::init();
    if (hi <= lo) return acc;
    // scalar head up to the first 16-byte boundary
    const uintptr_t addr = reinterpret_cast<uintptr_t>(base + lo);
    dim_t head = (dim_t)(((16 - (addr & 15)) & 15) / sizeof(In));
    if (head > hi - lo) head = hi - lo;
    if (tid < head) acc = P::join(acc, P::load(prm, ldg_stream<In>(base + lo + tid), idx_bias + (uint)(lo + tid)));
    const dim_t vlo = lo + head;
    const dim_t nvec = (hi - vlo) / V;
    const Vec16 *vp = reinterpret_cast<const Vec16 *>(base + vlo);
    dim_t v = tid;
    if (P::kSplitAcc && v + (dim_t)(kUnroll - 1) * nthr < nvec) {
        // one accumulator per in-flight vector: kUnroll independent dependency chains instead of one chain through
        // 4 x 16 bytes (the index reductions' compare/select chain left DRAM at 5.97 TB/s against 6.75 for sums)
        Acc part[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) part[u] = P::init();
        for (; v + (dim_t)(kUnroll - 1) * nthr < nvec; v += (dim_t)kUnroll * nthr) {
            Vec16 r[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; u++) r[u] = ldg_stream16(vp + v + (dim_t)u * nthr);
#pragma unroll
            for (int u = 0; u < kUnroll; u++)
                part[u] = P::fold16(prm, part[u], r[u], idx_bias + (uint)(vlo + (v + (dim_t)u * nthr) * V));
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) acc = P::join(acc, part[u]);
    }
    for (; v + (dim_t)(kUnroll - 1) * nthr < nvec; v += (dim_t)kUnroll * nthr) {   // single chain (plain value reductions)
        Vec16 r[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) r[u] = ldg_stream16(vp + v + (dim_t)u * nthr);
#pragma unroll
        for (int u = 0; u < kUnroll; u++)
            acc = P::fold16(prm, acc, r[u], idx_bias + (uint)(vlo + (v + (dim_t)u * nthr) * V));
    }
    for (; v < nvec; v += nthr) acc = P::fold16(prm, acc, ldg_stream16(vp + v), idx_bias + (uint)(vlo + v * V));
    const dim_t tlo = vlo + nvec * V;
    if (tlo + tid < hi) acc = P::join(acc, P::load(prm, ldg_stream<In>(base + tlo + tid), idx_bias + (uint)(tlo + tid)));
    return acc;
}

struct ColGeom {
    dim_t len;          // elements per column (dims[0])
    dim_t ncols;        // dims[1]*dims[2]*dims[3]
    dim_t d1, d2;       // to unravel a column id
    dim_t s1, s2, s3;   // element strides of dims 1..3
};
AFB_DI dim_t col_offset(const ColGeom &g, dim_t c) {
    const dim_t j = c % g.d1, r = c / g.d1;
    const dim_t k = r % g.d2, l = r / g.d2;
    return j * g.s1 + k * g.s2 + l * g.s3;
}

// ------------------------------------------------------------------------------------------------
// span_kernel: long columns / whole array
// ------------------------------------------------------------------------------------------------
// mode "columns": one output per column; the column is cut into `bpc` chunks, one CTA each
//                 (grid = ncols * bpc; CTAs of a cluster share a column because csize divides bpc).
// mode "all":     one output for everything; the work items (column, chunk) are dealt round-robin to a
//                 persistent grid, every CTA keeps ONE running partial.
template<typename P>
struct SpanArgs {
    const typename P::In *in;
    ColGeom g;
    dim_t chunk;        // elements per (column, chunk) work item
    int bpc;            // chunks per column
    int all;            // 1: whole-array mode
    int csize;          // cluster size (1 = no cluster)
    typename P::Params prm;
    typename P::Out out;
    Slot<typename P::Acc> *partials;   // [ncols * bpc / csize]  or  [gridDim.x / csize] in "all" mode
    uint *tickets;               // [ncols] or [1]
    // "all" mode, optional: the final value also goes to a pinned host record and its flag is raised (no D2H copy, no
    // stream synchronisation on the host side: common.hpp HostRet)
    void *host_wire;
    uint host_seq;
    // "all" mode, optional (comm.world > 1): `in` is one shard of an array spread over several GPUs.  The thread that
    // holds the shard's partial exchanges it with the other ranks over peer memory (peer_comm.cuh) and folds the world's
    // partials in rank order BEFORE the stores above: every rank ends with the global result (comm_mode 0), or with the
    // fold of the LOWER ranks only -- the carry-in of a sharded scan (comm_mode 1; rank 0 gets the identity).
    PeerComm comm;
    uint idx_base;      // global index of the shard's first element (index reductions)
    int comm_mode;
};

// ONE thread: the result as the output type (value at +0, index at +32) -> the pinned wire record of the call
template<typename P>
AFB_DI void span_host_send(const SpanArgs<P> &a, typename P::Acc v) {
    alignas(16) uint w[16];
#pragma unroll
    for (int j = 0; j < 16; j++) w[j] = 0u;
    P::store(P::host_out(a.out, w), 0, v);
    host_ret_send(a.host_wire, a.host_seq, w);
}

// the cross-rank step of span_kernel: called by the whole first warp, `mine` is lane 0's; the result is valid in lane 0
template<typename P>
AFB_DI typename P::Acc span_exchange(const SpanArgs<P> &a, typename P::Acc mine, PeerRecord *sm) {
    typedef typename P::Acc Acc;
    constexpr int NW = (int)((sizeof(Acc) + 3) / 4);
    PeerRecord rec;
#pragma unroll
    for (int j = 0; j < 16; j++) rec.w[j] = 0u;
    rec_put(rec, P::globalize(mine, a.idx_base));
    peer_exchange_warp<NW>(a.comm, rec, sm);
    const int upto = a.comm_mode == 1 ? a.comm.rank : a.comm.world;
    Acc g = P::init();
    for (int r = 0; r < upto; r++) {
        const Acc v = rec_get<Acc>(sm[r]);
        g = (r == 0) ? v : P::join(g, v);
    }
    return g;
}

// 4-byte accumulators (the f32 / s32 / u32 sums, counts): 8 CTAs of 256 threads per SM = 32 registers per thread, full
// occupancy for the streaming loop; wider accumulators and the index reductions keep the registers they need
// (5 CTAs = 51 registers, 4 CTAs = 64: about what ptxas chose unprompted -- an explicit "1" makes it spend up to 92)
template<typename P> struct SpanMinCtas {
    static constexpr int value = sizeof(typename P::Acc) <= 4 ? 8 : sizeof(typename P::Acc) <= 8 ? 5 : 4;
};

template<typename P>
__global__ void __launch_bounds__(kThreads, SpanMinCtas<P>::value) span_kernel(const SpanArgs<P> a) {
    typedef typename P::Acc Acc;
    __shared__ Acc sm[32];
    __shared__ Acc sm_cluster[8];
    __shared__ int sm_last;
    __shared__ PeerRecord sm_peer[kPeerMax];   // sharded arrays: the world's records (peer_comm.cuh)
    const int tid = threadIdx.x;

    Acc acc = P::init();
    dim_t col = 0;
    if (a.all) {
        const dim_t nwork = a.g.ncols * a.bpc;
        for (dim_t w = blockIdx.x; w < nwork; w += gridDim.x) {
            const dim_t c = w / a.bpc, part = w % a.bpc;
            const dim_t lo = part * a.chunk, hi = min(lo + a.chunk, a.g.len);
            acc = P::join(acc, reduce_span<P>(a.prm, a.in + col_offset(a.g, c), lo, hi, (uint)(c * a.g.len), tid, kThreads));
        }
    } else {
        col = blockIdx.x / a.bpc;
        const dim_t part = blockIdx.x % a.bpc;
        const dim_t lo = part * a.chunk, hi = min(lo + a.chunk, a.g.len);
        acc = reduce_span<P>(a.prm, a.in + col_offset(a.g, col), lo, hi, 0u, tid, kThreads);
    }
    acc = block_reduce<P>(acc, sm);  // thread 0

    // ---- level 2: CTAs of one cluster combine through distributed shared memory
    uint nparts = a.all ? gridDim.x : (uint)a.bpc;  // partials per output before clustering
    uint my_part = a.all ? blockIdx.x : (uint)(blockIdx.x % a.bpc);
    if (a.csize > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        const unsigned rank = cluster.block_rank();
        if (tid == 0) {
            Acc *leader = cluster.map_shared_rank(sm_cluster, 0);
            leader[rank] = acc;
        }
        cluster.sync();
        if (rank != 0) return;
        if (tid == 0) {
            acc = sm_cluster[0];
            for (int r = 1; r < a.csize; r++) acc = P::join(acc, sm_cluster[r]);
        }
        nparts /= a.csize;
        my_part /= a.csize;
    }

    // ---- level 3: one partial per (cluster of) CTA(s); the last arriver folds them in index order
    if (nparts == 1) {
        if (a.comm.world > 1 && tid < 32) acc = span_exchange<P>(a, acc, sm_peer);
        if (tid == 0) {
            P::store(a.out, (size_t)col, acc);
            if (a.host_wire) span_host_send<P>(a, acc);
        }
        return;
    }
    Slot<Acc> *slots = a.partials + (a.all ? 0 : (size_t)col * nparts);
    uint *ticket = a.tickets + (a.all ? 0 : col);
    if (tid == 0) {
        slots[my_part].v = acc;
        __threadfence();
        const uint t = atomicAdd(ticket, 1u);
        sm_last = (t == nparts - 1);
    }
    __syncthreads();
    if (!sm_last) return;
    __threadfence();
    Acc fin = P::init();
    for (uint i = tid; i < nparts; i += kThreads) fin = P::join(fin, load_cg(slots + i));
    fin = block_reduce<P>(fin, sm);
    if (a.comm.world > 1 && tid < 32) fin = span_exchange<P>(a, fin, sm_peer);
    if (tid == 0) {
        P::store(a.out, (size_t)col, fin);
        *ticket = 0u;  // self-reset: the next launch on this stream finds it clean
        if (a.host_wire) span_host_send<P>(a, fin);
    }
}

// ------------------------------------------------------------------------------------------------
// warp_cols_kernel: short columns, G (power of two, <= 32) lanes per column
// ------------------------------------------------------------------------------------------------
template<typename P>
struct WarpColsArgs {
    const typename P::In *in;
    ColGeom g;
    int group;  // lanes per column
    typename P::Params prm;
    typename P::Out out;
};

template<typename P>
__global__ void __launch_bounds__(kThreads) warp_cols_kernel(const WarpColsArgs<P> a) {
    typedef typename P::Acc Acc;
    typedef typename P::In In;
    const int G = a.group;
    const int sub = threadIdx.x % G;
    const dim_t groups_per_block = kThreads / G;
    const dim_t ngroups_total = (dim_t)gridDim.x * groups_per_block;
    const dim_t iters = (a.g.ncols + ngroups_total - 1) / ngroups_total;  // uniform trip count: shuffles stay converged
    for (dim_t it = 0; it < iters; it++) {
        const dim_t c = (it * gridDim.x + blockIdx.x) * groups_per_block + threadIdx.x / G;
        Acc acc = P::init();
        if (c < a.g.ncols) {
            const In *base = a.in + col_offset(a.g, c);
            dim_t i = sub;
            for (; i + 3 * (dim_t)G < a.g.len; i += 4 * (dim_t)G) {
                const In x0 = ldg_stream<In>(base + i), x1 = ldg_stream<In>(base + i + G);
                const In x2 = ldg_stream<In>(base + i + 2 * G), x3 = ldg_stream<In>(base + i + 3 * G);
                acc = P::join(acc, P::load(a.prm, x0, (uint)i));
                acc = P::join(acc, P::load(a.prm, x1, (uint)(i + G)));
                acc = P::join(acc, P::load(a.prm, x2, (uint)(i + 2 * G)));
                acc = P::join(acc, P::load(a.prm, x3, (uint)(i + 3 * G)));
            }
            for (; i < a.g.len; i += G) acc = P::join(acc, P::load(a.prm, ldg_stream<In>(base + i), (uint)i));
        }
        for (int d = G >> 1; d > 0; d >>= 1) {
            Acc o = shfl_down(acc, d);
            if (sub + d < G) acc = P::join(acc, o);
        }
        if (sub == 0 && c < a.g.ncols) P::store(a.out, (size_t)c, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// dim_kernel: reduce along dim > 0
// ------------------------------------------------------------------------------------------------
// VEC consecutive elements (4, 8 or 16 bytes) in one load
template<typename In, int VEC> struct alignas(sizeof(In) * VEC >= 16 ? 16 : sizeof(In) * VEC) Pack {
    In e[VEC];
};
template<typename In, int VEC> AFB_DI Pack<In, VEC> load_pack(const In *p) {
    Pack<In, VEC> r;
    constexpr int B = (int)sizeof(In) * VEC;
    static_assert(B == 4 || B == 8 || B == 16 || VEC == 1, "pack must be 4, 8 or 16 bytes");
    if constexpr (VEC == 1) {
        r.e[0] = ldg_stream<In>(p);
    } else if constexpr (B == 16) {
        const Vec16 v = ldg_stream16(p);
        r = *reinterpret_cast<const Pack<In, VEC> *>(&v);
    } else if constexpr (B == 8) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        r = *reinterpret_cast<const Pack<In, VEC> *>(&v);
    } else {
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(p));
        r = *reinterpret_cast<const Pack<In, VEC> *>(&v);
    }
    return r;
}
// widest pack the dim kernels use: at most 4 elements and at most 16 bytes
template<typename In> struct DimVec { static constexpr int value = sizeof(In) >= 16 ? 1 : (sizeof(In) == 8 ? 2 : 4); };

struct DimGeom {
    dim_t d0;           // contiguous extent
    dim_t len, sd;      // reduced extent and its stride
    dim_t na, sa;       // first batch dim (extent, input stride)
    dim_t nb, sb;       // second batch dim
    dim_t oa, ob;       // output strides of the two batch dims (output stride along dim 0 is 1)
};

template<typename P>
struct DimArgs {
    const typename P::In *in;
    DimGeom g;
    int splits;         // CTAs along the reduced axis (= cluster size in y)
    typename P::Params prm;
    typename P::Out out;
    // two-stage split (gsplits > 1, then splits == gridDim.y == gsplits and there is no cluster): every CTA leaves its
    // 32 x VEC partial in `partials`, the last arriver of a tile (ticket) folds the gsplits partials in split order
    int gsplits;
    Slot<typename P::Acc> *partials;
    uint *tickets;
};

constexpr int kDimRows = 8;  // warps per CTA; each warp owns every kDimRows-th row of the CTA's slice

template<typename P, int VEC>
__global__ void __launch_bounds__(32 * kDimRows) dim_kernel(const DimArgs<P> a) {
    typedef typename P::Acc Acc;
    typedef typename P::In In;
    __shared__ Acc sm[kDimRows][32 * VEC + 1];
    __shared__ Acc sm_cluster[8][32 * VEC];
    __shared__ int sm_last;
    const int lane = threadIdx.x, wy = threadIdx.y;
    const dim_t x0 = ((dim_t)blockIdx.x * 32 + lane) * VEC;  // first dim-0 element of this thread
    const int split = blockIdx.y;
    const dim_t rows_per_split = (a.g.len + a.splits - 1) / a.splits;
    const dim_t r_lo = split * rows_per_split, r_hi = min(r_lo + rows_per_split, a.g.len);

    for (dim_t batch = blockIdx.z; batch < a.g.na * a.g.nb; batch += gridDim.z) {
        const dim_t ia = batch % a.g.na, ib = batch / a.g.na;
        const In *base = a.in + ia * a.g.sa + ib * a.g.sb + x0;
        Acc acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; k++) acc[k] = P::init();
        if (x0 < a.g.d0) {
            dim_t r = r_lo + wy;
            for (; r + 3 * kDimRows < r_hi; r += 4 * kDimRows) {
                Pack<In, VEC> v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = load_pack<In, VEC>(base + (r + u * kDimRows) * a.g.sd);
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int k = 0; k < VEC; k++)
                        acc[k] = P::join(acc[k], P::load(a.prm, v[u].e[k], (uint)(r + u * kDimRows)));
            }
            for (; r < r_hi; r += kDimRows) {
                const Pack<In, VEC> v = load_pack<In, VEC>(base + r * a.g.sd);
#pragma unroll
                for (int k = 0; k < VEC; k++) acc[k] = P::join(acc[k], P::load(a.prm, v.e[k], (uint)r));
            }
        }
        // combine the kDimRows warps of the CTA
        __syncthreads();
#pragma unroll
        for (int k = 0; k < VEC; k++) sm[wy][lane * VEC + k] = acc[k];
        __syncthreads();
        if (wy == 0) {
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                Acc v = sm[0][lane * VEC + k];
                for (int y = 1; y < kDimRows; y++) v = P::join(v, sm[y][lane * VEC + k]);
                acc[k] = v;
            }
        }
        // narrow dim 0 (few x blocks): many splits of the reduced axis, combined through global partials
        if (a.gsplits > 1) {
            const size_t tile = (size_t)batch * gridDim.x + blockIdx.x;
            Slot<Acc> *part = a.partials + tile * (size_t)a.gsplits * (32 * VEC);
            if (wy == 0) {
#pragma unroll
                for (int k = 0; k < VEC; k++) part[(size_t)split * (32 * VEC) + lane * VEC + k].v = acc[k];
                __threadfence();
            }
            __syncthreads();
            if (wy == 0 && lane == 0) sm_last = (atomicAdd(a.tickets + tile, 1u) == (uint)a.gsplits - 1u);
            __syncthreads();
            if (!sm_last) continue;            // CTA-uniform
            __threadfence();
#pragma unroll
            for (int k = 0; k < VEC; k++) acc[k] = P::init();
            for (int s = wy; s < a.gsplits; s += kDimRows)
#pragma unroll
                for (int k = 0; k < VEC; k++) acc[k] = P::join(acc[k], load_cg(part + (size_t)s * (32 * VEC) + lane * VEC + k));
            __syncthreads();
#pragma unroll
            for (int k = 0; k < VEC; k++) sm[wy][lane * VEC + k] = acc[k];
            __syncthreads();
            if (wy == 0) {
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    Acc v = sm[0][lane * VEC + k];
                    for (int y = 1; y < kDimRows; y++) v = P::join(v, sm[y][lane * VEC + k]);
                    acc[k] = v;
                }
                if (lane == 0) a.tickets[tile] = 0u;   // self-reset: the next launch finds it clean
            }
        } else
        // combine the CTAs of the cluster (splits of the reduced axis)
        if (a.splits > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            const unsigned rank = cluster.block_rank();
            if (wy == 0) {
                Acc(*leader)[32 * VEC] = cluster.map_shared_rank(sm_cluster, 0);
#pragma unroll
                for (int k = 0; k < VEC; k++) leader[rank][lane * VEC + k] = acc[k];
            }
            cluster.sync();
            if (rank == 0 && wy == 0) {
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    Acc v = sm_cluster[0][lane * VEC + k];
                    for (int s = 1; s < a.splits; s++) v = P::join(v, sm_cluster[s][lane * VEC + k]);
                    acc[k] = v;
                }
            }
            cluster.sync();  // leader's smem must stay valid until read; also orders the next batch iteration
            if (rank != 0) continue;
        }
        if (wy == 0 && x0 < a.g.d0) {
            const size_t o = (size_t)(ia * a.g.oa + ib * a.g.ob + x0);
#pragma unroll
            for (int k = 0; k < VEC; k++) P::store(a.out, o + k, acc[k]);
        }
    }
}

}  // namespace afb
