// reduce_host.cuh -- launch-shape selection for the reduction kernels (host side).
//
// Grid sizing follows the B200 rules: persistent / split grids are multiples of the SM count (148),
// clusters of 2 pack the 148 SMs exactly, and every launch is a SINGLE kernel (the reference needs a
// second launch whenever a line is split, /root/reference/src/backend/cuda/kernel/reduce.hpp:151-188,470-504).
#pragma once

#include <cstdlib>

#include "reduce_kernels.cuh"

namespace afb {

inline int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

template<typename K, typename A>
inline void launch_with_cluster(K kernel, dim3 grid, dim3 block, dim3 cluster, cudaStream_t s, const A &args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster.x;
    attr[0].val.clusterDim.y = cluster.y;
    attr[0].val.clusterDim.z = cluster.z;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    AFB_CUDA(cudaLaunchKernelEx(&cfg, kernel, args));
}

inline ColGeom col_geom(const View &in) {
    ColGeom g;
    g.len = in.dims[0];
    g.ncols = in.dims[1] * in.dims[2] * in.dims[3];
    g.d1 = in.dims[1] > 0 ? in.dims[1] : 1;
    g.d2 = in.dims[2] > 0 ? in.dims[2] : 1;
    g.s1 = in.strides[1];
    g.s2 = in.strides[2];
    g.s3 = in.strides[3];
    return g;
}

inline bool view_is_linear(const View &v) {
    dim_t count = 1;
    for (int i = 0; i < 4; i++) {
        if (v.dims[i] > 1 && v.strides[i] != count) return false;
        count *= v.dims[i];
    }
    return true;
}

// ---- reduce along dim 0 (one output per column) ---------------------------------------------------
template<typename P>
void run_cols(const LaunchCtx &c, const View &in, const typename P::Params &prm, const typename P::Out &out) {
    typedef typename P::Acc Acc;
    typedef typename P::In In;
    ColGeom g = col_geom(in);
    if (g.ncols == 0 || g.len == 0) return;
    const dim_t short_len = env_int("AFB200_SHORT_COL", 2048);
    if (g.len < short_len && g.ncols >= 64) {
        WarpColsArgs<P> a;
        a.in = (const In *)in.ptr;
        a.g = g;
        int G = 1;
        while (G < 32 && G < g.len) G <<= 1;
        a.group = G;
        a.prm = prm;
        a.out = out;
        const dim_t groups_per_block = kThreads / G;
        dim_t blocks = (g.ncols + groups_per_block - 1) / groups_per_block;
        const dim_t cap = (dim_t)c.sms * 32;
        if (blocks > cap) blocks = cap;
        warp_cols_kernel<P><<<(unsigned)blocks, kThreads, 0, c.stream>>>(a);
        note_launches(1);
        post_launch_check("warp_cols_kernel");
        return;
    }
    // long columns: one CTA per (column, chunk)
    SpanArgs<P> a;
    a.in = (const In *)in.ptr;
    a.g = g;
    a.all = 0;
    // ONE wave: the grid never exceeds what is resident at once (a 48-register instantiation holds 5 CTAs per SM, not 8:
    // asking for 8 left a thin second wave that cannot saturate HBM -- measured -12 % on the 2^30 f32 sum)
    static int resident[64] = {0};
    const int dv = c.device & 63;
    if (!resident[dv]) {
        int occ = 0;
        AFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, span_kernel<P>, kThreads, 0));
        resident[dv] = occ < 1 ? 1 : occ;
    }
    int ctas_per_sm = env_int("AFB200_REDUCE_CTAS_PER_SM", 8);
    if (ctas_per_sm > resident[dv]) ctas_per_sm = resident[dv];
    const dim_t target = (dim_t)c.sms * ctas_per_sm;
    dim_t bpc = 1;
    if (g.ncols < target) {
        bpc = (target + g.ncols - 1) / g.ncols;
        const dim_t min_chunk = 16384;  // elements: keep every CTA streaming long enough to amortise its tail
        dim_t max_bpc = g.len / min_chunk;
        if (max_bpc < 1) max_bpc = 1;
        if (bpc > max_bpc) bpc = max_bpc;
    }
    int csize = env_int("AFB200_CLUSTER", 2);
    if (csize < 1 || csize > 8) csize = 1;
    if (bpc < 2) csize = 1;
    if (csize > 1) bpc = (bpc + csize - 1) / csize * csize;
    a.bpc = (int)bpc;
    a.csize = csize;
    // chunk: multiple of 32 vectors so chunk boundaries stay 16-byte aligned relative to the column start
    const dim_t quantum = (dim_t)VecN<In>::value * 32;
    dim_t chunk = (g.len + bpc - 1) / bpc;
    chunk = (chunk + quantum - 1) / quantum * quantum;
    a.chunk = chunk;
    a.prm = prm;
    a.out = out;
    a.partials = nullptr;
    a.tickets = nullptr;
    a.host_wire = nullptr;
    a.host_seq = 0;
    a.comm = peer_comm_of(nullptr);
    a.idx_base = 0;
    a.comm_mode = 0;
    if (bpc / csize > 1) {
        a.partials = (Slot<Acc> *)c.scratch->need_partials((size_t)g.ncols * (bpc / csize) * sizeof(Slot<Acc>));
        a.tickets = c.scratch->need_tickets((size_t)g.ncols);
    }
    const dim_t nblocks = g.ncols * bpc;
    if (nblocks > 0x7fffffffLL) fail(AF_ERR_SIZE, "reduce: too many columns (%lld)", (long long)g.ncols);
    if (csize > 1)
        launch_with_cluster(span_kernel<P>, dim3((unsigned)nblocks), dim3(kThreads), dim3(csize, 1, 1), c.stream, a);
    else
        span_kernel<P><<<(unsigned)nblocks, kThreads, 0, c.stream>>>(a);
    note_launches(1);
    post_launch_check("span_kernel");
}

// ---- whole-array reduce: ONE output --------------------------------------------------------------
// `in` must be a linear view (the API layer materialises non-linear views with small dim 0 first) or a
// strided view with long columns.
template<typename P>
void run_all(const LaunchCtx &c, const View &in, const typename P::Params &prm, const typename P::Out &out,
             const HostRet *ret = nullptr, const ShardLink *link = nullptr) {
    typedef typename P::Acc Acc;
    typedef typename P::In In;
    ColGeom g = col_geom(in);
    const dim_t n = g.len * g.ncols;
    if (n == 0) fail(AF_ERR_INTERNAL, "run_all on an empty view");
    if (view_is_linear(in)) {  // flatten: one long column
        g.len = n;
        g.ncols = 1;
        g.d1 = g.d2 = 1;
        g.s1 = g.s2 = g.s3 = 0;
    }
    SpanArgs<P> a;
    a.in = (const In *)in.ptr;
    a.g = g;
    a.all = 1;
    // ONE wave: the grid never exceeds what is resident at once (a 48-register instantiation holds 5 CTAs per SM, not 8:
    // asking for 8 left a thin second wave that cannot saturate HBM -- measured -12 % on the 2^30 f32 sum)
    static int resident[64] = {0};
    const int dv = c.device & 63;
    if (!resident[dv]) {
        int occ = 0;
        AFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, span_kernel<P>, kThreads, 0));
        resident[dv] = occ < 1 ? 1 : occ;
    }
    int ctas_per_sm = env_int("AFB200_REDUCE_CTAS_PER_SM", 8);
    if (ctas_per_sm > resident[dv]) ctas_per_sm = resident[dv];
    const dim_t target = (dim_t)c.sms * ctas_per_sm;
    const dim_t quantum = (dim_t)VecN<In>::value * 32;
    const dim_t min_chunk = 16384;
    // chunks per column
    dim_t bpc = (target + g.ncols - 1) / g.ncols;
    dim_t max_bpc = g.len / min_chunk;
    if (max_bpc < 1) max_bpc = 1;
    if (bpc > max_bpc) bpc = max_bpc;
    dim_t chunk = (g.len + bpc - 1) / bpc;
    chunk = (chunk + quantum - 1) / quantum * quantum;
    bpc = (g.len + chunk - 1) / chunk;
    a.bpc = (int)bpc;
    a.chunk = chunk;
    const dim_t nwork = g.ncols * bpc;
    dim_t grid = nwork < target ? nwork : target;
    int csize = env_int("AFB200_CLUSTER", 2);
    if (csize < 1 || csize > 8) csize = 1;
    if (grid < 2 * csize) csize = 1;
    grid = grid / csize * csize;
    a.csize = csize;
    a.prm = prm;
    a.out = out;
    a.partials = (Slot<Acc> *)c.scratch->need_partials((size_t)(grid / csize) * sizeof(Slot<Acc>));
    a.tickets = c.scratch->need_tickets(1);
    a.host_wire = ret ? ret->wire : nullptr;
    a.host_seq = ret ? ret->seq : 0u;
    a.comm = peer_comm_of(link);
    a.idx_base = link ? link->idx_base : 0u;
    a.comm_mode = link ? link->mode : 0;
    if (csize > 1)
        launch_with_cluster(span_kernel<P>, dim3((unsigned)grid), dim3(kThreads), dim3(csize, 1, 1), c.stream, a);
    else
        span_kernel<P><<<(unsigned)grid, kThreads, 0, c.stream>>>(a);
    note_launches(1);
    post_launch_check("span_kernel(all)");
}

// ---- reduce along dim > 0 ---------------------------------------------------------------------------
// CTAs of dim_kernel<P, VEC> resident per SM (cached per device)
template<typename P, int VEC> int dim_resident(int device) {
    static int resident[64] = {0};
    const int dv = device & 63;
    if (!resident[dv]) {
        int occ = 0;
        AFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dim_kernel<P, VEC>, 32 * kDimRows, 0));
        resident[dv] = occ < 1 ? 1 : occ;
    }
    return resident[dv];
}

template<typename P>
void run_dim(const LaunchCtx &c, const View &in, int dim, const typename P::Params &prm, const typename P::Out &out) {
    typedef typename P::In In;
    DimArgs<P> a;
    a.in = (const In *)in.ptr;
    DimGeom &g = a.g;
    g.d0 = in.dims[0];
    g.len = in.dims[dim];
    g.sd = in.strides[dim];
    int others[2], n = 0;
    for (int i = 1; i < 4; i++)
        if (i != dim) others[n++] = i;
    g.na = in.dims[others[0]];
    g.sa = in.strides[others[0]];
    g.nb = in.dims[others[1]];
    g.sb = in.strides[others[1]];
    // dense output strides with dims[dim] = 1
    dim_t od[4] = {in.dims[0], in.dims[1], in.dims[2], in.dims[3]}, os[4];
    od[dim] = 1;
    dense_strides(od, os);
    g.oa = os[others[0]];
    g.ob = os[others[1]];
    if (g.d0 == 0 || g.len == 0 || g.na * g.nb == 0) return;
    a.prm = prm;
    a.out = out;

    constexpr int V = DimVec<In>::value;
    bool vec_ok = V > 1 && (g.d0 % V == 0) && (reinterpret_cast<uintptr_t>(in.ptr) % (sizeof(In) * V) == 0) &&
                  (g.sd % V == 0) && (g.na == 1 || g.sa % V == 0) && (g.nb == 1 || g.sb % V == 0);
    if (env_int("AFB200_DIM_NOVEC", 0)) vec_ok = false;
    dim_t gz = g.na * g.nb;
    if (gz > 65535) gz = 65535;
    // narrow matrices (few rows along dim 0, e.g. the row blocks of a matrix partitioned over 8 GPUs): 16-byte vectors
    // would leave most SMs without a CTA even with 8 cluster splits -> trade vector width for CTAs
    int vec = vec_ok ? V : 1;
    constexpr int VH = (V >= 4 && sizeof(In) * (V / 2) >= 4) ? V / 2 : 1;
    // (tools/dim_sweep.py on B200, 8192 columns: 8192 rows 4349 GB/s with 16-byte vs 3526 with 8-byte vectors;
    //  4096 rows 2473 vs 2901; 1024 rows 641 vs 1024; scalar lanes never win)
    const dim_t want_ctas = (dim_t)c.sms * 3;
    // two-stage split below (keeps the full vectors): ONE full wave of CTAs -- 1.5 waves measured 30 % slower than 1
    int gsplit_per_sm = env_int("AFB200_DIM_GSPLIT_CTAS", 4);
    if (gsplit_per_sm > 0) {
        const int occ = (vec == V && V > 1) ? dim_resident<P, V>(c.device) : dim_resident<P, 1>(c.device);
        if (gsplit_per_sm > occ) gsplit_per_sm = occ;
    }
    if (gsplit_per_sm <= 0 && vec_ok && VH > 1 && ((g.d0 + 32 * V - 1) / (32 * V)) * gz * 8 < want_ctas) vec = VH;
    { const int f = env_int("AFB200_DIM_VEC", 0); if (f == 1 || (f == VH && vec_ok) || (f == V && vec_ok)) vec = f; }
    const dim_t gx = (g.d0 + 32 * vec - 1) / (32 * vec);
    // split the reduced axis across a cluster when the grid would not fill the machine
    int splits = 1;
    const dim_t want = (dim_t)c.sms * 4;
    while (splits < 8 && gx * gz * splits < want && g.len / (splits * 2) >= 4 * kDimRows) splits *= 2;
    if (env_int("AFB200_CLUSTER", 2) <= 1) splits = 1;
    // ... and when even 8 cluster splits leave SMs idle (1024 rows x 8192 columns: 64 CTAs, 1.75 TB/s): as many splits as it
    // takes to give every SM `per_sm` CTAs, combined through global partials by the last arriver of each tile
    a.gsplits = 1;
    a.partials = nullptr;
    a.tickets = nullptr;
    {
        typedef typename P::Acc Acc;
        const int per_sm = gsplit_per_sm;
        const dim_t ntiles = gx * g.na * g.nb;
        if (per_sm > 0 && gx * gz * 8 < (dim_t)c.sms * 4 && g.na * g.nb <= 65535 && g.len >= 16 * kDimRows) {
            dim_t gs = ((dim_t)c.sms * per_sm + ntiles - 1) / ntiles;
            const dim_t max_gs = g.len / (2 * kDimRows);          // at least two rows per warp
            if (gs > max_gs) gs = max_gs;
            if (gs > 1024) gs = 1024;
            if (gs > 8) {
                splits = (int)gs;
                a.gsplits = (int)gs;
                a.partials = (Slot<Acc> *)c.scratch->need_partials((size_t)ntiles * gs * 32 * vec * sizeof(Slot<Acc>));
                a.tickets = c.scratch->need_tickets((size_t)ntiles);
            }
        }
    }
    a.splits = splits;
    const bool clustered = splits > 1 && a.gsplits == 1;
    dim3 grid((unsigned)gx, (unsigned)splits, (unsigned)gz), block(32, kDimRows, 1);
    if (gx > 0x7fffffffLL) fail(AF_ERR_SIZE, "reduce: dim 0 too large");
    if (vec == V && V > 1) {
        if (clustered) launch_with_cluster(dim_kernel<P, V>, grid, block, dim3(1, splits, 1), c.stream, a);
        else dim_kernel<P, V><<<grid, block, 0, c.stream>>>(a);
    } else if (vec == VH && VH > 1) {
        if (clustered) launch_with_cluster(dim_kernel<P, VH>, grid, block, dim3(1, splits, 1), c.stream, a);
        else dim_kernel<P, VH><<<grid, block, 0, c.stream>>>(a);
    } else {
        if (clustered) launch_with_cluster(dim_kernel<P, 1>, grid, block, dim3(1, splits, 1), c.stream, a);
        else dim_kernel<P, 1><<<grid, block, 0, c.stream>>>(a);
    }
    note_launches(1);
    post_launch_check("dim_kernel");
}

template<typename P>
void run_reduce(const LaunchCtx &c, const View &in, int dim, const typename P::Params &prm, const typename P::Out &out) {
    if (dim == 0) run_cols<P>(c, in, prm, out);
    else run_dim<P>(c, in, dim, prm, out);
}

}  // namespace afb
