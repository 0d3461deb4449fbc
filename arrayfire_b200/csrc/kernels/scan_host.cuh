// scan_host.cuh -- launch-shape selection for the scan kernels (host side).
#pragma once

#include "reduce_host.cuh"
#include "chain_kernels.cuh"
#include "pipe_kernels.cuh"
#include "scan_kernels.cuh"

namespace afb {

inline bool aligned_to(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template<int OP, typename Ti, typename To, bool KEYED, typename Tk>
void run_scan(const LaunchCtx &c, const View &in, const View *key, int dim, bool inclusive, void *out,
              const void *carry, const void *carry_key, void *summary) {
    typedef Rule<OP, Ti, To> R;
    typedef SOps<R, To, KEYED> S;
    typedef typename S::State State;
    const dim_t n = in.dims[0] * in.dims[1] * in.dims[2] * in.dims[3];
    if (n == 0) return;
    if (dim == 0) {
        ColGeom g = col_geom(in);
        constexpr int E = ScanVec<Ti, To>::value;
        constexpr dim_t TILE = (dim_t)kScanThreads * kScanSlots * E;
        const dim_t short_len = env_int("AFB200_SHORT_SCAN", 1024);
        if (!carry && !summary && g.len <= short_len && g.ncols >= 32) {
            WarpScanArgs<Ti, To, Tk> a;
            a.in = (const Ti *)in.ptr;
            a.key = KEYED ? (const Tk *)key->ptr : nullptr;
            a.out = (To *)out;
            a.g = g;
            a.ks1 = KEYED ? key->strides[1] : 0;
            a.ks2 = KEYED ? key->strides[2] : 0;
            a.ks3 = KEYED ? key->strides[3] : 0;
            a.inclusive = inclusive ? 1 : 0;
            dim_t blocks = (g.ncols + kScanWarps - 1) / kScanWarps;
            const dim_t cap = (dim_t)c.sms * 32;
            if (blocks > cap) blocks = cap;
            warp_scan_cols_kernel<OP, Ti, To, KEYED, Tk><<<(unsigned)blocks, kScanThreads, 0, c.stream>>>(a);
            note_launches(1);
            post_launch_check("warp_scan_cols_kernel");
            return;
        }
        // ---- persistent TMA-fed chained scan: needs 16-byte aligned columns of in / key / out
        {
            constexpr int UC = KEYED ? 4 : 8;
            constexpr dim_t CT = (dim_t)kChainConsumers * UC * E;
            bool ok = aligned_to(in.ptr, 16) && aligned_to(out, 16) && (!KEYED || aligned_to(key->ptr, 16));
            if (g.ncols > 1) {
                ok = ok && ((g.len * sizeof(To)) % 16 == 0);
                for (int i = 1; i < 4; i++) {
                    if (in.dims[i] > 1) {
                        ok = ok && ((in.strides[i] * sizeof(Ti)) % 16 == 0);
                        if (KEYED) ok = ok && ((key->strides[i] * sizeof(Tk)) % 16 == 0);
                    }
                }
            }
            // ---- 4-byte values (and keys): software-pipelined TMA-in / TMA-out kernel, one CTA per SM
            if constexpr (((sizeof(Ti) == sizeof(To) && (sizeof(To) == 4 || sizeof(To) == 8)) || (!KEYED && sizeof(Ti) < 4 && sizeof(To) == 4)) &&
                          (sizeof(Tk) == 4 || sizeof(Tk) == 8)) {
                // measured on B200 (profiles/r01*): plain scans like few, fat warps (12 x 20 elements, 7 stages: 4.8 TB/s vs
                // 4.1 TB/s with 16 x 12); by-key scans are latency-bound in pass 1 and like more warps (16 x 12 elements,
                // 6 value + 3 key stages: 4.7 TB/s vs 4.1 TB/s with 12 warps)
                // 8-byte values (f64, s64, u64, c32) keep the same stage bytes with half the elements per thread
                constexpr int ES = (int)sizeof(To);
                constexpr bool NARROW = sizeof(Ti) < sizeof(To);   // promoting scans: smaller tiles, the input has its own ring
                constexpr int EPT = NARROW ? 12 : (KEYED ? 12 : 20) * 4 / ES;
#ifndef AFB_SBK_NW
#define AFB_SBK_NW 16
#define AFB_SBK_ST 6
#define AFB_SBK_DL 3
#endif
                constexpr int NWP = KEYED ? ((sizeof(Tk) > sizeof(To)) ? 12 : AFB_SBK_NW) : 12;   // 8-byte keys of 4-byte values: smaller tiles
                constexpr int ST = KEYED ? ((sizeof(Tk) > sizeof(To)) ? 6 : AFB_SBK_ST) : 7;
                constexpr int DL = (KEYED && sizeof(Tk) <= sizeof(To)) ? AFB_SBK_DL : 3;   // measured: D = 2 (one more tile of prefetch instead) costs the plain scan 25 % (0.140 vs 0.110 ms)
                typedef PipeLayout<KEYED || NARROW, EPT, ST, NWP, DL, (int)sizeof(To), NARROW ? (int)sizeof(Ti) : (int)sizeof(Tk)> PL;
                static_assert(PL::total <= 225 * 1024, "pipe scan: the stage rings must fit one SM");
                if (ok && g.len >= 4 * (dim_t)PL::TILE && !env_int("AFB200_NO_PIPE", 0)) {
                    auto kern = pipe_scan_kernel<OP, Ti, To, KEYED, Tk, EPT, ST, NWP, DL>;
                    static bool configured[64] = {false};
                    const int dv = c.device & 63;
                    if (!configured[dv]) {
                        AFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PL::total));
                        configured[dv] = true;
                    }
                    PipeScanArgs<Ti, To, Tk> a;
                    a.in = (const Ti *)in.ptr;
                    a.key = KEYED ? (const Tk *)key->ptr : nullptr;
                    a.out = (To *)out;
                    a.g = g;
                    a.ks1 = KEYED ? key->strides[1] : 0;
                    a.ks2 = KEYED ? key->strides[2] : 0;
                    a.ks3 = KEYED ? key->strides[3] : 0;
                    a.tiles_per_col = (g.len + PL::TILE - 1) / PL::TILE;
                    a.ntiles = a.tiles_per_col * g.ncols;
                    a.inclusive = inclusive ? 1 : 0;
                    a.summary = (g.ncols == 1) ? summary : nullptr;
                    if (a.summary && carry) fail(AF_ERR_INTERNAL, "scan: shard summary and carry-in are mutually exclusive");
                    if (a.summary) {
                        AFB_CUDA(cudaMemsetAsync(summary, 0, 24, c.stream));
                        AFB_CUDA(cudaMemsetAsync((char *)summary + 24, 0xff, 8, c.stream));
                    }
                    a.descs = (unsigned long long *)c.scratch->need_wtiles((size_t)a.ntiles * 8 * (sizeof(To) / 4), c.stream);
                    a.gen = c.scratch->next_wepoch(c.stream);
                    a.carry = (const To *)carry;
                    a.carry_key = (const Tk *)carry_key;
                    a.ahead = env_int("AFB200_PIPE_AHEAD", ST - DL - 1);
                    if (a.ahead > ST - DL - 1) a.ahead = ST - DL - 1;
                    if (a.ahead < 1) a.ahead = 1;
                    a.window = env_int("AFB200_PIPE_WINDOW", kPipeWindow);
                    if (a.window < 1 || a.window > kPipeWindow) a.window = kPipeWindow;
                    dim_t grid = (dim_t)c.sms;
                    if (grid > a.ntiles) grid = a.ntiles;
                    a.ticket = c.scratch->chain_ticket;
                    a.ticket_base = c.scratch->chain_base;
                    c.scratch->chain_base += (unsigned long long)a.ntiles + 2ULL * (unsigned long long)grid;  // two tickets in flight per CTA
                    static unsigned long long *dbg_buf = nullptr;
                    const bool dbg = env_int("AFB200_PIPE_DEBUG", 0) != 0;
                    if (dbg && !dbg_buf) AFB_CUDA(cudaMalloc(&dbg_buf, 1024 * 16 * 8));
                    a.debug = dbg ? dbg_buf : nullptr;
                    if (dbg) AFB_CUDA(cudaMemsetAsync(dbg_buf, 0, 1024 * 16 * 8, c.stream));
                    kern<<<(unsigned)grid, pipe_threads(NWP), PL::total, c.stream>>>(a);
                    note_launches(1);
                    post_launch_check("pipe_scan_kernel");
                    if (dbg) {
                        static unsigned long long h[1024 * 16];
                        AFB_CUDA(cudaStreamSynchronize(c.stream));
                        AFB_CUDA(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
                        const char *names[12] = {"lb_total", "lb_resolve", "polls", "tiles", "dma_wait_p2", "dma_wait_read",
                                                 "dma_wait_p1key", "dma_total", "cmp_wait_full", "cmp_wait_lb", "cmp_p1", "cmp_p2"};
                        fprintf(stderr, "[pipe_scan debug] grid=%d ntiles=%lld keyed=%d\n", (int)grid, (long long)a.ntiles, (int)KEYED);
                        for (int q = 0; q < 12; q++) {
                            double sum = 0, mx = 0;
                            for (int b = 0; b < (int)grid; b++) { double v = (double)h[b * 16 + q]; sum += v; if (v > mx) mx = v; }
                            fprintf(stderr, "   %-14s mean %.0f  max %.0f\n", names[q], sum / (double)grid, mx);
                        }
                    }
                    return;
                }
            }
            if (ok && g.len >= CT && !env_int("AFB200_NO_CHAIN", 0)) {
                typedef ChainSmem<Ti, Tk, KEYED, (int)CT> Lay;
                auto kern = chain_scan_kernel<OP, Ti, To, KEYED, Tk, UC>;
                static int resident[64] = {0};  // CTAs per SM, per device (0: not configured yet)
                const int dv = c.device & 63;
                if (!resident[dv]) {
                    AFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay::total));
                    int nb = 0;
                    AFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kChainThreads, Lay::total));
                    resident[dv] = nb > 0 ? nb : 1;
                }
                ChainScanArgs<Ti, To, Tk> a;
                a.in = (const Ti *)in.ptr;
                a.key = KEYED ? (const Tk *)key->ptr : nullptr;
                a.out = (To *)out;
                a.g = g;
                a.ks1 = KEYED ? key->strides[1] : 0;
                a.ks2 = KEYED ? key->strides[2] : 0;
                a.ks3 = KEYED ? key->strides[3] : 0;
                a.tiles_per_col = (g.len + CT - 1) / CT;
                a.ntiles = a.tiles_per_col * g.ncols;
                a.inclusive = inclusive ? 1 : 0;
                a.summary = (g.ncols == 1) ? summary : nullptr;
                if (a.summary && carry) fail(AF_ERR_INTERNAL, "scan: shard summary and carry-in are mutually exclusive");
                if (a.summary) {
                    AFB_CUDA(cudaMemsetAsync(summary, 0, 24, c.stream));
                    AFB_CUDA(cudaMemsetAsync((char *)summary + 24, 0xff, 8, c.stream));
                }
                if (sizeof(To) == 4) {
                    a.descs = c.scratch->need_wtiles((size_t)a.ntiles * 8, c.stream);
                    a.gen = c.scratch->next_wepoch(c.stream);
                } else {
                    a.descs = c.scratch->need_tiles((size_t)a.ntiles * sizeof(TileDesc<State>), c.stream);
                    a.gen = c.scratch->next_epoch(c.stream);
                }
                a.carry = (const To *)carry;
                a.carry_key = (const Tk *)carry_key;
                int ctas = env_int("AFB200_CHAIN_CTAS_PER_SM", resident[dv]);
                if (ctas > resident[dv]) ctas = resident[dv];
                dim_t grid = (dim_t)c.sms * ctas;
                if (grid > a.ntiles) grid = a.ntiles;
                a.ticket = c.scratch->chain_ticket;
                a.ticket_base = c.scratch->chain_base;
                c.scratch->chain_base += (unsigned long long)a.ntiles + (unsigned long long)grid;  // every CTA draws one past the end
                kern<<<(unsigned)grid, kChainThreads, Lay::total, c.stream>>>(a);
                note_launches(1);
                post_launch_check("chain_scan_kernel");
                return;
            }
        }
        LookbackArgs<Ti, To, Tk> a;
        a.in = (const Ti *)in.ptr;
        a.key = KEYED ? (const Tk *)key->ptr : nullptr;
        a.out = (To *)out;
        a.g = g;
        a.ks1 = KEYED ? key->strides[1] : 0;
        a.ks2 = KEYED ? key->strides[2] : 0;
        a.ks3 = KEYED ? key->strides[3] : 0;
        a.tiles_per_col = (g.len + TILE - 1) / TILE;
        a.ntiles = a.tiles_per_col * g.ncols;
        if (a.ntiles >= 0xffffffffLL) fail(AF_ERR_SIZE, "scan: too many tiles");
        a.inclusive = inclusive ? 1 : 0;
        const size_t in_al = sizeof(Ti) * E >= 16 ? 16 : sizeof(Ti) * E;
        const size_t out_al = sizeof(To) * E >= 16 ? 16 : sizeof(To) * E;
        const size_t key_al = sizeof(Tk) * E >= 16 ? 16 : sizeof(Tk) * E;
        bool vec_ok = aligned_to(in.ptr, in_al) && aligned_to(out, out_al);
        if (KEYED) vec_ok = vec_ok && aligned_to(key->ptr, key_al);
        if (g.ncols > 1) {
            vec_ok = vec_ok && (g.len % E == 0);
            for (int i = 1; i < 4; i++) {
                if (in.dims[i] > 1) vec_ok = vec_ok && (in.strides[i] % E == 0) && (!KEYED || key->strides[i] % E == 0);
            }
        }
        a.vec_ok = vec_ok ? 1 : 0;
        a.tiles = c.scratch->need_tiles((size_t)a.ntiles * sizeof(TileDesc<State>), c.stream);
        a.epoch = c.scratch->next_epoch(c.stream);
        a.tile_counter = c.scratch->tile_counter;
        a.carry = (const To *)carry;
        a.carry_key = (const Tk *)carry_key;
        a.summary = (g.ncols == 1) ? summary : nullptr;
        if (a.summary && carry) fail(AF_ERR_INTERNAL, "scan: shard summary and carry-in are mutually exclusive");
        if (a.summary) {  // the last tile writes the tail; first_head starts as "none" (~0)
            AFB_CUDA(cudaMemsetAsync(summary, 0, 24, c.stream));
            AFB_CUDA(cudaMemsetAsync((char *)summary + 24, 0xff, 8, c.stream));
        }
        lookback_kernel<OP, Ti, To, KEYED, Tk><<<(unsigned)a.ntiles, kScanThreads, 0, c.stream>>>(a);
        note_launches(1);
        post_launch_check("lookback_kernel");
        return;
    }
    // dim > 0: the tiled single-pass kernel (scanned axis split, look-back carry) for plain scans with 4-byte outputs
    if constexpr (!KEYED && sizeof(To) == 4) {
        if (!env_int("AFB200_NO_DIMSCAN", 0) && in.dims[0] > 0 && in.dims[dim] > 0) {
            DimScanArgs<Ti, To> d;
            d.in = (const Ti *)in.ptr;
            d.out = (To *)out;
            dim_t os[4];
            dense_strides(in.dims, os);
            d.d0 = in.dims[0];
            d.len = in.dims[dim];
            d.sd = in.strides[dim];
            d.osd = os[dim];
            int others[2], m = 0;
            for (int i = 1; i < 4; i++)
                if (i != dim) others[m++] = i;
            d.na = in.dims[others[0]];
            d.sa = in.strides[others[0]];
            d.osa = os[others[0]];
            d.nb = in.dims[others[1]];
            d.sb = in.strides[others[1]];
            d.osb = os[others[1]];
            d.inclusive = inclusive ? 1 : 0;
            bool vec = (d.d0 % 4 == 0) && aligned_to(in.ptr, sizeof(Ti) * 4) && aligned_to(out, 16) && (d.sd % 4 == 0) &&
                       (d.na == 1 || d.sa % 4 == 0) && (d.nb == 1 || d.sb % 4 == 0);
            const int V = vec ? 4 : 1;
            d.nxb = (d.d0 + 32 * V - 1) / (32 * V);
            d.nchains = d.nxb * d.na * d.nb;
            const dim_t splits = (d.len + kDsRows - 1) / kDsRows;
            d.ntiles = d.nchains * splits;
            d.descs = (unsigned long long *)c.scratch->need_wtiles((size_t)d.ntiles * 32 * V * 8, c.stream);
            d.gen = c.scratch->next_wepoch(c.stream);
            static int resident[64][2] = {{0}};
            const int dv = c.device & 63;
            if (!resident[dv][vec ? 1 : 0]) {
                int occ = 0;
                if (vec) AFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dim_scan_kernel<OP, Ti, To, 4>, 32 * kDsY, 0));
                else AFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dim_scan_kernel<OP, Ti, To, 1>, 32 * kDsY, 0));
                resident[dv][vec ? 1 : 0] = occ < 1 ? 1 : occ;
            }
            dim_t grid = (dim_t)c.sms * resident[dv][vec ? 1 : 0];
            if (grid > d.ntiles) grid = d.ntiles;
            d.ticket = c.scratch->chain_ticket;
            d.ticket_base = c.scratch->chain_base;
            c.scratch->chain_base += (unsigned long long)d.ntiles + (unsigned long long)grid;   // every CTA draws one past the end
            if (vec) dim_scan_kernel<OP, Ti, To, 4><<<(unsigned)grid, dim3(32, kDsY), 0, c.stream>>>(d);
            else dim_scan_kernel<OP, Ti, To, 1><<<(unsigned)grid, dim3(32, kDsY), 0, c.stream>>>(d);
            note_launches(1);
            post_launch_check("dim_scan_kernel");
            return;
        }
    }
    StripScanArgs<Ti, To, Tk> a;
    a.in = (const Ti *)in.ptr;
    a.key = KEYED ? (const Tk *)key->ptr : nullptr;
    a.out = (To *)out;
    dim_t os[4];
    dense_strides(in.dims, os);
    a.d0 = in.dims[0];
    a.len = in.dims[dim];
    a.sd = in.strides[dim];
    a.ksd = KEYED ? key->strides[dim] : 0;
    a.osd = os[dim];
    int others[2], m = 0;
    for (int i = 1; i < 4; i++)
        if (i != dim) others[m++] = i;
    a.na = in.dims[others[0]];
    a.sa = in.strides[others[0]];
    a.ksa = KEYED ? key->strides[others[0]] : 0;
    a.osa = os[others[0]];
    a.nb = in.dims[others[1]];
    a.sb = in.strides[others[1]];
    a.ksb = KEYED ? key->strides[others[1]] : 0;
    a.osb = os[others[1]];
    a.inclusive = inclusive ? 1 : 0;
    const dim_t gx = (a.d0 + 31) / 32;
    dim_t gy = a.na * a.nb;
    if (gy > 65535) gy = 65535;
    if (gx > 0x7fffffffLL) fail(AF_ERR_SIZE, "scan: dim 0 too large");
    strip_scan_kernel<OP, Ti, To, KEYED, Tk><<<dim3((unsigned)gx, (unsigned)gy), dim3(32, kStripY), 0, c.stream>>>(a);
    note_launches(1);
    post_launch_check("strip_scan_kernel");
}

// (Ti -> To) pairs the C API can reach (/root/reference/src/api/c/scan.cpp:133-149,175-215)
#define AFB_SCAN_PAIRS(X)                                                                             \
    X(f32, float, f32, float) X(f64, double, f64, double) X(c32, cfloat, c32, cfloat)                 \
    X(c64, cdouble, c64, cdouble) X(s32, int, s32, int) X(u32, uint, u32, uint) X(s64, intl, s64, intl) \
    X(u64, uintl, u64, uintl) X(s16, short, s32, int) X(u16, ushort, u32, uint) X(s8, schar, s32, int) \
    X(u8, uchar, u32, uint) X(b8, b8_t, u32, uint)

#define AFB_SBK_VALS(X)                                                                               \
    X(f32, float) X(f64, double) X(c32, cfloat) X(c64, cdouble) X(s32, int) X(u32, uint) X(s64, intl) X(u64, uintl)

template<int OP>
bool scan_family(const LaunchCtx &c, af_dtype ti, af_dtype to, const View &in, int dim, bool inclusive, void *out,
                 const void *carry) {
#define X(ci, Ti, co, To)                                                                         \
    if (ti == ci && to == co) {                                                                   \
        run_scan<OP, Ti, To, false, int>(c, in, nullptr, dim, inclusive, out, carry, nullptr, nullptr);    \
        return true;                                                                              \
    }
    AFB_SCAN_PAIRS(X)
#undef X
    return false;
}

// keys are only compared for equality: s32/u32 share the 32-bit kernels, s64/u64 the 64-bit ones
template<int OP>
bool scan_by_key_family(const LaunchCtx &c, af_dtype t, af_dtype tk, const View &key, const View &in, int dim,
                        bool inclusive, void *out, const void *carry, const void *carry_key, void *summary) {
    const bool k32 = (tk == s32 || tk == u32), k64 = (tk == s64 || tk == u64);
    if (!k32 && !k64) return false;
#define X(code, T)                                                                                     \
    if (t == code) {                                                                                   \
        if (k32) run_scan<OP, T, T, true, uint>(c, in, &key, dim, inclusive, out, carry, carry_key, summary);   \
        else run_scan<OP, T, T, true, uintl>(c, in, &key, dim, inclusive, out, carry, carry_key, summary);      \
        return true;                                                                                   \
    }
    AFB_SBK_VALS(X)
#undef X
    return false;
}

bool scan_add(const LaunchCtx &, af_dtype, af_dtype, const View &, int, bool, void *, const void *);
bool scan_mul(const LaunchCtx &, af_dtype, af_dtype, const View &, int, bool, void *, const void *);
bool scan_min(const LaunchCtx &, af_dtype, af_dtype, const View &, int, bool, void *, const void *);
bool scan_max(const LaunchCtx &, af_dtype, af_dtype, const View &, int, bool, void *, const void *);
bool sbk_add(const LaunchCtx &, af_dtype, af_dtype, const View &, const View &, int, bool, void *, const void *, const void *, void *);
bool sbk_mul(const LaunchCtx &, af_dtype, af_dtype, const View &, const View &, int, bool, void *, const void *, const void *, void *);
bool sbk_min(const LaunchCtx &, af_dtype, af_dtype, const View &, const View &, int, bool, void *, const void *, const void *, void *);
bool sbk_max(const LaunchCtx &, af_dtype, af_dtype, const View &, const View &, int, bool, void *, const void *, const void *, void *);

}  // namespace afb
