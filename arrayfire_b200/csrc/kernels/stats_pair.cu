// stats_pair.cu -- reductions over TWO arrays read together: af_corrcoef (five sums in one pass over x and y;
// /root/reference/src/api/c/corrcoef.cpp:39-60 makes three full-size temporaries and five separate reductions) and
// af_cov (per column: sum of (x - mean(x)) (y - mean(y)); covariance.cpp:41-63 -- the means are those of the WHOLE arrays,
// as in the reference).  Grid = (splits, columns); every CTA leaves K partial sums, the last arriver of a column folds them
// in split order (deterministic) -- the same ticket scheme as span_kernel.
#include "reduce_host.cuh"

namespace afb {

constexpr int kPairThreads = 256;

template<typename Tc, int K> struct PairAcc { Tc s[K]; };

template<typename Ti, typename Tc, int MODE> struct PairRule;
// MODE 0: corrcoef sums {x, y, x^2, y^2, x y}
template<typename Ti, typename Tc> struct PairRule<Ti, Tc, 0> {
    static constexpr int K = 5;
    AFB_DI static void add(Tc *s, Ti xi, Ti yi, Tc, Tc) {
        const Tc x = conv<Tc, Ti>(xi), y = conv<Tc, Ti>(yi);
        s[0] += x; s[1] += y; s[2] += x * x; s[3] += y * y; s[4] += x * y;
    }
};
// MODE 1: covariance numerator
template<typename Ti, typename Tc> struct PairRule<Ti, Tc, 1> {
    static constexpr int K = 1;
    AFB_DI static void add(Tc *s, Ti xi, Ti yi, Tc mx, Tc my) { s[0] += (conv<Tc, Ti>(xi) - mx) * (conv<Tc, Ti>(yi) - my); }
};

template<typename Ti, typename Tc>
struct PairArgs {
    const Ti *x, *y;
    dim_t n, ncols;            // dense columns of n elements
    dim_t chunk;               // elements per split
    int splits;
    const Tc *mx, *my;         // MODE 1: device scalars (whole-array means)
    Tc scale;                  // applied to the K results
    Tc *out;                   // [ncols][K]
    Tc *partials;              // [ncols][splits][K]
    uint *tickets;             // [ncols]
};

template<typename Ti, typename Tc, int MODE>
__global__ void __launch_bounds__(kPairThreads) pair_kernel(const PairArgs<Ti, Tc> a) {
    typedef PairRule<Ti, Tc, MODE> Rl;
    constexpr int K = Rl::K;
    __shared__ Tc sm[32][K];
    __shared__ int sm_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const dim_t col = blockIdx.y;
    const int split = blockIdx.x;
    const dim_t lo = (dim_t)split * a.chunk, hi = min(lo + a.chunk, a.n);
    const Ti *x = a.x + col * a.n, *y = a.y + col * a.n;
    const Tc mx = a.mx ? *a.mx : (Tc)0, my = a.my ? *a.my : (Tc)0;
    Tc s[K];
#pragma unroll
    for (int k = 0; k < K; k++) s[k] = (Tc)0;
    dim_t i = lo + tid;
    for (; i + 3 * kPairThreads < hi; i += 4 * kPairThreads) {   // four independent loads of each array in flight
        Ti xv[4], yv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            xv[u] = ldg_stream<Ti>(x + i + u * kPairThreads);
            yv[u] = ldg_stream<Ti>(y + i + u * kPairThreads);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) Rl::add(s, xv[u], yv[u], mx, my);
    }
    for (; i < hi; i += kPairThreads) Rl::add(s, ldg_stream<Ti>(x + i), ldg_stream<Ti>(y + i), mx, my);
    // CTA fold
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s[k] += __shfl_down_sync(0xffffffffu, s[k], d);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < K; k++) sm[warp][k] = s[k];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            Tc v = lane < kPairThreads / 32 ? sm[lane][k] : (Tc)0;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
            s[k] = v;
        }
    }
    Tc *part = a.partials + (size_t)col * a.splits * K;
    if (tid == 0) {
        if (a.splits == 1) {
#pragma unroll
            for (int k = 0; k < K; k++) a.out[col * K + k] = s[k] * a.scale;
            sm_last = 0;
        } else {
#pragma unroll
            for (int k = 0; k < K; k++) part[(size_t)split * K + k] = s[k];
            __threadfence();
            sm_last = (atomicAdd(a.tickets + col, 1u) == (uint)a.splits - 1u);
        }
    }
    __syncthreads();
    if (!sm_last) return;
    __threadfence();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            Tc v = (Tc)0;
            for (int sp = lane; sp < a.splits; sp += 32) v += __ldcg(part + (size_t)sp * K + k);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
            if (lane == 0) a.out[col * K + k] = v * a.scale;
        }
        if (lane == 0) a.tickets[col] = 0u;
    }
}

template<typename Ti, typename Tc, int MODE>
static void run_pair(const LaunchCtx &c, const void *x, const void *y, dim_t n, dim_t ncols, const void *mx, const void *my,
                     double scale, void *out) {
    constexpr int K = PairRule<Ti, Tc, MODE>::K;
    PairArgs<Ti, Tc> a;
    a.x = (const Ti *)x;
    a.y = (const Ti *)y;
    a.n = n;
    a.ncols = ncols;
    dim_t splits = ((dim_t)c.sms * 6 + ncols - 1) / ncols;
    const dim_t max_splits = (n + 4 * kPairThreads - 1) / (4 * kPairThreads);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 1024) splits = 1024;
    a.splits = (int)splits;
    a.chunk = (n + splits - 1) / splits;
    a.mx = (const Tc *)mx;
    a.my = (const Tc *)my;
    a.scale = (Tc)scale;
    a.out = (Tc *)out;
    a.partials = (Tc *)c.scratch->need_partials((size_t)ncols * splits * K * sizeof(Tc));
    a.tickets = c.scratch->need_tickets((size_t)ncols);
    if (ncols > 65535) fail(AF_ERR_SIZE, "cov: too many columns (%lld)", (long long)ncols);
    dim3 grid((unsigned)splits, (unsigned)ncols, 1);
    pair_kernel<Ti, Tc, MODE><<<grid, kPairThreads, 0, c.stream>>>(a);
    note_launches(1);
    post_launch_check("pair_kernel");
}

// (Ti -> Tc) pairs of corrcoef.cpp:84-96 / covariance.cpp:92-104 (cov has no b8 case)
#define AFB_PAIR_TYPES(X)                                                                                     \
    X(f32, float, float) X(f64, double, double) X(s32, int, float) X(u32, uint, float) X(s64, intl, double)   \
    X(u64, uintl, double) X(s16, short, float) X(u16, ushort, float) X(s8, schar, float) X(u8, uchar, float)  \
    X(b8, b8_t, float)

// out: 5 sums {x, y, x^2, y^2, xy} of the compute type (float or double)
bool launch_corr_sums(const LaunchCtx &c, af_dtype t, const void *x, const void *y, dim_t n, void *out) {
#define X(code, Ti, Tc)                                                         \
    if (t == code) {                                                            \
        run_pair<Ti, Tc, 0>(c, x, y, n, 1, nullptr, nullptr, 1.0, out);         \
        return true;                                                            \
    }
    AFB_PAIR_TYPES(X)
#undef X
    return false;
}

// out[j] = scale * sum_i (x[i,j] - *mx) (y[i,j] - *my); mx / my: device scalars of the compute type
bool launch_cov_cols(const LaunchCtx &c, af_dtype t, const void *x, const void *y, dim_t n, dim_t ncols, const void *mx,
                     const void *my, double scale, void *out) {
#define X(code, Ti, Tc)                                                         \
    if (t == code) {                                                            \
        run_pair<Ti, Tc, 1>(c, x, y, n, ncols, mx, my, scale, out);             \
        return true;                                                            \
    }
    AFB_PAIR_TYPES(X)
#undef X
    return false;
}

}  // namespace afb
