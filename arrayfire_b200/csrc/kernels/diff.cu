// diff.cu -- small element-wise kernels of SURVEY.md section 8(f)-3, the callers next to the scan / compaction path:
//   diff1 / diff2 along a dimension       /root/reference/src/backend/cpu/kernel/diff.hpp:18-81
//   "differs from its predecessor" mask   std::unique of a sorted vector, /root/reference/src/backend/cpu/set.cpp:32-51
//   gather by u32 index                   (the compaction's second half: values at the indices af_where produced)
// One thread per output element, lanes along the contiguous dim 0 (coalesced); HBM-bound, no staging needed.
#include "device_ops.cuh"

namespace afb {

template<typename T> AFB_DI T dsub(T a, T b) { return (T)(a - b); }
template<> AFB_DI cfloat dsub<cfloat>(cfloat a, cfloat b) { return cfloat{a.x - b.x, a.y - b.y}; }
template<> AFB_DI cdouble dsub<cdouble>(cdouble a, cdouble b) { return cdouble{a.x - b.x, a.y - b.y}; }
template<> AFB_DI b8_t dsub<b8_t>(b8_t a, b8_t b) { return b8_t{(signed char)(a.v - b.v)}; }
template<typename T> AFB_DI T dadd(T a, T b) { return (T)(a + b); }
template<> AFB_DI cfloat dadd<cfloat>(cfloat a, cfloat b) { return cfloat{a.x + b.x, a.y + b.y}; }
template<> AFB_DI cdouble dadd<cdouble>(cdouble a, cdouble b) { return cdouble{a.x + b.x, a.y + b.y}; }
template<> AFB_DI b8_t dadd<b8_t>(b8_t a, b8_t b) { return b8_t{(signed char)(a.v + b.v)}; }

struct DiffGeom {
    dim_t od[4];       // output dims
    dim_t is[4];       // input strides
    dim_t step;        // input stride of the differenced dim
};

template<typename T, int ORDER>
__global__ void __launch_bounds__(256) diff_kernel(const T *__restrict__ in, T *__restrict__ out, const DiffGeom g, dim_t n) {
    for (dim_t o = (dim_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (dim_t)gridDim.x * blockDim.x) {
        dim_t r = o;
        const dim_t i = r % g.od[0]; r /= g.od[0];
        const dim_t j = r % g.od[1]; r /= g.od[1];
        const dim_t k = r % g.od[2];
        const dim_t l = r / g.od[2];
        const T *p = in + i * g.is[0] + j * g.is[1] + k * g.is[2] + l * g.is[3];
        const T a = ldg_stream<T>(p), b = ldg_stream<T>(p + g.step);
        if (ORDER == 1) {
            out[o] = dsub<T>(b, a);                                   // diff.hpp:40
        } else {
            const T c = ldg_stream<T>(p + 2 * g.step);
            out[o] = dsub<T>(dsub<T>(dadd<T>(c, a), b), b);           // diff.hpp:73-74: in[k] + in[i] - in[j] - in[j]
        }
    }
}

template<typename T>
__global__ void __launch_bounds__(256) differs_kernel(const T *__restrict__ in, unsigned char *__restrict__ mask, dim_t n) {
    for (dim_t i = (dim_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (dim_t)gridDim.x * blockDim.x) {
        bool d = true;
        if (i > 0) {
            const T a = ldg_stream<T>(in + i - 1), b = ldg_stream<T>(in + i);
            d = !(a == b);                                            // std::unique keeps x[i] unless x[i-1] == x[i]
        }
        mask[i] = d ? 1 : 0;
    }
}
template<> __global__ void __launch_bounds__(256) differs_kernel<b8_t>(const b8_t *__restrict__ in, unsigned char *__restrict__ mask, dim_t n) {
    for (dim_t i = (dim_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (dim_t)gridDim.x * blockDim.x)
        mask[i] = (i == 0 || in[i - 1].v != in[i].v) ? 1 : 0;
}

template<typename T>
__global__ void __launch_bounds__(256) gather_kernel(const T *__restrict__ in, const unsigned *__restrict__ idx, T *__restrict__ out, dim_t n) {
    for (dim_t i = (dim_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (dim_t)gridDim.x * blockDim.x) out[i] = in[idx[i]];
}

static unsigned grid_for(const LaunchCtx &c, dim_t n) {
    dim_t g = (n + 255) / 256;
    const dim_t cap = (dim_t)c.sms * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

#define AFB_DIFF_TYPES(X)                                                                             \
    X(f32, float) X(f64, double) X(c32, cfloat) X(c64, cdouble) X(s32, int) X(u32, uint) X(s64, intl) \
    X(u64, uintl) X(s16, short) X(u16, ushort) X(s8, schar) X(u8, uchar) X(b8, b8_t)

// out: dense, dims = in.dims with dims[dim] - order
bool launch_diff(const LaunchCtx &c, af_dtype t, const View &in, int dim, int order, void *out) {
    DiffGeom g;
    dim_t n = 1;
    for (int i = 0; i < 4; i++) {
        g.od[i] = in.dims[i] - (i == dim ? order : 0);
        g.is[i] = in.strides[i];
        n *= g.od[i];
    }
    g.step = in.strides[dim];
    if (n <= 0) return true;
#define X(code, T)                                                                                                  \
    if (t == code) {                                                                                                \
        if (order == 1) diff_kernel<T, 1><<<grid_for(c, n), 256, 0, c.stream>>>((const T *)in.ptr, (T *)out, g, n); \
        else diff_kernel<T, 2><<<grid_for(c, n), 256, 0, c.stream>>>((const T *)in.ptr, (T *)out, g, n);            \
        note_launches(1);                                                                                           \
        post_launch_check("diff_kernel");                                                                           \
        return true;                                                                                                \
    }
    AFB_DIFF_TYPES(X)
#undef X
    return false;
}

// mask[i] = (i == 0 || in[i - 1] != in[i]) over a dense vector
bool launch_differs(const LaunchCtx &c, af_dtype t, const void *in, dim_t n, unsigned char *mask) {
    if (n <= 0) return true;
#define X(code, T)                                                                              \
    if (t == code) {                                                                            \
        differs_kernel<T><<<grid_for(c, n), 256, 0, c.stream>>>((const T *)in, mask, n);        \
        note_launches(1);                                                                       \
        post_launch_check("differs_kernel");                                                    \
        return true;                                                                            \
    }
    X(f32, float) X(f64, double) X(s32, int) X(u32, uint) X(s64, intl) X(u64, uintl) X(s16, short) X(u16, ushort)
    X(s8, schar) X(u8, uchar) X(b8, b8_t)
#undef X
    return false;
}

// out[i] = in[idx[i]]
bool launch_gather(const LaunchCtx &c, af_dtype t, const void *in, const unsigned *idx, dim_t n, void *out) {
    if (n <= 0) return true;
    const size_t sz = dtype_size(t);
#define G(T)                                                                                              \
    {                                                                                                     \
        gather_kernel<T><<<grid_for(c, n), 256, 0, c.stream>>>((const T *)in, idx, (T *)out, n);          \
        note_launches(1);                                                                                 \
        post_launch_check("gather_kernel");                                                               \
        return true;                                                                                      \
    }
    if (sz == 1) G(uchar)
    if (sz == 2) G(ushort)
    if (sz == 4) G(uint)
    if (sz == 8) G(uintl)
#undef G
    return false;
}

}  // namespace afb
