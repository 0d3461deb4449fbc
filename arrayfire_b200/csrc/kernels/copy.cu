// copy.cu -- the small data-movement kernels around the hot path: dense copy of a strided view
// (af_copy_array, non-linear inputs of the flattening ops), constant fill (af_constant) and the
// widening casts af_scan_by_key applies to narrow integer values
// (/root/reference/src/api/c/scan.cpp:258-274 castArray).
#include "scan_host.cuh"

namespace afb {

template<typename W>
__global__ void copy_strided_kernel(const W *in, W *out, dim_t d0, dim_t d1, dim_t d2, dim_t n, dim_t s0, dim_t s1,
                                    dim_t s2, dim_t s3) {
    for (dim_t i = (dim_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (dim_t)gridDim.x * blockDim.x) {
        const dim_t x = i % d0, r = i / d0;
        const dim_t y = r % d1, r2 = r / d1;
        const dim_t z = r2 % d2, w = r2 / d2;
        out[i] = in[x * s0 + y * s1 + z * s2 + w * s3];
    }
}

void launch_copy_strided(const LaunchCtx &c, af_dtype t, const View &in, void *out) {
    const dim_t n = in.dims[0] * in.dims[1] * in.dims[2] * in.dims[3];
    if (n == 0) return;
    dim_t blocks = (n + 255) / 256;
    const dim_t cap = (dim_t)c.sms * 16;
    if (blocks > cap) blocks = cap;
    const dim_t d0 = in.dims[0], d1 = in.dims[1], d2 = in.dims[2];
    const dim_t s0 = in.strides[0], s1 = in.strides[1], s2 = in.strides[2], s3 = in.strides[3];
    switch (dtype_size(t)) {
        case 1: copy_strided_kernel<uchar><<<(unsigned)blocks, 256, 0, c.stream>>>((const uchar *)in.ptr, (uchar *)out, d0, d1, d2, n, s0, s1, s2, s3); break;
        case 2: copy_strided_kernel<ushort><<<(unsigned)blocks, 256, 0, c.stream>>>((const ushort *)in.ptr, (ushort *)out, d0, d1, d2, n, s0, s1, s2, s3); break;
        case 4: copy_strided_kernel<uint><<<(unsigned)blocks, 256, 0, c.stream>>>((const uint *)in.ptr, (uint *)out, d0, d1, d2, n, s0, s1, s2, s3); break;
        case 8: copy_strided_kernel<uintl><<<(unsigned)blocks, 256, 0, c.stream>>>((const uintl *)in.ptr, (uintl *)out, d0, d1, d2, n, s0, s1, s2, s3); break;
        default: copy_strided_kernel<double2><<<(unsigned)blocks, 256, 0, c.stream>>>((const double2 *)in.ptr, (double2 *)out, d0, d1, d2, n, s0, s1, s2, s3); break;
    }
    note_launches(1);
    post_launch_check("copy_strided_kernel");
}

// af_tile (/root/reference/src/api/c/tile.cpp:41-84, backend/common/tile.hpp): out(x, y, z, w) = in(x % d0, y % d1, z % d2, w % d3);
// one thread per output element, consecutive threads along dim 0 (coalesced stores, loads coalesced within each repeat)
template<typename W>
__global__ void tile_kernel(const W *in, W *out, dim_t o0, dim_t o1, dim_t o2, dim_t n, dim_t d0, dim_t d1, dim_t d2, dim_t d3,
                            dim_t s0, dim_t s1, dim_t s2, dim_t s3) {
    for (dim_t i = (dim_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (dim_t)gridDim.x * blockDim.x) {
        const dim_t x = i % o0, r = i / o0;
        const dim_t y = r % o1, r2 = r / o1;
        const dim_t z = r2 % o2, w = r2 / o2;
        out[i] = in[(x % d0) * s0 + (y % d1) * s1 + (z % d2) * s2 + (w % d3) * s3];
    }
}

void launch_tile(const LaunchCtx &c, af_dtype t, const View &in, const dim_t odims[4], void *out) {
    const dim_t n = odims[0] * odims[1] * odims[2] * odims[3];
    if (n == 0) return;
    dim_t blocks = (n + 255) / 256;
    const dim_t cap = (dim_t)c.sms * 16;
    if (blocks > cap) blocks = cap;
#define AFB_TILE(W)                                                                                                        \
    tile_kernel<W><<<(unsigned)blocks, 256, 0, c.stream>>>((const W *)in.ptr, (W *)out, odims[0], odims[1], odims[2], n,    \
                                                           in.dims[0], in.dims[1], in.dims[2], in.dims[3], in.strides[0],  \
                                                           in.strides[1], in.strides[2], in.strides[3])
    switch (dtype_size(t)) {
        case 1: AFB_TILE(uchar); break;
        case 2: AFB_TILE(ushort); break;
        case 4: AFB_TILE(uint); break;
        case 8: AFB_TILE(uintl); break;
        default: AFB_TILE(double2); break;
    }
#undef AFB_TILE
    note_launches(1);
    post_launch_check("tile_kernel");
}

template<typename T> __global__ void fill_kernel(T *out, dim_t n, T v) {
    for (dim_t i = (dim_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (dim_t)gridDim.x * blockDim.x) out[i] = v;
}
template<typename T> static void fill(const LaunchCtx &c, void *out, dim_t n, T v) {
    dim_t blocks = (n + 255) / 256;
    const dim_t cap = (dim_t)c.sms * 16;
    if (blocks > cap) blocks = cap;
    fill_kernel<T><<<(unsigned)blocks, 256, 0, c.stream>>>((T *)out, n, v);
    note_launches(1);
    post_launch_check("fill_kernel");
}

void launch_fill(const LaunchCtx &c, af_dtype t, void *out, dim_t n, double re, double im) {
    if (n == 0) return;
    switch (t) {
        case f32: fill<float>(c, out, n, (float)re); break;
        case f64: fill<double>(c, out, n, re); break;
        case c32: fill<float2>(c, out, n, make_float2((float)re, (float)im)); break;
        case c64: fill<double2>(c, out, n, make_double2(re, im)); break;
        case s32: fill<int>(c, out, n, (int)re); break;
        case u32: fill<uint>(c, out, n, (uint)re); break;
        case s64: fill<intl>(c, out, n, (intl)re); break;
        case u64: fill<uintl>(c, out, n, (uintl)re); break;
        case s16: fill<short>(c, out, n, (short)re); break;
        case u16: fill<ushort>(c, out, n, (ushort)re); break;
        case s8: fill<schar>(c, out, n, (schar)re); break;
        case u8: fill<uchar>(c, out, n, (uchar)re); break;
        case b8: fill<schar>(c, out, n, (schar)(re != 0)); break;
        case f16: fill<__half>(c, out, n, __float2half_rn((float)re)); break;
    }
}

template<typename Ti, typename To>
__global__ void cast_kernel(const Ti *in, To *out, dim_t d0, dim_t d1, dim_t d2, dim_t n, dim_t s1, dim_t s2, dim_t s3) {
    for (dim_t i = (dim_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (dim_t)gridDim.x * blockDim.x) {
        const dim_t x = i % d0, r = i / d0;
        const dim_t y = r % d1, r2 = r / d1;
        const dim_t z = r2 % d2, w = r2 / d2;
        out[i] = (To)in[x + y * s1 + z * s2 + w * s3];
    }
}
template<typename Ti, typename To> static void cast(const LaunchCtx &c, const View &in, void *out) {
    const dim_t n = in.dims[0] * in.dims[1] * in.dims[2] * in.dims[3];
    if (n == 0) return;
    dim_t blocks = (n + 255) / 256;
    const dim_t cap = (dim_t)c.sms * 16;
    if (blocks > cap) blocks = cap;
    cast_kernel<Ti, To><<<(unsigned)blocks, 256, 0, c.stream>>>((const Ti *)in.ptr, (To *)out, in.dims[0], in.dims[1],
                                                                 in.dims[2], n, in.strides[1], in.strides[2], in.strides[3]);
    note_launches(1);
    post_launch_check("cast_kernel");
}

bool launch_cast(const LaunchCtx &c, af_dtype ti, af_dtype to, const View &in, void *out) {
    if (to == s32) {
        if (ti == s16) { cast<short, int>(c, in, out); return true; }
        if (ti == s8) { cast<schar, int>(c, in, out); return true; }
    } else if (to == u32) {
        if (ti == u16) { cast<ushort, uint>(c, in, out); return true; }
        if (ti == u8) { cast<uchar, uint>(c, in, out); return true; }
        if (ti == b8) { cast<schar, uint>(c, in, out); return true; }
    }
    return false;
}

}  // namespace afb
