// stats_kernels.cuh -- the statistics built on the reduction path: mean, variance / standard
// deviation, norms.  Each is ONE streaming pass over the input through the reduction kernels of reduce_kernels.cuh
// (span_kernel / warp_cols_kernel / dim_kernel) with its element-wise part fused into the load -- the reference composes
// them from element-wise temporaries + plain reductions:
//   mean      /root/reference/src/backend/cpu/kernel/mean.hpp:18-127 (running mean, one serial loop per output)
//   var/stdev /root/reference/src/api/c/var.cpp:52-155, stdev.cpp:44-96  (mean, tile, sub, mul, reduce, div [, sqrt]:
//             two passes over the input plus three full-size temporaries)
//   norm      /root/reference/src/api/c/norm.cpp:45-117 (abs, [mul | pow], reduce)
// Floating-point results agree with the reference within the tolerance BASELINE.json states (1e-5 relative f32, 1e-12
// f64); they are not bit-identical because the association order differs (as for every parallel sum).
#pragma once

#include "reduce_host.cuh"

namespace afb {

template<typename T> struct base_of { typedef T type; };
template<> struct base_of<cfloat> { typedef float type; };
template<> struct base_of<cdouble> { typedef double type; };

template<typename T> AFB_DI T sub_(T a, T b) { return a - b; }
template<> AFB_DI cfloat sub_<cfloat>(cfloat a, cfloat b) { return cfloat{a.x - b.x, a.y - b.y}; }
template<> AFB_DI cdouble sub_<cdouble>(cdouble a, cdouble b) { return cdouble{a.x - b.x, a.y - b.y}; }
template<typename T, typename B> AFB_DI T scale_(T a, B s) { return a * s; }
template<> AFB_DI cfloat scale_<cfloat, float>(cfloat a, float s) { return cfloat{a.x * s, a.y * s}; }
template<> AFB_DI cdouble scale_<cdouble, double>(cdouble a, double s) { return cdouble{a.x * s, a.y * s}; }
template<typename T> AFB_DI T sqrt_(T a) { return sqrt(a); }
template<> AFB_DI float sqrt_<float>(float a) { return sqrtf(a); }
template<> AFB_DI cfloat sqrt_<cfloat>(cfloat a) { return a; }     // af_stdev has no complex case (stdev.cpp:88-93)
template<> AFB_DI cdouble sqrt_<cdouble>(cdouble a) { return a; }
// value as the compute type: complex stays complex, everything else goes to float / double
template<typename Tc, typename Ti> struct ToCompute { AFB_DI static Tc go(Ti x) { return conv<Tc, Ti>(x); } };

// ------------------------------------------------------------------------------------------------
// mean: (sum, count) in the compute type; the division happens once, in the store
// ------------------------------------------------------------------------------------------------
template<typename Tc> struct MeanAcc {
    Tc s;
    uint n;
};

template<typename Ti, typename Tc, typename To> struct MeanPolicy {
    static constexpr bool kSplitAcc = false;
    typedef Ti In;
    typedef MeanAcc<Tc> Acc;
    typedef typename base_of<Tc>::type B;
    struct Params {};
    struct Out { To *val; };
    __host__ __device__ static Out host_out(const Out &, void *rec) { return Out{(To *)rec}; }
    AFB_DI static Acc init() { return Acc{zero_<Tc>(), 0u}; }
    AFB_DI static Acc load(const Params &, In x, uint) { return Acc{ToCompute<Tc, Ti>::go(x), 1u}; }
    AFB_DI static Acc join(Acc a, Acc b) { return Acc{add_<Tc>(a.s, b.s), a.n + b.n}; }
    AFB_DI static Acc globalize(Acc a, uint) { return a; }
    AFB_DI static void store(const Out &o, size_t i, Acc a) {
        o.val[i] = conv<To, Tc>(a.n ? scale_<Tc, B>(a.s, (B)1 / (B)a.n) : zero_<Tc>());
    }
    AFB_DI static Acc fold16(const Params &p, Acc acc, const Vec16 &v, uint idx0) {
        constexpr int V = VecN<In>::value;
        const In *e = reinterpret_cast<const In *>(&v);
#pragma unroll
        for (int k = 0; k < V; k++) acc.s = add_<Tc>(acc.s, ToCompute<Tc, Ti>::go(e[k]));
        acc.n += V;
        return acc;
    }
};

// ------------------------------------------------------------------------------------------------
// variance: (mean, M2 = sum of (x - mean)^2, count).  One element joins by Welford's update, two partials merge by the
// pairwise formula (Chan et al.); complex uses the plain product (x - m)(x - m), as the reference's diff * diff does.
// ------------------------------------------------------------------------------------------------
template<typename Tc> struct VarAcc {
    Tc m, m2;
    uint n;
};

template<typename To> struct VarOut {
    To *val;        // variance (or standard deviation)
    To *mean;       // optional second output (af_meanvar)
    double norm;    // 1 / (N - bias)
    int take_sqrt;
};

template<typename Ti, typename Tc, typename To> struct VarPolicy {
    static constexpr bool kSplitAcc = false;
    typedef Ti In;
    typedef VarAcc<Tc> Acc;
    typedef typename base_of<Tc>::type B;
    struct Params {};
    typedef VarOut<To> Out;
    __host__ __device__ static Out host_out(const Out &o, void *rec) { return Out{(To *)rec, nullptr, o.norm, o.take_sqrt}; }
    AFB_DI static Acc init() { return Acc{zero_<Tc>(), zero_<Tc>(), 0u}; }
    AFB_DI static Acc load(const Params &, In x, uint) { return Acc{ToCompute<Tc, Ti>::go(x), zero_<Tc>(), 1u}; }
    AFB_DI static Acc step(Acc a, Tc x) {
        a.n += 1u;
        const Tc d = sub_<Tc>(x, a.m);
        a.m = add_<Tc>(a.m, scale_<Tc, B>(d, (B)1 / (B)a.n));
        a.m2 = add_<Tc>(a.m2, mul_<Tc>(d, sub_<Tc>(x, a.m)));
        return a;
    }
    AFB_DI static Acc join(Acc a, Acc b) {
        if (b.n == 1u) return step(a, b.m);
        if (b.n == 0u) return a;
        if (a.n == 0u) return b;
        const uint n = a.n + b.n;
        const B fb = (B)b.n / (B)n;
        const Tc d = sub_<Tc>(b.m, a.m);
        Acc r;
        r.n = n;
        r.m = add_<Tc>(a.m, scale_<Tc, B>(d, fb));
        r.m2 = add_<Tc>(add_<Tc>(a.m2, b.m2), scale_<Tc, B>(mul_<Tc>(d, d), fb * (B)a.n));
        return r;
    }
    AFB_DI static Acc globalize(Acc a, uint) { return a; }
    AFB_DI static void store(const Out &o, size_t i, Acc a) {
        Tc v = scale_<Tc, B>(a.m2, (B)o.norm);
        if (o.take_sqrt) v = sqrt_<Tc>(v);
        o.val[i] = conv<To, Tc>(v);
        if (o.mean) o.mean[i] = conv<To, Tc>(a.m);
    }
    // a 16-byte vector at a time: mean and M2 of its V elements with compile-time weights, then ONE pairwise merge -- one
    // division per vector instead of one per element (the per-element Welford update ran at 3.1 TB/s on f32 columns)
    AFB_DI static Acc fold16(const Params &p, Acc acc, const Vec16 &v, uint idx0) {
        constexpr int V = VecN<In>::value;
        const In *e = reinterpret_cast<const In *>(&v);
        if constexpr (V == 1) {
            return step(acc, ToCompute<Tc, Ti>::go(e[0]));
        } else {
            Tc x[V];
            Tc sum = zero_<Tc>();
#pragma unroll
            for (int k = 0; k < V; k++) {
                x[k] = ToCompute<Tc, Ti>::go(e[k]);
                sum = add_<Tc>(sum, x[k]);
            }
            const Tc mb = scale_<Tc, B>(sum, (B)1 / (B)V);
            Tc m2b = zero_<Tc>();
#pragma unroll
            for (int k = 0; k < V; k++) {
                const Tc d = sub_<Tc>(x[k], mb);
                m2b = add_<Tc>(m2b, mul_<Tc>(d, d));
            }
            if (acc.n == 0u) return Acc{mb, m2b, (uint)V};
            const uint n = acc.n + (uint)V;
            const B fb = (B)V / (B)n;
            const Tc d = sub_<Tc>(mb, acc.m);
            Acc r;
            r.n = n;
            r.m = add_<Tc>(acc.m, scale_<Tc, B>(d, fb));
            r.m2 = add_<Tc>(add_<Tc>(acc.m2, m2b), scale_<Tc, B>(mul_<Tc>(d, d), fb * (B)acc.n));
            return r;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// norms: |x| (hypot for complex), optionally squared / raised to p, fused into the load of a sum or max reduction
// ------------------------------------------------------------------------------------------------
enum NormPre { NORM_ABS = 0, NORM_SQ = 1, NORM_POW = 2 };

struct NormParams {
    int pre;
    double p;
};

template<typename T> AFB_DI typename base_of<typename compute_of<T>::type>::type abs_of(T x) {
    return x < (T)0 ? -x : x;
}
template<> AFB_DI float abs_of<__half>(__half x) { return fabsf(__half2float(x)); }
template<> AFB_DI float abs_of<cfloat>(cfloat x) { return mag(x); }
template<> AFB_DI double abs_of<cdouble>(cdouble x) { return mag(x); }

// OP is OP_ADD or OP_MAX; R is the real result type (float / double)
template<int OP, typename Ti, typename R> struct NormPolicy {
    static constexpr bool kSplitAcc = false;
    typedef Ti In;
    typedef R Acc;
    typedef NormParams Params;
    struct Out { R *val; };
    __host__ __device__ static Out host_out(const Out &, void *rec) { return Out{(R *)rec}; }
    AFB_DI static Acc init() { return (R)0; }       // |x| >= 0: zero is the identity of both sum and max
    AFB_DI static Acc load(const Params &p, In x, uint) {
        R a = (R)abs_of<Ti>(x);
        if (p.pre == NORM_SQ) a = a * a;
        else if (p.pre == NORM_POW) a = (R)pow((double)a, p.p);
        return a;
    }
    AFB_DI static Acc join(Acc a, Acc b) {
        if (OP == OP_ADD) return a + b;
        return (b > a || b != b) ? b : a;            // a NaN propagates, as std::max over the serial loop would not hide it
    }
    AFB_DI static Acc globalize(Acc a, uint) { return a; }
    AFB_DI static void store(const Out &o, size_t i, Acc a) { o.val[i] = a; }
    AFB_DI static Acc fold16(const Params &p, Acc acc, const Vec16 &v, uint idx0) {
        constexpr int V = VecN<In>::value;
        const In *e = reinterpret_cast<const In *>(&v);
#pragma unroll
        for (int k = 0; k < V; k++) acc = join(acc, load(p, e[k], idx0 + k));
        return acc;
    }
};

}  // namespace afb
