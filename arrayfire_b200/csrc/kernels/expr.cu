// expr.cu -- SURVEY.md section 8(f)-4: a reduction whose input is an UNEVALUATED element-wise expression.
// The reference builds such expressions lazily (its JIT, src/backend/common/jit/*.hpp) and compiles a fused kernel with
// NVRTC when a reduction consumes them -- e.g. the pi benchmark, examples/benchmarks/pi.cpp:28-31:
//     sum<float>( sqrt(x*x + y*y) < 1 )
// Here the expression is a short postfix program (api_expr.cpp flattens the tree) that the reduction kernel INTERPRETS in
// its load path: no run-time compiler, no temporaries -- x and y are streamed once and only the count leaves the chip.
// Every thread evaluates the program on W = 8 (f32) / 4 (f64) consecutive elements at a time, one decode per W elements: the top of the value stack lives in registers,
// the rest in shared memory ([depth][w][thread]: conflict-free), so one decoded instruction is amortised over W elements.
// The same interpreter materialises an expression into an array when a non-fused consumer needs it.
// (Measured and rejected: prefetching the next iteration's operands into registers -- 116 registers, 1.11 -> 1.67 ms on the
// pi expression over 2 x 2^28 f32: the interpreter is bound by instruction issue, not by load latency.)
#include "device_ops.cuh"

namespace afb {

constexpr int kExprThreads = 128;
constexpr int kExprDepth = 8;       // value-stack depth (entries below the register-resident top)

template<typename Tc> struct ExprW { static constexpr int value = 32 / (int)sizeof(Tc); };   // 8 floats / 4 doubles per decode

template<typename Tc> AFB_DI void expr_load(const ExprProg &p, int leaf, dim_t i, dim_t n, Tc (&v)[ExprW<Tc>::value]) {
    constexpr int W = ExprW<Tc>::value;
    const void *base = p.leaf[leaf];
    const int t = p.leaf_type[leaf];
    if (t == 0) {
        const float *f = (const float *)base;
        if (p.aligned && i + W <= n) {
#pragma unroll
            for (int q = 0; q < W / 4; q++) {
                const float4 v4 = __ldg(reinterpret_cast<const float4 *>(f + i) + q);
                v[4 * q] = (Tc)v4.x; v[4 * q + 1] = (Tc)v4.y; v[4 * q + 2] = (Tc)v4.z; v[4 * q + 3] = (Tc)v4.w;
            }
        } else {
#pragma unroll
            for (int w = 0; w < W; w++) v[w] = i + w < n ? (Tc)__ldg(f + i + w) : (Tc)0;
        }
    } else if (t == 1) {
        const double *f = (const double *)base;
#pragma unroll
        for (int w = 0; w < W; w++) v[w] = i + w < n ? (Tc)__ldg(f + i + w) : (Tc)0;
    } else {
        const signed char *f = (const signed char *)base;
#pragma unroll
        for (int w = 0; w < W; w++) v[w] = i + w < n ? (Tc)__ldg(f + i + w) : (Tc)0;
    }
}

// evaluates the program for elements [i, i + W); result in top[]
template<typename Tc>
AFB_DI void expr_eval(const ExprProg &p, dim_t i, dim_t n, Tc (&top)[ExprW<Tc>::value], Tc *stk /* [depth][W][threads] */) {
    constexpr int W = ExprW<Tc>::value;
    const int tid = threadIdx.x;
    int sp = 0;          // entries in shared memory
#pragma unroll 1
    for (int pc = 0; pc < p.n; pc++) {
        const int opf = p.op[pc], arg = p.arg[pc];
        const int op = opf & EX_OPMASK;
        if (op == EX_LEAF || op == EX_CONST) {
            if (pc > 0) {      // spill the current top
#pragma unroll
                for (int w = 0; w < W; w++) stk[(sp * W + w) * kExprThreads + tid] = top[w];
                sp++;
            }
            if (op == EX_LEAF) expr_load<Tc>(p, arg, i, n, top);
            else {
#pragma unroll
                for (int w = 0; w < W; w++) top[w] = (Tc)p.cst[arg];
            }
            continue;
        }
        // the switch sits OUTSIDE the W loops: one decode per W elements
#define EX_UN(expr)                              \
    _Pragma("unroll") for (int w = 0; w < W; w++) { \
        const Tc a = top[w];                     \
        top[w] = (expr);                         \
    }                                            \
    break;
        if (op >= EX_FIRST_UNARY) {
            switch (op) {
                case EX_SQRT: EX_UN(sqrt(a))
                case EX_ABS: EX_UN(a < (Tc)0 ? -a : a)
                case EX_NEG: EX_UN(-a)
                case EX_SQR: EX_UN(a * a)
                default: EX_UN((a == (Tc)0) ? (Tc)1 : (Tc)0)   // EX_NOT
            }
            continue;
        }
#undef EX_UN
        // binary: left operand from shared memory, right operand = top -- unless one side is a scalar constant
        Tc l[W], r[W];
        if (opf & EX_RCONST) {
            const Tc cv = (Tc)p.cst[arg];
#pragma unroll
            for (int w = 0; w < W; w++) { l[w] = top[w]; r[w] = cv; }
        } else if (opf & EX_LCONST) {
            const Tc cv = (Tc)p.cst[arg];
#pragma unroll
            for (int w = 0; w < W; w++) { l[w] = cv; r[w] = top[w]; }
        } else {
            sp--;
#pragma unroll
            for (int w = 0; w < W; w++) { l[w] = stk[(sp * W + w) * kExprThreads + tid]; r[w] = top[w]; }
        }
#define EX_BIN(expr)                             \
    _Pragma("unroll") for (int w = 0; w < W; w++) { \
        const Tc a = l[w], b = r[w];             \
        top[w] = (expr);                         \
    }                                            \
    break;
        switch (op) {
            case EX_ADD: EX_BIN(a + b)
            case EX_SUB: EX_BIN(a - b)
            case EX_MUL: EX_BIN(a * b)
            case EX_DIV: EX_BIN(a / b)
            case EX_LT: EX_BIN(a < b ? (Tc)1 : (Tc)0)
            case EX_GT: EX_BIN(a > b ? (Tc)1 : (Tc)0)
            case EX_LE: EX_BIN(a <= b ? (Tc)1 : (Tc)0)
            case EX_GE: EX_BIN(a >= b ? (Tc)1 : (Tc)0)
            case EX_EQ: EX_BIN(a == b ? (Tc)1 : (Tc)0)
            case EX_NEQ: EX_BIN(a != b ? (Tc)1 : (Tc)0)
            case EX_AND: EX_BIN((a != (Tc)0 && b != (Tc)0) ? (Tc)1 : (Tc)0)
            case EX_OR: EX_BIN((a != (Tc)0 || b != (Tc)0) ? (Tc)1 : (Tc)0)
            case EX_MIN: EX_BIN((b < a) ? b : a)
            case EX_MAX: EX_BIN((a < b) ? b : a)
            default: EX_BIN(pow(a, b))                          // EX_POW
        }
#undef EX_BIN
    }
}

// ---- materialise: out[i] = expr(i); out_type 0 f32, 1 f64, 2 b8
template<typename Tc>
__global__ void __launch_bounds__(kExprThreads) expr_eval_kernel(const ExprProg p, void *out, int out_type, dim_t n) {
    constexpr int W = ExprW<Tc>::value;
    extern __shared__ __align__(16) unsigned char expr_smem[];
    Tc *stk = reinterpret_cast<Tc *>(expr_smem);
    const dim_t stride = (dim_t)gridDim.x * kExprThreads * W;
    for (dim_t i = ((dim_t)blockIdx.x * kExprThreads + threadIdx.x) * W; i < n; i += stride) {
        Tc top[W];
        expr_eval<Tc>(p, i, n, top, stk);
#pragma unroll
        for (int w = 0; w < W; w++) {
            if (i + w >= n) break;
            if (out_type == 0) ((float *)out)[i + w] = (float)top[w];
            else if (out_type == 1) ((double *)out)[i + w] = (double)top[w];
            else ((signed char *)out)[i + w] = (signed char)(top[w] != (Tc)0);
        }
    }
}

// ---- fused reduction: mode EXR_SUM (value sum), EXR_COUNT (non-zero count), EXR_MIN, EXR_MAX, EXR_ANY, EXR_ALL.
// Result: `Tc` for sum / min / max, u32 for count, b8 for any / all; device `out` + pinned host record.
template<typename Tc> struct ExprPart {
    Tc v;
    uint c;
};

template<typename Tc>
__global__ void __launch_bounds__(kExprThreads) expr_reduce_kernel(const ExprProg p, int mode, dim_t n, void *out,
                                                                    ExprPart<Tc> *partials, uint *ticket, void *host_wire,
                                                                    uint host_seq) {
    constexpr int W = ExprW<Tc>::value;
    extern __shared__ __align__(16) unsigned char expr_smem[];
    Tc *stk = reinterpret_cast<Tc *>(expr_smem);
    __shared__ ExprPart<Tc> sm_w[kExprThreads / 32];
    __shared__ int sm_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Tc vinit = mode == EXR_MIN ? (Tc)INFINITY : (mode == EXR_MAX ? (Tc)-INFINITY : (Tc)0);
    Tc acc = vinit;
    uint cnt = 0;
    const dim_t stride = (dim_t)gridDim.x * kExprThreads * W;
    for (dim_t i = ((dim_t)blockIdx.x * kExprThreads + tid) * W; i < n; i += stride) {
        Tc top[W];
        expr_eval<Tc>(p, i, n, top, stk);
#pragma unroll
        for (int w = 0; w < W; w++) {
            if (i + w >= n) break;
            const Tc x = top[w];
            if (mode == EXR_SUM) acc += x;
            else if (mode == EXR_MIN) acc = (x != x) ? acc : (x < acc ? x : acc);     // NaN -> identity, as af_min
            else if (mode == EXR_MAX) acc = (x != x) ? acc : (acc < x ? x : acc);
            cnt += (x != (Tc)0) ? 1u : 0u;
        }
    }
    auto fold = [&](ExprPart<Tc> a, ExprPart<Tc> b) {
        ExprPart<Tc> r;
        r.c = a.c + b.c;
        r.v = mode == EXR_MIN ? (b.v < a.v ? b.v : a.v) : (mode == EXR_MAX ? (a.v < b.v ? b.v : a.v) : a.v + b.v);
        return r;
    };
    ExprPart<Tc> me{acc, cnt};
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        ExprPart<Tc> o;
        o.v = __shfl_down_sync(0xffffffffu, me.v, d);
        o.c = __shfl_down_sync(0xffffffffu, me.c, d);
        me = fold(me, o);
    }
    if (lane == 0) sm_w[warp] = me;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kExprThreads / 32; w++) me = fold(me, sm_w[w]);
        partials[blockIdx.x] = me;
        __threadfence();
        sm_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!sm_last || tid != 0) return;
    __threadfence();
    ExprPart<Tc> fin{vinit, 0u};
    for (uint b = 0; b < gridDim.x; b++) {                  // <= a few thousand partials, fixed order
        ExprPart<Tc> q;
        q.v = __ldcg(&partials[b].v);
        q.c = __ldcg(&partials[b].c);
        fin = fold(fin, q);
    }
    *ticket = 0u;
    alignas(16) uint w[10];
#pragma unroll
    for (int j = 0; j < 10; j++) w[j] = 0u;
    if (mode == EXR_COUNT) {
        *(uint *)out = fin.c;
        w[0] = fin.c;
    } else if (mode == EXR_ANY || mode == EXR_ALL) {
        const signed char r = mode == EXR_ANY ? (fin.c > 0u) : ((dim_t)fin.c == n);
        *(signed char *)out = r;
        w[0] = (uint)(unsigned char)r;
    } else {
        *(Tc *)out = fin.v;
        memcpy(w, &fin.v, sizeof(Tc));
    }
    if (host_wire) host_ret_send(host_wire, host_seq, w);
}

static unsigned expr_grid(const LaunchCtx &c, dim_t n, int w) {
    dim_t g = (n + (dim_t)kExprThreads * w - 1) / ((dim_t)kExprThreads * w);
    const dim_t cap = (dim_t)c.sms * 6;      // 32 KB of value stack per CTA of 128 threads: 6 CTAs per SM
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

void launch_expr_eval(const LaunchCtx &c, const ExprProg &p, bool dbl, void *out, int out_type, dim_t n) {
    if (n <= 0) return;
    if (dbl) {
        const size_t sm = (size_t)kExprDepth * ExprW<double>::value * kExprThreads * sizeof(double);
        static bool attr = false;
        if (!attr) { AFB_CUDA(cudaFuncSetAttribute(expr_eval_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); attr = true; }
        expr_eval_kernel<double><<<expr_grid(c, n, ExprW<double>::value), kExprThreads, sm, c.stream>>>(p, out, out_type, n);
    } else {
        const size_t sm = (size_t)kExprDepth * ExprW<float>::value * kExprThreads * sizeof(float);
        expr_eval_kernel<float><<<expr_grid(c, n, ExprW<float>::value), kExprThreads, sm, c.stream>>>(p, out, out_type, n);
    }
    note_launches(1);
    post_launch_check("expr_eval_kernel");
}

void launch_expr_reduce(const LaunchCtx &c, const ExprProg &p, bool dbl, int mode, dim_t n, void *out, const HostRet *ret) {
    uint *ticket = c.scratch->need_tickets(3) + 2;   // tickets 0 / 1 belong to the whole-array and by-key reductions
    if (dbl) {
        const size_t sm = (size_t)kExprDepth * ExprW<double>::value * kExprThreads * sizeof(double);
        const unsigned grid = expr_grid(c, n, ExprW<double>::value);
        auto *parts = (ExprPart<double> *)c.scratch->need_partials((size_t)grid * sizeof(ExprPart<double>));
        expr_reduce_kernel<double><<<grid, kExprThreads, sm, c.stream>>>(p, mode, n, out, parts, ticket, ret ? ret->wire : nullptr,
                                                                         ret ? ret->seq : 0u);
    } else {
        const size_t sm = (size_t)kExprDepth * ExprW<float>::value * kExprThreads * sizeof(float);
        const unsigned grid = expr_grid(c, n, ExprW<float>::value);
        auto *parts = (ExprPart<float> *)c.scratch->need_partials((size_t)grid * sizeof(ExprPart<float>));
        expr_reduce_kernel<float><<<grid, kExprThreads, sm, c.stream>>>(p, mode, n, out, parts, ticket, ret ? ret->wire : nullptr,
                                                                        ret ? ret->seq : 0u);
    }
    note_launches(1);
    post_launch_check("expr_reduce_kernel");
}

}  // namespace afb
