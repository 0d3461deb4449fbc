// expr.cu -- SURVEY.md section 8(f)-4: a reduction whose input is an UNEVALUATED element-wise expression.
// The reference builds such expressions lazily (its JIT, src/backend/common/jit/*.hpp) and compiles a fused kernel with
// NVRTC when a reduction consumes them -- e.g. the pi benchmark, examples/benchmarks/pi.cpp:28-31:
//     sum<float>( sqrt(x*x + y*y) < 1 )
// Here the expression is a short postfix program (api_expr.cpp flattens the tree) that the reduction kernel INTERPRETS in
// its load path: no run-time compiler, no temporaries -- x and y are streamed once and only the count leaves the chip.
// Every thread evaluates the program on W = 32 (f32) / 16 (f64) elements at a time -- eight 16-byte chunks, interleaved with
// the other threads of the CTA so that every load coalesces -- with ONE decode per W elements: the top of the value stack
// lives in registers, the rest in shared memory ([depth][chunk][thread] 16-byte vectors: conflict-free, sized by the
// program's real depth).  (r02a: W = 8 consecutive elements: 74 executed instructions per element of the pi expression,
// most of them decode, branches and register moves between the cases.)
// The same interpreter materialises an expression into an array when a non-fused consumer needs it.
// (Measured and rejected: prefetching the next iteration's operands into registers -- 116 registers, 1.11 -> 1.67 ms on the
// pi expression over 2 x 2^28 f32: the interpreter is bound by instruction issue, not by load latency.)
#include "device_ops.cuh"

namespace afb {

constexpr int kExprThreads = 128;
constexpr int kExprDepth = 8;       // value-stack depth (entries below the register-resident top)
constexpr int kExprChunks = 8;      // 16-byte chunks per thread and decode: 128 bytes = 32 floats / 16 doubles

// A CTA evaluates blocks of kExprThreads * W elements; chunk q of thread t covers the C = 16 / sizeof(Tc) elements from
// block + (q * kExprThreads + t) * C on (consecutive lanes touch consecutive 16-byte vectors: every load / store coalesces).
template<typename Tc> struct ExprW {
    static constexpr int C = 16 / (int)sizeof(Tc);
    static constexpr int value = C * kExprChunks;
};
struct alignas(16) ExprVec { unsigned int w[4]; };
inline size_t expr_stack_bytes(int depth) { return (size_t)(depth < 1 ? 1 : depth) * kExprChunks * kExprThreads * sizeof(ExprVec); }

// integer operands (leaf types 3.. of ExprProg::leaf_type): only the double interpreter carries them -- every 8- / 16- / 32-bit
// integer is exact in f64, so comparisons and logic on them give the reference's answers bit for bit; flag 16 rounds the
// value through float first (an s32 / u32 operand compared with an f32 one: the reference casts it to float, implicit.cpp:20-45)
AFB_DI void expr_load_int(const void *base, int t, dim_t blk, dim_t n, bool whole, double (&v)[ExprW<double>::value]) {
    constexpr int C = ExprW<double>::C;
    const int bt = t & 15, tid = threadIdx.x;
    const bool via_float = (t & 16) != 0;
#pragma unroll
    for (int q = 0; q < kExprChunks; q++) {
#pragma unroll
        for (int k = 0; k < C; k++) {
            const dim_t i = blk + ((dim_t)q * kExprThreads + tid) * C + k;
            double x = 0.0;
            if (whole || i < n) {
                switch (bt) {
                    case 3: x = (double)__ldg((const int *)base + i); break;
                    case 4: x = (double)__ldg((const unsigned int *)base + i); break;
                    case 5: x = (double)__ldg((const short *)base + i); break;
                    case 6: x = (double)__ldg((const unsigned short *)base + i); break;
                    default: x = (double)__ldg((const unsigned char *)base + i); break;   // 7: u8 (s8 shares the b8 path)
                }
            }
            v[q * C + k] = via_float ? (double)(float)x : x;
        }
    }
}

template<typename Tc> AFB_DI void expr_load(const ExprProg &p, int leaf, dim_t blk, dim_t n, bool whole, Tc (&v)[ExprW<Tc>::value]) {
    constexpr int C = ExprW<Tc>::C;
    const void *base = p.leaf[leaf];
    const int t = p.leaf_type[leaf];
    const int tid = threadIdx.x;
    if constexpr (sizeof(Tc) == 8) {
        if (t >= 3) {
            expr_load_int(base, t, blk, n, whole, v);
            return;
        }
    }
    if (whole && p.aligned) {
        if (t == 0) {
            const float *f = (const float *)base + blk + (dim_t)tid * C;
#pragma unroll
            for (int q = 0; q < kExprChunks; q++) {
                if (C == 4) {
                    const float4 x = __ldg(reinterpret_cast<const float4 *>(f + (dim_t)q * kExprThreads * C));
                    v[q * C] = (Tc)x.x; v[q * C + 1] = (Tc)x.y; v[q * C + 2 % C] = (Tc)x.z; v[q * C + 3 % C] = (Tc)x.w;
                } else {
                    const float2 x = __ldg(reinterpret_cast<const float2 *>(f + (dim_t)q * kExprThreads * C));
                    v[q * C] = (Tc)x.x; v[q * C + 1] = (Tc)x.y;
                }
            }
        } else if (t == 1) {
            const double *f = (const double *)base + blk + (dim_t)tid * C;
#pragma unroll
            for (int q = 0; q < kExprChunks; q++) {
                if (C == 2) {
                    const double2 x = __ldg(reinterpret_cast<const double2 *>(f + (dim_t)q * kExprThreads * C));
                    v[q * C] = (Tc)x.x; v[q * C + 1] = (Tc)x.y;
                } else {
#pragma unroll
                    for (int k = 0; k < C; k++) v[q * C + k] = (Tc)__ldg(f + (dim_t)q * kExprThreads * C + k);
                }
            }
        } else {
            const signed char *f = (const signed char *)base + blk + (dim_t)tid * C;
#pragma unroll
            for (int q = 0; q < kExprChunks; q++) {
                if (C == 4) {
                    const char4 x = __ldg(reinterpret_cast<const char4 *>(f + (dim_t)q * kExprThreads * C));
                    v[q * C] = (Tc)x.x; v[q * C + 1] = (Tc)x.y; v[q * C + 2 % C] = (Tc)x.z; v[q * C + 3 % C] = (Tc)x.w;
                } else {
                    const char2 x = __ldg(reinterpret_cast<const char2 *>(f + (dim_t)q * kExprThreads * C));
                    v[q * C] = (Tc)x.x; v[q * C + 1] = (Tc)x.y;
                }
            }
        }
        return;
    }
#pragma unroll
    for (int q = 0; q < kExprChunks; q++) {
#pragma unroll
        for (int k = 0; k < C; k++) {
            const dim_t i = blk + ((dim_t)q * kExprThreads + tid) * C + k;
            Tc x = (Tc)0;
            if (i < n) x = t == 0 ? (Tc)__ldg((const float *)base + i) : (t == 1 ? (Tc)__ldg((const double *)base + i) : (Tc)__ldg((const signed char *)base + i));
            v[q * C + k] = x;
        }
    }
}

// one binary instruction on W elements.  KIND 0: left operand popped from the shared-memory stack (16 bytes per chunk),
// right operand = top; KIND 1: top (op) constant; KIND 2: constant (op) top.  The switch sits OUTSIDE the element loops.
template<typename Tc, int KIND>
AFB_DI void expr_binary(int op, Tc (&top)[ExprW<Tc>::value], const ExprVec *slot, Tc cv) {
    constexpr int C = ExprW<Tc>::C;
    auto run = [&](auto f) {
#pragma unroll
        for (int q = 0; q < kExprChunks; q++) {
            Tc o[C];
            if (KIND == 0) {
                const ExprVec x = slot[q * kExprThreads];
                memcpy(o, &x, sizeof(x));
            }
#pragma unroll
            for (int k = 0; k < C; k++) {
                Tc &t = top[q * C + k];
                const Tc a = KIND == 0 ? o[k] : (KIND == 1 ? t : cv);
                const Tc b = KIND == 0 ? t : (KIND == 1 ? cv : t);
                t = f(a, b);
            }
        }
    };
    switch (op) {
        case EX_ADD: run([](Tc a, Tc b) { return a + b; }); break;
        case EX_SUB: run([](Tc a, Tc b) { return a - b; }); break;
        case EX_MUL: run([](Tc a, Tc b) { return a * b; }); break;
        case EX_DIV: run([](Tc a, Tc b) { return a / b; }); break;
        case EX_LT: run([](Tc a, Tc b) { return a < b ? (Tc)1 : (Tc)0; }); break;
        case EX_GT: run([](Tc a, Tc b) { return a > b ? (Tc)1 : (Tc)0; }); break;
        case EX_LE: run([](Tc a, Tc b) { return a <= b ? (Tc)1 : (Tc)0; }); break;
        case EX_GE: run([](Tc a, Tc b) { return a >= b ? (Tc)1 : (Tc)0; }); break;
        case EX_EQ: run([](Tc a, Tc b) { return a == b ? (Tc)1 : (Tc)0; }); break;
        case EX_NEQ: run([](Tc a, Tc b) { return a != b ? (Tc)1 : (Tc)0; }); break;
        case EX_AND: run([](Tc a, Tc b) { return (a != (Tc)0 && b != (Tc)0) ? (Tc)1 : (Tc)0; }); break;
        case EX_OR: run([](Tc a, Tc b) { return (a != (Tc)0 || b != (Tc)0) ? (Tc)1 : (Tc)0; }); break;
        case EX_MIN: run([](Tc a, Tc b) { return (b < a) ? b : a; }); break;
        case EX_MAX: run([](Tc a, Tc b) { return (a < b) ? b : a; }); break;
        default: run([](Tc a, Tc b) { return (Tc)pow(a, b); }); break;   // EX_POW
    }
}

// evaluates the program for the block starting at element `blk`; result in top[]
template<typename Tc>
AFB_DI void expr_eval(const ExprProg &p, dim_t blk, dim_t n, bool whole, Tc (&top)[ExprW<Tc>::value], ExprVec *stk /* [depth][chunk][thread] */) {
    constexpr int W = ExprW<Tc>::value, C = ExprW<Tc>::C;
    ExprVec *sptr = stk + threadIdx.x;   // this thread's column of the next free stack entry
#pragma unroll 1
    for (int pc = 0; pc < p.n; pc++) {
        const int opf = p.op[pc], arg = p.arg[pc];
        const int op = opf & EX_OPMASK;
#ifdef AFB_EXPR_TRACE
        if (threadIdx.x == 0 && blockIdx.x == 0) printf("pc %d opf %x arg %d sp %d n %d depth %d\n", pc, opf, arg, sp, p.n, p.depth);
#endif
        if (op == EX_LEAF || op == EX_CONST) {
            if (pc > 0) {      // spill the current top
#pragma unroll
                for (int q = 0; q < kExprChunks; q++) {
                    ExprVec x;
                    memcpy(&x, &top[q * C], sizeof(x));
                    sptr[q * kExprThreads] = x;
                }
                sptr += kExprChunks * kExprThreads;
            }
            if (op == EX_LEAF) expr_load<Tc>(p, arg, blk, n, whole, top);
            else {
                const Tc cv = (Tc)p.cst[arg];
#pragma unroll
                for (int w = 0; w < W; w++) top[w] = cv;
            }
            continue;
        }
        if (op >= EX_FIRST_UNARY) {
            auto run = [&](auto f) {
#pragma unroll
                for (int w = 0; w < W; w++) top[w] = f(top[w]);
            };
            switch (op) {
                case EX_SQRT: run([](Tc a) { return (Tc)sqrt(a); }); break;
                case EX_ABS: run([](Tc a) { return a < (Tc)0 ? -a : a; }); break;
                case EX_NEG: run([](Tc a) { return -a; }); break;
                case EX_SQR: run([](Tc a) { return a * a; }); break;
                default: run([](Tc a) { return (a == (Tc)0) ? (Tc)1 : (Tc)0; }); break;   // EX_NOT
            }
            continue;
        }
        if (opf & EX_RCONST) expr_binary<Tc, 1>(op, top, nullptr, (Tc)p.cst[arg]);
        else if (opf & EX_LCONST) expr_binary<Tc, 2>(op, top, nullptr, (Tc)p.cst[arg]);
        else {
            sptr -= kExprChunks * kExprThreads;
            expr_binary<Tc, 0>(op, top, sptr, (Tc)0);
        }
    }
}

// ---- materialise: out[i] = expr(i); out_type 0 f32, 1 f64, 2 b8
template<typename Tc>
__global__ void __launch_bounds__(kExprThreads) expr_eval_kernel(const ExprProg p, void *out, int out_type, dim_t n) {
    constexpr int W = ExprW<Tc>::value, C = ExprW<Tc>::C;
    extern __shared__ __align__(16) unsigned char expr_smem[];
    ExprVec *stk = reinterpret_cast<ExprVec *>(expr_smem);
    const int tid = threadIdx.x;
    const dim_t per = (dim_t)kExprThreads * W;
    for (dim_t blk = (dim_t)blockIdx.x * per; blk < n; blk += (dim_t)gridDim.x * per) {
        const bool whole = blk + per <= n;
        Tc top[W];
        expr_eval<Tc>(p, blk, n, whole, top, stk);
#pragma unroll
        for (int q = 0; q < kExprChunks; q++) {
            const dim_t i0 = blk + ((dim_t)q * kExprThreads + tid) * C;
            if (whole && out_type == 0 && C == 4) {
                *reinterpret_cast<float4 *>((float *)out + i0) = make_float4((float)top[q * C], (float)top[q * C + 1], (float)top[q * C + 2 % C], (float)top[q * C + 3 % C]);
            } else if (whole && out_type == 1 && C == 2) {
                *reinterpret_cast<double2 *>((double *)out + i0) = make_double2((double)top[q * C], (double)top[q * C + 1]);
            } else {
#pragma unroll
                for (int k = 0; k < C; k++) {
                    if (i0 + k >= n) break;
                    const Tc x = top[q * C + k];
                    if (out_type == 0) ((float *)out)[i0 + k] = (float)x;
                    else if (out_type == 1) ((double *)out)[i0 + k] = (double)x;
                    else ((signed char *)out)[i0 + k] = (signed char)(x != (Tc)0);
                }
            }
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
    constexpr int W = ExprW<Tc>::value, C = ExprW<Tc>::C;
    extern __shared__ __align__(16) unsigned char expr_smem[];
    ExprVec *stk = reinterpret_cast<ExprVec *>(expr_smem);
    __shared__ ExprPart<Tc> sm_w[kExprThreads / 32];
    __shared__ int sm_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Tc vinit = mode == EXR_MIN ? (Tc)INFINITY : (mode == EXR_MAX ? (Tc)-INFINITY : (Tc)0);
    Tc acc = vinit;
    uint cnt = 0;
    const dim_t per = (dim_t)kExprThreads * W;
    for (dim_t blk = (dim_t)blockIdx.x * per; blk < n; blk += (dim_t)gridDim.x * per) {
        const bool whole = blk + per <= n;
        Tc top[W];
        expr_eval<Tc>(p, blk, n, whole, top, stk);
        if (whole && (mode == EXR_COUNT || mode == EXR_ANY || mode == EXR_ALL)) {
#pragma unroll
            for (int w = 0; w < W; w++) cnt += (top[w] != (Tc)0) ? 1u : 0u;
        } else {
#pragma unroll
            for (int q = 0; q < kExprChunks; q++) {
#pragma unroll
                for (int k = 0; k < C; k++) {
                    if (!whole && blk + ((dim_t)q * kExprThreads + tid) * C + k >= n) continue;
                    const Tc x = top[q * C + k];
                    if (mode == EXR_SUM) acc += x;
                    else if (mode == EXR_MIN) acc = (x != x) ? acc : (x < acc ? x : acc);     // NaN -> identity, as af_min
                    else if (mode == EXR_MAX) acc = (x != x) ? acc : (acc < x ? x : acc);
                    cnt += (x != (Tc)0) ? 1u : 0u;
                }
            }
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

template<typename K> static unsigned expr_grid(const LaunchCtx &c, K kern, dim_t n, int w, size_t smem, int *cache /* [kExprDepth + 1], per kernel */) {
    const int slot = (int)(smem / expr_stack_bytes(1));
    if (!cache[slot]) {
        AFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expr_stack_bytes(kExprDepth)));
        int per_sm = 0;
        AFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kExprThreads, smem));
        cache[slot] = per_sm < 1 ? 1 : per_sm;
    }
    dim_t g = (n + (dim_t)kExprThreads * w - 1) / ((dim_t)kExprThreads * w);
    const dim_t cap = (dim_t)c.sms * cache[slot];      // one wave of resident CTAs, each loops over its blocks
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int occ[64][4][kExprDepth + 2];   // resident CTAs per SM by device, kernel and stack depth (the shared-memory attribute is per device)

void launch_expr_eval(const LaunchCtx &c, const ExprProg &p, bool dbl, void *out, int out_type, dim_t n) {
    if (n <= 0) return;
    const size_t sm = expr_stack_bytes(p.depth);
    if (dbl) {
        expr_eval_kernel<double><<<expr_grid(c, expr_eval_kernel<double>, n, ExprW<double>::value, sm, occ[c.device & 63][0]), kExprThreads, sm, c.stream>>>(p, out, out_type, n);
    } else {
        expr_eval_kernel<float><<<expr_grid(c, expr_eval_kernel<float>, n, ExprW<float>::value, sm, occ[c.device & 63][1]), kExprThreads, sm, c.stream>>>(p, out, out_type, n);
    }
    note_launches(1);
    post_launch_check("expr_eval_kernel");
}

void launch_expr_reduce(const LaunchCtx &c, const ExprProg &p, bool dbl, int mode, dim_t n, void *out, const HostRet *ret) {
    uint *ticket = c.scratch->need_tickets(3) + 2;   // tickets 0 / 1 belong to the whole-array and by-key reductions
    const size_t sm = expr_stack_bytes(p.depth);
    if (getenv("AFB200_EXPR_TRACE")) {
        fprintf(stderr, "[expr] n=%d depth=%d aligned=%d dbl=%d smem=%zu:", p.n, p.depth, p.aligned, (int)dbl, sm);
        for (int i = 0; i < p.n; i++) fprintf(stderr, " %02x/%d", p.op[i], p.arg[i]);
        fprintf(stderr, "\n");
    }
    if (dbl) {
        const unsigned grid = expr_grid(c, expr_reduce_kernel<double>, n, ExprW<double>::value, sm, occ[c.device & 63][2]);
        auto *parts = (ExprPart<double> *)c.scratch->need_partials((size_t)grid * sizeof(ExprPart<double>));
        expr_reduce_kernel<double><<<grid, kExprThreads, sm, c.stream>>>(p, mode, n, out, parts, ticket, ret ? ret->wire : nullptr,
                                                                         ret ? ret->seq : 0u);
    } else {
        const unsigned grid = expr_grid(c, expr_reduce_kernel<float>, n, ExprW<float>::value, sm, occ[c.device & 63][3]);
        auto *parts = (ExprPart<float> *)c.scratch->need_partials((size_t)grid * sizeof(ExprPart<float>));
        expr_reduce_kernel<float><<<grid, kExprThreads, sm, c.stream>>>(p, mode, n, out, parts, ticket, ret ? ret->wire : nullptr,
                                                                        ret ? ret->seq : 0u);
    }
    note_launches(1);
    post_launch_check("expr_reduce_kernel");
}

}  // namespace afb
