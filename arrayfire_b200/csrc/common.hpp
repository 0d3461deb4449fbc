// common.hpp -- shared host-side declarations of libafb200: the array handle, the
// error model, the device/stream manager, the memory pool, and the type-erased
// launcher interface between the C-API layer (api_*.cpp) and the kernels (kernels/*.cu).
//
// Reference counterparts (studied for behaviour, not translated):
//   handle / ArrayInfo   src/backend/common/ArrayInfo.hpp:26-48, src/backend/cuda/Array.hpp:141-309
//   error model          src/backend/common/err_common.hpp:149-208, src/api/c/error.cpp:19-45
//   device + stream      src/backend/cuda/platform.cpp:354-408,546-556
//   memory manager       src/backend/common/DefaultMemoryManager.cpp:158-282
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>

#include "../../include/afb200.h"

namespace afb {

// ---------------------------------------------------------------------------------------------
// errors: C++ exceptions inside, af_err + thread-local message at the C boundary
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    af_err code;
    Error(af_err c, const std::string &m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] void fail(af_err code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
af_err store_error(af_err code, const char *msg) noexcept;  // sets the thread-local message
af_err current_exception_to_err() noexcept;                 // the CATCHALL of this library

#define AFB_ARG(index, cond)                                                                        \
    do {                                                                                            \
        if (!(cond)) ::afb::fail(AF_ERR_ARG, "%s: invalid argument %d, expected: %s", __func__, index, #cond); \
    } while (0)
#define AFB_TYPE_ERROR(index, type) \
    ::afb::fail(AF_ERR_TYPE, "%s: invalid type (%d) for argument %d", __func__, (int)(type), index)
#define AFB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            ::afb::fail(e_ == cudaErrorMemoryAllocation ? AF_ERR_NO_MEM : AF_ERR_RUNTIME,           \
                        "CUDA error %d (%s) at %s:%d: %s", (int)e_, cudaGetErrorName(e_), __FILE__, __LINE__, #call); \
    } while (0)
// every public entry point is an NVTX range named after the function (visible in Nsight Systems / Compute timelines, like
// the reference's AF_TRACE spans; a no-op costing a few nanoseconds when no tool is attached)
struct NvtxScope {
    explicit NvtxScope(const char *name);
    ~NvtxScope();
};
#define AFB_API_BEGIN try { ::afb::NvtxScope afb_nvtx_scope_(__func__);
#define AFB_API_END                                    \
    }                                                  \
    catch (...) { return ::afb::current_exception_to_err(); } \
    return AF_SUCCESS;

// ---------------------------------------------------------------------------------------------
// dtype helpers
// ---------------------------------------------------------------------------------------------
inline size_t dtype_size(af_dtype t) {
    switch (t) {
        case f32: case s32: case u32: return 4;
        case f64: case c32: case s64: case u64: return 8;
        case c64: return 16;
        case b8: case u8: case s8: return 1;
        case s16: case u16: case f16: return 2;
    }
    return 0;
}
inline bool dtype_valid(int t) { return t >= 0 && t <= (int)s8; }

// ---------------------------------------------------------------------------------------------
// devices, streams, scratch
// ---------------------------------------------------------------------------------------------
struct Scratch;  // per-device persistent workspace for partials / tickets / tile descriptors

int device_count();
int active_device();            // thread-local, default AF_CUDA_DEFAULT_DEVICE or 0
void set_active_device(int d);  // also cudaSetDevice
cudaStream_t stream_of(int device);
int sm_count(int device);
Scratch &scratch_of(int device);
std::mutex &op_mutex(int device);  // serialises multi-kernel enqueue sequences that share Scratch
void note_launches(unsigned n);    // launch accounting for afx_launch_count
unsigned long long launches_total(bool reset);
std::string take_last_error();     // fetch + clear the calling thread's error message
void post_launch_check(const char *what);  // cudaGetLastError (+ sync when AF_SYNCHRONOUS_CALLS=1)

// ---------------------------------------------------------------------------------------------
// memory pool
// ---------------------------------------------------------------------------------------------
struct Buffer {
    void *ptr = nullptr;
    size_t bytes = 0;       // allocated (rounded) size
    int device = 0;
    bool pooled = true;     // came from pool_alloc (false: adopted by af_device_array)
    bool user_locked = false;  // af_lock_array / af_get_device_ptr: the user owns the lifetime
    ~Buffer();
};
std::shared_ptr<Buffer> pool_alloc(size_t bytes);            // on the active device
std::shared_ptr<Buffer> adopt_device_ptr(void *ptr, size_t bytes);
void *pool_alloc_raw(size_t bytes);                          // af_alloc_device
void pool_free_raw(void *ptr);
void pool_gc(int device);
void pool_stats(int device, size_t *alloc_bytes, size_t *alloc_buffers, size_t *lock_bytes, size_t *lock_buffers);
void pool_set_step(size_t step);
size_t pool_get_step();

// ---------------------------------------------------------------------------------------------
// the array handle
// ---------------------------------------------------------------------------------------------
// The first bytes mirror the reference's ArrayInfo so that code which peeks at
// `devId` through an af_array (the unified loader's backend check,
// src/api/unified/symbol_manager.hpp:90-103) keeps working.
struct Info {
    unsigned devId;  // device | (AF_BACKEND_CUDA << 8), ArrayInfo.cpp:48-52,87-99
    af_dtype type;
    dim_t dims[4];
    dim_t offset;  // elements
    dim_t strides[4];
    bool is_sparse;
};

constexpr unsigned kMagic = 0xAFB20001u;

struct ExprNode;   // api_expr.cpp

struct Array {
    Info info;
    unsigned magic = kMagic;
    std::shared_ptr<Buffer> buf;  // shared between retained handles and views
    std::shared_ptr<ExprNode> expr;   // set (and buf empty) while the array is an unevaluated element-wise expression

    int device() const { return (int)(info.devId & 0xff); }
    dim_t elements() const { return info.dims[0] * info.dims[1] * info.dims[2] * info.dims[3]; }
    int ndims() const;     // dim4::ndims semantics (0 for empty, 1 for scalar)
    bool is_linear() const;
    const void *data() const { return (const char *)buf->ptr + (size_t)info.offset * dtype_size(info.type); }
    void *data() { return (char *)buf->ptr + (size_t)info.offset * dtype_size(info.type); }
};

Array *from_handle(af_array h);              // validates; throws AF_ERR_ARG on null/bad magic; EVALUATES a lazy expression
Array *peek_handle(af_array h);              // the same without evaluating (metadata getters, the fused consumers)
void materialize(Array &a);                  // api_expr.cpp: expression -> dense buffer (in place)
Array *new_array(af_dtype t, const dim_t dims[4]);  // dense, freshly allocated on the active device
Array *new_view(const Array &parent, const dim_t dims[4], dim_t offset, const dim_t strides[4]);
Array *retain(const Array &a);
void dense_strides(const dim_t dims[4], dim_t strides[4]);
void check_same_device(const Array &a);      // an array living on another device is MOVED to the active one (cuda/Array.cpp:61-74)

// ---------------------------------------------------------------------------------------------
// type-erased kernel launchers (kernels/*.cu).  Views carry pointers with the offset applied.
// Internal op codes: ADD..MAX coincide with af_binary_op.
// ---------------------------------------------------------------------------------------------
enum Op { OP_ADD = 0, OP_MUL = 1, OP_MIN = 2, OP_MAX = 3, OP_NOTZERO = 4, OP_AND = 5, OP_OR = 6 };

struct View {
    const void *ptr;
    dim_t dims[4];
    dim_t strides[4];
};
View view_of(const Array &a);

struct LaunchCtx {
    int device;
    cudaStream_t stream;
    int sms;
    Scratch *scratch;
};
LaunchCtx ctx_for_active_device();
// blocking copies on dedicated streams (runtime.cpp): overlap with compute and with the opposite direction
void upload_sync(int device, void *dst, const void *src, size_t bytes);
void download_sync(int device, void *dst, const void *src, size_t bytes);

// A scalar coming back from a kernel WITHOUT cudaMemcpyAsync + cudaStreamSynchronize: the kernel's last thread writes the
// payload (<= 40 bytes) into a pinned, device-mapped wire record as five 16-byte lines {data, seq, data, seq} -- flag in
// line, so the device needs NO system-scope fence (two MEMBAR.SYS cost 4.3 us per op, measured) -- and the host polls the
// lines until both flag words of each carry `seq`, then unpacks the payload into `host` (wait_host_ret).  `dev` is a
// 64-byte device slot paired with the record for kernels that also keep the value on the device (the hit count of
// af_where).  Records come from a small ring (Scratch::next_host_ret, under the device's op mutex); `seq` is monotone per
// device and never 0.
constexpr int kHostRetLines = 5;   // 10 payload words: value at +0 (<= 16 bytes), index at +32
struct HostRet {
    void *host;        // the unpacked payload (plain host memory; valid after wait_host_ret)
    void *wire;        // pinned wire record, as seen by host AND device (UVA): kHostRetLines x 16 bytes
    unsigned seq;
    void *dev;         // device slot
};
void wait_host_ret(const LaunchCtx &c, const HostRet &r);   // spins; surfaces kernel failures as af_err instead of hanging

// One shard of an array spread over several GPUs (one process per GPU): what a kernel needs to exchange its O(1) partial
// with the other ranks over peer-mapped symmetric buffers and fold the world's partials itself (kernels/peer_comm.cuh).
struct ShardLink {
    unsigned long long bufs[8];   // device address of every rank's symmetric buffer (>= 2048 bytes, zeroed once)
    int rank, world;
    unsigned epoch;               // 1, 2, 3, ... identical on every rank for the same op
    unsigned idx_base;            // global index of the shard's first element
    int mode;                     // reductions: 0 fold of all ranks, 1 fold of the LOWER ranks (carry-in of a sharded scan)
};

// out: dense array with dims = in.dims but dims[dim] = 1.  Returns false if (op,ti,to) is not instantiated.
bool launch_reduce(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, int dim, bool change_nan,
                   double nanval, void *out);
// out: device pointer to ONE element of `to`
// `ret` (optional): the result is ALSO published to the pinned record (value at host+0, index at host+32)
// `link` (optional): `in` is one shard; the kernel exchanges partials with the other ranks and folds them in rank order
bool launch_reduce_all(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, bool change_nan,
                       double nanval, void *out, const HostRet *ret = nullptr, const ShardLink *link = nullptr);
bool launch_ireduce(const LaunchCtx &c, Op op, af_dtype t, const View &in, int dim, void *out_val, unsigned *out_loc);
bool launch_ireduce_all(const LaunchCtx &c, Op op, af_dtype t, const View &in, void *out_val, unsigned *out_loc,
                        const HostRet *ret = nullptr, const ShardLink *link = nullptr);
// af_max_ragged: like launch_ireduce(OP_MAX) but line j only looks at its first min(len, rlen[j]) elements
bool launch_ireduce_ragged(const LaunchCtx &c, Op op, af_dtype t, const View &in, int dim, const unsigned *rlen,
                           void *out_val, unsigned *out_loc);
// carry / carry_key: optional device pointers (sharded scans); out is dense with in.dims
bool launch_scan(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, int dim, bool inclusive,
                 void *out, const void *carry);
bool launch_scan_by_key(const LaunchCtx &c, Op op, af_dtype t, af_dtype tk, const View &key, const View &in, int dim,
                        bool inclusive, void *out, const void *carry, const void *carry_key, void *summary);
// out[i] = op(*carry_host, out[i]) for i < count, in place (cross-shard fix-up of a by-key scan)
bool launch_scan_fixup(const LaunchCtx &c, Op op, af_dtype t, void *out, dim_t count, const void *carry_host);
// writes hit indices (+base) to `out` (capacity = elements) and the count to *d_count (device, 8 bytes)
// `ret` (optional): the count is also published to the pinned record (u64 at host+0)
// with `link` the kernel also exchanges the shard counts: the record then holds {count, offset of this shard's slice in the
// global output, global count} (3 x u64)
bool launch_where(const LaunchCtx &c, af_dtype t, const View &in, unsigned *out, unsigned long long *d_count,
                  unsigned base, const HostRet *ret = nullptr, const ShardLink *link = nullptr);
// sharded scan_by_key, after the carry-less local scan (summary mode): ONE kernel exchanges (tail total, has-head, first
// key, last key) with the other ranks, derives this shard's carry and repairs its first segment in place
bool launch_sbk_exchange_fixup(const LaunchCtx &c, Op op, af_dtype t, af_dtype tk, void *out, dim_t n, const void *keys,
                               const void *summary, const ShardLink &link);
// ---- statistics on the reduction path (kernels/stats_*.cu).  Output dtypes follow stat_out_type() (api_stats.cpp).
bool launch_mean(const LaunchCtx &c, af_dtype ti, const View &in, int dim, void *out);
// out: ONE element of the compute type (f32 / f64 / c32 / c64; f16 computes in f32), also published through `ret`
bool launch_mean_all(const LaunchCtx &c, af_dtype ti, const View &in, void *out, const HostRet *ret);
// variance (take_sqrt: standard deviation) = norm * sum (x - mean)^2 in one pass; out_mean (optional): the means
bool launch_var(const LaunchCtx &c, af_dtype ti, const View &in, int dim, void *out_var, void *out_mean, double norm,
                bool take_sqrt);
bool launch_var_all(const LaunchCtx &c, af_dtype ti, const View &in, void *out, double norm, bool take_sqrt,
                    const HostRet *ret);
// sum (OP_ADD) or max (OP_MAX) of |x| (pre 0), |x|^2 (pre 1) or |x|^p (pre 2); dim < 0: whole array.  Result type: f32 for
// f32 / c32 / f16 inputs, f64 for f64 / c64
bool launch_norm_reduce(const LaunchCtx &c, Op op, af_dtype ti, const View &in, int dim, int pre, double p, void *out,
                        const HostRet *ret);
// x, y: dense arrays of n elements; out: {sum x, sum y, sum x^2, sum y^2, sum xy} in f32 (f64 for f64 / s64 / u64)
bool launch_corr_sums(const LaunchCtx &c, af_dtype t, const void *x, const void *y, dim_t n, void *out);
// out[j] = scale * sum_i (x[i,j] - *mx)(y[i,j] - *my) over dense [n x ncols] arrays; mx / my device scalars
bool launch_cov_cols(const LaunchCtx &c, af_dtype t, const void *x, const void *y, dim_t n, dim_t ncols, const void *mx,
                     const void *my, double scale, void *out);
// ---- callers next to the scan / compaction path (kernels/diff.cu)
bool launch_diff(const LaunchCtx &c, af_dtype t, const View &in, int dim, int order, void *out);
bool launch_differs(const LaunchCtx &c, af_dtype t, const void *in, dim_t n, unsigned char *mask);
bool launch_gather(const LaunchCtx &c, af_dtype t, const void *in, const unsigned *idx, dim_t n, void *out);
// ---- lazily evaluated element-wise expressions (api_expr.cpp builds them, kernels/expr.cu interprets them)
enum ExprOp {
    EX_LEAF = 0, EX_CONST = 1,
    EX_ADD, EX_SUB, EX_MUL, EX_DIV, EX_LT, EX_GT, EX_LE, EX_GE, EX_EQ, EX_NEQ, EX_AND, EX_OR, EX_MIN, EX_MAX, EX_POW,
    EX_SQRT, EX_ABS, EX_NEG, EX_NOT, EX_SQR,
    EX_FIRST_UNARY = EX_SQRT,
    // flags on a binary op: one operand is the constant cst[arg] instead of a stack entry (no push / pop for scalars)
    EX_RCONST = 0x40, EX_LCONST = 0x80, EX_OPMASK = 0x3f
};
enum ExprReduce { EXR_SUM = 0, EXR_COUNT = 1, EXR_MIN = 2, EXR_MAX = 3, EXR_ANY = 4, EXR_ALL = 5 };
constexpr int kExprMaxOps = 48, kExprMaxLeaves = 8, kExprMaxConsts = 8;
struct ExprProg {                 // a postfix program over dense, equally shaped operands
    int n;                        // instructions
    int nleaf;                    // distinct operands
    int depth;                    // value-stack entries needed besides the register-resident top
    unsigned char op[kExprMaxOps], arg[kExprMaxOps];
    const void *leaf[kExprMaxLeaves];
    int leaf_type[kExprMaxLeaves];   // 0 f32, 1 f64, 2 b8 / s8, 3 s32, 4 u32, 5 s16, 6 u16, 7 u8 (3.. : double interpreter only); +16: through float
    double cst[kExprMaxConsts];
    int aligned;                  // every operand is 16-byte aligned
};
// out_type: 0 f32, 1 f64, 2 b8; dbl: interpret in double (any f64 operand) instead of float
void launch_expr_eval(const LaunchCtx &c, const ExprProg &p, bool dbl, void *out, int out_type, dim_t n);
// result (device `out`, also published through ret): float / double for sum, min, max; u32 for count; b8 for any / all
void launch_expr_reduce(const LaunchCtx &c, const ExprProg &p, bool dbl, int mode, dim_t n, void *out, const HostRet *ret);
// ---- reductions by key (kernels/rbk_*.cu): keys = n dense 32-bit keys, runs of equal consecutive keys are segments
// pass 1: number of runs -> published through `ret` (u64); tile_base[t] (device, ntiles + 1 words) = runs starting before tile t
dim_t rbk_tile_elems();
dim_t rbk_scratch_elems(dim_t ntiles);   // u32 words of the pass-1 scratch: tile bases + one run-head bit per key
void launch_rbk_count(const LaunchCtx &c, const void *keys, dim_t n, unsigned *tile_base, const HostRet &ret);
// pass 2: vals reduced per run along `dim` into vals_out (dense, dims[dim] = n_reduced); keys_out (may be null) gets one key per run
bool launch_rbk_reduce(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const void *keys, dim_t n, const View &vals,
                       int dim, bool change_nan, double nanval, const unsigned *tile_base, void *keys_out, void *vals_out,
                       dim_t n_reduced);
// all-gather of one 64-byte record per rank over peer-mapped symmetric buffers; the world x 64 bytes land in pinned
// host memory at host_out and *host_flag becomes `epoch` (kernels/peer.cu)
void launch_peer_exchange(const LaunchCtx &c, const void *payload, const unsigned long long *bufs, int rank, int world,
                          unsigned epoch, void *host_out, unsigned *host_flag);
// dense copy of a strided view (same dtype), value fill, dtype cast (dense -> dense)
void launch_copy_strided(const LaunchCtx &c, af_dtype t, const View &in, void *out);
void launch_tile(const LaunchCtx &c, af_dtype t, const View &in, const dim_t odims[4], void *out);   // af_tile
void launch_fill(const LaunchCtx &c, af_dtype t, void *out, dim_t n, double re, double im);
bool launch_cast(const LaunchCtx &c, af_dtype ti, af_dtype to, const View &in, void *out);

// ---------------------------------------------------------------------------------------------
// per-device persistent scratch
// ---------------------------------------------------------------------------------------------
struct Scratch {
    int device = -1;
    void *partials = nullptr;      size_t partials_bytes = 0;   // block partials of split reductions
    unsigned *tickets = nullptr;   size_t tickets_count = 0;    // self-resetting arrival counters (zeroed once)
    void *tiles = nullptr;         size_t tiles_bytes = 0;      // decoupled-lookback tile descriptors (scans)
    void *wtiles = nullptr;        size_t wtiles_bytes = 0;     // packed single-word descriptors (where)
    unsigned *tile_counter = nullptr;                           // dynamic tile-id ticket (self-resetting)
    unsigned long long *chain_ticket = nullptr;                 // monotone 64-bit ticket of the persistent chained kernels
    unsigned long long chain_base = 0;                          // host mirror: its value once all enqueued launches are done
    unsigned epoch = 0, wepoch = 0;                             // generation tags of the two descriptor arrays
    void *host_pinned = nullptr;   size_t host_pinned_bytes = 0;  // pinned: [0,4096) peer exchange, [4096,12288) HostRet wire ring
    void *dev_scalar = nullptr;                                 // 64 device slots of 64 bytes (paired with the HostRet ring)
    unsigned ret_seq = 0;                                       // last HostRet sequence number handed out
    unsigned char ret_plain[64][64] = {};                       // unpacked payloads of the HostRet ring
    HostRet next_host_ret();                                    // call under the device's op mutex

    void *need_partials(size_t bytes);
    unsigned *need_tickets(size_t count);
    void ensure_basics();                            // tile_counter, dev_scalar, host_pinned
    void *need_tiles(size_t bytes, cudaStream_t s);  // zero-filled on (re)allocation
    unsigned next_epoch(cudaStream_t s);             // bumps epoch; re-zeroes `tiles` on wrap-around
    void *need_wtiles(size_t bytes, cudaStream_t s);
    unsigned next_wepoch(cudaStream_t s);
};

}  // namespace afb
