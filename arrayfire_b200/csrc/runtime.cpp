// runtime.cpp -- error model, device/stream manager, caching memory pool and
// array-handle helpers of libafb200.
//
// Behavioural references (paths under /root/reference):
//   thread-local error string fetched+cleared by af_get_last_error   src/api/c/error.cpp:19-37,
//                                                                     src/backend/common/err_common.cpp:205-208
//   thread-local active device, one non-default stream per device    src/backend/cuda/platform.cpp:354-408
//   pool policy: round up to a step (1024 B), exact-size free lists,  src/backend/common/DefaultMemoryManager.cpp:76,158-282
//   garbage-collect and retry once on out-of-memory
// The implementation is this library's own: a mutex-protected size-bucketed pool per
// device plus a persistent scratch arena so that hot ops allocate nothing per call.
#include <nvtx3/nvToolsExt.h>

#include "common.hpp"

#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <map>
#include <unordered_map>
#include <vector>

namespace afb {

// ------------------------------------------------------------------ errors
static thread_local std::string tls_error;

void fail(af_err code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Error(code, buf);
}

af_err store_error(af_err code, const char *msg) noexcept {
    try {
        tls_error = msg ? msg : "";
        const char *pe = getenv("AF_PRINT_ERRORS");
        if (pe && pe[0] != '0') fprintf(stderr, "afb200 error %d: %s\n", (int)code, tls_error.c_str());
    } catch (...) {}
    return code;
}

af_err current_exception_to_err() noexcept {
    try {
        throw;
    } catch (const Error &e) {
        return store_error(e.code, e.what());
    } catch (const std::bad_alloc &) {
        return store_error(AF_ERR_NO_MEM, "host allocation failed");
    } catch (const std::exception &e) {
        return store_error(AF_ERR_INTERNAL, e.what());
    } catch (...) {
        return store_error(AF_ERR_UNKNOWN, "unknown exception");
    }
}

std::string take_last_error() {
    std::string s;
    s.swap(tls_error);
    return s;
}

// ------------------------------------------------------------------ devices
namespace {

struct Pool {
    std::mutex mu;
    std::multimap<size_t, void *> free_list;                 // size -> cached pointer
    std::unordered_map<void *, size_t> live;                 // every pointer we cudaMalloc'ed and handed out
    std::unordered_map<void *, size_t> raw_user;             // af_alloc_device pointers (user-locked)
    size_t cached_bytes = 0, live_bytes = 0;
};

struct Device {
    int id = -1;
    cudaStream_t stream = nullptr;
    // dedicated copy streams: an upload / a download blocks its caller but not the other direction nor the compute stream
    // (two host threads can keep both PCIe directions busy); ordered against the compute stream by events
    cudaStream_t up_stream = nullptr, down_stream = nullptr;
    cudaEvent_t up_ev = nullptr, down_ev = nullptr;
    std::mutex up_mu, down_mu;
    int sms = 0;
    Scratch scratch;
    Pool pool;
    std::mutex op_mu;
    std::once_flag once;
    std::once_flag scratch_once;
};

int g_ndev = -1;
std::once_flag g_once;
Device *g_devs = nullptr;
std::atomic<size_t> g_step{1024};
std::atomic<unsigned long long> g_launches{0};
bool g_sync_calls = false, g_mem_debug = false, g_trace = false;

void init_globals() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        n = 0;
        cudaGetLastError();
    }
    g_ndev = n;
    g_devs = new Device[n > 0 ? n : 1];
    const char *s = getenv("AF_SYNCHRONOUS_CALLS");
    g_sync_calls = s && s[0] == '1';
    const char *m = getenv("AF_MEM_DEBUG");
    g_mem_debug = m && m[0] == '1';
    const char *t = getenv("AF_TRACE");
    g_trace = t && (strstr(t, "all") || strstr(t, "kernel") || strstr(t, "mem"));
}

Device &dev(int d) {
    std::call_once(g_once, init_globals);
    if (g_ndev <= 0) fail(AF_ERR_DRIVER, "no CUDA device available: libafb200 has no CPU fallback");
    if (d < 0 || d >= g_ndev) fail(AF_ERR_ARG, "invalid device id %d (have %d)", d, g_ndev);
    Device &D = g_devs[d];
    std::call_once(D.once, [&] {
        int prev = 0;
        cudaGetDevice(&prev);
        AFB_CUDA(cudaSetDevice(d));
        D.id = d;
        AFB_CUDA(cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking));
        AFB_CUDA(cudaStreamCreateWithFlags(&D.up_stream, cudaStreamNonBlocking));
        AFB_CUDA(cudaStreamCreateWithFlags(&D.down_stream, cudaStreamNonBlocking));
        AFB_CUDA(cudaEventCreateWithFlags(&D.up_ev, cudaEventDisableTiming));
        AFB_CUDA(cudaEventCreateWithFlags(&D.down_ev, cudaEventDisableTiming));
        AFB_CUDA(cudaDeviceGetAttribute(&D.sms, cudaDevAttrMultiProcessorCount, d));
        D.scratch.device = d;
        cudaSetDevice(prev);
    });
    return D;
}

thread_local int tls_device = -1;


}  // namespace

NvtxScope::NvtxScope(const char *name) { nvtxRangePushA(name); }
NvtxScope::~NvtxScope() { nvtxRangePop(); }

int device_count() {
    std::call_once(g_once, init_globals);
    return g_ndev < 0 ? 0 : g_ndev;
}

int active_device() {
    if (tls_device < 0) {
        int d = 0;
        const char *e = getenv("AF_CUDA_DEFAULT_DEVICE");
        if (e) d = atoi(e);
        if (device_count() > 0 && (d < 0 || d >= device_count())) d = 0;
        set_active_device(d);
    }
    return tls_device;
}

void set_active_device(int d) {
    dev(d);
    AFB_CUDA(cudaSetDevice(d));
    tls_device = d;
}

cudaStream_t stream_of(int device) { return dev(device).stream; }

// host -> device, synchronous for the caller (the host buffer may be reused on return).  The copy runs on the upload stream
// after everything already enqueued on the compute stream (the destination may be a pool buffer an earlier kernel still
// reads); kernels enqueued after the return see the data because the host has waited for the copy.
void upload_sync(int device, void *dst, const void *src, size_t bytes) {
    Device &D = dev(device);
    std::lock_guard<std::mutex> lk(D.up_mu);
    AFB_CUDA(cudaEventRecord(D.up_ev, D.stream));
    AFB_CUDA(cudaStreamWaitEvent(D.up_stream, D.up_ev, 0));
    AFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, D.up_stream));
    AFB_CUDA(cudaStreamSynchronize(D.up_stream));
}
// device -> host, synchronous for the caller: waits (on the download stream) for the compute work enqueued so far
void download_sync(int device, void *dst, const void *src, size_t bytes) {
    Device &D = dev(device);
    std::lock_guard<std::mutex> lk(D.down_mu);
    AFB_CUDA(cudaEventRecord(D.down_ev, D.stream));
    AFB_CUDA(cudaStreamWaitEvent(D.down_stream, D.down_ev, 0));
    AFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, D.down_stream));
    AFB_CUDA(cudaStreamSynchronize(D.down_stream));
}
int sm_count(int device) { return dev(device).sms; }
Scratch &scratch_of(int device) { return dev(device).scratch; }
std::mutex &op_mutex(int device) { return dev(device).op_mu; }
void note_launches(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
unsigned long long launches_total(bool reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

void post_launch_check(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fail(AF_ERR_INTERNAL, "kernel launch failed (%s): %s", what, cudaGetErrorString(e));
    if (g_trace) fprintf(stderr, "[afb200 kernel] %s\n", what);
    if (g_sync_calls) {
        e = cudaStreamSynchronize(stream_of(active_device()));
        if (e != cudaSuccess) fail(AF_ERR_INTERNAL, "kernel failed (%s): %s", what, cudaGetErrorString(e));
    }
}

LaunchCtx ctx_for_active_device() {
    int d = active_device();
    Device &D = dev(d);
    AFB_CUDA(cudaSetDevice(d));
    std::call_once(D.scratch_once, [&] { D.scratch.ensure_basics(); });   // callers may race here without the op mutex
    return LaunchCtx{d, D.stream, D.sms, &D.scratch};
}

// ------------------------------------------------------------------ pool
static size_t round_up(size_t bytes) {
    size_t step = g_step.load();
    if (bytes == 0) bytes = 1;
    return (bytes + step - 1) / step * step;
}

// The reference's garbage-collection thresholds (common/DefaultMemoryManager.cpp:117-126, 88-89, 175-187): before an
// allocation, free the cached buffers when the bytes in use reach max_bytes = max(0.75 * memsize, memsize - 1 GB), or when
// the number of buffers (in use + cached) reaches AF_MAX_BUFFERS (default 1000).
static size_t pool_max_bytes(Device &D) {
    static size_t cached[64] = {0};
    size_t &m = cached[D.id & 63];
    if (!m) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || total_b == 0) {
            cudaGetLastError();
            m = (size_t)1 << 30;
        } else {
            const double a = (double)total_b * 0.75, b = (double)total_b - (double)((size_t)1 << 30);
            m = (size_t)(a > b ? a : b);
        }
    }
    return m;
}
static size_t pool_max_buffers() {
    static size_t n = 0;
    if (!n) {
        const char *e = getenv("AF_MAX_BUFFERS");
        long v = e ? atol(e) : 1000;
        n = (size_t)(v < 1 ? 1 : v);
    }
    return n;
}

static void *pool_take(Device &D, size_t rounded) {
    Pool &P = D.pool;
    bool run_gc = false;
    {
        std::lock_guard<std::mutex> lk(P.mu);
        run_gc = !P.free_list.empty() &&
                 (P.live_bytes >= pool_max_bytes(D) || P.live.size() + P.free_list.size() >= pool_max_buffers());
    }
    if (run_gc) pool_gc(D.id);
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.free_list.find(rounded);
        if (it != P.free_list.end()) {
            void *p = it->second;
            P.free_list.erase(it);
            P.cached_bytes -= rounded;
            P.live[p] = rounded;
            P.live_bytes += rounded;
            return p;
        }
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, rounded);
    if (e != cudaSuccess) {  // garbage-collect the free lists and retry once
        cudaGetLastError();
        pool_gc(D.id);
        e = cudaMalloc(&p, rounded);
        if (e != cudaSuccess) {
            cudaGetLastError();
            fail(AF_ERR_NO_MEM, "device %d out of memory allocating %zu bytes", D.id, rounded);
        }
    }
    std::lock_guard<std::mutex> lk(P.mu);
    P.live[p] = rounded;
    P.live_bytes += rounded;
    return p;
}

static void pool_give(int device, void *p) {
    Device &D = g_devs[device];
    Pool &P = D.pool;
    size_t sz = 0;
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.live.find(p);
        if (it == P.live.end()) return;
        sz = it->second;
        P.live.erase(it);
        P.live_bytes -= sz;
        if (!g_mem_debug) {
            P.free_list.emplace(sz, p);
            P.cached_bytes += sz;
            return;
        }
    }
    // AF_MEM_DEBUG=1: no caching.  The stream may still be using the block.
    cudaStreamSynchronize(D.stream);
    cudaFree(p);
}

Buffer::~Buffer() {
    if (!ptr || user_locked) return;  // a user-locked buffer is owned by the user from now on
    if (pooled) {
        pool_give(device, ptr);
    } else {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(device);
        cudaStreamSynchronize(g_devs[device].stream);
        cudaFree(ptr);
        cudaSetDevice(prev);
    }
}

std::shared_ptr<Buffer> pool_alloc(size_t bytes) {
    int d = active_device();
    Device &D = dev(d);
    AFB_CUDA(cudaSetDevice(d));
    auto b = std::make_shared<Buffer>();
    b->bytes = round_up(bytes);
    b->device = d;
    b->ptr = pool_take(D, b->bytes);
    b->pooled = true;
    return b;
}

std::shared_ptr<Buffer> adopt_device_ptr(void *ptr, size_t bytes) {
    auto b = std::make_shared<Buffer>();
    b->ptr = ptr;
    b->bytes = bytes;
    b->device = active_device();
    // a pointer that came from af_alloc_device goes back to the pool; a foreign one is cudaFree'd
    Pool &P = dev(b->device).pool;
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.raw_user.find(ptr);
    if (it != P.raw_user.end()) {
        P.raw_user.erase(it);
        b->pooled = true;
    } else {
        b->pooled = false;
    }
    return b;
}

void *pool_alloc_raw(size_t bytes) {
    int d = active_device();
    Device &D = dev(d);
    AFB_CUDA(cudaSetDevice(d));
    size_t r = round_up(bytes);
    void *p = pool_take(D, r);
    std::lock_guard<std::mutex> lk(D.pool.mu);
    D.pool.raw_user[p] = r;
    return p;
}

void pool_free_raw(void *ptr) {
    if (!ptr) return;
    int d = active_device();
    Device &D = dev(d);
    {
        std::lock_guard<std::mutex> lk(D.pool.mu);
        D.pool.raw_user.erase(ptr);
    }
    pool_give(d, ptr);
}

void pool_gc(int device) {
    Device &D = dev(device);
    std::vector<void *> victims;
    {
        std::lock_guard<std::mutex> lk(D.pool.mu);
        for (auto &kv : D.pool.free_list) victims.push_back(kv.second);
        D.pool.free_list.clear();
        D.pool.cached_bytes = 0;
    }
    if (victims.empty()) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    cudaStreamSynchronize(D.stream);
    for (void *p : victims) cudaFree(p);
    cudaSetDevice(prev);
}

void pool_stats(int device, size_t *ab, size_t *an, size_t *lb, size_t *ln) {
    Device &D = dev(device);
    std::lock_guard<std::mutex> lk(D.pool.mu);
    if (ab) *ab = D.pool.live_bytes + D.pool.cached_bytes;
    if (an) *an = D.pool.live.size() + D.pool.free_list.size();
    if (lb) *lb = D.pool.live_bytes;
    if (ln) *ln = D.pool.live.size();
}

void pool_set_step(size_t step) { g_step.store(step ? step : 1024); }
size_t pool_get_step() { return g_step.load(); }

// ------------------------------------------------------------------ scratch
void *Scratch::need_partials(size_t bytes) {
    if (bytes > partials_bytes) {
        // the stream may still read the old block: free it stream-ordered
        if (partials) { AFB_CUDA(cudaStreamSynchronize(stream_of(device))); cudaFree(partials); }
        size_t nb = bytes < (1u << 20) ? (1u << 20) : bytes * 2;
        AFB_CUDA(cudaMalloc(&partials, nb));
        partials_bytes = nb;
    }
    return partials;
}

unsigned *Scratch::need_tickets(size_t count) {
    if (count > tickets_count) {
        cudaStream_t s = stream_of(device);
        if (tickets) { AFB_CUDA(cudaStreamSynchronize(s)); cudaFree(tickets); }
        size_t n = count < 65536 ? 65536 : count * 2;
        AFB_CUDA(cudaMalloc((void **)&tickets, n * sizeof(unsigned)));
        AFB_CUDA(cudaMemsetAsync(tickets, 0, n * sizeof(unsigned), s));
        tickets_count = n;
    }
    return tickets;
}

void Scratch::ensure_basics() {
    if (tile_counter) return;
    cudaStream_t s = stream_of(device);
    AFB_CUDA(cudaMalloc((void **)&tile_counter, 256));
    AFB_CUDA(cudaMemsetAsync(tile_counter, 0, 256, s));
    chain_ticket = reinterpret_cast<unsigned long long *>(tile_counter + 32);  // same 256-byte block, own 128-byte line
    chain_base = 0;
    AFB_CUDA(cudaMalloc(&dev_scalar, 4096));
    AFB_CUDA(cudaMemsetAsync(dev_scalar, 0, 4096, s));
    AFB_CUDA(cudaMallocHost(&host_pinned, 12288));
    memset(host_pinned, 0, 12288);
    host_pinned_bytes = 12288;
}

constexpr unsigned kRetSlots = 64;

HostRet Scratch::next_host_ret() {
    if (++ret_seq == 0) ++ret_seq;
    const unsigned slot = ret_seq % kRetSlots;
    HostRet r;
    r.host = ret_plain[slot];
    r.wire = (unsigned char *)host_pinned + 4096 + (size_t)slot * 128;
    r.seq = ret_seq;
    r.dev = (unsigned char *)dev_scalar + (size_t)slot * 64;
    return r;
}

void wait_host_ret(const LaunchCtx &c, const HostRet &r) {
    // every line is one 16-byte device store {data, seq, data, seq}; a 64-bit half that shows the right flag carries valid
    // data (8-byte reads of aligned pinned memory are atomic), so the two halves are polled one after the other
    const volatile unsigned long long *wire = (const volatile unsigned long long *)r.wire;
    unsigned *out = (unsigned *)r.host;
    unsigned long long spins = 0;
    for (int h = 0; h < 2 * kHostRetLines; h++) {
        unsigned long long v;
        while ((unsigned)((v = wire[h]) >> 32) != r.seq) {
            if ((++spins & 0x3ffff) == 0) {
                cudaError_t e = cudaStreamQuery(c.stream);
                if (e != cudaSuccess && e != cudaErrorNotReady) AFB_CUDA(e);
                if (e == cudaSuccess && (unsigned)(wire[h] >> 32) != r.seq)
                    fail(AF_ERR_INTERNAL, "a kernel finished without publishing its host result (seq %u, flag %u)", r.seq,
                         (unsigned)(wire[h] >> 32));
            }
        }
        out[h] = (unsigned)v;
    }
    std::atomic_thread_fence(std::memory_order_acquire);
}

void *Scratch::need_wtiles(size_t bytes, cudaStream_t s) {
    bytes += 64;   // the walkers read descriptors in 16-byte pairs and may touch the word after the last tile
    if (bytes > wtiles_bytes) {
        if (wtiles) { AFB_CUDA(cudaStreamSynchronize(s)); cudaFree(wtiles); }
        size_t nb = bytes < (1u << 20) ? (1u << 20) : bytes * 2;
        AFB_CUDA(cudaMalloc(&wtiles, nb));
        AFB_CUDA(cudaMemsetAsync(wtiles, 0, nb, s));
        wtiles_bytes = nb;
        wepoch = 0;
    }
    return wtiles;
}

unsigned Scratch::next_wepoch(cudaStream_t s) {
    if (++wepoch >= (1u << 29) - 1) {  // 29 bits: the packed descriptors keep a flag bit below the generation
        AFB_CUDA(cudaMemsetAsync(wtiles, 0, wtiles_bytes, s));
        wepoch = 1;
    }
    return wepoch;
}

void *Scratch::need_tiles(size_t bytes, cudaStream_t s) {
    if (bytes > tiles_bytes) {
        if (tiles) { AFB_CUDA(cudaStreamSynchronize(s)); cudaFree(tiles); }
        size_t nb = bytes < (4u << 20) ? (4u << 20) : bytes * 2;
        AFB_CUDA(cudaMalloc(&tiles, nb));
        AFB_CUDA(cudaMemsetAsync(tiles, 0, nb, s));
        tiles_bytes = nb;
        epoch = 0;
    }
    return tiles;
}

unsigned Scratch::next_epoch(cudaStream_t s) {
    // descriptor status words hold (epoch << 2 | state); 30 bits of epoch
    if (++epoch >= (1u << 30) - 1) {
        AFB_CUDA(cudaMemsetAsync(tiles, 0, tiles_bytes, s));
        epoch = 1;
    }
    return epoch;
}

// ------------------------------------------------------------------ array handle
int Array::ndims() const {
    dim_t n = elements();
    if (n == 0) return 0;
    if (n == 1) return 1;
    if (info.dims[3] != 1) return 4;
    if (info.dims[2] != 1) return 3;
    if (info.dims[1] != 1) return 2;
    return 1;
}

bool Array::is_linear() const {
    int nd = ndims();
    if (nd == 1) return info.strides[0] == 1;
    dim_t count = 1;
    for (int i = 0; i < nd; i++) {
        if (count != info.strides[i]) return false;
        count *= info.dims[i];
    }
    return true;
}

void dense_strides(const dim_t dims[4], dim_t strides[4]) {
    strides[0] = 1;
    for (int i = 1; i < 4; i++) strides[i] = strides[i - 1] * dims[i - 1];
}

Array *peek_handle(af_array h) {
    if (!h) fail(AF_ERR_ARG, "null af_array handle");
    Array *a = reinterpret_cast<Array *>(h);
    if (a->magic != kMagic) fail(AF_ERR_INVALID_ARRAY, "af_array handle was not created by libafb200");
    return a;
}

Array *from_handle(af_array h) {
    Array *a = peek_handle(h);
    if (a->expr) materialize(*a);
    return a;
}

static Array *blank(af_dtype t, const dim_t dims[4], int device) {
    Array *a = new Array();
    a->info.devId = (unsigned)device | ((unsigned)AF_BACKEND_CUDA << 8);
    a->info.type = t;
    for (int i = 0; i < 4; i++) a->info.dims[i] = dims[i];
    a->info.offset = 0;
    dense_strides(dims, a->info.strides);
    a->info.is_sparse = false;
    return a;
}

Array *new_array(af_dtype t, const dim_t dims[4]) {
    for (int i = 0; i < 4; i++)
        if (dims[i] < 0) fail(AF_ERR_SIZE, "negative dimension");
    Array *a = blank(t, dims, active_device());
    try {
        a->buf = pool_alloc((size_t)a->elements() * dtype_size(t));
    } catch (...) {
        delete a;
        throw;
    }
    return a;
}

Array *new_view(const Array &parent, const dim_t dims[4], dim_t offset, const dim_t strides[4]) {
    Array *a = blank(parent.info.type, dims, parent.device());
    a->info.offset = offset;
    for (int i = 0; i < 4; i++) a->info.strides[i] = strides[i];
    a->buf = parent.buf;
    return a;
}

Array *retain(const Array &p) {
    Array *a = new Array();
    a->info = p.info;
    a->buf = p.buf;
    a->expr = p.expr;
    return a;
}

// An array used on a device other than the one it lives on is MOVED to the active device, like the reference's
// checkAndMigrate (/root/reference/src/backend/cuda/Array.cpp:61-74): a fresh buffer from the active device's pool, one
// peer copy ordered after the work already enqueued on the source device's stream, and the handle now belongs to the
// active device (other handles / views of the old buffer keep it).  The caller blocks until the copy has landed: the old
// buffer may return to its pool the moment this handle lets go of it.
void check_same_device(const Array &ca) {
    const int act = active_device();
    if (ca.device() == act) return;
    Array &a = const_cast<Array &>(ca);
    if (a.expr)
        fail(AF_ERR_DEVICE, "an unevaluated expression built on device %d cannot be consumed on device %d", a.device(), act);
    const int src = a.device();
    if (a.buf && a.buf->ptr && a.buf->bytes) {
        Device &S = dev(src), &D = dev(act);
        std::shared_ptr<Buffer> moved = pool_alloc(a.buf->bytes);   // on the active device
        cudaEvent_t ev = nullptr;
        AFB_CUDA(cudaSetDevice(src));
        AFB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e = cudaEventRecord(ev, S.stream);
        cudaSetDevice(act);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(D.stream, ev, 0);
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(moved->ptr, act, a.buf->ptr, src, a.buf->bytes, D.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
        cudaEventDestroy(ev);
        if (e != cudaSuccess) fail(AF_ERR_DEVICE, "moving an array from device %d to device %d failed: %s", src, act, cudaGetErrorString(e));
        a.buf = moved;
        if (g_trace) fprintf(stderr, "[afb200] array moved from device %d to device %d (%zu bytes)\n", src, act, moved->bytes);
    }
    a.info.devId = (a.info.devId & ~0xffu) | (unsigned)act;
}

View view_of(const Array &a) {
    View v;
    v.ptr = a.elements() ? a.data() : nullptr;
    for (int i = 0; i < 4; i++) {
        v.dims[i] = a.info.dims[i];
        v.strides[i] = a.info.strides[i];
    }
    return v;
}

}  // namespace afb
