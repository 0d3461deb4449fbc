// rbk_logic.cu -- af_count_by_key (T -> u32), af_all_true_by_key / af_any_true_by_key (T -> b8) for all 14 dtypes
// (/root/reference/src/api/c/reduce.cpp:122-154), plus pass 1 (run counting) and the dispatcher.
#include "reduce_dispatch.cuh"
#include "rbk_kernels.cuh"
namespace afb {
bool rbk_family_logic(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const void *keys, dim_t n, const View &vals, int dim,
                      bool cn, double nv, const unsigned *tile_base, void *keys_out, void *vals_out, dim_t n_reduced) {
    if (op == OP_NOTZERO) {
        if (to != u32) return false;
#define X(code, T)                                                                                               \
    if (ti == code) {                                                                                            \
        run_rbk<OP_NOTZERO, T, uint>(c, keys, n, vals, dim, cn, nv, tile_base, keys_out, vals_out, n_reduced);   \
        return true;                                                                                             \
    }
        AFB_ALL_TYPES(X)
#undef X
        return false;
    }
    if (to != b8) return false;
#define X(code, T)                                                                                               \
    if (ti == code) {                                                                                            \
        if (op == OP_AND) run_rbk<OP_AND, T, b8_t>(c, keys, n, vals, dim, cn, nv, tile_base, keys_out, vals_out, n_reduced); \
        else run_rbk<OP_OR, T, b8_t>(c, keys, n, vals, dim, cn, nv, tile_base, keys_out, vals_out, n_reduced);   \
        return true;                                                                                             \
    }
    AFB_ALL_TYPES(X)
#undef X
    return false;
}

dim_t rbk_tile_elems() { return kRbkTile; }
dim_t rbk_scratch_elems(dim_t ntiles) { return (dim_t)rbk_scratch_words(ntiles); }

void launch_rbk_count(const LaunchCtx &c, const void *keys, dim_t n, unsigned *tile_base, const HostRet &ret) {
    RbkCountArgs a;
    a.keys = (const uint *)keys;
    a.n = n;
    a.ntiles = (n + kRbkTile - 1) / kRbkTile;
    a.tile_base = tile_base;
    a.bits = reinterpret_cast<unsigned char *>(tile_base + rbk_bits_word_offset(a.ntiles));
    a.ticket = c.scratch->need_tickets(2) + 1;   // ticket 0 belongs to the whole-array reductions
    a.host_wire = ret.wire;
    a.host_seq = ret.seq;
    dim_t grid = a.ntiles;
    const dim_t cap = (dim_t)c.sms * 4;      // 58 registers: 4 resident CTAs per SM, one wave
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    rbk_count_kernel<<<(unsigned)grid, kRbkThreads, 0, c.stream>>>(a);
    note_launches(1);
    post_launch_check("rbk_count_kernel");
}

bool launch_rbk_reduce(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const void *keys, dim_t n, const View &vals,
                       int dim, bool cn, double nv, const unsigned *tile_base, void *keys_out, void *vals_out,
                       dim_t n_reduced) {
    switch (op) {
        case OP_ADD: case OP_MUL: return rbk_family_sum(c, op, ti, to, keys, n, vals, dim, cn, nv, tile_base, keys_out, vals_out, n_reduced);
        case OP_MIN: case OP_MAX: return rbk_family_minmax(c, op, ti, to, keys, n, vals, dim, cn, nv, tile_base, keys_out, vals_out, n_reduced);
        case OP_NOTZERO: case OP_AND: case OP_OR: return rbk_family_logic(c, op, ti, to, keys, n, vals, dim, cn, nv, tile_base, keys_out, vals_out, n_reduced);
    }
    return false;
}
}  // namespace afb
