// reduce_logic.cu -- af_count (T -> u32), af_all_true / af_any_true (T -> b8) for all 14 dtypes.
#include "reduce_dispatch.cuh"
namespace afb {
bool reduce_family_logic(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, int dim, bool all,
                         bool cn, double nv, void *out, const HostRet *ret, const ShardLink *link) {
    if (op == OP_NOTZERO) {
        if (to != u32) return false;
#define X(code, T)                                                          \
    if (ti == code) {                                                       \
        reduce_one<OP_NOTZERO, T, uint>(c, in, dim, all, cn, nv, out, ret, link);      \
        return true;                                                        \
    }
        AFB_ALL_TYPES(X)
#undef X
        return false;
    }
    if (to != b8) return false;
#define X(code, T)                                                                      \
    if (ti == code) {                                                                   \
        if (op == OP_AND) reduce_one<OP_AND, T, b8_t>(c, in, dim, all, cn, nv, out, ret, link);    \
        else reduce_one<OP_OR, T, b8_t>(c, in, dim, all, cn, nv, out, ret, link);                  \
        return true;                                                                    \
    }
    AFB_ALL_TYPES(X)
#undef X
    return false;
}
}  // namespace afb
