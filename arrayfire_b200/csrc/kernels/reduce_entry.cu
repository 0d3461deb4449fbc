// reduce_entry.cu -- launch_reduce / launch_reduce_all: route an (op, Ti, To) request to its family.
#include "reduce_dispatch.cuh"
namespace afb {
static bool route(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, int dim, bool all, bool cn,
                  double nv, void *out, const HostRet *ret, const ShardLink *link) {
    switch (op) {
        case OP_MIN: case OP_MAX: return reduce_family_minmax(c, op, ti, to, in, dim, all, cn, nv, out, ret, link);
        case OP_ADD: return reduce_family_sum(c, op, ti, to, in, dim, all, cn, nv, out, ret, link);
        case OP_MUL: return reduce_family_prod(c, op, ti, to, in, dim, all, cn, nv, out, ret, link);
        case OP_NOTZERO: case OP_AND: case OP_OR: return reduce_family_logic(c, op, ti, to, in, dim, all, cn, nv, out, ret, link);
    }
    return false;
}
bool launch_reduce(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, int dim, bool cn, double nv,
                   void *out) {
    return route(c, op, ti, to, in, dim, false, cn, nv, out, nullptr, nullptr);
}
bool launch_reduce_all(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, bool cn, double nv,
                       void *out, const HostRet *ret, const ShardLink *link) {
    return route(c, op, ti, to, in, 0, true, cn, nv, out, ret, link);
}
}  // namespace afb
