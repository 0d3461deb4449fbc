// scan_entry.cu -- launch_scan / launch_scan_by_key routing, plus the b8 count-scan of af_accum.
#include "scan_host.cuh"
namespace afb {
bool launch_scan(const LaunchCtx &c, Op op, af_dtype ti, af_dtype to, const View &in, int dim, bool inclusive,
                 void *out, const void *carry) {
    switch (op) {
        case OP_ADD: return scan_add(c, ti, to, in, dim, inclusive, out, carry);
        case OP_MUL: return scan_mul(c, ti, to, in, dim, inclusive, out, carry);
        case OP_MIN: return scan_min(c, ti, to, in, dim, inclusive, out, carry);
        case OP_MAX: return scan_max(c, ti, to, in, dim, inclusive, out, carry);
        case OP_NOTZERO:  // af_accum on b8 counts non-zeros (/root/reference/src/api/c/scan.cpp:146-148)
            if (ti == b8 && to == u32) {
                run_scan<OP_NOTZERO, b8_t, uint, false, int>(c, in, nullptr, dim, inclusive, out, carry, nullptr, nullptr);
                return true;
            }
            return false;
        default: return false;
    }
}
bool launch_scan_by_key(const LaunchCtx &c, Op op, af_dtype t, af_dtype tk, const View &key, const View &in, int dim,
                        bool inclusive, void *out, const void *carry, const void *carry_key, void *summary) {
    switch (op) {
        case OP_ADD: return sbk_add(c, t, tk, key, in, dim, inclusive, out, carry, carry_key, summary);
        case OP_MUL: return sbk_mul(c, t, tk, key, in, dim, inclusive, out, carry, carry_key, summary);
        case OP_MIN: return sbk_min(c, t, tk, key, in, dim, inclusive, out, carry, carry_key, summary);
        case OP_MAX: return sbk_max(c, t, tk, key, in, dim, inclusive, out, carry, carry_key, summary);
        default: return false;
    }
}
}  // namespace afb
