// sbk_max.cu -- af_scan_by_key instantiations for op MAX.
#include "scan_host.cuh"
namespace afb {
bool sbk_max(const LaunchCtx &c, af_dtype t, af_dtype tk, const View &key, const View &in, int dim, bool inclusive,
            void *out, const void *carry, const void *carry_key, void *summary) {
    return scan_by_key_family<OP_MAX>(c, t, tk, key, in, dim, inclusive, out, carry, carry_key, summary);
}
}  // namespace afb
