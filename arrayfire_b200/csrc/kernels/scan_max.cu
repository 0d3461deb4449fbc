// scan_max.cu -- af_scan / af_accum instantiations for op MAX.
#include "scan_host.cuh"
namespace afb {
bool scan_max(const LaunchCtx &c, af_dtype ti, af_dtype to, const View &in, int dim, bool inclusive, void *out,
             const void *carry) {
    return scan_family<OP_MAX>(c, ti, to, in, dim, inclusive, out, carry);
}
}  // namespace afb
