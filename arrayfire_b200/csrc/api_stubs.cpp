// api_stubs.cpp -- every af_* symbol the reference's C++ wrapper sources of this path reference but that lies OUTSIDE
// the reduce / scan / where path (element-wise arithmetic, joins, tiling, gathers, I/O ...).
//
// The wrapper's array.cpp / data.cpp / index.cpp are compiled into each reference backend library whole
// (/root/reference/src/api/cpp/CMakeLists.txt:1-4), so a backend that hosts them must resolve all of their symbols even
// when a program never calls them.  Each stub fails the way SURVEY.md section 8b prescribes for a subset backend: it
// returns AF_ERR_NOT_SUPPORTED and leaves a message for af_get_last_error -- it never fabricates a result and there is
// no CPU fallback behind it.
//
// C linkage symbols carry no signature; the stubs read none of their arguments, and on the SysV x86-64 / AAPCS64
// calling conventions the caller owns argument clean-up, so one parameterless definition serves every prototype.
#include "common.hpp"

#define AFB_STUB(name)                                                                                              \
    extern "C" __attribute__((visibility("default"))) af_err name(void) {                                          \
        return ::afb::store_error(AF_ERR_NOT_SUPPORTED, #name ": not provided by libafb200 (outside the reduce / "   \
                                                              "scan / where path this backend implements)");        \
    }

// element-wise arithmetic / logic used by af::array's operators (include/af/arith.h)
AFB_STUB(af_mod)

AFB_STUB(af_bitnot)
AFB_STUB(af_bitand) AFB_STUB(af_bitor) AFB_STUB(af_bitxor) AFB_STUB(af_bitshiftl) AFB_STUB(af_bitshiftr)
AFB_STUB(af_real) AFB_STUB(af_imag) AFB_STUB(af_cplx) AFB_STUB(af_cplx2) AFB_STUB(af_conjg) 
// data movement / generation (include/af/data.h, index.h)
AFB_STUB(af_reorder) AFB_STUB(af_join) AFB_STUB(af_join_many) AFB_STUB(af_shift) AFB_STUB(af_flip)
AFB_STUB(af_lower) AFB_STUB(af_upper) AFB_STUB(af_pad) AFB_STUB(af_range) AFB_STUB(af_iota) AFB_STUB(af_identity)
AFB_STUB(af_diag_create) AFB_STUB(af_diag_extract) AFB_STUB(af_lookup) AFB_STUB(af_assign_gen) AFB_STUB(af_assign_seq)
AFB_STUB(af_select) AFB_STUB(af_select_scalar_l) AFB_STUB(af_select_scalar_r)
AFB_STUB(af_select_scalar_l_long) AFB_STUB(af_select_scalar_l_ulong)
AFB_STUB(af_select_scalar_r_long) AFB_STUB(af_select_scalar_r_ulong)
AFB_STUB(af_replace) AFB_STUB(af_replace_scalar) AFB_STUB(af_replace_scalar_long) AFB_STUB(af_replace_scalar_ulong)
AFB_STUB(af_transpose) AFB_STUB(af_transpose_inplace)
// printing / file I/O (include/af/util.h)
AFB_STUB(af_print_array_gen) AFB_STUB(af_array_to_string) AFB_STUB(af_save_array)
AFB_STUB(af_read_array_index) AFB_STUB(af_read_array_key) AFB_STUB(af_read_array_key_check)
