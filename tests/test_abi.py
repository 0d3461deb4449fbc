"""C-ABI checks that need no GPU: libafb200.so loads and exports every symbol include/afb200.h declares,
argument validation that happens before any device work returns the reference's error codes, and the
product never links or loads anything under oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "afb200.h")
LIB = os.path.join(ROOT, "arrayfire_b200", "libafb200.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(LIB)


def declared_symbols():
    text = open(HDR).read()
    return sorted(set(re.findall(r"AFAPI\s+[\w\s\*]+?\b(af[xcu]*_\w+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for s in ("af_sum", "af_sum_nan", "af_product", "af_min", "af_max", "af_count", "af_all_true", "af_any_true",
              "af_sum_all", "af_min_all", "af_max_all", "af_count_all", "af_sum_all_array", "af_imin", "af_imax",
              "af_imin_all", "af_imax_all", "af_accum", "af_scan", "af_scan_by_key", "af_where", "af_create_array",
              "af_release_array", "af_get_data_ptr", "af_index", "af_get_last_error", "afcu_get_stream"):
        assert s in syms, s
    assert len(syms) > 100


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_no_extra_dependencies_on_the_oracle(lib):
    out = subprocess.check_output(["ldd", LIB]).decode()
    assert "oracle" not in out and "afref" not in out
    for src in os.listdir(os.path.join(ROOT, "arrayfire_b200", "csrc")):
        p = os.path.join(ROOT, "arrayfire_b200", "csrc", src)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read().replace("parity oracle", ""), src
    assert "oracle" not in open(os.path.join(ROOT, "arrayfire_b200", "__init__.py")).read()


def test_error_codes_without_a_device(lib):
    """null handles / bad args are rejected before any CUDA call, with the reference's af_err values"""
    lib.af_err_to_string.restype = ctypes.c_char_p
    out = ctypes.c_void_p()
    assert lib.af_sum(ctypes.byref(out), None, 0) == 202          # AF_ERR_ARG: null array
    assert lib.af_sum(ctypes.byref(out), None, 7) == 202          # dim out of range
    assert lib.af_release_array(None) == 0                        # releasing a null handle is fine
    msg, n = ctypes.c_char_p(), ctypes.c_longlong()
    lib.af_sum(ctypes.byref(out), None, 9)
    lib.af_get_last_error.restype = None
    lib.af_get_last_error(ctypes.byref(msg), ctypes.byref(n))
    assert n.value > 0 and b"invalid argument" in msg.value
    lib.af_get_last_error(ctypes.byref(msg), ctypes.byref(n))     # fetched + cleared
    assert n.value == 0
    assert lib.af_err_to_string(204) == b"Function does not support this data type"
    # the entry points beyond the north-star path validate their arguments the same way (AF_ERR_ARG before any device work)
    m = ctypes.c_void_p()
    assert lib.af_tile(ctypes.byref(out), None, 1, 1, 1, 1) == 202 and lib.af_tile(None, None, 1, 1, 1, 1) == 202
    assert lib.af_var_weighted(ctypes.byref(out), None, None, ctypes.c_longlong(4)) == 202      # var.cpp:267 dim in 0..3
    assert lib.af_var_weighted(None, None, None, ctypes.c_longlong(0)) == 202
    assert lib.af_mean_weighted(ctypes.byref(out), None, None, ctypes.c_longlong(-1)) == 202    # mean.cpp:98
    assert lib.af_meanvar(ctypes.byref(m), ctypes.byref(out), None, None, 2, ctypes.c_longlong(0)) == 202
    assert lib.af_eq(ctypes.byref(out), None, None, False) == 202
    maj, mi, pa = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.af_get_version(ctypes.byref(maj), ctypes.byref(mi), ctypes.byref(pa)) == 0 and (maj.value, mi.value) == (3, 10)


def test_no_cpu_fallback(lib):
    """without a usable CUDA device the product fails loudly (AF_ERR_DRIVER), it never computes on the CPU"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    arr = ctypes.c_void_p()
    dims = (ctypes.c_longlong * 1)(4)
    data = (ctypes.c_float * 4)(1, 2, 3, 4)
    assert lib.af_create_array(ctypes.byref(arr), data, 1, dims, 0) == 102  # AF_ERR_DRIVER
    assert lib.af_init() == 102


def _prototypes(text):
    """{name: (return type, [parameter types])} of every `AFAPI <ret> name(args);`, parameter names and `const` dropped"""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for m in re.finditer(r"AFAPI\s+(\w+)\s+(\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        parts = [p.strip() for p in m.group(3).split(",")]
        if parts in ([""], ["void"]):
            parts = []
        types = []
        for p in parts:
            p = re.sub(r"\s+", " ", re.sub(r"\bconst\b", " ", p)).strip()
            m2 = re.match(r"^(.*?)(\w+)$", p)
            if m2 and m2.group(1).strip():
                p = m2.group(1)
            types.append(p.replace(" ", ""))
        out[m.group(2)] = (m.group(1), types)
    return out


def test_prototypes_match_the_reference_headers():
    """the drop-in claim, checked textually: every af_* / afcu_* prototype of include/afb200.h has the return type and the
    parameter type list of the same function in the reference's include/af/*.h (needs /root/reference: CPU box only)"""
    import glob
    ref_dir = "/root/reference/include/af"
    if not os.path.isdir(ref_dir):
        pytest.skip("/root/reference is not present")
    ref = {}
    for f in glob.glob(os.path.join(ref_dir, "*.h")):
        ref.update(_prototypes(open(f).read()))
    ours = _prototypes(open(HDR).read())
    checked = 0
    for name, sig in sorted(ours.items()):
        if name.startswith("afx_"):
            continue                      # this library's own extensions (sharded / issue-collect entry points)
        assert name in ref, "%s is not a function of the reference's C API" % name
        if name == "afcu_get_stream":     # cudaStream_t is a pointer typedef: void ** is the same ABI without <cuda_runtime.h>
            assert sig == ("af_err", ["void**", "int"]) and ref[name] == ("af_err", ["cudaStream_t*", "int"])
        else:
            assert sig == ref[name], (name, sig, ref[name])
        checked += 1
    assert checked >= 140


def _enum_values(text):
    """{enumerator: value} of every `typedef enum { ... }` of a header (C rules: previous + 1 unless assigned)"""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for body in re.findall(r"typedef\s+enum\s*\w*\s*\{(.*?)\}", text, flags=re.S):
        val = -1
        for item in body.split(","):
            item = re.sub(r"#(if|endif|else)[^\n]*\n?", "", item.strip() + "\n").strip()
            if not item:
                continue
            m = re.match(r"^(\w+)\s*(?:=\s*(.+))?$", item, flags=re.S)
            assert m, item
            if m.group(2) is not None:
                try:
                    val = int(m.group(2).strip(), 0)
                except ValueError:
                    val = eval(m.group(2).strip(), {}, dict(out))   # e.g. AF_BACKEND_CUDA = 2 written via another enumerator
            else:
                val += 1
            out[m.group(1)] = val
    return out


def test_enum_values_match_the_reference_header():
    """af_err codes, af_dtype, af_binary_op, af_var_bias, af_norm_type, af_backend ...: same names, same values as
    /root/reference/include/af/defines.h (a caller compiled against the reference's headers passes these integers)"""
    ref_hdr = "/root/reference/include/af/defines.h"
    if not os.path.isfile(ref_hdr):
        pytest.skip("/root/reference is not present")
    ours, ref = _enum_values(open(HDR).read()), _enum_values(open(ref_hdr).read())
    assert len(ours) >= 50
    for name, value in ours.items():
        assert name in ref, name
        assert ref[name] == value, (name, value, ref[name])


def test_python_mirror_and_oracle_use_the_header_values(lib):
    """the ctypes mirror (arrayfire_b200) and the oracle's Python front (oracle/pyoracle.py) spell af_dtype / af_binary_op /
    af_var_bias / af_norm_type with the integers of include/afb200.h"""
    import arrayfire_b200 as af
    from oracle import pyoracle as po
    e = _enum_values(open(HDR).read())
    for n in ("f32", "c32", "f64", "c64", "b8", "s32", "u32", "u8", "s64", "u64", "s16", "u16", "f16", "s8"):
        assert getattr(af, n) == e[n] and getattr(po, n.upper()) == e[n], n
    for n in ("AF_BINARY_ADD", "AF_BINARY_MUL", "AF_BINARY_MIN", "AF_BINARY_MAX", "AF_VARIANCE_DEFAULT", "AF_VARIANCE_SAMPLE",
              "AF_VARIANCE_POPULATION", "AF_NORM_VECTOR_1", "AF_NORM_VECTOR_INF", "AF_NORM_VECTOR_2", "AF_NORM_VECTOR_P",
              "AF_NORM_MATRIX_1", "AF_NORM_MATRIX_INF", "AF_NORM_MATRIX_2", "AF_NORM_MATRIX_L_PQ"):
        assert getattr(af, n) == e[n] and getattr(po, n) == e[n], n
