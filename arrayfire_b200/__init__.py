"""arrayfire_b200 -- host-side mirror of ArrayFire's ``af::`` interface for the reduce / scan /
scan_by_key / where path, bound over the C-ABI of libafb200.so (include/afb200.h).

This module is plumbing only: every function is a thin ctypes call into the ``af_*`` symbols the
reference's C++ wrapper binds (/root/reference/src/api/cpp/reduce.cpp:17-346, scan.cpp:15-33,
where.cpp:16-24, array.cpp), with the same names, argument meaning, defaults (``dim=-1`` picks the first
non-singleton dimension, src/api/cpp/common.hpp:27-38) and error behaviour (an ``af_err`` becomes an
``AfError`` carrying the thread-local message of af_get_last_error, like af::exception).

There is NO fallback: if libafb200.so is missing or no CUDA device is usable, import / first use fails.
"""
import ctypes
import os

import numpy as np

__all__ = [
    "Array", "AfError", "lib", "f32", "c32", "f64", "c64", "b8", "s32", "u32", "u8", "s64", "u64", "s16", "u16", "f16",
    "s8", "AF_BINARY_ADD", "AF_BINARY_MUL", "AF_BINARY_MIN", "AF_BINARY_MAX", "sum", "product", "min", "max", "count",
    "all_true", "any_true", "sum_all", "product_all", "min_all", "max_all", "count_all", "all_true_all", "any_true_all",
    "imin", "imax", "imin_all", "imax_all", "accum", "scan", "scan_by_key", "where", "sync", "constant", "info_string",
    "get_stream", "launch_count", "device_gc", "seq", "span", "sum_by_key", "product_by_key", "min_by_key", "max_by_key",
    "count_by_key", "all_true_by_key", "any_true_by_key", "max_ragged", "cast", "moddims", "flat",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AFB200_LIB") or os.path.join(_HERE, "libafb200.so")   # AFB200_LIB: A-B runs of two builds (tools/)

# af_dtype, include/af/defines.h:210-233
f32, c32, f64, c64, b8, s32, u32, u8, s64, u64, s16, u16, f16, s8 = range(14)
AF_BINARY_ADD, AF_BINARY_MUL, AF_BINARY_MIN, AF_BINARY_MAX = range(4)
_NP = {f32: np.float32, c32: np.complex64, f64: np.float64, c64: np.complex128, b8: np.int8, s32: np.int32,
       u32: np.uint32, u8: np.uint8, s64: np.int64, u64: np.uint64, s16: np.int16, u16: np.uint16, f16: np.float16,
       s8: np.int8}
_SIZE = {k: np.dtype(v).itemsize for k, v in _NP.items()}
DTYPE_NAMES = ["f32", "c32", "f64", "c64", "b8", "s32", "u32", "u8", "s64", "u64", "s16", "u16", "f16", "s8"]


class AfError(RuntimeError):
    """af::exception: carries the af_err code and the library's message."""

    def __init__(self, code, message):
        super().__init__("af_err %d: %s" % (code, message))
        self.code = code
        self.message = message


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)


lib = _load()
lib.af_err_to_string.restype = ctypes.c_char_p
lib.af_get_last_error.restype = None
_vp = ctypes.c_void_p


class af_seq(ctypes.Structure):
    _fields_ = [("begin", ctypes.c_double), ("end", ctypes.c_double), ("step", ctypes.c_double)]


def seq(begin, end, step=1.0):
    """af::seq(begin, end, step) -- end is inclusive (include/af/seq.h:20-30)."""
    return af_seq(float(begin), float(end), float(step))


span = af_seq(1.0, 1.0, 0.0)  # af_span


def _check(err):
    if err != 0:
        msg = ctypes.c_char_p()
        n = ctypes.c_longlong()
        lib.af_get_last_error(ctypes.byref(msg), ctypes.byref(n))
        text = msg.value.decode(errors="replace") if msg.value else lib.af_err_to_string(err).decode()
        if msg.value:
            lib.af_free_host(msg)
        raise AfError(err, text)


def _dims_arr(dims):
    return (ctypes.c_longlong * len(dims))(*[int(d) for d in dims])


def _code_of(np_dtype, dtype=None):
    if dtype is not None:
        return dtype
    if np_dtype == np.bool_:
        return b8
    for code in (f32, c32, f64, c64, s32, u32, u8, s64, u64, s16, u16, f16, s8):
        if np_dtype == _NP[code]:
            return code
    raise TypeError("unsupported dtype %s" % np_dtype)


class Array:
    """Owner of one ``af_array`` handle (af::array, include/af/array.h:37-38)."""

    def __init__(self, handle, keepalive=None):
        self.handle = _vp(handle) if not isinstance(handle, _vp) else handle
        self._keepalive = keepalive

    # ---- construction -------------------------------------------------------------------------
    @staticmethod
    def from_numpy(a, dtype=None):
        """af::array(dims, host_ptr): copies a column-major host array to the device (af_create_array)."""
        a = np.asarray(a)
        code = _code_of(a.dtype, dtype)
        if a.ndim == 0:
            a = a.reshape(1)
        if a.ndim > 4:
            raise ValueError("at most 4 dimensions")
        host = np.asfortranarray(a.view(np.int8) if a.dtype == np.bool_ else a)
        h = _vp()
        if host.size == 0:
            _check(lib.af_create_handle(ctypes.byref(h), host.ndim, _dims_arr(host.shape), code))
        else:
            _check(lib.af_create_array(ctypes.byref(h), _vp(host.ctypes.data), host.ndim, _dims_arr(host.shape), code))
        return Array(h)

    @staticmethod
    def from_device_ptr(ptr, dims, dtype, keepalive=None):
        """af_device_array + af_lock_array: wraps foreign device memory (e.g. a torch tensor) WITHOUT copying
        and without taking ownership (the documented recipe for memory the library must not free)."""
        h = _vp()
        _check(lib.af_device_array(ctypes.byref(h), _vp(ptr), len(dims), _dims_arr(dims), dtype))
        _check(lib.af_lock_array(h))
        return Array(h, keepalive)

    @staticmethod
    def from_torch(t, dtype=None):
        """Zero-copy view of a contiguous CUDA torch tensor; dims are reversed so that torch's last
        (fastest) axis becomes dim 0, ArrayFire's fastest axis."""
        import torch
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("need a contiguous CUDA tensor")
        table = {torch.float32: f32, torch.float64: f64, torch.int32: s32, torch.int64: s64, torch.uint8: u8,
                 torch.int8: s8, torch.int16: s16, torch.float16: f16, torch.bool: b8, torch.complex64: c32,
                 torch.complex128: c64}
        for name, code in (("uint32", u32), ("uint64", u64), ("uint16", u16)):
            if hasattr(torch, name):
                table[getattr(torch, name)] = code
        code = dtype if dtype is not None else table[t.dtype]
        dims = list(reversed(t.shape)) or [1]
        return Array.from_device_ptr(t.data_ptr(), dims, code, keepalive=t)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h is not None and h.value and lib is not None:
            try:
                lib.af_release_array(h)
            except Exception:
                pass

    # ---- queries --------------------------------------------------------------------------------
    @property
    def dims(self):
        d = [ctypes.c_longlong() for _ in range(4)]
        _check(lib.af_get_dims(*[ctypes.byref(x) for x in d], self.handle))
        return tuple(x.value for x in d)

    @property
    def strides(self):
        d = [ctypes.c_longlong() for _ in range(4)]
        _check(lib.af_get_strides(*[ctypes.byref(x) for x in d], self.handle))
        return tuple(x.value for x in d)

    @property
    def offset(self):
        o = ctypes.c_longlong()
        _check(lib.af_get_offset(ctypes.byref(o), self.handle))
        return o.value

    @property
    def dtype(self):
        t = ctypes.c_int()
        _check(lib.af_get_type(ctypes.byref(t), self.handle))
        return t.value

    @property
    def elements(self):
        n = ctypes.c_longlong()
        _check(lib.af_get_elements(ctypes.byref(n), self.handle))
        return n.value

    @property
    def device(self):
        d = ctypes.c_int(0)
        _check(lib.af_get_device_id(ctypes.byref(d), self.handle))
        return d.value

    @property
    def numdims(self):
        n = ctypes.c_uint()
        _check(lib.af_get_numdims(ctypes.byref(n), self.handle))
        return n.value

    def is_linear(self):
        r = ctypes.c_bool()
        _check(lib.af_is_linear(ctypes.byref(r), self.handle))
        return r.value

    def device_ptr(self):
        """af_get_device_ptr (locks the buffer: the caller now shares ownership)."""
        p = _vp()
        _check(lib.af_get_device_ptr(ctypes.byref(p), self.handle))
        return p.value

    def unlock(self):
        _check(lib.af_unlock_array(self.handle))

    def to_numpy(self):
        """array::host(): a column-major numpy array of shape dims (synchronises)."""
        dims, code = self.dims, self.dtype
        out = np.zeros(dims, dtype=_NP[code], order="F")
        if out.size:
            _check(lib.af_get_data_ptr(_vp(out.ctypes.data), self.handle))
        return out

    def host_into(self, out):
        """array::host(ptr): copy into a caller-provided host buffer (e.g. pinned memory); synchronises."""
        n = self.elements * _SIZE[self.dtype]
        if out.nbytes < n or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("host buffer too small or not contiguous")
        if n:
            _check(lib.af_get_data_ptr(_vp(out.ctypes.data), self.handle))
        return out

    def scalar(self):
        out = np.zeros(1, dtype=_NP[self.dtype])
        _check(lib.af_get_scalar(_vp(out.ctypes.data), self.handle))
        return out[0]

    def retain(self):
        h = _vp()
        _check(lib.af_retain_array(ctypes.byref(h), self.handle))
        return Array(h, self._keepalive)

    def copy(self):
        h = _vp()
        _check(lib.af_copy_array(ctypes.byref(h), self.handle))
        return Array(h)

    def index(self, *seqs):
        """a(seq, seq, ...): a view sharing the parent's buffer (af_index, src/api/c/index.cpp:84-124)."""
        arr = (af_seq * len(seqs))(*seqs)
        h = _vp()
        _check(lib.af_index(ctypes.byref(h), self.handle, len(seqs), arr))
        return Array(h, self._keepalive)


# ---- helpers -------------------------------------------------------------------------------------
def _fnsd(a, dim):
    """getFNSD, src/api/cpp/common.hpp:27-38"""
    if dim >= 0:
        return dim
    d = a.dims
    for i in range(4):
        if d[i] > 1:
            return i
    return 0


def _dimwise(fn, a, dim, *extra):
    h = _vp()
    _check(fn(ctypes.byref(h), a.handle, ctypes.c_int(_fnsd(a, dim)), *extra))
    return Array(h)


def _all(fn, a, *extra):
    re, im = ctypes.c_double(), ctypes.c_double()
    _check(fn(ctypes.byref(re), ctypes.byref(im), a.handle, *extra))
    if a.dtype in (c32, c64):
        return complex(re.value, im.value)
    return re.value


# ---- reductions along a dimension (af::sum & co., src/api/cpp/reduce.cpp:17-110) --------------
def sum(a, dim=-1, nanval=None):
    if nanval is None:
        return _dimwise(lib.af_sum, a, dim)
    h = _vp()  # the nanval overload passes dim through unmodified (reduce.cpp:23-27)
    _check(lib.af_sum_nan(ctypes.byref(h), a.handle, ctypes.c_int(dim), ctypes.c_double(nanval)))
    return Array(h)


def product(a, dim=-1, nanval=None):
    if nanval is None:
        return _dimwise(lib.af_product, a, dim)
    h = _vp()
    _check(lib.af_product_nan(ctypes.byref(h), a.handle, ctypes.c_int(dim), ctypes.c_double(nanval)))
    return Array(h)


def min(a, dim=-1):
    return _dimwise(lib.af_min, a, dim)


def max(a, dim=-1):
    return _dimwise(lib.af_max, a, dim)


def count(a, dim=-1):
    return _dimwise(lib.af_count, a, dim)


def all_true(a, dim=-1):
    return _dimwise(lib.af_all_true, a, dim)


def any_true(a, dim=-1):
    return _dimwise(lib.af_any_true, a, dim)


# ---- whole-array reductions (af::sum<T>(a) & co., reduce.cpp:199-346) -----------------------------
def sum_all(a, nanval=None):
    if nanval is None:
        return _all(lib.af_sum_all, a)
    return _all(lib.af_sum_nan_all, a, ctypes.c_double(nanval))


def product_all(a, nanval=None):
    if nanval is None:
        return _all(lib.af_product_all, a)
    return _all(lib.af_product_nan_all, a, ctypes.c_double(nanval))


def min_all(a):
    return _all(lib.af_min_all, a)


def max_all(a):
    return _all(lib.af_max_all, a)


def count_all(a):
    return _all(lib.af_count_all, a)


def all_true_all(a):
    return _all(lib.af_all_true_all, a)


def any_true_all(a):
    return _all(lib.af_any_true_all, a)


def _all_array(fn, a, *extra):
    h = _vp()
    _check(fn(ctypes.byref(h), a.handle, *extra))
    return Array(h)


def sum_all_array(a, nanval=None):
    if nanval is None:
        return _all_array(lib.af_sum_all_array, a)
    return _all_array(lib.af_sum_nan_all_array, a, ctypes.c_double(nanval))


def product_all_array(a, nanval=None):
    if nanval is None:
        return _all_array(lib.af_product_all_array, a)
    return _all_array(lib.af_product_nan_all_array, a, ctypes.c_double(nanval))


def min_all_array(a):
    return _all_array(lib.af_min_all_array, a)


def max_all_array(a):
    return _all_array(lib.af_max_all_array, a)


def count_all_array(a):
    return _all_array(lib.af_count_all_array, a)


def all_true_all_array(a):
    return _all_array(lib.af_all_true_all_array, a)


def any_true_all_array(a):
    return _all_array(lib.af_any_true_all_array, a)


# ---- index-returning reductions (af::min/max(val, idx, in, dim), reduce.cpp:167-197) --------------
def _ireduce(fn, a, dim):
    v, i = _vp(), _vp()
    _check(fn(ctypes.byref(v), ctypes.byref(i), a.handle, ctypes.c_int(_fnsd(a, dim))))
    return Array(v), Array(i)


def imin(a, dim=-1):
    return _ireduce(lib.af_imin, a, dim)


def imax(a, dim=-1):
    return _ireduce(lib.af_imax, a, dim)


def _ireduce_all(fn, a):
    re, im, idx = ctypes.c_double(), ctypes.c_double(), ctypes.c_uint()
    _check(fn(ctypes.byref(re), ctypes.byref(im), ctypes.byref(idx), a.handle))
    val = complex(re.value, im.value) if a.dtype in (c32, c64) else re.value
    return val, idx.value


def imin_all(a):
    return _ireduce_all(lib.af_imin_all, a)


def imax_all(a):
    return _ireduce_all(lib.af_imax_all, a)


# ---- reductions by key (af::sumByKey & co., src/api/cpp/reduce.cpp:28-166) ---------------------------
def _by_key(fn, keys, vals, dim, *extra, fnsd=True):
    ko, vo = _vp(), _vp()
    d = _fnsd(vals, dim) if fnsd else dim
    _check(fn(ctypes.byref(ko), ctypes.byref(vo), keys.handle, vals.handle, ctypes.c_int(d), *extra))
    return Array(ko), Array(vo)


def sum_by_key(keys, vals, dim=-1, nanval=None):
    if nanval is None:
        return _by_key(lib.af_sum_by_key, keys, vals, dim)
    return _by_key(lib.af_sum_by_key_nan, keys, vals, dim, ctypes.c_double(nanval), fnsd=False)


def product_by_key(keys, vals, dim=-1, nanval=None):
    if nanval is None:
        return _by_key(lib.af_product_by_key, keys, vals, dim)
    return _by_key(lib.af_product_by_key_nan, keys, vals, dim, ctypes.c_double(nanval), fnsd=False)


def min_by_key(keys, vals, dim=-1):
    return _by_key(lib.af_min_by_key, keys, vals, dim)


def max_by_key(keys, vals, dim=-1):
    return _by_key(lib.af_max_by_key, keys, vals, dim)


def count_by_key(keys, vals, dim=-1):
    return _by_key(lib.af_count_by_key, keys, vals, dim)


def all_true_by_key(keys, vals, dim=-1):
    return _by_key(lib.af_all_true_by_key, keys, vals, dim)


def any_true_by_key(keys, vals, dim=-1):
    return _by_key(lib.af_any_true_by_key, keys, vals, dim)


def max_ragged(a, ragged_len, dim):
    """af::max(val, idx, in, ragged_len, dim), src/api/cpp/reduce.cpp:120-127"""
    v, i = _vp(), _vp()
    _check(lib.af_max_ragged(ctypes.byref(v), ctypes.byref(i), a.handle, ragged_len.handle, ctypes.c_int(dim)))
    return Array(v), Array(i)


def cast(a, dtype):
    """array::as(type) -> af_cast"""
    h = _vp()
    _check(lib.af_cast(ctypes.byref(h), a.handle, ctypes.c_int(dtype)))
    return Array(h)


def moddims(a, dims):
    h = _vp()
    _check(lib.af_moddims(ctypes.byref(h), a.handle, ctypes.c_uint(len(dims)), _dims_arr(dims)))
    return Array(h, a._keepalive)


def flat(a):
    h = _vp()
    _check(lib.af_flat(ctypes.byref(h), a.handle))
    return Array(h, a._keepalive)


def tile(a, x, y=1, z=1, w=1):
    """af::tile (src/api/cpp/data.cpp; api/c/tile.cpp:82-90)"""
    h = _vp()
    _check(lib.af_tile(ctypes.byref(h), a.handle, ctypes.c_uint(x), ctypes.c_uint(y), ctypes.c_uint(z), ctypes.c_uint(w)))
    return Array(h)


# ---- scans (af::accum / scan / scanByKey default dim = 0, src/api/cpp/scan.cpp:15-33) --------------
def accum(a, dim=0):
    h = _vp()
    _check(lib.af_accum(ctypes.byref(h), a.handle, ctypes.c_int(dim)))
    return Array(h)


def scan(a, dim=0, op=AF_BINARY_ADD, inclusive_scan=True):
    h = _vp()
    _check(lib.af_scan(ctypes.byref(h), a.handle, ctypes.c_int(dim), ctypes.c_int(op), ctypes.c_bool(inclusive_scan)))
    return Array(h)


def scan_by_key(key, a, dim=0, op=AF_BINARY_ADD, inclusive_scan=True):
    h = _vp()
    _check(lib.af_scan_by_key(ctypes.byref(h), key.handle, a.handle, ctypes.c_int(dim), ctypes.c_int(op),
                              ctypes.c_bool(inclusive_scan)))
    return Array(h)


def where(a):
    h = _vp()
    _check(lib.af_where(ctypes.byref(h), a.handle))
    return Array(h)


# ---- statistics on the reduction path (af::mean / var / stdev / norm / cov / corrcoef, src/api/cpp/{mean,var,stdev}.cpp) ------
AF_VARIANCE_DEFAULT, AF_VARIANCE_SAMPLE, AF_VARIANCE_POPULATION = 0, 1, 2
(AF_NORM_VECTOR_1, AF_NORM_VECTOR_INF, AF_NORM_VECTOR_2, AF_NORM_VECTOR_P, AF_NORM_MATRIX_1, AF_NORM_MATRIX_INF,
 AF_NORM_MATRIX_2, AF_NORM_MATRIX_L_PQ) = range(8)


def _stat_dim(fn, a, dim, *extra):
    h = _vp()
    _check(fn(ctypes.byref(h), a.handle, *extra, ctypes.c_longlong(_fnsd(a, dim))))
    return Array(h)


def mean(a, dim=-1):
    return _stat_dim(lib.af_mean, a, dim)


def mean_weighted(a, weights, dim=-1):
    """af::mean(in, weights, dim) (api/c/mean.cpp:95-149): a weights dim of 1 is repeated up to the input's"""
    h = _vp()
    _check(lib.af_mean_weighted(ctypes.byref(h), a.handle, weights.handle, ctypes.c_longlong(_fnsd(a, dim))))
    return Array(h)


def mean_all(a):
    return _all(lib.af_mean_all, a)


def var(a, dim=-1, bias=AF_VARIANCE_POPULATION):
    return _stat_dim(lib.af_var_v2, a, dim, ctypes.c_int(bias))


def var_all(a, bias=AF_VARIANCE_POPULATION):
    return _all(lib.af_var_all_v2, a, ctypes.c_int(bias))


def stdev(a, dim=-1, bias=AF_VARIANCE_POPULATION):
    return _stat_dim(lib.af_stdev_v2, a, dim, ctypes.c_int(bias))


def stdev_all(a, bias=AF_VARIANCE_POPULATION):
    return _all(lib.af_stdev_all_v2, a, ctypes.c_int(bias))


def meanvar(a, dim=-1, bias=AF_VARIANCE_POPULATION, weights=None):
    """af::meanvar (api/cpp/var.cpp; api/c/var.cpp:418-470): weights=None is the reference's empty weights array"""
    m, v = _vp(), _vp()
    _check(lib.af_meanvar(ctypes.byref(m), ctypes.byref(v), a.handle, weights.handle if weights is not None else None, ctypes.c_int(bias),
                          ctypes.c_longlong(_fnsd(a, dim))))
    return Array(m), Array(v)


def var_weighted(a, weights, dim=-1):
    """af::var(in, weights, dim) (api/c/var.cpp:263-325)"""
    h = _vp()
    _check(lib.af_var_weighted(ctypes.byref(h), a.handle, weights.handle, ctypes.c_longlong(_fnsd(a, dim))))
    return Array(h)


def norm(a, kind=AF_NORM_VECTOR_2, p=1.0, q=1.0):
    out = ctypes.c_double()
    _check(lib.af_norm(ctypes.byref(out), a.handle, ctypes.c_int(kind), ctypes.c_double(p), ctypes.c_double(q)))
    return out.value


def corrcoef(x, y):
    re, im = ctypes.c_double(), ctypes.c_double()
    _check(lib.af_corrcoef(ctypes.byref(re), ctypes.byref(im), x.handle, y.handle))
    return re.value


def cov(x, y, bias=AF_VARIANCE_POPULATION):
    h = _vp()
    _check(lib.af_cov_v2(ctypes.byref(h), x.handle, y.handle, ctypes.c_int(bias)))
    return Array(h)


# ---- callers next to the scan / compaction path: af::sat, af::diff1, af::diff2, af::setUnique (sorted input) ----------------
def sat(a):
    h = _vp()
    _check(lib.af_sat(ctypes.byref(h), a.handle))
    return Array(h)


def diff1(a, dim=0):
    h = _vp()
    _check(lib.af_diff1(ctypes.byref(h), a.handle, ctypes.c_int(dim)))
    return Array(h)


def diff2(a, dim=0):
    h = _vp()
    _check(lib.af_diff2(ctypes.byref(h), a.handle, ctypes.c_int(dim)))
    return Array(h)


def set_unique(a, is_sorted=False):
    h = _vp()
    _check(lib.af_set_unique(ctypes.byref(h), a.handle, ctypes.c_bool(is_sorted)))
    return Array(h)


# ---- lazily evaluated element-wise expressions (af::operator+ ... af::sqrt): fused into whole-array reductions ------------
def _binary(fn, a, b):
    h = _vp()
    _check(fn(ctypes.byref(h), a.handle, b.handle, ctypes.c_bool(False)))
    return Array(h)


def _unary(fn, a):
    h = _vp()
    _check(fn(ctypes.byref(h), a.handle))
    return Array(h)


def _operand(a, like):
    """a scalar becomes a (lazy) constant of `like`'s dims, typed as the C++ operators do (array op double)"""
    if isinstance(a, Array):
        return a
    if isinstance(a, int) and not isinstance(a, bool) and like.dtype in (s32, u32, s16, u16, s8, u8):
        t = like.dtype if like.dtype == u32 else s32      # array op int: an s32 scalar (u32 against a u32 array), array.cpp implicit_dtype
    else:
        t = like.dtype if like.dtype in (f32, f64) else f32
    return constant(float(a), [int(d) for d in like.dims], t)


def _bin(fn):
    def op(a, b):
        like = a if isinstance(a, Array) else b
        return _binary(fn, _operand(a, like), _operand(b, like))
    return op


add, sub, mul, div = _bin(lib.af_add), _bin(lib.af_sub), _bin(lib.af_mul), _bin(lib.af_div)
lt, gt, le, ge, eq, neq = _bin(lib.af_lt), _bin(lib.af_gt), _bin(lib.af_le), _bin(lib.af_ge), _bin(lib.af_eq), _bin(lib.af_neq)
logical_and, logical_or = _bin(lib.af_and), _bin(lib.af_or)
minof, maxof, pow = _bin(lib.af_minof), _bin(lib.af_maxof), _bin(lib.af_pow)   # noqa: A001


def sqrt(a):
    return _unary(lib.af_sqrt, a)


def abs(a):   # noqa: A001
    return _unary(lib.af_abs, a)


def logical_not(a):
    return _unary(lib.af_not, a)


# ---- issue / collect (include/afb200.h afx_*_async): enqueue many ops, collect their O(1) host results afterwards --------
OP_ADD, OP_MUL, OP_MIN, OP_MAX, OP_COUNT, OP_ALL, OP_ANY = range(7)


class Pending:
    """an enqueued whole-array op whose scalar / index / hit count has not been collected yet; result() waits on the
    pinned host record the kernel's last thread writes (no cudaMemcpy, no stream synchronisation)"""

    def __init__(self, handle, kind, cplx):
        self.handle, self.kind, self.cplx = handle, kind, cplx

    def result(self):
        h, self.handle = self.handle, None
        if h is None:
            raise AfError(202, "this pending op was already collected")   # AF_ERR_ARG
        if self.kind == "where":
            out, off, tot = _vp(), ctypes.c_longlong(), ctypes.c_longlong()
            _check(lib.afx_pending_where(h, ctypes.byref(out), ctypes.byref(off), ctypes.byref(tot)))
            return Array(out), off.value, tot.value
        re, im, idx = ctypes.c_double(), ctypes.c_double(), ctypes.c_uint()
        _check(lib.afx_pending_scalar(h, ctypes.byref(re), ctypes.byref(im), ctypes.byref(idx)))
        v = complex(re.value, im.value) if self.cplx else re.value
        return (v, idx.value) if self.kind == "indexed" else v

    def __del__(self):
        if getattr(self, "handle", None) is not None and lib is not None:
            lib.afx_pending_release(self.handle)
            self.handle = None


def reduce_all_async(a, op, comm=None, nanval=None):
    """af_<op>_all (op: OP_ADD .. OP_ANY) enqueued; with `comm` (afx_comm_create) `a` is this rank's shard of a global array"""
    h = _vp()
    _check(lib.afx_reduce_all_async(ctypes.byref(h), comm, a.handle, ctypes.c_int(op), ctypes.c_bool(nanval is not None),
                                    ctypes.c_double(0.0 if nanval is None else nanval)))
    return Pending(h, "scalar", a.dtype in (c32, c64) and op in (OP_ADD, OP_MUL, OP_MIN, OP_MAX))


def ireduce_all_async(a, op, comm=None, base=0):
    h = _vp()
    _check(lib.afx_ireduce_all_async(ctypes.byref(h), comm, a.handle, ctypes.c_int(op), ctypes.c_uint(base)))
    return Pending(h, "indexed", a.dtype in (c32, c64))


def where_async(a, comm=None, base=0):
    """af_where enqueued; result() -> (indices [+ base], offset of the slice in the global output, global hit count)"""
    h = _vp()
    _check(lib.afx_where_async(ctypes.byref(h), comm, a.handle, ctypes.c_uint(base)))
    return Pending(h, "where", False)


# ---- device ------------------------------------------------------------------------------------------
def sync(device=-1):
    _check(lib.af_sync(ctypes.c_int(device)))


def set_device(device):
    _check(lib.af_set_device(ctypes.c_int(device)))


def get_device():
    d = ctypes.c_int()
    _check(lib.af_get_device(ctypes.byref(d)))
    return d.value


def device_count():
    n = ctypes.c_int()
    _check(lib.af_get_device_count(ctypes.byref(n)))
    return n.value


def constant(val, dims, dtype=f32):
    h = _vp()
    _check(lib.af_constant(ctypes.byref(h), ctypes.c_double(val), len(dims), _dims_arr(dims), dtype))
    return Array(h)


def info_string():
    s = ctypes.c_char_p()
    _check(lib.af_info_string(ctypes.byref(s), ctypes.c_bool(False)))
    text = s.value.decode()
    lib.af_free_host(s)
    return text


def get_stream(device=None):
    """afcu_get_stream: the cudaStream_t every op of `device` is enqueued on (as an int)."""
    s = _vp()
    _check(lib.afcu_get_stream(ctypes.byref(s), ctypes.c_int(get_device() if device is None else device)))
    return s.value or 0


def launch_count(reset=False):
    n = ctypes.c_ulonglong()
    _check(lib.afx_launch_count(ctypes.byref(n), ctypes.c_bool(reset)))
    return n.value


def device_gc():
    _check(lib.af_device_gc())
