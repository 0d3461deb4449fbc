"""numpy front-end to the CPU checkers under oracle/ -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (arrayfire_b200, libafb200.so) never does.

Two layers are restated here:

* ``Oracle`` -- ctypes binding of the ``orc_*`` (restatement, oracle/oracle.cpp) or
  ``ref_*`` (the reference's own CPU kernels, oracle/_ref/libafref.so) entry points
  declared in oracle_abi.h.  This is the backend level (``detail::reduce<op,Ti,To>``).
* ``AfOracle`` -- the C-API level on top of it: the dtype promotion tables and the
  early-outs of /root/reference/src/api/c/reduce.cpp:82-121,190-228,305-378,571-860,
  918-1148, scan.cpp:118-283 and where.cpp:34-65, so a test can ask "what would
  af_sum(x, dim) return" and get (dtype code, numpy array).

Arrays are numpy arrays in Fortran (column-major) order with at most 4 dims, or
strided views of such arrays whose first-axis stride is one element (the reference
guarantees strides[0]==1 for every view, src/backend/cuda/Array.cpp:399-417).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# af_dtype codes, include/af/defines.h:210-233
F32, C32, F64, C64, B8, S32, U32, U8, S64, U64, S16, U16, F16, S8 = range(14)
DTYPE_NAMES = ["f32", "c32", "f64", "c64", "b8", "s32", "u32", "u8", "s64", "u64", "s16", "u16", "f16", "s8"]
NP_OF = {
    F32: np.float32, C32: np.complex64, F64: np.float64, C64: np.complex128, B8: np.int8, S32: np.int32,
    U32: np.uint32, U8: np.uint8, S64: np.int64, U64: np.uint64, S16: np.int16, U16: np.uint16,
    F16: np.float16, S8: np.int8,
}
OP_ADD, OP_MUL, OP_MIN, OP_MAX, OP_NOTZERO, OP_AND, OP_OR = range(7)
# af_binary_op, include/af/defines.h:433-438
AF_BINARY_ADD, AF_BINARY_MUL, AF_BINARY_MIN, AF_BINARY_MAX = range(4)


def code_of(arr, dtype=None):
    """af_dtype code of a numpy array (b8 must be asked for explicitly or be np.bool_)."""
    if dtype is not None:
        return dtype
    dt = arr.dtype
    if dt == np.bool_:
        return B8
    for code in (F32, C32, F64, C64, S32, U32, U8, S64, U64, S16, U16, F16, S8):
        if dt == NP_OF[code]:
            return code
    raise TypeError("unsupported numpy dtype %s" % dt)


def dims4(shape):
    d = list(shape) + [1] * (4 - len(shape))
    assert len(d) == 4, "at most 4 dims"
    return d


def af_ndims(d):
    """dim4::ndims(), src/backend/common/dim4.cpp:48-58."""
    n = d[0] * d[1] * d[2] * d[3]
    if n == 0:
        return 0
    if n == 1:
        return 1
    for i in (3, 2, 1):
        if d[i] != 1:
            return i + 1
    return 1


def as_view(arr):
    """(pointer, dims[4], strides[4] in elements, keepalive) of a column-major numpy view."""
    a = np.asarray(arr)
    if a.ndim == 0:
        a = a.reshape(1)
    if a.dtype == np.bool_:
        a = a.view(np.int8)
    d = dims4(a.shape)
    item = a.dtype.itemsize
    st = [s // item for s in a.strides] + [0] * (4 - a.ndim)
    # give trailing unit dims the dense continuation stride, as the reference does
    dense = 1
    for i in range(4):
        if i >= a.ndim:
            st[i] = max(dense, 1)
        dense = max(dense, st[i] * d[i]) if d[i] > 0 else dense
    if a.size > 0 and d[0] > 1:
        assert st[0] == 1, "first-axis stride must be 1 (pass Fortran-ordered arrays)"
    return a.ctypes.data, d, st, a


def _dim_arr(v):
    return (ctypes.c_longlong * 4)(*[int(x) for x in v])


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(HERE, "liboracle.so")
    ref = os.path.join(HERE, "_ref", "libafref.so")
    need = force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "oracle.cpp"))
    have_src = os.path.exists("/root/reference/src/backend/cpu/kernel/reduce.hpp")
    if have_src and (force or not os.path.exists(ref)
                     or os.path.getmtime(ref) < os.path.getmtime(os.path.join(HERE, "ref_driver.cpp"))):
        need = True
    if need:
        subprocess.check_call(["make", "-s", "-C", HERE])


class Oracle:
    """Backend-level checker: ``kind`` is "port" (restatement) or "reference" (oracle/_ref)."""

    def __init__(self, kind="port"):
        self.kind = kind
        if kind == "port":
            path, self.prefix = os.path.join(HERE, "liboracle.so"), "orc"
        elif kind == "reference":
            path, self.prefix = os.path.join(HERE, "_ref", "libafref.so"), "ref"
        else:
            raise ValueError(kind)
        if not os.path.exists(path):
            build()
        self.lib = ctypes.CDLL(path)
        self.lib[self.prefix + "_where"].restype = ctypes.c_longlong
        assert self.lib[self.prefix + "_is_reference"]() == (1 if kind == "reference" else 0)

    @staticmethod
    def available(kind):
        p = os.path.join(HERE, "liboracle.so") if kind == "port" else os.path.join(HERE, "_ref", "libafref.so")
        return os.path.exists(p)

    def _fn(self, name):
        return self.lib[self.prefix + "_" + name]

    # -- backend-level entry points (numpy in, numpy out) ---------------------------------
    def reduce(self, op, x, dim, to, ti=None, change_nan=False, nanval=0.0):
        ptr, d, st, keep = as_view(x)
        ti = code_of(np.asarray(x), ti)
        od = list(d)
        od[dim] = 1
        out = np.zeros(od, dtype=NP_OF[to], order="F")
        rc = self._fn("reduce")(op, ti, to, ctypes.c_void_p(ptr), _dim_arr(d), _dim_arr(st), dim,
                                int(change_nan), ctypes.c_double(nanval), ctypes.c_void_p(out.ctypes.data))
        if rc != 0:
            raise TypeError("reduce: no (op=%d, %s -> %s) instantiation" % (op, DTYPE_NAMES[ti], DTYPE_NAMES[to]))
        return out

    def reduce_all(self, op, x, to, ti=None, change_nan=False, nanval=0.0):
        ptr, d, st, keep = as_view(x)
        ti = code_of(np.asarray(x), ti)
        out = np.zeros(1, dtype=NP_OF[to])
        rc = self._fn("reduce_all")(op, ti, to, ctypes.c_void_p(ptr), _dim_arr(d), _dim_arr(st),
                                    int(change_nan), ctypes.c_double(nanval), ctypes.c_void_p(out.ctypes.data))
        if rc != 0:
            raise TypeError("reduce_all: no (op=%d, %s -> %s) instantiation" % (op, DTYPE_NAMES[ti], DTYPE_NAMES[to]))
        return out

    def ireduce(self, op, x, dim, t=None):
        ptr, d, st, keep = as_view(x)
        t = code_of(np.asarray(x), t)
        od = list(d)
        od[dim] = 1
        val = np.zeros(od, dtype=NP_OF[t], order="F")
        loc = np.zeros(od, dtype=np.uint32, order="F")
        rc = self._fn("ireduce")(op, t, ctypes.c_void_p(ptr), _dim_arr(d), _dim_arr(st), dim,
                                 ctypes.c_void_p(val.ctypes.data), ctypes.c_void_p(loc.ctypes.data))
        if rc != 0:
            raise TypeError("ireduce: unsupported")
        return val, loc

    def ireduce_all(self, op, x, t=None):
        ptr, d, st, keep = as_view(x)
        t = code_of(np.asarray(x), t)
        val = np.zeros(1, dtype=NP_OF[t])
        loc = np.zeros(1, dtype=np.uint32)
        rc = self._fn("ireduce_all")(op, t, ctypes.c_void_p(ptr), _dim_arr(d), _dim_arr(st),
                                     ctypes.c_void_p(val.ctypes.data), ctypes.c_void_p(loc.ctypes.data))
        if rc != 0:
            raise TypeError("ireduce_all: unsupported")
        return val[0], int(loc[0])

    def scan(self, op, x, dim, to, inclusive=True, ti=None):
        ptr, d, st, keep = as_view(x)
        ti = code_of(np.asarray(x), ti)
        out = np.zeros(d, dtype=NP_OF[to], order="F")
        rc = self._fn("scan")(op, ti, to, ctypes.c_void_p(ptr), _dim_arr(d), _dim_arr(st), dim, int(inclusive),
                              ctypes.c_void_p(out.ctypes.data))
        if rc != 0:
            raise TypeError("scan: no (op=%d, %s -> %s) instantiation" % (op, DTYPE_NAMES[ti], DTYPE_NAMES[to]))
        return out

    def scan_by_key(self, op, key, x, dim, inclusive=True, t=None, tk=None):
        kptr, kd, kst, kkeep = as_view(key)
        ptr, d, st, keep = as_view(x)
        assert kd == d
        t = code_of(np.asarray(x), t)
        tk = code_of(np.asarray(key), tk)
        out = np.zeros(d, dtype=NP_OF[t], order="F")
        rc = self._fn("scan_by_key")(op, t, tk, ctypes.c_void_p(kptr), _dim_arr(kst), ctypes.c_void_p(ptr),
                                     _dim_arr(d), _dim_arr(st), dim, int(inclusive), ctypes.c_void_p(out.ctypes.data))
        if rc != 0:
            raise TypeError("scan_by_key: unsupported (%s, key %s)" % (DTYPE_NAMES[t], DTYPE_NAMES[tk]))
        return out

    def reduce_by_key(self, op, keys, x, dim, to, ti=None, tk=None, change_nan=False, nanval=0.0):
        """(keys_out, vals_out): runs of equal consecutive keys reduced along `dim` (two calls: count, then reduce)"""
        ptr, d, st, keep = as_view(x)
        ti = code_of(np.asarray(x), ti)
        k = np.ascontiguousarray(np.asarray(keys).reshape(-1))
        tk = code_of(k, tk)
        assert k.size == d[dim]
        fn = self._fn("reduce_by_key")
        fn.restype = ctypes.c_longlong
        args = (op, ti, to, tk, ctypes.c_void_p(k.ctypes.data), ctypes.c_void_p(ptr), _dim_arr(d), _dim_arr(st), dim,
                int(change_nan), ctypes.c_double(nanval))
        runs = fn(*args, None, None)
        if runs < 0:
            raise TypeError("reduce_by_key: no (op=%d, %s -> %s, key %s) instantiation" % (op, DTYPE_NAMES[ti], DTYPE_NAMES[to], DTYPE_NAMES[tk]))
        od = list(d)
        od[dim] = runs
        kout = np.zeros(runs, dtype=k.dtype)
        vout = np.zeros(od, dtype=NP_OF[to], order="F")
        fn(*args, ctypes.c_void_p(kout.ctypes.data), ctypes.c_void_p(vout.ctypes.data))
        return kout, vout

    def ireduce_ragged(self, op, x, dim, rlen, t=None):
        ptr, d, st, keep = as_view(x)
        t = code_of(np.asarray(x), t)
        od = list(d)
        od[dim] = 1
        r = np.asfortranarray(np.asarray(rlen, dtype=np.uint32).reshape(od, order="F"))
        val = np.zeros(od, dtype=NP_OF[t], order="F")
        loc = np.zeros(od, dtype=np.uint32, order="F")
        rc = self._fn("ireduce_ragged")(op, t, ctypes.c_void_p(ptr), _dim_arr(d), _dim_arr(st), dim,
                                        ctypes.c_void_p(r.ctypes.data), ctypes.c_void_p(val.ctypes.data),
                                        ctypes.c_void_p(loc.ctypes.data))
        if rc != 0:
            raise TypeError("ireduce_ragged: unsupported")
        return val, loc

    def where(self, x, t=None):
        ptr, d, st, keep = as_view(x)
        t = code_of(np.asarray(x), t)
        n = d[0] * d[1] * d[2] * d[3]
        out = np.zeros(max(n, 1), dtype=np.uint32)
        cnt = self._fn("where")(t, ctypes.c_void_p(ptr), _dim_arr(d), _dim_arr(st), ctypes.c_void_p(out.ctypes.data))
        if cnt < 0:
            raise TypeError("where: unsupported dtype")
        return out[:cnt].copy()


# ---------------------------------------------------------------------------------------------
# C-API level (what af_sum & co. return): promotion tables + early-outs
# ---------------------------------------------------------------------------------------------
_SUM_TO = {F32: F32, F64: F64, C32: C32, C64: C64, U32: U32, S32: S32, U64: U64, S64: S64,
           U16: U32, S16: S32, U8: U32, S8: S32, F16: F32}  # api/c/reduce.cpp:322-371
_SCAN_TO = {F32: F32, F64: F64, C32: C32, C64: C64, U32: U32, S32: S32, U64: U64, S64: S64,
            U16: U32, S16: S32, U8: U32, S8: S32, B8: U32}  # api/c/scan.cpp:133-149,175-215
_SBK_TO = {F32: F32, F64: F64, C32: C32, C64: C64, S16: S32, S32: S32, S8: S32, U64: U64, S64: S64,
           U16: U32, U32: U32, U8: U32, B8: U32}  # api/c/scan.cpp:244-276
_BINOP = {AF_BINARY_ADD: OP_ADD, AF_BINARY_MUL: OP_MUL, AF_BINARY_MIN: OP_MIN, AF_BINARY_MAX: OP_MAX}


# output dtype of mean / var / stdev along a dimension (api/c/mean.cpp:71-86)
STAT_TO = {F32: F32, S32: F32, U32: F32, S16: F32, U16: F32, S8: F32, U8: F32, B8: F32, F64: F64, S64: F64, U64: F64,
           C32: C32, C64: C64, F16: F16}


def _stat_methods():
    def mean(self, x, dim, ti=None):
        ptr, d, st, keep = as_view(x)
        ti = code_of(np.asarray(x), ti)
        od = list(d)
        od[dim] = 1
        out = np.zeros(od, dtype=NP_OF[STAT_TO[ti]], order="F")
        rc = self._fn("mean")(ti, ctypes.c_void_p(ptr), _dim_arr(d), _dim_arr(st), dim, ctypes.c_void_p(out.ctypes.data))
        if rc != 0:
            raise TypeError("mean: no instantiation for %s" % DTYPE_NAMES[ti])
        return out

    def mean_all(self, x, ti=None):
        ptr, d, st, keep = as_view(x)
        ti = code_of(np.asarray(x), ti)
        out = (ctypes.c_double * 2)()
        rc = self._fn("mean_all")(ti, ctypes.c_void_p(ptr), _dim_arr(d), _dim_arr(st), out)
        if rc != 0:
            raise TypeError("mean_all: no instantiation for %s" % DTYPE_NAMES[ti])
        return complex(out[0], out[1]) if ti in (C32, C64) else out[0]

    def var(self, x, dim, denom, take_sqrt=False, want_mean=False, ti=None):
        ptr, d, st, keep = as_view(x)
        ti = code_of(np.asarray(x), ti)
        od = list(d)
        od[dim] = 1
        out = np.zeros(od, dtype=NP_OF[STAT_TO[ti]], order="F")
        mo = np.zeros(od, dtype=NP_OF[STAT_TO[ti]], order="F") if want_mean else None
        rc = self._fn("var")(ti, ctypes.c_void_p(ptr), _dim_arr(d), _dim_arr(st), dim, ctypes.c_double(denom), int(take_sqrt),
                             ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(mo.ctypes.data if want_mean else None))
        if rc != 0:
            raise TypeError("var: no instantiation for %s" % DTYPE_NAMES[ti])
        return (mo, out) if want_mean else out

    def var_all(self, x, denom, take_sqrt=False, ti=None):
        ptr, d, st, keep = as_view(x)
        ti = code_of(np.asarray(x), ti)
        out = (ctypes.c_double * 2)()
        rc = self._fn("var_all")(ti, ctypes.c_void_p(ptr), _dim_arr(d), _dim_arr(st), ctypes.c_double(denom), int(take_sqrt), out)
        if rc != 0:
            raise TypeError("var_all: no instantiation for %s" % DTYPE_NAMES[ti])
        return complex(out[0], out[1]) if ti in (C32, C64) else out[0]

    Oracle.mean, Oracle.mean_all, Oracle.var, Oracle.var_all = mean, mean_all, var, var_all


_stat_methods()

AF_VARIANCE_DEFAULT, AF_VARIANCE_SAMPLE, AF_VARIANCE_POPULATION = 0, 1, 2
(AF_NORM_VECTOR_1, AF_NORM_VECTOR_INF, AF_NORM_VECTOR_2, AF_NORM_VECTOR_P, AF_NORM_MATRIX_1, AF_NORM_MATRIX_INF,
 AF_NORM_MATRIX_2, AF_NORM_MATRIX_L_PQ) = range(8)


class AfArgError(Exception):
    """AF_ERR_ARG (202)"""


class AfTypeError(Exception):
    """AF_ERR_TYPE (204)"""


class AfOracle:
    """Expected results of the af_* C entry points, as (dtype code, numpy array) pairs."""

    def __init__(self, kind="port"):
        self.o = Oracle(kind)

    @staticmethod
    def _check_dim(dim):
        if not (0 <= dim < 4):
            raise AfArgError("dim")

    def _dimwise(self, fn, x, dim, dtype):
        """shared prologue of reduce_promote / reduce_common / reduce_type."""
        self._check_dim(dim)
        _, d, _, keep = as_view(x)
        ti = code_of(np.asarray(x), dtype)
        if dim >= af_ndims(d):
            return ti, np.array(keep, copy=True, order="F").reshape(d, order="F")  # retain(in): unpromoted
        return fn(ti)

    # sum / product (+ _nan) ---------------------------------------------------------------
    def _promote(self, op, x, dim, dtype, change_nan, nanval):
        def go(ti):
            if ti == B8:  # api/c/reduce.cpp:358-366
                if op == OP_MUL:
                    return B8, self.o.reduce(OP_AND, x, dim, B8, ti, change_nan, nanval)
                return U32, self.o.reduce(OP_NOTZERO, x, dim, U32, ti, change_nan, nanval)
            if ti not in _SUM_TO:
                raise AfTypeError
            return _SUM_TO[ti], self.o.reduce(op, x, dim, _SUM_TO[ti], ti, change_nan, nanval)
        return self._dimwise(go, x, dim, dtype)

    def sum(self, x, dim, dtype=None):
        return self._promote(OP_ADD, x, dim, dtype, False, 0.0)

    def product(self, x, dim, dtype=None):
        return self._promote(OP_MUL, x, dim, dtype, False, 0.0)

    def sum_nan(self, x, dim, nanval, dtype=None):
        return self._promote(OP_ADD, x, dim, dtype, True, nanval)

    def product_nan(self, x, dim, nanval, dtype=None):
        return self._promote(OP_MUL, x, dim, dtype, True, nanval)

    # min / max ----------------------------------------------------------------------------
    def min(self, x, dim, dtype=None):
        return self._dimwise(lambda ti: (ti, self.o.reduce(OP_MIN, x, dim, ti, ti)), x, dim, dtype)

    def max(self, x, dim, dtype=None):
        return self._dimwise(lambda ti: (ti, self.o.reduce(OP_MAX, x, dim, ti, ti)), x, dim, dtype)

    # count / all / any --------------------------------------------------------------------
    def count(self, x, dim, dtype=None):
        return self._dimwise(lambda ti: (U32, self.o.reduce(OP_NOTZERO, x, dim, U32, ti)), x, dim, dtype)

    def all_true(self, x, dim, dtype=None):
        return self._dimwise(lambda ti: (B8, self.o.reduce(OP_AND, x, dim, B8, ti)), x, dim, dtype)

    def any_true(self, x, dim, dtype=None):
        return self._dimwise(lambda ti: (B8, self.o.reduce(OP_OR, x, dim, B8, ti)), x, dim, dtype)

    # whole-array forms: return the 1-element array of the accumulation type ----------------
    def _all_promote(self, op, x, dtype, change_nan, nanval, scalar_api):
        _, d, _, keep = as_view(x)
        ti = code_of(np.asarray(x), dtype)
        if ti == B8:
            if op == OP_MUL:
                return B8, self.o.reduce_all(OP_AND, x, B8, ti, change_nan, nanval)
            return U32, self.o.reduce_all(OP_NOTZERO, x, U32, ti, change_nan, nanval)
        if ti not in _SUM_TO:
            raise AfTypeError
        if scalar_api and ti in (C32, C64):
            change_nan = False  # api/c/reduce.cpp:762-773 drops change_nan/nanval for complex scalars
        return _SUM_TO[ti], self.o.reduce_all(op, x, _SUM_TO[ti], ti, change_nan, nanval)

    def sum_all(self, x, dtype=None, scalar_api=True):
        return self._all_promote(OP_ADD, x, dtype, False, 0.0, scalar_api)

    def product_all(self, x, dtype=None, scalar_api=True):
        return self._all_promote(OP_MUL, x, dtype, False, 0.0, scalar_api)

    def sum_nan_all(self, x, nanval, dtype=None, scalar_api=True):
        return self._all_promote(OP_ADD, x, dtype, True, nanval, scalar_api)

    def product_nan_all(self, x, nanval, dtype=None, scalar_api=True):
        return self._all_promote(OP_MUL, x, dtype, True, nanval, scalar_api)

    def _all_common(self, op, x, dtype):
        _, d, _, keep = as_view(x)
        if af_ndims(d) == 0:
            raise AfArgError("ndims > 0")  # api/c/reduce.cpp:646
        ti = code_of(np.asarray(x), dtype)
        return ti, self.o.reduce_all(op, x, ti, ti)

    def min_all(self, x, dtype=None):
        return self._all_common(OP_MIN, x, dtype)

    def max_all(self, x, dtype=None):
        return self._all_common(OP_MAX, x, dtype)

    def count_all(self, x, dtype=None):
        return U32, self.o.reduce_all(OP_NOTZERO, x, U32, code_of(np.asarray(x), dtype))

    def all_true_all(self, x, dtype=None):
        return B8, self.o.reduce_all(OP_AND, x, B8, code_of(np.asarray(x), dtype))

    def any_true_all(self, x, dtype=None):
        return B8, self.o.reduce_all(OP_OR, x, B8, code_of(np.asarray(x), dtype))

    # ireduce ------------------------------------------------------------------------------
    def _ireduce(self, op, x, dim, dtype):
        if not (0 <= dim < 4):
            raise AfArgError("dim")
        _, d, _, keep = as_view(x)
        if af_ndims(d) == 0:
            raise AfArgError("ndims > 0")  # api/c/reduce.cpp:957
        ti = code_of(np.asarray(x), dtype)
        if dim >= af_ndims(d):  # api/c/reduce.cpp:959-963
            return ti, np.array(keep, copy=True, order="F").reshape(d, order="F"), np.zeros(d, np.uint32, order="F")
        v, l = self.o.ireduce(op, x, dim, ti)
        return ti, v, l

    def imin(self, x, dim, dtype=None):
        return self._ireduce(OP_MIN, x, dim, dtype)

    def imax(self, x, dim, dtype=None):
        return self._ireduce(OP_MAX, x, dim, dtype)

    def _ireduce_all(self, op, x, dtype):
        _, d, _, keep = as_view(x)
        if af_ndims(d) == 0:
            raise AfArgError("ndims > 0")
        ti = code_of(np.asarray(x), dtype)
        if ti == F16:
            raise AfTypeError  # api/c/reduce.cpp:1095-1131 has no f16 case
        return self.o.ireduce_all(op, x, ti)

    def imin_all(self, x, dtype=None):
        return self._ireduce_all(OP_MIN, x, dtype)

    def imax_all(self, x, dtype=None):
        return self._ireduce_all(OP_MAX, x, dtype)

    # scans --------------------------------------------------------------------------------
    def accum(self, x, dim, dtype=None):
        def go(ti):
            if ti not in _SCAN_TO:
                raise AfTypeError
            op = OP_NOTZERO if ti == B8 else OP_ADD  # api/c/scan.cpp:146-148
            return _SCAN_TO[ti], self.o.scan(op, x, dim, _SCAN_TO[ti], True, ti)
        return self._dimwise(go, x, dim, dtype)

    def scan(self, x, dim, op=AF_BINARY_ADD, inclusive=True, dtype=None):
        def go(ti):
            if ti not in _SCAN_TO:
                raise AfTypeError
            if op not in _BINOP:
                raise AfArgError("op")
            return _SCAN_TO[ti], self.o.scan(_BINOP[op], x, dim, _SCAN_TO[ti], inclusive, ti)
        return self._dimwise(go, x, dim, dtype)

    def scan_by_key(self, key, x, dim, op=AF_BINARY_ADD, inclusive=True, dtype=None, key_dtype=None):
        self._check_dim(dim)
        _, d, _, keep = as_view(x)
        _, kd, _, kkeep = as_view(key)
        ti = code_of(np.asarray(x), dtype)
        if dim >= af_ndims(d):
            return ti, np.array(keep, copy=True, order="F").reshape(d, order="F")
        if kd != d:
            raise AfArgError("key dims")
        if ti not in _SBK_TO:
            raise AfTypeError
        if op not in _BINOP:
            raise AfArgError("op")
        to = _SBK_TO[ti]
        tk = code_of(np.asarray(key), key_dtype)
        if tk not in (S32, U32, S64, U64):
            raise AfTypeError
        cast = np.asarray(keep).astype(NP_OF[to], order="F") if to != ti else keep  # castArray<Ti>(in)
        return to, self.o.scan_by_key(_BINOP[op], key, cast, dim, inclusive, to, tk)

    # reductions by key (api/c/reduce.cpp:122-304, 500-554) -------------------------------------------
    def _by_key(self, fam, op, keys, x, dim, dtype, key_dtype, change_nan=False, nanval=0.0):
        self._check_dim(dim)
        _, d, _, keep = as_view(x)
        kd = dims4(np.asarray(keys).shape)
        ti = code_of(np.asarray(x), dtype)
        tk = code_of(np.asarray(keys), key_dtype)
        nons = sum(1 for v in kd if v not in (0, 1))
        if not (nons == 1 and sum(1 for v in kd if v == 1) == 3):
            raise AfArgError("keys must be a vector")
        klen = int(np.prod(kd)) if fam == "type" else kd[0]
        if d[dim] != klen:
            raise AfArgError("key length")
        if tk not in (S32, U32):
            raise AfTypeError
        if fam == "promote":
            if ti == B8:
                op, to = OP_NOTZERO, U32
            elif ti in _SUM_TO:
                to = _SUM_TO[ti]
            else:
                raise AfTypeError
        elif fam == "common":
            to = ti
        else:
            to = U32 if op == OP_NOTZERO else B8
        kout, vout = self.o.reduce_by_key(op, keys, x, dim, to, ti, tk, change_nan, nanval)
        return tk, kout, to, vout

    def sum_by_key(self, keys, x, dim, dtype=None, key_dtype=None, nanval=None):
        return self._by_key("promote", OP_ADD, keys, x, dim, dtype, key_dtype, nanval is not None, nanval or 0.0)

    def product_by_key(self, keys, x, dim, dtype=None, key_dtype=None, nanval=None):
        return self._by_key("promote", OP_MUL, keys, x, dim, dtype, key_dtype, nanval is not None, nanval or 0.0)

    def min_by_key(self, keys, x, dim, dtype=None, key_dtype=None):
        return self._by_key("common", OP_MIN, keys, x, dim, dtype, key_dtype)

    def max_by_key(self, keys, x, dim, dtype=None, key_dtype=None):
        return self._by_key("common", OP_MAX, keys, x, dim, dtype, key_dtype)

    def count_by_key(self, keys, x, dim, dtype=None, key_dtype=None):
        return self._by_key("type", OP_NOTZERO, keys, x, dim, dtype, key_dtype)

    def all_true_by_key(self, keys, x, dim, dtype=None, key_dtype=None):
        return self._by_key("type", OP_AND, keys, x, dim, dtype, key_dtype)

    def any_true_by_key(self, keys, x, dim, dtype=None, key_dtype=None):
        return self._by_key("type", OP_OR, keys, x, dim, dtype, key_dtype)

    def max_ragged(self, x, rlen, dim, dtype=None):
        """af_max_ragged (api/c/reduce.cpp:1002-1078)"""
        if not (0 <= dim < 4):
            raise AfArgError("dim")
        _, d, _, keep = as_view(x)
        if af_ndims(d) == 0:
            raise AfArgError("ndims > 0")
        ti = code_of(np.asarray(x), dtype)
        if dim >= af_ndims(d):
            return ti, np.array(keep, copy=True, order="F").reshape(d, order="F"), np.zeros(d, np.uint32, order="F")
        v, l = self.o.ireduce_ragged(OP_MAX, x, dim, rlen, ti)
        return ti, v, l

    def where(self, x, dtype=None):
        _, d, _, keep = as_view(x)
        ti = code_of(np.asarray(x), dtype)
        if af_ndims(d) == 0:
            return np.zeros(0, np.uint32)  # api/c/where.cpp:39-41
        if ti == F16:
            raise AfTypeError
        return self.o.where(x, ti)


    # ---- statistics ---------------------------------------------------------------
    def mean(self, x, dim, dtype=None):
        """af_mean (api/c/mean.cpp:65-93)"""
        if not (0 <= dim <= 3):
            raise AfArgError("dim")
        ti = code_of(np.asarray(x), dtype)
        return STAT_TO[ti], self.o.mean(x, dim, ti)

    def mean_all(self, x, dtype=None):
        return self.o.mean_all(x, code_of(np.asarray(x), dtype))

    def var(self, x, dim, bias=AF_VARIANCE_POPULATION, dtype=None, want_mean=False):
        """af_var_v2 / af_meanvar: 1 / (population ? N : N - 1) -- DEFAULT divides by N - 1 here (api/c/var.cpp:118-121)"""
        if not (0 <= dim <= 3):
            raise AfArgError("dim")
        ti = code_of(np.asarray(x), dtype)
        n = dims4(np.asarray(x).shape)[dim]
        denom = n if bias == AF_VARIANCE_POPULATION else n - 1
        return STAT_TO[ti], self.o.var(x, dim, denom, False, want_mean, ti)

    @staticmethod
    def meanvar_weighted(x, w, dim, bias=AF_VARIANCE_POPULATION):
        """af_meanvar with weights / af_var_weighted (api/c/var.cpp:104-142, 263-325), restated step by step in the result type
        (f32; f64 for f64 / s64 / u64 inputs): mean = sum(x w) / sum(w) (mean.cpp, weighted); the squared deviations from that
        mean are summed UNWEIGHTED (var.cpp:134-136) and scaled by 1 / (sum(w) - (bias == SAMPLE)) (var.cpp:122-130)."""
        if not (0 <= dim <= 3):
            raise AfArgError("dim")
        x, w = np.asarray(x), np.asarray(w)
        if w.dtype not in (np.float32, np.float64):
            raise AfArgError("weights")
        tc = np.float64 if x.dtype in (np.float64, np.int64, np.uint64) else np.float32
        d = dims4(x.shape)
        xc, wc = x.astype(tc).reshape(d, order="F"), w.astype(tc).reshape(d, order="F")
        wsum = wc.sum(axis=dim, keepdims=True, dtype=tc)
        mean = ((xc * wc).sum(axis=dim, keepdims=True, dtype=tc) / wsum).astype(tc)
        diff = xc - mean
        red = (diff * diff).sum(axis=dim, keepdims=True, dtype=tc)
        norm = tc(1) / (wsum - tc(1) if bias == AF_VARIANCE_SAMPLE else wsum)
        return mean, (norm * red).astype(tc)

    def var_all(self, x, bias=AF_VARIANCE_POPULATION, dtype=None):
        """af_var_all_v2: N - (bias == SAMPLE) (api/c/var.cpp:71-73)"""
        return self.o.var_all(x, np.asarray(x).size - (1 if bias == AF_VARIANCE_SAMPLE else 0), False, code_of(np.asarray(x), dtype))

    def stdev(self, x, dim, bias=AF_VARIANCE_POPULATION, dtype=None):
        """af_stdev_v2: N - (bias == SAMPLE) (api/c/stdev.cpp:80-81); no complex / f16 case"""
        ti = code_of(np.asarray(x), dtype)
        if ti in (C32, C64, F16):
            raise AfTypeError
        n = dims4(np.asarray(x).shape)[dim]
        return STAT_TO[ti], self.o.var(x, dim, n - (1 if bias == AF_VARIANCE_SAMPLE else 0), True, False, ti)

    def stdev_all(self, x, bias=AF_VARIANCE_POPULATION, dtype=None):
        ti = code_of(np.asarray(x), dtype)
        if ti in (C32, C64, F16):
            raise AfTypeError
        return self.o.var_all(x, np.asarray(x).size - (1 if bias == AF_VARIANCE_SAMPLE else 0), True, ti)

    @staticmethod
    def norm(x, kind=AF_NORM_VECTOR_2, p=1.0, q=1.0):
        """af_norm (api/c/norm.cpp:45-151): |x| first, every reduction serial in the result type (f32 for f32 / c32 / f16)"""
        a = np.asarray(x)
        rt = np.float64 if a.dtype in (np.float64, np.complex128) else np.float32
        A = np.abs(a).astype(rt).reshape(dims4(a.shape)[:2], order="F") if a.ndim <= 2 else None
        if A is None:
            raise AfArgError("batch")
        if a.size == 0:
            return 0.0

        def total(v):   # reduce_all<af_add_t>: serial, in rt
            acc = rt(0)
            for e in v.reshape(-1, order="F"):
                acc = rt(e + acc)
            return acc

        def vec(pp):
            if pp == 1:
                return float(total(A))
            if np.isinf(pp):
                return float(A.max())
            if pp == 2:
                return float(np.sqrt(total((A * A).astype(rt))))
            return float(total(np.power(A, rt(pp)).astype(rt)) ** (1.0 / pp))

        if kind == AF_NORM_VECTOR_2:
            return vec(2)
        if kind == AF_NORM_VECTOR_1:
            return vec(1)
        if kind == AF_NORM_VECTOR_INF:
            return vec(np.inf)
        if kind == AF_NORM_VECTOR_P:
            return vec(p)
        if kind == AF_NORM_MATRIX_1:
            return float(max(total(A[:, j]) for j in range(A.shape[1])))
        if kind == AF_NORM_MATRIX_INF:
            return float(max(total(A[i, :]) for i in range(A.shape[0])))
        if kind == AF_NORM_MATRIX_L_PQ:
            if p == 1:
                cols = np.array([total(A[:, j]) for j in range(A.shape[1])], dtype=rt)
            else:
                cols = np.array([total(np.power(A[:, j], rt(p)).astype(rt)) for j in range(A.shape[1])], dtype=rt)
                cols = np.power(cols, rt(1.0 / p)).astype(rt)
            if q == 1:
                return float(total(cols))
            return float(total(np.power(cols, rt(q)).astype(rt)) ** (1.0 / q))
        raise AfTypeError

    @staticmethod
    def corrcoef(x, y):
        """af_corrcoef (api/c/corrcoef.cpp:39-60): five serial sums in the compute type, then the closed formula"""
        a, b = np.asarray(x), np.asarray(y)
        rt = np.float64 if a.dtype in (np.float64, np.int64, np.uint64) else np.float32
        xa, ya = a.astype(rt).reshape(-1, order="F"), b.astype(rt).reshape(-1, order="F")

        def total(v):
            acc = rt(0)
            for e in v:
                acc = rt(e + acc)
            return acc
        n = rt(xa.size)
        xs, ys, xx, yy, xy = total(xa), total(ya), total(xa * xa), total(ya * ya), total(xa * ya)
        return float((n * xy - xs * ys) / (np.sqrt(n * xx - xs * xs) * np.sqrt(n * yy - ys * ys)))

    def cov(self, x, y, bias=AF_VARIANCE_POPULATION):
        """af_cov_v2 (api/c/covariance.cpp:41-63): the means are those of the WHOLE arrays; sums along dim 0"""
        a, b = np.asarray(x), np.asarray(y)
        rt = np.float64 if a.dtype in (np.float64, np.int64, np.uint64) else np.float32
        d = dims4(a.shape)
        n = d[0] - 1 if bias == AF_VARIANCE_SAMPLE else d[0]
        mx, my = rt(self.o.mean_all(a)), rt(self.o.mean_all(b))
        prod = ((a.astype(rt) - mx) * (b.astype(rt) - my)).astype(rt).reshape(d[:2], order="F")
        out = np.zeros((1, d[1]), dtype=rt, order="F")
        for j in range(d[1]):
            acc = rt(0)
            for e in prod[:, j]:
                acc = rt(e + acc)
            out[0, j] = acc / rt(n)
        return (F64 if rt is np.float64 else F32), out

    # ---- callers next to the path: thin numpy restatements of serial loops --------------------
    def sat(self, x, dtype=None):
        """af_sat (api/c/sat.cpp:32-61): cast (b8 -> s32, narrow ints as the scan promotes), scan dim 0, scan dim 1"""
        ti = code_of(np.asarray(x), dtype)
        if ti in (C32, C64, F16):
            raise AfTypeError
        if af_ndims(dims4(np.asarray(x).shape)) < 2:
            raise AfArgError("ndims")
        a = np.asarray(x)
        if ti == B8:
            a = np.asfortranarray(a.astype(np.int32))
            ti = S32
        t1, r1 = self.scan(a, 0, AF_BINARY_ADD, True, dtype=ti)
        return self.scan(r1, 1, AF_BINARY_ADD, True, dtype=t1)

    @staticmethod
    def diff(x, dim, order):
        """af_diff1 / af_diff2 (cpu/kernel/diff.hpp:18-81): same dtype, wrap-around for integers"""
        a = np.asarray(x).reshape(dims4(np.asarray(x).shape), order="F")
        if a.shape[dim] < order + 1:
            return np.zeros((0,), dtype=a.dtype)
        sl = lambda lo, hi: tuple(slice(lo, hi) if i == dim else slice(None) for i in range(4))   # noqa: E731
        n = a.shape[dim]
        with np.errstate(over="ignore"):
            if order == 1:
                return (a[sl(1, n)] - a[sl(0, n - 1)]).astype(a.dtype)
            return (((a[sl(2, n)] + a[sl(0, n - 2)]).astype(a.dtype) - a[sl(1, n - 1)]).astype(a.dtype) - a[sl(1, n - 1)]).astype(a.dtype)

    @staticmethod
    def set_unique_sorted(x):
        """std::unique of an already sorted vector (cpu/set.cpp:32-51)"""
        a = np.asarray(x).reshape(-1, order="F")
        if a.size <= 1:
            return a.copy()
        keep = np.ones(a.size, dtype=bool)
        keep[1:] = ~(a[1:] == a[:-1])
        return a[keep]
