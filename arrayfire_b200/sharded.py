"""Sharded (one process per GPU) whole-array ops: the multi-GPU layer of the path.

New functionality -- the reference has no multi-GPU path: an array belongs to exactly one
device (/root/reference/src/backend/cuda/Array.cpp:61-74).  Here a flattened array of N elements is split
into contiguous shards, shard r = [r*N/G, (r+1)*N/G) resident on rank r's GPU; every op runs the SAME
single-GPU kernels on the local shard (through the afx_* entry points of include/afb200.h) and combines
O(1) per-shard partials across ranks with ONE small NCCL all-gather (torch.distributed), folded in rank
order so results are deterministic and follow the reference's serial semantics:

  sum/product/count/any/all, min/max   partial scalar            fold with the op, rank order
  imin / imax                           (value, global index)     max -> lowest index among ties, min -> highest
  scan                                  shard total               exclusive fold over lower ranks = carry-in
  scan_by_key                           (tail total, single-seg,  carry reaches only the first segment of a shard:
                                         first key, last key)      local scan + fix-up of that prefix
  where / count                         hit count                 exclusive prefix = output offset (indices carry
                                                                    the shard base, output stays sharded)

The combine rules are plain functions over gathered tensors (testable on CPU with the gloo backend,
tests/test_sharded_gloo.py); the GPU binding at the bottom feeds them from libafb200.
"""
import numpy as np

OP_SUM, OP_PRODUCT, OP_MIN, OP_MAX, OP_COUNT, OP_ALL, OP_ANY = range(7)


def shard_bounds(n, world, rank):
    """[lo, hi) of shard `rank`: sizes differ by at most one element, lower ranks get the extra ones."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# ---------------------------------------------------------------------------------------------
# combine rules (numpy, host side; inputs are the all-gathered per-rank partials in rank order)
# ---------------------------------------------------------------------------------------------
def fold_partials(op, parts):
    """rank-ordered fold of scalar partials with the reduction's own operator"""
    parts = np.asarray(parts)
    acc = parts[0]
    for p in parts[1:]:
        if op in (OP_SUM, OP_COUNT):
            acc = (acc + p).astype(parts.dtype) if parts.dtype.kind in "iu" else parts.dtype.type(acc + p)
        elif op == OP_PRODUCT:
            acc = (acc * p).astype(parts.dtype) if parts.dtype.kind in "iu" else parts.dtype.type(acc * p)
        elif op == OP_MIN:
            acc = p if _less(p, acc) else acc
        elif op == OP_MAX:
            acc = p if _less(acc, p) else acc
        elif op == OP_ALL:
            acc = parts.dtype.type(bool(acc) and bool(p))
        elif op == OP_ANY:
            acc = parts.dtype.type(bool(acc) or bool(p))
        else:
            raise ValueError(op)
    return acc


def _less(a, b):
    if np.iscomplexobj(a) or np.iscomplexobj(b):
        return abs(a) < abs(b)
    return a < b


def _key(v, is_b8=False):
    if is_b8:
        return float(v > 0)
    if np.iscomplexobj(v):
        return float(abs(v))
    return float(v)


def fold_arg(op, vals, gidx, is_b8=False):
    """(value, GLOBAL index) pairs in rank order -> the reference's MinMaxOp result
    (/root/reference/src/backend/cpu/kernel/ireduce.hpp:29-63): compare as double; on ties max keeps the
    lowest index, min the highest.  Shard partials already obey the rule locally and global indices grow
    with the rank, so the same comparison across ranks reproduces the serial answer."""
    bv, bi = vals[0], int(gidx[0])
    for v, i in zip(vals[1:], gidx[1:]):
        ka, kb = _key(bv, is_b8), _key(v, is_b8)
        if op == OP_MIN:
            take = kb < ka or (kb == ka and int(i) > bi)
        else:
            take = kb > ka or (kb == ka and int(i) < bi)
        if take:
            bv, bi = v, int(i)
    return bv, bi


def exclusive_fold(op, totals, rank):
    """carry-in of shard `rank` for a plain scan: fold of the totals of ranks < rank (None for rank 0)"""
    if rank == 0:
        return None
    return fold_partials({0: OP_SUM, 1: OP_PRODUCT, 2: OP_MIN, 3: OP_MAX}[op], np.asarray(totals)[:rank])


def segmented_carry(op, tail_totals, single_segment, first_keys, last_keys, rank):
    """carry-in of shard `rank` for scan_by_key, or None.

    tail_totals[r]   reduction of shard r's LAST segment (computed without any carry)
    single_segment[r] shard r consists of one segment only
    The carry exists when the key continues across the boundary; it is the tail total of shard rank-1,
    extended by that shard's own carry when it is a single segment (a run spanning several shards)."""
    binop = {0: OP_SUM, 1: OP_PRODUCT, 2: OP_MIN, 3: OP_MAX}[op]
    carry = None
    for r in range(1, rank + 1):
        if last_keys[r - 1] != first_keys[r]:
            carry = None
            continue
        prev = tail_totals[r - 1]
        if carry is not None and single_segment[r - 1]:
            prev = fold_partials(binop, np.asarray([carry, prev]))
        carry = prev
    return carry


def where_offsets(counts):
    """exclusive prefix of the per-shard hit counts and their total"""
    counts = np.asarray(counts, dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(counts)[:-1]])
    return off, int(counts.sum())


# ---------------------------------------------------------------------------------------------
# process-group plumbing
# ---------------------------------------------------------------------------------------------
class Group:
    """thin wrapper over torch.distributed (NCCL on GPUs, gloo in the CPU tests)"""

    def __init__(self, dist=None, device=None):
        self.dist = dist
        self.device = device
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1

    def all_gather_bytes(self, payload):
        """payload: 1-D uint8 torch tensor (same length on every rank) on self.device -> [world, len] tensor"""
        import torch
        if self.world == 1:
            return payload.reshape(1, -1)
        out = torch.empty(self.world * payload.numel(), dtype=torch.uint8, device=payload.device)
        self.dist.all_gather_into_tensor(out, payload)
        return out.reshape(self.world, -1)


# af_dtype code -> numpy dtype (b8 as int8), mirrors arrayfire_b200._NP without importing the library
_NP_OF = {0: np.float32, 1: np.complex64, 2: np.float64, 3: np.complex128, 4: np.int8, 5: np.int32, 6: np.uint32,
          7: np.uint8, 8: np.int64, 9: np.uint64, 10: np.int16, 11: np.uint16, 12: np.float16, 13: np.int8}
_B8 = 4
_BINOP_TO_RED = {0: OP_SUM, 1: OP_PRODUCT, 2: OP_MIN, 3: OP_MAX}
SUMMARY_BYTES = 32   # layout of afx_scan_shard_local's summary, include/afb200.h
PAYLOAD_BYTES = 64


class LocalGpu:
    """shard-local primitives on this rank's GPU: the afx_* entry points of libafb200 (include/afb200.h).
    Every method leaves its O(1) partial in a uint8 torch tensor on the device, enqueued on the library's
    stream, so the NCCL all-gather that follows needs no host synchronisation."""

    def __init__(self):
        import ctypes

        import torch

        import arrayfire_b200 as af
        self.af, self.torch, self.ct = af, torch, ctypes
        self.dev = torch.device("cuda", af.get_device())
        self.stream = torch.cuda.ExternalStream(af.get_stream(), device=self.dev)

    def enable_peer_exchange(self, dist):
        """set up the NVLink peer path of the O(1) combine (kernels/peer.cu): a symmetric, peer-mapped 2 KB buffer per
        rank (torch.distributed._symmetric_memory).  Returns False -- and the NCCL all-gather stays in use -- when the
        platform cannot map peer memory."""
        self._peer = None
        self._comm = None
        try:
            import torch.distributed._symmetric_memory as symm
            world, rank = dist.get_world_size(), dist.get_rank()
            if world < 2 or world > 8:
                return False
            t = symm.empty(2048, dtype=self.torch.uint8, device=self.dev)
            t.zero_()
            hdl = symm.rendezvous(t, dist.group.WORLD)
            self.torch.cuda.synchronize()
            dist.barrier()
            ptrs = (self.ct.c_ulonglong * 8)(*([int(p) for p in hdl.buffer_ptrs] + [0] * (8 - world)))
            self._peer = {"tensor": t, "handle": hdl, "ptrs": ptrs, "rank": rank, "world": world, "epoch": 0,
                          "out": (self.ct.c_ubyte * (8 * PAYLOAD_BYTES))()}
            # a second symmetric buffer for the FUSED ops (afx_sharded_*: the exchange happens inside the compute kernels,
            # include/afb200.h); its epochs advance independently of the afx_peer_exchange buffer above
            t2 = symm.empty(4096, dtype=self.torch.uint8, device=self.dev)
            t2.zero_()
            hdl2 = symm.rendezvous(t2, dist.group.WORLD)
            self.torch.cuda.synchronize()
            dist.barrier()
            ptrs2 = (self.ct.c_ulonglong * 8)(*([int(p) for p in hdl2.buffer_ptrs] + [0] * (8 - world)))
            comm = self.ct.c_void_p()
            self.af._check(self.af.lib.afx_comm_create(self.ct.byref(comm), self.ct.c_int(rank), self.ct.c_int(world), ptrs2))
            self._comm = comm
            self._comm_keep = (t2, hdl2, ptrs2)
            return True
        except Exception as ex:  # noqa: BLE001 -- any failure here only means "keep using NCCL"
            import sys
            print("[afb200] peer exchange unavailable, using NCCL all-gather: %r" % (ex,), file=sys.stderr)
            self._peer = None
            self._comm = None
            return False

    # ---- fused ops: ONE C call per sharded op, the cross-GPU combine happens inside the kernels ----------------------
    def fused_reduce_all(self, a, op):
        af, ct = self.af, self.ct
        re, im = ct.c_double(), ct.c_double()
        af._check(af.lib.afx_sharded_reduce_all(ct.byref(re), ct.byref(im), self._comm, a.handle, ct.c_int(op), ct.c_bool(False),
                                                ct.c_double(0.0)))
        return complex(re.value, im.value) if a.dtype in (af.c32, af.c64) else re.value

    def fused_ireduce_all(self, a, op, base):
        af, ct = self.af, self.ct
        re, im, idx = ct.c_double(), ct.c_double(), ct.c_uint()
        af._check(af.lib.afx_sharded_ireduce_all(ct.byref(re), ct.byref(im), ct.byref(idx), self._comm, a.handle, ct.c_int(op),
                                                 ct.c_uint(base)))
        v = complex(re.value, im.value) if a.dtype in (af.c32, af.c64) else re.value
        return v, idx.value

    def fused_scan(self, key, a, binop, inclusive):
        af, ct = self.af, self.ct
        h = ct.c_void_p()
        af._check(af.lib.afx_sharded_scan(ct.byref(h), self._comm, key.handle if key is not None else None, a.handle,
                                          ct.c_int(binop), ct.c_bool(inclusive)))
        return af.Array(h)

    def fused_where(self, a, base):
        af, ct = self.af, self.ct
        h, off, tot = ct.c_void_p(), ct.c_longlong(), ct.c_longlong()
        af._check(af.lib.afx_sharded_where(ct.byref(h), ct.byref(off), ct.byref(tot), self._comm, a.handle, ct.c_uint(base)))
        return af.Array(h), off.value, tot.value

    def peer_gather(self, buf):
        """[world, PAYLOAD_BYTES] numpy array of every rank's 64-byte record, exchanged by ONE kernel over peer memory"""
        p = self._peer
        p["epoch"] += 1
        af, ct = self.af, self.ct
        af._check(af.lib.afx_peer_exchange(p["out"], ct.c_void_p(buf.data_ptr()), p["ptrs"], ct.c_int(p["rank"]),
                                           ct.c_int(p["world"]), ct.c_uint(p["epoch"])))
        return np.frombuffer(p["out"], dtype=np.uint8, count=p["world"] * PAYLOAD_BYTES).reshape(p["world"], PAYLOAD_BYTES).copy()

    def payload(self):
        """the per-op 64-byte partial record.  ONE persistent device buffer: every op overwrites the fields it
        reads back, and ops on the library's stream are ordered, so no per-op allocation / memset is needed."""
        if getattr(self, "_payload", None) is None:
            with self.on_stream():     # every op that touches the payload runs on the LIBRARY's stream
                self._payload = self.torch.zeros(PAYLOAD_BYTES, dtype=self.torch.uint8, device=self.dev)
        return self._payload

    def put_base(self, buf, base):
        """buf[48:56] <- this shard's global offset (so that no second exchange is needed to learn the other ranks' bases)"""
        cache = self.__dict__.setdefault("_base_t", {})
        if base not in cache:   # purely local: no collective hides behind this cache
            cache[base] = self.torch.frombuffer(bytearray(np.uint64(base).tobytes()), dtype=self.torch.uint8).to(self.dev)
        with self.on_stream():
            buf[48:56] = cache[base]

    def to_host(self, gathered):
        """[world, PAYLOAD_BYTES] device tensor -> numpy, through a persistent pinned landing zone (one async copy +
        one stream synchronisation; `.cpu()` would stage through pageable memory)"""
        n = gathered.numel()
        if getattr(self, "_pinned", None) is None or self._pinned.numel() < n:
            self._pinned = self.torch.empty(max(n, 8 * PAYLOAD_BYTES), dtype=self.torch.uint8, pin_memory=True)
        dst = self._pinned[:n]
        dst.copy_(gathered.reshape(-1), non_blocking=True)
        self.stream.synchronize()
        return dst.numpy().reshape(gathered.shape[0], -1).copy()

    def on_stream(self):
        return self.torch.cuda.stream(self.stream)

    def elements(self, a):
        return a.elements

    def dtype(self, a):
        return a.dtype

    def reduce_partial(self, buf, a, op):
        """buf[0:16] <- reduction of the shard; returns the af_dtype code of the partial"""
        af, ct = self.af, self.ct
        t = ct.c_int()
        af._check(af.lib.afx_reduce_all_partial(ct.c_void_p(buf.data_ptr()), ct.byref(t), a.handle, ct.c_int(op),
                                                ct.c_bool(False), ct.c_double(0.0)))
        return t.value

    def ireduce_partial(self, buf, a, op):
        """buf[0:16] <- value, buf[16:20] <- shard-local u32 index"""
        af, ct = self.af, self.ct
        af._check(af.lib.afx_ireduce_all_partial(ct.c_void_p(buf.data_ptr()), ct.c_void_p(buf.data_ptr() + 16), a.handle,
                                                 ct.c_int(op)))

    def scan_with_carry(self, a, binop, inclusive, carry_bytes):
        """scan of the shard seeded with the running value of the lower ranks (host bytes or None)"""
        af, ct = self.af, self.ct
        h = ct.c_void_p()
        carry_ptr = None
        if carry_bytes is not None:
            # zero-copy: the kernel's walkers read the carry once, straight from pinned host memory.  The slot may be
            # rewritten by the next sharded scan only after that scan's own gather, which waits (stream order) for this
            # kernel -- so a single slot suffices.
            if getattr(self, "_carry_pinned", None) is None:
                self._carry_pinned = self.torch.zeros(64, dtype=self.torch.uint8, pin_memory=True)
            self._carry_pinned[:len(carry_bytes)] = self.torch.frombuffer(bytearray(carry_bytes), dtype=self.torch.uint8)
            carry_ptr = ct.c_void_p(self._carry_pinned.data_ptr())
        af._check(af.lib.afx_scan_shard(ct.byref(h), None, a.handle, ct.c_int(binop), ct.c_bool(inclusive), carry_ptr, None))
        return af.Array(h)

    def scan_with_device_carry(self, a, binop, inclusive, gathered, part_code, out_code, rank):
        """NCCL fall-back only (no peer memory): the carry-in is folded on the HOST from the gathered totals (the caller's
        path when this returns None).  A device-side fold belongs to the fused kernels (afx_sharded_scan); round 1 had an
        eager-torch fold here, removed so that no library-external kernel sits on the data path."""
        return None

    def scan_by_key_local(self, buf, key, a, binop, inclusive):
        """carry-less segmented scan; buf[0:32] <- summary, buf[32:40] <- first key, buf[40:48] <- last key"""
        af, ct = self.af, self.ct
        h = ct.c_void_p()
        af._check(af.lib.afx_scan_shard_local(ct.byref(h), ct.c_void_p(buf.data_ptr()), key.handle, a.handle, ct.c_int(binop),
                                              ct.c_bool(inclusive)))
        kb = af._SIZE[key.dtype]
        kv = self.bytes_view(key)
        with self.on_stream():         # ordered after the scan kernel and before the gather, both on the library's stream
            buf[32:32 + kb] = kv[:kb]
            buf[40:40 + kb] = kv[-kb:]
        return af.Array(h)

    def scan_fixup(self, out, carry_bytes, count, binop):
        af, ct = self.af, self.ct
        raw = (ct.c_ubyte * 16)(*bytearray(carry_bytes.ljust(16, b"\0")))
        af._check(af.lib.afx_scan_shard_fixup(out.handle, raw, ct.c_longlong(count), ct.c_int(binop)))

    def where_shard(self, a, base):
        af, ct = self.af, self.ct
        h, n = ct.c_void_p(), ct.c_longlong()
        af._check(af.lib.afx_where_shard(ct.byref(h), ct.byref(n), a.handle, ct.c_uint(base)))
        return af.Array(h), n.value

    def where_launch(self, a, base):
        """enqueue the shard's compaction; returns (opaque handle, 64-byte uint8 device view whose first 8 bytes
        will hold the hit count) -- nothing is synchronised"""
        af, ct = self.af, self.ct
        h, p = ct.c_void_p(), ct.c_void_p()
        af._check(af.lib.afx_where_shard_launch(ct.byref(h), ct.byref(p), a.handle, ct.c_uint(base)))
        if getattr(self, "_slot_ptr", None) != p.value:
            ptr = p.value

            class _Holder:
                __cuda_array_interface__ = {"shape": (PAYLOAD_BYTES,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                            "strides": None}
            self._slot = self.torch.as_tensor(_Holder(), device=self.dev)
            self._slot_ptr = ptr
        return h, self._slot

    def where_finish(self, h, count):
        af, ct = self.af, self.ct
        af._check(af.lib.afx_where_shard_finish(ct.byref(h), ct.c_longlong(count)))
        return af.Array(h)

    def bytes_view(self, a):
        """uint8 torch view of a linear array's device memory (no copy)"""
        ptr = a.device_ptr()
        a.unlock()
        n = a.elements * self.af._SIZE[a.dtype]

        class _Holder:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
        return self.torch.as_tensor(_Holder(), device=self.dev)

    def host_scalar_payload(self, value_bytes):
        buf = self.torch.zeros(PAYLOAD_BYTES, dtype=self.torch.uint8)
        buf[:len(value_bytes)] = self.torch.frombuffer(bytearray(value_bytes), dtype=self.torch.uint8)
        return buf.to(self.dev)


class ShardedOps:
    """Whole-array ops over an array sharded across the ranks of a process group.  `local` supplies the
    shard-local primitives (LocalGpu in production; the CPU tests plug in an oracle-backed stand-in and the
    gloo backend, tests/test_sharded_gloo.py) -- the packing, the all-gather and the rank-ordered combine
    below are the same code in both cases."""

    def __init__(self, dist=None, local=None):
        self.local = local if local is not None else LocalGpu()
        self.group = Group(dist)
        self.rank, self.world = self.group.rank, self.group.world
        import os
        if (dist is not None and self.world > 1 and hasattr(self.local, "enable_peer_exchange")
                and os.environ.get("AFB200_NO_PEER", "0") != "1"):
            ok = self.local.enable_peer_exchange(dist)
            # every rank must take the same path: fall back together
            import torch
            flag = torch.tensor([1 if ok else 0], device=self.local.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self.local._peer = None
                self.local._comm = None
            if os.environ.get("AFB200_NO_FUSED", "0") == "1":   # A-B switch: round-1 path (separate exchange kernel + host fold)
                self.local._comm = None

    # -- helpers -----------------------------------------------------------------------------------
    def _gather(self, buf):
        """[world, PAYLOAD_BYTES] uint8 numpy array of every rank's payload (one collective, one D2H)"""
        if getattr(self.local, "_peer", None) is not None:
            return self.local.peer_gather(buf)
        with self.local.on_stream():
            g = self.group.all_gather_bytes(buf)
            if hasattr(self.local, "to_host"):
                return self.local.to_host(g)
            return g.cpu().numpy()

    # -- reductions ----------------------------------------------------------------------------------
    def reduce_all(self, op, a):
        """af_<op>_all over the global array; the scalar comes back as the promoted numpy type"""
        if getattr(self.local, "_comm", None) is not None:
            return self.local.fused_reduce_all(a, _AFX_OP[op])
        buf = self.local.payload()
        code = self.local.reduce_partial(buf, a, _AFX_OP[op])
        g = self._gather(buf)
        dt = np.dtype(_NP_OF[code])
        parts = np.array([g[r, :dt.itemsize].view(dt)[0] for r in range(self.world)], dtype=dt)
        return fold_partials(op, parts)

    def ireduce_all(self, op, a, base):
        """(value, GLOBAL u32 index) of af_imin_all / af_imax_all; `base` is this shard's global offset"""
        if getattr(self.local, "_comm", None) is not None:
            return self.local.fused_ireduce_all(a, _AFX_OP[op], int(base))
        buf = self.local.payload()
        self.local.ireduce_partial(buf, a, _AFX_OP[op])
        self.local.put_base(buf, int(base))      # every rank's base travels in its own record
        g = self._gather(buf)
        bases = [int(g[r, 48:56].view(np.uint64)[0]) for r in range(self.world)]
        code = self.local.dtype(a)
        dt = np.dtype(_NP_OF[code])
        vals = [g[r, :dt.itemsize].view(dt)[0] for r in range(self.world)]
        gidx = [int(g[r, 16:20].view(np.uint32)[0]) + bases[r] for r in range(self.world)]
        v, i = fold_arg(op, vals, gidx, is_b8=(code == _B8))
        return v, i & 0xffffffff

    # -- issue / collect -----------------------------------------------------------------------------------
    # The *_async forms enqueue the op (kernel + in-kernel exchange) and return at once; .result() collects the O(1) host
    # result later.  A step that issues all its ops first and collects afterwards keeps every GPU busy: no host round trip
    # between two ops, and the skew between ranks is absorbed by queued work instead of idling the early rank.
    class _Done:
        def __init__(self, value):
            self.value = value

        def result(self):
            return self.value

    def reduce_all_async(self, op, a):
        comm = getattr(self.local, "_comm", None)
        if comm is None:
            return self._Done(self.reduce_all(op, a))
        return self.local.af.reduce_all_async(a, _AFX_OP[op], comm)

    def ireduce_all_async(self, op, a, base):
        comm = getattr(self.local, "_comm", None)
        if comm is None:
            return self._Done(self.ireduce_all(op, a, base))
        return self.local.af.ireduce_all_async(a, _AFX_OP[op], comm, int(base))

    def where_async(self, a, base):
        comm = getattr(self.local, "_comm", None)
        if comm is None:
            return self._Done(self.where(a, base))
        return self.local.af.where_async(a, comm, int(base))

    # -- scans -----------------------------------------------------------------------------------------
    def scan(self, a, binop, inclusive):
        """af_scan along the flattened axis: shard totals are exchanged first, then every shard is scanned
        once with its carry-in (the second read of a shard that fits the 126 MB L2 is served from L2)."""
        if getattr(self.local, "_comm", None) is not None and self.local.dtype(a) != _B8:
            return self.local.fused_scan(None, a, binop, inclusive)
        buf = self.local.payload()
        red = _BINOP_TO_RED[binop]
        code = self.local.reduce_partial(buf, a, _AFX_OP[red])
        if hasattr(self.local, "scan_with_device_carry") and getattr(self.local, "_peer", None) is None:
            with self.local.on_stream():
                gd = self.group.all_gather_bytes(buf)
                out = self.local.scan_with_device_carry(a, binop, inclusive, gd, code, scan_out_type(self.local.dtype(a)), self.rank)
            if out is not None:
                return out
        g = self._gather(buf)
        dt = np.dtype(_NP_OF[code])
        out_dt = np.dtype(_NP_OF[scan_out_type(self.local.dtype(a))])
        totals = np.array([g[r, :dt.itemsize].view(dt)[0] for r in range(self.world)], dtype=dt).astype(out_dt)
        carry = exclusive_fold(binop, totals, self.rank)
        return self.local.scan_with_carry(a, binop, inclusive, None if carry is None else np.asarray(carry, out_dt).tobytes())

    def scan_by_key(self, key, a, binop, inclusive):
        """af_scan_by_key along the flattened axis: local carry-less scan, one all-gather of the shard
        summaries, fix-up of the first segment only."""
        if getattr(self.local, "_comm", None) is not None:
            return self.local.fused_scan(key, a, binop, inclusive)
        buf = self.local.payload()
        out = self.local.scan_by_key_local(buf, key, a, binop, inclusive)
        g = self._gather(buf)
        out_dt = np.dtype(_NP_OF[sbk_out_type(self.local.dtype(a))])
        kdt = np.dtype(_NP_OF[self.local.dtype(key)])
        tails = [g[r, :out_dt.itemsize].view(out_dt)[0] for r in range(self.world)]
        single = [int(g[r, 16:20].view(np.uint32)[0]) == 0 for r in range(self.world)]
        first_keys = [g[r, 32:32 + kdt.itemsize].view(kdt)[0] for r in range(self.world)]
        last_keys = [g[r, 40:40 + kdt.itemsize].view(kdt)[0] for r in range(self.world)]
        carry = segmented_carry(binop, np.asarray(tails, out_dt), single, first_keys, last_keys, self.rank)
        if carry is not None:
            first_head = int(g[self.rank, 24:32].view(np.uint64)[0])
            count = min(first_head, self.local.elements(a))
            self.local.scan_fixup(out, np.asarray(carry, out_dt).tobytes(), count, binop)
        return out

    # -- compaction ------------------------------------------------------------------------------------
    def where(self, a, base):
        """af_where: returns (this shard's indices [global numbering], offset of the slice in the global
        output, global hit count).  The output stays sharded."""
        if getattr(self.local, "_comm", None) is not None:
            return self.local.fused_where(a, base)
        if hasattr(self.local, "where_launch"):
            # kernel, all-gather of the device-resident counts and ONE read-back / synchronisation
            h, slot = self.local.where_launch(a, base)
            g = self._gather(slot)
            counts = [int(g[r, :8].view(np.int64)[0]) for r in range(self.world)]
            idx = self.local.where_finish(h, counts[self.rank])
        else:
            idx, n = self.local.where_shard(a, base)
            g = self._gather(self.local.host_scalar_payload(np.int64(n).tobytes()))
            counts = [int(g[r, :8].view(np.int64)[0]) for r in range(self.world)]
        off, total = where_offsets(counts)
        return idx, int(off[self.rank]), total


_AFX_OP = {OP_SUM: 0, OP_PRODUCT: 1, OP_MIN: 2, OP_MAX: 3, OP_COUNT: 4, OP_ALL: 5, OP_ANY: 6}


def scan_out_type(code):
    """output dtype of af_scan / af_accum (/root/reference/src/api/c/scan.cpp:133-149)"""
    return {10: 5, 13: 5, 11: 6, 7: 6, 4: 6}.get(code, code)


sbk_out_type = scan_out_type  # af_scan_by_key casts narrow integers the same way (scan.cpp:244-276)
