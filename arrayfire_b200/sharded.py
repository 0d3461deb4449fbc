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
        out = torch.empty((self.world, payload.numel()), dtype=torch.uint8, device=payload.device)
        self.dist.all_gather_into_tensor(out, payload)
        return out
