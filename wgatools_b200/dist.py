"""Record sharding and the one exchange step of the hot path: one rank per GPU, the input cut into
contiguous byte ranges in rank order, outputs concatenated in rank order; `stat` is the only command whose partial
results meet.

The reduction itself is C++ behind the C ABI (include/wga_b200.h `wga_stat_allreduce`, wgatools_b200/csrc/comm.cu):
pair dictionary by all-gather of key hashes, ONE ncclAllReduce(sum) of the [n_pairs x 11] u64 counters, one
ncclAllReduce(min) of the [n_pairs x 2] starts, and the order-dependent f32 inv_size carried rank to rank
(src/tools/stat.rs:167-223).  torch.distributed is plumbing here: it launches the ranks and carries the
ncclUniqueId; on CPU (gloo, the tests) it also provides the three host-buffer primitives of `wga_transport`.
"""
import struct

import numpy as np

from .api import Comm, StatAcc


def parse_blob(blob: bytes):
    """inverse of wga_stat_acc_export (wgatools_b200/csrc/stat_acc.hpp export_blob)"""
    off = 0
    each, n = struct.unpack_from("<QQ", blob, off)
    off += 16
    keys, vals, invs = [], np.zeros((n, 13), dtype=np.uint64), np.zeros(n, dtype=np.float32)
    for i in range(n):
        lr, lq = struct.unpack_from("<II", blob, off)
        off += 8
        rn = blob[off:off + lr]
        qn = blob[off + lr:off + lr + lq]
        off += lr + lq
        rs, qs = struct.unpack_from("<QQ", blob, off)
        off += 16
        vals[i] = np.frombuffer(blob, dtype="<u8", count=13, offset=off)
        off += 104
        invs[i] = struct.unpack_from("<f", blob, off)[0]
        off += 8
        keys.append((rn, rs, qn, qs))
    (ni,) = struct.unpack_from("<Q", blob, off)
    off += 8
    inv_pair = np.frombuffer(blob, dtype="<u4", count=ni, offset=off).copy()
    off += 4 * ni
    inv_num = np.frombuffer(blob, dtype="<u8", count=ni, offset=off).copy()
    return int(each), keys, vals, invs, inv_pair, inv_num


def build_blob(each, keys, vals, invs, inv_pair, inv_num) -> bytes:
    out = [struct.pack("<QQ", int(each), len(keys))]
    for i, (rn, rs, qn, qs) in enumerate(keys):
        out.append(struct.pack("<II", len(rn), len(qn)))
        out.append(rn)
        out.append(qn)
        out.append(struct.pack("<QQ", rs, qs))
        out.append(np.asarray(vals[i], dtype="<u8").tobytes())
        out.append(struct.pack("<fI", float(invs[i]), 0))
    out.append(struct.pack("<Q", len(inv_pair)))
    out.append(np.asarray(inv_pair, dtype="<u4").tobytes())
    out.append(np.asarray(inv_num, dtype="<u8").tobytes())
    return b"".join(out)


def make_comm(dist, ctx=None) -> Comm:
    """NCCL communicator inside the C-ABI library when a Context (GPU) is given and the process group is nccl;
    otherwise the host-buffer transport over torch.distributed (gloo)."""
    if ctx is not None and dist.get_backend() == "nccl":
        return Comm.nccl(ctx, dist)
    return Comm.host(dist)


def allreduce_stat(acc: StatAcc, dist=None, comm: Comm = None, ctx=None) -> StatAcc:
    """merged-mode stat accumulators of all ranks -> the accumulator of the whole file, in place, identical on every
    rank (wga_stat_allreduce).  Pass a `comm` made once with make_comm() when calling this repeatedly."""
    own = comm is None
    if own:
        comm = make_comm(dist, ctx)
    try:
        return comm.allreduce_stat(acc)
    finally:
        if own:
            comm.close()


def shard_range(data, rank: int, world: int, kind: str = "maf"):
    """Byte range [lo, hi) of `data` that rank `rank` of `world` processes: contiguous ranges snapped to
    record boundaries; concatenating the ranks' outputs in rank order reproduces the single-process output (chain ids:
    see chain_id_base).
    MAF: a range starts at a line that does not begin with 's', so s-line groups are never split (maf.rs:395-413).  The
    file's first line -- which the reader swallows (maf.rs:25-36) -- belongs to the range that starts at byte 0 and is not
    empty: pass WGA_F_FIRST_SLAB exactly when first_slab(lo, hi) says so (with more ranks than records several ranges
    start at byte 0; all but the last of them are empty).
    PAF: a range starts right after a newline."""
    n = len(data)

    def snap(pos):
        if pos <= 0:
            return 0
        if pos >= n:
            return n
        nl = data.find(b"\n", pos - 1)           # first line start at or after pos
        if nl < 0:
            return n
        p = nl + 1
        if kind == "maf":
            while p < n and data[p:p + 1] == b"s":  # skip the rest of a split s-line group
                nl = data.find(b"\n", p)
                if nl < 0:
                    return n
                p = nl + 1
        return min(p, n)
    return snap(n * rank // world), snap(n * (rank + 1) // world)


def first_slab(lo: int, hi: int) -> bool:
    """does the byte range [lo, hi) hold the file's first line?"""
    return lo == 0 and hi > 0


def chain_id_base(n_records: int, dist) -> int:
    """maf2chain / paf2chain number the chains of the WHOLE file (converter.rs:82-88 `enumerate`): rank r passes
    id_base = number of records of ranks 0..r-1, i.e. an exclusive scan of the per-rank record counts."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor([int(n_records)], dtype=torch.int64, device=dev)
    every = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(every, mine)
    return int(sum(int(t.item()) for t in every[:rank]))
