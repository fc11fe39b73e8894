"""Cross-rank reduction of `stat` partials: one rank per GPU, records sharded by contiguous
byte ranges in rank order; the only exchange of the whole hot path.

  1. ranks all-gather their Pair keys (ref_name, ref_size, query_name, query_size) -- a few kB;
  2. every rank derives the same global pair list (first-seen order over ranks 0..N-1 = file order);
  3. ONE all-reduce(sum) of the dense [n_pairs x 11] u64 counter table and one all-reduce(min) of the
     [n_pairs x 2] start coordinates (NCCL when the tensors live on the GPU, gloo in the CPU tests);
  4. the order-dependent f32 inv_size needs the '-'-strand records in file order: the (pair, numerator)
     lists are all-gathered and concatenated in rank order (SURVEY hard part 5 / Q6).
torch.distributed is plumbing here; all compute stays behind the C ABI.
"""
import struct

import numpy as np

from .api import StatAcc

_I64_MAX = np.iinfo(np.int64).max


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


def allreduce_stat(acc: StatAcc, dist, device=None) -> StatAcc:
    """merged-mode stat accumulators of all ranks -> the global accumulator (identical on every rank)"""
    import torch
    world = dist.get_world_size()
    each, keys, vals, invs, inv_pair, inv_num = parse_blob(acc.export())
    if each:
        raise ValueError("--each rows are concatenated in rank order, not reduced")
    gathered = [None] * world
    dist.all_gather_object(gathered, keys)
    gkeys, index = [], {}
    for ks in gathered:  # rank order == file order
        for k in ks:
            if k not in index:
                index[k] = len(gkeys)
                gkeys.append(k)
    P = len(gkeys)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    sums = torch.zeros((max(P, 1), 11), dtype=torch.int64)
    mins = torch.full((max(P, 1), 2), _I64_MAX, dtype=torch.int64)
    local = np.array([index[k] for k in keys], dtype=np.int64)
    if len(keys):
        sums[local] = torch.from_numpy(vals[:, :11].astype(np.int64))
        mins[local] = torch.from_numpy(vals[:, 11:13].astype(np.int64))
    sums, mins = sums.to(dev), mins.to(dev)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)  # the single counter all-reduce
    dist.all_reduce(mins, op=dist.ReduceOp.MIN)
    sums, mins = sums.cpu().numpy(), mins.cpu().numpy()
    inv_lists = [None] * world
    dist.all_gather_object(inv_lists, (local[inv_pair].astype(np.uint32) if len(inv_pair) else np.zeros(0, np.uint32),
                                       inv_num))
    g_pair = np.concatenate([p for p, _ in inv_lists]) if world else np.zeros(0, np.uint32)
    g_num = np.concatenate([n for _, n in inv_lists]) if world else np.zeros(0, np.uint64)
    gvals = np.zeros((P, 13), dtype=np.uint64)
    if P:
        gvals[:, :11] = sums[:P].astype(np.uint64)
        gvals[:, 11:13] = mins[:P].astype(np.uint64)
    out = StatAcc(False)
    out.import_(build_blob(0, gkeys, gvals, np.zeros(P, np.float32), g_pair, g_num))
    return out


def shard_range(data, rank: int, world: int, kind: str = "maf"):
    """Byte range [lo, hi) of `data` that rank `rank` of `world` processes: contiguous ranges snapped to
    record boundaries, concatenating the ranks' outputs in rank order reproduces the single-process output.
    MAF: a range starts at a line that does not begin with 's', so s-line groups are never split (maf.rs:395-413); only
    rank 0's range holds the file's first line, which the reader swallows (pass WGA_F_FIRST_SLAB there only).
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
