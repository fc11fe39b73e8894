"""world_size-2 gloo test of the only exchange step of the path: the stat counter all-reduce
(wgatools_b200/dist.py).  The per-rank partial tables are built from the oracle's per-record counters (no GPU
here); the merged TSV must equal the oracle's stat of the whole file, including the order-dependent f32
inv_size."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _partial_blob(text: bytes):
    """emulates what wga_maf_stat leaves in a StatAcc for one shard (merged mode)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc
    from wgatools_b200 import dist as wd
    lines = text.split(b"\n")[1:]
    recs, cur = [], []
    for l in lines:
        if l.startswith(b"s"):
            cur.append(l.split())
        elif cur:
            recs.append(cur)
            cur = []
    if cur:
        recs.append(cur)
    cnt = orc.maf_record_counts(text, len(recs))
    keys, index, vals, inv_pair, inv_num = [], {}, [], [], []
    for r, c in zip(recs, cnt):
        t, q = r[0], r[1]
        rs, qs = int(t[5]), int(q[5])
        key = (t[1], rs, q[1], qs)
        neg = q[4] == b"-"
        qstart = qs - int(q[2]) - int(q[3]) if neg else int(q[2])
        if key not in index:
            index[key] = len(keys)
            keys.append(key)
            vals.append([0] * 11 + [rs, qs])
        v = vals[index[key]]
        for i in range(11):
            v[i] += c[i]
        v[11] = min(v[11], int(t[2]))
        v[12] = min(v[12], qstart)
        if neg:
            inv_pair.append(index[key])
            inv_num.append(2 * (c[0] + c[1]) + c[5] + c[9] + c[3] + c[7])
    return wd.build_blob(0, keys, np.array(vals, dtype=np.uint64).reshape(len(keys), 13), np.zeros(len(keys), np.float32),
                         np.array(inv_pair, dtype=np.uint32), np.array(inv_num, dtype=np.uint64))


def _worker(rank, world, port, shards, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import wgatools_b200 as w
    from wgatools_b200 import dist as wd
    acc = w.StatAcc(False)
    acc.import_(_partial_blob(shards[rank]))
    merged = wd.allreduce_stat(acc, dist)
    q.put((rank, merged.write()))
    dist.barrier()
    dist.destroy_process_group()


def test_stat_allreduce_gloo_world2():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as g
    g.build()
    import bench_support as bs
    import oracle_lib as orc
    text = bs.gen_maf(bs.maf_params(600, seed=77, p_neg=0.45))
    # shard at an s-line-group boundary; the second shard gets a dummy first line (only rank 0 has the header)
    cut = text.find(b"\n\n", len(text) // 2) + 2
    shards = [text[:cut], b"#shard\n" + text[cut:]]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, shards, q)) for r in range(2)]
    for p in ps:
        p.start()
    outs = dict(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = orc.stat_maf(text)
    assert outs[0] == want and outs[1] == want


def _blob_worker(rank, world, port, blobs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import wgatools_b200 as w
    from wgatools_b200 import dist as wd
    comm = wd.make_comm(dist)
    out = []
    for case in blobs:
        acc = w.StatAcc(False)
        acc.import_(case[rank])
        comm.allreduce_stat(acc)
        out.append(acc.write())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_stat_allreduce_dictionary_paths_world3():
    """pair-dictionary edge cases of wga_stat_allreduce: identical key sets in different order (the digest shortcut),
    a rank with no records, pairs only a later rank holds, starts >= 2^63 under the unsigned min.  Expected = one
    accumulator fed with the ranks' partials in rank order (the single-process merge of the same host code; the
    world-2 test above pins the reduction against the oracle)."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    import wgatools_b200 as w
    from wgatools_b200 import dist as wd
    rng = np.random.default_rng(11)

    def part(names, n_inv, big=False, num_hi=1 << 33, hot=None):
        keys = [(f"ref.{a}".encode(), 1000 + a, f"qry.{b}".encode(), 2000 + b) for a, b in names]
        vals = rng.integers(0, 1 << 40, size=(len(keys), 13), dtype=np.uint64)
        vals[:, 11] = rng.integers(0, 900, size=len(keys))
        vals[:, 12] = rng.integers(0, 1900, size=len(keys))
        if big and len(keys):
            vals[0, 11] = np.uint64((1 << 63) + 5)
        inv_pair = rng.integers(0, max(len(keys), 1), size=n_inv if keys else 0, dtype=np.uint32)
        inv_num = rng.integers(1, num_hi, size=len(inv_pair), dtype=np.uint64) | np.uint64(1)   # odd: halves are inexact in f32
        if hot is not None and len(inv_pair):   # one pair far beyond 2^24 among exact ones: only it takes the rank-order ring
            sel = inv_pair == inv_pair[hot % len(inv_pair)]
            inv_num[sel] = rng.integers(1 << 23, 1 << 26, size=int(sel.sum()), dtype=np.uint64) | np.uint64(1)
        return wd.build_blob(0, keys, vals, np.zeros(len(keys), np.float32), inv_pair, inv_num)

    base = [(a, b) for a in range(12) for b in range(9)]
    perm1 = [base[i] for i in rng.permutation(len(base))]
    perm2 = [base[i] for i in rng.permutation(len(base))]
    cases = [
        [part(base, 400), part(perm1, 300), part(perm2, 500)],                      # same set, three orders
        [part(base[:40], 100), part([], 0), part(base[20:], 200)],                  # empty rank, late pairs
        [part([], 0), part(perm1[:30], 50, big=True), part(perm1[10:50], 60)],      # rank 0 empty
        [part([], 0), part([], 0), part([], 0)],                                    # nothing anywhere
        # every pair's numerators stay below 2^24 in total: the order-free integer column settles inv_size, no ring
        [part(base, 400, num_hi=9000), part(perm1, 300, num_hi=9000), part(perm2, 500, num_hi=9000)],
        [part(base, 400, num_hi=9000, hot=3), part(perm1[:50], 300, num_hi=9000), part(perm2, 500, num_hi=9000, hot=7)],  # mixed
    ]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    ps = [ctx.Process(target=_blob_worker, args=(r, 3, port, cases, q)) for r in range(3)]
    for p in ps:
        p.start()
    outs = dict(q.get(timeout=180) for _ in range(3))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for ci, case in enumerate(cases):
        one = w.StatAcc(False)
        for blob in case:
            one.import_(blob)
        want = one.write()
        for r in range(3):
            assert outs[r][ci] == want, (ci, r)
        assert ci == 3 or want.count(b"\n") > 10


def test_stat_allreduce_large_tables_world2():
    """the threaded host paths of wga_stat_allreduce (>= 16384 pairs, >= 200000 '-' records): uniform key sets in two orders
    (threaded counter table and rebuild), and a non-uniform pair of ranks (sequential rebuild from the name blobs); numerators
    small enough for the order-free integer column in the first case, a few pairs over 2^24 in the second (compact ring)."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    import wgatools_b200 as w
    from wgatools_b200 import dist as wd
    rng = np.random.default_rng(23)
    P = 20000

    def part(ids, n_inv, num_hi, hot=0):
        keys = [(b"g%d#chr%d" % (i % 40, i // 40 % 25), 1000 + i % 7, b"h%d#c%d" % (i // 1000, i % 1000), 2000 + i % 5) for i in ids]
        vals = rng.integers(0, 1 << 30, size=(len(keys), 13), dtype=np.uint64)
        inv_pair = rng.integers(0, len(keys), size=n_inv, dtype=np.uint32)
        inv_num = rng.integers(1, num_hi, size=n_inv, dtype=np.uint64) | np.uint64(1)
        if hot:
            sel = inv_pair < hot
            inv_num[sel] = rng.integers(1 << 22, 1 << 25, size=int(sel.sum()), dtype=np.uint64) | np.uint64(1)
        return wd.build_blob(0, keys, vals, np.zeros(len(keys), np.float32), inv_pair, inv_num)

    ids = np.arange(P)
    cases = [
        [part(ids, 250000, 9000), part(rng.permutation(P), 250000, 9000)],
        [part(ids[:17000], 220000, 9000, hot=5), part(rng.permutation(P)[:18000], 230000, 9000, hot=9)],
    ]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + os.getpid() % 2000
    ps = [ctx.Process(target=_blob_worker, args=(r, 2, port, cases, q)) for r in range(2)]
    for p in ps:
        p.start()
    outs = dict(q.get(timeout=300) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for ci, case in enumerate(cases):
        one = w.StatAcc(False)
        for blob in case:
            one.import_(blob)
        want = one.write()
        assert outs[0][ci] == want and outs[1][ci] == want, ci
        assert want.count(b"\n") > 17000


def _shard_worker(rank, world, port, text, paf, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as orc
    from wgatools_b200 import dist as wd
    lo, hi = wd.shard_range(text, rank, world, "maf")
    mine = text[lo:hi] if rank == 0 else b"#not the file header\n" + text[lo:hi]   # only rank 0 holds the real first line
    out = orc.maf2paf(mine)                     # the per-rank command (the CPU oracle stands in for the GPU here)
    plo, phi = wd.shard_range(paf, rank, world, "paf")
    outs = [None] * world
    dist.all_gather_object(outs, (lo, hi, out, plo, phi))
    if rank == 0:
        q.put(outs)
    dist.barrier()
    dist.destroy_process_group()


def test_record_sharding_gloo_world2():
    """byte ranges snapped to record boundaries; outputs concatenated in rank order == one process"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as g
    g.build()
    import bench_support as bs
    import oracle_lib as orc
    from wgatools_b200 import dist as wd
    text = bs.gen_maf(bs.maf_params(300, seed=5, p_neg=0.2))
    paf = bs.gen_paf(bs.paf_params(200, seed=6))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    ps = [ctx.Process(target=_shard_worker, args=(r, 2, port, text, paf, q)) for r in range(2)]
    for p in ps:
        p.start()
    outs = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert outs[0][0] == 0 and outs[0][1] == outs[1][0] and outs[1][1] == len(text)
    assert 0 < outs[0][1] < len(text) and text[outs[0][1]:outs[0][1] + 1] != b"s"
    assert outs[0][2] + outs[1][2] == orc.maf2paf(text)
    assert outs[0][3] == 0 and outs[0][4] == outs[1][3] and outs[1][4] == len(paf) and paf[outs[0][4] - 1:outs[0][4]] == b"\n"
    # more ranks than records / odd worlds still tile the file
    for world in (1, 3, 7, 50):
        r = [wd.shard_range(text, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == len(text) and all(r[k][1] == r[k + 1][0] for k in range(world - 1))
