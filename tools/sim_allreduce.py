#!/usr/bin/env python
"""Host-side timing of wga_stat_allreduce without a GPU: W gloo processes, each with a synthetic config-4-shaped accumulator
(P pairs met in a rank-specific order, NREC '-' records), WGA_DEBUG_TIMING laps on rank 0.
usage: tools/sim_allreduce.py [world=4] [pairs=102400] [neg_records=1000000]"""
import os
import sys
import time

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, port, P, NREC, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["WGA_DEBUG_TIMING"] = "1"
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import wgatools_b200 as w
    from wgatools_b200 import dist as wd
    rng = np.random.default_rng(100 + rank)
    order = rng.permutation(P)
    keys = [(b"g%d#chr%d" % (p % 16, (p // 16) % 20 + 1), 250000000, b"g%d#chr%d" % ((p // 320) % 16, (p // 5120) % 20 + 1), 250000000)
            for p in order]
    vals = rng.integers(1, 1000, size=(P, 13)).astype(np.uint64)
    inv_pair = rng.integers(0, P, size=NREC).astype(np.uint32)
    inv_num = rng.integers(1000, 60000, size=NREC).astype(np.uint64)
    blob = wd.build_blob(0, keys, vals, np.zeros(P, np.float32), inv_pair, inv_num)
    comm = wd.make_comm(dist)
    times = []
    out = None
    for it in range(4):
        acc = w.StatAcc(False)
        acc.import_(blob)
        dist.barrier()
        t0 = time.perf_counter()
        comm.allreduce_stat(acc)
        times.append(time.perf_counter() - t0)
        t1 = time.perf_counter()
        out = acc.write()
        tw = time.perf_counter() - t1
    if rank == 0:
        print(f"world {world}: allreduce {[round(t * 1e3, 1) for t in times]} ms, TSV write {tw * 1e3:.1f} ms, {len(out)} bytes", flush=True)
    import hashlib; q.put((rank, hashlib.sha1(out).hexdigest()))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 102400
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 1000000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, W, 29621, P, N, q)) for r in range(W)]
    for p in ps:
        p.start()
    hs = [q.get() for _ in ps]
    for p in ps:
        p.join()
    assert len({h for _, h in hs}) == 1, "ranks disagree"
