"""Hardware test of the N > 1 path: one process per GPU, the input cut with dist.shard_range, every
rank runs the commands on its range through the C ABI, the stat partials meet in wga_stat_allreduce over NCCL
(communicator created inside the library).  The merged TSV must be byte-identical to the oracle's stat of the whole
file -- including the order-dependent f32 inv_size -- and the per-rank PAF / chain outputs, concatenated in rank order,
must equal the oracle's.  Skipped unless at least two GPUs are visible (`gpurun --gpus 2`)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, maf, paf, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import ctypes
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import wgatools_b200 as w
    from wgatools_b200 import dist as wd
    from wgatools_b200.api import _host_ptr
    ctx = w.Context(rank, torch.cuda.current_stream().cuda_stream)
    comm = wd.make_comm(dist, ctx)
    res = {}
    # ---- MAF: maf2paf / maf2chain per range, stat reduced
    lo, hi = wd.shard_range(maf, rank, world, "maf")
    mine = maf[lo:hi]
    keep, base = _host_ptr(mine)
    fl = 1 if wd.first_slab(lo, hi) else 0
    p, n = ctx.maf2paf_raw(ctypes.c_void_p(base), len(mine), 0, fl)
    res["paf"] = ctypes.string_at(p, n) if n else b""
    acc = w.StatAcc(False)
    ctx.maf_stat_raw(ctypes.c_void_p(base), len(mine), 0, fl, acc)
    n_local = acc.npairs()
    comm.allreduce_stat(acc)
    res["stat_maf"] = acc.write()
    res["pairs"] = (n_local, acc.npairs())
    sp, nrec = w.api.Span(), ctypes.c_uint64()
    ctx._check(ctx.L.wga_maf2chain(ctx.h, ctypes.c_void_p(base), len(mine), 0, fl, None, 0, ctypes.byref(sp), ctypes.byref(nrec)))
    id_base = wd.chain_id_base(nrec.value, dist)
    ctx._check(ctx.L.wga_maf2chain(ctx.h, ctypes.c_void_p(base), len(mine), 0, fl, None, id_base, ctypes.byref(sp), ctypes.byref(nrec)))
    res["chain"] = ctypes.string_at(sp.data, sp.len) if sp.len else b""
    # ---- PAF: stat -f paf per range, reduced (several rounds: the communicator and its staging buffers are reused)
    lo, hi = wd.shard_range(paf, rank, world, "paf")
    minep = paf[lo:hi]
    keep2, basep = _host_ptr(minep)
    for _ in range(3):
        accp = w.StatAcc(False)
        ctx.paf_stat_raw(ctypes.c_void_p(basep), len(minep), 0, 0, accp)
        comm.allreduce_stat(accp)
        res["stat_paf"] = accp.write()
    res["nccl_calls"] = ctx.L.wga_ctx_nccl_calls(ctx.h)
    q.put((rank, res))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_sharded_commands_and_stat_allreduce_nccl():
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    import bench_support as bs
    import oracle_lib as orc
    import torch.multiprocessing as mp
    maf = bs.gen_maf(bs.maf_params(3000, seed=91, p_neg=0.4))
    paf = bs.gen_paf(bs.paf_params(20000, seed=92, ops_median=40.0))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, world, port, maf, paf, q)) for r in range(world)]
    for p in ps:
        p.start()
    outs = dict(q.get(timeout=300) for _ in range(world))
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    want_stat, want_paf_stat = orc.stat_maf(maf), orc.stat_paf(paf)
    for r in range(world):
        assert outs[r]["stat_maf"] == want_stat, f"rank {r}: merged MAF stat differs from the oracle"
        assert outs[r]["stat_paf"] == want_paf_stat, f"rank {r}: merged PAF stat differs from the oracle"
        assert outs[r]["nccl_calls"] > 0
    assert b"".join(outs[r]["paf"] for r in range(world)) == orc.maf2paf(maf)
    assert b"".join(outs[r]["chain"] for r in range(world)) == orc.maf2chain(maf)
    # the dictionary really was merged: the global table has at least as many pairs as any rank's
    assert all(outs[r]["pairs"][1] >= outs[r]["pairs"][0] for r in range(world))
    assert len({outs[r]["pairs"][1] for r in range(world)}) == 1
