#!/usr/bin/env python
"""quick device-resident timing of paf2chain / plot data on the config-4 shape (2 M records): usage tools/time_paf2chain.py [records]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
g.build()
import bench_support as bs
import wgatools_b200 as w
from wgatools_b200._lib import Span
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
paf = bs.gen_paf(bs.paf_params(n, seed=20260928))
d = torch.frombuffer(bytearray(paf), dtype=torch.uint8).cuda()
ctx = w.Context(0)
L = ctx.L
sp = Span()
ptr = ctypes.c_void_p(d.data_ptr())
def timed(fn, k=5):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3
ms = timed(lambda: ctx._check(L.wga_paf2chain(ctx.h, ptr, len(paf), 1, 1, 0, ctypes.byref(sp), None)))
print(f"paf2chain {ms:.2f} ms  ({len(paf) / ms / 1e6:.1f} GB/s in, {sp.len / 1e9:.2f} GB out)")
ms = timed(lambda: ctx._check(L.wga_paf_plotdata(ctx.h, ptr, len(paf), 1, 1, 50, ctypes.byref(sp))))
print(f"plotdata  {ms:.2f} ms  ({len(paf) / ms / 1e6:.1f} GB/s in)")
import oracle_lib as orc
smp = paf[:paf.find(b"\n", 30_000_000) + 1]
ctx._check(L.wga_paf2chain(ctx.h, ctypes.c_char_p(smp), len(smp), 0, 1, 0, ctypes.byref(sp), None))
print("paf2chain parity on a 30 MB prefix:", ctypes.string_at(sp.data, sp.len) == orc.paf2chain(smp))
