import sys, os, ctypes, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc, bench_support as bs, wgatools_b200 as w
from bench_support.paf_legs import make_fasta
ctx = w.Context(0)
fa, fai = make_fasta(3, 2, 300000, seed=9)
F = w.Fasta(ctx, fa, fai)
small = (b"chain 9 g0#chr1 300000 + 2 12 g1#chr2 300000 - 3 12 5\n3 1 2\n4\n\nchain 9 g2#chr1 300000 - 0 5 g0#chr2 300000 + 10 15 6\n5\n")
for t in (small,):
    sp = w.api.Span()
    keep, base = w.api._host_ptr(t)
    rc = ctx.L.wga_chain2maf(ctx.h, ctypes.c_void_p(base), len(t), 0, 1, F.h, F.h, ctypes.byref(sp))
    print("rc", rc, "len", sp.len, ctx.L.wga_last_error(ctx.h, None))
    if sp.len: print(ctypes.string_at(sp.data, sp.len))
    print(orc.chain2maf(t, fa, fai, fa, fai))
    rc = ctx.L.wga_chain2paf(ctx.h, ctypes.c_void_p(base), len(t), 0, 1, ctypes.byref(sp))
    print("rc", rc, "len", sp.len)
