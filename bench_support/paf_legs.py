#!/usr/bin/env python
"""Second-wave measurements (scaled): `stat -f paf`, `paf2maf`, `call -f paf`, `pafpseudo`.

Not the headline workload (that is bench.py's config 2); bench.py runs these legs as `--config 4` / `--config 5` and, in
its default single-GPU run, reports them under the `side_configs` key.  One JSON object per command:
  value     = input bytes / device time with the PAF (and FASTA) resident in HBM   (CUDA events on the ctx stream)
  e2e       = same through the host-buffer C-ABI call (H2D + D2H inside)
  cpu       = the CPU port (oracle) on a bounded sample of the same input, one thread (these reference paths
              are sequential: src/tools/caller.rs:283-292, src/converter.rs:196, src/tools/stat.rs:92-99 is rayon)
Usage (GPU box): python bench_support/paf_legs.py [--records 2000000] [--genomes 4] [--chrs 5] [--chr-mbp 1.0] [--fa-records 50000]
(pafpseudo writes one target-sized row per (target, query) pair, so the FASTA-backed part is kept to a few hundred pairs)
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402


def make_fasta(n_genomes, n_chr, chr_size, seed=5):
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    width = 60
    parts, fai = [], []
    pos = 0
    for g in range(n_genomes):
        for c in range(n_chr):
            name = b"g%d#chr%d" % (g, c + 1)
            head = b">" + name + b"\n"
            seq = alpha[rng.integers(0, 4, chr_size, dtype=np.uint8)]
            full = chr_size // width
            body = np.empty((full, width + 1), dtype=np.uint8)
            body[:, :width] = seq[:full * width].reshape(full, width)
            body[:, width] = 10
            rest = seq[full * width:].tobytes()
            blob = head + body.tobytes() + (rest + b"\n" if rest else b"")
            fai.append(b"%s\t%d\t%d\t%d\t%d\n" % (name, chr_size, pos + len(head), width, width + 1))
            parts.append(blob)
            pos += len(blob)
    return b"".join(parts), b"".join(fai)


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=2_000_000, help="config-4 PAF records for stat -f paf")
    ap.add_argument("--genomes", type=int, default=4)
    ap.add_argument("--chrs", type=int, default=5)
    ap.add_argument("--chr-mbp", type=float, default=1.0)
    ap.add_argument("--fa-records", type=int, default=50_000, help="config-5 PAF records against the FASTA")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu-mb", type=float, default=60.0)
    ap.add_argument("--only", default="", help="stat | fasta : run one half only (profiling)")
    return ap.parse_args(argv)


def run(args, ctx=None, emit=None):
    """runs the legs `args.only` selects on the current CUDA device; every JSON object goes to emit() (default: print)"""
    import torch
    import bench_support as bs
    import oracle_lib as orc
    import wgatools_b200 as w
    if ctx is None:
        ctx = w.Context(torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    L = ctx.L
    if emit is None:
        def emit(line):
            print(json.dumps(line), flush=True)

    def timed(fn, steps):
        fn()
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def resident(data):
        d = torch.zeros(len(data) + 64, dtype=torch.uint8, device="cuda")
        d[:len(data)].copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
        torch.cuda.synchronize()
        return d

    def sample_lines(data, mb):
        n = int(mb * 1e6)
        if n >= len(data):
            return data
        return data[:data.rfind(b"\n", 0, n) + 1]

    def report(name, nbytes, ms_dev, ms_e2e, out_bytes, cpu_gbs, cpu_n, extra=None):
        line = {"command": name, "metric": "alignment GB/s (input PAF bytes)", "value": nbytes / ms_dev / 1e6, "unit": "GB/s",
                "ms_per_step": ms_dev, "e2e": {"value": nbytes / ms_e2e / 1e6, "unit": "GB/s", "ms_per_step": ms_e2e,
                                               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": out_bytes},
                "input_bytes": nbytes, "output_bytes": out_bytes,
                "cpu_baseline": {"value": cpu_gbs, "unit": "GB/s", "cores": 1, "kind": "port",
                                 "sample": f"first {cpu_n / 1e6:.0f} MB of the same PAF, C restatement of the reference algorithm"}}
        if extra:
            line.update(extra)
        emit(line)

    # ---------------- config 4 (scaled): stat -f paf
    if args.only == "fasta":
        args.records = 1000
    paf = bs.gen_paf(bs.paf_params(args.records, seed=20260928))
    d = resident(paf)
    def pinned_copy(data):  # pinned host memory: the e2e leg should see PCIe, not a pageable staging copy
        ptr = L.wga_host_alloc(len(data) + 64)
        ctypes.memmove(ptr, data, len(data))
        return ptr
    hp = pinned_copy(paf)

    def st_dev():
        acc = w.StatAcc(False)
        ctx.paf_stat_raw(ctypes.c_void_p(d.data_ptr()), len(paf), 1, 1, acc)
        return acc

    def st_host():
        acc = w.StatAcc(False)
        ctx.paf_stat_raw(ctypes.c_void_p(hp), len(paf), 0, 1, acc)
        return acc
    ms_dev, ms_e2e = timed(st_dev, args.steps), timed(st_host, args.steps)
    smp = sample_lines(paf, args.cpu_mb)
    t0 = time.perf_counter()
    orc.stat_paf(smp)
    cpu = len(smp) / (time.perf_counter() - t0) / 1e9
    report("stat -f paf (config 4 shape)", len(paf), ms_dev, ms_e2e, 0, cpu, len(smp), {"records": args.records})
    # paf2chain on the same PAF (the only command of the reference with a published timing)
    spc = w.api.Span()

    def pc(ptr, mem):
        def run():
            ctx._check(L.wga_paf2chain(ctx.h, ptr, len(paf), mem, 1, 0, ctypes.byref(spc), None))
        return run
    ms_dev, ms_e2e = timed(pc(ctypes.c_void_p(d.data_ptr()), 1), args.steps), timed(pc(ctypes.c_void_p(hp), 0), args.steps)
    t0 = time.perf_counter()
    orc.paf2chain(smp)
    cpu = len(smp) / (time.perf_counter() - t0) / 1e9
    report("paf2chain (config 4 shape)", len(paf), ms_dev, ms_e2e, spc.len, cpu, len(smp), {"records": args.records})
    # validate -f on the same PAF
    spr, spf = w.api.Span(), w.api.Span()

    def vd(ptr, mem):
        def run():
            ctx._check(L.wga_paf_validate(ctx.h, ptr, len(paf), mem, 1, 1, ctypes.byref(spr), ctypes.byref(spf)))
        return run
    ms_dev, ms_e2e = timed(vd(ctypes.c_void_p(d.data_ptr()), 1), args.steps), timed(vd(ctypes.c_void_p(hp), 0), args.steps)
    t0 = time.perf_counter()
    orc.paf_validate(smp, True)
    cpu = len(smp) / (time.perf_counter() - t0) / 1e9
    report("validate -f (config 4 shape)", len(paf), ms_dev, ms_e2e, spf.len, cpu, len(smp), {"records": args.records})
    # filter -a (pair-sum filter, cut-off = the median pair sum would need a pre-pass: 1 Mbp keeps roughly half) and the
    # base-level plot data on the same PAF
    spx = w.api.Span()
    cut = ctypes.c_uint64(1_000_000)

    def fl(ptr, mem):
        def run():
            ctx._check(L.wga_paf_filter(ctx.h, ptr, len(paf), mem, 1, 0, 0, ctypes.byref(cut), ctypes.byref(spx)))
        return run
    ms_dev, ms_e2e = timed(fl(ctypes.c_void_p(d.data_ptr()), 1), args.steps), timed(fl(ctypes.c_void_p(hp), 0), args.steps)
    t0 = time.perf_counter()
    orc.paf_filter(smp, min_align=1_000_000)
    cpu = len(smp) / (time.perf_counter() - t0) / 1e9
    report("filter -f paf -a 1000000 (config 4 shape)", len(paf), ms_dev, ms_e2e, spx.len, cpu, len(smp), {"records": args.records})

    def pdat(ptr, mem):
        def run():
            ctx._check(L.wga_paf_plotdata(ctx.h, ptr, len(paf), mem, 1, 50, ctypes.byref(spx)))
        return run
    ms_dev, ms_e2e = timed(pdat(ctypes.c_void_p(d.data_ptr()), 1), args.steps), timed(pdat(ctypes.c_void_p(hp), 0), args.steps)
    t0 = time.perf_counter()
    orc.paf_plotdata(smp, 50)
    cpu = len(smp) / (time.perf_counter() - t0) / 1e9
    report("dotplot base-level data -l 50 (config 4 shape)", len(paf), ms_dev, ms_e2e, spx.len, cpu, len(smp), {"records": args.records})
    # pafcov: one BED line per target base, so a PAF over a small target set (10 x 2 Mbp)
    pafc = bs.gen_paf(bs.paf_params(200_000, seed=20260930, n_genomes=2, n_chr=5, chr_size=2_000_000, ops_median=150.0, ops_max=2000))
    dc = resident(pafc)
    hpc = pinned_copy(pafc)
    spv = w.api.Span()

    def cv(ptr, mem):
        def run():
            ctx._check(L.wga_paf_cov(ctx.h, ptr, len(pafc), mem, 1, ctypes.byref(spv)))
        return run
    ms_dev, ms_e2e = timed(cv(ctypes.c_void_p(dc.data_ptr()), 1), args.steps), timed(cv(ctypes.c_void_p(hpc), 0), args.steps)
    smpc = sample_lines(pafc, 20.0)
    t0 = time.perf_counter()
    orc.pafcov(smpc)
    cpu = len(smpc) / (time.perf_counter() - t0) / 1e9
    report("pafcov (10 targets x 2 Mbp)", len(pafc), ms_dev, ms_e2e, spv.len, cpu, len(smpc), {"records": 200_000})
    del dc
    del d
    if args.only == "stat":
        return

    # ---------------- config 5 (scaled): FASTA-backed commands
    chr_size = int(args.chr_mbp * 1e6)
    fa, fai = make_fasta(args.genomes, args.chrs, chr_size)
    F = w.Fasta(ctx, fa, fai)
    paf5 = bs.gen_paf(bs.paf_params(args.fa_records, seed=20260929, n_genomes=args.genomes, n_chr=args.chrs, chr_size=chr_size,
                                    ops_median=150.0, ops_max=2000))
    d5 = resident(paf5)
    hp5 = pinned_copy(paf5)
    sp, sp2 = w.api.Span(), w.api.Span()
    outn = {}

    def mk(fn_name, dev):
        ptr, mem = (ctypes.c_void_p(d5.data_ptr()), 1) if dev else (ctypes.c_void_p(hp5), 0)

        def run():
            if fn_name == "paf2maf":
                ctx._check(L.wga_paf2maf(ctx.h, ptr, len(paf5), mem, 1, F.h, F.h, ctypes.byref(sp)))
            elif fn_name == "call":
                ctx._check(L.wga_paf_call(ctx.h, ptr, len(paf5), mem, 1, F.h, F.h, 1, 0, ctypes.byref(sp)))
            else:
                ctx._check(L.wga_paf_pseudo(ctx.h, ptr, len(paf5), mem, 1, F.h, None, ctypes.byref(sp), ctypes.byref(sp2)))
            outn[fn_name] = sp.len
        return run
    smp5 = sample_lines(paf5, min(args.cpu_mb, 4.0))  # paf2maf writes ~80x its input
    for name, label, cpu_fn in (("paf2maf", "paf2maf (config 5 shape)", lambda s: orc.paf2maf(s, fa, fai, fa, fai)),
                                ("call", "call -f paf -s -l0", lambda s: orc.call_paf(s, fa, fai, fa, fai, snp=True, svlen=0)),
                                ("pseudo", "pafpseudo -f all.fa", lambda s: orc.paf_pseudo(s, fa, fai))):
        ms_dev, ms_e2e = timed(mk(name, True), args.steps), timed(mk(name, False), args.steps)
        walk_ms, walk_bytes, _ = ctx.last_walk()
        t0 = time.perf_counter()
        cpu_fn(smp5)
        cpu = len(smp5) / (time.perf_counter() - t0) / 1e9
        extra = {"records": paf5.count(b"\n"), "fasta_bytes": len(fa)}
        if name in ("paf2maf", "pseudo") and walk_ms > 0:
            extra["expand_kernels"] = {"ms": walk_ms, "output_GBps": walk_bytes / walk_ms / 1e6}
        report(label, len(paf5), ms_dev, ms_e2e, outn[name], cpu, len(smp5), extra)


def main():
    import torch
    import __graft_entry__ as g
    g.build()
    torch.cuda.set_device(0)
    run(parse_args())


if __name__ == "__main__":
    main()
