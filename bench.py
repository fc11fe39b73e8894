#!/usr/bin/env python
"""bench.py -- alignment GB/s (input MAF bytes) of the B200 hot path: maf2paf + stat + call over one
synthetic 2-genome MAF (BASELINE.json configs[1] shape: ~1 M blocks / 3 Gbp).

A "step" is one pass of the three commands over the batch.  `value` times the pass with the MAF text
already resident in HBM (inputs larger than L2, so no flush is needed); `e2e` times the same pass
through the host-buffer pipeline (pinned host text in, host PAF / VCF / stat table out: H2D and D2H
copies inside the timed region).  `roofline` is the dense column-walk kernel: algorithmic bytes
(2 B per column + 48 B per work item) over its CUDA-event time, against MEASURED_PEAKS.json.
`cpu_baseline` is the oracle (C restatement of the reference algorithm) on a bounded sample of the same
workload -- and the GPU outputs of that very sample are compared with the oracle's byte for byte
(`parity`); a mismatch makes the run fail.  `--impl reference` times that CPU path alone.

N > 1 (torchrun, one rank per GPU): the job is ONE file of N x blocks records, rank r holding the r-th contiguous
range; every timed step ends with the stat all-reduce (wga_stat_allreduce: NCCL inside the C-ABI
library) and rank 0 printing the TSV, and the merged table is checked against a single-process pass over the whole
file.  `sharded` is the strong-scaling leg on the config-4 shape: ONE all-to-all PAF cut with dist.shard_range,
`stat -f paf` per range + the all-reduce in every step, TSV byte-identical to the single-GPU run and to the oracle.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--blocks B] [--config 2|3|4|5] [--impl reference]
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DEFAULT_BLOCKS = 1_000_000
SEED = 20260926
PAF_SEED = 20260928


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--blocks", type=int, default=int(os.environ.get("WGA_BENCH_BLOCKS", 0)))
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-sample-mb", type=float, default=float(os.environ.get("WGA_BENCH_CPU_MB", 1500)))
    ap.add_argument("--no-call", action="store_true")
    ap.add_argument("--config", type=int, default=int(os.environ.get("WGA_BENCH_CONFIG", 2)), choices=(2, 3, 4, 5),
                    help="2 = the headline workload (configs[1]); 3 = long-record shape of configs[2] (30 GB with --blocks 98000); "
                         "4 / 5 = the PAF-side legs (stat -f paf, paf2chain, validate, pafcov / paf2maf, call -f paf, pafpseudo)")
    ap.add_argument("--no-indep", action="store_true", help="profiling runs only: skip the independent-commands leg")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer legs")
    ap.add_argument("--no-side", action="store_true", help="skip the config 4 / 5 side legs of the default single-GPU run")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded stat -f paf leg")
    ap.add_argument("--paf-records", type=int, default=int(os.environ.get("WGA_BENCH_PAF_RECORDS", 2_000_000)),
                    help="records of the ONE config-4-shaped PAF the sharded leg cuts across the ranks")
    ap.add_argument("--p-mis", type=float, default=None, help="override the mismatch rate of the generator (fast-path dependence)")
    ap.add_argument("--slab-mb", type=int, default=int(os.environ.get("WGA_BENCH_SLAB_MB", 256)))
    a = ap.parse_args()
    if a.blocks <= 0:
        a.blocks = DEFAULT_BLOCKS if a.config == 2 else 20_000
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def walk_traffic(blocks):
    """dram__bytes_read.sum + dram__bytes_write.sum per k_walk launch from the committed ncu --set full capture of this
    workload (profiles/walk_traffic.json); null when the capture was taken on another workload size"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "walk_traffic.json")))
        return t["dram_bytes_per_launch"] if int(t["workload_blocks"]) == int(blocks) else None
    except Exception:
        return None


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def gen_kwargs(args):
    return {} if args.p_mis is None else {"p_mis": args.p_mis}


def make_workload(blocks, seed, pinned_alloc, config=2, threads=None, **over):
    import bench_support as bs
    params = bs.maf_params(blocks, seed=seed, config=config, threads=threads, **over)
    keep, nbytes, ncols = bs.gen_maf_into(params, pinned_alloc)
    return keep, nbytes, ncols


def bind_to_gpu_numa(dev):
    """pin this rank to the CPUs of its GPU's NUMA node (so that the pinned MAF buffer is allocated and first touched
    there): at N=8 every rank streams ~52 GB/s over PCIe and a remote-socket buffer would go through UPI.
    Best effort: returns the node or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(dev)
        pci = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{pci}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def cpu_sample_len(host_ptr, nbytes, want):
    """largest safe cut <= want bytes (whole MAF blocks only)"""
    if want >= nbytes:
        return nbytes
    buf = (ctypes.c_char * want).from_address(host_ptr)
    data = bytes(buf[max(0, want - (1 << 22)):want])
    # last blank line inside the window: blocks are separated by "\n\n"
    k = data.rfind(b"\n\n")
    return want - len(data) + k + 2 if k >= 0 else want


def run_cpu(host_ptr, nbytes, sample_bytes, threads, do_call=True, keep=False):
    """times the CPU port of the reference on text[0:sample] with the reference's own parallel structure at
    `-t threads`: maf2paf and stat pull records from a serial reader (par_bridge, src/converter.rs:41-48,
    src/tools/stat.rs:67-81) into a worker pool; call is strictly sequential (src/tools/caller.rs:62).
    keep: also return the three outputs (the bench compares the GPU's outputs of the same sample with them)."""
    import oracle_lib as orc
    orc.lib()
    n = cpu_sample_len(host_ptr, nbytes, int(sample_bytes))
    text = ctypes.string_at(host_ptr, n)
    t0 = time.perf_counter()
    paf = orc.maf2paf_mt(text, threads)
    tsv = orc.stat_maf_mt(text, threads)
    vcf = orc.call_maf(text, snp=True, svlen=0) if do_call else None
    dt = time.perf_counter() - t0
    return n / dt / 1e9, n, dt, ((paf, tsv, vcf) if keep else None)


CPU_KIND_NOTE = ("C restatement of the reference algorithm (Rust toolchain unavailable) with the reference's rayon "
                 "structure: serial record reader feeding a worker pool for maf2paf/stat, sequential call")


def build_checker():
    """the oracle and the generators only -- the reference arm must not load the product library"""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "bench_support"), "-s"])


def workload_text(args, per_gpu=True):
    if args.config == 3:
        return (f"configs[2] shape: maf2paf+stat+call(-s -l0) over {args.blocks} long blocks{'/GPU' if per_gpu else ''} (lognormal median "
                "50 kb, SVs to 100 kb, SURVEY 8(d) cfg 3; 98000 blocks = 29.5 GB)")
    return (f"configs[1]: maf2paf+stat+call(-s -l0) over a synthetic 2-genome MAF, {args.blocks} blocks{'/GPU' if per_gpu else ''} "
            f"(lognormal median 2 kb, SURVEY 8(d) cfg 2), seed {SEED}+rank")


def reference_arm(args, rank, metric):
    if rank != 0:
        return
    import numpy as np
    build_checker()
    import bench_support as bs
    cores = os.cpu_count() or 1
    # bounded sample: sized so that the whole (warmup + steps) run ends within ~2-3 minutes at the ~0.2 GB/s
    # the three commands reach together (the sequential `call` dominates)
    budget_s = 150.0 / max(args.warmup + args.steps, 1)
    sample_mb = min(1500.0, max(40.0, budget_s * 150.0))
    per_block = 6700 if args.config == 2 else 305000
    sample_blocks = max(200 if args.config == 3 else 1000, min(args.blocks, int(sample_mb * 1e6 / per_block)))

    def alloc(n):
        a = np.empty(n, dtype=np.uint8)
        return a.ctypes.data, a
    keep, nbytes, ncols = bs.gen_maf_into(bs.maf_params(sample_blocks, seed=SEED, config=args.config, **gen_kwargs(args)), alloc)
    ptr = keep.ctypes.data
    do_call = not args.no_call
    for _ in range(args.warmup):
        run_cpu(ptr, nbytes, nbytes, cores, do_call)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_cpu(ptr, nbytes, nbytes, cores, do_call)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    v = nbytes / dt / 1e9
    sample = (f"{sample_blocks} blocks ({nbytes / 1e6:.0f} MB) of the same generator per step (the workload's "
              f"{args.blocks} blocks would take minutes per step on the CPU); " + CPU_KIND_NOTE + f", -t {cores}")
    config = {"workload": workload_text(args), "blocks_per_gpu": args.blocks, "blocks_timed_per_step": sample_blocks,
              "commands": ["maf2paf", "stat", "call -s -l0"] if do_call else ["maf2paf", "stat"],
              "index": "every command parses the sample itself (the reference re-reads the file per command)"}
    line = {"impl": "reference", "metric": metric, "value": v, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def snap_newline(view, pos):
    """first line start at or after pos (numpy uint8 view of a PAF) -- dist.shard_range's snap for kind='paf'"""
    import numpy as np
    n = len(view)
    if pos <= 0:
        return 0
    if pos >= n:
        return n
    p = pos - 1
    while p < n:
        w = view[p:min(n, p + (1 << 20))]
        k = np.flatnonzero(w == 10)
        if len(k):
            return min(p + int(k[0]) + 1, n)
        p += len(w)
    return n


def sharded_paf_leg(args, ctx, comm, rank, world, dist, timed_fn, all_cpus):
    """the config-4 shape across ranks: ONE all-to-all PAF (16 genomes x 20 chromosomes, `=`/X/I/D CIGARs) made of `world`
    consecutive parts of --paf-records records each; rank r holds part r (its byte range of the file, cut at a line start:
    dist.shard_range semantics).  Every timed step = wga_paf_stat on the rank's range + wga_stat_allreduce + the TSV on
    rank 0.  Weak scaling (2.02 GB per GPU); rank 0 then runs ONE single-GPU pass over the whole file and compares."""
    import numpy as np
    import torch
    import bench_support as bs
    import wgatools_b200 as w
    L = ctx.L
    keep = {}

    def pinned(n):
        p = L.wga_host_alloc(n + 64)
        if not p:
            raise MemoryError("wga_host_alloc failed")
        keep["p"] = p
        return p, p
    threads = max(4, len(os.sched_getaffinity(0)) // max(1, (world + 1) // 2))
    _, n_mine = bs.gen_paf_into(bs.paf_params(args.paf_records, seed=PAF_SEED + rank, threads=threads), pinned)
    host = keep["p"]
    view = np.ctypeslib.as_array((ctypes.c_uint8 * n_mine).from_address(host))
    assert n_mine and view[n_mine - 1] == 10, "a part must end at a line end"
    d = torch.zeros(n_mine + 64, dtype=torch.uint8, device="cuda")
    d[:n_mine].copy_(torch.from_numpy(view))
    torch.cuda.synchronize()
    sizes = torch.tensor([n_mine], dtype=torch.int64, device="cuda")
    if world > 1:
        allsz = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(allsz, sizes)
        nbytes = int(sum(int(t.item()) for t in allsz))
    else:
        nbytes = n_mine
    out = {"tsv": b""}

    def step(mem, ptr, keep_bytes=False):
        acc = w.StatAcc(False)
        ctx.paf_stat_raw(ctypes.c_void_p(ptr), n_mine, mem, 0, acc)
        if world > 1:
            comm.allreduce_stat(acc)
        if rank == 0:
            if keep_bytes:
                out["tsv"] = acc.write()       # a Python copy of the 15 MB table: only for the comparison below
            else:
                out["span"] = acc.write_span()  # the TSV, formatted, inside the accumulator: what a host would write out
        return acc
    step(1, d.data_ptr(), True)
    for _ in range(2):
        step(1, d.data_ptr())
    nc0 = L.wga_ctx_nccl_calls(ctx.h)
    ms_dev = timed_fn(lambda: step(1, d.data_ptr()), args.steps)
    nc1 = L.wga_ctx_nccl_calls(ctx.h)
    step(0, host)
    ms_e2e = timed_fn(lambda: step(0, host), args.steps)
    # the exchange alone: the per-range tables are rebuilt outside the timed region
    ms_ar = 0.0
    if world > 1:
        accs = []
        for _ in range(args.steps):
            a = w.StatAcc(False)
            ctx.paf_stat_raw(ctypes.c_void_p(d.data_ptr()), n_mine, 1, 0, a)
            accs.append(a)
        it = iter(accs)
        ms_ar = timed_fn(lambda: comm.allreduce_stat(next(it)), args.steps)
    res = {"workload": f"config 4 shape: stat -f paf over ONE all-to-all PAF of {world} x {args.paf_records} records ({nbytes / 1e9:.2f} GB), "
                       f"rank r holds the r-th of {world} consecutive byte range(s); every step = wga_paf_stat + wga_stat_allreduce + TSV "
                       "on rank 0",
           "scaling": "weak", "value": nbytes / (ms_dev / args.steps * 1e-3) / 1e9, "unit": "GB/s",
           "ms_per_step": ms_dev / args.steps,
           "e2e": {"value": nbytes / (ms_e2e / args.steps * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_e2e / args.steps,
                   "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 0},
           "allreduce_ms": ms_ar / args.steps, "nccl_calls_per_step": (nc1 - nc0) / max(args.steps, 1),
           "input_bytes": nbytes, "range_bytes_rank0": n_mine}
    del d
    L.wga_host_free(ctypes.c_void_p(host))
    torch.cuda.empty_cache()
    if rank == 0:
        os.sched_setaffinity(0, all_cpus)
        # parity 1: ONE single-GPU pass over the WHOLE file (all parts regenerated into one host buffer, in rank order)
        whole = L.wga_host_alloc(nbytes + 64)
        if not whole:
            raise MemoryError("wga_host_alloc failed")
        off = 0
        for r in range(world):
            def into(n, base=whole + off):
                return base, base
            _, pn = bs.gen_paf_into(bs.paf_params(args.paf_records, seed=PAF_SEED + r), into)
            off += pn
        assert off == nbytes
        one = w.StatAcc(False)
        ctx.paf_stat_raw(ctypes.c_void_p(whole), nbytes, 0, 0, one)
        tsv1 = one.write()
        ok = out["tsv"] == tsv1
        res["parity_vs_single_gpu"] = bool(ok)
        res["tsv_rows"] = tsv1.count(b"\n") - 1
        # parity 2: the CPU oracle over the whole file (bounded: ~15 s per 2 GB)
        if nbytes <= float(os.environ.get("WGA_BENCH_ORACLE_MB", 2600)) * 1e6:
            import oracle_lib as orc
            t0 = time.perf_counter()
            want = orc.stat_paf(ctypes.string_at(whole, nbytes))
            res["parity_vs_oracle"] = bool(out["tsv"] == want)
            res["oracle_seconds"] = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": nbytes / res["oracle_seconds"] / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                   "sample": "the whole PAF, one thread (stat_paf, src/tools/stat.rs:87-105)"}
            ok = ok and res["parity_vs_oracle"]
        res["parity"] = bool(ok)
        L.wga_host_free(ctypes.c_void_p(whole))
    return res


VERBOSE = bool(os.environ.get("WGA_BENCH_VERBOSE"))
_T0 = time.perf_counter()


def note(msg):
    if VERBOSE:
        sys.stderr.write(f"[bench {time.perf_counter() - _T0:7.2f}s rank {os.environ.get('RANK', 0)}] {msg}\n")
        sys.stderr.flush()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    do_call = not args.no_call
    metric = "alignment GB/s (input MAF bytes) for call+stat+maf2paf"

    if args.impl == "reference":
        reference_arm(args, rank, metric)
        return

    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) go to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    else:
        torch.cuda.set_device(0)
    import numpy as np
    import wgatools_b200 as w
    from wgatools_b200._lib import CallOpts, WGA_F_REUSE_INDEX, WGA_F_SHARE_WALK
    dev = torch.cuda.current_device()
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa(dev)   # pinned input buffers must live on the socket the GPU hangs off
    stream = torch.cuda.current_stream().cuda_stream
    ctx = w.Context(dev, stream)
    L = ctx.L
    comm = w.Comm.nccl(ctx, dist) if world > 1 else None   # ncclCommInitRank inside the C-ABI library

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_fn(fn, steps):
        """K calls of fn bracketed by barrier + synchronize, CUDA events on the ctx stream, max over ranks (ms total)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- explicit PAF-side configs (single GPU): one JSON line per command
    if args.config in (4, 5):
        if rank == 0:
            from bench_support import paf_legs
            pa = paf_legs.parse_args(["--steps", str(args.steps), "--only", "stat" if args.config == 4 else "fasta"])
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
            paf_legs.run(pa, ctx)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    config = {"workload": workload_text(args) + (f"; the job is ONE file of {world} x {args.blocks} blocks, rank r holds the r-th "
                                                  "contiguous range, the stat all-reduce closes every step" if world > 1 else ""),
              "blocks_per_gpu": args.blocks, "l2": "inputs larger than L2 (no flush needed)",
              "commands": ["maf2paf", "stat", "call -s -l0"] if do_call else ["maf2paf", "stat"],
              "index": "one tokenisation and ONE dense column walk per step: maf2paf runs it in the caller classification and keeps "
                       "counters + run list (WGA_F_SHARE_WALK); stat and call pick them up (WGA_F_REUSE_INDEX)"}
    if args.p_mis is not None:
        config["p_mis"] = args.p_mis

    # ---- workload in pinned host memory + a resident device copy
    def pinned(n):
        p = L.wga_host_alloc(n + 64)
        if not p:
            raise MemoryError("wga_host_alloc failed")
        return p, p
    t_gen = time.perf_counter()
    # generator threads: this rank's share of the CPUs of its NUMA node (ranks of one node generate at the same time)
    gen_threads = max(4, len(os.sched_getaffinity(0)) // max(1, (world + 1) // 2))
    host_ptr, nbytes, ncols = make_workload(args.blocks, SEED + rank, pinned, args.config, gen_threads, **gen_kwargs(args))
    t_gen = time.perf_counter() - t_gen
    h_view = torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(host_ptr)))
    d_text = torch.zeros(nbytes + 64, dtype=torch.uint8, device="cuda")
    d_text[:nbytes].copy_(h_view)
    torch.cuda.synchronize()
    opts = CallOpts(snp=1, inv=0, svlen_cutoff=0, chunk_size=1000000, query_name=None)

    out_bytes = {"paf": 0, "vcf": 0}
    walk = {"ms": 0.0, "bytes": 0, "launches": 0}
    last = {"tsv": b""}
    mode = {"share": True}

    def collect_walk():
        ms, by, nl = ctx.last_walk()
        walk["ms"] += ms; walk["bytes"] += by; walk["launches"] += nl

    def finish_stat(acc):
        """the step's exchange: stat totals over all ranks, TSV printed by rank 0"""
        if world > 1:
            comm.allreduce_stat(acc)
        if rank == 0:
            last["tsv"] = acc.write()

    def step(mem, ptr, collect=False):
        f1 = 1 | (WGA_F_SHARE_WALK if mode["share"] else 0)
        f2 = 1 | (WGA_F_REUSE_INDEX if mode["share"] else 0)
        p, n = ctx.maf2paf_raw(ctypes.c_void_p(ptr), nbytes, mem, f1)
        out_bytes["paf"] = n
        if collect:
            collect_walk()
        acc = w.StatAcc(False)
        ctx.maf_stat_raw(ctypes.c_void_p(ptr), nbytes, mem, f2, acc)
        if collect:
            collect_walk()
        if do_call:
            p, n = ctx.maf_call_raw(ctypes.c_void_p(ptr), nbytes, mem, f2, opts)
            out_bytes["vcf"] = n
            if collect:
                collect_walk()
        finish_stat(acc)
        return acc

    note(f"workload ready: {nbytes} bytes, gen {t_gen:.1f}s")
    step(1, d_text.data_ptr())
    note("first step done")
    for _ in range(max(args.warmup - 1, 0)):
        step(1, d_text.data_ptr())
    sampler = ClockSampler(dev)
    sampler.start()
    # independent commands (every command tokenises and walks the slab itself, like three separate invocations) ...
    mode["share"] = False
    ms_indep = float("nan")
    if not args.no_indep:
        step(1, d_text.data_ptr())
        ms_indep = timed_fn(lambda: step(1, d_text.data_ptr()), args.steps)
    # ... and the job as the product runs it: one tokenisation and one walk per step, shared by the three commands
    mode["share"] = True
    step(1, d_text.data_ptr())
    own0, cub0 = ctx.launch_counts()
    nc0 = L.wga_ctx_nccl_calls(ctx.h)
    ms_dev = timed_fn(lambda: step(1, d_text.data_ptr(), True), args.steps)
    own1, cub1 = ctx.launch_counts()
    nc1 = L.wga_ctx_nccl_calls(ctx.h)
    tsv_job = last["tsv"]
    note(f"device-resident legs done: {ms_dev / args.steps:.2f} ms/step")
    # e2e: host buffers, copies inside the timed region.
    # (1) the public pipeline: ONE asynchronous slab-wise upload per step (ResidentMaf = wga_text_open), the three
    #     commands run on every slab as it arrives, outputs copied back to pinned host memory inside each call;
    # (2) three independent host-buffer calls (the file crosses PCIe three times), kept as e2e_separate_calls.
    d2h = {"n": 0}

    def step_resident():
        acc = w.StatAcc(False)
        tot = 0
        with w.ResidentMaf((host_ptr, nbytes), ctx, args.slab_mb << 20) as rm:
            for p, n, fl in rm.slabs():
                tot += ctx.maf2paf_raw(p, n, 1, fl | WGA_F_SHARE_WALK)[1]
                ctx.maf_stat_raw(p, n, 1, fl | WGA_F_REUSE_INDEX, acc)
                if do_call:
                    tot += ctx.maf_call_raw(p, n, 1, fl | WGA_F_REUSE_INDEX, opts)[1]
        d2h["n"] = tot
        finish_stat(acc)
        return acc

    sep_steps = max(1, min(args.steps, 3))
    ms_e2e = ms_e2e_sep = float("nan")
    if not args.no_e2e:
        for _ in range(2):
            step_resident()
        ms_e2e = timed_fn(step_resident, args.steps)
        mode["share"] = False
        step(0, host_ptr)
        ms_e2e_sep = timed_fn(lambda: step(0, host_ptr), sep_steps)
        mode["share"] = True
    tsv_e2e = last["tsv"]
    note("host-buffer legs done")
    # the exchange alone (N > 1): per-rank tables rebuilt outside the timed region
    ms_ar = 0.0
    if world > 1:
        accs = []
        for _ in range(args.steps):
            a = w.StatAcc(False)
            ctx.maf_stat_raw(ctypes.c_void_p(d_text.data_ptr()), nbytes, 1, 1, a)
            accs.append(a)
        it = iter(accs)
        ms_ar = timed_fn(lambda: comm.allreduce_stat(next(it)), args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    per_step = ms_dev / args.steps
    value = world * nbytes / (per_step * 1e-3) / 1e9
    e2e_value = world * nbytes / (ms_e2e / args.steps * 1e-3) / 1e9
    peak, peak_src = peaks()
    achieved = (walk["bytes"] / max(walk["launches"], 1)) / (walk["ms"] / max(walk["launches"], 1) * 1e-3) / 1e9 if walk["ms"] > 0 else 0.0
    line = {"metric": metric, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": d2h["n"],
                    "steps": args.steps, "ms_per_step": ms_e2e / args.steps,
                    "mode": f"ResidentMaf: one async upload per step in {args.slab_mb} MiB slabs, the three commands run "
                            "per slab while the next slab uploads; outputs D2H inside every call"},
            "e2e_separate_calls": {"value": world * nbytes / (ms_e2e_sep / sep_steps * 1e-3) / 1e9, "unit": "GB/s",
                                   "h2d_bytes_per_step": nbytes * (3 if do_call else 2), "steps": sep_steps,
                                   "d2h_bytes_per_step": out_bytes["paf"] + out_bytes["vcf"],
                                   "mode": "three independent host-buffer calls (file uploaded, tokenised and walked three times): "
                                           "the structure the CPU baseline has"},
            "value_independent_commands": {"value": world * nbytes / (ms_indep / args.steps * 1e-3) / 1e9, "unit": "GB/s",
                                           "ms_per_step": ms_indep / args.steps,
                                           "mode": "each of the three commands tokenises and walks the slab itself (no flags)"},
            "gpu_launches": own1 - own0, "cub_scan_calls": cub1 - cub0, "nccl_calls": nc1 - nc0,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "k_walk_tma<MODE_RUNS_ALL> (dense column walk, TMA-staged; the one launch of every timed step)",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": walk_traffic(args.blocks) if args.config == 2 else None,
                         "launches": walk["launches"], "avg_launch_ms": walk["ms"] / max(walk["launches"], 1),
                         "alg_bytes_per_launch": walk["bytes"] / max(walk["launches"], 1)},
            "input_bytes_per_gpu": nbytes, "columns_per_gpu": ncols, "gen_seconds": t_gen, "numa_node": numa}
    failed = []
    note("exchange-only leg done")
    if world > 1:
        line["allreduce_ms"] = ms_ar / args.steps
    # ---- parity of the multi-GPU exchange: the merged table against ONE process reading the whole file
    if world > 1 and rank == 0:
        os.sched_setaffinity(0, all_cpus)
        t0 = time.perf_counter()
        one = w.StatAcc(False)
        ctx.maf_stat_raw(ctypes.c_void_p(host_ptr), nbytes, 0, 1, one)
        for r in range(1, world):
            part, pn, _ = make_workload(args.blocks, SEED + r, pinned, args.config, None, **gen_kwargs(args))
            ctx.maf_stat_raw(ctypes.c_void_p(part), pn, 0, 1, one)
            L.wga_host_free(ctypes.c_void_p(part))
        want = one.write()
        ok = tsv_job == want and (args.no_e2e or tsv_e2e == want)
        line["stat_allreduce_parity"] = {"parity": bool(ok), "tsv_rows": want.count(b"\n") - 1, "seconds": time.perf_counter() - t0,
                                         "checked": f"TSV of the {world}-rank job (device-resident and host-buffer legs) == one process "
                                                    f"running wga_maf_stat over all {world} ranges in file order"}
        if not ok:
            failed.append("stat all-reduce differs from the single-process table")
    # ---- CPU baseline on a bounded sample + byte parity of the GPU outputs of that very sample
    if rank == 0 and world == 1:
        os.sched_setaffinity(0, all_cpus)   # the CPU baseline gets every host core back
        cores = os.cpu_count() or 1
        gbs, n, dt, outs = run_cpu(host_ptr, nbytes, args.cpu_sample_mb * 1e6, cores, do_call, keep=True)
        line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
                                "sample": f"first {n / 1e6:.0f} MB of the same MAF ({dt:.1f} s); " + CPU_KIND_NOTE
                                          + f", -t {cores}"}
        p, ln = ctx.maf2paf_raw(ctypes.c_void_p(host_ptr), n, 0, 1 | WGA_F_SHARE_WALK)
        ok_paf = ctypes.string_at(p, ln) == outs[0]
        acc = w.StatAcc(False)
        ctx.maf_stat_raw(ctypes.c_void_p(host_ptr), n, 0, 1, acc)
        ok_tsv = acc.write() == outs[1]
        ok_vcf = True
        if do_call:
            p, ln = ctx.maf_call_raw(ctypes.c_void_p(host_ptr), n, 0, 1, opts)
            ok_vcf = ctx.vcf_header("sample") + ctypes.string_at(p, ln) == outs[2]
        line["parity_checked_bytes"] = n
        line["parity"] = bool(ok_paf and ok_tsv and ok_vcf)
        line["parity_detail"] = {"maf2paf": bool(ok_paf), "stat": bool(ok_tsv), "call": bool(ok_vcf),
                                 "against": "the oracle's PAF / TSV / VCF of the cpu_baseline sample, byte for byte"}
        if not line["parity"]:
            failed.append("GPU outputs differ from the oracle on the cpu_baseline sample")
    note("parity checks done")
    # free the MAF before the PAF legs
    del d_text
    del h_view
    L.wga_host_free(ctypes.c_void_p(host_ptr))
    torch.cuda.empty_cache()
    if not args.no_sharded and args.config == 2:
        sh = sharded_paf_leg(args, ctx, comm, rank, world, dist, timed_fn, all_cpus)
        line["sharded"] = sh
        note("sharded leg done")
        if rank == 0 and not sh.get("parity", True):
            failed.append("sharded stat -f paf differs from the single-GPU / oracle table")
    if rank == 0 and world == 1 and not args.no_side and args.config == 2:
        from bench_support import paf_legs
        side = []
        try:
            paf_legs.run(paf_legs.parse_args(["--steps", "3", "--records", "1000", "--only", "fasta"]), ctx, side.append)
            line["side_configs"] = {"note": "config 5 shape, scaled (4 genomes x 5 x 1 Mbp FASTA resident in HBM; 50 k records): "
                                            "paf2maf / call -f paf / pafpseudo; `python bench.py --config 4|5` prints the full legs",
                                    "legs": [s for s in side if "config 5" in s["command"] or "call -f" in s["command"] or "pafpseudo" in s["command"]]}
        except Exception as e:  # a side leg must not take the headline line down
            line["side_configs"] = {"error": repr(e)}
    if rank == 0:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        comm.close()
        dist.destroy_process_group()
    if failed:
        sys.stderr.write("bench.py: PARITY FAILURE: " + "; ".join(failed) + "\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
