#!/usr/bin/env python
"""bench.py -- alignment GB/s (input MAF bytes) of the B200 hot path: maf2paf + stat + call over one
synthetic 2-genome MAF (BASELINE.json configs[1] shape: ~1 M blocks / 3 Gbp).

A "step" is one pass of the three commands over the batch.  `value` times the pass with the MAF text
already resident in HBM (inputs larger than L2, so no flush is needed); `e2e` times the same pass
through the host-buffer C-ABI entry points (pinned host text in, host PAF / VCF / stat partials out:
H2D and D2H copies inside the timed region).  `roofline` is the dense column-walk kernel: algorithmic
bytes (2 B per column + 32 B per work item) over its CUDA-event time, against MEASURED_PEAKS.json.
`cpu_baseline` is the oracle (C restatement of the reference algorithm) on a bounded sample of the same
workload.  `--impl reference` times that CPU path alone, on every host core it can use.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--blocks B] [--impl reference]
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--blocks", type=int, default=int(os.environ.get("WGA_BENCH_BLOCKS", DEFAULT_BLOCKS)))
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-sample-mb", type=float, default=float(os.environ.get("WGA_BENCH_CPU_MB", 1500)))
    ap.add_argument("--no-call", action="store_true")
    ap.add_argument("--config", type=int, default=2, choices=(2, 3),
                    help="2 = the headline workload (configs[1]); 3 = long-record shape of configs[2] (side measurement)")
    ap.add_argument("--no-indep", action="store_true", help="profiling runs only: skip the independent-commands leg")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer legs")
    ap.add_argument("--slab-mb", type=int, default=int(os.environ.get("WGA_BENCH_SLAB_MB", 256)))
    return ap.parse_args()


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


def make_workload(blocks, seed, pinned_alloc, config=2, threads=None):
    import bench_support as bs
    params = bs.maf_params(blocks, seed=seed, config=config, threads=threads)
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


def cpu_sample_len(host_ptr, nbytes, want, L=None):
    """largest safe cut <= want bytes (whole MAF blocks only)"""
    if want >= nbytes:
        return nbytes
    buf = (ctypes.c_char * want).from_address(host_ptr)
    data = bytes(buf[max(0, want - (1 << 22)):want])
    # last blank line inside the window: blocks are separated by "\n\n"
    k = data.rfind(b"\n\n")
    return want - len(data) + k + 2 if k >= 0 else want


def run_cpu(host_ptr, nbytes, sample_bytes, threads, do_call=True):
    """times the CPU port of the reference on text[0:sample] with the reference's own parallel structure at
    `-t threads`: maf2paf and stat pull records from a serial reader (par_bridge, src/converter.rs:41-48,
    src/tools/stat.rs:67-81) into a worker pool; call is strictly sequential (src/tools/caller.rs:62)."""
    import oracle_lib as orc
    orc.lib()
    n = cpu_sample_len(host_ptr, nbytes, int(sample_bytes))
    text = ctypes.string_at(host_ptr, n)
    t0 = time.perf_counter()
    orc.maf2paf_mt(text, threads)
    orc.stat_maf_mt(text, threads)
    if do_call:
        orc.call_maf(text, snp=True, svlen=0)
    dt = time.perf_counter() - t0
    return n / dt / 1e9, n, dt


CPU_KIND_NOTE = ("C restatement of the reference algorithm (Rust toolchain unavailable) with the reference's rayon "
                 "structure: serial record reader feeding a worker pool for maf2paf/stat, sequential call")


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    do_call = not args.no_call
    metric = "alignment GB/s (input MAF bytes) for call+stat+maf2paf"
    config = {"workload": f"configs[1]: maf2paf+stat+call(-s -l0) over a synthetic 2-genome MAF, {args.blocks} blocks/GPU "
                          f"(lognormal median 2 kb, SURVEY 8(d) cfg 2), seed {SEED}+rank" if args.config == 2 else
                          f"configs[2] shape (side measurement): {args.blocks} long blocks/GPU (lognormal median 50 kb, SVs, SURVEY 8(d) cfg 3)",
              "blocks_per_gpu": args.blocks, "l2": "inputs larger than L2 (no flush needed)",
              "commands": ["maf2paf", "stat", "call -s -l0"] if do_call else ["maf2paf", "stat"],
              "index": "the slab is tokenised on the device once per step (by maf2paf) and reused by stat and call (WGA_F_REUSE_INDEX)"}

    if args.impl == "reference":
        if rank != 0:
            return
        import numpy as np
        import bench_support as bs
        import __graft_entry__ as g
        g.build()
        cores = os.cpu_count() or 1
        # bounded sample: sized so that the whole (warmup + steps) run ends within ~2-3 minutes at the ~0.15 GB/s
        # the three commands reach together (the sequential `call` dominates)
        budget_s = 150.0 / max(args.warmup + args.steps, 1)
        sample_mb = min(1500.0, max(40.0, budget_s * 150.0))
        sample_blocks = max(1000, min(args.blocks, int(sample_mb * 1e6 / 6700)))

        def alloc(n):
            a = np.empty(n, dtype=np.uint8)
            return a.ctypes.data, a
        keep, nbytes, ncols = bs.gen_maf_into(bs.maf_params(sample_blocks, seed=SEED, config=2), alloc)
        ptr = keep.ctypes.data
        for _ in range(args.warmup):
            run_cpu(ptr, nbytes, nbytes, cores, do_call)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run_cpu(ptr, nbytes, nbytes, cores, do_call)
        dt = (time.perf_counter() - t0) / max(args.steps, 1)
        v = nbytes / dt / 1e9
        line = {"impl": "reference", "metric": metric, "value": v, "unit": "GB/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "GB/s", "cores": cores, "kind": "port",
                                 "sample": f"{sample_blocks} blocks ({nbytes / 1e6:.0f} MB) of the same generator per step; "
                                           + CPU_KIND_NOTE + f", -t {cores}"},
                "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    else:
        torch.cuda.set_device(0)
    import wgatools_b200 as w
    from wgatools_b200._lib import CallOpts
    dev = torch.cuda.current_device()
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa(dev)   # pinned input buffers must live on the socket the GPU hangs off
    stream = torch.cuda.current_stream().cuda_stream
    ctx = w.Context(dev, stream)
    L = ctx.L

    # ---- workload in pinned host memory + a resident device copy
    def pinned(n):
        p = L.wga_host_alloc(n + 64)
        if not p:
            raise MemoryError("wga_host_alloc failed")
        return p, p
    t_gen = time.perf_counter()
    # generator threads: this rank's share of the CPUs of its NUMA node (ranks of one node generate at the same time)
    gen_threads = max(4, len(os.sched_getaffinity(0)) // max(1, (world + 1) // 2))
    host_ptr, nbytes, ncols = make_workload(args.blocks, SEED + rank, pinned, args.config, gen_threads)
    t_gen = time.perf_counter() - t_gen
    import numpy as np
    h_view = torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(host_ptr)))
    d_text = torch.zeros(nbytes + 64, dtype=torch.uint8, device="cuda")
    d_text[:nbytes].copy_(h_view)
    torch.cuda.synchronize()
    opts = CallOpts(snp=1, inv=0, svlen_cutoff=0, chunk_size=1000000, query_name=None)

    out_bytes = {"paf": 0, "vcf": 0}
    walk = {"ms": 0.0, "bytes": 0, "launches": 0}

    reuse = {"f": 0}   # WGA_F_REUSE_INDEX for stat/call on device-resident text: the slab is tokenised once per step

    def step(mem, ptr, collect=False):
        p, n = ctx.maf2paf_raw(ctypes.c_void_p(ptr), nbytes, mem, 1)
        out_bytes["paf"] = n
        if collect:
            ms, by, nl = ctx.last_walk()
            walk["ms"] += ms; walk["bytes"] += by; walk["launches"] += nl
        acc = w.StatAcc(False)
        ctx.maf_stat_raw(ctypes.c_void_p(ptr), nbytes, mem, 1 | reuse["f"], acc)
        if collect:
            ms, by, nl = ctx.last_walk()
            walk["ms"] += ms; walk["bytes"] += by; walk["launches"] += nl
        if do_call:
            p, n = ctx.maf_call_raw(ctypes.c_void_p(ptr), nbytes, mem, 1 | reuse["f"], opts)
            out_bytes["vcf"] = n
            if collect:
                ms, by, nl = ctx.last_walk()
                walk["ms"] += ms; walk["bytes"] += by; walk["launches"] += nl
        return acc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(mem, ptr, steps, collect):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            step(mem, ptr, collect)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    try:
        step(1, d_text.data_ptr())
    except w.WGAError as e:
        if e.code == 102:  # call not built yet
            do_call = False
            config["commands"] = ["maf2paf", "stat"]
        else:
            raise
    for _ in range(max(args.warmup - 1, 0)):
        step(1, d_text.data_ptr())
    sampler = ClockSampler(dev)
    sampler.start()
    # independent commands (every command tokenises the slab again, like three separate invocations) ...
    ms_indep = float("nan") if args.no_indep else timed(1, d_text.data_ptr(), args.steps, False)
    # ... and the job as the product runs it: one tokenisation per step, shared by the three commands
    from wgatools_b200._lib import WGA_F_REUSE_INDEX
    reuse["f"] = WGA_F_REUSE_INDEX
    step(1, d_text.data_ptr())
    own0, cub0 = ctx.launch_counts()
    ms_dev = timed(1, d_text.data_ptr(), args.steps, True)
    own1, cub1 = ctx.launch_counts()
    reuse["f"] = 0
    # e2e: host buffers, copies inside the timed region.
    # (1) the public pipeline: ONE asynchronous slab-wise upload per step (ResidentMaf = wga_text_open), the three
    #     commands run on every slab as it arrives, outputs copied back to pinned host memory inside each call;
    # (2) three independent host-buffer calls (the file crosses PCIe three times), kept as e2e_separate_calls.
    d2h = {"n": 0}

    def step_resident(_mem, ptr, _collect=False):
        acc = w.StatAcc(False)
        tot = 0
        with w.ResidentMaf((ptr, nbytes), ctx, args.slab_mb << 20) as rm:
            for p, n, fl in rm.slabs():
                tot += ctx.maf2paf_raw(p, n, 1, fl)[1]
                ctx.maf_stat_raw(p, n, 1, fl | WGA_F_REUSE_INDEX, acc)
                if do_call:
                    tot += ctx.maf_call_raw(p, n, 1, fl | WGA_F_REUSE_INDEX, opts)[1]
        d2h["n"] = tot
        return acc

    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e = ms_e2e_sep = float("nan")
    if not args.no_e2e:
        _step = step
        step = step_resident
        for _ in range(2):
            step(0, host_ptr)
        ms_e2e = timed(0, host_ptr, e2e_steps, False)
        step = _step
        step(0, host_ptr)
        ms_e2e_sep = timed(0, host_ptr, e2e_steps, False)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- stat totals across ranks: one NCCL all-reduce of the counter table
    if world > 1:
        from wgatools_b200 import dist as wdist
        acc = step(1, d_text.data_ptr())
        wdist.allreduce_stat(acc, dist)

    per_step = ms_dev / args.steps
    value = world * nbytes / (per_step * 1e-3) / 1e9
    e2e_value = world * nbytes / (ms_e2e / e2e_steps * 1e-3) / 1e9
    peak, peak_src = peaks()
    achieved = (walk["bytes"] / max(walk["launches"], 1)) / (walk["ms"] / max(walk["launches"], 1) * 1e-3) / 1e9 if walk["ms"] > 0 else 0.0
    line = {"metric": metric, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": d2h["n"],
                    "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                    "mode": f"ResidentMaf: one async upload per step in {args.slab_mb} MiB slabs, the three commands run "
                            "per slab while the next slab uploads; outputs D2H inside every call"},
            "e2e_separate_calls": {"value": world * nbytes / (ms_e2e_sep / e2e_steps * 1e-3) / 1e9, "unit": "GB/s",
                                   "h2d_bytes_per_step": nbytes * (3 if do_call else 2),
                                   "d2h_bytes_per_step": out_bytes["paf"] + out_bytes["vcf"],
                                   "mode": "three independent host-buffer calls (file uploaded three times)"},
            "value_independent_commands": {"value": world * nbytes / (ms_indep / args.steps * 1e-3) / 1e9, "unit": "GB/s",
                                           "ms_per_step": ms_indep / args.steps,
                                           "mode": "each of the three commands tokenises the slab itself (no WGA_F_REUSE_INDEX)"},
            "gpu_launches": own1 - own0, "cub_scan_calls": cub1 - cub0,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "k_walk_tma (dense column walk, TMA-staged; all launches of the timed steps)",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": walk_traffic(args.blocks) if args.config == 2 else None,
                         "launches": walk["launches"], "avg_launch_ms": walk["ms"] / max(walk["launches"], 1),
                         "alg_bytes_per_launch": walk["bytes"] / max(walk["launches"], 1)},
            "input_bytes_per_gpu": nbytes, "columns_per_gpu": ncols, "gen_seconds": t_gen, "numa_node": numa}
    if rank == 0:
        if world == 1:
            os.sched_setaffinity(0, all_cpus)   # the CPU baseline gets every host core back
            cores = os.cpu_count() or 1
            gbs, n, dt = run_cpu(host_ptr, nbytes, args.cpu_sample_mb * 1e6, cores, do_call)
            line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n / 1e6:.0f} MB of the same MAF ({dt:.1f} s); " + CPU_KIND_NOTE
                                              + f", -t {cores}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
