"""Seeded synthetic MAF / PAF generators (bench + test infrastructure; shapes from SURVEY.md 8(d))."""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libwga_synth.so")


class MafParams(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("n_blocks", ctypes.c_uint64),
                ("len_median", ctypes.c_double), ("len_sigma", ctypes.c_double),
                ("len_min", ctypes.c_uint64), ("len_max", ctypes.c_uint64),
                ("p_mis", ctypes.c_double), ("p_gap", ctypes.c_double), ("gap_mean", ctypes.c_double),
                ("p_sv", ctypes.c_double), ("sv_min", ctypes.c_uint64), ("sv_max", ctypes.c_uint64),
                ("p_neg", ctypes.c_double), ("p_bothgap", ctypes.c_double), ("p_lower", ctypes.c_double),
                ("n_chr", ctypes.c_uint32), ("chr_size", ctypes.c_uint64), ("threads", ctypes.c_int32)]


class PafParams(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("n_records", ctypes.c_uint64),
                ("ops_median", ctypes.c_double), ("ops_sigma", ctypes.c_double),
                ("ops_min", ctypes.c_uint64), ("ops_max", ctypes.c_uint64),
                ("eq_mean", ctypes.c_double), ("p_neg", ctypes.c_double),
                ("n_genomes", ctypes.c_uint32), ("n_chr", ctypes.c_uint32),
                ("chr_size", ctypes.c_uint64), ("m_only", ctypes.c_int32), ("threads", ctypes.c_int32)]


def build():
    src = os.path.join(_DIR, "wga_synth.c")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-s"])


_lib = None


def _L():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.wga_synth_maf_plan.argtypes = [ctypes.POINTER(MafParams), ctypes.POINTER(ctypes.c_uint64),
                                         ctypes.POINTER(ctypes.c_uint64)]
        L.wga_synth_maf_plan.restype = ctypes.c_void_p
        L.wga_synth_maf_fill.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.wga_synth_plan_free.argtypes = [ctypes.c_void_p]
        L.wga_synth_paf_plan.argtypes = [ctypes.POINTER(PafParams), ctypes.POINTER(ctypes.c_uint64)]
        L.wga_synth_paf_plan.restype = ctypes.c_void_p
        L.wga_synth_paf_fill.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.wga_synth_paf_plan_free.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


def maf_params(n_blocks, seed=20260926, config=2, threads=None, **over):
    """config 2: 2-genome MAF (median 2 kb blocks); config 3: human-vs-human (median 50 kb, SVs)."""
    threads = threads or os.cpu_count() or 1
    if config == 3:
        d = dict(len_median=50000.0, len_sigma=1.5, len_min=200, len_max=5_000_000, p_mis=1e-3,
                 p_gap=1.5e-4, gap_mean=3.0, p_sv=1e-5, sv_min=50, sv_max=100_000, p_neg=0.01,
                 p_bothgap=0.0, p_lower=0.0)
    else:
        d = dict(len_median=2000.0, len_sigma=1.0, len_min=50, len_max=200_000, p_mis=1e-3,
                 p_gap=2e-4, gap_mean=3.0, p_sv=0.0, sv_min=50, sv_max=100_000, p_neg=0.02,
                 p_bothgap=0.0, p_lower=0.0)
    d.update(over)
    return MafParams(seed=seed, n_blocks=n_blocks, n_chr=22, chr_size=250_000_000, threads=threads, **d)


def gen_maf_into(params, alloc):
    """alloc(nbytes) -> (address, keepalive).  Returns (keepalive, nbytes, ncols)."""
    L = _L()
    tb, tc = ctypes.c_uint64(), ctypes.c_uint64()
    plan = L.wga_synth_maf_plan(ctypes.byref(params), ctypes.byref(tb), ctypes.byref(tc))
    addr, keep = alloc(tb.value)
    L.wga_synth_maf_fill(plan, addr)
    L.wga_synth_plan_free(plan)
    return keep, tb.value, tc.value


def gen_maf(params) -> bytes:
    def alloc(n):
        a = np.empty(n, dtype=np.uint8)
        return a.ctypes.data, a
    arr, n, _ = gen_maf_into(params, alloc)
    return arr.tobytes()


def paf_params(n_records, seed=20260928, threads=None, **over):
    threads = threads or os.cpu_count() or 1
    d = dict(ops_median=150.0, ops_sigma=1.2, ops_min=1, ops_max=200_000, eq_mean=400.0, p_neg=0.5,
             n_genomes=16, n_chr=20, chr_size=250_000_000, m_only=0)
    d.update(over)
    return PafParams(seed=seed, n_records=n_records, threads=threads, **d)


def gen_paf_into(params, alloc):
    L = _L()
    tb = ctypes.c_uint64()
    plan = L.wga_synth_paf_plan(ctypes.byref(params), ctypes.byref(tb))
    addr, keep = alloc(tb.value)
    L.wga_synth_paf_fill(plan, addr)
    L.wga_synth_paf_plan_free(plan)
    return keep, tb.value


def gen_paf(params) -> bytes:
    def alloc(n):
        a = np.empty(n, dtype=np.uint8)
        return a.ctypes.data, a
    arr, n = gen_paf_into(params, alloc)
    return arr.tobytes()
