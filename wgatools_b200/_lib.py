"""Loads the in-tree CUDA library (wgatools_b200/libwga_b200.so) and declares the C ABI of include/wga_b200.h.

There is no CPU fallback: if the library is missing this module raises, and if no CUDA device is present
wga_ctx_create fails with WGA_E_CUDA.
"""
import ctypes
import os

_DIR = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("WGA_B200_LIB") or os.path.join(_DIR, "libwga_b200.so")  # override: A/B builds

WGA_MEM_HOST, WGA_MEM_DEVICE = 0, 1
WGA_F_FIRST_SLAB = 1
WGA_F_OUT_HOST = 2
WGA_F_REUSE_INDEX = 4
WGA_F_SHARE_WALK = 8


class Span(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("len", ctypes.c_uint64)]


class CallOpts(ctypes.Structure):
    _fields_ = [("snp", ctypes.c_int32), ("inv", ctypes.c_int32), ("svlen_cutoff", ctypes.c_uint64),
                ("chunk_size", ctypes.c_uint64), ("query_name", ctypes.c_char_p), ("query_match", ctypes.c_char_p)]


# wga_transport (include/wga_b200.h): host-buffer collective primitives provided by the caller
ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64)
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int)
SENDRECV_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p,
                               ctypes.c_uint64, ctypes.c_int)


class Transport(ctypes.Structure):
    _fields_ = [("user", ctypes.c_void_p), ("rank", ctypes.c_int), ("world", ctypes.c_int), ("allgather", ALLGATHER_FN),
                ("allreduce_u64", ALLREDUCE_FN), ("sendrecv", SENDRECV_FN)]


# every symbol include/wga_b200.h declares: name -> (restype, argtypes)
_P = ctypes.POINTER
ABI = {
    "wga_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, _P(ctypes.c_void_p)]),
    "wga_ctx_destroy": (None, [ctypes.c_void_p]),
    "wga_last_error": (ctypes.c_char_p, [ctypes.c_void_p, _P(ctypes.c_uint64)]),
    "wga_ctx_launch_counts": (None, [ctypes.c_void_p, _P(ctypes.c_uint64), _P(ctypes.c_uint64)]),
    "wga_ctx_nccl_calls": (ctypes.c_uint64, [ctypes.c_void_p]),
    "wga_nccl_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "wga_comm_init_nccl": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _P(ctypes.c_void_p)]),
    "wga_comm_from_nccl": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _P(ctypes.c_void_p)]),
    "wga_comm_from_transport": (ctypes.c_int, [_P(Transport), _P(ctypes.c_void_p)]),
    "wga_comm_destroy": (None, [ctypes.c_void_p]),
    "wga_comm_rank": (ctypes.c_int, [ctypes.c_void_p]),
    "wga_comm_world": (ctypes.c_int, [ctypes.c_void_p]),
    "wga_comm_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "wga_stat_allreduce": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "wga_ctx_last_walk": (None, [ctypes.c_void_p, _P(ctypes.c_float), _P(ctypes.c_uint64), _P(ctypes.c_uint32)]),
    "wga_host_alloc": (ctypes.c_void_p, [ctypes.c_uint64]),
    "wga_host_free": (None, [ctypes.c_void_p]),
    "wga_maf_safe_cut": (ctypes.c_uint64, [ctypes.c_void_p, ctypes.c_uint64]),
    "wga_text_open": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, _P(ctypes.c_void_p)]),
    "wga_text_slabs": (ctypes.c_uint32, [ctypes.c_void_p]),
    "wga_text_slab": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, _P(ctypes.c_void_p), _P(ctypes.c_uint64),
                                    _P(ctypes.c_uint32)]),
    "wga_text_whole": (ctypes.c_int, [ctypes.c_void_p, _P(ctypes.c_void_p), _P(ctypes.c_uint64)]),
    "wga_text_close": (None, [ctypes.c_void_p]),
    "wga_maf2paf": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32,
                                  ctypes.c_char_p, _P(Span)]),
    "wga_stat_acc_create": (ctypes.c_void_p, [ctypes.c_int]),
    "wga_stat_acc_destroy": (None, [ctypes.c_void_p]),
    "wga_maf_stat": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32,
                                   ctypes.c_char_p, ctypes.c_void_p]),
    "wga_paf_stat": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32,
                                   ctypes.c_void_p]),
    "wga_stat_acc_write": (ctypes.c_int, [ctypes.c_void_p, _P(Span)]),
    "wga_stat_acc_export": (ctypes.c_int, [ctypes.c_void_p, _P(Span)]),
    "wga_stat_acc_import": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]),
    "wga_stat_acc_npairs": (ctypes.c_uint64, [ctypes.c_void_p]),
    "wga_vcf_header": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, _P(Span)]),
    "wga_vcf_header_contigs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, _P(ctypes.c_char_p), _P(ctypes.c_uint64),
                                              ctypes.c_uint64, _P(Span)]),
    "wga_maf_sline_names": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, _P(Span)]),
    "wga_maf_call": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32,
                                   _P(CallOpts), _P(Span)]),
    "wga_fasta_load": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                     _P(ctypes.c_void_p)]),
    "wga_fasta_free": (None, [ctypes.c_void_p]),
    "wga_paf2maf_header": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, _P(Span)]),
    "wga_paf2maf": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32,
                                  ctypes.c_void_p, ctypes.c_void_p, _P(Span)]),
    "wga_paf_call": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, _P(Span)]),
    "wga_paf_pseudo": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32,
                                     ctypes.c_void_p, ctypes.c_char_p, _P(Span), _P(Span)]),
    "wga_paf_validate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32,
                                       ctypes.c_int, _P(Span), _P(Span)]),
    "wga_maf2chain": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, ctypes.c_char_p,
                                    ctypes.c_uint64, _P(Span), _P(ctypes.c_uint64)]),
    "wga_paf2chain": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64,
                                    _P(Span), _P(ctypes.c_uint64)]),
    "wga_chain2paf": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, _P(Span)]),
    "wga_chain2maf_header": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, _P(Span)]),
    "wga_chain2maf": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32,
                                    ctypes.c_void_p, ctypes.c_void_p, _P(Span)]),
    "wga_paf_filter": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64,
                                      ctypes.c_uint64, ctypes.c_void_p, _P(Span)]),
    "wga_maf_plotdata": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, ctypes.c_char_p,
                                        ctypes.c_uint64, _P(Span)]),
    "wga_paf_plotdata": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64,
                                        _P(Span)]),
    "wga_paf_cov": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, _P(Span)]),
    "wga_last_record_counts": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, _P(ctypes.c_uint64)]),
}

_lib = None


def load():
    """Returns the loaded library; raises (loudly) when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). wgatools_b200 has no CPU fallback.")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
