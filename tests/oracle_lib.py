"""ctypes binding of the CPU oracle (oracle/libwga_oracle.so).

Test infrastructure only: nothing under wgatools_b200/ may import this module.
"""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_SO = os.path.join(ORACLE_DIR, "libwga_oracle.so")


class OrcBuf(ctypes.Structure):
    _fields_ = [("p", ctypes.c_void_p), ("n", ctypes.c_size_t), ("cap", ctypes.c_size_t)]


NAME_PRED = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)


class OrcCounts(ctypes.Structure):
    _fields_ = [("c", ctypes.c_uint64 * 11)]


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.msg = msg


def build():
    src = os.path.join(ORACLE_DIR, "wga_oracle.c")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.orc_buf_free.argtypes = [ctypes.POINTER(OrcBuf)]
        L.orc_maf2paf.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p,
                                  ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_stat_maf.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p,
                                   ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_stat_paf.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                                   ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_paf_filter.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64,
                                     ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_maf_plotdata.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_uint64, ctypes.POINTER(OrcBuf),
                                       ctypes.c_char_p, ctypes.c_size_t]
        L.orc_paf_plotdata.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.POINTER(OrcBuf), ctypes.c_char_p,
                                       ctypes.c_size_t]
        L.orc_cs_to_cigar.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_call_maf.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_uint64, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint64,
                                   ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_call_maf_ex.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_char_p,
                                      ctypes.c_char_p, NAME_PRED, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_uint64), ctypes.c_size_t,
                                      ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_maf2paf_mt.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_int,
                                     ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_stat_maf_mt.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                      ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_paf2maf.argtypes = [ctypes.c_char_p, ctypes.c_size_t] * 5 + [ctypes.c_char_p, ctypes.c_char_p,
                                                                             ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_call_paf.argtypes = [ctypes.c_char_p, ctypes.c_size_t] * 5 + [ctypes.c_int, ctypes.c_uint64, ctypes.c_char_p,
                                                                              ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_paf_pseudo.argtypes = [ctypes.c_char_p, ctypes.c_size_t] * 3 + [ctypes.c_char_p, ctypes.POINTER(OrcBuf),
                                                                                ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_paf_validate.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(OrcBuf), ctypes.POINTER(OrcBuf),
                                       ctypes.c_char_p, ctypes.c_size_t]
        L.orc_maf2chain.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.POINTER(OrcBuf), ctypes.c_char_p,
                                    ctypes.c_size_t]
        L.orc_paf2chain.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_pafcov.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_chain2paf.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_chain2maf.argtypes = [ctypes.c_char_p, ctypes.c_size_t] * 5 + [ctypes.c_char_p, ctypes.c_char_p,
                                                                               ctypes.POINTER(OrcBuf), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_parse_paf_cigar.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                                          ctypes.POINTER(OrcCounts), ctypes.c_char_p, ctypes.c_size_t]
        L.orc_parse_maf_seq_to_cigar.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t,
                                                 ctypes.c_int, ctypes.POINTER(OrcCounts),
                                                 ctypes.POINTER(OrcBuf)]
        L.orc_parse_maf_seq_to_cigar.restype = None
        L.orc_natord_compare.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        L.orc_fmt_f32.argtypes = [ctypes.c_float, ctypes.c_char_p]
        L.orc_fmt_f32.restype = None
        L.orc_maf_record_counts.argtypes = [ctypes.c_char_p, ctypes.c_size_t,
                                            ctypes.POINTER(OrcCounts), ctypes.c_size_t]
        L.orc_maf_record_counts.restype = ctypes.c_int64
        _lib = L
    return _lib


def _take(buf):
    data = ctypes.string_at(buf.p, buf.n) if buf.n else b""
    lib().orc_buf_free(ctypes.byref(buf))
    return data


def _run(fn, *args):
    buf = OrcBuf()
    err = ctypes.create_string_buffer(512)
    rc = fn(*args, ctypes.byref(buf), err, 512)
    data = _take(buf)
    if rc != 0:
        raise OracleError(rc, err.value.decode("utf-8", "replace"))
    return data


def _opt(s):
    return None if s is None else (s if isinstance(s, bytes) else s.encode())


def maf2paf(text: bytes, query_name=None) -> bytes:
    return _run(lib().orc_maf2paf, text, len(text), _opt(query_name))


def stat_maf(text: bytes, each=False, query_name=None) -> bytes:
    return _run(lib().orc_stat_maf, text, len(text), int(each), _opt(query_name))


def maf2paf_mt(text: bytes, threads, query_name=None) -> bytes:
    return _run(lib().orc_maf2paf_mt, text, len(text), _opt(query_name), int(threads))


def stat_maf_mt(text: bytes, threads, each=False, query_name=None) -> bytes:
    return _run(lib().orc_stat_maf_mt, text, len(text), int(each), _opt(query_name), int(threads))


def stat_paf(text: bytes, each=False) -> bytes:
    return _run(lib().orc_stat_paf, text, len(text), int(each))


def paf_filter(text: bytes, min_block=0, min_query=0, min_align=None) -> bytes:
    return _run(lib().orc_paf_filter, text, len(text), min_block, min_query, int(min_align is not None), min_align or 0)


def maf_plotdata(text: bytes, cutoff=0, query_name=None) -> bytes:
    return _run(lib().orc_maf_plotdata, text, len(text), _opt(query_name), cutoff)


def paf_plotdata(text: bytes, cutoff=0) -> bytes:
    return _run(lib().orc_paf_plotdata, text, len(text), cutoff)


def cs_to_cigar(cs: bytes) -> bytes:
    return _run(lib().orc_cs_to_cigar, cs, len(cs))


def call_maf(text: bytes, snp=False, inv=False, svlen=50, sample="sample", query_name=None,
             chunk_size=1000000) -> bytes:
    return _run(lib().orc_call_maf, text, len(text), int(snp), int(inv), svlen, _opt(sample),
                _opt(query_name), chunk_size)


def call_maf_ex(text: bytes, snp=False, inv=False, svlen=50, sample="sample", query_name=None, query_regex=None,
                chunk_size=1000000, mafindex=None) -> bytes:
    """call with --query-regex (Python's re.search stands in for Regex::is_match_at(name, 0)) and a parsed <maf>.index"""
    import re
    rx = re.compile(query_regex) if query_regex is not None else None
    pred = NAME_PRED(lambda p, n, ud: 1 if rx.search(ctypes.string_at(p, n)) else 0) if rx is not None else NAME_PRED()
    contigs = [(k, v["size"]) for k, v in mafindex.items() if v.get("isref")] if mafindex is not None else []
    names = (ctypes.c_char_p * max(len(contigs), 1))(*[c[0].encode() for c in contigs])
    sizes = (ctypes.c_uint64 * max(len(contigs), 1))(*[c[1] for c in contigs])
    return _run(lib().orc_call_maf_ex, text, len(text), int(snp), int(inv), svlen, _opt(sample), _opt(query_name), pred, None,
                chunk_size, 1 if mafindex is not None else 0, names, sizes, len(contigs))


def paf2maf(paf: bytes, t_fa: bytes, t_fai: bytes, q_fa: bytes, q_fai: bytes, t_path="t.fa", q_path="q.fa") -> bytes:
    return _run(lib().orc_paf2maf, paf, len(paf), t_fa, len(t_fa), t_fai, len(t_fai), q_fa, len(q_fa), q_fai, len(q_fai),
                _opt(t_path), _opt(q_path))


def call_paf(paf: bytes, t_fa: bytes, t_fai: bytes, q_fa: bytes, q_fai: bytes, snp=False, svlen=50, sample="sample") -> bytes:
    return _run(lib().orc_call_paf, paf, len(paf), t_fa, len(t_fa), t_fai, len(t_fai), q_fa, len(q_fa), q_fai, len(q_fai),
                int(snp), int(svlen), _opt(sample))


def paf_pseudo(paf: bytes, fa: bytes = None, fai: bytes = None, target=None) -> dict:
    """{target name: bytes of {out_dir}/{target}.maf}, in first-occurrence order"""
    out, index = OrcBuf(), OrcBuf()
    err = ctypes.create_string_buffer(512)
    rc = lib().orc_paf_pseudo(paf, len(paf), fa, len(fa) if fa is not None else 0, fai, len(fai) if fai is not None else 0,
                              _opt(target), ctypes.byref(out), ctypes.byref(index), err, 512)
    data, idx = _take(out), _take(index)
    if rc != 0:
        raise OracleError(rc, err.value.decode("utf-8", "replace"))
    files, pos = {}, 0
    for line in idx.splitlines():
        name, n = line.rsplit(b"\t", 1)
        files[name] = data[pos:pos + int(n)]
        pos += int(n)
    return files


def paf_validate(paf: bytes, fix=False):
    """(report text, fixed PAF or b"")"""
    rep, fx = OrcBuf(), OrcBuf()
    err = ctypes.create_string_buffer(512)
    rc = lib().orc_paf_validate(paf, len(paf), int(fix), ctypes.byref(rep), ctypes.byref(fx), err, 512)
    a, b = _take(rep), _take(fx)
    if rc != 0:
        raise OracleError(rc, err.value.decode("utf-8", "replace"))
    return a, b


def maf2chain(text: bytes, query_name=None) -> bytes:
    return _run(lib().orc_maf2chain, text, len(text), _opt(query_name))


def paf2chain(text: bytes) -> bytes:
    return _run(lib().orc_paf2chain, text, len(text))


def chain2paf(text: bytes) -> bytes:
    return _run(lib().orc_chain2paf, text, len(text))


def chain2maf(text: bytes, t_fa: bytes, t_fai: bytes, q_fa: bytes, q_fai: bytes, t_path="t.fa", q_path="q.fa") -> bytes:
    return _run(lib().orc_chain2maf, text, len(text), t_fa, len(t_fa), t_fai, len(t_fai), q_fa, len(q_fa), q_fai, len(q_fai),
                _opt(t_path), _opt(q_path))


def pafcov(text: bytes) -> bytes:
    return _run(lib().orc_pafcov, text, len(text))


def make_fai(fa: bytes) -> bytes:
    """samtools-faidx-style index of an in-memory FASTA (uniform line width per record assumed)"""
    out, pos = [], 0
    lines = fa.split(b"\n")
    i = 0
    while i < len(lines):
        l = lines[i]
        if l.startswith(b">"):
            name = l[1:].split()[0]
            pos += len(l) + 1
            off, n, lb, lw = pos, 0, 0, 0
            i += 1
            while i < len(lines) and not lines[i].startswith(b">"):
                if lines[i]:
                    if lb == 0:
                        lb, lw = len(lines[i]), len(lines[i]) + 1
                    n += len(lines[i])
                pos += len(lines[i]) + 1
                i += 1
            out.append(b"%s\t%d\t%d\t%d\t%d" % (name, n, off, lb, lw))
        else:
            pos += len(l) + 1
            i += 1
    return b"\n".join(out) + b"\n"


def paf_cigar_counts(cg: bytes, neg=False):
    c = OrcCounts()
    err = ctypes.create_string_buffer(512)
    rc = lib().orc_parse_paf_cigar(cg, len(cg), int(neg), ctypes.byref(c), err, 512)
    if rc:
        raise OracleError(rc, err.value.decode())
    return list(c.c)


def maf_seq_to_cigar(t: bytes, q: bytes, neg=False):
    c = OrcCounts()
    buf = OrcBuf()
    n = min(len(t), len(q))
    lib().orc_parse_maf_seq_to_cigar(t, q, n, int(neg), ctypes.byref(c), ctypes.byref(buf))
    return list(c.c), _take(buf)


def natord(a: str, b: str) -> int:
    return lib().orc_natord_compare(a.encode(), b.encode())


def fmt_f32(v: float) -> str:
    out = ctypes.create_string_buffer(48)
    lib().orc_fmt_f32(v, out)
    return out.value.decode()


def maf_record_counts(text: bytes, cap: int):
    arr = (OrcCounts * max(cap, 1))()
    n = lib().orc_maf_record_counts(text, len(text), arr, cap)
    if n < 0:
        raise OracleError(-n, "maf_record_counts")
    return [list(arr[i].c) for i in range(min(n, cap))]
