"""Host-side mirror of the reference's operator interface for the hot path.

Same names, argument meaning and error behaviour as the Rust functions they stand in for:
  maf2paf(mafreader, writer, query_name)          src/converter.rs:29-33
  stat_maf(reader, writer, each, query_name)      src/tools/stat.rs:61-66
  stat_paf(reader, writer, each)                  src/tools/stat.rs:87-91
  call_var_maf(mafreader, mafindex, writer, ...)  src/tools/caller.rs:42-54
Errors surface as WGAError whose str() is the reference's `impl Display for WGAError` text
(src/errors.rs:8-86) and whose .code is the C-ABI status.  All compute happens in the CUDA library
behind include/wga_b200.h; nothing here touches sequence bytes.
"""
import ctypes
import io
import os

from . import _lib
from ._lib import WGA_F_FIRST_SLAB, WGA_F_OUT_HOST, WGA_MEM_DEVICE, WGA_MEM_HOST, CallOpts, Span

SLAB_BYTES = 1 << 30  # host inputs larger than this are cut into slabs at s-line-group boundaries


class WGAError(Exception):
    """src/errors.rs:8-86"""

    def __init__(self, code, msg, where=0):
        super().__init__(msg)
        self.code = code
        self.where = where


class Context:
    """One per (host thread, GPU).  stream: an int cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device=0, stream=None):
        self.L = _lib.load()
        h = ctypes.c_void_p()
        rc = self.L.wga_ctx_create(int(device), ctypes.c_void_p(stream) if stream else None, ctypes.byref(h))
        if rc != 0:
            raise WGAError(rc, "wga_ctx_create failed (no CUDA device? wgatools_b200 has no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.wga_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            where = ctypes.c_uint64()
            msg = self.L.wga_last_error(self.h, ctypes.byref(where))
            raise WGAError(rc, (msg or b"").decode("utf-8", "replace"), where.value)

    def launch_counts(self):
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        self.L.wga_ctx_launch_counts(self.h, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def last_walk(self):
        ms, by, n = ctypes.c_float(), ctypes.c_uint64(), ctypes.c_uint32()
        self.L.wga_ctx_last_walk(self.h, ctypes.byref(ms), ctypes.byref(by), ctypes.byref(n))
        return ms.value, by.value, n.value

    # ---- raw entry points (ptr = host address or device address) -------------------------------
    def maf2paf_raw(self, ptr, n, mem, flags, query_name=None):
        sp = Span()
        self._check(self.L.wga_maf2paf(self.h, ptr, n, mem, flags, _opt(query_name), ctypes.byref(sp)))
        return sp.data, sp.len

    def maf_stat_raw(self, ptr, n, mem, flags, acc, query_name=None):
        self._check(self.L.wga_maf_stat(self.h, ptr, n, mem, flags, _opt(query_name), acc.h))

    def paf_stat_raw(self, ptr, n, mem, flags, acc):
        self._check(self.L.wga_paf_stat(self.h, ptr, n, mem, flags, acc.h))

    def maf_call_raw(self, ptr, n, mem, flags, opts):
        sp = Span()
        self._check(self.L.wga_maf_call(self.h, ptr, n, mem, flags, ctypes.byref(opts), ctypes.byref(sp)))
        return sp.data, sp.len

    def vcf_header_contigs(self, sample, contigs):
        """header with the ##contig lines of a <maf>.index: contigs = [(name, size)] of its isref entries (caller.rs:339-357)"""
        n = len(contigs)
        names = (ctypes.c_char_p * max(n, 1))(*[c[0].encode() if isinstance(c[0], str) else c[0] for c in contigs])
        sizes = (ctypes.c_uint64 * max(n, 1))(*[int(c[1]) for c in contigs])
        sp = Span()
        self._check(self.L.wga_vcf_header_contigs(self.h, _opt(sample), names, sizes, n, ctypes.byref(sp)))
        return ctypes.string_at(sp.data, sp.len)

    def maf_sline_names(self, ptr, n, mem, flags):
        """distinct s-line names of one slab (first-occurrence order): what a host evaluates --query-regex on"""
        sp = Span()
        self._check(self.L.wga_maf_sline_names(self.h, ptr, n, mem, flags, ctypes.byref(sp)))
        return ctypes.string_at(sp.data, sp.len).split(b"\n")[:-1] if sp.len else []

    def vcf_header(self, sample="sample"):
        sp = Span()
        self._check(self.L.wga_vcf_header(self.h, _opt(sample), ctypes.byref(sp)))
        return ctypes.string_at(sp.data, sp.len)

    def last_record_counts(self):
        n = ctypes.c_uint64()
        self._check(self.L.wga_last_record_counts(self.h, None, 0, ctypes.byref(n)))
        import numpy as np
        out = np.zeros((n.value, 11), dtype=np.uint64)
        if n.value:
            self._check(self.L.wga_last_record_counts(self.h, out.ctypes.data, n.value, ctypes.byref(n)))
        return out


class Fasta:
    """FASTA resident in HBM (stands in for rust_htslib::faidx::Reader, src/converter.rs:183-184)"""

    def __init__(self, ctx, fa: bytes, fai: bytes):
        self.ctx = ctx
        h = ctypes.c_void_p()
        ctx._check(ctx.L.wga_fasta_load(ctx.h, fa, len(fa), fai, len(fai), ctypes.byref(h)))
        self.h = h

    @classmethod
    def from_path(cls, ctx, path):
        # utils.rs:470-492: both the FASTA and its .fai must exist
        for p in (path, path + ".fai"):
            if not os.path.exists(p):
                raise WGAError(-1, f"File path `{p}` not exist")
        with open(path, "rb") as f, open(path + ".fai", "rb") as g:
            return cls(ctx, f.read(), g.read())

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.ctx.L.wga_fasta_free(self.h)
                self.h = None
        except Exception:
            pass


class StatAcc:
    def __init__(self, each=False):
        self.L = _lib.load()
        self.h = ctypes.c_void_p(self.L.wga_stat_acc_create(int(bool(each))))

    def write(self) -> bytes:
        sp = Span()
        rc = self.L.wga_stat_acc_write(self.h, ctypes.byref(sp))
        if rc:
            raise WGAError(rc, "wga_stat_acc_write failed")
        return ctypes.string_at(sp.data, sp.len) if sp.len else b""

    def write_span(self):
        """(address, length) of the TSV inside the accumulator (valid until the next call on it): what a host writes out"""
        sp = Span()
        rc = self.L.wga_stat_acc_write(self.h, ctypes.byref(sp))
        if rc:
            raise WGAError(rc, "wga_stat_acc_write failed")
        return sp.data, sp.len

    def export(self) -> bytes:
        sp = Span()
        rc = self.L.wga_stat_acc_export(self.h, ctypes.byref(sp))
        if rc:
            raise WGAError(rc, "wga_stat_acc_export failed")
        return ctypes.string_at(sp.data, sp.len)

    def import_(self, blob: bytes):
        rc = self.L.wga_stat_acc_import(self.h, blob, len(blob))
        if rc:
            raise WGAError(rc, "wga_stat_acc_import failed")

    def npairs(self):
        return self.L.wga_stat_acc_npairs(self.h)

    def __del__(self):
        try:
            if self.h:
                self.L.wga_stat_acc_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Comm:
    """Communicator for the one exchange step of the path, the stat all-reduce (include/wga_b200.h wga_comm_*).

    Comm.nccl(ctx, dist)  -- production: an ncclComm_t created INSIDE the C-ABI library on ctx's device; torch.distributed
                             only carries the 128-byte ncclUniqueId from rank 0 to the others (plumbing);
    Comm.host(dist)       -- the same C++ reduction over caller-provided host-buffer primitives (wga_transport), here
                             torch.distributed on CPU tensors (gloo): used by the CPU tests, world_size 2."""

    def __init__(self, L, h, keep=None):
        self.L, self.h, self._keep = L, h, keep

    @classmethod
    def nccl(cls, ctx, dist):
        import torch
        L = ctx.L
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = (ctypes.c_uint8 * 128)()
        if rank == 0:
            rc = L.wga_nccl_unique_id(ident)
            if rc:
                raise WGAError(rc, "wga_nccl_unique_id failed (libnccl.so.2 not found?)")
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0)
        h = ctypes.c_void_p()
        ctx._check(L.wga_comm_init_nccl(ctx.h, box[0], rank, world, ctypes.byref(h)))
        return cls(L, h)

    @classmethod
    def host(cls, dist):
        import numpy as np
        import torch
        L = _lib.load()
        rank, world = dist.get_rank(), dist.get_world_size()

        def view(ptr, n, dtype=np.uint8):
            return np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(ptr)).view(dtype)

        def allgather(_u, send, recv, n):
            try:
                mine = torch.from_numpy(view(send, n).copy())
                outs = [torch.empty(n, dtype=torch.uint8) for _ in range(world)]
                dist.all_gather(outs, mine)
                dst = view(recv, n * world)
                for r, t in enumerate(outs):
                    dst[r * n:(r + 1) * n] = t.numpy()
                return 0
            except Exception:
                return 104

        def allreduce(_u, buf, n, op):
            try:
                a = view(buf, n * 8, np.uint64)
                if op == 0:   # wrapping u64 sum == wrapping i64 sum
                    t = torch.from_numpy(a.view(np.int64).copy())
                    dist.all_reduce(t, op=dist.ReduceOp.SUM)
                    a[:] = t.numpy().view(np.uint64)
                else:         # unsigned min through the order-preserving map x ^ 2^63 -> i64
                    t = torch.from_numpy((a ^ np.uint64(1 << 63)).view(np.int64).copy())
                    dist.all_reduce(t, op=dist.ReduceOp.MIN)
                    a[:] = t.numpy().view(np.uint64) ^ np.uint64(1 << 63)
                return 0
            except Exception:
                return 104

        def sendrecv(_u, send, sb, dst, recv, rb, src):
            try:
                req = None
                if dst >= 0 and sb:
                    req = dist.isend(torch.from_numpy(view(send, sb).copy()), dst)
                if src >= 0 and rb:
                    t = torch.empty(rb, dtype=torch.uint8)
                    dist.recv(t, src)
                    view(recv, rb)[:] = t.numpy()
                if req is not None:
                    req.wait()
                return 0
            except Exception:
                return 104

        fns = (_lib.ALLGATHER_FN(allgather), _lib.ALLREDUCE_FN(allreduce), _lib.SENDRECV_FN(sendrecv))
        tr = _lib.Transport(None, rank, world, *fns)
        h = ctypes.c_void_p()
        rc = L.wga_comm_from_transport(ctypes.byref(tr), ctypes.byref(h))
        if rc:
            raise WGAError(rc, "wga_comm_from_transport failed")
        return cls(L, h, keep=fns)

    def allreduce_stat(self, acc):
        """collective; acc (merged mode) becomes the accumulator of the whole file on every rank"""
        rc = self.L.wga_stat_allreduce(self.h, acc.h)
        if rc:
            raise WGAError(rc, (self.L.wga_comm_last_error(self.h) or b"").decode("utf-8", "replace"))
        return acc

    def close(self):
        if getattr(self, "h", None):
            self.L.wga_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _opt(s):
    return None if s is None else (s if isinstance(s, bytes) else str(s).encode())


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _read_all(src) -> bytes:
    if isinstance(src, (bytes, bytearray, memoryview)):
        return bytes(src)
    if isinstance(src, (str, os.PathLike)):
        with open(src, "rb") as f:
            return f.read()
    return src.read()


class MAFReader:
    """src/parser/maf.rs:15-44.  new() consumes the first line as the header (maf.rs:25-36)."""

    def __init__(self, reader):
        self.data = _read_all(reader)
        nl = self.data.find(b"\n")
        self.header = (self.data if nl < 0 else self.data[:nl + 1]).decode("utf-8", "replace")

    @classmethod
    def from_path(cls, path):
        if not os.path.exists(path):
            raise WGAError(-1, f"File path `{path}` not exist")
        return cls(path)


class PAFReader:
    """src/parser/paf.rs:13-46"""

    def __init__(self, reader):
        self.data = _read_all(reader)


class ResidentMaf:
    """A MAF uploaded ONCE for several commands (the reference re-reads the file per command).

    The upload is asynchronous and slab-wise (include/wga_b200.h, wga_text_open): iterating `slabs()` hands out
    each slab as soon as it has arrived, so commands on slab k overlap the PCIe transfer of slab k+1.
    maf2paf / stat_maf / call_var_maf accept a ResidentMaf wherever they take a MAFReader.
    `data`: bytes-like, or (host_address, length) of caller-owned (ideally pinned) memory."""

    def __init__(self, data, ctx=None, slab_bytes=0):
        self.ctx = ctx or default_context()
        if isinstance(data, tuple):
            self._keep, base, n = None, int(data[0]), int(data[1])
        else:
            data = _read_all(data)
            self._keep, base = _host_ptr(data)
            n = len(data)
        h = ctypes.c_void_p()
        self.ctx._check(self.ctx.L.wga_text_open(self.ctx.h, ctypes.c_void_p(base), n, int(slab_bytes), ctypes.byref(h)))
        self.h = h
        self.nbytes = n

    def slabs(self):
        """yield (device_ptr, length, flags); flags carry WGA_F_OUT_HOST so outputs come back in host memory"""
        L = self.ctx.L
        for k in range(L.wga_text_slabs(self.h)):
            dev, ln, fl = ctypes.c_void_p(), ctypes.c_uint64(), ctypes.c_uint32()
            self.ctx._check(L.wga_text_slab(self.h, k, ctypes.byref(dev), ctypes.byref(ln), ctypes.byref(fl)))
            yield ctypes.c_void_p(dev.value), ln.value, fl.value | WGA_F_OUT_HOST

    def whole(self):
        dev, ln = ctypes.c_void_p(), ctypes.c_uint64()
        self.ctx._check(self.ctx.L.wga_text_whole(self.h, ctypes.byref(dev), ctypes.byref(ln)))
        return ctypes.c_void_p(dev.value), ln.value, WGA_F_FIRST_SLAB | WGA_F_OUT_HOST

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.wga_text_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _maf_slabs(src, ctx):
    """(ptr, length, mem, flags) of every slab of a MAFReader (host, cut here) or a ResidentMaf (device)"""
    if isinstance(src, ResidentMaf):
        if src.ctx is not ctx:
            raise WGAError(101, "ResidentMaf belongs to another Context")
        for ptr, n, flags in src.slabs():
            yield ptr, n, WGA_MEM_DEVICE, flags
        return
    keep, base = _host_ptr(src.data)
    for off, n, flags in _slabs(src.data, ctx.L):
        yield ctypes.c_void_p(base + off), n, WGA_MEM_HOST, flags


def _slabs(data: bytes, L):
    """yield (offset, length, flags): slabs cut right before a line that is not an s-line"""
    n = len(data)
    if n <= SLAB_BYTES:
        yield 0, n, WGA_F_FIRST_SLAB
        return
    buf = ctypes.c_char_p(data)
    base = ctypes.cast(buf, ctypes.c_void_p).value
    off = 0
    first = True
    while off < n:
        want = min(SLAB_BYTES, n - off)
        cut = want
        if off + want < n:
            cut = L.wga_maf_safe_cut(ctypes.c_void_p(base + off), want)
            if cut == 0:  # an s-line group larger than a slab: take the remainder in one piece
                cut = n - off
        yield off, cut, (WGA_F_FIRST_SLAB if first else 0)
        first = False
        off += cut


def _host_ptr(data: bytes):
    buf = ctypes.c_char_p(data)
    return buf, ctypes.cast(buf, ctypes.c_void_p).value


def maf2paf(mafreader: MAFReader, writer, query_name=None, ctx=None):
    """src/converter.rs:29-54: every MAF record -> one PAF row (file order, i.e. the `-t 1` order)."""
    ctx = ctx or default_context()
    for p, n, mem, flags in _maf_slabs(mafreader, ctx):
        ptr, ln = ctx.maf2paf_raw(p, n, mem, flags, query_name)
        if ln:
            writer.write(ctypes.string_at(ptr, ln))


def stat_maf(reader: MAFReader, writer, each=False, query_name=None, ctx=None):
    """src/tools/stat.rs:61-84"""
    ctx = ctx or default_context()
    acc = StatAcc(each)
    for p, n, mem, flags in _maf_slabs(reader, ctx):
        ctx.maf_stat_raw(p, n, mem, flags, acc, query_name)
    writer.write(acc.write())


def stat_paf(reader: PAFReader, writer, each=False, ctx=None):
    """src/tools/stat.rs:87-105"""
    ctx = ctx or default_context()
    acc = StatAcc(each)
    keep, base = _host_ptr(reader.data)
    ctx.paf_stat_raw(ctypes.c_void_p(base), len(reader.data), WGA_MEM_HOST, WGA_F_FIRST_SLAB, acc)
    writer.write(acc.write())


def call_var_maf(mafreader: MAFReader, mafindex, writer, if_snp=False, if_inv=False, svlen_cutoff=50,
                 _between=False, sample=None, query_name=None, query_regex=None, chunk_size=None, ctx=None):
    """src/tools/caller.rs:42-157.
    mafindex: the parsed <maf>.index (tools/index.rs:78-86: {name: {"size": .., "isref": ..}}) or None -- its isref
    entries become the header's ##contig lines (caller.rs:339-357).
    query_regex: a pattern string or compiled pattern.  The regex engine stays on the host, as it would in the Rust
    wrapper: the slab's distinct s-line names come back from the library, the host keeps the names the
    regex matches (`Regex::is_match_at(name, 0)` = an unanchored search, maf.rs:267-271; Python's `re.search` stands in
    for the Rust crate here) and the library picks, per block, the first s-line whose name is in that list."""
    ctx = ctx or default_context()
    if mafindex is not None:
        contigs = [(name, item["size"]) for name, item in mafindex.items() if item.get("isref")]
        writer.write(ctx.vcf_header_contigs(sample or "sample", contigs))
    else:
        writer.write(ctx.vcf_header(sample or "sample"))
    rx = None
    if query_name is None and query_regex is not None:
        import re
        rx = re.compile(query_regex.encode() if isinstance(query_regex, str) else query_regex) if not hasattr(query_regex, "search") \
            else query_regex
    for p, n, mem, flags in _maf_slabs(mafreader, ctx):
        match = None
        if rx is not None:
            names = ctx.maf_sline_names(p, n, mem, flags)
            hit = [nm for nm in names if rx.search(nm if isinstance(rx.pattern, bytes) else nm.decode("utf-8", "replace"))]
            match = b"".join(nm + b"\n" for nm in hit)
        opts = CallOpts(snp=int(bool(if_snp)), inv=int(bool(if_inv)), svlen_cutoff=int(svlen_cutoff),
                        chunk_size=int(1000000 if chunk_size is None else chunk_size), query_name=_opt(query_name),
                        query_match=match)
        ptr, ln = ctx.maf_call_raw(p, n, mem, flags, opts)
        if ln:
            writer.write(ctypes.string_at(ptr, ln))


def paf2maf(pafreader: PAFReader, writer, t_fa_path, q_fa_path, ctx=None, target=None, query=None):
    """src/converter.rs:176-265.  target/query: preloaded Fasta objects (else loaded from the two paths)."""
    ctx = ctx or default_context()
    target = target or Fasta.from_path(ctx, t_fa_path)
    query = query or Fasta.from_path(ctx, q_fa_path)
    sp = Span()
    ctx._check(ctx.L.wga_paf2maf_header(ctx.h, _opt(t_fa_path), _opt(q_fa_path), ctypes.byref(sp)))
    writer.write(ctypes.string_at(sp.data, sp.len))
    keep, base = _host_ptr(pafreader.data)
    ctx._check(ctx.L.wga_paf2maf(ctx.h, ctypes.c_void_p(base), len(pafreader.data), WGA_MEM_HOST, WGA_F_FIRST_SLAB, target.h,
                                 query.h, ctypes.byref(sp)))
    if sp.len:
        writer.write(ctypes.string_at(sp.data, sp.len))


def call_var_paf(pafreader: PAFReader, t_fa_path, q_fa_path, writer, if_snp=False, svlen_cutoff=50, _between=False,
                 sample=None, ctx=None, target=None, query=None):
    """src/tools/caller.rs:268-302.  Nothing is written when the call fails (the reference writes the header after
    all records were processed).  target/query: preloaded Fasta objects (else loaded from the two paths)."""
    ctx = ctx or default_context()
    target = target or Fasta.from_path(ctx, t_fa_path)
    query = query or Fasta.from_path(ctx, q_fa_path)
    keep, base = _host_ptr(pafreader.data)
    sp = Span()
    ctx._check(ctx.L.wga_paf_call(ctx.h, ctypes.c_void_p(base), len(pafreader.data), WGA_MEM_HOST, WGA_F_FIRST_SLAB, target.h,
                                  query.h, int(bool(if_snp)), int(svlen_cutoff), ctypes.byref(sp)))
    body = ctypes.string_at(sp.data, sp.len) if sp.len else b""
    writer.write(ctx.vcf_header(sample or "sample"))
    writer.write(body)


def pseudo_maf_files(reader: PAFReader, fa=None, target=None, ctx=None) -> dict:
    """{target name (bytes): bytes of {out_dir}/{target}.maf} in first-occurrence order; fa: a Fasta or None"""
    ctx = ctx or default_context()
    keep, base = _host_ptr(reader.data)
    out, index = Span(), Span()
    ctx._check(ctx.L.wga_paf_pseudo(ctx.h, ctypes.c_void_p(base), len(reader.data), WGA_MEM_HOST, WGA_F_FIRST_SLAB,
                                    fa.h if fa is not None else None, _opt(target), ctypes.byref(out), ctypes.byref(index)))
    files, pos = {}, 0
    idx = ctypes.string_at(index.data, index.len) if index.len else b""
    for line in idx.splitlines():
        name, n = line.rsplit(b"\t", 1)
        files[name] = ctypes.string_at(out.data + pos, int(n))
        pos += int(n)
    return files


def generate_pesudo_maf(reader: PAFReader, out_dir, fa_path=None, target=None, ctx=None, fa=None):
    """src/tools/pseudomaf.rs:18-75 (sic): one file {out_dir}/{target}.maf per target"""
    ctx = ctx or default_context()
    if fa is None and fa_path is not None:
        fa = Fasta.from_path(ctx, fa_path)
    for name, data in pseudo_maf_files(reader, fa, target, ctx).items():
        with open(os.path.join(out_dir, name.decode("utf-8", "replace") + ".maf"), "wb") as f:
            f.write(data)


def maf2chain(mafreader, writer, query_name=None, ctx=None):
    """src/converter.rs:57-92: one chain per MAF record; chain ids enumerate the records of the file"""
    ctx = ctx or default_context()
    base = 0
    for p, n, mem, flags in _maf_slabs(mafreader, ctx):
        sp, nrec = Span(), ctypes.c_uint64()
        ctx._check(ctx.L.wga_maf2chain(ctx.h, p, n, mem, flags, _opt(query_name), base, ctypes.byref(sp), ctypes.byref(nrec)))
        if sp.len:
            writer.write(ctypes.string_at(sp.data, sp.len))
        base += nrec.value


def paf2chain(pafreader: PAFReader, writer, ctx=None):
    """src/converter.rs:148-173"""
    ctx = ctx or default_context()
    keep, base = _host_ptr(pafreader.data)
    sp = Span()
    ctx._check(ctx.L.wga_paf2chain(ctx.h, ctypes.c_void_p(base), len(pafreader.data), WGA_MEM_HOST, WGA_F_FIRST_SLAB, 0, ctypes.byref(sp),
                                   None))
    if sp.len:
        writer.write(ctypes.string_at(sp.data, sp.len))


class ChainReader:
    """src/parser/chain.rs:17-44: the whole file is read into memory and parsed record by record"""

    def __init__(self, reader):
        self.data = _read_all(reader)


def chain2paf(chainreader: ChainReader, writer, ctx=None):
    """src/converter.rs:391-416: one PAF row (12 columns + cg:Z:) per chain, file order"""
    ctx = ctx or default_context()
    keep, base = _host_ptr(chainreader.data)
    sp = Span()
    ctx._check(ctx.L.wga_chain2paf(ctx.h, ctypes.c_void_p(base), len(chainreader.data), WGA_MEM_HOST, WGA_F_FIRST_SLAB, ctypes.byref(sp)))
    if sp.len:
        writer.write(ctypes.string_at(sp.data, sp.len))


def chain2maf(chainreader: ChainReader, writer, t_fa_path, q_fa_path, ctx=None, target=None, query=None):
    """src/converter.rs:268-357.  target/query: preloaded Fasta objects (else loaded from the two paths)."""
    ctx = ctx or default_context()
    target = target or Fasta.from_path(ctx, t_fa_path)
    query = query or Fasta.from_path(ctx, q_fa_path)
    sp = Span()
    ctx._check(ctx.L.wga_chain2maf_header(ctx.h, _opt(t_fa_path), _opt(q_fa_path), ctypes.byref(sp)))
    writer.write(ctypes.string_at(sp.data, sp.len))
    keep, base = _host_ptr(chainreader.data)
    ctx._check(ctx.L.wga_chain2maf(ctx.h, ctypes.c_void_p(base), len(chainreader.data), WGA_MEM_HOST, WGA_F_FIRST_SLAB, target.h,
                                   query.h, ctypes.byref(sp)))
    if sp.len:
        writer.write(ctypes.string_at(sp.data, sp.len))


def pafcov(reader: PAFReader, writer, ctx=None):
    """src/tools/pafcov.rs:13-60: per-base BED coverage of every target"""
    ctx = ctx or default_context()
    keep, base = _host_ptr(reader.data)
    sp = Span()
    ctx._check(ctx.L.wga_paf_cov(ctx.h, ctypes.c_void_p(base), len(reader.data), WGA_MEM_HOST, WGA_F_FIRST_SLAB, ctypes.byref(sp)))
    if sp.len:
        writer.write(ctypes.string_at(sp.data, sp.len))


PLOTDATA_CSV_HEADER = b"ref_start,ref_end,query_start,query_end,cigar,ref_chro,query_chro\n"


def generate_maf_basedata(reader: MAFReader, writer, cutoff=0, query_name=None, ctx=None):
    """src/tools/dotplot.rs:358-379 + render_output(.., Csv): the base-level plot segments of every record as csv"""
    ctx = ctx or default_context()
    keep, base = _host_ptr(reader.data)
    sp = Span()
    ctx._check(ctx.L.wga_maf_plotdata(ctx.h, ctypes.c_void_p(base), len(reader.data), WGA_MEM_HOST, WGA_F_FIRST_SLAB, _opt(query_name),
                                      cutoff, ctypes.byref(sp)))
    if sp.len:
        writer.write(PLOTDATA_CSV_HEADER)
        writer.write(ctypes.string_at(sp.data, sp.len))


def generate_paf_basedata(reader: PAFReader, writer, cutoff=0, ctx=None):
    """src/tools/dotplot.rs:338-356 + render_output(.., Csv)"""
    ctx = ctx or default_context()
    keep, base = _host_ptr(reader.data)
    sp = Span()
    ctx._check(ctx.L.wga_paf_plotdata(ctx.h, ctypes.c_void_p(base), len(reader.data), WGA_MEM_HOST, WGA_F_FIRST_SLAB, cutoff,
                                      ctypes.byref(sp)))
    if sp.len:
        writer.write(PLOTDATA_CSV_HEADER)
        writer.write(ctypes.string_at(sp.data, sp.len))


def _paf_filter(reader, writer, min_block_size, min_query_size, min_align_size, ctx):
    ctx = ctx or default_context()
    keep, base = _host_ptr(reader.data)
    sp = Span()
    cut = None if min_align_size is None else ctypes.byref(ctypes.c_uint64(min_align_size))
    ctx._check(ctx.L.wga_paf_filter(ctx.h, ctypes.c_void_p(base), len(reader.data), WGA_MEM_HOST, WGA_F_FIRST_SLAB, min_block_size,
                                    min_query_size, cut, ctypes.byref(sp)))
    if sp.len:
        writer.write(ctypes.string_at(sp.data, sp.len))


def filter_paf(reader: PAFReader, writer, min_block_size, min_query_size, ctx=None):
    """src/tools/filter.rs:41-63: drop records with target_end - target_start < min_block_size or query_length < min_query_size"""
    _paf_filter(reader, writer, min_block_size, min_query_size, None, ctx)


def filter_paf_align_pair(reader: PAFReader, writer, filt_align_size, ctx=None):
    """src/tools/filter.rs:108-160 (`filter -f paf -a`): keep the records of (query, target) pairs whose summed
    target_end - target_start reaches filt_align_size"""
    _paf_filter(reader, writer, 0, 0, filt_align_size, ctx)


def parallel_validatepaf(reader: PAFReader, writer, fix_writer=None, fix_flag=None, ctx=None):
    """src/tools/validate.rs:44-71: report -> writer; with a fix_writer the corrected records -> fix_writer"""
    ctx = ctx or default_context()
    fix = fix_writer is not None if fix_flag is None else bool(fix_flag)
    keep, base = _host_ptr(reader.data)
    rep, fx = Span(), Span()
    ctx._check(ctx.L.wga_paf_validate(ctx.h, ctypes.c_void_p(base), len(reader.data), WGA_MEM_HOST, WGA_F_FIRST_SLAB, int(fix),
                                      ctypes.byref(rep), ctypes.byref(fx)))
    writer.write(ctypes.string_at(rep.data, rep.len))
    if fix_writer is not None and fx.len:
        fix_writer.write(ctypes.string_at(fx.data, fx.len))


def run_bytes(fn, *args, **kw) -> bytes:
    """convenience: run one of the functions above into a bytes object (the writer argument is inserted
    where the reference signature has it: 2nd for maf2paf/stat_*, 3rd for call_var_maf, 4th for call_var_paf)"""
    w = io.BytesIO()
    pos = 2 if fn is call_var_maf else 3 if fn is call_var_paf else 1
    fn(*args[:pos], w, *args[pos:], **kw)
    return w.getvalue()
