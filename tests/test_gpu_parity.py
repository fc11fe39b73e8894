"""GPU parity tests: the CUDA path (through the C ABI / the host mirror) against the CPU oracle on the same
inputs -- byte-identical PAF / TSV, identical integers.  Run on the B200 box: pytest -m gpu."""
import ctypes
import io
import os
import random
import re

import numpy as np
import pytest

import bench_support as bs
import oracle_lib as orc

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
TEST_MAF = open(os.path.join(G, "ref_test.maf"), "rb").read()


@pytest.fixture(scope="module")
def w():
    import __graft_entry__ as g
    g.build()
    import wgatools_b200 as w
    return w


@pytest.fixture(scope="module")
def ctx(w):
    return w.Context(0)


def paf(w, ctx, text, qn=None):
    return w.run_bytes(w.maf2paf, w.MAFReader(text), qn, ctx=ctx)


def stat(w, ctx, text, each=False, qn=None):
    return w.run_bytes(w.stat_maf, w.MAFReader(text), each, qn, ctx=ctx)


def check_both(w, ctx, text, qn=None):
    assert paf(w, ctx, text, qn) == orc.maf2paf(text, qn)
    assert stat(w, ctx, text, False, qn) == orc.stat_maf(text, False, qn)
    assert stat(w, ctx, text, True, qn) == orc.stat_maf(text, True, qn)


def test_golden_test_maf(w, ctx):
    out = paf(w, ctx, TEST_MAF).decode()
    assert out.endswith("cg:Z:109=1D243=1X12=1X138=1X177=1X31=1X133=8I18=1X100=2X7=1X22=\n")
    check_both(w, ctx, TEST_MAF)
    c = ctx.last_record_counts()
    assert c.tolist() == [[990, 9, 1, 8, 1, 1, 0, 0, 0, 0, 0]]


def _maf(blocks, header=b"##maf version=1\n"):
    out = [header]
    for lines in blocks:
        out.append(b"a score=0\n")
        out += [l + b"\n" for l in lines]
        out.append(b"\n")
    return b"".join(out)


def test_empty_and_header_only(w, ctx):
    for t in (b"", b"##maf\n", b"##maf", b"\n\n\n", b"##maf\na score=0\n\n"):
        assert paf(w, ctx, t) == orc.maf2paf(t) == b""
        assert stat(w, ctx, t) == orc.stat_maf(t) == b""


def test_reader_quirks(w, ctx):
    blk = [b"s t 0 4 + 100 ACGT", b"s q 0 4 + 100 ACGA"]
    cases = [
        b"s t 0 4 + 100 AAAA\n" + b"\n".join(blk) + b"\n",                       # Q9 first line swallowed
        _maf([blk]) + b"\n".join(blk) + b"\ni something\n" + b"\n".join(blk),     # no trailing newline
        _maf([[b"s t 0 4 + 100 A-cGTTT", b"s q 0 4 + 100 A-CG"]]),                # Q14 zip, Q1, Q2
        b"#h\r\ns t +5 4 + 100 ACGT\r\ns q 0 4 - 100 ACGT\r\n",                   # '+' ints, CRLF, '-' strand
        _maf([[b"s\tt.x\t0\t4\t+\t100\tACGT", b"s  q\"y   0 4 + 100 AC-T  "]]),   # tabs, runs of blanks, quote
        _maf([[b"s t 0 4 + 100 ACGT", b"s q 0 4 + 100 ACGT", b"s r 0 4 + 100 AAAA"]]),
        _maf([[b"s t 0 2 + 100 ----AC", b"s q 0 4 + 100 GGTTAC"]]),               # leading insertion
        _maf([[b"s t 0 6 + 100 ACGTAC", b"s q 0 2 + 100 AC----"]]),               # trailing deletion
        _maf([[b"s t 0 1 + 100 A", b"s q 0 1 + 100 C"]]),
        _maf([[b"s t 0 0 + 100 -", b"s q 0 0 + 100 -"]]),
    ]
    for t in cases:
        check_both(w, ctx, t)


def test_errors_match_oracle(w, ctx):
    bad = [
        (_maf([[b"s t 0 4 + 100 ACGT X", b"s q 0 4 + 100 ACGT"]]), 4),
        (_maf([[b"s t 0 4 + 100", b"s q 0 4 + 100 ACGT"]]), 3),
        (_maf([[b"s t x 4 + 100 ACGT", b"s q 0 4 + 100 ACGT"]]), 1),
        (_maf([[b"s t 0 4 * 100 ACGT", b"s q 0 4 + 100 ACGT"]]), 2),
        (_maf([[b"s t 0 -4 + 100 ACGT", b"s q 0 4 + 100 ACGT"]]), 1),
        (_maf([[b"s t 0 4 + 99999999999999999999999 ACGT", b"s q 0 4 + 100 ACGT"]]), 1),
        (_maf([[b"s t 0 4 + 100 ACGT"]]), 8),
        (_maf([[b"s", b"s q 0 4 + 100 ACGT"]]), 3),
        # the first error in file order wins: bad int in block 2 precedes the 1-s-line block 3
        (_maf([[b"s t 0 4 + 100 ACGT", b"s q 0 4 + 100 ACGT"], [b"s t 0 4 + 100 ACGT", b"s q zz 4 + 100 ACGT"],
               [b"s t 0 4 + 100 ACGT"]]), 1),
        (_maf([[b"s t 0 4 + 100 ACGT"], [b"s t 0 4 + 100 ACGT", b"s q zz 4 + 100 ACGT"]]), 8),
    ]
    for text, code in bad:
        for fn_o, fn_g in ((orc.maf2paf, lambda t: paf(w, ctx, t)), (orc.stat_maf, lambda t: stat(w, ctx, t))):
            with pytest.raises(orc.OracleError) as eo:
                fn_o(text)
            with pytest.raises(w.WGAError) as eg:
                fn_g(text)
            assert eo.value.code == code == eg.value.code, text
            if code != 8:
                assert str(eg.value) == eo.value.msg, text
    # after an error the context keeps working
    check_both(w, ctx, TEST_MAF)


def test_query_name(w, ctx):
    t = _maf([[b"s t 0 4 + 100 ACGT", b"s q1 0 4 + 100 ACGA", b"s q2 0 4 - 100 AC-T"],
              [b"s t 9 4 + 100 ACGT", b"s q2 0 4 + 100 ACGA", b"s q1 0 4 + 100 AC-T"]])
    for qn in (None, "q1", "q2", "t"):
        check_both(w, ctx, t, qn)
    with pytest.raises(w.WGAError) as e:
        paf(w, ctx, t, "zz")
    assert e.value.code == 7 and str(e.value) == "Query name:zz not found in MAF"


@pytest.mark.parametrize("seed,kw", [
    (1, {}), (2, dict(p_bothgap=2e-3, p_lower=0.3)), (3, dict(p_neg=0.5, p_gap=5e-3, gap_mean=6.0)),
    (4, dict(len_median=60000.0, len_sigma=1.2, len_max=400000, p_sv=2e-4, sv_min=50, sv_max=40000)),
    (5, dict(p_mis=0.2, p_gap=0.05)), (6, dict(len_median=80.0, len_min=1, len_sigma=1.5)),
])
def test_synthetic_parity(w, ctx, seed, kw):
    n = 400 if kw.get("len_median", 2000.0) < 10000 else 60
    text = bs.gen_maf(bs.maf_params(n, seed=seed, **kw))
    check_both(w, ctx, text)
    cnt = ctx.last_record_counts()
    assert cnt.tolist() == orc.maf_record_counts(text, len(cnt) + 5)


def test_random_small_blocks(w, ctx):
    rng = random.Random(11)
    alpha = "ACGTacgtN-"
    for it in range(40):
        blocks = []
        for b in range(rng.randrange(1, 30)):
            L = rng.choice([1, 2, 3, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 511, 512, 513, 1000])
            mode = rng.random()
            t = "".join(rng.choice(alpha if mode < 0.5 else "AC-") for _ in range(L))
            q = "".join((tc if rng.random() < 0.7 else rng.choice(alpha)) for tc in t)
            if rng.random() < 0.2:
                q = q[:rng.randrange(1, L + 1)]
            st = rng.choice("+-")
            blocks.append([f"s tgt{b % 3} {rng.randrange(1000)} {sum(c != '-' for c in t)} + 5000 {t}".encode(),
                           f"s qry{b % 2} {rng.randrange(1000)} {sum(c != '-' for c in q)} {st} 7000 {q}".encode()])
        check_both(w, ctx, _maf(blocks))


def test_alignment_offsets(w, ctx):
    # every (target, query) start misalignment modulo 16, runs crossing 16-byte and 512-byte boundaries
    rng = random.Random(5)
    for pad_t in range(0, 16):
        for pad_q in (0, 1, 3, 4, 7, 8, 13, 15):
            L = rng.choice([5, 40, 530, 1100])
            t = "".join(rng.choice("ACGT-") for _ in range(L))
            q = "".join((c if rng.random() < 0.8 else rng.choice("ACGT-")) for c in t)
            text = _maf([[("s " + "t" * (1 + pad_t) + f" 1 1 + 99 {t}").encode(),
                          ("s " + "q" * (1 + pad_q) + f" 1 1 + 99 {q}").encode()]])
            assert paf(w, ctx, text) == orc.maf2paf(text), (pad_t, pad_q, L)


def test_long_record_multi_item(w, ctx):
    # records longer than one work item (16384 columns), runs spanning item boundaries
    rng = random.Random(9)
    L = 16384 * 3 + 777
    t = bytearray(rng.choice(b"ACGT") for _ in range(L))
    q = bytearray(t)
    for lo, hi, kind in ((16380, 16390, "D"), (32768 - 3, 32768 + 40, "I"), (16384 * 3 - 1, 16384 * 3 + 1, "X"),
                         (100, 101, "X"), (16384, 16385, "D")):
        for i in range(lo, hi):
            if kind == "D": q[i] = ord("-")
            elif kind == "I": t[i] = ord("-")
            else: q[i] = ord("A") if t[i] != ord("A") else ord("C")
    text = _maf([[b"s t 0 1 + 9 " + bytes(t), b"s q 0 1 - 9 " + bytes(q)],
                 [b"s t 0 1 + 9 " + bytes(t[:16384]), b"s q 0 1 + 9 " + bytes(q[:16384])],
                 [b"s t 0 1 + 9 " + bytes(t[:16385]), b"s q 0 1 + 9 " + bytes(q[:16385])]])
    check_both(w, ctx, text)


def _dev_bytes(ptr, n):
    import torch

    class Raw:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(Raw(), device="cuda").cpu().numpy().tobytes() if n else b""


def test_device_resident_input(w, ctx):
    import torch
    text = bs.gen_maf(bs.maf_params(500, seed=21))
    d = torch.zeros(len(text) + 64, dtype=torch.uint8, device="cuda")
    d[:len(text)].copy_(torch.frombuffer(bytearray(text), dtype=torch.uint8))
    torch.cuda.synchronize()
    ptr, n = ctx.maf2paf_raw(ctypes.c_void_p(d.data_ptr()), len(text), 1, 1)
    assert _dev_bytes(ptr, n) == orc.maf2paf(text)
    acc = w.StatAcc(False)
    ctx.maf_stat_raw(ctypes.c_void_p(d.data_ptr()), len(text), 1, 1, acc)
    assert acc.write() == orc.stat_maf(text)


def test_multi_slab_matches_single(w, ctx):
    text = bs.gen_maf(bs.maf_params(800, seed=33, p_neg=0.3))
    old = w.api.SLAB_BYTES
    try:
        w.api.SLAB_BYTES = 200_000
        check_both(w, ctx, text)
    finally:
        w.api.SLAB_BYTES = old


def test_resident_maf_pipeline(w, ctx):
    """one upload, three commands, small slabs (unaligned slab starts -> realign copy), slab-major and command-major"""
    text = bs.gen_maf(bs.maf_params(1500, seed=35, p_neg=0.3))
    want_paf, want_stat = orc.maf2paf(text), orc.stat_maf(text)
    want_vcf = orc.call_maf(text, snp=True, svlen=0)
    for slab in (70_000, 1 << 20, 0):
        with w.ResidentMaf(text, ctx, slab) as rm:
            if slab == 70_000:
                assert ctx.L.wga_text_slabs(rm.h) > 10
            # command-major through the reference-shaped functions
            assert w.run_bytes(w.maf2paf, rm, ctx=ctx) == want_paf
            assert w.run_bytes(w.stat_maf, rm, ctx=ctx) == want_stat
            assert w.run_bytes(w.call_var_maf, rm, None, True, False, 0, ctx=ctx) == want_vcf
            # slab-major through the raw calls
            from wgatools_b200._lib import CallOpts
            opts = CallOpts(snp=1, inv=0, svlen_cutoff=0, chunk_size=1000000, query_name=None)
            acc = w.StatAcc(False)
            out_paf, out_vcf = [], [ctx.vcf_header("sample")]
            from wgatools_b200._lib import WGA_F_REUSE_INDEX
            for p, n, fl in rm.slabs():     # as bench.py's e2e leg: tokenise once per slab, reuse for stat and call
                ptr, ln = ctx.maf2paf_raw(p, n, 1, fl)
                out_paf.append(ctypes.string_at(ptr, ln) if ln else b"")
                ctx.maf_stat_raw(p, n, 1, fl | WGA_F_REUSE_INDEX, acc)
                ptr, ln = ctx.maf_call_raw(p, n, 1, fl | WGA_F_REUSE_INDEX, opts)
                out_vcf.append(ctypes.string_at(ptr, ln) if ln else b"")
            assert b"".join(out_paf) == want_paf
            assert acc.write() == want_stat
            assert b"".join(out_vcf) == want_vcf
            # the whole text in one call once everything has arrived
            p, n, fl = rm.whole()
            ptr, ln = ctx.maf2paf_raw(p, n, 1, fl)
            assert ctypes.string_at(ptr, ln) == want_paf
    # empty text
    with w.ResidentMaf(b"", ctx) as rm:
        assert w.run_bytes(w.maf2paf, rm, ctx=ctx) == b""


def test_reuse_index_across_commands(w, ctx):
    """WGA_F_REUSE_INDEX: stat / call reuse the tokenisation of the preceding command on the same device slab"""
    import torch
    from wgatools_b200._lib import CallOpts, WGA_F_REUSE_INDEX
    opts = CallOpts(snp=1, inv=1, svlen_cutoff=0, chunk_size=1000000, query_name=None)
    texts = [bs.gen_maf(bs.maf_params(700, seed=61, p_neg=0.3)), bs.gen_maf(bs.maf_params(500, seed=62)), TEST_MAF]
    cap = max(len(t) for t in texts) + 64
    d = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    for text in texts:
        d.zero_()
        d[:len(text)].copy_(torch.frombuffer(bytearray(text), dtype=torch.uint8))
        torch.cuda.synchronize()
        ptr = ctypes.c_void_p(d.data_ptr())
        p, n = ctx.maf2paf_raw(ptr, len(text), 1, 1)                      # tokenises (same pointer, new bytes)
        assert _dev_bytes(p, n) == orc.maf2paf(text)
        acc = w.StatAcc(False)
        ctx.maf_stat_raw(ptr, len(text), 1, 1 | WGA_F_REUSE_INDEX, acc)
        assert acc.write() == orc.stat_maf(text)
        p, n = ctx.maf_call_raw(ptr, len(text), 1, 1 | WGA_F_REUSE_INDEX, opts)
        assert ctx.vcf_header("sample") + _dev_bytes(p, n) == orc.call_maf(text, snp=True, inv=True, svlen=0)
        # a different query_name / a different length falls back to what each command needs
        if text is TEST_MAF:
            p, n = ctx.maf2paf_raw(ptr, len(text), 1, 1 | WGA_F_REUSE_INDEX, "query.chr8")
            assert _dev_bytes(p, n) == orc.maf2paf(text, "query.chr8")
        cut = text.rfind(b"\na ", 0, len(text) // 2) + 1
        p, n = ctx.maf2paf_raw(ptr, cut, 1, 1 | WGA_F_REUSE_INDEX)
        assert _dev_bytes(p, n) == orc.maf2paf(text[:cut])
    # errors found while tokenising are reported again by the commands that reuse the index
    bad = TEST_MAF.replace(b"\t+\t", b"\t*\t", 1)
    d.zero_()
    d[:len(bad)].copy_(torch.frombuffer(bytearray(bad), dtype=torch.uint8))
    torch.cuda.synchronize()
    ptr = ctypes.c_void_p(d.data_ptr())
    msgs = []
    for flags, fn in ((1, lambda f: ctx.maf2paf_raw(ptr, len(bad), 1, f)),
                      (1 | WGA_F_REUSE_INDEX, lambda f: ctx.maf_stat_raw(ptr, len(bad), 1, f, w.StatAcc(False))),
                      (1 | WGA_F_REUSE_INDEX, lambda f: ctx.maf_call_raw(ptr, len(bad), 1, f, opts))):
        with pytest.raises(w.WGAError) as e:
            fn(flags)
        msgs.append((e.value.code, str(e.value)))
    assert msgs[0] == msgs[1] == msgs[2] and msgs[0][0] == 2


def test_share_walk_one_pass(w, ctx):
    """WGA_F_SHARE_WALK: maf2paf runs the dense walk once in the caller classification and keeps counters + run list;
    stat then needs no walk and call starts at its chunk grid.  Outputs stay byte-identical to the oracle, with and
    without both-gap columns (where the EXT and the caller run lists differ), in either order of stat / call."""
    import torch
    from wgatools_b200._lib import CallOpts, WGA_F_REUSE_INDEX, WGA_F_SHARE_WALK
    texts = [
        (bs.gen_maf(bs.maf_params(600, seed=71, p_neg=0.3, p_bothgap=2e-3, p_mis=5e-3, p_gap=2e-3)), None),
        (bs.gen_maf(bs.maf_params(40, seed=72, len_median=60000.0, len_sigma=0.5, len_max=200000, p_bothgap=1e-3, p_neg=0.5)), None),
        (bs.gen_maf(bs.maf_params(500, seed=73, p_neg=0.1)), None),            # no both-gap columns: lists coincide
        (TEST_MAF, None),
        (TEST_MAF, "query.chr8"),
        # hand-made: = W = / X W X / I W I / leading and trailing W, adjacent blocks
        (b"##maf version=1\na score=0\ns t.c 10 8 + 100 AC-GT--ACG-\ns q.c 5 8 - 90 AC-CT--A-GA\n\n"
         b"a\ns t.c 30 4 + 100 --AC-GT-\ns q.c 40 5 + 90 --ACTG--\n\na\ns t.c 50 3 + 100 A-C-G\ns q.c 60 3 + 90 A-T-G\n", None),
    ]
    cap = max(len(t) for t, _ in texts) + 64
    d = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    for text, qn in texts:
        opts = CallOpts(snp=1, inv=1, svlen_cutoff=0, chunk_size=20000, query_name=qn.encode() if qn else None)
        d.zero_()
        d[:len(text)].copy_(torch.frombuffer(bytearray(text), dtype=torch.uint8))
        torch.cuda.synchronize()
        ptr = ctypes.c_void_p(d.data_ptr())
        want_vcf = orc.call_maf(text, snp=True, inv=True, svlen=0, query_name=qn, chunk_size=20000)
        for order in ("stat,call", "call,stat"):
            p, n = ctx.maf2paf_raw(ptr, len(text), 1, 1 | WGA_F_SHARE_WALK, qn)
            assert _dev_bytes(p, n) == orc.maf2paf(text, qn)
            assert ctx.last_walk()[2] == 1
            cnt = ctx.last_record_counts()
            for cmd in order.split(","):
                if cmd == "stat":
                    for each in (False, True):
                        acc = w.StatAcc(each)
                        ctx.maf_stat_raw(ptr, len(text), 1, 1 | WGA_F_REUSE_INDEX, acc, qn)
                        assert acc.write() == orc.stat_maf(text, each, qn)
                        assert ctx.last_walk()[2] == 0, "stat after a shared walk must not walk again"
                        assert (ctx.last_record_counts() == cnt).all()
                else:
                    p, n = ctx.maf_call_raw(ptr, len(text), 1, 1 | WGA_F_REUSE_INDEX, opts)
                    assert ctx.vcf_header("sample") + _dev_bytes(p, n) == want_vcf
                    assert ctx.last_walk()[2] == 0, "call after a shared walk must not walk again"
        # another query name than the one the walk was shared for: every command falls back to its own pass
        if qn is None and text is TEST_MAF:
            ctx.maf2paf_raw(ptr, len(text), 1, 1 | WGA_F_SHARE_WALK)
            acc = w.StatAcc(False)
            ctx.maf_stat_raw(ptr, len(text), 1, 1 | WGA_F_REUSE_INDEX, acc, "query.chr8")
            assert acc.write() == orc.stat_maf(text, False, "query.chr8") and ctx.last_walk()[2] == 1
    # a block `call` sees differently (third row shorter than the target row -> the reference panics in
    # create_chunk_record): maf2paf / stat are fine, call must not pick up the shared records
    rag = (b"##maf version=1\na\ns t.c 0 8 + 100 ACGTACGT\ns q.c 0 8 + 90 ACGTACGA\ns r.c 0 4 + 80 ACGT\n\n"
           b"a\ns t.c 20 4 + 100 ACGT\ns q.c 30 4 + 90 AGGT\n")
    d.zero_()
    d[:len(rag)].copy_(torch.frombuffer(bytearray(rag), dtype=torch.uint8))
    torch.cuda.synchronize()
    ptr = ctypes.c_void_p(d.data_ptr())
    opts = CallOpts(snp=1, inv=0, svlen_cutoff=0, chunk_size=1000000, query_name=None)
    p, n = ctx.maf2paf_raw(ptr, len(rag), 1, 1 | WGA_F_SHARE_WALK)
    assert _dev_bytes(p, n) == orc.maf2paf(rag)
    acc = w.StatAcc(False)
    ctx.maf_stat_raw(ptr, len(rag), 1, 1 | WGA_F_REUSE_INDEX, acc)
    assert acc.write() == orc.stat_maf(rag) and ctx.last_walk()[2] == 0
    with pytest.raises(orc.OracleError) as oe:
        orc.call_maf(rag, snp=True, svlen=0)
    with pytest.raises(w.WGAError) as e:
        ctx.maf_call_raw(ptr, len(rag), 1, 1 | WGA_F_REUSE_INDEX, opts)
    assert e.value.code == 8 and oe.value.code == 8


def test_stat_export_import_roundtrip(w, ctx):
    text = bs.gen_maf(bs.maf_params(600, seed=44, p_neg=0.4))
    L = ctx.L
    cut = L.wga_maf_safe_cut(text, len(text) // 2)
    assert 0 < cut < len(text)
    a, b = w.StatAcc(False), w.StatAcc(False)
    ka, pa = w.api._host_ptr(text)
    ctx.maf_stat_raw(ctypes.c_void_p(pa), cut, 0, 1, a)
    ctx.maf_stat_raw(ctypes.c_void_p(pa + cut), len(text) - cut, 0, 0, b)
    m = w.StatAcc(False)
    m.import_(a.export())
    m.import_(b.export())
    assert m.write() == orc.stat_maf(text)


# ------------------------------------------------------------------------------------------- call
def vcf(w, ctx, text, **kw):
    return w.run_bytes(w.call_var_maf, w.MAFReader(text), None, kw.get("snp", False), kw.get("inv", False),
                       kw.get("svlen", 50), False, kw.get("sample"), kw.get("query_name"), None,
                       kw.get("chunk_size"), ctx=ctx)


def check_call(w, ctx, text, **kw):
    want = orc.call_maf(text, snp=kw.get("snp", False), inv=kw.get("inv", False), svlen=kw.get("svlen", 50),
                        sample=kw.get("sample") or "sample", query_name=kw.get("query_name"),
                        chunk_size=kw.get("chunk_size") or 1000000)
    got = vcf(w, ctx, text, **kw)
    if got != want:
        gl, wl = got.splitlines(), want.splitlines()
        for i, (a, b) in enumerate(zip(gl, wl)):
            assert a == b, (i, a, b, kw)
        assert len(gl) == len(wl), (len(gl), len(wl), kw)
    assert got == want


def test_call_golden(w, ctx):
    check_call(w, ctx, TEST_MAF, snp=True, svlen=0)
    check_call(w, ctx, TEST_MAF)
    check_call(w, ctx, TEST_MAF, snp=True, svlen=0, inv=True, sample="hg002", chunk_size=100)
    got = vcf(w, ctx, TEST_MAF, snp=True, svlen=0).decode().splitlines()
    readme = open(os.path.join(G, "ref_readme_call_s_l0.vcf")).read().splitlines()
    for a, b in zip(got, readme):
        if a.startswith("#") or "SVTYPE" in a:
            assert a == b
        else:
            assert a.split("\t")[:8] == b.split("\t")[:8]


def test_call_quirks(w, ctx):
    t = b"ACGTACGTAC" + b"GGGGGGGG" + b"ACGT"
    q = b"ACGTACGTAC" + b"--------" + b"ACGT"
    for strand in (b"+", b"-"):
        text = _maf([[b"s t 0 22 + 100 " + t, b"s q 0 14 " + strand + b" 100 " + q]])
        for cs in (1000, 10, 11, 14, 3, 1):
            for K in (0, 3, 8, 9):
                check_call(w, ctx, text, svlen=K, chunk_size=cs, inv=True, snp=True)
    one = _maf([[b"s t 0 4 + 100 ACGT"]])
    assert vcf(w, ctx, one) == orc.call_maf(one)
    three = _maf([[b"s t 0 4 + 100 ACGT", b"s q1 0 4 + 100 AC-T", b"s q2 0 4 - 100 A-GA"],
                  [b"s t 0 4 + 100 ACGT", b"s q1 0 4 + 100 ACGT"]])
    for qn in (None, "q2", "zz", "t"):
        check_call(w, ctx, three, svlen=0, snp=True, inv=True, query_name=qn)
    # both-gap columns (W), adjacent I/D, leading / trailing gaps, lower case upper-cased on output
    text = _maf([[b"s t 0 9 + 100 --AC-GT--AaC---GT", b"s q 0 9 + 100 GGAC-G-TTAcCTTT--"]])
    check_call(w, ctx, text, svlen=0, snp=True)
    check_call(w, ctx, text, svlen=1, snp=True, inv=True)
    # a base outside ACGTN in REF/ALT is an error (noodles)
    bad = _maf([[b"s t 0 4 + 100 ACRT", b"s q 0 4 + 100 ACGT"]])
    with pytest.raises(orc.OracleError) as eo:
        orc.call_maf(bad, snp=True)
    with pytest.raises(w.WGAError) as eg:
        vcf(w, ctx, bad, snp=True)
    assert eo.value.code == eg.value.code == 10
    assert vcf(w, ctx, bad) == orc.call_maf(bad)  # no SNP rows requested: the base is never printed
    with pytest.raises(w.WGAError) as e:
        vcf(w, ctx, TEST_MAF, chunk_size=0)
    assert e.value.code == 101


def test_call_query_regex_and_contigs(w, ctx):
    """--query-regex (caller.rs:91-104, maf.rs:267-271) through wga_maf_sline_names + wga_call_opts.query_match, and the
    ##contig lines of a <maf>.index (caller.rs:339-357) through wga_vcf_header_contigs: the host mirror against the
    oracle's predicate restatement, on hand-made multi-row blocks and on a generated 3-row MAF with many names."""
    import random
    hand = _maf([[b"s t.chr1 0 4 + 100 ACGT", b"s hg002#1 0 4 + 100 ACGA", b"s hg002#2 0 4 + 100 AC-T"],
                 [b"s t.chr1 10 4 + 100 ACGT", b"s other 0 4 + 100 AGGT"],
                 [b"s t.chr2 0 4 + 100 ACGT", b"s zz 0 4 - 100 ACGT", b"s hg002#2 9 4 + 100 TCGT"],
                 [b"s t.chr2 0 4 + 100 ACGT"]])
    rng = random.Random(5)
    blocks = []
    for b in range(400):
        t = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(20, 200)))
        rows = [b"s ref.chr%d %d %d + 100000 " % (b % 7, b * 10, len(t)) + t]
        for k in range(rng.randrange(1, 4)):
            q = bytearray(t)
            for i in range(len(q)):
                if rng.random() < 0.05:
                    q[i] = rng.choice(b"ACGT-")
            n = sum(1 for c in q if c != 45)
            rows.append(b"s sp%d#%d.ctg%d %d %d %s 100000 " % (rng.randrange(4), rng.randrange(2), rng.randrange(30), b, n,
                                                              rng.choice([b"+", b"-"])) + bytes(q))
        blocks.append(rows)
    big = _maf(blocks)
    idx = {"ref.chr10": {"size": 30, "isref": True}, "ref.chr2": {"size": 20, "isref": True}, "sp1#0.ctg1": {"size": 5, "isref": False},
           "ref.chr1": {"size": 100000, "isref": True}}
    for text, pats in ((hand, [rb"^hg002#.*", rb"#2", rb"chr", rb"nomatch", rb"^(zz|other)$", rb"."]),
                       (big, [rb"^sp2#", rb"#1\.ctg1\d$", rb"^sp[03]#0", rb"ctg7$", rb"^ref", rb"xyz"])):
        for pat in pats:
            for mi in (None, idx):
                got = w.run_bytes(w.call_var_maf, w.MAFReader(text), mi, True, True, 0, query_regex=pat, ctx=ctx)
                assert got == orc.call_maf_ex(text, snp=True, inv=True, svlen=0, query_regex=pat, mafindex=mi), pat
        # --query-name wins over the regex (caller.rs:76)
        got = w.run_bytes(w.call_var_maf, w.MAFReader(text), None, True, False, 0, query_name="zz", query_regex=rb"hg002", ctx=ctx)
        assert got == orc.call_maf(text, snp=True, svlen=0, query_name="zz")
    # the names the host evaluates the regex on: distinct, first-occurrence order
    from wgatools_b200._lib import WGA_F_FIRST_SLAB, WGA_MEM_HOST
    hbuf = ctypes.c_char_p(hand)
    names = ctx.maf_sline_names(ctypes.cast(hbuf, ctypes.c_void_p), len(hand), WGA_MEM_HOST, WGA_F_FIRST_SLAB)
    assert names == [b"t.chr1", b"hg002#1", b"hg002#2", b"other", b"t.chr2", b"zz"]
    with pytest.raises(w.WGAError) as e:
        ctx.vcf_header_contigs("s", [("bad,name", 3)])
    assert e.value.code == 101


def test_call_random_chunks(w, ctx):
    rng = random.Random(17)
    for it in range(60):
        blocks = []
        for b in range(rng.randrange(1, 12)):
            L = rng.choice([1, 2, 5, 17, 40, 64, 130, 600])
            kind = rng.random()
            t, q = [], []
            while len(t) < L:
                r = rng.random()
                n = rng.randrange(1, 9)
                if r < 0.5:
                    for _ in range(n):
                        c = rng.choice("ACGT"); t.append(c); q.append(c)
                elif r < 0.65:
                    for _ in range(n):
                        c = rng.choice("ACGT"); t.append(c); q.append(rng.choice([x for x in "ACGT" if x != c]))
                elif r < 0.8:
                    t += ["-"] * n; q += [rng.choice("ACGT") for _ in range(n)]
                elif r < 0.95:
                    q += ["-"] * n; t += [rng.choice("ACGT") for _ in range(n)]
                else:
                    t += ["-"] * n; q += ["-"] * n
            t, q = "".join(t[:L]), "".join(q[:L])
            st = rng.choice("+-")
            nt, nq = sum(c != "-" for c in t), sum(c != "-" for c in q)
            blocks.append([f"s chr{b % 2} {rng.randrange(100, 900)} {nt} + 5000 {t}".encode(),
                           f"s q.{b % 3} {rng.randrange(100, 900)} {nq} {st} 7000 {q}".encode()])
        text = _maf(blocks)
        for kw in (dict(snp=True, svlen=0, inv=True, chunk_size=rng.choice([1, 2, 3, 7, 16, 50, 1000])),
                   dict(snp=False, svlen=rng.choice([1, 2, 4]), inv=True, chunk_size=rng.choice([5, 9, 33])),
                   dict(snp=True, svlen=3, inv=False, chunk_size=rng.choice([4, 10, 100]))):
            check_call(w, ctx, text, **kw)


@pytest.mark.parametrize("seed,kw,ckw", [
    (1, {}, dict(snp=True, svlen=0)),
    (2, dict(p_bothgap=2e-3, p_neg=0.3), dict(snp=True, svlen=0, inv=True)),
    (3, dict(len_median=60000.0, len_sigma=1.2, len_max=400000, p_sv=3e-4, sv_min=50, sv_max=30000, p_neg=0.3),
     dict(inv=True, chunk_size=20000)),
    (4, dict(len_median=60000.0, len_sigma=1.2, len_max=400000, p_sv=3e-4, sv_min=50, sv_max=30000, p_neg=0.3),
     dict(snp=True, svlen=2, inv=True, chunk_size=5000)),
    (5, dict(p_mis=0.1, p_gap=0.05, p_neg=0.5), dict(snp=True, svlen=1, inv=True, chunk_size=300)),
])
def test_call_synthetic(w, ctx, seed, kw, ckw):
    n = 300 if kw.get("len_median", 2000.0) < 10000 else 40
    text = bs.gen_maf(bs.maf_params(n, seed=100 + seed, **kw))
    check_call(w, ctx, text, **ckw)


def test_dense_whitespace_falls_back_to_exact_index(w, ctx):
    # > 128 whitespace bytes inside one 4 KiB tile overflows the single-pass index -> exact two-pass path
    blk = [b"s t 0 4 + 100 ACGT", b"s q 0 4 + 100 AC-A"]
    text = b"##maf\n" + b"\n" * 6000 + _maf([blk], header=b"") + b" \t " * 3000 + b"\n" + _maf([blk, blk], header=b"")
    check_both(w, ctx, text)
    check_call(w, ctx, text, snp=True, svlen=0)
    big = bs.gen_maf(bs.maf_params(300, seed=5)) + b"\n" * 9000 + b"a\n" + b"\n".join(blk) + b"\n"
    check_both(w, ctx, big)


def test_cli_matches_oracle(w, ctx, tmp_path):
    import subprocess
    import sys
    maf = tmp_path / "in.maf"
    text = bs.gen_maf(bs.maf_params(120, seed=9, p_neg=0.2))
    maf.write_bytes(text)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    def run(*args):
        return subprocess.run([sys.executable, "-m", "wgatools_b200", *args], cwd=root, capture_output=True)
    r = run("m2p", str(maf))
    assert r.returncode == 0 and r.stdout == orc.maf2paf(text)
    out = tmp_path / "o.tsv"
    assert run("-o", str(out), "st", str(maf)).returncode == 0 and out.read_bytes() == orc.stat_maf(text)
    assert run("-o", str(out), "st", str(maf)).returncode == 1  # refuses to overwrite without -r
    r = run("c", str(maf), "-s", "-l", "0", "-n", "hg1")
    assert r.returncode == 0 and r.stdout == orc.call_maf(text, snp=True, svlen=0, sample="hg1")
    r = run("m2p", str(maf), "-q", "nope")
    assert r.returncode == 1 and b"Query name:nope not found in MAF" in r.stderr


# ------------------------------------------------------------------------------------- stat -f paf
TEST_PAF = open(os.path.join(G, "ref_testdotplot.paf"), "rb").read()


def statp(w, ctx, text, each=False):
    return w.run_bytes(w.stat_paf, w.PAFReader(text), each, ctx=ctx)


def check_paf(w, ctx, text):
    assert statp(w, ctx, text) == orc.stat_paf(text)
    assert statp(w, ctx, text, True) == orc.stat_paf(text, True)


def test_paf_stat_golden(w, ctx):
    out = statp(w, ctx, TEST_PAF).decode().splitlines()
    assert out[1].split("\t") == "B 300 0 A 300 0 250 50 0.84 0.84 210 0 2 2 30 30 1 50.0 1 10 1 10".split()
    check_paf(w, ctx, TEST_PAF)


def _pafline(cg, strand="+", q="q1", t="t1", extra=("NM:i:3",), ints=(1000, 10, 200, 2000, 50, 400, 180, 200, 60)):
    f = [q, str(ints[0]), str(ints[1]), str(ints[2]), strand, t, str(ints[3]), str(ints[4]), str(ints[5]), str(ints[6]),
         str(ints[7]), str(ints[8])] + list(extra) + ([f"cg:Z:{cg}"] if cg is not None else [])
    return "\t".join(f).encode()


def test_paf_stat_quirks_and_errors(w, ctx):
    ok = [
        b"#comment\n" + _pafline("10M2I3D5X4=") + b"\n\n" + _pafline("7=", "-", "q2") + b"\r\n" + _pafline("3=1X", "-"),
        _pafline("100M") + b"\n" + _pafline("5=", extra=("tp:A:P", "cg:Z:9=")) + b"\n",  # first cg:Z: tag wins
        "\t".join(["q", "0x10", "1", "2", "+", "t", "0x20", "3", "4", "1", "1", "0", "cg:Z:1="]).encode(),  # csv hex ints
        _pafline("18446744073709551615=1X"),
        # csv Terminator::CRLF: a lone '\r' ends a record too; a comment runs to the next '\n' whatever '\r' it holds
        _pafline("10M") + b"\r" + _pafline("7=", "-", "q2") + b"\r\r\n\r" + _pafline("3=1X", "-") + b"\r",
        b"#c\rstill comment\r" + _pafline("9=") + b"\n" + _pafline("4=") + b"\r#x\r" + _pafline("8=") + b"\n\r" + _pafline("5=", "-"),
    ]
    for t in ok:
        check_paf(w, ctx, t)
    bad = [(_pafline("10S5M"), 5), (_pafline("5H"), 5), (_pafline("10MM"), 5), (_pafline("M"), 1), (_pafline("12"), 5),
           (_pafline("99999999999999999999999M"), 1), (_pafline(None), 6), (_pafline(""), 8),
           (b"q\t1\t2\n", 9), (_pafline("5=", strand="*"), 9), (_pafline("5=", ints=(1000, "x", 200, 2000, 50, 400, 180, 200, 60)), 9),
           (_pafline("5=") + b"\n" + _pafline("3=2Q1=") + b"\n" + _pafline("M"), 5),   # first bad record wins
           (_pafline("5=") + b"\n" + _pafline("3=Q") + b"\n" + b"q\t1\n", 1)]           # "3=Q": op ok? no: 'Q' has no digits
    for text, code in bad:
        with pytest.raises(orc.OracleError) as eo:
            orc.stat_paf(text)
        with pytest.raises(w.WGAError) as eg:
            statp(w, ctx, text)
        assert eo.value.code == eg.value.code, (text, eo.value.code, eg.value.code, eo.value.msg)
        if eo.value.code in (1, 5, 6):
            assert str(eg.value) == eo.value.msg, text
    check_paf(w, ctx, TEST_PAF)


@pytest.mark.parametrize("seed,kw", [(1, {}), (2, dict(m_only=1)), (3, dict(ops_median=4000.0, ops_sigma=1.0)),
                                     (4, dict(ops_median=3.0, ops_sigma=0.8, n_genomes=3, n_chr=2))])
def test_paf_stat_synthetic(w, ctx, seed, kw):
    n = 3000 if kw.get("ops_median", 150) < 1000 else 300
    text = bs.gen_paf(bs.paf_params(n, seed=seed, **kw))
    check_paf(w, ctx, text)
    cnt = ctx.last_record_counts()
    assert len(cnt) == n


def test_paf_stat_long_numbers_all_alignments(w, ctx):
    """lengths of 1..20 digits crossing the 16-byte pieces of the warp kernel at every alignment; 10+ digit lengths,
    malformed tokens and overflowing numbers take the exact path (per-record counters compared too)"""
    rng = random.Random(5)
    lines, cgs = [], []
    for i in range(600):
        ops = []
        for _ in range(rng.randrange(1, 40)):
            nd = rng.choice([1, 1, 2, 3, 5, 8, 9, 10, 12, 15, 16, 17, 19])
            v = rng.randrange(10 ** (nd - 1), 10 ** nd) if nd < 19 else rng.randrange(10 ** 18, 2 ** 62)
            if rng.random() < 0.3:
                v = b"%0*d" % (rng.randrange(nd, nd + 8), v)   # leading zeros
            else:
                v = b"%d" % v
            ops.append(v + bytes([rng.choice(b"M=XID")]))
        cg = b"".join(ops)
        cgs.append(cg)
        pad = b"q" * rng.randrange(1, 33)                       # shifts the CIGAR's alignment
        lines.append(b"\t".join([pad, b"100", b"0", b"50", rng.choice([b"+", b"-"]), b"t%d" % (i % 3), b"100", b"0", b"50", b"1", b"2",
                                 b"60", b"cg:Z:" + cg]) + b"\n")
    text = b"".join(lines)
    acc = w.StatAcc(True)
    buf = ctypes.c_char_p(text)
    ctx.paf_stat_raw(ctypes.cast(buf, ctypes.c_void_p), len(text), 0, 1, acc)
    cnt = ctx.last_record_counts()
    for i, cg in enumerate(cgs):
        want = orc.paf_cigar_counts(cg, lines[i].split(b"\t")[4] == b"-")
        assert [int(x) & (2 ** 64 - 1) for x in cnt[i]] == [int(x) & (2 ** 64 - 1) for x in want], (i, cg)
    # one malformed token anywhere -> the reference's error, from the lowest record
    for badcg in (b"12MM3=", b"5=7", b"3=Q4M", b"5=99999999999999999999M", b"4=5S"):
        t2 = b"".join(lines[:7]) + lines[7].replace(b"cg:Z:" + cgs[7], b"cg:Z:" + cgs[7] + badcg) + b"".join(lines[8:20])
        with pytest.raises(orc.OracleError) as eo:
            orc.stat_paf(t2)
        with pytest.raises(w.WGAError) as eg:
            w.run_bytes(w.stat_paf, w.PAFReader(t2), ctx=ctx)
        assert eo.value.code == eg.value.code and str(eg.value) == eo.value.msg, badcg


# ------------------------------------------------------------------------- larger inputs (seconds of oracle time)
def test_large_config2_exact(w, ctx):
    """100 k blocks (0.67 GB) of the bench workload: all three commands byte-identical to the oracle."""
    text = bs.gen_maf(bs.maf_params(100_000, seed=20260926))
    assert paf(w, ctx, text) == orc.maf2paf(text)
    assert stat(w, ctx, text) == orc.stat_maf(text)
    assert vcf(w, ctx, text, snp=True, svlen=0) == orc.call_maf(text, snp=True, svlen=0)
    assert vcf(w, ctx, text) == orc.call_maf(text)


def test_large_config3_long_records_exact(w, ctx):
    """human-vs-human shape (SURVEY 8(d) cfg 3): blocks up to 5 M columns so that the 1 M-column chunk grid, SVs and
    multi-item records are all exercised; default call and call -s -l0 --inv."""
    text = bs.gen_maf(bs.maf_params(400, seed=31, config=3, p_neg=0.2))
    assert len(text) > 50_000_000
    assert paf(w, ctx, text) == orc.maf2paf(text)
    assert stat(w, ctx, text) == orc.stat_maf(text)
    assert vcf(w, ctx, text, inv=True) == orc.call_maf(text, inv=True)
    assert vcf(w, ctx, text, snp=True, svlen=0, inv=True) == orc.call_maf(text, snp=True, svlen=0, inv=True)


def test_full_size_properties(w, ctx):
    """size-independent invariants on a 1.3 GB slab (too big to diff comfortably): one PAF row per record, NM and
    block length consistent with the per-record counters, stat totals = column sums, one SNP row per mismatch."""
    text = bs.gen_maf(bs.maf_params(200_000, seed=5))
    old_slab = w.api.SLAB_BYTES
    w.api.SLAB_BYTES = 1 << 40  # one slab, so that last_record_counts covers the whole file
    try:
        out = paf(w, ctx, text)
    finally:
        w.api.SLAB_BYTES = old_slab
    cnt = ctx.last_record_counts().astype(np.int64)
    rows = out.split(b"\n")[:-1]
    assert len(rows) == len(cnt) == 200_000
    f = np.array([[int(r.split(b"\t")[9]), int(r.split(b"\t")[10]), int(r.split(b"\t")[12][5:])] for r in rows[::97]])
    c = cnt[::97]
    assert (f[:, 0] == c[:, 0]).all() and (f[:, 1] == c[:, [0, 1, 3, 5, 7, 9]].sum(1)).all() and (f[:, 2] == f[:, 1] - f[:, 0]).all()
    st = stat(w, ctx, text).decode().splitlines()[1:]
    tot = np.array([[int(x) for i, x in enumerate(l.split("\t")) if i in (10, 11, 12, 13, 14, 15, 18, 19, 20, 21)] for l in st]).sum(0)
    want = cnt[:, [0, 1, 2, 4, 3, 5, 6, 7, 8, 9]].sum(0)
    assert (tot == want).all()
    v = vcf(w, ctx, text, snp=True, svlen=10**9)
    assert v.count(b"\n") - 8 == int(cnt[:, 1].sum())  # only SNP rows when no indel passes the cutoff


# ----------------------------------------------------------------------------------------- paf2maf
def _revcomp(s):
    return s.translate(bytes.maketrans(b"ACGTacgtNn", b"TGCAtgcaNn"))[::-1]


def _wrap(name, seq, width):
    return b">" + name + b" some description\n" + b"\n".join(seq[i:i + width] for i in range(0, len(seq), width)) + b"\n"


def make_paf2maf_case(rng, n_rec=40, n_tchr=3, n_qchr=4, width_t=60, width_q=70, soft=True):
    """target genome + query genome assembled from mutated (and for '-' records reverse-complemented) target segments,
    with a PAF whose cg:Z: strings are consistent with both FASTAs"""
    alpha = b"ACGT"
    tch = [bytes(rng.choice(alpha) for _ in range(rng.randrange(3000, 9000))) for _ in range(n_tchr)]
    if soft:
        tch = [c[:500] + c[500:900].lower() + c[900:] for c in tch]
    qparts = [[] for _ in range(n_qchr)]
    recs = []
    for r in range(n_rec):
        ti = rng.randrange(n_tchr)
        L = rng.randrange(1, 1500)
        ts = rng.randrange(0, len(tch[ti]) - L)
        tseg = tch[ti][ts:ts + L]
        q = bytearray()
        ops = []
        i = 0
        while i < L:
            u = rng.random()
            n = min(L - i, rng.randrange(1, 60))
            if u < 0.6:
                ops.append((b"=", n)); q += tseg[i:i + n]; i += n
            elif u < 0.75:
                n = min(n, 3)
                for k in range(n):
                    c = tseg[i + k:i + k + 1].upper()
                    q += bytes([rng.choice([x for x in alpha if bytes([x]) != c])])
                ops.append((b"X", n)); i += n
            elif u < 0.88:
                ops.append((b"D", n)); i += n
            else:
                ops.append((b"I", n)); q += bytes(rng.choice(alpha) for _ in range(n))
        if rng.random() < 0.3:
            ops = [(b"M" if o in b"=X" else o, n) for o, n in ops]  # M-only spelling
        merged = []
        for o, n in ops:
            if merged and merged[-1][0] == o:
                merged[-1][1] += n
            else:
                merged.append([o, n])
        neg = rng.random() < 0.4
        qi = rng.randrange(n_qchr)
        spacer = bytes(rng.choice(alpha) for _ in range(rng.randrange(0, 50)))
        qparts[qi].append(spacer)
        qs = sum(len(p) for p in qparts[qi])
        qparts[qi].append(_revcomp(bytes(q)) if neg else bytes(q))
        recs.append((qi, qs, qs + len(q), neg, ti, ts, ts + L, b"".join(b"%d%s" % (n, o) for o, n in merged)))
    qch = [b"".join(p) + b"ACGT" for p in qparts]
    t_fa = b"".join(_wrap(b"tgt.chr%d" % i, c, width_t) for i, c in enumerate(tch))
    q_fa = b"".join(_wrap(b"qry#%d" % i, c, width_q) for i, c in enumerate(qch))
    paf = b"".join(b"qry#%d\t%d\t%d\t%d\t%s\ttgt.chr%d\t%d\t%d\t%d\t%d\t%d\t%d\ttp:A:P\tcg:Z:%s\n" %
                   (qi, len(qch[qi]), qs, qe, b"-" if neg else b"+", ti, len(tch[ti]), ts, te, 1, 2, rng.randrange(0, 61), cg)
                   for (qi, qs, qe, neg, ti, ts, te, cg) in recs)
    return paf, t_fa, orc.make_fai(t_fa), q_fa, orc.make_fai(q_fa)


def p2m(w, ctx, paf, t_fa, t_fai, q_fa, q_fai):
    out = io.BytesIO()
    w.paf2maf(w.PAFReader(paf), out, "t.fa", "q.fa", ctx=ctx, target=w.Fasta(ctx, t_fa, t_fai), query=w.Fasta(ctx, q_fa, q_fai))
    return out.getvalue()


def test_paf2maf_parity(w, ctx):
    rng = random.Random(2026)
    for it in range(6):
        case = make_paf2maf_case(rng, n_rec=60, width_t=rng.choice([50, 60, 80]), width_q=rng.choice([60, 61, 100]), soft=it % 2 == 0)
        want = orc.paf2maf(*case)
        got = p2m(w, ctx, *case)
        assert got == want
        # the result is a MAF whose maf2paf round-trips the alignment counters of the PAF CIGARs
        back = orc.maf2paf(got)
        assert back.count(b"\n") == case[0].count(b"\n")


def test_paf2maf_edges_and_errors(w, ctx):
    t_fa = b">chrT desc\nACGTACGTAC\nGGGGCCCCAA\nTTTT\n>other\nAAAA\n"
    q_fa = b">chrQ\nACGTACGTACGGCCCCAATT\nTTAA\n>bad\nACGTRYAC\n"
    ti, qi = orc.make_fai(t_fa), orc.make_fai(q_fa)
    def line(cg, strand=b"+", q=b"chrQ", t=b"chrT", qs=1, qe=13, ts=2, te=14, tag=True):
        return b"\t".join([q, b"24", b"%d" % qs, b"%d" % qe, strand, t, b"24", b"%d" % ts, b"%d" % te, b"9", b"14", b"60"] +
                          ([b"cg:Z:" + cg] if tag else [])) + b"\n"
    ok = [line(b"4=1X2I3=2D2="), line(b"4=1X2I3=2D2=", b"-"), line(b"5M"),           # CIGAR shorter than the spans
          line(b"3M2I20M"), line(b"12M3D"), line(b"2I12M", ts=20, te=30),            # spans clipped at the sequence end
          b"#c\n" + line(b"12=") + b"\n" + line(b"6=6X", b"-"), line(b"1=", qs=0, qe=1, ts=0, te=1), line(b"3=20D1="),
          line(b"3=20D1=", b"-"), line(b"2I3D2I8=")]
    for paf in ok:
        assert p2m(w, ctx, paf, t_fa, ti, q_fa, qi) == orc.paf2maf(paf, t_fa, ti, q_fa, qi), paf
    bad = [(line(b"10S2M"), 5), (line(b"10MM"), 5), (line(b"M"), 1), (line(b""), 8), (line(b"5=", tag=False), 6),
           (line(b"5=", t=b"nope"), 11), (line(b"5=", q=b"nope"), 11), (line(b"8=", b"-", q=b"bad", qs=0, qe=8), 10),
           (line(b"20M2I"), 8), (line(b"20M2D"), 8), (b"a\tb\n", 9),
           (line(b"12=") + line(b"3=2Q") + line(b"M"), 5)]
    for paf, code in bad:
        with pytest.raises(orc.OracleError) as eo:
            orc.paf2maf(paf, t_fa, ti, q_fa, qi)
        with pytest.raises(w.WGAError) as eg:
            p2m(w, ctx, paf, t_fa, ti, q_fa, qi)
        assert eo.value.code == code == eg.value.code, (paf, eo.value.code, eg.value.code, eo.value.msg)
        if code in (1, 5, 6, 10):
            assert str(eg.value) == eo.value.msg, paf


# ---------------------------------------------------------------- call -f paf (row a15)
def cpaf(w, ctx, paf, t_fa, t_fai, q_fa, q_fai, snp=False, svlen=50, sample=None):
    out = io.BytesIO()
    w.call_var_paf(w.PAFReader(paf), "t.fa", "q.fa", out, snp, svlen, False, sample, ctx=ctx,
                   target=w.Fasta(ctx, t_fa, t_fai), query=w.Fasta(ctx, q_fa, q_fai))
    return out.getvalue()


def test_call_paf_parity(w, ctx):
    rng = random.Random(77)
    for it in range(6):
        case = make_paf2maf_case(rng, n_rec=80, width_t=rng.choice([50, 60, 80]), width_q=rng.choice([60, 61, 100]), soft=it % 2 == 0)
        for snp, svlen in ((True, 0), (False, 0), (True, 5), (False, 50)):
            want = orc.call_paf(*case, snp=snp, svlen=svlen, sample="s1")
            got = cpaf(w, ctx, *case, snp=snp, svlen=svlen, sample="s1")
            assert got == want, (it, snp, svlen)
            if snp and svlen == 0:
                assert want.count(b"\n") > 200


def test_call_paf_edges_and_errors(w, ctx):
    t_fa = b">chrT\nACGTACGTACGGGGCCCCAATTTT\n>rT\nRCGTACGTAC\n"
    q_fa = b">chrQ\nACGTACGTACGGCCCCAATTTTAA\n>yQ\nACGTAYGTACGG\n"
    ti, qi = orc.make_fai(t_fa), orc.make_fai(q_fa)

    def line(cg, strand=b"+", q=b"chrQ", t=b"chrT", qs=1, qe=13, ts=2, te=14, tag=True):
        return b"\t".join([q, b"24", b"%d" % qs, b"%d" % qe, strand, t, b"24", b"%d" % ts, b"%d" % te, b"9", b"14", b"60"] +
                          ([b"cg:Z:" + cg] if tag else [])) + b"\n"
    ok = [line(b"4=1X2I3=2D2="), line(b"4=1X2I3=2D2=", b"-"), line(b"2I4=2D3I2="), line(b"4=1X3S2I3=") + line(b"4=1X"),
          line(b"4=1X2II3=") + line(b"2=2X", b"-"), line(b"4=X2=") + line(b"3=1X"), line(b"3=99999999999999999999X") + line(b"3=1X"),
          line(b"2=3X3=", q=b"yQ", qs=2, qe=10),           # ALT 'Y' ends the record after one SNP row
          line(b"2=3X3=", b"-", q=b"yQ", qs=2, qe=10) + line(b"1=2X"),
          line(b"2X3D", t=b"rT", ts=1, te=8), b"#c\n\n" + line(b"12M"), line(b"11=1X"), line(b"12=1X"),   # the inclusive extra base
          line(b"8=4D", ts=10, te=22) + line(b"8=5I", qs=10, qe=23), line(b"0M5X")]
    for paf in ok:
        for snp, svlen in ((True, 0), (False, 0), (True, 1), (False, 2)):
            assert cpaf(w, ctx, paf, t_fa, ti, q_fa, qi, snp, svlen) == orc.call_paf(paf, t_fa, ti, q_fa, qi, snp=snp, svlen=svlen), (paf, snp, svlen)
    bad = [(line(b"4=1X") + line(b"4=20D"), 8), (line(b"12=3X"), 8), (line(b"0=2I"), 8), (line(b"0=2D"), 8), (line(b""), 8),
           (line(b"5=", tag=False), 6), (line(b"5=", b"-", t=b"rT", ts=0, te=5, tag=False), 10), (line(b"5=", b"-", t=b"rT", ts=0, te=5), 10),
           (line(b"5=", t=b"nope"), 11), (line(b"5=", q=b"nope"), 11), (b"a\tb\n", 9), (line(b"5=", b"-", ts=30, te=40), 8),
           (line(b"8=6D", ts=10, te=22), 8), (line(b"4=") + line(b"13=1X") + line(b"", tag=False), 8)]
    for paf, code in bad:
        with pytest.raises(orc.OracleError) as eo:
            orc.call_paf(paf, t_fa, ti, q_fa, qi, snp=True, svlen=0)
        with pytest.raises(w.WGAError) as eg:
            cpaf(w, ctx, paf, t_fa, ti, q_fa, qi, True, 0)
        assert eo.value.code == code == eg.value.code, (paf, eo.value.code, eg.value.code, eo.value.msg)
        if code in (6, 10):
            assert str(eg.value) == eo.value.msg, paf
    # errors that only exist with --snp do not fire without it
    assert cpaf(w, ctx, line(b"12=3X"), t_fa, ti, q_fa, qi, False, 0) == orc.call_paf(line(b"12=3X"), t_fa, ti, q_fa, qi, snp=False, svlen=0)


# ---------------------------------------------------------------- pafpseudo (row a17)
def _dedupe_starts(paf):
    seen, out = set(), []
    for l in paf.splitlines(keepends=True):
        f = l.split(b"\t")
        k = (f[0], f[5], f[7])
        if k not in seen:
            seen.add(k)
            out.append(l)
    return b"".join(out)


def test_paf_pseudo_parity(w, ctx):
    rng = random.Random(99)
    for it in range(5):
        paf, t_fa, _, q_fa, _ = make_paf2maf_case(rng, n_rec=120, width_t=rng.choice([50, 60, 80]), width_q=rng.choice([60, 61, 100]),
                                                  soft=it % 2 == 0)
        if it % 2:
            paf = _dedupe_starts(paf)
        fa = t_fa + q_fa
        fai = orc.make_fai(fa)
        F = w.Fasta(ctx, fa, fai)
        want = orc.paf_pseudo(paf, fa, fai)
        got = w.pseudo_maf_files(w.PAFReader(paf), F, ctx=ctx)
        assert list(got) == list(want)
        for k in want:
            assert got[k] == want[k], (it, k)
        want = orc.paf_pseudo(paf)
        got = w.pseudo_maf_files(w.PAFReader(paf), None, ctx=ctx)
        assert list(got) == list(want) and all(got[k] == want[k] for k in want)
        one = list(want)[-1]
        assert w.pseudo_maf_files(w.PAFReader(paf), F, one.decode(), ctx=ctx) == orc.paf_pseudo(paf, fa, fai, target=one.decode())


def test_paf_pseudo_edges_and_errors(w, ctx, tmp_path):
    fa = b">chrT\nACGTACGTACGGGGCCCCAATTTT\n>chrQ\nACGTACGTACGGCCCCAATTTTAA\n>q2\nTTTTGGGGCCCCAAAA\n>other\nACGTACGTACGTACGTACGTACGT\n>bad\nACGTRYACGT\n"
    fai = orc.make_fai(fa)
    F = w.Fasta(ctx, fa, fai)

    def line(cg, strand=b"+", q=b"chrQ", t=b"chrT", qs=1, qe=13, ts=2, te=14, qsz=24, tag=True):
        return b"\t".join([q, b"%d" % qsz, b"%d" % qs, b"%d" % qe, strand, t, b"24", b"%d" % ts, b"%d" % te, b"9", b"14", b"60"] +
                          ([b"cg:Z:" + cg] if tag else [])) + b"\n"
    paf = (line(b"4=1X2I3=2D2=") + line(b"4=", q=b"q2", qs=0, qe=4, ts=16, te=20, qsz=16) +
           line(b"3=1D2=", b"-", q=b"q2", qs=4, qe=9, ts=4, te=10, qsz=16) +
           line(b"6=", qs=14, qe=20, ts=10, te=16) + line(b"4=", qs=3, qe=7, ts=11, te=15))
    paf2 = line(b"2H4=2N", t=b"other", ts=0, te=4, qs=0, qe=4) + paf
    ok = [paf, paf2, b"#x\n\n" + line(b"5M3S4M"), line(b"4=9I", qs=0, qe=20), line(b"3M") + line(b"20M", ts=0, te=1),
          line(b"12=") + line(b"", ts=3, te=9) + line(b"M", ts=4, te=8, tag=False)]     # contained records are never looked at
    for p in ok:
        for f, ff, ffi in ((F, fa, fai), (None, None, None)):
            want = orc.paf_pseudo(p, ff, ffi)
            got = w.pseudo_maf_files(w.PAFReader(p), f, ctx=ctx)
            assert list(got) == list(want) and all(got[k] == want[k] for k in want), (p, f is None)
    assert w.pseudo_maf_files(w.PAFReader(paf2), F, "nope", ctx=ctx) == {}
    assert w.pseudo_maf_files(w.PAFReader(b""), F, ctx=ctx) == {}
    # the reference-shaped entry writes {out_dir}/{target}.maf
    w.generate_pesudo_maf(w.PAFReader(paf2), str(tmp_path), None, None, ctx=ctx, fa=F)
    assert sorted(os.listdir(tmp_path)) == ["chrT.maf", "other.maf"]
    assert (tmp_path / "chrT.maf").read_bytes() == orc.paf_pseudo(paf2, fa, fai)[b"chrT"]
    bad = [(line(b"4=9I"), 8), (line(b"20=3D"), 8), (line(b"5MM"), 5), (line(b"M"), 1), (line(b""), 8), (line(b"4=", q=b"zz"), 11),
           (line(b"4=", t=b"zz"), 11), (line(b"12=") + line(b"2=", ts=4, te=20, qs=0, qe=2), 8), (line(b"4=", tag=False), 6),
           (line(b"8=", b"-", q=b"bad", qs=0, qe=8, qsz=10), 10), (b"a\tb\n" + line(b"4="), 9), (line(b"4=", ts=30, te=40), 8),
           (line(b"4=", t=b"other") + line(b"4=", q=b"zz", ts=3, te=5, t=b"other"), 11)]
    for p, code in bad:
        with pytest.raises(orc.OracleError) as eo:
            orc.paf_pseudo(p, fa, fai)
        with pytest.raises(w.WGAError) as eg:
            w.pseudo_maf_files(w.PAFReader(p), F, ctx=ctx)
        assert eo.value.code == code == eg.value.code, (p, eo.value.code, eg.value.code, eo.value.msg)
    # equal target_start inside one (target, query) group: placed as binary_search_by + insert do (pseudomaf.rs:91-94)
    two = line(b"4=") + line(b"5=")
    assert w.pseudo_maf_files(w.PAFReader(two), F, ctx=ctx) == orc.paf_pseudo(two, fa, fai)


def test_paf_pseudo_equal_target_starts(w, ctx):
    """tie classes of every size between smaller and greater records of the group, in every insertion order: the order
    of the reference's sorted insert (the three-way binary search of core::slice, rustc 1.75 by the reference's flake.lock)"""
    fa = b">T\n" + b"ACGT" * 64 + b"\n>Q\n" + b"TTGCA" * 60 + b"\n>R\n" + b"GATTACA" * 40 + b"\n"
    fai = orc.make_fai(fa)
    F = w.Fasta(ctx, fa, fai)
    rng = random.Random(17)
    def line(q, ts, n, qs, strand=b"+"):
        return b"\t".join([q, b"280", b"%d" % qs, b"%d" % (qs + n), strand, b"T", b"256", b"%d" % ts, b"%d" % (ts + n), b"1", b"1", b"60",
                           b"cg:Z:%d=" % n]) + b"\n"
    for it in range(40):
        recs = []
        for q in (b"Q", b"R"):
            starts = [rng.choice([0, 10, 10, 10, 40, 40, 90, 120, 120]) for _ in range(rng.randrange(1, 14))]
            recs += [line(q, ts, rng.randrange(1, 30), rng.randrange(0, 100), rng.choice([b"+", b"-"])) for ts in starts]
        rng.shuffle(recs)
        paf = b"".join(recs)
        for use_fa in (True, False):
            want = orc.paf_pseudo(paf, fa, fai) if use_fa else orc.paf_pseudo(paf)
            got = w.pseudo_maf_files(w.PAFReader(paf), F if use_fa else None, ctx=ctx)
            assert list(got) == list(want) and all(got[k] == want[k] for k in want), (it, paf)
    # a big group: 3000 records, 60 distinct starts
    big = b"".join(line(b"Q", 4 * rng.randrange(60), rng.randrange(1, 12), rng.randrange(0, 200)) for _ in range(3000))
    want = orc.paf_pseudo(big)
    got = w.pseudo_maf_files(w.PAFReader(big), None, ctx=ctx)
    assert list(got) == list(want) and all(got[k] == want[k] for k in want)


# ---------------------------------------------------------------- validate (8(f) item 3)
def _validate(w, ctx, paf, fix=True):
    rep, fx = io.BytesIO(), io.BytesIO()
    w.parallel_validatepaf(w.PAFReader(paf), rep, fx if fix else None, ctx=ctx)
    return rep.getvalue(), fx.getvalue()


def test_validate_parity(w, ctx):
    assert _validate(w, ctx, TEST_PAF) == orc.paf_validate(TEST_PAF, True)
    assert _validate(w, ctx, b"") == orc.paf_validate(b"", True)
    rng = random.Random(8)
    for seed, kw in ((1, {}), (2, dict(m_only=1)), (3, dict(ops_median=2000.0, ops_sigma=1.0))):
        text = bs.gen_paf(bs.paf_params(1500 if seed < 3 else 200, seed=seed, **kw))
        lines = text.splitlines(keepends=True)
        for i in rng.sample(range(len(lines)), len(lines) // 5):   # falsify some ends, add awkward tags
            f = lines[i].rstrip(b"\n").split(b"\t")
            which = rng.randrange(4)
            if which in (0, 2):
                f[3] = b"%d" % (int(f[3]) + rng.randrange(1, 50))
            if which in (1, 2):
                f[8] = b"0x%x" % (int(f[8]) + rng.randrange(1, 9))
            if which == 3:
                f.insert(12, b'zz:Z:say "hi"')
                f.append(b"")
            lines[i] = b"\t".join(f) + b"\n"
        text = b"#c\n" + b"".join(lines)
        want = orc.paf_validate(text, True)
        assert _validate(w, ctx, text) == want
        assert want[0].count(b"\n") > 10
        assert _validate(w, ctx, text, fix=False) == (want[0], b"")
    for bad, code in ((TEST_PAF.replace(b"25M10I", b"25M10S", 1), 8), (TEST_PAF.replace(b"\tcg:Z:25M", b"\tcx:Z:25M", 1), 8),
                      (TEST_PAF.replace(b"\t300\t0\t200\t+", b"\t300\t0\tx\t+", 1), 9)):
        with pytest.raises(orc.OracleError) as eo:
            orc.paf_validate(bad, True)
        with pytest.raises(w.WGAError) as eg:
            _validate(w, ctx, bad)
        assert eo.value.code == code == eg.value.code


# ---------------------------------------------------------------- chain (8(f) item 2)
def test_chain_parity(w, ctx):
    m2c = lambda t, qn=None: w.run_bytes(w.maf2chain, w.MAFReader(t), qn, ctx=ctx)
    p2c = lambda t: w.run_bytes(w.paf2chain, w.PAFReader(t), ctx=ctx)
    assert m2c(TEST_MAF) == orc.maf2chain(TEST_MAF)
    assert p2c(TEST_PAF) == orc.paf2chain(TEST_PAF)
    assert m2c(b"") == b"" and p2c(b"") == b""
    maf = _maf([[b"s t 10 9 + 100 --ACGTAC-GTT", b"s q 5 8 - 50 GGAC-TACAGT-", b"s q2 0 9 + 60 TTACGTAC-GT-"],
                [b"s t 0 4 + 100 ACGT", b"s q 0 4 + 50 ACGA"], [b"s t 0 3 + 100 ---ACG", b"s q 0 3 + 50 TTT---"]])
    for qn in (None, "q"):
        assert m2c(maf, qn) == orc.maf2chain(maf, qn)
    for seed, kw in ((71, dict(p_neg=0.4)), (72, dict(len_median=30000.0, len_sigma=1.2, p_neg=0.2)), (73, dict(p_gap=0.05, p_bothgap=0.01))):
        text = bs.gen_maf(bs.maf_params(600 if seed != 72 else 60, seed=seed, **kw))
        want = orc.maf2chain(text)
        assert m2c(text) == want
        assert want == orc.paf2chain(orc.maf2paf(text))          # the same chains through PAF
        assert p2c(paf(w, ctx, text)) == want
    old = w.api.SLAB_BYTES
    try:                                                          # chain ids keep counting across slabs
        w.api.SLAB_BYTES = 150_000
        text = bs.gen_maf(bs.maf_params(700, seed=74, p_neg=0.3))
        assert m2c(text) == orc.maf2chain(text)
    finally:
        w.api.SLAB_BYTES = old
    for seed, kw in ((1, {}), (2, dict(m_only=1)), (3, dict(ops_median=3000.0, ops_sigma=1.0))):
        text = bs.gen_paf(bs.paf_params(1200 if seed < 3 else 150, seed=seed, **kw))
        assert p2c(text) == orc.paf2chain(text)
    line = lambda cg, strand=b"+", tag=True: b"\t".join([b"q", b"1000", b"100", b"200", strand, b"t", b"2000", b"300", b"420", b"1", b"2",
                                                        b"60"] + ([b"cg:Z:" + cg] if tag else [])) + b"\n"
    # op lengths on both sides of 2^16 inside one record (the row writer switches between its packed segmented scan and the
    # general three-scan path step by step, carrying the run / gap-group state across), gap groups across 32-op boundaries
    rng = random.Random(77)
    for _ in range(60):
        n_ops = rng.choice([3, 31, 32, 33, 64, 65, 150, 400])
        big_p = rng.choice([0.0, 0.02, 0.3, 1.0])
        ops, prev = [], None
        for k in range(n_ops):
            op = rng.choice(b"=XID" if rng.random() < 0.6 else b"ID=")
            ln = rng.randrange(65536, 400000) if rng.random() < big_p else rng.choice([1, 2, 7, 65535, rng.randrange(1, 3000)])
            ops.append(b"%d%c" % (ln, op))
        cg = b"".join(ops)
        for strand in (b"+", b"-"):
            l = b"\t".join([b"q", b"4000000000", b"100", b"200", strand, b"t", b"4000000000", b"300", b"420", b"1", b"2", b"60", b"cg:Z:" + cg]) + b"\n"
            assert p2c(l * 3) == orc.paf2chain(l * 3), cg
    for cg in (b"3I2D5I40=2X10I8=4D6=7I2D3I", b"5I3D", b"7M", b"1D1I1M1I1D"):
        for strand in (b"+", b"-"):
            assert p2c(line(cg, strand)) == orc.paf2chain(line(cg, strand))
    for bad, code in ((line(b"5M2S"), 5), (line(b"5MM"), 5), (line(b"M"), 1), (line(b""), 8), (line(b"5M", tag=False), 6), (b"x\ty\n", 9)):
        with pytest.raises(orc.OracleError) as eo:
            orc.paf2chain(bad)
        with pytest.raises(w.WGAError) as eg:
            p2c(bad)
        assert eo.value.code == code == eg.value.code
        if code in (1, 5, 6):
            assert str(eg.value) == eo.value.msg
    with pytest.raises(w.WGAError) as e:
        m2c(_maf([[b"s t 0 4 + 100 ACGT"]]))
    assert e.value.code == 8


def test_chain_reader_parity(w, ctx):
    """chain2paf / chain2maf (converter.rs:268-416) with the on-device chain reader (chain_in.cuh) against the oracle:
    well-formed files of every shape, the nom grammar's quirks, and every error class with its code"""
    c2p = lambda t: w.run_bytes(w.chain2paf, w.ChainReader(t), ctx=ctx)
    ch = orc.paf2chain(TEST_PAF)
    assert c2p(ch) == orc.chain2paf(ch)
    assert c2p(b"") == b""
    two = (b"chain 255 t 100 + 0 10 q 90 - 5 15 0\n4\t1\t2\n3\n\n# a note without the letter\n\n"
           b"chain 255 t 100 + 20 30 q 90 + 1 2 1\n7 2\n5")
    H = b"chain 1 t 9 + 0 1 q 9 + 0 1 0\n"
    good = [two, b"chain255 t 9 + 0 1 q 9 + 0 1 7 extra fields\r\n1 0 0 9\r\n", H + b"1\n2 3\nh\n9 9 9\n" + H + b"5\n",
            H + b"0\n", H + b"18446744073709551615 1 1\n1\n", b'chain 1e3 "t" 9 + 0 1 q"x 9 - 0 1 0\n4 0 3\n2\n\n\n\n' + H + b"1\n",
            H + b"1\n" * 3000 + b"\n" + H + b"2 1\n" * 70000 + b"9\n"]
    for t in good:
        assert c2p(t) == orc.chain2paf(t), t[:80]
    bads = [b"xchain 1 t 9 + 0 1 q 9 + 0 1 0\n1\n", b"xchain\n", H + b"\n", H + b"\n" + b"#" * 12, H[:-1],
            b"chain a t 9 + 0 1 q 9 + 0 1 0\n1\n", b"chain 1 t x + 0 1 q 9 + 0 1 0\n1\n", b"chain 1 t 9 * 0 1 q 9 + 0 1 0\n1\n",
            b"chain 1 t 9 + 0 1 q 9 + 0 1\n1\n", H + b"1\n2 x\n", H + b"1\n \n", H + b"1\n\n# a comment\n" + H + b"1\n",
            H + b"1\n12 c\n", H + b"1\n" + H + b"x\n" + H + b"1\n", H + b"1\n\nzzz\n" + b"chain 1 t 9 + 0 1\n1\n" + b"chain x\n",
            b"\n" + H + b"1\n", H + b"1\n" + b"chain 1 t 9 + 0 1 q 9 + 0 1 0\n" + b"#" * 20 + b"\n", H + b"99999999999999999999\n"]
    for t in bads:
        with pytest.raises(orc.OracleError) as eo:
            orc.chain2paf(t)
        with pytest.raises(w.WGAError) as eg:
            c2p(t)
        assert eg.value.code == eo.value.code, (t, eo.value, eg.value)
        if eo.value.code in (1, 2, 12, 13):
            assert str(eg.value) == eo.value.msg, t
    # synthetic: the chains of generated PAFs / MAFs (long chains, both strands), then back
    for seed, kw in ((1, {}), (2, dict(m_only=1)), (3, dict(ops_median=3000.0, ops_sigma=1.0))):
        text = orc.paf2chain(bs.gen_paf(bs.paf_params(1500 if seed < 3 else 200, seed=seed, **kw)))
        got = c2p(text)
        assert got == orc.chain2paf(text)
        assert w.run_bytes(w.paf2chain, w.PAFReader(got), ctx=ctx) == orc.paf2chain(got)   # and back (not the identity on '-': chain.rs:133-138)
    text = orc.maf2chain(bs.gen_maf(bs.maf_params(800, seed=75, p_neg=0.3)))
    assert c2p(text) == orc.chain2paf(text)
    assert c2p(text.replace(b"\t", b" ")) == orc.chain2paf(text.replace(b"\t", b" "))
    # ---- chain2maf
    from bench_support.paf_legs import make_fasta
    fa, fai = make_fasta(3, 2, 300000, seed=9)
    F = w.Fasta(ctx, fa, fai)
    c2m = lambda t: w.run_bytes(w.chain2maf, w.ChainReader(t), "t.fa", "q.fa", ctx=ctx, target=F, query=F)
    want = lambda t: orc.chain2maf(t, fa, fai, fa, fai)
    paf5 = bs.gen_paf(bs.paf_params(400, seed=93, n_genomes=3, n_chr=2, chr_size=300000, ops_median=20.0, ops_max=200, p_neg=0.0))
    chain5 = orc.paf2chain(paf5)
    assert c2m(chain5) == want(chain5)
    assert orc.maf2chain(want(chain5)) == chain5                                  # README.md:225 round trip
    # '-' records: paf2chain's header quirk (chain.rs:133-138) leaves query coordinates that do not match the blocks, and
    # chain2maf then panics in insert_str on the first such record -- the same record here
    chain6 = orc.paf2chain(bs.gen_paf(bs.paf_params(400, seed=94, n_genomes=3, n_chr=2, chr_size=300000, ops_median=20.0, ops_max=200,
                                                    p_neg=0.5)))
    with pytest.raises(orc.OracleError) as eo:
        want(chain6)
    with pytest.raises(w.WGAError) as eg:
        c2m(chain6)
    assert eo.value.code == eg.value.code == 8
    small = (b"chain 9 g0#chr1 30000 + 2 12 g1#chr2 30000 - 3 12 5\n3 1 2\n4\n\nchain 9 g2#chr1 30000 - 0 5 g0#chr2 30000 + 10 15 6\n5\n"
             b"chain 9 g0#chr1 300000 + 299990 300000 g0#chr1 300000 + 0 10 7\n4 0 0\n6\n")
    assert c2m(small) == want(small)
    assert c2m(b"") == want(b"")
    fa_bad = fa + b">k\nACGTRACGT\n"
    fai_bad = orc.make_fai(fa_bad)
    FB = w.Fasta(ctx, fa_bad, fai_bad)
    for t in (b"chain 9 g0#chr1 30000 + 2 4 g1#chr1 30000 + 3 5 0\n9 0 1\n",            # insert past the end of the target row
              b"chain 9 zz 20 + 2 4 g0#chr1 30000 + 3 5 0\n2\n",                          # sequence not in the .fai
              b"chain 9 g0#chr1 30000 + 2 4 k 9 - 0 5 0\n2\n",                            # reverse complement of a non-ACGTN base
              small + b"chain 9 g0#chr1 30000 + 2 4 k 9 - 0 5 0\n2\n" + b"chain x\n1\n",    # later stage error BEFORE a later reader error
              small + b"chain x\n1\n" + b"chain 9 g0#chr1 30000 + 2 4 k 9 - 0 5 0\n2\n",    # reader error before a later stage error
              b"chain 9 g0#chr1 30000 + 2 4 k 9 - 0 5 0\nx\n"):                           # same record: the reader comes first
        with pytest.raises(orc.OracleError) as eo:
            orc.chain2maf(t, fa_bad, fai_bad, fa_bad, fai_bad)
        with pytest.raises(w.WGAError) as eg:
            w.run_bytes(w.chain2maf, w.ChainReader(t), "t.fa", "q.fa", ctx=ctx, target=FB, query=FB)
        assert eg.value.code == eo.value.code, (t, eo.value, eg.value)


# ---------------------------------------------------------------- fuzz: many tiny MAFs with awkward text
def _fuzz_maf(rng):
    ws = lambda: rng.choice([b" ", b"\t", b"  ", b" \t", b"\t\t "])
    nl = lambda: rng.choice([b"\n", b"\n", b"\n", b"\r\n"])
    out = [rng.choice([b"##maf version=1", b"a score=0", b"s hdr 0 1 + 5 A", b""]) + nl()]
    for _ in range(rng.randrange(1, 7)):
        out.append(rng.choice([b"a score=1", b"", b"# c", b"a", b"i x y", b"q t 99", b"e z"]) + nl())
        n = rng.randrange(0, 70)
        k = rng.choice([1] + [2] * 14 + [3] * 5)
        names = [b"t.chr1"] + [rng.choice([b"q", b"q", b"q2", b"query.long_name-1"]) for _ in range(k - 1)]
        for j in range(k):
            ln = n if rng.random() < 0.85 else max(0, n + rng.randrange(-3, 4))
            alpha = b"ACGTACGTACGTNacgt-" if rng.random() < 0.7 else b"AC-"
            seq = bytes(rng.choice(alpha) for _ in range(ln))
            if j == 0 and rng.random() < 0.5:
                seq = seq.replace(b"-", b"A")
            ali = sum(1 for c in seq if c != 45)
            start = rng.randrange(0, 1000)
            size = start + ali + rng.randrange(0, 50)
            fields = [b"s", names[j], b"%d" % start, b"%d" % ali, rng.choice([b"+", b"+", b"-"]), b"%d" % size, seq]
            if rng.random() < 0.008:
                fields[rng.randrange(2, 6)] = rng.choice([b"x1", b"", b"-3", b"*"])
            if rng.random() < 0.006:
                fields.append(b"extra")
            if rng.random() < 0.006:
                fields = fields[:rng.randrange(2, 7)]
            line = rng.choice([b"", b"", b""]) + b"".join(f + ws() for f in fields[:-1]) + fields[-1]
            out.append(line + nl())
        if rng.random() < 0.8:
            out.append(nl())
    text = b"".join(out)
    return text if rng.random() < 0.9 else text.rstrip(b"\r\n")


def test_fuzz_small_mafs(w, ctx):
    rng = random.Random(20260924)
    from wgatools_b200._lib import CallOpts
    n_err = n_ok = 0
    for it in range(250):
        text = _fuzz_maf(rng)
        qn = rng.choice([None, None, None, None, None, "q", "q2"])
        cases = [
            ("maf2paf", lambda: orc.maf2paf(text, qn), lambda: w.run_bytes(w.maf2paf, w.MAFReader(text), qn, ctx=ctx)),
            ("stat", lambda: orc.stat_maf(text, False, qn), lambda: w.run_bytes(w.stat_maf, w.MAFReader(text), False, qn, ctx=ctx)),
            ("stat-each", lambda: orc.stat_maf(text, True, qn), lambda: w.run_bytes(w.stat_maf, w.MAFReader(text), True, qn, ctx=ctx)),
            ("chain", lambda: orc.maf2chain(text, qn), lambda: w.run_bytes(w.maf2chain, w.MAFReader(text), qn, ctx=ctx)),
            ("call", lambda: orc.call_maf(text, snp=True, inv=True, svlen=1, query_name=qn, chunk_size=7),
             lambda: w.run_bytes(w.call_var_maf, w.MAFReader(text), None, True, True, 1, False, None, qn, None, 7, ctx=ctx)),
        ]
        for name, fo, fg in cases:
            want = err_o = None
            try:
                want = fo()
            except orc.OracleError as e:
                err_o = e
            if err_o is None:
                assert fg() == want, (it, name, text)
                n_ok += 1
            else:
                with pytest.raises(w.WGAError) as eg:
                    fg()
                assert eg.value.code == err_o.code, (it, name, text, err_o.msg, str(eg.value))
                if err_o.code in (1, 2, 3, 4, 7) and name != "call":
                    assert str(eg.value) == err_o.msg, (it, name, text)
                n_err += 1
    assert n_ok > 300 and n_err > 50


# ---------------------------------------------------------------- pafcov (8(f) item 3)
def test_pafcov_parity(w, ctx):
    cov = lambda t: w.run_bytes(w.pafcov, w.PAFReader(t), ctx=ctx)
    line = lambda cg, t=b"t", tl=20, ts=2, tag=True: b"\t".join([b"q", b"100", b"0", b"10", b"+", t, b"%d" % tl, b"%d" % ts, b"12", b"1", b"2",
                                                                 b"60"] + ([b"cg:Z:" + cg] if tag else [])) + b"\n"
    small = line(b"3=2X2I4M3D5=") + line(b"30M", ts=15, tl=999) + line(b"2=1N2H1=", t=b"u", tl=7, ts=0) + line(b"4=", ts=100) + \
        line(b"1=", t=b"zero", tl=0, ts=0)
    assert cov(small) == orc.pafcov(small)
    assert cov(TEST_PAF) == orc.pafcov(TEST_PAF)
    assert cov(b"") == b"" and cov(b"#only a comment\n") == b""
    for seed, kw in ((1, dict(chr_size=30_000, n_genomes=2, n_chr=2)), (2, dict(chr_size=200_000, n_genomes=1, n_chr=3, m_only=1))):
        text = bs.gen_paf(bs.paf_params(400, seed=seed, ops_median=40.0, **kw))
        want = orc.pafcov(text)
        assert cov(text) == want
        assert max(int(l.rsplit(b"\t", 1)[1]) for l in want.splitlines()) >= 2
    for bad, code in ((line(b"5MM"), 5), (line(b"M"), 1), (line(b""), 8), (line(b"3=", tag=False), 6), (b"a\tb\n", 9),
                      (line(b"3=") + line(b"2=9999999999999999999999M"), 1)):
        with pytest.raises(orc.OracleError) as eo:
            orc.pafcov(bad)
        with pytest.raises(w.WGAError) as eg:
            cov(bad)
        assert eo.value.code == code == eg.value.code
        if code in (1, 5, 6):
            assert str(eg.value) == eo.value.msg


# ---------------------------------------------------------------- cs:Z: standing in for cg:Z: (row a7, paf.rs:122-140,159-218)
def _cg_to_cs(rng, cg, noise=True):
    """a cs:Z: spelling whose conversion gives the ops of `cg` back (M/= -> `:N` pieces, X -> `*xy` runs, I/D -> +seq / -seq),
    sprinkled with the things cs_to_cigar skips or ignores"""
    import re as _re
    out = []
    for n, op in _re.findall(rb"(\d+)([M=XID])", cg):
        n = int(n)
        if op in b"M=":
            a = rng.randrange(0, n + 1) if n > 1 and rng.random() < 0.5 else n
            out.append(b":%d" % a)
            if a != n:
                out.append(b":%0*d" % (rng.randrange(1, 4), n - a))
        elif op == b"X":
            out.append(b"".join(b"*" + bytes(rng.choice(b"acgt") for _ in range(2)) for _ in range(n)))
        else:
            out.append((b"+" if op == b"I" else b"-") + bytes(rng.choice(b"acgtACGTn") for _ in range(n)))
        if noise and rng.random() < 0.15:
            out.append(rng.choice([b"=ACGT", b"~", b"*AT", b":", b" ", b"=a", b"*a"]))
    return b"".join(out)


def _cs_paf(rng, paf, frac=0.7):
    lines = []
    for l in paf.splitlines(keepends=True):
        f = l.rstrip(b"\n").split(b"\t")
        if len(f) > 12 and rng.random() < frac:
            for k in range(12, len(f)):
                if f[k].startswith(b"cg:Z:") and not re.search(rb"(^|\D)0+[MIDX=]", f[k][5:]):
                    f[k] = b"cs:Z:" + _cg_to_cs(rng, f[k][5:])
                    if rng.random() < 0.3:
                        f.append(b"cs:Z::999")       # only the FIRST cs:Z: tag counts
                    break
        lines.append(b"\t".join(f) + b"\n")
    return b"".join(lines)


def test_cs_tag_all_paf_commands(w, ctx):
    rng = random.Random(4242)
    # stat / paf2chain / pafcov / validate over the synthetic config-4 shape, most records re-spelled with cs:Z:
    for seed, kw in ((1, {}), (3, dict(ops_median=1500.0, ops_sigma=1.0))):
        text = _cs_paf(rng, bs.gen_paf(bs.paf_params(800 if seed == 1 else 120, seed=seed, **kw)))
        assert text.count(b"cs:Z:") > 50
        check_paf(w, ctx, text)
        assert w.run_bytes(w.paf2chain, w.PAFReader(text), ctx=ctx) == orc.paf2chain(text)
        assert _validate(w, ctx, text) == orc.paf_validate(text, True)
    small = _cs_paf(rng, bs.gen_paf(bs.paf_params(200, seed=9, ops_median=30.0, chr_size=30_000, n_genomes=2, n_chr=2)))
    assert w.run_bytes(w.pafcov, w.PAFReader(small), ctx=ctx) == orc.pafcov(small)
    # the FASTA commands
    for it in range(3):
        case = make_paf2maf_case(rng, n_rec=60, soft=it % 2 == 0)
        paf = _cs_paf(rng, case[0])
        assert paf.count(b"cs:Z:") > 10
        assert p2m(w, ctx, paf, *case[1:]) == orc.paf2maf(paf, *case[1:])
        assert cpaf(w, ctx, paf, *case[1:], snp=True, svlen=0, sample="s") == orc.call_paf(paf, *case[1:], snp=True, svlen=0, sample="s")
        pp = _dedupe_starts(paf)
        fa = case[1] + case[3]
        fai = orc.make_fai(fa)
        want = orc.paf_pseudo(pp, fa, fai)
        got = w.pseudo_maf_files(w.PAFReader(pp), w.Fasta(ctx, fa, fai), ctx=ctx)
        assert list(got) == list(want) and all(got[k] == want[k] for k in want)
    # quirks and the two panics
    base = b"q\t100\t0\t12\t+\tt\t100\t0\t13\t1\t2\t60\t"
    ok = [base + b"cs:Z::6-a*ac+tt:3\n", base + b"tp:A:P\tcs:Z:=ACGT:3*AT*at~~:*ga-N+nn:0:0\n",
          base + b"cs:Z::3\tcg:Z:7=\n",                          # cg:Z: wins wherever it stands
          base + b"cs:Z::18446744073709551615\n", base + b"cs:Z::6\n" + base.replace(b"+", b"-") + b"cs:Z:*ac*ac:2-acgt\n"]
    for t in ok:
        check_paf(w, ctx, t)
    for t, code in [(base + b"cs:Z:junk\tcs:Z::6\n", 8), (base + b"cs:Z:\n", 8), (base + b"cs:Z::18446744073709551616\n", 8),
                    (base + b"cs:Z::5\n" + base + b"cs:Z::99999999999999999999999\n" + base + b"\n", 8)]:
        with pytest.raises(orc.OracleError) as eo:
            orc.stat_paf(t)
        with pytest.raises(w.WGAError) as eg:
            statp(w, ctx, t)
        assert eo.value.code == code == eg.value.code, (t, eo.value.code, eg.value.code)


# ---------------------------------------------------------------- filter -f paf (8(f) item 3, filter.rs:41-63,92-160)
def test_paf_filter_parity(w, ctx):
    flt = lambda t, b=0, q=0: w.run_bytes(w.filter_paf, w.PAFReader(t), b, q, ctx=ctx)
    fla = lambda t, a: w.run_bytes(w.filter_paf_align_pair, w.PAFReader(t), a, ctx=ctx)
    assert flt(TEST_PAF) == orc.paf_filter(TEST_PAF) == TEST_PAF
    assert fla(TEST_PAF, 0) == TEST_PAF and fla(b"", 5) == b"" and flt(b"#only a comment\n") == b""
    rng = random.Random(31)
    for seed, kw in ((1, {}), (4, dict(ops_median=3.0, ops_sigma=0.8, n_genomes=3, n_chr=2)), (5, dict(n_genomes=2, n_chr=1))):
        text = bs.gen_paf(bs.paf_params(3000, seed=seed, **kw))
        lines = text.splitlines(keepends=True)
        for i in rng.sample(range(len(lines)), len(lines) // 10):       # awkward tags, csv hex ints, a wrapping span
            f = lines[i].rstrip(b"\n").split(b"\t")
            which = rng.randrange(4)
            if which == 0:
                f.insert(12, b'zz:Z:say "hi"')
            elif which == 1:
                f[8] = b"0x%x" % int(f[8])
            elif which == 2:
                f[7], f[8] = f[8], f[7]
            else:
                f = f[:12]
            lines[i] = b"\t".join(f) + b"\n"
        text = b"#c\n" + b"".join(lines)
        spans = sorted(int(l.split(b"\t")[8], 0) - int(l.split(b"\t")[7], 0) for l in lines)
        for b, q in ((0, 0), (spans[len(spans) // 2], 0), (0, 250000001), (spans[len(spans) // 4], 1000)):
            want = orc.paf_filter(text, b, q)
            assert flt(text, b, q) == want
        sums = {}
        for l in lines:
            f = l.split(b"\t")
            sums[(f[0], f[5])] = (sums.get((f[0], f[5]), 0) + int(f[8], 0) - int(f[7], 0)) % 2 ** 64
        cuts = sorted(sums.values())
        for a in (0, cuts[len(cuts) // 2], cuts[-1], cuts[-1] + 1, 2 ** 64 - 1):
            want = orc.paf_filter(text, min_align=a)
            assert fla(text, a) == want, (seed, a)
        assert 0 < orc.paf_filter(text, min_align=cuts[len(cuts) // 2]).count(b"\n") < len(lines)
    for bad, code in ((b"q\t1\t2\n", 9), (TEST_PAF + b"a\tb\tc\td\t+\tt\t1\t2\t3\t4\t5\t6\n", 9)):
        for a in (None, 3):
            with pytest.raises(orc.OracleError) as eo:
                orc.paf_filter(bad, min_align=a)
            with pytest.raises(w.WGAError) as eg:
                fla(bad, a) if a is not None else flt(bad)
            assert eo.value.code == code == eg.value.code


# ---------------------------------------------------------------- dotplot base-level plot data (8(f) item 4, cigar.rs:807-985)
def test_plotdata_parity(w, ctx):
    pm = lambda t, c=0, qn=None: w.run_bytes(w.generate_maf_basedata, w.MAFReader(t), c, qn, ctx=ctx)
    pp = lambda t, c=0: w.run_bytes(w.generate_paf_basedata, w.PAFReader(t), c, ctx=ctx)
    for c in (0, 5, 10, 19, 20, 10 ** 9):
        assert pp(TEST_PAF, c) == orc.paf_plotdata(TEST_PAF, c)
        assert pm(TEST_MAF, c) == orc.maf_plotdata(TEST_MAF, c)
    assert pp(b"") == b"" and pm(b"##maf version=1\n") == b"" and pp(b"#x\n") == b""
    for seed in (3, 4):
        maf = bs.gen_maf(bs.maf_params(400, seed=seed, p_bothgap=1e-3))
        for c in (0, 1, 3, 50):
            assert pm(maf, c) == orc.maf_plotdata(maf, c), (seed, c)
    long_maf = bs.gen_maf(bs.maf_params(6, seed=11, len_median=200000.0, len_sigma=0.3))
    assert pm(long_maf, 2) == orc.maf_plotdata(long_maf, 2)
    rng = random.Random(12)
    for seed, kw in ((1, {}), (3, dict(ops_median=1500.0, ops_sigma=1.0)), (4, dict(ops_median=3.0, ops_sigma=0.8, n_genomes=3, n_chr=2))):
        text = bs.gen_paf(bs.paf_params(1500 if seed != 3 else 150, seed=seed, **kw))
        lines = text.splitlines(keepends=True)
        for i in rng.sample(range(len(lines)), len(lines) // 8):       # ops the match arm ignores, names the csv writer quotes
            f = lines[i].rstrip(b"\n").split(b"\t")
            k = [j for j in range(12, len(f)) if f[j].startswith(b"cg:Z:")][0]
            which = rng.randrange(4)
            if which == 0:
                f[k] = f[k] + b"7N3S" + rng.choice([b"12=", b"4I", b""])
            elif which == 1:
                f[k] = b"cg:Z:5H" + f[k][5:]
            elif which == 2:
                f[0] = b'a,b'
            else:
                f[5] = b'say"hi'
            lines[i] = b"\t".join(f) + b"\n"
        text = b"".join(lines)
        for c in (0, 2, 100):
            assert pp(text, c) == orc.paf_plotdata(text, c), (seed, c)
    base = b"q\t100\t7\t12\t-\tt\t100\t3\t13\t1\t2\t60\t"
    for t, code in [(base + b"cg:Z:5MM\n", 5), (base + b"cg:Z:M\n", 1), (base + b"cg:Z:\n", 8), (base + b"NM:i:1\n", 6), (b"a\tb\n", 9),
                    (base + b"cg:Z:3=\n" + base + b"cg:Z:2=99999999999999999999999M\n", 1)]:
        with pytest.raises(orc.OracleError) as eo:
            orc.paf_plotdata(t)
        with pytest.raises(w.WGAError) as eg:
            pp(t)
        assert eo.value.code == code == eg.value.code, (t, eo.value.code, eg.value.code)


# ---------------------------------------------------------------- quoted csv fields in PAF input (paf.rs:22-31, csv quoting on)
def _quote_fields(rng, paf, frac=0.5):
    """quote some fields the way a csv writer with QuoteStyle::Always would: "content" (content has no special byte)"""
    out = []
    for l in paf.splitlines(keepends=True):
        if l.startswith(b"#") or not l.strip():
            out.append(l)
            continue
        f = l.rstrip(b"\n").split(b"\t")
        f = [b'"' + x + b'"' if rng.random() < frac else x for x in f]
        out.append(b"\t".join(f) + b"\n")
    return b"".join(out)


def test_paf_quoted_fields(w, ctx):
    rng = random.Random(77)
    plain = b"#c\n" + bs.gen_paf(bs.paf_params(600, seed=11, ops_median=40.0, chr_size=30_000, n_genomes=2, n_chr=2))
    q = _quote_fields(rng, plain)
    assert q.count(b'"') > 2000
    # the reader hands the deserializer the unquoted content: every command gives what it gives on the plain file
    assert orc.stat_paf(q) == orc.stat_paf(plain)
    assert statp(w, ctx, q) == orc.stat_paf(q)
    assert statp(w, ctx, q, True) == orc.stat_paf(q, True)
    assert w.run_bytes(w.paf2chain, w.PAFReader(q), ctx=ctx) == orc.paf2chain(q)
    assert _validate(w, ctx, q) == orc.paf_validate(q, True)
    assert w.run_bytes(w.filter_paf_align_pair, w.PAFReader(q), 1000, ctx=ctx) == orc.paf_filter(q, min_align=1000)
    assert w.run_bytes(w.pafcov, w.PAFReader(q), ctx=ctx) == orc.pafcov(q)
    assert w.run_bytes(w.generate_paf_basedata, w.PAFReader(q), 3, ctx=ctx) == orc.paf_plotdata(q, 3)
    case = make_paf2maf_case(rng, n_rec=40)
    qp = _quote_fields(rng, case[0])
    assert p2m(w, ctx, qp, *case[1:]) == orc.paf2maf(qp, *case[1:])
    assert cpaf(w, ctx, qp, *case[1:], snp=True, svlen=0, sample="s") == orc.call_paf(qp, *case[1:], snp=True, svlen=0, sample="s")
    fa = case[1] + case[3]
    fai = orc.make_fai(fa)
    want = orc.paf_pseudo(qp, fa, fai)
    got = w.pseudo_maf_files(w.PAFReader(qp), w.Fasta(ctx, fa, fai), ctx=ctx)
    assert list(got) == list(want) and all(got[k] == want[k] for k in want)
    # an error behind the quotes is still the reference's error
    bad = q + b'"q"\t"x"\t2\n'
    with pytest.raises(orc.OracleError) as eo:
        orc.stat_paf(bad)
    with pytest.raises(w.WGAError) as eg:
        statp(w, ctx, bad)
    assert eo.value.code == eg.value.code == 9
    # what stays refused: a quote, a tab or a line break inside the quotes, text behind the closing quote
    line = plain.splitlines(keepends=True)[1]
    f = line.rstrip(b"\n").split(b"\t")
    for weird in (b'"a""b"', b'"a\tb"', b'"ab"c', b'"#ab"'):
        g = [weird] + f[1:]
        with pytest.raises(w.WGAError) as e:
            statp(w, ctx, plain + b"\t".join(g) + b"\n")
        assert e.value.code == 102, weird
