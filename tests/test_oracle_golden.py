"""Pins the CPU oracle (oracle/wga_oracle.c) to the golden vectors the reference ships.

Sources (all under /root/reference, copied as data fixtures into tests/golden/):
  test/test.maf            -> tests/golden/ref_test.maf
  test/testdotplot.paf     -> tests/golden/ref_testdotplot.paf
  README.md:324-343        -> tests/golden/ref_readme_call_s_l0.vcf  (`wgatools call test/test.maf -s -l0`)
The reference has no unit tests of its own; these are all the vectors that exist.
"""
import os

import numpy as np
import random

import pytest
from hypothesis import given, settings, strategies as st

import oracle_lib as orc
import pyspec

G = os.path.join(os.path.dirname(__file__), "golden")
TEST_MAF = open(os.path.join(G, "ref_test.maf"), "rb").read()
TEST_PAF = open(os.path.join(G, "ref_testdotplot.paf"), "rb").read()
README_VCF = open(os.path.join(G, "ref_readme_call_s_l0.vcf")).read().splitlines()

CIGAR = "109=1D243=1X12=1X138=1X177=1X31=1X133=8I18=1X100=2X7=1X22="
STAT_HDR = ("ref_name\tref_size\tref_start\tquery_name\tquery_size\tquery_start\taligned_size\tunaligned_size\t"
            "identity\tsimilarity\tmatched\tmismatched\tins_event\tdel_event\tins_size\tdel_size\tinv_event\t"
            "inv_size\tinv_ins_event\tinv_ins_size\tinv_del_event\tinv_del_size")


def test_maf2paf_test_maf():
    out = orc.maf2paf(TEST_MAF).decode()
    assert out == ("query.chr8\t183119688\t181989421\t181990428\t+\tref.chr8\t182411202\t181469925\t181470925\t"
                   "990\t1008\t255\tNM:i:18\tcg:Z:" + CIGAR + "\n")


def test_stat_test_maf():
    out = orc.stat_maf(TEST_MAF).decode().splitlines()
    assert out[0] == STAT_HDR
    assert out[1].split("\t") == ("ref.chr8 182411202 181469925 query.chr8 183119688 181989421 1000 182410202 "
                                  "0.99 0.999 990 9 1 1 8 1 0 0.0 0 0 0 0").split()
    assert len(out) == 2


def test_stat_each_leaves_unaligned_zero():
    row = orc.stat_maf(TEST_MAF, each=True).decode().splitlines()[1].split("\t")
    assert row[7] == "0" and row[6] == "1000"


def test_stat_paf_testdotplot():
    out = orc.stat_paf(TEST_PAF).decode().splitlines()
    assert out[0] == STAT_HDR
    assert out[1].split("\t") == "B 300 0 A 300 0 250 50 0.84 0.84 210 0 2 2 30 30 1 50.0 1 10 1 10".split()


def test_call_readme_vcf():
    got = orc.call_maf(TEST_MAF, snp=True, svlen=0).decode().splitlines()
    assert len(got) == len(README_VCF) == 19
    for a, b in zip(got, README_VCF):
        fa, fb = a.split("\t"), b.split("\t")
        if a.startswith("#") or "SVTYPE" in a:
            assert a == b  # header + INS/DEL rows byte-identical
        else:
            # SNP rows: README predates the GT:QI format for SNPs (caller.rs:582-593); cols 1-8 pinned
            assert fa[:8] == fb[:8]
            assert fa[8] == "GT:QI" and fa[9].startswith("1|1:query.chr8@") and fa[9].endswith("@P")


def test_call_default_svlen50_emits_nothing_for_test_maf():
    got = orc.call_maf(TEST_MAF).decode().splitlines()
    assert len(got) == 8  # header only


def test_natord():
    names = ["chr10", "chr2", "chr1", "chrX", "chr01", "chr1a"]
    import functools
    s = sorted(names, key=functools.cmp_to_key(orc.natord))
    assert s.index("chr1") < s.index("chr2") < s.index("chr10")
    assert orc.natord("a 1", "a1") == 0  # whitespace skipped
    assert orc.natord("x9", "x10") < 0 and orc.natord("x10", "x9") > 0
    assert orc.natord("1.001", "1.002") < 0 and orc.natord("1.010", "1.02") < 0  # fractional (leading zero)


def test_fmt_f32():
    import numpy as np
    assert orc.fmt_f32(0.99) == "0.99"
    assert orc.fmt_f32(float(np.float32(990) / np.float32(1000))) == "0.99"
    assert orc.fmt_f32(0.0) == "0.0" and orc.fmt_f32(50.0) == "50.0" and orc.fmt_f32(1.0) == "1.0"
    assert orc.fmt_f32(float("nan")) == "NaN"
    assert orc.fmt_f32(1.5e10) == "15000000000.0"
    assert orc.fmt_f32(float(np.float32(1e-7))) == "1e-7"
    assert orc.fmt_f32(float(np.float32(1.25e20))) == "1.25e20"
    rng = random.Random(1)
    for _ in range(2000):
        a, b = rng.randrange(0, 10**6), rng.randrange(1, 10**6)
        v = np.float32(a) / np.float32(b)
        s = orc.fmt_f32(float(v))
        assert np.float32(s) == v
        ref = np.format_float_positional(v, unique=True, trim="0")
        if 1e-5 <= float(v) < 1e13:
            assert s == ref


def test_paf_cigar_quirks():
    assert orc.paf_cigar_counts(b"10M5X")[:2] == [10, 5]  # Q3: M counts as matched
    assert orc.paf_cigar_counts(b"3I4D", neg=True) == [0, 0, 0, 0, 0, 0, 1, 3, 1, 4, 1]
    for bad in (b"10S5M", b"5H", b"5N", b"3P"):
        with pytest.raises(orc.OracleError) as e:
            orc.paf_cigar_counts(bad)
        assert e.value.code == 5  # Q4: ops outside M=XID are errors
    with pytest.raises(orc.OracleError) as e:
        orc.paf_cigar_counts(b"10MM")
    assert e.value.code == 5 and "`MM`" in e.value.msg
    with pytest.raises(orc.OracleError) as e:
        orc.paf_cigar_counts(b"M")
    assert e.value.code == 1  # empty length
    with pytest.raises(orc.OracleError) as e:
        orc.paf_cigar_counts(b"12")
    assert e.value.code == 5 and "``" in e.value.msg  # digits without op
    with pytest.raises(orc.OracleError) as e:
        orc.paf_cigar_counts(b"99999999999999999999999M")
    assert e.value.code == 1


def _maf(blocks, header=b"##maf version=1\n"):
    out = [header]
    for lines in blocks:
        out.append(b"a score=0\n")
        for l in lines:
            out.append(l + b"\n")
        out.append(b"\n")
    return b"".join(out)


def test_maf_reader_quirks():
    blk = [b"s t 0 4 + 100 ACGT", b"s q 0 4 + 100 ACGA"]
    # Q9: the first line is always swallowed, even if it is an s-line
    txt = b"s t 0 4 + 100 AAAA\n" + b"\n".join(blk) + b"\n"
    assert orc.maf2paf(txt).count(b"\n") == 1
    # non-'s' lines split blocks; 'a' line optional
    txt = _maf([blk]) + b"\n".join(blk) + b"\ni something\n" + b"\n".join(blk)
    assert orc.maf2paf(txt).count(b"\n") == 3
    # Q10: 1-s-line block panics in maf2paf/stat, is skipped by call
    one = _maf([[b"s t 0 4 + 100 ACGT"]])
    with pytest.raises(orc.OracleError) as e:
        orc.maf2paf(one)
    assert e.value.code == 8
    assert len(orc.call_maf(one).splitlines()) == 8
    # surplus / missing fields, bad ints, bad strand
    for bad, code in ((b"s t 0 4 + 100 ACGT X", 4), (b"s t 0 4 + 100", 3), (b"s t x 4 + 100 ACGT", 1),
                      (b"s t 0 4 * 100 ACGT", 2), (b"s t 0 -4 + 100 ACGT", 1)):
        with pytest.raises(orc.OracleError) as e:
            orc.stat_maf(_maf([[bad, b"s q 0 4 + 100 ACGT"]]))
        assert e.value.code == code, bad
    # Q14 zip truncation, Q1 both-gap '=' , Q2 case-sensitive
    txt = _maf([[b"s t 0 4 + 100 A-cGTTT", b"s q 0 4 + 100 A-CG"]])
    assert orc.maf2paf(txt).split(b"cg:Z:")[1] == b"2=1X1=\n"
    # '+' accepted by u64 parse; CRLF line ends
    txt = b"#h\r\ns t +5 4 + 100 ACGT\r\ns q 0 4 - 100 ACGT\r\n"
    row = orc.maf2paf(txt).split(b"\t")
    assert row[7] == b"5" and row[2] == b"96" and row[3] == b"100" and row[4] == b"-"


def test_query_name_selection():
    txt = _maf([[b"s t 0 4 + 100 ACGT", b"s q1 0 4 + 100 ACGA", b"s q2 0 4 + 100 AC-T"]])
    assert orc.maf2paf(txt, query_name="q2").startswith(b"q2\t")
    assert orc.maf2paf(txt).startswith(b"q1\t")
    with pytest.raises(orc.OracleError) as e:
        orc.maf2paf(txt, query_name="zz")
    assert e.value.code == 7
    assert len(orc.call_maf(txt, query_name="zz").splitlines()) == 8  # call skips instead


def test_query_regex_and_contig_lines():
    """caller.rs:91-104 (--query-regex: first s-line whose name matches -- the target line included -- else the block is
    skipped; ignored when --query-name is given) and caller.rs:339-357 (##contig lines of the <maf>.index: isref entries,
    natord order, behind the FORMAT lines).  Hand-derived from the source; the regex crate itself is played by Python's re."""
    txt = _maf([[b"s t.chr1 0 4 + 100 ACGT", b"s hg002#1 0 4 + 100 ACGA", b"s hg002#2 0 4 + 100 AC-T"],
                [b"s t.chr1 10 4 + 100 ACGT", b"s other 0 4 + 100 AGGT"],
                [b"s t.chr2 0 4 + 100 ACGT", b"s zz 0 4 + 100 ACGT", b"s hg002#2 9 4 + 100 TCGT"]])
    body = lambda v: [l for l in v.splitlines() if not l.startswith(b"#")]
    # regex picks hg002#1 in block 1, nothing in block 2 (skipped), hg002#2 in block 3
    got = body(orc.call_maf_ex(txt, snp=True, svlen=0, query_regex=rb"^hg002#.*"))
    assert [l.split(b"\t")[:2] for l in got] == [[b"t.chr1", b"4"], [b"t.chr2", b"1"]]
    assert got[0].endswith(b"1|1:hg002#1@3@P") and got[1].endswith(b"1|1:hg002#2@9@P")
    # an unanchored search (is_match_at(name, 0)): "#2" finds hg002#2 (a deletion at POS 2); "chr" matches the TARGET line itself
    assert body(orc.call_maf_ex(txt, snp=True, svlen=0, query_regex=rb"#2"))[0].startswith(b"t.chr1\t2\t.\tCG\tC\t")
    assert body(orc.call_maf_ex(txt, snp=True, svlen=0, query_regex=rb"chr")) == []  # target vs itself: no variant
    # --query-name wins over the regex
    assert orc.call_maf_ex(txt, snp=True, svlen=0, query_name="other", query_regex=rb"hg002") == \
        orc.call_maf(txt, snp=True, svlen=0, query_name="other")
    # contig lines: natord (chr2 < chr10), only isref entries, behind the FORMAT lines and in front of #CHROM
    idx = {"t.chr10": {"size": 30, "isref": True}, "t.chr2": {"size": 20, "isref": True}, "q.chr1": {"size": 5, "isref": False},
           "t.chr1": {"size": 100, "isref": True}}
    hdr = [l for l in orc.call_maf_ex(txt, mafindex=idx).splitlines() if l.startswith(b"#")]
    assert hdr[-4:] == [b"##contig=<ID=t.chr1,length=100>", b"##contig=<ID=t.chr2,length=20>", b"##contig=<ID=t.chr10,length=30>",
                        b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tsample"]
    assert hdr[-5].startswith(b"##FORMAT=<ID=GT")
    assert orc.call_maf_ex(txt, mafindex={}) == orc.call_maf(txt)  # an empty index adds nothing
    with pytest.raises(orc.OracleError):
        orc.call_maf_ex(txt, mafindex={"bad,name": {"size": 1, "isref": True}})


seqs = st.text(alphabet="ACGTacgtN-", min_size=1, max_size=80)


@settings(max_examples=300, deadline=None, derandomize=True)
@given(seqs, seqs, st.booleans())
def test_cigar_vs_pyspec(t, q, neg):
    n = min(len(t), len(q))
    c, cg = orc.maf_seq_to_cigar(t.encode()[:n], q.encode()[:n], neg)
    c2, cg2 = pyspec.maf_seq_to_cigar(t[:n], q[:n], neg)
    assert c == c2 and cg.decode() == cg2


def test_call_chunk_quirks():
    # (i) a gap run starting exactly at a chunk start is dropped; forced boundary inside '=' run
    t = b"ACGTACGTAC" + b"GGGGGGGG" + b"ACGT"
    q = b"ACGTACGTAC" + b"--------" + b"ACGT"
    txt = _maf([[b"s t 0 22 + 100 " + t, b"s q 0 14 + 100 " + q]])
    full = orc.call_maf(txt, svlen=3, chunk_size=1000).decode().splitlines()[8:]
    assert len(full) == 1 and "SVTYPE=DEL;SVLEN=8" in full[0]
    cut = orc.call_maf(txt, svlen=3, chunk_size=10).decode().splitlines()[8:]
    assert cut == []  # window [0,10) ends at the run start -> next chunk starts on the gap: dropped
    # (ii) one INV row per chunk for '-' strand with --inv
    txt = _maf([[b"s t 0 22 + 100 " + t, b"s q 0 14 - 100 " + q]])
    rows = orc.call_maf(txt, svlen=3, inv=True, chunk_size=1000).decode().splitlines()[8:]
    assert sum("<INV>" in r for r in rows) == len(pyspec.chunk_grid(t.decode(), q.decode(), 1000, 3))
    assert any("INV_NEST=TRUE;SVTYPE=DEL" in r and r.endswith("@N") for r in rows)


@settings(max_examples=200, deadline=None, derandomize=True)
@given(st.text(alphabet="AC-", min_size=1, max_size=60), st.integers(1, 12), st.integers(0, 4), st.data())
def test_chunk_grid_inv_rows(t, C, K, data):
    q = data.draw(st.text(alphabet="AC-", min_size=len(t), max_size=len(t)))
    nt = sum(ch != "-" for ch in t)
    nq = sum(ch != "-" for ch in q)
    txt = _maf([[f"s t 5 {nt} + 1000 {t}".encode(), f"s q 7 {nq} - 1000 {q}".encode()]])
    rows = orc.call_maf(txt, svlen=K, inv=True, chunk_size=C).decode().splitlines()[8:]
    grid = pyspec.chunk_grid(t, q, C, K)
    want = sum(1 for (s, e) in grid if any(ch != "-" for ch in t[s:e]))
    assert sum("<INV>" in r for r in rows) == want


def test_call_paf_hand_derived():
    """caller.rs:610-822 on a hand-walked record (inclusive faidx end, no revcomp, ungated <INV>, Q12/Q13/Q15)"""
    t_fa = b">chrT\nACGTACGTACGGGGCCCCAATTTT\n"
    q_fa = b">chrQ\nACGTACGTACGGCCCCAATTTTAA\n"
    ti, qi = orc.make_fai(t_fa), orc.make_fai(q_fa)

    def line(cg, strand=b"+"):
        return b"\t".join([b"chrQ", b"24", b"1", b"13", strand, b"chrT", b"24", b"2", b"14", b"9", b"14", b"60", b"cg:Z:" + cg]) + b"\n"
    hdr = orc.call_maf(b"##maf version=1\n", snp=True, svlen=0)  # header only
    got = orc.call_paf(line(b"4=1X2I3=2D2="), t_fa, ti, q_fa, qi, snp=True, svlen=0)
    assert got == hdr + (
        b"chrT\t7\t.\tG\tC\t.\t.\t.\tGT:QI\t1|1:chrQ@5@P\n"
        b"chrT\t7\t.\tG\tCGT\t.\t.\tSVTYPE=INS;SVLEN=2;END=7\tGT:QI\t1|1:chrQ@6@8@P\n"
        b"chrT\t10\t.\tCGG\tG\t.\t.\tSVTYPE=DEL;SVLEN=2;END=12\tGT:QI\t1|1:chrQ@11@11@P\n")
    got = orc.call_paf(line(b"4=1X2I3=2D2=", b"-"), t_fa, ti, q_fa, qi, snp=False, svlen=1)
    assert got == hdr + (
        b"chrT\t3\t.\tG\t<INV>\t.\t.\tSVTYPE=INV;END=14\tGT:QI\t1|1:chrQ@1@13@N\n"
        b"chrT\t7\t.\tG\tCGT\t.\t.\tINV_NEST=TRUE;SVTYPE=INS;SVLEN=2;END=7\tGT:QI\t1|1:chrQ@6@8@N\n"
        b"chrT\t10\t.\tCGG\tG\t.\t.\tINV_NEST=TRUE;SVTYPE=DEL;SVLEN=2;END=12\tGT:QI\t1|1:chrQ@11@11@N\n")
    # svlen cutoff is strict (>), leading indel has no anchor (after_m false), I right after D is skipped too
    assert orc.call_paf(line(b"4=1X2I3=2D2="), t_fa, ti, q_fa, qi, snp=False, svlen=2) == hdr
    assert orc.call_paf(line(b"2I4=2D3I2="), t_fa, ti, q_fa, qi, snp=False, svlen=0) == hdr + (
        b"chrT\t6\t.\tCGT\tG\t.\t.\tSVTYPE=DEL;SVLEN=2;END=8\tGT:QI\t1|1:chrQ@7@7@P\n")
    # a failing token silently ends the record (rows before it are kept); the next record is still processed
    two = line(b"4=1X3S2I3=") + line(b"4=1X")
    assert orc.call_paf(two, t_fa, ti, q_fa, qi, snp=True, svlen=0) == hdr + 2 * b"chrT\t7\t.\tG\tC\t.\t.\t.\tGT:QI\t1|1:chrQ@5@P\n"
    # slices past the fetched strings are panics; nothing is written
    for cg in (b"4=20D", b"12=3X", b"0=2I"):
        with pytest.raises(orc.OracleError) as e:
            orc.call_paf(line(b"4=1X") + line(cg), t_fa, ti, q_fa, qi, snp=True, svlen=0)
        assert e.value.code == 8


def test_paf_pseudo_hand_derived():
    """pseudomaf.rs:78-205 + cigar.rs:744-804 on hand-walked records: gap fill, head-overlap trim, contained record
    skipped, '-' strand reverse-complemented, I deleted / D gap-inserted, and the FASTA-less 1/0/- spelling"""
    fa = b">chrT\nACGTACGTACGGGGCCCCAATTTT\n>chrQ\nACGTACGTACGGCCCCAATTTTAA\n>q2\nTTTTGGGGCCCCAAAA\n"
    fai = orc.make_fai(fa)

    def line(cg, strand=b"+", q=b"chrQ", t=b"chrT", qs=1, qe=13, ts=2, te=14, qsz=24):
        return b"\t".join([q, b"%d" % qsz, b"%d" % qs, b"%d" % qe, strand, t, b"24", b"%d" % ts, b"%d" % te, b"9", b"14", b"60",
                           b"cg:Z:" + cg]) + b"\n"
    paf = (line(b"4=1X2I3=2D2=") + line(b"4=", q=b"q2", qs=0, qe=4, ts=16, te=20, qsz=16) +
           line(b"3=1D2=", b"-", q=b"q2", qs=4, qe=9, ts=4, te=10, qsz=16) +
           line(b"6=", qs=14, qe=20, ts=10, te=16) + line(b"4=", qs=3, qe=7, ts=11, te=15))
    got = orc.paf_pseudo(paf, fa, fai)
    assert list(got) == [b"chrT"]
    assert got[b"chrT"] == (b"a score=0\n"
                            b"s\tchrT\t0\t24\t+\t24\tACGTACGTACGGGGCCCCAATTTT\n"
                            b"s\tchrQ\t0\t24\t+\t24\t--CGTACACG--GCTT--------\n"
                            b"s\tq2\t0\t16\t+\t16\t----GCC-CC------TTTT----\n\n")
    assert orc.paf_pseudo(paf)[b"chrT"] == (b"a score=0\n"
                                            b"s\tchrT\t0\t24\t+\t24\tNNNNNNNNNNNNNNNNNNNNNNNN\n"
                                            b"s\tchrQ\t0\t24\t+\t24\t--11110111--1111--------\n"
                                            b"s\tq2\t0\t16\t+\t16\t----111-11------1111----\n\n")
    # --target filter, several targets in first-occurrence order, ops outside M=XIDS are ignored (Q4)
    paf2 = line(b"2H4=2N", t=b"other", ts=0, te=4, qs=0, qe=4) + paf
    assert list(orc.paf_pseudo(paf2)) == [b"other", b"chrT"]
    assert orc.paf_pseudo(paf2)[b"other"] == b"a score=0\ns\tother\t0\t24\t+\t24\t" + b"N" * 24 + b"\ns\tchrQ\t0\t24\t+\t24\t1111" + b"-" * 20 + b"\n\n"
    assert list(orc.paf_pseudo(paf2, target="chrT")) == [b"chrT"]
    assert orc.paf_pseudo(paf2, target="nope") == {}
    for bad, code in ((line(b"4=9I"), 8), (line(b"20=3D"), 8), (line(b"5MM"), 5), (line(b"M"), 1), (line(b""), 8),
                      (line(b"4=", q=b"zz"), 11), (line(b"12=") + line(b"2=", ts=4, te=20, qs=0, qe=2), 8)):
        with pytest.raises(orc.OracleError) as e:
            orc.paf_pseudo(bad, fa, fai)
        assert e.value.code == code, (bad, e.value.code, e.value.msg)


def test_validate_golden_and_hand_derived():
    """validate.rs:73-125 on the reference's own PAF (consistent: nothing to fix, -f output == input) and on records
    whose ends were falsified by hand"""
    rep, fx = orc.paf_validate(TEST_PAF, True)
    assert rep == (b"Total records: 2\nQuery invalid records: 0\nTarget invalid records: 0\nQuery invalid list:\n"
                   b"Target invalid list:\n\n")
    assert fx == TEST_PAF
    # 25M10I15M20D30M20I30M10D70M: query 0 + 170 + 30 = 200, target 0 + 170 + 30 = 200
    bad = TEST_PAF.replace(b"A\t300\t0\t200\t+\tB\t300\t0\t200", b"A\t300\t0\t190\t+\tB\t300\t0\t0x7", 1)
    rep, fx = orc.paf_validate(bad, True)
    assert rep == (b"Total records: 2\nQuery invalid records: 1\nTarget invalid records: 1\nQuery invalid list:\nA:0-190\n"
                   b"Target invalid list:\nB:0-7\n\n")
    assert fx == TEST_PAF                      # both ends repaired; the hex target_end comes back in decimal
    with pytest.raises(orc.OracleError) as e:  # get_stat().unwrap()
        orc.paf_validate(TEST_PAF.replace(b"25M10I", b"25M10S", 1))
    assert e.value.code == 8


def test_chain_hand_derived():
    """converter.rs:57-92,148-173 + chain.rs:92-204 + cigar.rs:155-295,435-490 on hand-walked records"""
    # the reference's own PAF: 25M10I15M20D30M20I30M10D70M and, on '-', 10M10D10M10I20M
    assert orc.paf2chain(TEST_PAF) == (b"chain\t255\tB\t300\t+\t0\t200\tA\t300\t+\t0\t200\t0\n25\t0\t10\n15\t20\t0\n30\t0\t20\n30\t10\t0\n70\n\n"
                                       b"chain\t255\tB\t300\t+\t200\t250\tA\t300\t-\t50\t250\t1\n10\t10\t0\n10\t0\t10\n20\n\n")
    # head / tail indels are trimmed off the coordinates; the tail keeps only the LAST I and the LAST D run; on '-' the
    # new query end is computed from the new query start (chain.rs:133-138)
    line = lambda cg, strand: b"\t".join([b"q", b"1000", b"100", b"200", strand, b"t", b"2000", b"300", b"420", b"1", b"2", b"60",
                                         b"cg:Z:" + cg]) + b"\n"
    assert orc.paf2chain(line(b"3I2D5I40=2X10I8=4D6=7I2D3I", b"+")) == (
        b"chain\t255\tt\t2000\t+\t302\t418\tq\t1000\t+\t108\t197\t0\n42\t0\t10\n8\t4\t0\n6\n\n")
    assert orc.paf2chain(line(b"3I2D5I40=2X10I8=4D6=7I2D3I", b"-")) == (
        b"chain\t255\tt\t2000\t+\t302\t418\tq\t1000\t-\t808\t189\t0\n42\t0\t10\n8\t4\t0\n6\n\n")
    assert orc.paf2chain(line(b"5I3D", b"+")) == b"chain\t255\tt\t2000\t+\t303\t417\tq\t1000\t+\t105\t195\t0\n0\n\n"
    for bad, code in ((line(b"5M2S", b"+"), 5), (line(b"5MM", b"+"), 5), (line(b"M", b"+"), 1), (line(b"", b"+"), 8)):
        with pytest.raises(orc.OracleError) as e:
            orc.paf2chain(bad)
        assert e.value.code == code
    # MAF: the same chain as converting to PAF first (same coordinates, '='/'X' merge into one ungapped block)
    maf = _maf([[b"s t 10 9 + 100 --ACGTAC-GTT", b"s q 5 8 - 50 GGAC-TACAGT-"],
                [b"s t 0 4 + 100 ACGT", b"s q 0 4 + 50 ACGA"]])
    assert orc.maf2chain(maf) == orc.paf2chain(orc.maf2paf(maf))
    assert orc.maf2chain(maf) == (b"chain\t255\tt\t100\t+\t10\t18\tq\t50\t-\t7\t43\t0\n2\t1\t0\n3\t0\t1\n2\n\n"
                                  b"chain\t255\tt\t100\t+\t0\t4\tq\t50\t+\t0\t4\t1\n4\n\n")
    assert orc.maf2chain(TEST_MAF) == orc.paf2chain(orc.maf2paf(TEST_MAF))


def test_pafcov_hand_derived():
    """pafcov.rs:13-60 + cigar.rs:710-741: M/= cover, I/S stay, everything else advances; clipped to the first record's
    target_length; one BED line per base"""
    line = lambda cg, t=b"t", tl=20, ts=2: b"\t".join([b"q", b"100", b"0", b"10", b"+", t, b"%d" % tl, b"%d" % ts, b"12", b"1", b"2", b"60",
                                                     b"cg:Z:" + cg]) + b"\n"
    got = orc.pafcov(line(b"3=2X2I4M3D5=") + line(b"30M", ts=15, tl=999) + line(b"2=1N2H1=", t=b"u", tl=7, ts=0))
    cov_t = [0, 0, 1, 1, 1, 0, 0, 1, 1, 1, 1, 0, 0, 0, 1, 2, 2, 2, 2, 1]
    cov_u = [1, 1, 0, 0, 0, 1, 0]
    want = b"".join(b"t\t%d\t%d\t%d\n" % (i, i + 1, c) for i, c in enumerate(cov_t)) + \
        b"".join(b"u\t%d\t%d\t%d\n" % (i, i + 1, c) for i, c in enumerate(cov_u))
    assert got == want
    assert orc.pafcov(b"") == b""
    for bad, code in ((line(b"5MM"), 5), (line(b"M"), 1), (line(b""), 8)):
        with pytest.raises(orc.OracleError) as e:
            orc.pafcov(bad)
        assert e.value.code == code


_ops = st.lists(st.tuples(st.integers(0, 40), st.sampled_from("M=XID")), min_size=1, max_size=14)


def _paf_line(q, qsize, qs, qe, strand, t, tsize, ts, te, cg):
    return ("\t".join(map(str, [q, qsize, qs, qe, strand, t, tsize, ts, te, 1, 2, 60, "cg:Z:" + cg])) + "\n").encode()


@settings(max_examples=300, deadline=None, derandomize=True)
@given(_ops, st.booleans(), st.integers(0, 50), st.integers(0, 50))
def test_paf_chain_and_counts_vs_pyspec(ops, neg, qs, ts):
    cg = "".join(f"{n}{o}" for n, o in ops)
    strand = "-" if neg else "+"
    c = pyspec.paf_counts(cg, neg)
    assert orc.paf_cigar_counts(cg.encode(), neg) == c
    qe = qs + c[0] + c[1] + c[3] + c[7]
    te = ts + c[0] + c[1] + c[5] + c[9]
    rec = ("q", 5000, qs, qe, strand, "t", 6000, ts, te, cg)
    line = _paf_line(*rec)
    assert orc.paf2chain(line + line).decode() == pyspec.paf_chain(rec, 0) + pyspec.paf_chain(rec, 1)
    rep, fx = orc.paf_validate(line, True)        # consistent by construction: nothing to fix
    assert rep.startswith(b"Total records: 1\nQuery invalid records: 0\nTarget invalid records: 0\n") and fx == line
    rep, _ = orc.paf_validate(_paf_line("q", 5000, qs, qe + 1, strand, "t", 6000, ts, te + 2, cg), False)
    assert rep == (f"Total records: 1\nQuery invalid records: 1\nTarget invalid records: 1\nQuery invalid list:\nq:{qs}-{qe + 1}\n"
                   f"Target invalid list:\nt:{ts}-{te + 2}\n\n").encode()


def test_chain_reader_chain2paf_chain2maf():
    """src/parser/chain.rs:206-383 + converter.rs:268-416 + cigar.rs:554-626.  Pins: the README's round trip
    (README.md:225 `paf2chain | chain2maf | maf2chain | chain2paf`) on the reference's own PAF, the nom grammar's quirks
    one by one, and the independent Python spelling in tests/pyspec.py."""
    ch = orc.paf2chain(TEST_PAF)
    # chain2paf of the reference's PAF through paf2chain: '='/'X'/'M' merged, I = target_diff, D = query_diff
    assert orc.chain2paf(ch) == (b"A\t300\t0\t200\t+\tB\t300\t0\t200\t170\t200\t255\tcg:Z:25M10I15M20D30M20I30M10D70M\n"
                                 b"A\t300\t50\t250\t-\tB\t300\t200\t250\t40\t50\t255\tcg:Z:10M10D10M10I20M\n")
    assert orc.chain2paf(ch).decode() == pyspec.chain2paf(ch.decode())
    # a chain round trip is the identity on chains: paf2chain(chain2paf(x)) == x (ids restart at 0 either way)
    assert orc.paf2chain(orc.chain2paf(ch)) == ch
    # quirks of the grammar (each line of the docstring in oracle/wga_oracle.c)
    two = (b"chain 255 t 100 + 0 10 q 90 - 5 15 0\n4\t1\t2\n3\n\n# a note without the letter\n\n"
           b"chain 255 t 100 + 20 30 q 90 + 1 2 1\n7 2\n5")
    assert orc.chain2paf(two) == (b"q\t90\t5\t15\t-\tt\t100\t0\t10\t7\t8\t255\tcg:Z:4M2I1D3M\n"          # size qdiff tdiff: I before D
                                  b"q\t90\t1\t2\t+\tt\t100\t20\t30\t7\t9\t255\tcg:Z:7M2D\n")          # last line without '\n': dropped
    assert orc.chain2paf(two).decode() == pyspec.chain2paf(two.decode())
    assert orc.chain2paf(b"") == b""
    assert orc.chain2paf(b"chain255 t 9 + 0 1 q 9 + 0 1 7 extra fields\r\n1 0 0 9\r\n") == b"q\t9\t0\t1\t+\tt\t9\t0\t1\t1\t1\t255\tcg:Z:1M\n"
    H = b"chain 1 t 9 + 0 1 q 9 + 0 1 0\n"
    for bad, code in ((b"xchain 1 t 9 + 0 1 q 9 + 0 1 0\n1\n", 11), (b"xchain\n", 8),      # tag("chain"); short rest -> input[..10] panics
                      (H + b"\n", 8), (H + b"\n" + b"#" * 12, 11),                            # fold_many1 needs one data line
                      (H[:-1], 8),                                                             # header that ends the file
                      (b"chain a t 9 + 0 1 q 9 + 0 1 0\n1\n", 13), (b"chain 1 t x + 0 1 q 9 + 0 1 0\n1\n", 1),
                      (b"chain 1 t 9 * 0 1 q 9 + 0 1 0\n1\n", 2), (b"chain 1 t 9 + 0 1 q 9 + 0 1\n1\n", 12),
                      (H + b"1\n2 x\n", 1), (H + b"1\n \n", 12),                            # data line errors
                      (H + b"1\n\n# a comment\n" + H + b"1\n", 11),                         # 'c' in the skipped part: tag fails there
                      (H + b"1\n12 c\n", 8)):
        with pytest.raises(orc.OracleError) as e:
            orc.chain2paf(bad)
        assert e.value.code == code, (bad, e.value)
        with pytest.raises(ValueError):
            pyspec.chain2paf(bad.decode())
    assert orc.chain2paf(H + b"1\n2 3\nh\n9 9 9\n" + H + b"5\n") == (b"q\t9\t0\t1\t+\tt\t9\t0\t1\t3\t6\t255\tcg:Z:1M2M3D\n"
                                                                  b"q\t9\t0\t1\t+\tt\t9\t0\t1\t5\t5\t255\tcg:Z:5M\n")   # lines after 'h' are skipped
    # chain2maf: rows through the FASTA, '-' strand reverse-complemented, gaps by the data lines
    fa = b">t\nACGTACGTAC\nGTACGTACGT\n>q\nTTGCATGCAT\nGCATGCAAGG\n"
    fai = orc.make_fai(fa)
    seqs = {"t": "ACGTACGTACGTACGTACGT", "q": "TTGCATGCATGCATGCAAGG"}
    c2 = b"chain 9 t 20 + 2 12 q 20 - 3 12 5\n3 1 2\n4\n\nchain 9 t 20 - 0 5 q 20 + 10 15 6\n5\n"
    got = orc.chain2maf(c2, fa, fai, fa, fai)
    assert got.decode() == pyspec.chain2maf(c2.decode(), seqs)
    assert got == (b"#maf version=1.6 convert_from=chain t_seq_path=t.fa q_seq_path=q.fa\n"
                   b"a score=255\ns\tt\t2\t10\t+\t20\tGTA--CGTACGT\ns\tq\t8\t9\t-\t20\tGCATG-CATG\n\n"
                   b"a score=255\ns\tt\t0\t5\t-\t20\tACGTA\ns\tq\t10\t5\t+\t20\tGCATG\n\n")
    for bad, code in ((b"chain 9 t 20 + 2 4 q 20 + 3 5 0\n9 0 1\n", 8),            # insert_str past the end of the target row
                      (b"chain 9 zz 20 + 2 4 q 20 + 3 5 0\n2\n", 11),              # sequence not in the .fai
                      (b"chain 9 t 20 + 2 4 k 20 - 0 5 0\n2\n", 10)):              # revcomp of a non-ACGTN base
        with pytest.raises(orc.OracleError) as e:
            orc.chain2maf(bad, fa + b">k\nACGTR\n", orc.make_fai(fa + b">k\nACGTR\n"), fa + b">k\nACGTR\n", orc.make_fai(fa + b">k\nACGTR\n"))
        assert e.value.code == code, (bad, e.value)
    # README.md:225 round trip on the reference's PAF with a FASTA that fits its coordinates: chain -> maf -> chain -> paf
    rng = np.random.default_rng(5)
    fa2 = b"".join(b">" + n + b"\n" + bytes(rng.choice(list(b"ACGT"), 300).astype(np.uint8)) + b"\n" for n in (b"A", b"B"))
    fai2 = orc.make_fai(fa2)
    maf = orc.chain2maf(ch, fa2, fai2, fa2, fai2)
    assert orc.maf2chain(maf) == ch                      # gapped rows carry exactly the chain's blocks
    assert orc.chain2paf(orc.maf2chain(maf)) == orc.chain2paf(ch)


@settings(max_examples=150, deadline=None, derandomize=True)
@given(st.lists(st.tuples(st.booleans(), st.lists(st.tuples(st.integers(0, 30), st.integers(0, 9), st.integers(0, 9)), min_size=1, max_size=6),
                          st.sampled_from(["\n", "\n\n", "\n# zzz\n", "\r\n\n"])), min_size=1, max_size=5), st.booleans())
def test_chain2paf_vs_pyspec(recs, final_newline):
    text = ""
    for i, (neg, lines, sep) in enumerate(recs):
        text += f"chain 4{i}.5 t{i} 1000 + 10 {10 + i} q 900 {'-' if neg else '+'} 5 50 {i}\n"
        text += "\n".join("\t".join(str(x) for x in (l if j + 1 < len(lines) else l[:1 + (l[1] + l[2] > 0) * 2])) for j, l in enumerate(lines))
        text += sep
    if not final_newline:
        text = text.rstrip("\n")
    try:
        want = pyspec.chain2paf(text)
    except ValueError:
        with pytest.raises(orc.OracleError):
            orc.chain2paf(text.encode())
        return
    assert orc.chain2paf(text.encode()).decode() == want


@settings(max_examples=200, deadline=None, derandomize=True)
@given(st.lists(st.tuples(st.sampled_from(["a", "b"]), st.integers(0, 30),
                          st.lists(st.tuples(st.integers(0, 12), st.sampled_from("M=XIDSNH")), min_size=1, max_size=8)),
                min_size=1, max_size=6), st.integers(1, 40))
def test_pafcov_vs_pyspec(recs, tlen):
    lines, spec = [], []
    for t, ts, ops in recs:
        cg = "".join(f"{n}{o}" for n, o in ops)
        lines.append(_paf_line("q", 100, 0, 10, "+", t, tlen if t == "a" else tlen + 3, ts, ts + 1, cg))
        spec.append((t, tlen if t == "a" else tlen + 3, ts, cg))
    assert orc.pafcov(b"".join(lines)).decode() == pyspec.paf_cov(spec)


@settings(max_examples=200, deadline=None, derandomize=True)
@given(st.data())
def test_paf_pseudo_vs_pyspec(data):
    """random consistent records (CIGAR spans == coordinate spans) over a tiny FASTA, one target, two queries"""
    rnd = data.draw(st.randoms(use_true_random=False))
    bases = "ACGTacgtN"
    tlen = rnd.randrange(20, 60)
    fasta = {"t": "".join(rnd.choice(bases) for _ in range(tlen)), "qa": "".join(rnd.choice(bases) for _ in range(80)),
             "qb": "".join(rnd.choice(bases) for _ in range(70))}
    recs, used = [], set()
    for _ in range(rnd.randrange(1, 6)):
        q = rnd.choice(["qa", "qb"])
        ts = rnd.randrange(0, tlen - 5)
        if (q, ts) in used:
            continue
        used.add((q, ts))
        ops, tcons, qcons = [], 0, 0
        for _ in range(rnd.randrange(1, 6)):
            op = rnd.choice("M=XIDS" if rnd.random() < 0.9 else "NH")
            n = rnd.randrange(1, 6)
            if op in "M=XD" and ts + tcons + n > tlen:
                break
            ops.append(f"{n}{op}")
            tcons += n if op in "M=XD" else 0
            qcons += n if op in "M=XIS" else 0
        if not ops or tcons == 0:
            continue
        qs = rnd.randrange(0, len(fasta[q]) - qcons) if len(fasta[q]) > qcons else 0
        recs.append((q, len(fasta[q]), qs, qs + max(qcons, 1), rnd.choice("+-"), "t", tlen, ts, ts + tcons, "".join(ops)))
    if not recs:
        return
    paf = b"".join(_paf_line(*r) for r in recs)
    fa = "".join(f">{k}\n{v}\n" for k, v in fasta.items()).encode()
    fai = orc.make_fai(fa)
    for use_fa in (True, False):
        try:
            want = pyspec.pseudo_maf(recs, fasta if use_fa else None)
        except AssertionError:
            with pytest.raises(orc.OracleError) as e:
                orc.paf_pseudo(paf, fa if use_fa else None, fai if use_fa else None)
            assert e.value.code == 8
            continue
        got = orc.paf_pseudo(paf, fa if use_fa else None, fai if use_fa else None)
        assert list(got) == [b"t"] and got[b"t"].decode() == want


@settings(max_examples=300, deadline=None, derandomize=True)
@given(st.data())
def test_call_paf_vs_pyspec(data):
    rnd = data.draw(st.randoms(use_true_random=False))
    bases = "ACGTacgtN" if rnd.random() < 0.8 else "ACGTRY"
    ft = {"t": "".join(rnd.choice(bases) for _ in range(60))}
    fq = {"q": "".join(rnd.choice(bases) for _ in range(60))}
    recs = []
    for _ in range(rnd.randrange(1, 4)):
        ops = "".join(f"{rnd.randrange(0, 7)}{rnd.choice('M=XXIDDI' if rnd.random() < 0.95 else 'SN')}" for _ in range(rnd.randrange(1, 7)))
        ts, qs = rnd.randrange(0, 30), rnd.randrange(0, 30)
        recs.append(("q", 60, qs, qs + rnd.randrange(0, 30), rnd.choice("+-"), "t", 60, ts, ts + rnd.randrange(0, 30), ops))
    paf = b"".join(_paf_line(*r) for r in recs)
    tfa, qfa = f">t\n{ft['t']}\n".encode(), f">q\n{fq['q']}\n".encode()
    snp, svlen = rnd.random() < 0.7, rnd.choice([0, 0, 1, 3])
    hdr = orc.call_maf(b"##maf version=1\n")
    try:
        want = hdr.decode() + "".join("".join(pyspec.call_paf_rows(r, ft, fq, snp, svlen)) for r in recs)
        err = None
    except AssertionError:
        err = 8
    except ValueError:
        err = 10
    if err is None:
        assert orc.call_paf(paf, tfa, orc.make_fai(tfa), qfa, orc.make_fai(qfa), snp=snp, svlen=svlen).decode() == want
    else:
        with pytest.raises(orc.OracleError) as e:
            orc.call_paf(paf, tfa, orc.make_fai(tfa), qfa, orc.make_fai(qfa), snp=snp, svlen=svlen)
        assert e.value.code == err


# ---------------------------------------------------------------- cs:Z: -> CIGAR (row a7, paf.rs:159-218)
def test_cs_to_cigar_reference_doc_example():
    """the worked example of the reference's own doc comment (paf.rs:156-158) -- a reference-held vector"""
    assert orc.cs_to_cigar(b":6-ata:10+gtc:4*at*tg:3") == b"6M3D10M3I4M2X3M"
    assert pyspec.cs_to_cigar(":6-ata:10+gtc:4*at*tg:3") == "6M3D10M3I4M2X3M"
    # quirks: `=ACGT` is matched but ignored, garbage is skipped, upper-case `*AT` is no token, `:` needs digits
    assert orc.cs_to_cigar(b"=ACGT:3*AT*at~~:*ga-N+nn:0:0") == b"3M2X1D2I"
    assert orc.cs_to_cigar(b"") == b"" and orc.cs_to_cigar(b"=ACGT") == b""
    with pytest.raises(orc.OracleError) as e:
        orc.cs_to_cigar(b":18446744073709551616")
    assert e.value.code == 8
    assert orc.cs_to_cigar(b":18446744073709551615") == b"18446744073709551615M"
    line = b"q\t100\t0\t12\t+\tt\t100\t0\t13\t1\t2\t60\ttp:A:P\tcs:Z:junk\tcs:Z::6-a*ac+tt:3\n"
    with pytest.raises(orc.OracleError) as e:   # the FIRST cs tag is taken: "junk" -> "" -> the empty-CIGAR panic
        orc.stat_paf(line)
    assert e.value.code == 8
    line2 = line.replace(b"cs:Z:junk\t", b"")
    assert orc.stat_paf(line2) == orc.stat_paf(line2.replace(b"cs:Z::6-a*ac+tt:3", b"cg:Z:6M1D1X2I3M"))
    assert orc.stat_paf(line2.replace(b"tp:A:P", b"cg:Z:9=")) == orc.stat_paf(line2.replace(b"cs:Z::6-a*ac+tt:3", b"cg:Z:9="))


@settings(max_examples=500, deadline=None, derandomize=True)
@given(st.text(alphabet=":*=+-0123459acgtnACGTZ~ é", max_size=60))
def test_cs_to_cigar_vs_pyspec(cs):
    want = pyspec.cs_to_cigar(cs)
    if want is None:
        with pytest.raises(orc.OracleError):
            orc.cs_to_cigar(cs.encode())
    else:
        assert orc.cs_to_cigar(cs.encode()).decode() == want


# ---------------------------------------------------------------- filter -f paf (8(f) item 3, filter.rs:41-63,92-160)
def test_paf_filter_hand_derived():
    paf = open(os.path.join(G, "ref_testdotplot.paf"), "rb").read()
    rows = paf.splitlines(keepends=True)
    assert orc.paf_filter(paf) == paf                                   # nothing to drop, csv round trip is the identity here
    # records of test/testdotplot.paf: target spans and query lengths by hand
    spans = [int(r.split(b"\t")[8]) - int(r.split(b"\t")[7]) for r in rows]
    qlens = [int(r.split(b"\t")[1]) for r in rows]
    for b, q in ((0, 0), (max(spans), 0), (max(spans) + 1, 0), (0, max(qlens) + 1), (min(spans) + 1, min(qlens))):
        want = b"".join(r for r, s_, l in zip(rows, spans, qlens) if not (s_ < b or l < q))
        assert orc.paf_filter(paf, b, q) == want
    # -a: pair sums over the whole file, -b/-q ignored
    def line(q, t, ts, te, extra=b""):
        return b"\t".join([q, b"100", b"0", b"10", b"+", t, b"200", b"%d" % ts, b"%d" % te, b"5", b"10", b"60"]) + extra + b"\n"
    f = line(b"a", b"x", 0, 10) + line(b"b", b"x", 0, 30) + b"#c\n" + line(b"a", b"x", 50, 65, b"\tcg:Z:15=\tzz:Z:say \"hi\"") + \
        line(b"a", b"y", 0, 24) + line(b"b", b"x", 5, 4)                # the last one wraps: 2^64 - 1 added to (b, x)
    got = orc.paf_filter(f, min_align=25)
    assert got == line(b"a", b"x", 0, 10) + line(b"b", b"x", 0, 30) + line(b"a", b"x", 50, 65, b"\tcg:Z:15=\t\"zz:Z:say \"\"hi\"\"\"") + line(b"b", b"x", 5, 4)
    assert orc.paf_filter(f, min_align=26) == line(b"b", b"x", 0, 30) + line(b"b", b"x", 5, 4)
    assert orc.paf_filter(f, 1000, 1000, min_align=0) == orc.paf_filter(f)
    with pytest.raises(orc.OracleError) as e:
        orc.paf_filter(f + b"q\t1\n", min_align=0)
    assert e.value.code == 9


# ---------------------------------------------------------------- dotplot base-level plot data (8(f) item 4, cigar.rs:807-985)
def _csv_to_values(csv_text):
    rows = csv_text.decode().splitlines()
    assert rows[0] == "ref_start,ref_end,query_start,query_end,cigar,ref_chro,query_chro"
    out = []
    for r in rows[1:]:
        rs, re_, qs, qe, c, tn, qn = r.split(",")
        out.append({"cigar": c, "query_chro": qn, "query_end": int(qe), "query_start": int(qs), "ref_chro": tn, "ref_end": int(re_),
                    "ref_start": int(rs)})
    return out


def test_plotdata_reference_html_golden():
    """test/test.html is the reference's own `dotplot -f paf` output: its vega `data.values` are the BasePlotdata of record 1
    of test/testdotplot.paf (9 segments M I M D M I M D M) -- a true reference output for the coordinate walk"""
    import json
    html = open(os.path.join(G, "ref_test.html")).read()
    i = html.index('"values":') + len('"values":')
    values, _ = json.JSONDecoder().raw_decode(html[i:])
    assert [v["cigar"] for v in values] == list("MIMDMIMDM")
    paf = open(os.path.join(G, "ref_testdotplot.paf"), "rb").read()
    got = _csv_to_values(orc.paf_plotdata(paf, 0))
    assert got[:9] == values
    # record 2 is on the '-' strand: query_start / query_end change places (cigar.rs:807-811), offsets still count up
    assert got[9:] == [
        {"cigar": "M", "query_chro": "A", "query_end": 200, "query_start": 210, "ref_chro": "B", "ref_end": 210, "ref_start": 200},
        {"cigar": "D", "query_chro": "A", "query_end": 210, "query_start": 210, "ref_chro": "B", "ref_end": 220, "ref_start": 210},
        {"cigar": "M", "query_chro": "A", "query_end": 210, "query_start": 220, "ref_chro": "B", "ref_end": 230, "ref_start": 220},
        {"cigar": "I", "query_chro": "A", "query_end": 220, "query_start": 230, "ref_chro": "B", "ref_end": 230, "ref_start": 230},
        {"cigar": "M", "query_chro": "A", "query_end": 230, "query_start": 250, "ref_chro": "B", "ref_end": 250, "ref_start": 230}]
    # -l 10: nothing is longer than the cut-off except the 20D / 20I of record 1; short indels stretch the open M segment
    got = _csv_to_values(orc.paf_plotdata(paf, 10))
    assert [(v["cigar"], v["ref_start"], v["ref_end"], v["query_start"], v["query_end"]) for v in got] == [
        ("M", 0, 40, 0, 50), ("D", 40, 60, 50, 50), ("M", 60, 90, 50, 80), ("I", 90, 90, 80, 100), ("M", 90, 200, 100, 200),
        ("M", 200, 250, 250, 200)]
    # the MAF route over the same alignment gives the same segments (test.maf block 1 is another alignment: just shape-check it)
    maf = open(os.path.join(G, "ref_test.maf"), "rb").read()
    rows = orc.maf_plotdata(maf, 0).decode().splitlines()
    assert rows[0].startswith("ref_start,") and len(rows) > 2
    assert orc.maf_plotdata(b"##maf version=1\n", 0) == b"" and orc.paf_plotdata(b"", 0) == b""


_pops = st.lists(st.tuples(st.integers(0, 30), st.sampled_from("M=XIDNSH")), min_size=1, max_size=16)


@settings(max_examples=400, deadline=None, derandomize=True)
@given(st.lists(st.tuples(_pops, st.booleans(), st.integers(0, 40), st.integers(0, 40)), min_size=1, max_size=4), st.integers(0, 12))
def test_plotdata_vs_pyspec(recs, cutoff):
    lines, spec = [], []
    for k, (ops, neg, ts, qs) in enumerate(recs):
        cg = "".join(f"{n}{o}" for n, o in ops)
        tn, qn = f"t{k},x" if k == 1 else f"t{k}", f'q"{k}' if k == 2 else f"q{k}"
        lines.append(_paf_line(qn, 5000, qs, qs + 1, "-" if neg else "+", tn, 6000, ts, ts + 1, cg))
        tq = '"' + tn + '"' if "," in tn else tn
        qq = '"' + qn.replace('"', '""') + '"' if '"' in qn else qn
        spec.append((tq, qq, ts, qs, neg, ops))
    assert orc.paf_plotdata(b"".join(lines), cutoff).decode() == pyspec.plot_csv(spec, cutoff)


def test_validate_report_labels_match_readme():
    """README.md:366-372 shows the report of `wgatools validate wrong.paf`: the five labels and their order are the
    reference's own text (the numbers there belong to a file the reference does not ship)"""
    bad = b"q\t100\t0\t9\t+\tt\t100\t0\t12\t5\t10\t60\tcg:Z:10=\n"      # query end 9 != 0 + 10, target end 12 != 0 + 10
    rep, _ = orc.paf_validate(bad, False)
    lines = rep.decode().split("\n")
    assert lines[0] == "Total records: 1" and lines[1] == "Query invalid records: 1" and lines[2] == "Target invalid records: 1"
    assert lines[3] == "Query invalid list:" and lines[4] == "q:0-9" and lines[5] == "Target invalid list:" and lines[6] == "t:0-12"
    for label in ("Total records: ", "Query invalid records: ", "Target invalid records: ", "Query invalid list:", "Target invalid list:"):
        assert label.encode() in rep
