"""No-GPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/wga_b200.h declares (no compute calls here -- there is no GPU on this box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from wgatools_b200 import _lib
    return _lib


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "wga_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(wga_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    L = ctypes.CDLL(built.SO_PATH)
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/wga_b200.h but not exported"
    assert names == set(built.ABI), names ^ set(built.ABI)


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import wgatools_b200 as w
    with pytest.raises(w.WGAError) as e:
        w.Context()
    assert e.value.code == 100
    with pytest.raises(w.WGAError):
        w.run_bytes(w.maf2paf, w.MAFReader(b"#h\ns t 0 1 + 9 A\ns q 0 1 + 9 A\n"))


def test_product_does_not_reference_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "wgatools_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "hd_core.h" and False, f"{f} mentions the oracle"


def test_safe_cut(built):
    L = built.load()
    txt = b"#h\na\ns t\ns q\n\na\ns t\ns q\n"
    # last non-'s' line start: the 'a' at offset 15? compute in python
    want = max(i for i in range(1, len(txt)) if txt[i - 1:i] == b"\n" and txt[i:i + 1] != b"s")
    assert L.wga_maf_safe_cut(txt, len(txt)) == want
    assert L.wga_maf_safe_cut(b"s a\ns b\n", 8) == 0


def test_vcf_header_contigs_host_only(built):
    """wga_vcf_header_contigs is host-side text (ctx may be NULL): natord order, isref entries as given, name grammar --
    against the oracle's restatement of caller.rs:339-357"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc
    L = built.load()
    idx = {"t.chr10": 30, "t.chr2": 20, "t.chr1": 100, "scaffold_7": 9, "Chr01": 4, "chrX": 7, "chr 9": 1}
    idx.pop("chr 9")
    names = (ctypes.c_char_p * len(idx))(*[k.encode() for k in idx])
    sizes = (ctypes.c_uint64 * len(idx))(*idx.values())
    sp = built.Span()
    assert L.wga_vcf_header_contigs(None, b"smp", names, sizes, len(idx), ctypes.byref(sp)) == 0
    got = ctypes.string_at(sp.data, sp.len)
    txt = b"##maf version=1\na\ns t 0 1 + 9 A\ns q 0 1 + 9 A\n"
    want = orc.call_maf_ex(txt, sample="smp", mafindex={k: {"size": v, "isref": True} for k, v in idx.items()})
    assert got == want
    assert L.wga_vcf_header_contigs(None, None, None, None, 0, ctypes.byref(sp)) == 0
    assert ctypes.string_at(sp.data, sp.len) == orc.call_maf(txt)
    for bad in (b"a,b", b"*x", b"=x", b"a b", b"", b"x<y"):
        nm = (ctypes.c_char_p * 1)(bad)
        assert L.wga_vcf_header_contigs(None, b"s", nm, sizes, 1, ctypes.byref(sp)) == 101
