"""CPU checks of the device building blocks (hd_core.h) and host formatters (host_fmt.hpp), compiled with
g++ on the no-GPU box: the byte-SIMD classifier against the scalar classifiers of the reference
(cigar.rs:298-328), number formatting/parsing, and the product's f32 / natord restatements against the
oracle's independent ones."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

import oracle_lib as orc
import pyspec

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhd_harness.so")


@pytest.fixture(scope="module")
def hd():
    src = os.path.join(HERE, "hd_harness.cc")
    deps = [src, os.path.join(HERE, "..", "wgatools_b200", "csrc", "hd_core.h"),
            os.path.join(HERE, "..", "wgatools_b200", "csrc", "host_fmt.hpp")]
    if not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, src])
    L = ctypes.CDLL(SO)
    L.hd_classify4.argtypes = [ctypes.c_uint32] * 3 + [ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
    L.hd_classify_prev.argtypes = [ctypes.c_uint32] * 2 + [ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
    L.hd_zero_bytes.restype = L.hd_ws_bytes.restype = L.hd_prev_flags.restype = L.hd_funnel.restype = ctypes.c_uint32
    L.hd_zero_bytes.argtypes = L.hd_ws_bytes.argtypes = [ctypes.c_uint32]
    L.hd_prev_flags.argtypes = [ctypes.c_uint32] * 2
    L.hd_funnel.argtypes = [ctypes.c_uint32] * 3
    L.hd_ndigits.argtypes = [ctypes.c_uint64]
    L.hd_write_u64.argtypes = [ctypes.c_char_p, ctypes.c_uint64]
    L.hd_parse_u64.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
    L.hd_fmt_f32.argtypes = [ctypes.c_float, ctypes.c_char_p]
    L.hd_natord.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    return L


def _word(bs):
    return int.from_bytes(bytes(bs), "little")


def test_zero_and_ws_bytes(hd):
    rng = random.Random(0)
    for _ in range(5000):
        bs = [rng.choice([0, 9, 10, 13, 14, 8, 32, 33, 0x80 | 9, 0xA0, 65, 45, 255]) for _ in range(4)]
        w = _word(bs)
        z = hd.hd_zero_bytes(w)
        ws = hd.hd_ws_bytes(w)
        for i, b in enumerate(bs):
            assert ((z >> (8 * i + 7)) & 1) == (b == 0)
            assert ((ws >> (8 * i + 7)) & 1) == (b == 32 or 9 <= b <= 13), (bs, hex(ws))
        assert z & 0x7F7F7F7F == 0 and ws & 0x7F7F7F7F == 0


CLS = "=XIDW"


def test_classify4_matches_reference_classifiers(hd):
    rng = random.Random(1)
    alpha = b"ACGTacgtN-"
    out = (ctypes.c_uint32 * 5)()
    for caller in (0, 1):
        for _ in range(5000):
            t = [rng.choice(alpha) for _ in range(4)]
            q = [rng.choice(alpha) for _ in range(4)]
            valid = [rng.random() < 0.8 for _ in range(4)]
            vm = sum(0x80 << (8 * i) for i in range(4) if valid[i])
            hd.hd_classify4(_word(t), _word(q), vm, caller, out)
            for i in range(4):
                want = (pyspec.cat_caller if caller else pyspec.cat_ext)(chr(t[i]), chr(q[i]))
                got = [CLS[c] for c in range(5) if (out[c] >> (8 * i + 7)) & 1]
                assert got == ([want] if valid[i] else []), (t, q, valid, caller)
            for c in range(5):
                assert out[c] & 0x7F7F7F7F == 0
        for _ in range(500):
            tb, qb = rng.choice(alpha), rng.choice(alpha)
            hd.hd_classify_prev(tb, qb, caller, out)
            want = (pyspec.cat_caller if caller else pyspec.cat_ext)(chr(tb), chr(qb))
            assert [CLS[c] for c in range(5) if out[c] >> 31] == [want]
            assert all(out[c] & 0x7FFFFFFF == 0 for c in range(5))


def test_nibble_flags_and_class_codes(hd):
    """the batch's byte-range masks (one multiply per word) and the per-column class codes the head emission reads"""
    hd.hd_nibble_flags.restype = hd.hd_class_codes.restype = ctypes.c_uint32
    hd.hd_nibble_flags.argtypes = [ctypes.c_uint32]
    hd.hd_class_codes.argtypes = [ctypes.c_uint32] * 3 + [ctypes.c_int]
    for n in range(64):  # only the low nibble counts
        assert hd.hd_nibble_flags(n) == sum(0x80 << (8 * i) for i in range(4) if (n >> i) & 1)
    # bytes [f, t) of a 16-byte chunk, word k: the nibble of the range mask == the old per-word range computation
    for f in range(17):
        for t in range(17):
            vm = ((1 << t) - 1) & ~((1 << f) - 1)
            for k in range(4):
                want = sum(0x80 << (8 * i) for i in range(4) if f <= 4 * k + i < t)
                assert hd.hd_nibble_flags(vm >> (4 * k)) == want
    rng = random.Random(4)
    alpha = b"ACGTacgtN-"
    for caller in (0, 1):
        for _ in range(3000):
            t = [rng.choice(alpha) for _ in range(4)]
            q = [rng.choice(alpha) for _ in range(4)]
            cw = hd.hd_class_codes(_word(t), _word(q), 0x80808080, caller)
            for i in range(4):
                want = (pyspec.cat_caller if caller else pyspec.cat_ext)(chr(t[i]), chr(q[i]))
                assert CLS[(cw >> (8 * i)) & 7] == want and (cw >> (8 * i)) & 0xF8 == 0


def test_prev_flags_and_funnel(hd):
    assert hd.hd_prev_flags(0x80000000, 0x00800080) == 0x80008080
    assert hd.hd_prev_flags(0, 0x80808080) == 0x80808000
    lo, hi = 0x44332211, 0x88776655
    assert [hd.hd_funnel(lo, hi, s) for s in range(4)] == [0x44332211, 0x55443322, 0x66554433, 0x77665544]


def test_decimal(hd):
    rng = random.Random(2)
    vals = [0, 9, 10, 99, 100, 999, 1000, 99999, 100000, 10**10 - 1, 10**10, 2**64 - 1, 2**32 - 1, 2**32, 2**32 + 1,
            10**9 - 1, 10**9, 4294967290, 42949672960, 10**19] + [10**k for k in range(20)] + [10**k - 1 for k in range(1, 20)] + \
           [rng.randrange(0, 2**64) >> rng.randrange(0, 64) for _ in range(3000)]
    buf = ctypes.create_string_buffer(32)
    for v in vals:
        assert hd.hd_ndigits(v) == len(str(v))
        n = hd.hd_write_u64(buf, v)
        assert buf.raw[:n] == str(v).encode()
    out = ctypes.c_uint64()
    for s, ok in ((b"0", 1), (b"+5", 1), (b"", 0), (b"+", 0), (b"-1", 0), (b"12a", 0), (b"18446744073709551615", 1),
                  (b"18446744073709551616", 0), (b"007", 1)):
        assert hd.hd_parse_u64(s, len(s), ctypes.byref(out)) == ok, s
        if ok:
            assert out.value == int(s)


def test_host_f32_and_natord_agree_with_oracle(hd):
    rng = random.Random(3)
    buf = ctypes.create_string_buffer(64)
    vals = [0.0, 1.0, 0.99, 0.999, 50.0, float("nan"), 1.5e10, 1e-7, 1.25e20, 3.4e38, 1e13, 9.9999e12, 1e-5, 9.99e-6]
    vals += [float(np.float32(rng.randrange(0, 10**7)) / np.float32(rng.randrange(1, 10**7))) for _ in range(3000)]
    vals += [float(np.float32(rng.random() * 10 ** rng.randrange(-12, 14))) for _ in range(2000)]
    for v in vals:
        hd.hd_fmt_f32(v, buf)
        assert buf.value.decode() == orc.fmt_f32(v), v
    words = ["chr1", "chr10", "chr2", "chr02", "chr 3", "x1.05", "x1.5", "a", "", "chrX", "ref.chr8", "ref.chr18", "9", "10", "010"]
    for a in words:
        for b in words:
            s = lambda x: (x > 0) - (x < 0)
            assert s(hd.hd_natord(a.encode(), b.encode())) == s(orc.natord(a, b)), (a, b)
