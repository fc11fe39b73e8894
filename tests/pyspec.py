"""Tiny pure-Python restatement of the reference's per-record arithmetic (second, independent
spelling of the oracle; used only to cross-check oracle/wga_oracle.c on small inputs).

Follows /root/reference/src/parser/cigar.rs:298-432, :629-707 and src/tools/caller.rs:159-265,388-608.
"""
from itertools import groupby


def cat_ext(a, b):  # cigar.rs:298-308
    if a == b:
        return "="
    if a == "-":
        return "I"
    if b == "-":
        return "D"
    return "X"


def cat_caller(a, b):  # cigar.rs:314-328
    if a == "-":
        return "W" if b == "-" else "I"
    if b == "-":
        return "D"
    return "=" if a == b else "X"


def maf_seq_to_cigar(t, q, neg):  # cigar.rs:344-432
    c = [0] * 11
    if neg:
        c[10] = 1
    cg = []
    for k, g in groupby(zip(t, q), key=lambda p: cat_ext(*p)):
        n = len(list(g))
        if k == "=":
            c[0] += n
        elif k == "X":
            c[1] += n
        elif k == "I":
            if neg:
                c[6] += 1; c[7] += n
            else:
                c[2] += 1; c[3] += n
        elif k == "D":
            if neg:
                c[8] += 1; c[9] += n
            else:
                c[4] += 1; c[5] += n
        cg.append(f"{n}{k}")
    return c, "".join(cg)


def chunk_grid(t, q, C, K):  # caller.rs:109-149,159-219
    L = len(t)
    out = []
    s = 0
    while s < L:
        p = min(s + C, L)
        e = p
        run = 0
        in_gap = False
        for i in range(s, p):
            if t[i] == "-" or q[i] == "-":
                in_gap = True
                run += 1
            elif in_gap:
                if run >= K:
                    e = i
                in_gap = False
                run = 0
        if in_gap and run >= K:
            e = p
            for j in range(p, L):
                if t[j] != "-" and q[j] != "-":
                    e = j
                    break
        out.append((s, e))
        s = e
    return out
