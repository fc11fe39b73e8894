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


# ---- second-wave commands on PAF records (independent spelling; small inputs only) -----------------------------
import re

_TOK = re.compile(r"(\d*)(\D*)")


def cigar_tokens(cg):  # cigar.rs:43-75: digits* then non-digits*; None when a token is malformed
    out, i = [], 0
    while i < len(cg):
        m = _TOK.match(cg, i)
        d, op = m.group(1), m.group(2)
        if len(op) != 1 or d == "" or int(d) >= 2 ** 64:
            return None
        out.append((int(d), op))
        i = m.end()
    return out


def paf_counts(cg, neg):  # cigar.rs:629-707
    c = [0] * 11
    if neg:
        c[10] = 1
    for n, op in cigar_tokens(cg):
        if op in "M=":
            c[0] += n
        elif op == "X":
            c[1] += n
        elif op == "I":
            c[6 if neg else 2] += 1; c[7 if neg else 3] += n
        elif op == "D":
            c[8 if neg else 4] += 1; c[9 if neg else 5] += n
        else:
            return None
    return c


def paf_chain(rec, idx):  # converter.rs:148-173, chain.rs:92-204, cigar.rs:202-295,458-490
    q, qsize, qs, qe, strand, t, tsize, ts, te, cg = rec
    toks = cigar_tokens(cg)
    hi = hd = ti = td = 0
    head = True
    for n, op in toks:
        if op in "M=X":
            ti = td = 0; head = False
        elif op == "I":
            hi += n if head else 0; ti = n
        else:
            hd += n if head else 0; td = n
    W = 2 ** 64  # u64 arithmetic wraps like release Rust
    ts, te = (ts + hd) % W, (te - td) % W
    if strand == "+":
        qs, qe = (qs + hi) % W, (qe - ti) % W
    else:
        qs = (qsize - (qe - hi)) % W
        qe = (qsize - (qs + ti)) % W
    out = [f"chain\t255\t{t}\t{tsize}\t+\t{ts}\t{te}\t{q}\t{qsize}\t{strand}\t{qs}\t{qe}\t{idx}"]
    size = qd = tdf = 0
    for n, op in toks:
        if op in "M=X":
            if size != 0 and tdf + qd != 0:
                out.append(f"{size}\t{qd}\t{tdf}"); size = 0
            size += n; tdf = qd = 0
        elif op == "I":
            tdf += n
        else:
            qd += n
    out.append(str(size))
    return "\n".join(out) + "\n\n"


def paf_cov(recs):  # pafcov.rs:13-60, cigar.rs:710-741; recs = (target, tlen, tstart, cigar)
    cov, order = {}, []
    for t, tlen, ts, cg in recs:
        if t not in cov:
            cov[t] = [0] * tlen; order.append(t)
        v, pos = cov[t], ts
        for n, op in cigar_tokens(cg):
            if op in "M=":
                for i in range(pos, min(pos + n, len(v))):
                    v[i] += 1
                pos += n
            elif op not in "IS":
                pos += n
    return "".join(f"{t}\t{i}\t{i + 1}\t{c}\n" for t in order for i, c in enumerate(cov[t]))
