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


_CS = re.compile(r"(:[0-9]+|\*[a-z][a-z]|[=\+\-][A-Za-z]+)")  # the reference's own regex, paf.rs:165


def cs_to_cigar(cs):  # paf.rs:159-218; None = a `:N` beyond usize (parse().unwrap() panics)
    out, last_op, last_len = [], "M", 0
    for part in _CS.findall(cs):
        k = part[0]
        if k == ":":
            n = int(part[1:])
            if n >= 2 ** 64:
                return None
            if last_op == "M":
                last_len = (last_len + n) % 2 ** 64
            else:
                if last_len:
                    out.append(f"{last_len}{last_op}")
                last_op, last_len = "M", n
        elif k in "-+":
            if last_len:
                out.append(f"{last_len}{last_op}")
            out.append(f"{len(part) - 1}{'D' if k == '-' else 'I'}")
            last_op, last_len = "M", 0
        elif k == "*":
            if last_op == "X":
                last_len += 1
            else:
                if last_len:
                    out.append(f"{last_len}{last_op}")
                last_op, last_len = "X", 1
    if last_len:
        out.append(f"{last_len}{last_op}")
    return "".join(out)


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


_COMP = str.maketrans("ACGTacgt", "TGCAtgca")


def pseudo_maf(recs, fasta=None):
    """pseudomaf.rs:78-237 + cigar.rs:744-804 for ONE target; recs = (q, qsize, qs, qe, strand, t, tsize, ts, te, cg) in file
    order, fasta = {name: sequence} or None.  Queries in first-occurrence order; raises on the reference's panics."""
    groups, order = {}, []
    for r in recs:
        g = groups.setdefault(r[0], [])
        if r[0] not in order:
            order.append(r[0])
        g.append(r)
    out = ["a score=0\n"]
    first = True
    tsize = 0
    for q in order:
        g = sorted(groups[q], key=lambda r: r[7])   # distinct target_start assumed (ties are unspecified)
        firstq, lte = True, 0
        for (qn, qsize, qs, qe, strand, t, tsz, ts, te, cg) in g:
            tsize = tsz
            if first:
                seq = fasta[t][0:tsz] if fasta is not None else "N" * tsz
                out.append(f"s\t{t}\t0\t{tsz}\t+\t{tsz}\t{seq}\n")
                first = False
            if firstq:
                out.append(f"s\t{qn}\t0\t{qsize}\t+\t{qsize}\t")
            overlap = 0
            if ts > lte:
                out.append("-" * (ts - lte))
            else:
                if lte > te:
                    continue
                overlap = lte - ts
            lte = te
            s = fasta[qn][qs:qe] if fasta is not None else ""
            if strand == "-":
                s = s[::-1].translate(_COMP)
            s = list(s)
            off = 0
            for n, op in cigar_tokens(cg):
                if op in "M=X":
                    if fasta is not None:
                        off += n
                    else:
                        s += ("1" if op != "X" else "0") * n
                elif op in "IS":
                    if fasta is not None:
                        assert off + n <= len(s), "drain panics"
                        del s[off:off + n]
                elif op == "D":
                    if fasta is not None:
                        assert off <= len(s), "insert_str panics"
                        s[off:off] = "-" * n
                        off += n
                    else:
                        s += "-" * n
            assert overlap <= len(s), "drain panics"
            out.append("".join(s[overlap:]))
            firstq = False
        assert lte <= tsize, "subtract overflows"
        out.append("-" * (tsize - lte) + "\n")
    out.append("\n")
    return "".join(out)


def call_paf_rows(rec, fasta_t, fasta_q, snp, svlen):
    """caller.rs:610-822 for one record -> list of VCF body rows; raises AssertionError on the reference's panics and
    ValueError on its real errors (<INV> row).  Sequences are fetched with an inclusive end, no reverse complement."""
    q, qsize, qs, qe, strand, t, tsize, ts, te, cg = rec
    tseq, qseq = fasta_t[t][ts:te + 1], fasta_q[q][qs:qe + 1]
    up = lambda s: s.upper()
    ok = lambda s: all(c in "ACGTN" for c in s.upper())
    sfx = "N" if strand == "-" else "P"
    rows = []
    if strand == "-":
        assert len(tseq) >= 1
        if not ok(tseq[0]):
            raise ValueError("bad base")
        rows.append(f"{t}\t{ts + 1}\t.\t{up(tseq[0])}\t<INV>\t.\t.\tSVTYPE=INV;END={te}\tGT:QI\t1|1:{q}@{qs}@{qe}@{sfx}\n")
    nest = "INV_NEST=TRUE;" if strand == "-" else ""
    tp, qp, after_m = ts, qs, False
    toks = []
    i = 0
    while i < len(cg):   # stop silently at the first malformed token (the closure's Result is discarded)
        m = _TOK.match(cg, i)
        d, op = m.group(1), m.group(2)
        if len(op) != 1 or d == "" or int(d) >= 2 ** 64:
            break
        toks.append((int(d), op)); i = m.end()
    for n, op in toks:
        if op in "M=":
            tp += n; qp += n; after_m = True
        elif op == "X":
            if snp:
                for _ in range(n):
                    a, b = tp - ts, qp - qs
                    assert a + 1 <= len(tseq) and b + 1 <= len(qseq)
                    if not (ok(tseq[a]) and ok(qseq[b])):
                        return rows
                    rows.append(f"{t}\t{tp + 1}\t.\t{up(tseq[a])}\t{up(qseq[b])}\t.\t.\t.\tGT:QI\t1|1:{q}@{qp}@{sfx}\n")
                    tp += 1; qp += 1
            else:
                tp += n; qp += n
            after_m = True
        elif op in "ID":
            if n > svlen and after_m:
                assert tp > ts and qp > qs
                a, b = tp - ts - 1, qp - qs - 1
                if op == "I":
                    assert a + 1 <= len(tseq) and b + n + 1 <= len(qseq)
                    ref, alt = tseq[a], qseq[b:b + n + 1]
                    info, qi = f"{nest}SVTYPE=INS;SVLEN={n};END={tp}", f"{q}@{qp}@{qp + n}@{sfx}"
                else:
                    assert a + n + 1 <= len(tseq) and b + 1 <= len(qseq)
                    ref, alt = tseq[a:a + n + 1], qseq[b]
                    info, qi = f"{nest}SVTYPE=DEL;SVLEN={n};END={tp + n}", f"{q}@{qp}@{qp}@{sfx}"
                if not (ok(ref) and ok(alt)):
                    return rows
                rows.append(f"{t}\t{tp}\t.\t{up(ref)}\t{up(alt)}\t.\t.\t{info}\tGT:QI\t1|1:{qi}\n")
            if op == "I":
                qp += n
            else:
                tp += n
            after_m = False
        else:
            return rows
    return rows


# ---------------------------------------------------------------- chain reader / chain2paf / chain2maf (second spelling)
def chain_records(text: str):
    """src/parser/chain.rs:352-383 chain_parser, applied to what is left of the text.  Independent of the C oracle: this
    one works on the whole string with str methods, the oracle walks bytes.  Raises ValueError(kind)."""
    out, rest = [], text
    while rest:
        if not rest.startswith("chain"):
            raise ValueError("nom" if len(rest) >= 10 else "panic")
        rest = rest[5:]
        nl = rest.find("\n")
        header = rest if nl < 0 else rest[:nl]
        if header.endswith("\r"):
            header = header[:-1]
        if "\r" in header:
            raise ValueError("nom")
        f = header.split()
        names = ["score", "target_name", "target_size", "target_strand", "target_start", "target_end", "query_name", "query_size",
                 "query_strand", "query_start", "query_end", "chain_id"]
        vals = {}
        for i, nm in enumerate(names):
            if i >= len(f):
                raise ValueError("missing:" + nm)
            if nm == "score":
                try:
                    float(f[i])
                    if "_" in f[i] or f[i].strip() != f[i] or f[i].lower().lstrip("+-") in ("infinity", "inf", "nan") and False:
                        raise ValueError
                except ValueError:
                    raise ValueError("float:" + f[i])
            elif nm.endswith("strand"):
                if f[i] not in "+-" or len(f[i]) != 1:
                    raise ValueError("strand:" + f[i])
            elif not nm.endswith("name"):
                if not _is_u64(f[i]):
                    raise ValueError("int:" + f[i])
            vals[nm] = f[i]
        if nl < 0:
            raise ValueError("panic")          # line_ending on an empty rest -> e.input[..10] panics
        rest = rest[nl + 1:]
        lines, err = [], None
        n_lines = 0
        while True:
            nl = rest.find("\n")
            if nl <= 0:
                break
            cand = rest[:nl]
            if any(c in "chain" for c in cand):
                break
            rest = rest[nl + 1:]
            n_lines += 1
            if err:
                continue
            t = cand.split()
            if not t:
                err = "missing:size"
                continue
            bad = [x for x in t[:3] if not _is_u64(x)]
            if bad:
                err = "int:" + bad[0]
                continue
            lines.append((int(t[0]), int(t[1]) if len(t) > 1 else 0, int(t[2]) if len(t) > 2 else 0))
        if n_lines == 0:
            raise ValueError("nom" if len(rest) >= 10 else "panic")
        if err:
            raise ValueError(err)
        c = rest.find("c")
        rest = "" if c < 0 else rest[c:]
        out.append((vals, lines))
    return out


def _is_u64(s):
    t = s[1:] if s.startswith("+") else s
    return t.isascii() and t.isdigit() and int(t) < (1 << 64)


def chain2paf(text: str) -> str:
    rows = []
    for h, lines in chain_records(text):
        match = sum(l[0] for l in lines)
        dele = sum(l[1] for l in lines)
        cg = "".join(f"{s}M" + (f"{td}I" if td else "") + (f"{qd}D" if qd else "") for s, qd, td in lines)
        rows.append("\t".join([h["query_name"], str(int(h["query_size"])), str(int(h["query_start"])), str(int(h["query_end"])),
                               h["query_strand"], h["target_name"], str(int(h["target_size"])), str(int(h["target_start"])),
                               str(int(h["target_end"])), str(match), str(match + dele), "255", "cg:Z:" + cg]) + "\n")
    return "".join(rows)


def chain2maf(text: str, fa: dict, t_path="t.fa", q_path="q.fa") -> str:
    """fa: name -> sequence (both FASTAs); well-formed inputs only (every name present, sizes consistent)"""
    comp = str.maketrans("ACGTacgt", "TGCAtgca")
    out = [f"#maf version=1.6 convert_from=chain t_seq_path={t_path} q_seq_path={q_path}\n"]
    for h, lines in chain_records(text):
        ts, te, qs, qe = int(h["target_start"]), int(h["target_end"]), int(h["query_start"]), int(h["query_end"])
        t = fa[h["target_name"]][ts:te]
        q = fa[h["query_name"]][qs:qe]
        neg = h["query_strand"] == "-"
        if neg:
            q = q[::-1].translate(comp)
        off = 0
        for size, qd, td in lines:
            off += size
            if td:
                assert off <= len(t)
                t = t[:off] + "-" * td + t[off:]
                off += td
            if qd:
                assert off <= len(q)
                q = q[:off] + "-" * qd + q[off:]
                off += qd
        qsz = int(h["query_size"])
        out.append("a score=255\n")
        out.append(f"s\t{h['target_name']}\t{ts}\t{te - ts}\t{h['target_strand']}\t{int(h['target_size'])}\t{t}\n")
        out.append(f"s\t{h['query_name']}\t{qsz - qe if neg else qs}\t{qe - qs}\t{h['query_strand']}\t{qsz}\t{q}\n\n")
    return "".join(out)


def plot_segments(ops, ref0, qry0, neg, cutoff):
    """cigar.rs:807-952 spelled by stretches instead of the reference's state machine: I / D ops longer than `cutoff` are
    segments of their own and cut the op list into stretches; a stretch yields one M segment from the offsets at its first
    M/=/X op to the offsets at its end; other ops are ignored.  ops = [(len, op)]; returns [(rs, re, qs, qe, c)]."""
    segs, stretch = [], []
    ref, qry = ref0, qry0

    def close():
        nonlocal ref, qry
        r, q = ref, qry
        start = None
        for n, op in stretch:
            if op in "M=X":
                if start is None:
                    start = (r, q)
                r += n; q += n
            elif op == "I":
                q += n
            elif op == "D":
                r += n
        if start is not None:
            segs.append((start[0], r, start[1], q, "M"))
        ref, qry = r, q
        stretch.clear()

    for n, op in ops:
        if op in "ID" and n > cutoff:
            close()
            if op == "I":
                segs.append((ref, ref, qry, qry + n, "I")); qry += n
            else:
                segs.append((ref, ref + n, qry, qry, "D")); ref += n
        else:
            stretch.append((n, op))
    close()
    if neg:
        segs = [(rs, re, qe, qs, c) for rs, re, qs, qe, c in segs]
    return segs


def plot_csv(recs, cutoff):  # recs = [(ref_chro, query_chro, ref_start, query_start, neg, ops)]
    rows = []
    for tname, qname, ref0, qry0, neg, ops in recs:
        for rs, re_, qs, qe, c in plot_segments(ops, ref0, qry0, neg, cutoff):
            rows.append(f"{rs},{re_},{qs},{qe},{c},{tname},{qname}\n")
    return ("ref_start,ref_end,query_start,query_end,cigar,ref_chro,query_chro\n" + "".join(rows)) if rows else ""
