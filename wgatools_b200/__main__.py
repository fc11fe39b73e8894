"""`python -m wgatools_b200 <subcommand>` -- the in-scope subcommands of the reference CLI with the same flags,
aliases and defaults (src/cli.rs:20-343), running on the GPU path.  Global flags: -o/--outfile ('-' = stdout),
-r/--rewrite, -t/--threads (accepted, unused: the work runs on the GPU), -v (verbosity, stderr only).

  maf2paf|m2p [INPUT] [-q/--query-name NAME]                         src/cli.rs:42-49
  stat|st     [INPUT] [-f/--format maf|paf] [-e/--each] [-q NAME]    src/cli.rs:187-200
  call|c      [INPUT] [-n/--sample S] [-s/--snp] [-i/--inv] [-l/--svlen N] [--query-name NAME] [--query-regex RE] [-c/--chunk-size N]
                                                                    src/cli.rs:132-174
Errors are printed like the reference (`main.rs:14-22`) and exit with status 1.
"""
import argparse
import gzip
import os
import sys

from . import api


def _read_input(path):
    # src/utils.rs:135-169: None/'-' = stdin; gzip sniffed by magic number or .gz extension
    if path is None or path == "-":
        if sys.stdin.isatty():
            raise api.WGAError(-1, "Empty stdin, please add `-h` for help")
        data = sys.stdin.buffer.read()
    else:
        if not os.path.exists(path):
            raise api.WGAError(-1, f"File path `{path}` not exist")
        with open(path, "rb") as f:
            data = f.read()
    if data[:3] == b"\x1f\x8b\x08" or (path or "").endswith(".gz"):
        data = gzip.decompress(data)
    return data


def _open_output(path, rewrite):
    # src/utils.rs:181-246
    if path == "-":
        return sys.stdout.buffer
    if os.path.exists(path) and not rewrite:
        raise api.WGAError(-1, f"File `{path}` already exists, please add `-r` to rewrite it.")
    if path.endswith(".gz"):
        return gzip.open(path, "wb", compresslevel=6)
    return open(path, "wb")


def main(argv=None):
    ap = argparse.ArgumentParser(prog="wgatools", description="B200-native hot path of wgatools")
    ap.add_argument("-o", "--outfile", default="-")
    ap.add_argument("-r", "--rewrite", action="store_true")
    ap.add_argument("-t", "--threads", type=int, default=1)
    ap.add_argument("-v", "--verbose", action="count", default=0)
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("maf2paf", aliases=["m2p"])
    p.add_argument("input", nargs="?")
    p.add_argument("-q", "--query-name")
    p = sub.add_parser("stat", aliases=["st"])
    p.add_argument("input", nargs="?")
    p.add_argument("-f", "--format", default="maf", choices=["maf", "paf"])
    p.add_argument("-e", "--each", action="store_true")
    p.add_argument("-q", "--query-name")
    p = sub.add_parser("call", aliases=["c"])
    p.add_argument("input", nargs="?")
    p.add_argument("-n", "--sample", default="sample")
    p.add_argument("-s", "--snp", action="store_true")
    p.add_argument("-i", "--inv", action="store_true")
    p.add_argument("-l", "--svlen", type=int, default=50)
    p.add_argument("-f", "--format", default="maf", choices=["maf"])
    p.add_argument("--query-name")
    p.add_argument("--query-regex")
    p.add_argument("-c", "--chunk-size", type=int, default=1000000)
    a = ap.parse_args(argv)
    try:
        data = _read_input(a.input)
        out = _open_output(a.outfile, a.rewrite)
        if a.cmd in ("maf2paf", "m2p"):
            api.maf2paf(api.MAFReader(data), out, a.query_name)
        elif a.cmd in ("stat", "st"):
            if a.format == "maf":
                api.stat_maf(api.MAFReader(data), out, a.each, a.query_name)
            else:
                api.stat_paf(api.PAFReader(data), out, a.each)
        else:
            mafindex = None  # "<input>.index" next to a file input, like utils.rs:414-431 (stdin: no index)
            if a.input and a.input != "-" and os.path.exists(a.input + ".index"):
                import json
                with open(a.input + ".index") as fh:
                    mafindex = json.load(fh)
            api.call_var_maf(api.MAFReader(data), mafindex, out, a.snp, a.inv, a.svlen, False, a.sample, a.query_name,
                             a.query_regex, a.chunk_size)
        if out is not sys.stdout.buffer:
            out.close()
        else:
            out.flush()
    except api.WGAError as e:
        print(f"ERROR {e}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
