#!/usr/bin/env python
"""Summarise the source view of an ncu report by function and by line.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv
    python tools/ncu_lines.py X.csv [fiasco_b200/csrc/tile_kernel.cu] [top]

Per function (the enclosing definition of each source line): share of executed warp instructions, of
all warp-stall samples, of the samples that are NOT barrier waits (= what the working warps do), and
the split of those.  Then the hottest lines by non-barrier samples.
"""
import csv
import re
import sys


def functions_of(src):
    """line number -> name of the function whose definition encloses it (crude: a line that starts a
    definition is one whose previous non-blank line carries the return type / template and which
    names an identifier followed by '(' at column 0)."""
    lines = open(src, errors="replace").read().split("\n")
    owner, cur = {}, "?"
    rx = re.compile(r"^([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    for i, ln in enumerate(lines, 1):
        m = rx.match(ln)
        if m and not ln.startswith(("if", "for", "while", "switch", "return", "static_assert")):
            cur = m.group(1)
        m2 = re.match(r"^__device__.*?([A-Za-z_][A-Za-z0-9_]*)\s*\(", ln)
        if m2:
            cur = m2.group(1)
        owner[i] = cur
    return owner, lines


def main():
    path = sys.argv[1]
    src = sys.argv[2] if len(sys.argv) > 2 else "fiasco_b200/csrc/tile_kernel.cu"
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    owner, text = functions_of(src)
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = None
    per_line = {}
    for r in rows:
        if len(r) > 10 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))
        ln = int(r[0])
        e = per_line.setdefault(ln, {})
        for k, v in d.items():
            if k.startswith("stall_") and "(Not Issued)" not in k or k in ("Instructions Executed", "# Samples"):
                try:
                    e[k] = e.get(k, 0) + float(v)
                except ValueError:
                    pass
    tot_inst = sum(e.get("Instructions Executed", 0) for e in per_line.values()) or 1
    keys = sorted({k for e in per_line.values() for k in e if k.startswith("stall_")})
    tot = {k: sum(e.get(k, 0) for e in per_line.values()) for k in keys}
    tot_s = sum(tot.values()) or 1
    nobar_tot = tot_s - tot.get("stall_barrier", 0) or 1
    print("warp instructions executed: %.0f; stall samples: %.0f" % (tot_inst, tot_s))
    print("shares of all samples: " + ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot_s)
                                                for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v / tot_s > 0.002))
    fn = {}
    for ln, e in per_line.items():
        f = fn.setdefault(owner.get(ln, "?"), {})
        for k, v in e.items():
            f[k] = f.get(k, 0) + v
    print("\n%-26s %6s %6s %7s | of non-barrier samples: %6s %6s %6s %7s %6s" % (
        "function", "inst%", "samp%", "nobar%", "longsb", "noinst", "wait", "shortsb", "select"))
    for name, f in sorted(fn.items(), key=lambda kv: -(sum(v for k, v in kv[1].items() if k.startswith("stall_"))
                                                       - kv[1].get("stall_barrier", 0))):
        s = sum(v for k, v in f.items() if k.startswith("stall_"))
        nb = s - f.get("stall_barrier", 0)
        if s / tot_s < 0.002 and f.get("Instructions Executed", 0) / tot_inst < 0.002:
            continue
        print("%-26s %6.2f %6.2f %7.2f | %29.2f %6.2f %6.2f %7.2f %6.2f" % (
            name[:26], 100 * f.get("Instructions Executed", 0) / tot_inst, 100 * s / tot_s, 100 * nb / nobar_tot,
            100 * f.get("stall_long_sb", 0) / nobar_tot, 100 * f.get("stall_no_inst", 0) / nobar_tot,
            100 * f.get("stall_wait", 0) / nobar_tot, 100 * f.get("stall_short_sb", 0) / nobar_tot,
            100 * f.get("stall_selected", 0) / nobar_tot))
    print("\nhottest lines by non-barrier samples (line: nobar%  inst%  barrier% | source)")
    order = sorted(per_line.items(), key=lambda kv: -(sum(v for k, v in kv[1].items() if k.startswith("stall_"))
                                                     - kv[1].get("stall_barrier", 0)))
    for ln, e in order[:top]:
        s = sum(v for k, v in e.items() if k.startswith("stall_"))
        nb = s - e.get("stall_barrier", 0)
        print("%5d: %5.2f %5.2f %5.2f | %s" % (ln, 100 * nb / nobar_tot, 100 * e.get("Instructions Executed", 0) / tot_inst,
                                               100 * e.get("stall_barrier", 0) / tot_s,
                                               text[ln - 1].strip()[:110] if ln - 1 < len(text) else ""))


if __name__ == "__main__":
    main()
