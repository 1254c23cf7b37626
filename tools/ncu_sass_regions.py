#!/usr/bin/env python3
"""Summarises the source (SASS) page of an `ncu --set full --import-source on` capture:
     ncu -i X.ncu-rep --page source --csv > X_source.csv ; python tools/ncu_sass_regions.py X_source.csv [top]
Splits the kernel's SASS into functions (regions end at RET / EXIT), and prints per region its share of the warp-stall
samples, its executed warp instructions, opcode mix and stall mix, then the totals by opcode.  Runs on the GPU box so that only
this text (not the tens of megabytes of the report) has to travel back."""
import csv
import sys
from collections import Counter, defaultdict


def main():
    fn = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 18
    rows = list(csv.reader(open(fn, newline="")))
    kernel = rows[0][1] if rows and rows[0] and rows[0][0] == "Kernel Name" else "?"
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {name: i for i, name in enumerate(hdr)}
    stall_cols = [(n, i) for n, i in col.items() if n.startswith("stall_") and "Not Issued" not in n]
    regions, cur = [], {"n": 0, "samples": 0, "exec": 0, "ops": Counter(), "stalls": Counter()}
    by_op_s, by_op_e = Counter(), Counter()
    tot_s = tot_e = 0

    def num(r, name):
        try:
            return int(float(r[col[name]]))
        except (ValueError, IndexError, KeyError):
            return 0

    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) - 2 or not r[0].startswith("0x"):
            continue
        src = r[col["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        base = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "IADD3", "LDL", "STL", "LDC", "LDCU", "SHF", "LDS", "STS", "LDG", "STG", "SHFL")) else op.split(".")[0]
        s, e = num(r, "# Samples"), num(r, "Instructions Executed")
        cur["n"] += 1; cur["samples"] += s; cur["exec"] += e
        cur["ops"][base] += e
        for n, i in stall_cols:
            try:
                cur["stalls"][n[6:]] += int(float(r[i]))
            except ValueError:
                pass
        by_op_s[base] += s; by_op_e[base] += e
        tot_s += s; tot_e += e
        if op.split(".")[0] in ("RET", "EXIT"):
            regions.append(cur)
            cur = {"n": 0, "samples": 0, "exec": 0, "ops": Counter(), "stalls": Counter()}
    if cur["n"]:
        regions.append(cur)
    print("kernel:", kernel[:160])
    print("total stall samples %d, executed warp instructions %d, %d SASS lines, %d regions" % (tot_s, tot_e, sum(g["n"] for g in regions), len(regions)))
    print(" samples  instrs   executed  top opcodes (share of the region's executed instructions) | top stall reasons")
    for g in sorted(regions, key=lambda g: -g["samples"])[:top]:
        ops = ", ".join("%s %d%%" % (o, round(100.0 * c / max(1, g["exec"]))) for o, c in g["ops"].most_common(5))
        st_tot = sum(g["stalls"].values()) or 1
        sts = ", ".join("%s %d%%" % (o, round(100.0 * c / st_tot)) for o, c in g["stalls"].most_common(4))
        print("  %5.1f%%  %6d  %8.3fG  %s | %s" % (100.0 * g["samples"] / max(1, tot_s), g["n"], g["exec"] / 1e9, ops, sts))
    print("\nby opcode: samples %, executed %")
    for o, c in by_op_s.most_common(22):
        print("%-18s %5.1f%%  %5.1f%%" % (o, 100.0 * c / max(1, tot_s), 100.0 * by_op_e[o] / max(1, tot_e)))


if __name__ == "__main__":
    main()
