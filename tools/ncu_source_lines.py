"""Joins the per-instruction counters of an ncu capture (--page source --csv --print-source sass) with the line table of the
cubin it profiled (nvdisasm -gi), so that executed warp instructions and stall samples can be read per source line of the
NVRTC module although ncu could not import its sources (they exist only as strings inside the library).

usage: python tools/ncu_source_lines.py REPORT.ncu-rep CUBIN KERNEL [--top N] [--by outer|inner]
  CUBIN: written by LINON_DUMP_CUBIN=... (linon_compile_check) for the same variant; the instruction sequences must agree.
Prints a table: source line (outermost frame inside KERNEL, or innermost inlined frame), warp instructions executed, share,
stall samples, share.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def sass_of_report(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    cols = rows[hdr]
    ia, isrc, iex, ismp = cols.index("Address"), cols.index("Source"), cols.index("Instructions Executed"), cols.index("# Samples")
    res = []
    for r in rows[hdr + 1:]:
        if len(r) <= iex or not r[ia].startswith("0x"):
            continue
        res.append((r[isrc].strip(), int(r[iex] or 0), int(r[ismp] or 0)))
    return res


def lines_of_cubin(cubin, kernel):
    dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.strip() == f".text.{kernel}:")
    res, cur = [], ("?", 0, "?", 0)
    file_re = re.compile(r'//## File "([^"]+)", line (\d+)')
    chain_open = False  # an inlined instruction is annotated by consecutive lines, innermost frame first, outermost last
    for l in dis[start + 1:]:
        if l.startswith("//---------------------") or l.lstrip().startswith(".section"):
            break
        if "//## File" in l:
            f, n = file_re.findall(l)[0]
            loc = (f.split("/")[-1], int(n))
            cur = (cur[0], cur[1], loc[0], loc[1]) if chain_open else (loc[0], loc[1], loc[0], loc[1])
            chain_open = True
            continue
        chain_open = False
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            res.append((m.group(2).strip(), cur))
    return res


def main():
    rep, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    by = sys.argv[sys.argv.index("--by") + 1] if "--by" in sys.argv else "outer"
    a, b = sass_of_report(rep, kernel), lines_of_cubin(cubin, kernel)
    if len(a) != len(b):
        raise SystemExit(f"instruction counts differ: report {len(a)}, cubin {len(b)} - not the same binary")
    bad = sum(1 for (s, _, _), (t, _) in zip(a, b) if s.split()[0].lstrip("@!P0123456789U ") [:3] != t.split()[0].lstrip("@!P0123456789U ")[:3])
    agg = defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for (s, ex, smp), (t, (fi, li, fo, lo)) in zip(a, b):
        key = (fo, lo) if by == "outer" else (fi, li)
        agg[key][0] += ex
        agg[key][1] += smp
        agg[key][2] += 1
        tot_i += ex
        tot_s += smp
    print(f"# {kernel}: {len(a)} SASS instructions, {tot_i} warp instructions executed, {tot_s} stall samples; opcode mismatches {bad}")
    print(f"| line ({by}) | static instr | warp instr executed | share | samples | share |")
    print("|---|---|---|---|---|---|")
    for (f, l), (ex, smp, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| {f}:{l} | {n} | {ex} | {100 * ex / max(tot_i, 1):.2f} % | {smp} | {100 * smp / max(tot_s, 1):.2f} % |")


if __name__ == "__main__":
    main()
