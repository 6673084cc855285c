"""Attribute an ncu source-page CSV (SASS view) to CUDA source lines through nvdisasm line info.
Usage: python tools/ncu_lines.py <src.csv> <kernel.sass from nvdisasm --print-line-info>"""
import collections
import csv
import re
import sys


def main():
    src_csv, sass = sys.argv[1], sys.argv[2]
    cur, amap = None, {}
    for ln in open(sass):
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
        if m:
            sites = re.findall(r'inlined at ".*?([^/"]+)", line (\d+)', ln)
            cur = [(m.group(1), int(m.group(2)))] + [(a, int(b)) for a, b in sites]
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            amap[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    col = {k: hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "# Samples", "L1 Wavefronts Shared",
                                     "L1 Wavefronts Shared Excessive", "L1 Tag Requests Global")}
    base = int(rows[2][col["Address"]], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
    ops = collections.Counter()
    tot = 0
    for r in rows[2:]:
        a = int(r[col["Address"]], 16) - base
        ch = amap.get(a)
        key = tuple(l for f, l in ch) if ch else ("?",)
        n = int(r[col["Instructions Executed"]] or 0)
        tot += n
        s = r[col["Source"]].split()
        op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
        ops[op] += n
        v = agg[key]
        v[0] += n
        v[1] += int(r[col["# Samples"]] or 0)
        v[2] += int(r[col["L1 Wavefronts Shared"]] or 0)
        v[3] += int(r[col["L1 Wavefronts Shared Excessive"]] or 0)
        v[4] += int(r[col["L1 Tag Requests Global"]] or 0)
    print("instructions %.3fG; opcode mix: %s" % (tot / 1e9, "  ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in ops.most_common(16))))
    print("%-28s %8s %8s %10s %10s %10s" % ("line (inlined-at chain)", "inst%", "samples", "smem wf M", "excess M", "gl tags M"))
    for k, v in sorted(agg.items(), key=lambda kv: -(kv[1][0]))[:45]:
        print("%-28s %8.2f %8d %10.1f %10.1f %10.1f" % (str(k), 100 * v[0] / tot, v[1], v[2] / 1e6, v[3] / 1e6, v[4] / 1e6))


if __name__ == "__main__":
    main()
