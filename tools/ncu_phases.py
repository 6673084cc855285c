"""Attribute an ncu source-page CSV of eval_kernel to its phases (prologue / forward / H / adjoint).
Usage: python tools/ncu_phases.py <src.csv> <kernel.sass from nvdisasm --print-line-info>"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    cur, amap = None, {}
    for ln in open(sys.argv[2]):
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
        if m:
            sites = re.findall(r'inlined at ".*?([^/"]+)", line (\d+)', ln)
            cur = [(m.group(1), int(m.group(2)))] + [(a, int(b)) for a, b in sites]
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            amap[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[1]
    iA, iS, iI, iSrc = (hdr.index(k) for k in ("Address", "# Samples", "Instructions Executed", "Source"))
    base = int(rows[2][iA], 16)
    src = open(os.path.join(ROOT, "qas_b200", "csrc", "kernels.cuh")).read().split("\n")

    def find(pat):
        for i, l in enumerate(src):
            if pat in l:
                return i + 1
        raise KeyError(pat)
    L_pro, L_fwd, L_h, L_adj = (find(x) for x in ("// ---- 1. fetch the compiled tape record", "// ---- 2. forward sweep",
                                                  "// ---- 3. lambda = H psi", "// ---- 4. adjoint sweep"))
    L_kernel = find("__global__ void __launch_bounds__(256, MINB) eval_kernel")

    def region(ch):
        if not ch:
            return "?"
        body = [l for f, l in ch if f == "kernels.cuh" and l >= L_kernel]
        if not body:
            return "?inner"
        L = body[-1]
        return "0.setup" if L < L_pro else "1.prologue" if L < L_fwd else "2.forward" if L < L_h else "3.Happly+E" if L < L_adj else "4.adjoint"
    agg = collections.defaultdict(lambda: [0, 0])
    tot = [0, 0]
    for r in rows[2:]:
        g = region(amap.get(int(r[iA], 16) - base))
        agg[g][0] += int(r[iI]); agg[g][1] += int(r[iS]); tot[0] += int(r[iI]); tot[1] += int(r[iS])
    for g, v in sorted(agg.items()):
        print("%-12s inst %5.1f%%  stall-samples %5.1f%%" % (g, 100 * v[0] / tot[0], 100 * v[1] / tot[1]))
    for r in sorted(rows[2:], key=lambda r: -int(r[iS]))[:16]:
        ch = amap.get(int(r[iA], 16) - base)
        print("%5.2f%% %-11s %-60s %s" % (100 * int(r[iS]) / tot[1], region(ch), r[iSrc][:60], ch[:2] if ch else ""))


if __name__ == "__main__":
    main()
