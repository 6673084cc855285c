"""Top CUDA source lines of an ncu --set full --import-source on capture, by executed warp instructions.
Usage: python tools/ncu_srclines.py <report.ncu-rep> [top N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
out, cur, hdr = [], None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            ie, te, sm = int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])
        except ValueError:
            continue
        if ie > 0:
            out.append((ie, te, sm, cur, int(r[0]), r[1].strip()[:100]))
out.sort(reverse=True)
tot = sum(o[0] for o in out)
tots = sum(o[2] for o in out) or 1
print("total warp instructions %d, stall samples %d" % (tot, tots))
for ie, te, sm, f, ln, src in out[:top]:
    print("%10d %5.1f%%  thr/inst %4.1f  samples %4.1f%%  %s:%d  %s" % (ie, 100.0 * ie / tot, te / ie, 100.0 * sm / tots, f, ln, src))
