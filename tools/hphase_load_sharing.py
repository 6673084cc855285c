"""How many psi loads of the compressed H phase could slices share?  For the circuits of the LiH-10q batch with support
dimension d = 8 / 9: in-span X-mask slices per circuit, and the number of distinct LANE parts of their compressed X masks
with today's mapping (a thread's amplitudes differ in the top bits) and with the best per-circuit choice of the register
coordinates.   python tools/hphase_load_sharing.py"""
import sys, itertools
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, ctape_interp as CT
from oracle import sim
model, pool, params, ks = bench.make_workload("lih_10q", 3000, 0)
terms = model.terms()
xm, im, V = CT.sign_tables(10, terms)
stats = {}
for k in ks:
    r = CT.compile_ctape(10, pool, [int(x) for x in k], xm)
    d = r["d"]
    if d < 8: continue
    xs = [x for x, s in r["slices"]]
    nsl = len(xs)
    T = 32 if d == 8 else 64
    lt = 5 if d == 8 else 6
    nreg = d - lt                      # register bits (amplitudes per thread = 2^nreg)
    # current mapping: register bits = top nreg bits
    top = len({x & ((1 << lt) - 1) for x in xs})
    best = top
    for regbits in itertools.combinations(range(d), nreg):
        mask = ((1 << d) - 1) & ~sum(1 << b for b in regbits)
        best = min(best, len({x & mask for x in xs}))
    st = stats.setdefault(d, [0, 0, 0, 0])
    st[0] += 1; st[1] += nsl; st[2] += top; st[3] += best
for d, (cnt, nsl, top, best) in sorted(stats.items()):
    print("d=%d: %d circuits, mean in-span slices %.1f, distinct lane-parts with top register bits %.1f, best choice of register bits %.1f" % (d, cnt, nsl / cnt, top / cnt, best / cnt))
