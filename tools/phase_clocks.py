"""Where does a circuit's latency go?  Runs one batch of a bench workload with the evaluation
kernel's phase clocks switched on and prints the mean SM cycles a team spends per circuit in each
phase.   QAS_B200_PHASE_CLOCKS=1 python tools/phase_clocks.py --workload lih_10q"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["QAS_B200_PHASE_CLOCKS"] = "1"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="lih_10q")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--no-grad", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import bench
    import __graft_entry__ as ge
    ge.build()
    model, pool, params, ks = bench.make_workload(args.workload, args.batch, 0)
    ev = model.evaluator(pool, 3, 0)
    for rep in range(3):
        ev.evaluate(ks, params, want_grad=not args.no_grad)
        out = (ctypes.c_uint64 * 4)()
        outc = (ctypes.c_uint64 * (8 * 16))()
        dmin = ctypes.c_int(0)
        rc = ev._lib.qas_debug_phase_clocks_c(ev._h, outc, 16, ctypes.byref(dmin), 0)
        assert rc == 0, ev._lib.qas_last_error(ev._h)
        rc = ev._lib.qas_debug_phase_clocks(ev._h, out, 1)
        assert rc == 0, ev._lib.qas_last_error(ev._h)
    st = ev.stats()
    names = ["prologue", "forward", "H+energy", "adjoint"]
    print("%s: B=%d; eval kernels %.3f ms" % (args.workload, len(ks), st["last_eval_kernel_ms"]))
    ncls = len(ks) - sum(int(outc[8 * b_ + 4]) for b_ in range(16))
    if sum(out):
        print("  classic kernel (full bin), %d circuits:" % ncls)
        for n_, v in zip(names, out):
            print("    %-10s %9.0f cycles/circuit  %5.1f%%" % (n_, v / max(1, ncls), 100.0 * v / sum(out)))
        print("    %-10s %9.0f cycles/circuit" % ("total", sum(out) / max(1, ncls)))
    for b_ in range(16):
        cnt = int(outc[8 * b_ + 4])
        if not cnt:
            continue
        v4 = [int(outc[8 * b_ + i]) for i in range(4)]
        print("  compressed kernel, slot %d (%d circuits, %.1f%%):" % (dmin.value + b_, cnt, 100.0 * cnt / len(ks)))
        for n_, v in zip(names, v4):
            print("    %-10s %9.0f cycles/circuit  %5.1f%%" % (n_, v / cnt, 100.0 * v / max(1, sum(v4))))
        print("    %-10s %9.0f cycles/circuit" % ("total", sum(v4) / cnt))


if __name__ == "__main__":
    main()
