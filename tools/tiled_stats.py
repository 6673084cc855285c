"""Tile-pass statistics of a workload on the CPU (host planner) and, with --gpu, the per-phase clocks of the tiled
kernel (QAS_B200_PHASE_CLOCKS=1) on one batch.   python tools/tiled_stats.py --workload kagome_16q [--gpu --batch 2048]"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="kagome_16q")
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--sample", type=int, default=16)
    ap.add_argument("--tb", type=int, default=11)
    ap.add_argument("--clow", type=int, default=2)
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--no-grad", action="store_true")
    args = ap.parse_args()
    if args.gpu:
        os.environ["QAS_B200_PHASE_CLOCKS"] = "1"
    import numpy as np
    import bench
    import __graft_entry__ as ge
    ge.build()
    model, pool, params, ks = bench.make_workload(args.workload, args.batch, 0)
    n = model.num_qubits if hasattr(model, "num_qubits") else pool.num_qubits
    import tape_interp as TI
    import tile_interp as TL
    from qas_b200.evaluator import compile_tape_host
    rows = []
    for b in range(min(args.sample, len(ks))):
        ops, _ = compile_tape_host(n, pool, [int(x) for x in ks[b]], 0)
        planned, ng, _ = TI.plan_passes(n, ops, 2, ("zero", 0), 0, swizzled=False)
        npf = npa = 0
        j = len(planned) - 1
        while j >= 0:
            cnt, _, _ = TL.plan_tile_pass(n, planned, j, -1, args.tb, args.clow)
            j -= cnt
            npf += 1
        j = 0
        while j < len(planned):
            cnt, _, _ = TL.plan_tile_pass(n, planned, j, +1, args.tb, args.clow)
            j += cnt
            npa += 1
        rows.append((len(planned), ng, npf, npa))
    r = np.array(rows, dtype=float)
    print("%s (n = %d, tb = %d, clow = %d), %d circuits: ops %.1f  groups %.1f  forward passes %.1f  adjoint passes %.1f  ops/pass %.1f"
          % (args.workload, n, args.tb, args.clow, len(rows), r[:, 0].mean(), r[:, 1].mean(), r[:, 2].mean(), r[:, 3].mean(),
             r[:, 0].mean() / max(1e-9, r[:, 3].mean())))
    if not args.gpu:
        return
    ev = model.evaluator(pool, 3, 0)
    for rep in range(2):
        ev.evaluate(ks, params, want_grad=not args.no_grad)
        out = (ctypes.c_uint64 * 8)()
        rc = ev._lib.qas_debug_phase_clocks_t(ev._h, out, 1)
        assert rc == 0, ev._lib.qas_last_error(ev._h)
    st = ev.stats()
    v = [int(x) for x in out]
    ncirc = max(1, v[6])
    print("B = %d: eval kernel %.3f ms (%.1f k evals/s), path %d, %d CTAs/SM, %d B smem/CTA" %
          (len(ks), st["last_eval_kernel_ms"], len(ks) / st["last_eval_kernel_ms"], st["path"], st["ctas_per_sm"], st["smem_bytes_per_cta"]))
    for name, x in zip(["prologue", "forward", "H+energy", "adjoint"], v[:4]):
        print("  %-10s %10.0f cycles/circuit  %5.1f%%" % (name, x / ncirc, 100.0 * x / max(1, sum(v[:4]))))
    print("  passes per circuit: forward %.1f, adjoint %.1f; cycles per forward pass %.0f, per adjoint pass %.0f" %
          (v[4] / ncirc, v[5] / ncirc, v[1] / max(1, v[4]), v[3] / max(1, v[5])))


if __name__ == "__main__":
    main()
