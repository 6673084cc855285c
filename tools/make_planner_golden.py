"""Fixture of the classic pass planner's output (tests/golden/planner_descriptors.npz): compiled tapes, group marks and
group descriptors (qas_compile_tape_host / qas_plan_passes_host) and final-support descriptors (qas_plan_hsupport_host) of
48 seeded circuits, written by the library as it was when the packed pivot tables replaced the unrolled ones -- their
outputs had been compared word by word on 400 random tapes.  tests/test_host.py regenerates the same
arrays and demands equality, so a later change of the planner that alters a single descriptor word shows up on the CPU.
    python tools/make_planner_golden.py            # rewrites the fixture
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def planner_outputs():
    import helpers
    from tape_interp import plan_passes
    from qas_b200 import _lib
    from qas_b200.evaluator import compile_tape_host
    from qas_b200.sampling import sample_structure_lists
    lib = _lib.load()
    rng = np.random.default_rng(20261020)
    words = []
    for trial in range(48):
        n = [4, 6, 7, 8, 8, 10, 12, 14][trial % 8]
        full_vocab = trial % 6 == 5
        pool = helpers.full_vocab_pool(n) if full_vocab else helpers.near_cx_pool(n)
        p = int(rng.integers(3, 8)) if full_vocab else int(rng.integers(4, 30))
        kind = ["zero", "basis", "plus"][trial % 3]
        init = (kind, int(rng.integers(0, 1 << n)) if kind == "basis" else 0)
        if full_vocab:
            k = [int(x) for x in rng.integers(0, len(pool), size=p)]
        else:
            k = [int(x) for x in sample_structure_lists(pool, p, {"CNOT": p // 2}, 1, seed=7000 + trial)[0]]
        ops, idx = compile_tape_host(n, pool, k, init[1] if kind == "basis" else 0)
        words += [np.array([n, len(ops), idx], dtype=np.uint32), ops.astype(np.uint32).ravel()]
        for gmax in (1, 2):
            for swz in (0, 1):
                o2, ng, gd = plan_passes(n, ops, gmax, init, idx, bool(swz))
                words += [np.array([ng], dtype=np.uint32), o2.astype(np.uint32).ravel(), gd.astype(np.uint32).ravel()]
        if kind != "plus":
            piv = np.array([0, 2, 0], dtype=np.int32)
            wv = np.array([1, 4, 0], dtype=np.uint32)
            hd = np.zeros(32, dtype=np.uint32)
            opsc = np.ascontiguousarray(ops, dtype=np.uint32)
            rc = lib.qas_plan_hsupport_host(n, opsc.ctypes.data, len(opsc), 1 if kind == "basis" else 0, int(idx), 2,
                                            piv.ctypes.data, wv.ctypes.data, hd.ctypes.data, 32)
            words += [np.array([rc & 0xffffffff], dtype=np.uint32), hd]
    return np.concatenate(words)


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden", "planner_descriptors.npz")
    w = planner_outputs()
    np.savez_compressed(out, words=w)
    print("wrote %s: %d words, %d bytes" % (out, len(w), os.path.getsize(out)))
