"""CPU tier: the tile-pass planner of the tiled kernel for 13..16-qubit registers (qas_plan_tile_pass, tape.h) and
the tiled algorithm itself (tests/tile_interp.py restates eval_tiled_kernel in NumPy) against the oracle."""
import numpy as np
import pytest

import helpers
import tape_interp as TI
import tile_interp as TL
from oracle import sim
from qas_b200.evaluator import compile_tape_host
from qas_b200.sampling import sample_structure_lists


def _rank_gf2(vs):
    vs, r = list(vs), 0
    for bit in range(20):
        piv = next((i for i in range(r, len(vs)) if (vs[i] >> bit) & 1), None)
        if piv is None:
            continue
        vs[r], vs[piv] = vs[piv], vs[r]
        for i in range(len(vs)):
            if i != r and (vs[i] >> bit) & 1:
                vs[i] ^= vs[r]
        r += 1
    return r


@pytest.mark.parametrize("n,tb,clow", [(9, 6, 2), (10, 7, 3), (8, 8, 2)])
def test_tile_pass_planner_frames(n, tb, clow):
    """Every pass gets an invertible frame whose first tb columns span the masks of its ops and the low bits; the
    passes partition the tape in both sweep directions and never split a pass group."""
    pool = helpers.near_cx_pool(n)
    p = 40
    for seed in range(4):
        k = [int(x) for x in sample_structure_lists(pool, p, {"CNOT": p // 2}, 1, seed=seed)[0]]
        ops, _ = compile_tape_host(n, pool, k, 0)
        planned, ng, _ = TI.plan_passes(n, ops, 2, ("zero", 0), 0, swizzled=False)
        for direction in (-1, +1):
            j = len(planned) - 1 if direction < 0 else 0
            taken = 0
            while 0 <= j < len(planned):
                cnt, col, piv = TL.plan_tile_pass(n, planned, j, direction, tb, clow)
                assert cnt >= 1
                assert _rank_gf2(col) == n
                assert col[:clow] == [1 << q for q in range(clow)]
                assert all((cv & ((1 << clow) - 1)) == 0 for cv in col[clow:])
                for q in range(n):
                    assert col[piv[col[q].bit_length() - 1]] == col[q]
                lo = j if direction > 0 else j - cnt + 1
                for i in range(lo, lo + cnt):                      # raises if a mask leaves the tile span
                    TL._translate(planned[i], n, tb, clow, col, piv)
                # whole groups only: the first op of the run starts a group, the last one ends one
                assert (int(planned[lo][4]) >> 24) & 3 >= 1 and (int(planned[lo + cnt - 1][4]) >> 26) & 3 >= 1
                j += direction * cnt
                taken += cnt
            assert taken == len(planned)


@pytest.mark.parametrize("case", range(6))
def test_tiled_algorithm_matches_oracle(case):
    rng = np.random.default_rng(100 + case)
    n, pool, p, init, tb, clow = [
        (8, helpers.near_cx_pool(8), 30, ("zero", 0), 5, 2),
        (7, helpers.near_cx_pool(7, helpers.ring(7)), 24, ("basis", 0b1011001), 5, 2),
        (7, helpers.near_cx_pool(7), 20, ("plus", 0), 4, 2),
        (6, helpers.full_vocab_pool(6), 14, ("basis", 37), 4, 2),
        (6, helpers.full_vocab_pool(6), 14, ("zero", 0), 5, 3),
        (6, helpers.near_cx_pool(6), 16, ("zero", 0), 6, 2)][case]
    c = len(pool)
    if case in (3, 4):
        k = [int(x) for x in rng.integers(0, c, size=p)]
    else:
        k = [int(x) for x in sample_structure_lists(pool, p, {"CNOT": p // 2}, 1, seed=case)[0]]
    params = rng.standard_normal((p, c, 3))
    words = ["".join(rng.choice(list("IXYZ"), size=n)) for _ in range(6)]
    terms = [(float(rng.standard_normal()), w) for w in words]
    ops, idx = compile_tape_host(n, pool, k, init[1] if init[0] == "basis" else 0)
    planned, ng, _ = TI.plan_passes(n, ops, 2, init, idx, swizzled=False)
    e, gr, npf, npa = TL.run_tiled(n, planned, init, idx, pool.pool, c, params, terms, tb, clow)
    assert npf >= 1 and (npa >= 1 or len(planned) == 0)
    assert abs(e - sim.energy(n, k, pool.pool, params, terms, init)) < 1e-11
    ga = sim.grad_adjoint(n, k, pool.pool, params, terms, init)
    got = np.zeros_like(ga)
    for (d, j), v in gr.items():
        got[d, k[d], j] = v
    assert np.max(np.abs(got - ga)) < 1e-10


def test_tiled_empty_tape():
    n, pool = 6, helpers.near_cx_pool(6)
    ph = [i for i in range(len(pool)) if list(pool[i].keys())[0] == "PlaceHolder"]
    k = [ph[0]] * 5
    params = np.zeros((5, len(pool), 3))
    terms = [(0.7, "ZIIIII"), (0.2, "IIIIII")]
    ops, idx = compile_tape_host(n, pool, k, 0)
    e, gr, npf, npa = TL.run_tiled(n, ops.reshape(-1, 5), ("zero", 0), idx, pool.pool, len(pool), params, terms, 4, 2)
    assert abs(e - 0.9) < 1e-14 and not gr and npf == 1 and npa == 0


def test_group_coset_bases_are_conflict_free_where_possible():
    """qas_tile_group_basis: the basis lies in the common kernel of the parity masks, has distinct ascending lowest-set-bit
    pivots, enumerates every coset of span(m) exactly once, and a quarter-warp (eight consecutive cosets) touches eight
    different 16-byte bank groups whenever the kernel has vectors with lowest set bits 0, 1 and 2."""
    rng = np.random.default_rng(11)
    tb = 9
    degs = []
    for trial in range(60):
        g = 1 + trial % 2
        while True:       # random masks with parity(m_a & r_b) = delta_ab
            ms = [int(rng.integers(1, 1 << tb)) for _ in range(g)]
            rs = [int(rng.integers(1, 1 << tb)) for _ in range(g)]
            if all(TL._parity(ms[a] & rs[b]) == (1 if a == b else 0) for a in range(g) for b in range(g)):
                break
        kb = TL.group_basis(g, tb, ms, rs)
        lsb = [(v & -v).bit_length() - 1 for v in kb]
        assert lsb == sorted(set(lsb)) and len(kb) == tb - g
        pp, xm = TL._cosets(g, tb, ms, rs, [bool(trial & 4)] * g)
        cover = np.zeros(1 << tb, dtype=int)
        for x in xm:
            cover[pp ^ x] += 1
        assert (cover == 1).all()
        q = (pp.reshape(-1, 8) & 7)
        deg = max(int((q == b).sum(axis=1).max()) for b in range(8))
        if lsb[:3] == [0, 1, 2]:
            assert deg == 1
        degs.append(deg)
    assert np.mean(degs) < 1.5
