"""CPU tier: the C-ABI library loads and exports its symbols, the host-side tape compiler and
the kernel algorithm (NumPy tape interpreter) agree with the oracle, and the host mirror of the
reference plugin surface (pool, indices, legality) behaves like the reference's."""
import ctypes
import re
import os

import numpy as np
import pytest

import helpers
import tape_interp as TI
from oracle import sim
from qas_b200 import _lib
from qas_b200.evaluator import compile_tape_host
from qas_b200.mcts.qml_mcts_state import QMLStateBasicGates
from qas_b200.qml_models.qml_gate_ops import QMLGate, QMLPool, SUPPORTED_OPS_DICT
from qas_b200.qml_models.utils import extractParamIndicesQML
from qas_b200.sampling import legal_mask, sample_structure_lists


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    assert lib.qas_abi_version() == 1
    header = open(os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "..", "include", "qas_b200.h")).read()
    declared = set(re.findall(r"\b(qas_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert declared == set(_lib.EXPORTS)
    header_sv = open(os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "..", "include", "qas_b200_sv.h")).read()
    declared_sv = set(re.findall(r"\b(qas_sv_[a-z0-9_]+)\s*\(", header_sv))
    for sym in declared_sv:
        assert hasattr(lib, sym), sym
    assert declared_sv == set(_lib.SV_EXPORTS)
    from qas_b200.mcts.native import MCTS_EXPORTS
    header_m = open(os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "..", "include", "qas_b200_mcts.h")).read()
    declared_m = set(re.findall(r"\b(qas_mcts_[a-z0-9_]+)\s*\(", header_m))
    for sym in declared_m:
        assert hasattr(lib, sym), sym
    assert declared_m == set(MCTS_EXPORTS)


def test_gate_id_table_matches_header():
    header = open(os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "..", "include", "qas_b200.h")).read()
    body = header.split("enum qas_gate_id {")[1].split("}")[0]
    names = [x.strip().split("=")[0].strip() for x in body.replace("\n", " ").split(",") if x.strip()]
    names = [x for x in names if x != "QAS_G_COUNT"]
    ours = sorted(set(_lib.GATE_IDS.values()))
    assert ours == list(range(len(names)))
    assert names[_lib.GATE_IDS["Rot"]] == "QAS_G_ROT"
    assert names[_lib.GATE_IDS["CNOT"]] == "QAS_G_CNOT"
    assert names[_lib.GATE_IDS["IsingZZ"]] == "QAS_G_ISINGZZ"


def test_pool_key_layout():
    # qml_gate_ops.py:167-176: key 2i = Rot on i, 2i+1 = PlaceHolder on i, 2n+j = CNOT on coupling[j]
    pool = helpers.near_cx_pool(4)
    assert len(pool) == 14
    assert pool[0] == {"Rot": [0]} and pool[1] == {"PlaceHolder": [0]} and pool[6] == {"Rot": [3]}
    assert pool[8] == {"CNOT": [0, 1]} and pool[13] == {"CNOT": [3, 2]}
    full = QMLPool(3, ["Rot"], ["CNOT"])
    assert full.coupling == [(0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1)]
    with pytest.raises(AssertionError):
        QMLPool(3, ["CNOT"], [])
    with pytest.raises(AssertionError):
        QMLPool(3, ["MultiRZ"], [])
    with pytest.raises(AssertionError):
        QMLPool(3, ["Rot"], ["CNOT"], True, [[0, 1]])


def test_extract_param_indices_and_gate_record():
    pool = helpers.near_cx_pool(4)
    assert extractParamIndicesQML([0, 8, 1, 2], pool) == [(0, 0, 0), (0, 0, 1), (0, 0, 2), (3, 2, 0), (3, 2, 1), (3, 2, 2)]
    with pytest.raises(AssertionError):
        extractParamIndicesQML([0, 14], pool)
    g = QMLGate("Rot", [2], [0.1, 0.2, 0.3])
    assert g.getOp().name == "Rot" and g.getQregPos() == [2] and g.num_params == 3
    with pytest.raises(AssertionError):
        QMLGate("Rot", [2], [0.1])
    assert SUPPORTED_OPS_DICT["PlaceHolder"].num_params == 0


@pytest.mark.parametrize("g", helpers.goldens(), ids=lambda g: g["fixture"])
def test_tape_compiler_and_kernel_algorithm_on_goldens(g):
    ib = g["init"][1] if g["init"][0] == "basis" else 0
    ops, idx = compile_tape_host(g["n_qubits"], g["pool"], g["k"], ib)
    # CNOTs and PlaceHolders are folded away: one op per remaining gate
    assert len(ops) == sum(1 for name, _, _ in g["op_list"] if name != "CNOT")
    e, gr = TI.run_tape(g["n_qubits"], ops, g["init"], idx, g["pool"], len(g["pool"]), g["params"], g["terms"])
    assert abs(e - sim.energy(g["n_qubits"], g["k"], g["pool"], g["params"], g["terms"], g["init"])) < 1e-11
    ga = sim.grad_adjoint(g["n_qubits"], g["k"], g["pool"], g["params"], g["terms"], g["init"])
    for (d, j), v in gr.items():
        assert abs(v - ga[d, g["k"][d], j]) < 1e-10


def test_tape_compiler_full_vocabulary_random():
    rng = np.random.default_rng(0)
    n = 4
    pool = helpers.full_vocab_pool(n)
    c = len(pool)
    terms = [(0.3, "XYZI"), (0.2, "IYII"), (-0.15, "XZYY"), (0.5, "ZZII"), (0.1, "IIII")]
    for trial in range(12):
        p = 10
        k = [int(x) for x in rng.integers(0, c, size=p)]
        params = rng.standard_normal((p, c, 3))
        init = [("zero", 0), ("basis", int(rng.integers(0, 16))), ("plus", 0)][trial % 3]
        ops, idx = compile_tape_host(n, pool, k, init[1] if init[0] == "basis" else 0)
        e, gr = TI.run_tape(n, ops, init, idx, pool.pool, c, params, terms)
        assert abs(e - sim.energy(n, k, pool.pool, params, terms, init)) < 1e-11
        ga = sim.grad_adjoint(n, k, pool.pool, params, terms, init)
        got = np.zeros_like(ga)
        for (d, j), v in gr.items():
            got[d, k[d], j] = v
        assert np.max(np.abs(got - ga)) < 1e-10


@pytest.mark.parametrize("gmax", [1, 2, 3])
def test_pass_planner_fused_support_restricted_passes_match_oracle(gmax):
    """Register-fused, support-restricted passes (qas_plan_groups + the kernel's coset enumeration,
    restated in tape_interp.run_tape_planned) give the oracle's energies and gradients; every group
    tiles the index space exactly once and nothing ever lives outside the tracked support."""
    rng = np.random.default_rng(5)
    fused_any, sparse_any = 0, 0
    cases = [(5, helpers.near_cx_pool(5, helpers.ring(5)), 14, ("zero", 0)),
             (6, helpers.near_cx_pool(6), 16, ("basis", 0b101100)),
             (7, helpers.near_cx_pool(7), 12, ("zero", 0)),
             (6, helpers.near_cx_pool(6), 10, ("plus", 0)),
             (4, helpers.full_vocab_pool(4), 10, ("basis", 9)),
             (4, helpers.full_vocab_pool(4), 10, ("zero", 0))]
    for trial, (n, pool, p, init) in enumerate(cases):
        c = len(pool)
        if trial < 4:
            k = [int(x) for x in sample_structure_lists(pool, p, {"CNOT": p // 2}, 1, seed=trial)[0]]
        else:
            k = [int(x) for x in rng.integers(0, c, size=p)]
        params = rng.standard_normal((p, c, 3))
        words = ["".join(rng.choice(list("IXYZ"), size=n)) for _ in range(6)]
        terms = [(float(rng.standard_normal()), w) for w in words]
        ops, idx = compile_tape_host(n, pool, k, init[1] if init[0] == "basis" else 0)
        for swizzled in (True, False):
            planned, ng, gds = TI.plan_passes(n, ops, gmax, init, idx, swizzled)
            assert ng <= len(ops)
            fused_any += len(ops) - ng
            sparse_any += sum(1 for d in gds if int(d[1]) < n - 3)
            # planning never touches parity masks, tables or depths; xor masks of 2x2 ops move to
            # slot coordinates in a swizzled plan
            assert np.array_equal(planned[:, 1:4], ops[:, 1:4])
            is_u1 = ((ops[:, 4] >> 16) & 0xF) == 0
            want_m = np.array([TI._swz(int(m)) if (swizzled and u) else int(m) for m, u in zip(ops[:, 0], is_u1)], dtype=np.uint32)
            assert np.array_equal(planned[:, 0], want_m)
            assert np.array_equal(planned[:, 4] & 0xFFFFFF, ops[:, 4])
            e, gr = TI.run_tape_planned(n, planned, gds, init, idx, pool.pool, c, params, terms, swizzled=swizzled)
            assert abs(e - sim.energy(n, k, pool.pool, params, terms, init)) < 1e-11
            ga = sim.grad_adjoint(n, k, pool.pool, params, terms, init)
            got = np.zeros_like(ga)
            for (d, j), v in gr.items():
                got[d, k[d], j] = v
            assert np.max(np.abs(got - ga)) < 1e-10
    assert (fused_any > 0 or gmax == 1) and sparse_any > 0


def test_tape_compiler_rejects_out_of_range_keys():
    pool = helpers.near_cx_pool(4)
    with pytest.raises(AssertionError):
        compile_tape_host(4, pool, [0, 14, 2])
    with pytest.raises(AssertionError):
        compile_tape_host(4, pool, [0, -1, 2])


def test_all_placeholder_circuit_has_empty_tape():
    pool = helpers.near_cx_pool(3)
    ops, _ = compile_tape_host(3, pool, [1, 3, 5, 1])
    assert len(ops) == 0


def test_sampler_only_draws_legal_actions_and_matches_state_class():
    pool = helpers.near_cx_pool(5, helpers.ring(5))
    p = 12
    lim = {"CNOT": 4}
    ks = sample_structure_lists(pool, p, lim, 64, seed=1)
    assert ks.shape == (64, p) and ks.dtype == np.int32
    for k in ks:
        st = QMLStateBasicGates([], pool, p, set(), lim)
        for a in k:
            assert int(a) in st.getLegalActions()
            st = st.takeAction(int(a))
        assert st.isTerminal() and st.getCurrK() == [int(a) for a in k]
        assert sum(1 for a in k if "CNOT" in pool[int(a)]) <= 4
    # first action is necessarily a parameterised 1-q gate
    assert all("Rot" in pool[int(k[0])] for k in ks)


def test_legality_rules_clifford_t():
    pool = QMLPool(2, ["T", "Tdg", "S", "Sdg", "PauliZ", "Hadamard", "PlaceHolder"], ["CNOT", "CZ"])
    names = {k: list(v.keys())[0] for k, v in pool.pool.items()}
    key = lambda nm, w: [k for k, v in pool.pool.items() if list(v) == [nm] and list(v[nm]) == w][0]  # noqa: E731
    st = QMLStateBasicGates([], pool, 6, set(), None).takeAction(key("T", [0]))
    legal = st.getLegalActions()
    assert key("Tdg", [0]) not in legal            # T then Tdg cancels
    assert key("T", [0]) not in legal              # two T = S and S is in the pool
    assert key("S", [0]) in legal
    assert key("CNOT", [1, 0]) not in legal        # control qubit 1 untouched
    assert key("CNOT", [0, 1]) in legal
    assert key("PlaceHolder", [1]) not in legal and key("PlaceHolder", [0]) in legal
    st2 = st.takeAction(key("CNOT", [0, 1]))
    assert key("CNOT", [0, 1]) not in st2.getLegalActions()
    assert names[key("CZ", [0, 1])] == "CZ" and key("CZ", [0, 1]) in st2.getLegalActions()
    st3 = st2.takeAction(key("Hadamard", [1]))
    assert key("Hadamard", [1]) not in st3.getLegalActions()
    assert key("CNOT", [0, 1]) in st3.getLegalActions()   # last op on qubit 1 is no longer the CNOT


def _par(x):
    return bin(int(x)).count("1") & 1


def test_final_support_descriptor_of_the_h_application():
    """qas_plan_hsupport_host (the planner code behind the support-restricted H application): the
    descriptor's vectors form a basis of the quotient space adapted to Dbar, the check masks read the
    Dbar-coset id, and the oracle's final state has no amplitude outside ubar(b0) ^ Dbar."""
    import ctypes
    from qas_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)
    restricted = 0
    for trial in range(40):
        n = [8, 9, 10][trial % 3]
        full_vocab = trial % 5 == 4          # controlled rotations, Ising gates, swaps ... as well
        pool = helpers.full_vocab_pool(n) if full_vocab else helpers.near_cx_pool(n)
        p = int(rng.integers(3, 7)) if full_vocab else int(rng.integers(4, 14))
        init = ("zero", 0) if trial % 4 else ("basis", int(rng.integers(0, 1 << n)))
        if full_vocab:
            k = [int(x) for x in rng.integers(0, len(pool), size=p)]
        else:
            k = [int(x) for x in sample_structure_lists(pool, p, {"CNOT": p // 2}, 1, seed=100 + trial)[0]]
        ops, idx = compile_tape_host(n, pool, k, init[1] if init[0] == "basis" else 0)
        # a register subspace W in reduced echelon form (highest-bit pivots ascending)
        rb_count = int(rng.integers(0, 4))
        W = []
        while len(W) < rb_count:
            x = int(rng.integers(1, 1 << n))
            for pv, v in W:
                if (x >> pv) & 1:
                    x ^= v
            if not x:
                continue
            pv = x.bit_length() - 1
            W = [(q, v ^ x if (v >> pv) & 1 else v) for q, v in W] + [(pv, x)]
        W.sort()
        piv = np.array([pv for pv, _ in W] + [0] * (3 - len(W)), dtype=np.int32)
        wv = np.array([v for _, v in W] + [0] * (3 - len(W)), dtype=np.uint32)
        hd = np.zeros(32, dtype=np.uint32)
        opsc = np.ascontiguousarray(ops, dtype=np.uint32)
        rc = lib.qas_plan_hsupport_host(n, opsc.ctypes.data, len(opsc), 1 if init[0] == "basis" else 0, int(idx),
                                        rb_count, piv.ctypes.data, wv.ctypes.data, hd.ctypes.data, 32)
        assert rc == 0
        nu = n - rb_count
        dbar = int(hd[1])
        assert 0 <= dbar <= nu

        def ubar(x):
            for pv, v in W:
                if (x >> pv) & 1:
                    x ^= v
            for pv, _ in reversed(W):
                x = ((x >> (pv + 1)) << pv) | (x & ((1 << pv) - 1))
            return x

        st = sim.run(n, sim.lower(k, pool.pool, rng.standard_normal((p, len(pool), 3))), init).reshape(-1)
        support = [i for i in range(1 << n) if abs(st[i]) > 1e-14]
        if dbar == nu:
            continue          # dense: nothing claimed
        restricted += 1
        X = [int(x) for x in hd[2:2 + nu]]
        chk = [int(x) for x in hd[2 + nu:2 + 2 * nu - dbar]]
        # the X's are a basis: 2^nu distinct xor-combinations
        span = {0}
        for x in X:
            span |= {s ^ x for s in span}
        assert len(span) == 1 << nu
        # Dbar part only touches bits <= its pivot, pivots ascending
        pivs = [x.bit_length() - 1 for x in X[:dbar]]
        assert pivs == sorted(pivs) and len(set(pivs)) == dbar
        # check masks: zero on Dbar, delta on the completion vectors
        for i, c_ in enumerate(chk):
            assert all(_par(c_ & x) == 0 for x in X[:dbar])
            assert [_par(c_ & x) for x in X[dbar:]] == [int(j == i) for j in range(nu - dbar)]
        # every amplitude of the final state sits on a coset u with u ^ ubar(b0) in Dbar
        dspan = {0}
        for x in X[:dbar]:
            dspan |= {s ^ x for s in dspan}
        assert int(hd[0]) == ubar(int(idx))
        assert all((ubar(i) ^ int(hd[0])) in dspan for i in support)
    assert restricted >= 10


# ------------------------------------------------------------------------------------------------
# support-compressed tapes (tape.h qas_compress_tape; kernel algorithm restated in tests/ctape_interp.py)
def _ctape_case(n, pool, p, lim, terms, seeds, k_override=None):
    import ctape_interp as CT
    c = len(pool)
    xs, imag, V = CT.sign_tables(n, terms)
    rng = np.random.default_rng(17)
    stats = {"compressed": 0, "full": 0, "dims": set(), "dropped_slices": 0}
    for seed in seeds:
        if k_override is None:
            k = [int(x) for x in sample_structure_lists(pool, p, lim, 1, seed=seed)[0]]
        else:
            k = k_override(rng)
        params = rng.standard_normal((p, c, 3))
        ct = CT.compile_ctape(n, pool, k, xs, gmax=2)
        if ct["d"] < 0:
            stats["full"] += 1
            # the classic tape comes back grouped and otherwise untouched
            ops, _ = compile_tape_host(n, pool, k, 0)
            assert np.array_equal(ct["ops"][:, :4], ops[:, :4])
            assert np.array_equal(ct["ops"][:, 4] & 0xFFFFFF, ops[:, 4])
            continue
        stats["compressed"] += 1
        stats["dims"].add(ct["d"])
        stats["dropped_slices"] += len(xs) - len(ct["slices"])
        assert 0 <= ct["d"] < n
        e, gr = CT.run_ctape(ct, n, pool.pool, c, params, xs, imag, V)
        assert abs(e - sim.energy(n, k, pool.pool, params, terms, ("zero", 0))) < 1e-11
        ga = sim.grad_adjoint(n, k, pool.pool, params, terms, ("zero", 0))
        got = np.zeros_like(ga)
        for (dd, j), v in gr.items():
            got[dd, k[dd], j] = v
        assert np.max(np.abs(got - ga)) < 1e-10
        # one-gate adjoint passes (evalc_kernel<..., AG = 1>): the gates of a two-gate group un-applied one at a time, the
        # second inside its own (smaller or equal) support prefix -- same cross matrices
        e1, gr1 = CT.run_ctape(ct, n, pool.pool, c, params, xs, imag, V, adjoint_gates=1)
        assert e1 == e and gr1.keys() == gr.keys()
        assert max([abs(gr1[key] - gr[key]) for key in gr] + [0.0]) < 1e-13
    return stats


def test_ctape_lih_and_h2o_batches_match_oracle():
    """c-tapes of LiH-10q / H2O-8q near_cx circuits: the compressed register evaluated with the kernel's
    algorithm gives the oracle's energy and every derivative; X masks outside the support span drop out."""
    from qas_b200.qml_models.qml_models_legacy import H2O, LiH
    st = _ctape_case(10, helpers.near_cx_pool(10), 20, {"CNOT": 10}, LiH.terms(), range(10))
    assert st["compressed"] >= 8 and len(st["dims"]) >= 3 and st["dropped_slices"] > 0
    st = _ctape_case(8, helpers.near_cx_pool(8), 50, {"CNOT": 25}, H2O.terms(), range(6))
    assert st["compressed"] + st["full"] == 6


def test_ctape_full_vocabulary_random_circuits():
    """Every gate of the vocabulary through the c-tape path: controlled gates whose control never fires on the
    support are dropped, parity-diagonal ops act on the prefix, fused groups use run-time projection."""
    n = 6
    pool = helpers.full_vocab_pool(n)
    c = len(pool)
    rng0 = np.random.default_rng(3)
    words = ["".join(rng0.choice(list("IXYZ"), size=n)) for _ in range(10)] + ["ZZIIII", "IIIIII", "XXIIII", "YYIIII"]
    terms = [(float(rng0.standard_normal()), w) for w in words]
    st = _ctape_case(n, pool, 9, None, terms, range(40), k_override=lambda rng: [int(x) for x in rng.integers(0, c, size=9)])
    assert st["compressed"] >= 10 and len(st["dims"]) >= 2


def test_ctape_edge_cases():
    import ctape_interp as CT
    n = 7
    pool = helpers.near_cx_pool(n)
    terms = [(0.7, "ZIIIIII"), (0.2, "XXIIIII"), (-0.4, "IIIIIII")]
    xs, imag, V = CT.sign_tables(n, terms)
    # all PlaceHolders: d = 0, one amplitude, only the diagonal slice survives
    k = [1] * 5
    ct = CT.compile_ctape(n, pool, k, xs)
    assert ct["d"] == 0 and len(ct["ops"]) == 0 and [s for _, s in ct["slices"]] == [0]
    e, gr = CT.run_ctape(ct, n, pool.pool, len(pool), np.zeros((5, len(pool), 3)), xs, imag, V)
    assert abs(e - 0.3) < 1e-15 and gr == {}
    # max_d below the circuit's span: not compressed
    k = [0, 2, 4, 6, 8]
    assert CT.compile_ctape(n, pool, k, xs, max_d=3)["d"] == CT.CT_TOO_BIG
    assert CT.compile_ctape(n, pool, k, xs, max_d=5)["d"] == 5


def test_ctape_device_formulation_matches_reference_formulation():
    """The c-tape compiler as the tape-compile kernel runs it (packed ops, reduced echelon basis in registers,
    one-shot reductions) produces bit-identical records to the reference formulation, on chemistry batches and on
    the full vocabulary, compressed and not."""
    import ctape_interp as CT
    from qas_b200.qml_models.qml_models_legacy import LiH
    rng = np.random.default_rng(9)
    cases = []
    pool10 = helpers.near_cx_pool(10)
    xs10, _, _ = CT.sign_tables(10, LiH.terms()[:60])
    for k in sample_structure_lists(pool10, 20, {"CNOT": 10}, 60, seed=4):
        cases.append((10, pool10, [int(x) for x in k], xs10, 12))
    pool7 = helpers.full_vocab_pool(7)
    xs7 = sorted({int(x) for x in rng.integers(0, 128, size=20)} | {0})
    for _ in range(60):
        cases.append((7, pool7, [int(x) for x in rng.integers(0, len(pool7), size=8)], xs7, 12))
    pool14 = helpers.near_cx_pool(14)
    xs14 = sorted({int(x) for x in rng.integers(0, 1 << 14, size=40)} | {0, 3, 5 << 7})
    for k in sample_structure_lists(pool14, 24, {"CNOT": 12}, 30, seed=5):
        cases.append((14, pool14, [int(x) for x in k], xs14, 9))
    seen = set()
    for n, pool, k, xs, max_d in cases:
        a = CT.compile_ctape(n, pool, k, xs, gmax=2, max_d=max_d, variant=0)
        b = CT.compile_ctape(n, pool, k, xs, gmax=2, max_d=max_d, variant=1)
        seen.add(a["d"] if a["d"] >= 0 else a["d"])
        assert a["d"] == b["d"] and a["ngroups"] == b["ngroups"]
        assert np.array_equal(a["ops"], b["ops"])
        assert a["basis"] == b["basis"] and a["slices"] == b["slices"]
    assert CT.CT_TOO_BIG in seen and len([d for d in seen if d >= 0]) >= 5


# ------------------------------------------------------------------------------------------------
# property tests (hypothesis): random registers, pools, depths and structure lists
from hypothesis import HealthCheck, given, settings, strategies as st_


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(n=st_.integers(3, 8), p=st_.integers(1, 14), seed=st_.integers(0, 10 ** 6), vocab=st_.booleans())
def test_property_tape_paths_agree_with_oracle(n, p, seed, vocab):
    """For any register size, depth and (legal or arbitrary) structure list from the all-zero state: classic tape +
    kernel algorithm == c-tape (both formulations, bit-identical records) + compressed kernel algorithm == oracle."""
    import ctape_interp as CT
    rng = np.random.default_rng(seed)
    pool = helpers.full_vocab_pool(n) if (vocab and n <= 5) else helpers.near_cx_pool(n)
    c = len(pool)
    if vocab and n <= 5:
        k = [int(x) for x in rng.integers(0, c, size=p)]
    else:
        k = [int(x) for x in sample_structure_lists(pool, p, {"CNOT": max(1, p // 2)}, 1, seed=seed)[0]]
    params = rng.standard_normal((p, c, 3))
    words = ["".join(rng.choice(list("IXYZ"), size=n)) for _ in range(5)] + ["I" * n]
    terms = [(float(rng.standard_normal()), w) for w in words]
    e_ref = sim.energy(n, k, pool.pool, params, terms, ("zero", 0))
    g_ref = sim.grad_adjoint(n, k, pool.pool, params, terms, ("zero", 0))
    ops, idx = compile_tape_host(n, pool, k, 0)
    e1, gr1 = TI.run_tape(n, ops, ("zero", 0), idx, pool.pool, c, params, terms)
    assert abs(e1 - e_ref) < 1e-10
    xs, imag, V = CT.sign_tables(n, terms)
    a = CT.compile_ctape(n, pool, k, xs, variant=0)
    b = CT.compile_ctape(n, pool, k, xs, variant=1)
    assert a["d"] == b["d"] and np.array_equal(a["ops"], b["ops"]) and a["slices"] == b["slices"] and a["basis"] == b["basis"]
    if a["d"] >= 0:
        e2, gr2 = CT.run_ctape(a, n, pool.pool, c, params, xs, imag, V)
        assert abs(e2 - e_ref) < 1e-10
        got = np.zeros_like(g_ref)
        for (dd, j), v in gr2.items():
            got[dd, k[dd], j] = v
        assert np.max(np.abs(got - g_ref)) < 1e-9


def test_vectorised_key_range_check_matches_the_scalar_rule():
    """qas_check_keys_host = the passes qas_eval_batch runs on large host batches (16 keys per instruction): the
    reference's asserts min(k) >= 0, max(k) < c (utils.py:12-15) for byte and int32 keys, and the int32 -> byte
    narrowing of pageable lists; every length 0..130 (vector bodies and tails), bad keys at random places."""
    import ctypes
    from qas_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(2)
    for trial in range(600):
        nk = int(rng.integers(0, 131)) if trial % 3 else int(rng.integers(1000, 5000))
        c = int(rng.integers(1, 257))
        k = rng.integers(0, c, size=nk).astype(np.int32)
        kind = trial % 4
        if nk and kind == 1:
            k[int(rng.integers(0, nk))] = c + int(rng.integers(0, 3))
        if nk and kind == 2:
            k[int(rng.integers(0, nk))] = -1 - int(rng.integers(0, 3))
        want = bool(np.any((k < 0) | (k >= c)))
        flag = ctypes.c_int(-1)
        kk = np.ascontiguousarray(k) if nk else np.zeros(1, dtype=np.int32)
        assert lib.qas_check_keys_host(kk.ctypes.data, nk, 4, c, None, ctypes.byref(flag)) == 0
        assert bool(flag.value) == want
        out = np.full(nk + 1, 0xAA, dtype=np.uint8)
        assert lib.qas_check_keys_host(kk.ctypes.data, nk, 4, c, out.ctypes.data, ctypes.byref(flag)) == 0
        assert bool(flag.value) == want and out[nk] == 0xAA
        if not want:
            assert np.array_equal(out[:nk], k.astype(np.uint8))
        k8 = (k & 0xFF).astype(np.uint8)
        kk8 = np.ascontiguousarray(k8) if nk else np.zeros(1, dtype=np.uint8)
        assert lib.qas_check_keys_host(kk8.ctypes.data, nk, 1, c, None, ctypes.byref(flag)) == 0
        assert bool(flag.value) == bool(np.any(k8 >= c))
    assert lib.qas_check_keys_host(None, 0, 4, 4, None, ctypes.byref(flag)) != 0


def test_planner_output_is_pinned_word_by_word():
    """Compiled tapes, group marks, group descriptors and final-support descriptors of 48 seeded circuits equal the
    committed fixture (tools/make_planner_golden.py): the planner was rewritten on packed pivot tables with bit-identical
    output, and every kernel-side enumeration depends on these words."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_planner_golden", os.path.join(root, "tools", "make_planner_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = np.load(os.path.join(root, "tests", "golden", "planner_descriptors.npz"))["words"]
    got = mod.planner_outputs()
    assert got.shape == want.shape and np.array_equal(got, want)

