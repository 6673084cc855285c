"""Generate the Pauli-sum data files under qas_b200/data/ (run once, in the build container).

The reference builds its Hamiltonians at import time with PySCF/OpenFermion/PennyLane
(qas/qml_models/qml_models_legacy.py:1237-1256, qas/qml_models/qchem_model.py:18-20),
none of which is installed here, and nothing under /root/reference exists on the GPU box.
This script therefore derives what can be derived from files the reference ships and
writes small JSON tables the product and the tests load:

  h2_4q            4-qubit H2 / sto-3g, the standard PennyLane Pauli sum;
                   min eigenvalue checked against plot_results.py:48
  lih_sto6g_10q    10-qubit LiH: rebuilt from results_and_plots-legacy/molecule_pyscf_sto-6g.hdf5
                   (core=[0], active=[1..5], Jordan-Wigner), SURVEY.md A.2
  lih_sto3g_12q    12-qubit LiH: full space of molecule_pyscf_sto-3g.hdf5 (stands in for the
                   downloaded qchem dataset of qchem_model.py:19; same geometry class)
  kagome_16q       kagome_heisenberg_model.py:20-52 (edge list is in-tree)
  qaoa_7q, qaoa_weighted_5q   graphs at qml_models_legacy.py:2347, :2514
  h2o_sto6g_8q     8-qubit H2O of qml_models_legacy.py:1245-1256: restricted Hartree-Fock over STO-6G
                   with tools/sto_ng.py (McMurchie-Davidson integrals), core = 3 orbitals, active =
                   4 orbitals / 4 electrons, Jordan-Wigner.  Pinned: the integral code reproduces the
                   SCF and FCI energies quoted at search_H2O_near_cx.py:45-49 to 3e-10, the shipped
                   LiH integrals to 3e-8, and the operator reproduces the last loss of the saved H2O
                   search result (20220504-074905_H2O_*.json) to 3e-6.
  h2_sto3g_4q      H2 / STO-3G at 0.742 A (qchem_model.py:18; the dataset itself is a download)
  h2o_sto3g_14q    H2O / STO-3G, r(OH) = 0.958 A (qchem_model.py:20), full space, 14 qubits; the
                   H-O-H angle of the dataset is not stated in the reference: 104.5 deg assumed
  h2_2q_standin    PARITY-UNPINNED stand-in: the textbook 2-qubit reduced H2 operator (the class
                   TwoQubitH2 is missing from the reference tree).

Usage:  python tools/make_hamiltonians.py [/root/reference]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fermion as F  # noqa: E402
import sto_ng as G  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "qas_b200", "data")

H2_4Q = [
    (-0.04207897647782276, "IIII"), (0.17771287465139946, "ZIII"), (0.1777128746513994, "IZII"),
    (-0.24274280513140462, "IIZI"), (-0.24274280513140462, "IIIZ"), (0.12293305056183798, "ZIZI"),
    (0.12293305056183798, "IZIZ"), (0.1676831945771896, "ZIIZ"), (0.1676831945771896, "IZZI"),
    (0.17059738328801052, "ZZII"), (0.17627640804319591, "IIZZ"), (-0.04475014401535161, "YYXX"),
    (-0.04475014401535161, "XXYY"), (0.04475014401535161, "YXXY"), (0.04475014401535161, "XYYX"),
]

KAGOME_EDGES = [(1, 2), (2, 3), (3, 5), (5, 8), (8, 11), (11, 14), (14, 13), (13, 12), (12, 10),
                (10, 7), (7, 4), (4, 1), (4, 2), (2, 5), (5, 11), (11, 13), (13, 10), (10, 4)]
QAOA_7Q = [(0, 1, 1.0), (0, 2, 1.0), (2, 3, 1.0), (1, 4, 1.0), (2, 4, 1.0), (0, 5, 1.0), (3, 6, 1.0),
           (1, 6, 1.0)]
QAOA_4Q = [(0, 1, 1.0), (0, 3, 1.0), (1, 2, 1.0), (2, 3, 1.0)]   # PennyLane max-cut tutorial ring
QAOA_W5Q = [(0, 1, 1.0), (0, 2, 2.0), (2, 3, 3.0), (1, 4, 4.0), (2, 4, 5.0), (0, 4, 6.0)]


def word(n, ops):
    w = ["I"] * n
    for q, ch in ops:
        w[q] = ch
    return "".join(w)


def merge(terms):
    acc = {}
    order = []
    for c, w in terms:
        if w not in acc:
            order.append(w)
            acc[w] = 0.0
        acc[w] += c
    return [(acc[w], w) for w in order]


def kagome():
    t = []
    for a, b in KAGOME_EDGES:
        for ch in "XYZ":
            t.append((1.0, word(16, [(a, ch), (b, ch)])))
    return t


def maxcut(n, graph):
    """loss = -sum_e w/2 (1 - Z_a Z_b)  (qml_models_legacy.py:2390-2393, :2557-2561)."""
    t = []
    for a, b, w in graph:
        t.append((-0.5 * w, "I" * n))
        t.append((0.5 * w, word(n, [(a, "Z"), (b, "Z")])))
    return merge(t)


def molecular(one, two, e_nuc, occupied, active):
    const, ob, tb = F.active_space(one, two, occupied, active)
    h1, h2 = F.spinorb_from_spatial(ob, tb)
    acc = F.jordan_wigner(const + e_nuc, h1, h2)
    return F.to_words(h1.shape[0], acc)


H2O_LEGACY_SYMBOLS = ["H", "O", "H"]
# qml_models_legacy.py:1245-1246 (an Angstrom -> Bohr factor applied to the listed numbers)
H2O_LEGACY_BOHR = np.array([0., 0., 0., 1.63234543, 0.86417176, 0., 3.36087791, 0., 0.]) * 1.88973


def hartree_fock_molecular(symbols, coords_bohr, basis, n_electrons, n_core, n_active):
    """RHF -> MO integrals -> active space -> Jordan-Wigner words; also returns E_SCF."""
    S, h, g, e_nuc = G.ao_integrals(symbols, coords_bohr, basis)
    e_scf, C, _ = G.rhf(S, h, g, n_electrons, e_nuc)
    one, two = G.mo_integrals(h, g, C)
    core = list(range(n_core))
    active = list(range(n_core, n_core + n_active))
    return molecular(one, two, e_nuc, core, active), e_scf


def h2o_sto3g_geometry(r_angstrom=0.958, angle_deg=104.5):
    r = r_angstrom * G.BOHR_PER_ANGSTROM
    t = np.deg2rad(angle_deg) / 2
    return ["O", "H", "H"], np.array([0., 0., 0., r * np.sin(t), r * np.cos(t), 0., -r * np.sin(t), r * np.cos(t), 0.])


def dump(name, n, terms, source, **extra):
    os.makedirs(OUT, exist_ok=True)
    d = {"name": name, "n_qubits": n, "source": source, "num_terms": len(terms)}
    d.update(extra)
    d["terms"] = [[float(c), w] for c, w in terms]
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(d, f, indent=0, separators=(",", ":"))
    xs = {tuple(i for i, ch in enumerate(w) if ch in "XY") for _, w in terms}
    print("%-18s n=%2d terms=%5d xmask-groups=%4d" % (name, n, len(terms), len(xs)))


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    dump("h2_4q", 4, H2_4Q, "standard PennyLane H2/sto-3g Pauli sum (SURVEY.md A.1)",
         e_fci_reference=-1.136189454088)
    dump("kagome_16q", 16, kagome(), "qas/qml_models/kagome_heisenberg_model.py:20-52")
    dump("qaoa_7q", 7, maxcut(7, QAOA_7Q), "qas/qml_models/qml_models_legacy.py:2347,2390-2393")
    dump("qaoa_weighted_5q", 5, maxcut(5, QAOA_W5Q), "qas/qml_models/qml_models_legacy.py:2514,2557-2561")
    dump("qaoa_ring_4q", 4, maxcut(4, QAOA_4Q),
         "4-node ring of the PennyLane max-cut tutorial the 4Q result files were produced with "
         "(the 4-qubit model class is no longer in the reference tree; graph inferred from "
         "classical_result = {0101, 1010: 4} in 20220406-*_QAOAVQCDemo_4Q_*.json)")
    dump("h2_2q_standin", 2,
         [(-1.052373245772859 + 0.7199689944489797, "II"), (0.39793742484318045, "IZ"),
          (-0.39793742484318045, "ZI"), (-0.01128010425623538, "ZZ"), (0.18093119978423156, "XX")],
         "STAND-IN (parity unpinned): textbook 2-qubit reduced H2 operator + nuclear repulsion; "
         "the reference class TwoQubitH2 is missing from its tree")

    e_nuc_lih = 3.0 / 2.969280527     # Z_Li*Z_H / R, geometry at qml_models_legacy.py:1242
    p6 = os.path.join(ref, "results_and_plots-legacy", "molecule_pyscf_sto-6g.hdf5")
    one, two = F.read_molecular_integrals(p6)
    dump("lih_sto6g_10q", 10, molecular(one, two, e_nuc_lih, [0], [1, 2, 3, 4, 5]),
         "results_and_plots-legacy/molecule_pyscf_sto-6g.hdf5, core=[0], active=[1..5], JW "
         "(qml_models_legacy.py:1241-1243)")
    p3 = os.path.join(ref, "results_and_plots-legacy", "molecule_pyscf_sto-3g.hdf5")
    one3, two3 = F.read_molecular_integrals(p3)
    m3 = one3.shape[0]
    dump("lih_sto3g_12q", 2 * m3, molecular(one3, two3, e_nuc_lih, [], list(range(m3))),
         "results_and_plots-legacy/molecule_pyscf_sto-3g.hdf5, full space, JW; stands in for "
         "qml.data.load('qchem','LiH','STO-3G') at qchem_model.py:19 (parity unpinned)")

    words, e_scf = hartree_fock_molecular(H2O_LEGACY_SYMBOLS, H2O_LEGACY_BOHR, "sto-6g", 10, 3, 4)
    dump("h2o_sto6g_8q", 8, words,
         "qml_models_legacy.py:1245-1256 rebuilt with tools/sto_ng.py: RHF/STO-6G, core=[0,1,2], "
         "active=[3..6] (4 electrons), JW; orbital phases: largest AO coefficient positive "
         "(the reference's phases are whatever its eigen-solver returned: not recoverable; the "
         "saved search result's energy moves by < 3e-6 across the 8 choices)",
         e_scf=e_scf, e_scf_quoted_angstrom_geometry=-75.1645758157887,
         e_fci_quoted_angstrom_geometry=-75.491788196432,
         e_casci_4e4o=G.sector_ground_energy(8, words, 2, 2))
    r = 0.742 * G.BOHR_PER_ANGSTROM
    words, e_scf = hartree_fock_molecular(["H", "H"], [0, 0, -r / 2, 0, 0, r / 2], "sto-3g", 2, 0, 2)
    dump("h2_sto3g_4q", 4, words,
         "qchem_model.py:18 (qml.data.load qchem H2 STO-3G bondlength 0.742): RHF/STO-3G with "
         "tools/sto_ng.py, full space, JW (the dataset itself is a network download)", e_scf=e_scf)
    sym, xyz = h2o_sto3g_geometry()
    words, e_scf = hartree_fock_molecular(sym, xyz, "sto-3g", 10, 0, 7)
    dump("h2o_sto3g_14q", 14, words,
         "qchem_model.py:20 (qml.data.load qchem H2O STO-3G bondlength 0.958): RHF/STO-3G with "
         "tools/sto_ng.py, full space, JW; H-O-H angle 104.5 deg ASSUMED (not stated in the "
         "reference; the dataset itself is a network download)", e_scf=e_scf)


if __name__ == "__main__":
    main()
