"""The BASELINE model classes (drop-in for the Pauli-sum energy models of
qas/qml_models/qml_models_legacy.py): same class names, ``name`` attributes, constructor and
five methods; only what they evaluate is declared here.

  FourQubitH2               :1258-1368  HF start |1100>, H2/sto-3g 15-term operator
  FourQubitH2_VaccumInitial :1370-1480  same operator, |0000> start
  LiH                       :1482-1592  10 qubits, |0..0> start (the BasisState line is commented out upstream)
  H2O                       :1594-1702  8 qubits, |0..0> start; operator rebuilt by RHF/STO-6G (tools/sto_ng.py), pinned to the
                                        reference's quoted SCF/FCI energies and its saved search result
  QAOAVQCDemo               :2328-2491  7 qubits, |+>^7 start, loss = -sum_e (1 - <ZZ>)/2, reward = -loss
  QAOAWeightedVQCDemo       :2493-2651  5 qubits, weighted edges
  VQLSDemo, VQLSDemo5Q      :1807-2326  Hadamard-test cost as two Pauli-sum observables (vqls_model.py)
  TwoQubitH2                missing upstream (imported by search-scripts-legacy/two_qubit_h2.py:3); stand-in

Upstream the QAOA objective runs one QNode per graph edge (:2377-2393); here the edge sum is
one diagonal Pauli-sum evaluated in the same launch as everything else.
"""
from qas_b200.qml_models.hamiltonian_model import HamiltonianModel
from qas_b200.qml_models.utils import extractParamIndicesQML  # noqa: F401  (re-exported upstream too)
from qas_b200.qml_models.vqls_model import VQLSDemo, VQLSDemo5Q  # noqa: F401  (:1807-2326 upstream)


class FourQubitH2(HamiltonianModel):
    name = "FourQubitH2"
    num_qubits = 4
    hamiltonian = "h2_4q"
    init_state = ("basis", 0b1100)     # qml.qchem.hf_state(2, 4)


class FourQubitH2_VaccumInitial(HamiltonianModel):
    name = "FourQubitH2_VaccumInitial"
    num_qubits = 4
    hamiltonian = "h2_4q"


class LiH(HamiltonianModel):
    name = "LiH"
    num_qubits = 10
    hamiltonian = "lih_sto6g_10q"


class H2O(HamiltonianModel):
    name = "H2O"
    num_qubits = 8
    hamiltonian = "h2o_sto6g_8q"


class TwoQubitH2(HamiltonianModel):
    name = "TwoQubitH2"
    num_qubits = 2
    hamiltonian = "h2_2q_standin"


class QAOAVQCDemo(HamiltonianModel):
    name = "QAOAVQCDemo_7Q"
    num_qubits = 7
    n_qubits = 7
    hamiltonian = "qaoa_7q"
    init_state = ("plus", 0)
    graph = [(0, 1), (0, 2), (2, 3), (1, 4), (2, 4), (0, 5), (3, 6), (1, 6)]


class QAOAWeightedVQCDemo(HamiltonianModel):
    name = "QAOAWeightedVQCDemo_5Q"
    num_qubits = 5
    n_qubits = 5
    hamiltonian = "qaoa_weighted_5q"
    init_state = ("plus", 0)
    graph = [(0, 1, 1), (0, 2, 2), (2, 3, 3), (1, 4, 4), (2, 4, 5), (0, 4, 6)]
