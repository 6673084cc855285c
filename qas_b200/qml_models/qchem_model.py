"""Dataset-backed chemistry models (drop-in for qas/qml_models/qchem_model.py:23-60).

Upstream downloads the Hamiltonians with ``qml.data.load`` at import time (:18-20), which is
impossible offline; the tables used here are documented in tools/make_hamiltonians.py
(H2: the 4-qubit sto-3g operator; LiH: 12-qubit full-space operator rebuilt from the
reference's own HDF5; H2O: 14-qubit parity-unpinned stand-in).
"""
from qas_b200.qml_models.hamiltonian_model import HamiltonianModel


class H2STO3G(HamiltonianModel):
    name = "H2_STO-3G"
    num_qubits = 4
    hamiltonian = "h2_4q"


class LiHSTO3G(HamiltonianModel):
    name = "LiH_STO-3G"
    num_qubits = 12
    hamiltonian = "lih_sto3g_12q"


class H2OSTO3G(HamiltonianModel):
    name = "H2O_STO-3G"
    num_qubits = 14
    hamiltonian = "h2o_14q_standin"
