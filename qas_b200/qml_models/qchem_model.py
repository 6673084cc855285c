"""Dataset-backed chemistry models (drop-in for qas/qml_models/qchem_model.py:23-60).

Upstream downloads the Hamiltonians with ``qml.data.load`` at import time (:18-20), which is
impossible offline; the tables used here are documented in tools/make_hamiltonians.py
(H2: RHF/STO-3G at 0.742 A; LiH: 12-qubit full-space operator rebuilt from the reference's own
HDF5, same bond length to 1e-3 A; H2O: RHF/STO-3G at r(OH) = 0.958 A with an assumed 104.5 deg
angle -- all three are parity-unpinned against the datasets themselves, which are downloads).
"""
from qas_b200.qml_models.hamiltonian_model import HamiltonianModel


class H2STO3G(HamiltonianModel):
    name = "H2_STO-3G"
    num_qubits = 4
    hamiltonian = "h2_sto3g_4q"


class LiHSTO3G(HamiltonianModel):
    name = "LiH_STO-3G"
    num_qubits = 12
    hamiltonian = "lih_sto3g_12q"


class H2OSTO3G(HamiltonianModel):
    name = "H2O_STO-3G"
    num_qubits = 14
    hamiltonian = "h2o_sto3g_14q"
