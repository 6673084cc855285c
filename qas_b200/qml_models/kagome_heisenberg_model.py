"""16-qubit kagome Heisenberg model (drop-in for qas/qml_models/kagome_heisenberg_model.py:54-161):
H = sum over the 18 lattice edges of XX + YY + ZZ, start state |0...0>."""
from qas_b200.qml_models.hamiltonian_model import HamiltonianModel


class KagomeHeisenberg(HamiltonianModel):
    name = "KagomeHeisenberg"
    num_qubits = 16
    hamiltonian = "kagome_16q"
