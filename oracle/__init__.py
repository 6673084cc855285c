"""CPU oracle for the QAS candidate-circuit evaluation path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy complex128 restatement of what the reference computes on its hot
path: lowering a structure list ``k`` to a gate sequence, evolving a
statevector, taking a Pauli-sum expectation value and differentiating it with respect to
the shared parameter super-set.  It exists to CHECK the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; nothing under ``qas_b200/`` does, and the product path raises
if the CUDA library is missing instead of falling back to this code.

Where the arithmetic lives upstream: un-vendored third-party packages
``pennylane==0.29.1`` (``default.qubit``), ``autograd==1.5`` and
``pennylane-lightning==0.29.0`` (pins: reference ``environment-updated-230314.yaml:142,264-265``).
None of them is installed here, so the oracle restates their *published* definitions
(gate matrices, ``<psi|H|psi>``, exact derivative) and is anchored on the reference's own
call sites and saved results.

Parity pinning status (see DESIGN.md "Oracle"):
  * QAOA 4q/5q/7q energies          -- pinned by the reference's result JSONs (<=1e-6, one
                                       optimiser step of lag between stored params and loss)
  * 4-qubit H2 energies             -- pinned (Hamiltonian ground state vs plot_results.py:48 to
                                       2e-11; stored circuits to the one-step lag, <=2e-5)
  * 10-qubit LiH energy             -- pinned (Hamiltonian rebuilt from the reference's own
                                       molecule_pyscf_sto-6g.hdf5; stored circuit to <=2e-5)
  * gradients                       -- PARITY UNPINNED by any stored vector (the reference
                                       saves none); the oracle's three independent derivative
                                       methods (adjoint, parameter-shift, central difference)
                                       must agree instead
  * 8-qubit H2O energy              -- pinned (Hamiltonian rebuilt by tools/sto_ng.py: its RHF and
                                       FCI energies match search_H2O_near_cx.py:45-49 to 3e-10,
                                       its LiH integrals the shipped HDF5 to 3e-8; stored circuit
                                       to <=5e-6; orbital phases not recoverable)
  * TwoQubitH2, H2/LiH/H2O sto-3g dataset models -- PARITY UNPINNED (class missing upstream /
                                       network downloads; rebuilt with the same RHF code where the
                                       geometry is known, assumptions documented)
"""
