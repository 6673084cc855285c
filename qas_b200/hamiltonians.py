"""Pauli-sum Hamiltonians of the QAS model zoo, as (coeff, word) tables.

Upstream these are PennyLane ``qml.Hamiltonian`` objects built at import time by
PySCF/OpenFermion (qas/qml_models/qml_models_legacy.py:1237-1256), downloaded datasets
(qchem_model.py:18-20) or in-file edge lists (kagome_heisenberg_model.py:20-52,
qml_models_legacy.py:2347,2514).  Here they are small JSON tables under ``qas_b200/data/``
generated once by ``tools/make_hamiltonians.py`` (which documents provenance, what each table
is pinned against, and which ones are parity-unpinned).
"""
import json
import os

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
_cache = {}


def load(name):
    """-> dict(name, n_qubits, terms=[(coeff, word)], source)."""
    if name not in _cache:
        with open(os.path.join(_DATA, name + ".json")) as f:
            d = json.load(f)
        d["terms"] = [(float(c), str(w)) for c, w in d["terms"]]
        _cache[name] = d
    return _cache[name]


def word_to_masks(word):
    """Pauli word (wire 0 first) -> (xmask, zmask); wire w is bit n-1-w; Y sets both."""
    n = len(word)
    x = z = 0
    for w, ch in enumerate(word):
        bit = 1 << (n - 1 - w)
        if ch in "XY":
            x |= bit
        if ch in "ZY":
            z |= bit
    return x, z


def pauli_terms(terms):
    """[(coeff, word)] -> [(xmask, zmask, coeff)] for the C ABI."""
    return [word_to_masks(w) + (float(c),) for c, w in terms]
