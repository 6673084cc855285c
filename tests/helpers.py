"""Shared builders for the tests: the pools / models of the BASELINE configs."""
import glob
import json
import os

import numpy as np

from qas_b200 import hamiltonians
from qas_b200.qml_models.qml_gate_ops import QMLPool

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ring(n):
    e = []
    for i in range(n):
        e += [[i, (i + 1) % n], [(i + 1) % n, i]]
    return e


def line(n):
    e = []
    for i in range(n - 1):
        e += [[i, i + 1], [i + 1, i]]
    return e


def near_cx_pool(n, coupling=None):
    """["Rot","PlaceHolder"] + ["CNOT"] on nearest neighbours (every chemistry/QAOA script)."""
    return QMLPool(n, ["Rot", "PlaceHolder"], ["CNOT"], False, coupling if coupling is not None else line(n))


# config -> (model class name, n, p, pool builder, gate limit, init param scale)
CONFIGS = {
    "h2_2q": ("TwoQubitH2", 2, 3, lambda: near_cx_pool(2, [[0, 1], [1, 0]]), {"CNOT": 2}),
    "h2_4q_hf": ("FourQubitH2", 4, 40, lambda: near_cx_pool(4), {"CNOT": 40}),
    "h2_4q_vac": ("FourQubitH2_VaccumInitial", 4, 30, lambda: near_cx_pool(4), {"CNOT": 15}),
    "qaoa_7q": ("QAOAVQCDemo", 7, 15, lambda: near_cx_pool(7, ring(7)), {"CNOT": 7}),
    "qaoa_w5q": ("QAOAWeightedVQCDemo", 5, 10, lambda: near_cx_pool(5, ring(5)), {"CNOT": 5}),
    "h2o_8q": ("H2O", 8, 50, lambda: near_cx_pool(8), {"CNOT": 25}),
    "lih_10q": ("LiH", 10, 20, lambda: near_cx_pool(10), {"CNOT": 10}),
}


def get_model(name):
    from qas_b200.qml_models import qml_models_legacy as L
    from qas_b200.qml_models import qchem_model as Q
    from qas_b200.qml_models import kagome_heisenberg_model as K
    for mod in (L, Q, K):
        if hasattr(mod, name):
            return getattr(mod, name)
    raise KeyError(name)


def goldens():
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.json"))):
        g = json.load(open(f))
        g["fixture"] = os.path.basename(f)[:-5]
        g["pool"] = {int(k): v for k, v in g["pool"].items()}
        g["params"] = np.array(g["params"], dtype=np.float64)
        g["terms"] = hamiltonians.load(g["hamiltonian"])["terms"]
        g["init"] = (g["init"][0], g["init"][1])
        out.append(g)
    return out


def full_vocab_pool(n):
    """Every one- and two-wire gate of the vocabulary on every qubit / ordered pair."""
    from qas_b200.qml_models.qml_gate_ops import SUPPORTED_OPS_DICT
    ones = [s.name for s in SUPPORTED_OPS_DICT.values() if s.num_wires == 1]
    twos = [s.name for s in SUPPORTED_OPS_DICT.values() if s.num_wires == 2]
    return QMLPool(n, ones, twos)


def dense_from_compact(ks, gc, c):
    B, p, l = gc.shape
    out = np.zeros((B, p, c, l))
    for b in range(B):
        for i in range(p):
            out[b, i, ks[b][i]] = gc[b, i]
    return out
