"""``QuantumGate`` / ``Pool`` under their reference module path (qas/ops.py); defined in interfaces.py."""
from qas_b200.interfaces import Pool, QuantumGate  # noqa: F401
