"""``ModelFromK`` under its reference module path (qas/models.py); defined in interfaces.py."""
from qas_b200.interfaces import ModelFromK  # noqa: F401
