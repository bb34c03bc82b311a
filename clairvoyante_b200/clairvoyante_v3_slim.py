"""Drop-in for reference clairvoyante/clairvoyante_v3_slim.py (callVar.py:33-34, train.py:23-24)."""
from .model import ClairvoyanteV3Slim as Clairvoyante  # noqa: F401
