"""Drop-in for reference clairvoyante/clairvoyante_v3.py: `import clairvoyante_v3 as cv; cv.Clairvoyante()`
(callVar.py:35-46, train.py:25-33)."""
from .model import ClairvoyanteV3 as Clairvoyante  # noqa: F401
