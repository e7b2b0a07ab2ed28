from . import pytorch  # noqa: F401
from .pytorch import SetTransformerEncoder  # noqa: F401
