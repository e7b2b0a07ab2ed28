"""TEST INFRASTRUCTURE ONLY — minimal stand-in for torch_geometric 2.0.1 (pin: Alchemy/setup.sh:3-5) so that the
reference's modules import unmodified.  Only the ops on the SignNet hot path carry semantics; the rest are stubs."""
from . import nn, utils  # noqa: F401
