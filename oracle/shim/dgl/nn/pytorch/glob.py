import torch


class SetTransformerEncoder(torch.nn.Module):
    """Constructible placeholder (TransformerDeepSigns is off the hot path: no shipped config selects it)."""

    def __init__(self, *a, **k):
        super().__init__()
