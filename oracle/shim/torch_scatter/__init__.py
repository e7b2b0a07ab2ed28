"""TEST INFRASTRUCTURE ONLY — stand-in for the third-party `torch_scatter` package (absent here, not vendored by
the reference) so the reference's own modules import unmodified on CPU.  Restates the published semantics of
torch_scatter.scatter (pinned by the reference to whatever conda `pyg=2.0.1` resolves, Alchemy/setup.sh:3-5):
`dim_size` defaults to index.max()+1; reduce in {sum, add, mean}; mean divides by counts clamped to 1
(floor division for integer inputs).  Call sites: Alchemy/sign_net/sign_net.py:100, model.py:58-61, transform.py:29.
"""
import torch


def _expand(index, src, dim):
    if dim < 0:
        dim += src.dim()
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src) if index.dim() == 1 else index, dim


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    index_e, dim = _expand(index, src, dim)
    if out is None:
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
        size = list(src.shape)
        size[dim] = dim_size
        out = src.new_zeros(size)
    if reduce in ("sum", "add"):
        return out.scatter_add_(dim, index_e, src)
    if reduce == "mean":
        out = out.scatter_add_(dim, index_e, src)
        cnt = torch.zeros(out.shape[dim], dtype=src.dtype, device=src.device)
        cnt.scatter_add_(0, index if index.dim() == 1 else index.select(dim, 0), torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp_(min=1)
        shape = [1] * out.dim()
        shape[dim] = -1
        if out.is_floating_point():
            return out / cnt.view(shape)
        return torch.div(out, cnt.view(shape), rounding_mode="floor")
    raise ValueError(reduce)


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")
