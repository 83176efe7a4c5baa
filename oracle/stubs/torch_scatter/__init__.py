"""TEST INFRASTRUCTURE ONLY (oracle/): stand-in for torch_scatter 2.1.2 `scatter`.

Reference call sites: mdqm9/thermo/ambient/models/cpainn.py:303-304 -
`scatter(src, index, dim=0)` with the default reduce="sum" and implicit
dim_size = index.max()+1.  Restated with `index_add_`, which on CPU accumulates the
rows in index order (edge order = (src,dst)-lexicographic after coalesce, so for a
fixed destination the sources arrive in ascending order)."""
import torch


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0 and reduce in ("sum", "add") and out is None
    if dim_size is None:
        dim_size = int(index.max().item()) + 1
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add_(0, index, src)
