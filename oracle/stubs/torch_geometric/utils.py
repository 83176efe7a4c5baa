"""`coalesce` (oracle stub): sort edges lexicographically by (row, col) and merge
duplicates with `reduce` - the behaviour the reference depends on at
mdqm9/thermo/utils.py:74-78 (reduce="max")."""
import torch


def coalesce(edge_index, edge_attr=None, reduce="sum"):
    n = int(edge_index.max().item()) + 1 if edge_index.numel() else 0
    key = edge_index[0] * n + edge_index[1]
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    out_index = torch.stack([uniq // n, uniq % n], dim=0)
    if edge_attr is None:
        return out_index
    if reduce == "max":
        out_attr = torch.full((uniq.numel(),), torch.iinfo(edge_attr.dtype).min, dtype=edge_attr.dtype)
        out_attr = out_attr.scatter_reduce(0, inv, edge_attr, reduce="amax")
    elif reduce in ("sum", "add"):
        out_attr = torch.zeros((uniq.numel(),), dtype=edge_attr.dtype).index_add_(0, inv, edge_attr)
    else:
        raise NotImplementedError(reduce)
    return out_index, out_attr
