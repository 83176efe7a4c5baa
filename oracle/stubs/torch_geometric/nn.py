"""`radius_graph` (oracle stub): all ordered pairs (j -> i), i != j, of the same graph with
||x_i - x_j|| < r; no self loops (torch_cluster 1.6.3 semantics as used at
mdqm9/thermo/utils.py:122 with max_num_neighbors=999999).  Edge order is irrelevant:
the reference always `coalesce`s afterwards."""
import torch


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target"):
    n = x.shape[0]
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long)
    d = torch.cdist(x.double(), x.double())
    same = batch[:, None] == batch[None, :]
    mask = same & (d < r)
    if not loop:
        mask &= ~torch.eye(n, dtype=torch.bool)
    tgt, src = torch.nonzero(mask, as_tuple=True)
    return torch.stack([src, tgt], dim=0)
