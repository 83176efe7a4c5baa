"""Attribute-bag `Data`/`Batch` (oracle stub, test infrastructure only).

Semantics mirrored from torch_geometric 2.6.1 as far as the reference relies on them:
  * `batch[name]` / `batch[name] = v` / `hasattr(batch, name)`        (embedding.py:78-84)
  * `del batch.attr` on a missing attribute is a no-op                (ode_wrapper.py:106-107)
  * `clone()` deep-copies every tensor, `to(device)` moves every tensor
  * `Batch.from_data_list`: node-level tensors are concatenated, keys containing
    "index" are offset by the cumulative node count, a `batch` vector and `ptr`
    are added; graph-level scalars are stacked.
  * `to_data_list()` splits node-level tensors back per graph
    (only `.x` / `.x0` are read: integrators.py:29, ode_wrapper.py:75).
"""
import copy

import torch


class Data:
    def __init__(self, **kwargs):
        object.__setattr__(self, "_store", {})
        for k, v in kwargs.items():
            self._store[k] = v

    # -- attribute / item access -------------------------------------------------
    def __getattr__(self, name):
        store = object.__getattribute__(self, "_store")
        if name in store:
            return store[name]
        if name == "edge_index":  # PyG returns None for an absent edge_index
            return None
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self._store[name] = value

    def __delattr__(self, name):
        self._store.pop(name, None)

    def __getitem__(self, name):
        return self._store[name]

    def __setitem__(self, name, value):
        self._store[name] = value

    def __contains__(self, name):
        return name in self._store

    def keys(self):
        return list(self._store.keys())

    # -- whole-object ops --------------------------------------------------------
    def clone(self):
        new = self.__class__()
        for k, v in self._store.items():
            new._store[k] = v.clone() if torch.is_tensor(v) else copy.deepcopy(v)
        return new

    def to(self, device):
        for k, v in list(self._store.items()):
            if torch.is_tensor(v):
                self._store[k] = v.to(device)
        return self

    @property
    def num_nodes(self):
        return self._store["x"].shape[0]


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        out = cls()
        keys = data_list[0].keys()
        n_nodes = [d.num_nodes for d in data_list]
        offsets = [0]
        for n in n_nodes:
            offsets.append(offsets[-1] + n)
        for k in keys:
            vals = [d[k] for d in data_list]
            if not torch.is_tensor(vals[0]):
                out._store[k] = vals
                continue
            if "index" in k:
                out._store[k] = torch.cat([v + off for v, off in zip(vals, offsets[:-1])], dim=-1)
            elif vals[0].dim() == 0:
                out._store[k] = torch.stack(vals)
            else:
                out._store[k] = torch.cat(vals, dim=0)
        out._store["batch"] = torch.cat(
            [torch.full((n,), i, dtype=torch.long) for i, n in enumerate(n_nodes)]
        )
        out._store["ptr"] = torch.tensor(offsets, dtype=torch.long)
        return out

    def to_data_list(self):
        ptr = self._store["ptr"].tolist()
        n_total = ptr[-1]
        n_graphs = len(ptr) - 1
        out = []
        for g in range(n_graphs):
            d = Data()
            lo, hi = ptr[g], ptr[g + 1]
            for k, v in self._store.items():
                if k in ("batch", "ptr") or not torch.is_tensor(v):
                    continue
                if "index" in k:
                    continue  # never read per-graph by the reference
                if v.dim() >= 1 and v.shape[0] == n_total:
                    d._store[k] = v[lo:hi]
                elif v.dim() >= 1 and v.shape[0] == n_graphs:
                    d._store[k] = v[g]
            out.append(d)
        return out
