"""Placeholder so `from torch_geometric.loader import DataLoader` resolves (oracle stub).
The sampling scripts' loaders need the real datasets (absent); the harness builds batches
directly."""


class DataLoader:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("oracle stub: DataLoader is not available (datasets are absent)")
