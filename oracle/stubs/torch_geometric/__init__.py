"""TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for torch_geometric 2.6.1.

The reference (olsson-group/thermodynamic-interpolation) never runs a PyG *kernel* on
its sampling hot path; it only uses `Batch` as a mutable attribute bag plus three
dataset-time helpers (`radius_graph`, `coalesce`, `DataLoader`).  This stub provides
exactly that surface so the reference's own `thermo/**` modules import and execute
unmodified in the build container (ti_env.yml:14 pins the real package; it is not
installable here - no network).  Nothing under `thermodynamic_interpolation_b200/`
may import this.
"""
from . import data, utils, nn, loader  # noqa: F401
