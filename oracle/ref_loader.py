"""TEST INFRASTRUCTURE ONLY (oracle/).  Imports the *unmodified* reference modules from
/root/reference (build container only - the path does not exist on the GPU box) behind the
stubs in oracle/stubs.  Used by oracle/make_golden.py and by the CPU tests that validate the
restatement in oracle/cpainn_oracle.py against the real thing.

`mdqm9/` and `adw/` both use the top-level package name `thermo`; they are loaded one at a time
(`load_mdqm9()` / `load_adw()` purge `thermo*` from sys.modules before importing).
"""
from __future__ import annotations

import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("TI_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "stubs")
_REPO = os.path.dirname(_HERE)


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mdqm9", "thermo"))


def _prepare(subtree: str):
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT} (only present in the build container)")
    for p in (_REPO, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in [m for m in sys.modules if m == "thermo" or m.startswith("thermo.") or m == "data" or m.startswith("data.")]:
        del sys.modules[name]
    # the reference appends un-normalised entries such as ".../mdqm9/thermo/.." (thermo/utils.py:3)
    trees = {os.path.normpath(os.path.join(REFERENCE_ROOT, other)) for other in ("mdqm9", "adw")}
    sys.path[:] = [p for p in sys.path if os.path.normpath(p) not in trees]
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, subtree))
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    importlib.invalidate_caches()


class _NS:
    pass


def load_mdqm9():
    """Returns a namespace with the reference's ambient/latent model + integrator modules."""
    _prepare("mdqm9")
    ns = _NS()
    ns.ambient_cpainn = importlib.import_module("thermo.ambient.models.cpainn")
    ns.ambient_embedding = importlib.import_module("thermo.ambient.models.embedding")
    ns.ambient_graph = importlib.import_module("thermo.ambient.models.graph")
    ns.ambient_ode_wrapper = importlib.import_module("thermo.ambient.models.ode_wrapper")
    ns.ambient_integrators = importlib.import_module("thermo.ambient.integrators")
    ns.latent_cpainn = importlib.import_module("thermo.latent.models.cpainn")
    ns.latent_ode_wrapper = importlib.import_module("thermo.latent.models.ode_wrapper")
    ns.latent_integrators = importlib.import_module("thermo.latent.integrators")
    ns.utils = importlib.import_module("thermo.utils")
    ns.ambient_losses = importlib.import_module("thermo.ambient.losses")
    ns.ambient_interpolants = importlib.import_module("thermo.ambient.interpolants")
    import torch_geometric
    ns.torch_geometric = torch_geometric
    return ns


def load_adw():
    _prepare("adw")
    ns = _NS()
    ns.simple = importlib.import_module("thermo.models.simple")
    ns.ode_wrapper = importlib.import_module("thermo.models.ode_wrapper")
    ns.integrators = importlib.import_module("thermo.integrators")
    return ns


def load_analysis():
    """mdqm9/analysis/utils/{ess,free_energy,sensititvity}.py import with plain numpy/scipy."""
    if not available():
        raise RuntimeError("reference not found")
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    ns = _NS()
    ns.ess = importlib.import_module("mdqm9.analysis.utils.ess")
    ns.free_energy = importlib.import_module("mdqm9.analysis.utils.free_energy")
    ns.sensitivity = importlib.import_module("mdqm9.analysis.utils.sensititvity")
    return ns
