"""B200-native (sm_100a) sampling hot path of olsson-group/thermodynamic-interpolation.

Layout mirrors the reference's three script trees for the files on the hot path only:

    ambient/   <- mdqm9/thermo/ambient   (integrators.MoleculeIntegrator, models.cpainn.cPaiNN, models.ode_wrapper.ODEWrapper)
    latent/    <- mdqm9/thermo/latent
    adw/       <- adw/thermo             (integrators.StandardIntegrator, models.simple.FCNetMultiBeta)

Everything numerical runs in libtib.so (csrc/, C ABI in include/tib.h).  There is no CPU or
PyTorch fallback: importing the engine without the built library raises.
"""
from .batch import MolBatch, synthetic_ambient_batch, synthetic_latent_batch  # noqa: F401

__all__ = ["MolBatch", "synthetic_ambient_batch", "synthetic_latent_batch"]
__version__ = "0.1.0"
