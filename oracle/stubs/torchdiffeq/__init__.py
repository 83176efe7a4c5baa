"""TEST INFRASTRUCTURE ONLY (oracle/): stand-in for torchdiffeq==0.2.5 exposing the two entry
points the reference imports (`from torchdiffeq import odeint_adjoint`).  The arithmetic lives in
oracle/ode_oracle.py (restated from the published algorithm; parity unpinned - see its header)."""
from oracle.ode_oracle import odeint, odeint_adjoint  # noqa: F401
