"""pnode_b200: B200-native neural-ODE integrator and discrete-adjoint engine behind pnode's Python API.

    from pnode_b200 import petsc_adjoint          # or, unmodified reference scripts:  from pnode import petsc_adjoint
    ode = petsc_adjoint.ODEPetsc(); ode.setupTS(u, func, ...); y = ode.odeint_adjoint(u0, t)

Host code is Python/PyTorch plumbing; all arithmetic of the integrator runs in the sm_100a kernels of
pnode_b200/csrc behind the C ABI of include/pnode_b200.h.  There is no CPU path.
"""
from . import petsc_adjoint  # noqa: F401
from .errors import Error  # noqa: F401
from .options import Options  # noqa: F401

__version__ = "0.1.0"
