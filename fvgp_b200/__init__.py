"""fvgp_b200: B200-native drop-in for the training hot path of lbl-camera/fvGP.

    from fvgp_b200 import GP, fvGP
    from fvgp_b200 import kernels          # the fvgp.kernels names

Importing the package never touches the GPU; the first compute call loads
`fvgp_b200/lib/libfvgp_b200.so` and fails loudly if it (or a CUDA device) is missing.
"""
from . import kernels
from ._lib import NativeLibraryError, NonPositiveDefiniteError
from .fvgp import fvGP
from .gp import GP

__version__ = "0.2.0"
__all__ = ["GP", "fvGP", "kernels", "NonPositiveDefiniteError", "NativeLibraryError", "install_as_fvgp"]


def install_as_fvgp():
    """Register this package under the reference's import name, so that unmodified user code

        from fvgp import GP, fvGP
        from fvgp.kernels import squared_exponential_kernel, get_distance_matrix

    runs on the B200 path.  Call once, before anything imports `fvgp`; refuses to shadow an already imported
    reference installation."""
    import sys
    have = sys.modules.get("fvgp")
    me = sys.modules[__name__]
    if have is not None and have is not me:
        raise ImportError("a different `fvgp` module is already imported; call install_as_fvgp() first")
    sys.modules["fvgp"] = me
    sys.modules["fvgp.kernels"] = kernels
    for name in ("gp", "fvgp", "gp_kv", "gp_prior", "gp_likelihood", "gp_marginal_likelihood", "gp_posterior",
                 "gp_training", "gp_data"):
        sys.modules["fvgp." + name] = sys.modules[__name__ + "." + name]
    return me
