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

__version__ = "0.1.0"
__all__ = ["GP", "fvGP", "kernels", "NonPositiveDefiniteError", "NativeLibraryError"]
