"""fvgp_b200: B200-native drop-in for the training hot path of lbl-camera/fvGP.

    from fvgp_b200 import GP, fvGP
    from fvgp_b200 import kernels          # the fvgp.kernels names

Importing the package never touches the GPU; the first compute call loads
`fvgp_b200/lib/libfvgp_b200.so` and fails loudly if it (or a CUDA device) is missing.
"""
import os as _os

# The population evaluator (ops.lml_population) overlaps up to 32 independent evaluation chains on separate streams;
# the driver multiplexes streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8), and chains that share
# a queue serialise.  Only takes effect when set before the CUDA context exists; a user setting wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import kernels  # noqa: E402
from ._lib import NativeLibraryError, NonPositiveDefiniteError
from .fvgp import fvGP
from .gp import GP

__version__ = "0.1.0"
__all__ = ["GP", "fvGP", "kernels", "NonPositiveDefiniteError", "NativeLibraryError"]
