"""onepose_b200 -- B200-native (sm_100a) GATsSPG 2D-3D matcher, drop-in for the reference's
``GATsSuperGlue`` forward.  The CUDA library is loaded lazily by the matcher; importing this
package on a machine without the built library works, constructing a matcher does not."""
from .matcher import GATsSuperGlue, LitModelGATsSPG  # noqa: F401
from . import features3d, pnp, synthetic  # noqa: F401

__all__ = ["GATsSuperGlue", "LitModelGATsSPG", "features3d", "pnp", "synthetic"]
