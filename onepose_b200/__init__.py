"""onepose_b200 -- B200-native (sm_100a) GATsSPG 2D-3D matcher, drop-in for the reference's
``GATsSuperGlue`` forward, plus the rows either side of it (SuperPoint extractor, per-object feature
construction, RANSAC-PnP).  The CUDA library is loaded lazily by the modules; importing this
package on a machine without the built library works, constructing a matcher / extractor does not."""
from .matcher import GATsSuperGlue, LitModelGATsSPG  # noqa: F401
from .extractor import SuperPoint  # noqa: F401
from . import features3d, pnp, synthetic  # noqa: F401

__all__ = ["GATsSuperGlue", "LitModelGATsSPG", "SuperPoint", "features3d", "pnp", "synthetic"]
