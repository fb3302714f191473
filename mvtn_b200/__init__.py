"""mvtn_b200 -- B200-native (sm_100a) implementation of MVTN's MVRenderer hot path.

    from mvtn_b200 import MVRenderer          # drop-in for models/renderer.py:MVRenderer

The arithmetic lives in libmvr_b200.so (csrc/*.cu, C ABI in include/mvr_b200.h); this package is the
Python host side that mirrors the reference's interface.  There is no CPU fallback.
"""
from .renderer import MVRenderer  # noqa: F401
from .augment import regularize_rendered_views, regualarize_rendered_views  # noqa: F401
from .structures import Meshes  # noqa: F401
from .ops import (HostPackedMeshes, PackedMeshes, camera_position_from_spherical_angles, collate_meshes,  # noqa: F401
                  look_at_view_transform, render_meshes, render_points)
from .cameras import (FoVOrthographicCameras, FoVPerspectiveCameras, OpenGLOrthographicCameras,  # noqa: F401
                      OpenGLPerspectiveCameras)

__version__ = "0.1.0"
