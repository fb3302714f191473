"""Camera objects returned by MVRenderer.forward (second element of the tuple, renderer.py:198).

The only consumer in the reference is render_and_save -> util.save_cameras -> plot_cameras
(util.py:604-618), which needs get_world_to_view_transform().inverse().transform_points(); ViewGCN also
reads camera positions.  PyTorch3D conventions: row vectors, X_view = X_world R + T.
"""
import torch


class Transform3d:
    """4x4 homogeneous transform in PyTorch3D's row-vector convention (points @ M)."""

    def __init__(self, matrix: torch.Tensor):
        self._matrix = matrix

    def get_matrix(self):
        return self._matrix

    def inverse(self):
        return Transform3d(torch.inverse(self._matrix))

    def compose(self, other: "Transform3d"):
        return Transform3d(self._matrix @ other._matrix)

    def transform_points(self, points: torch.Tensor):
        pts = points if points.dim() == 3 else points[None]
        ones = torch.ones(*pts.shape[:2], 1, dtype=pts.dtype, device=pts.device)
        out = torch.cat([pts, ones], dim=2) @ self._matrix
        out = out[..., :3] / out[..., 3:]
        return out if points.dim() == 3 else out[0]


class _CamerasBase:
    def __init__(self, R, T, centers=None):
        self.R, self.T = R, T
        self._centers = centers
        self.device = R.device

    def __len__(self):
        return self.R.shape[0]

    def get_world_to_view_transform(self):
        n = self.R.shape[0]
        m = torch.zeros((n, 4, 4), dtype=self.R.dtype, device=self.R.device)
        m[:, :3, :3] = self.R
        m[:, 3, :3] = self.T
        m[:, 3, 3] = 1.0
        return Transform3d(m)

    def get_camera_center(self):
        """[upstream] CamerasBase.get_camera_center: translation row of the inverted world-to-view matrix."""
        return self.get_world_to_view_transform().inverse().get_matrix()[:, 3, :3]

    def is_perspective(self):
        raise NotImplementedError


class FoVPerspectiveCameras(_CamerasBase):
    """OpenGLPerspectiveCameras(R, T) as built at renderer.py:84-87: fov 60 deg, znear 1, zfar 100."""

    def __init__(self, R, T, centers=None, fov=60.0, znear=1.0, zfar=100.0, aspect_ratio=1.0):
        super().__init__(R, T, centers)
        self.fov, self.znear, self.zfar, self.aspect_ratio = fov, znear, zfar, aspect_ratio

    def is_perspective(self):
        return True

    def get_znear(self):
        return self.znear


class FoVOrthographicCameras(_CamerasBase):
    """OpenGLOrthographicCameras(R, T, znear=0.01) as built at renderer.py:127-128."""

    def __init__(self, R, T, centers=None, znear=0.01, zfar=100.0):
        super().__init__(R, T, centers)
        self.znear, self.zfar = znear, zfar

    def is_perspective(self):
        return False

    def get_znear(self):
        return self.znear


OpenGLPerspectiveCameras = FoVPerspectiveCameras
OpenGLOrthographicCameras = FoVOrthographicCameras
