"""render_and_save helpers (renderer.py:200-207 -> util.py:654-672 save_cameras / save_grid).
Not on the hot path; torchvision / matplotlib are imported lazily and matplotlib is optional."""
import numpy as np
import torch


def save_grid(image_batch, save_path, **kwargs):
    """util.py:667-672: torchvision.utils.save_image of a (M,3,H,W) batch as one grid."""
    from torchvision.utils import save_image
    save_image(image_batch.detach().float().cpu(), save_path, **kwargs)


def get_camera_wireframe(scale: float = 0.3):
    """util.py:586-601: a wireframe of a camera frustum (21 points)."""
    a = 0.5 * torch.tensor([-2, 1.5, 4])
    b = 0.5 * torch.tensor([2, 1.5, 4])
    c = 0.5 * torch.tensor([-2, -1.5, 4])
    d = 0.5 * torch.tensor([2, -1.5, 4])
    C = torch.zeros(3)
    F = torch.tensor([0, 0, 3])
    camera_points = [a, b, d, c, a, C, b, d, C, c, C, F]
    return torch.stack([x.float() for x in camera_points]) * scale


def camera_wireframes_world(cameras, scale: float = 0.3):
    """util.py:604-611: wireframes moved to world space with the inverse world-to-view transform."""
    wires = get_camera_wireframe(scale).to(cameras.R.device)[None]
    cam_trans = cameras.get_world_to_view_transform().inverse()
    return cam_trans.transform_points(wires.expand(len(cameras), -1, -1))


def save_cameras(cameras, save_path, scale=0.22, dpi=200):
    """util.py:654-664.  With matplotlib present: the same 3-D plot; otherwise the wireframes are saved
    as an .npy next to the requested path (matplotlib is not installed in the build image)."""
    wires = camera_wireframes_world(cameras, scale).detach().cpu().numpy()
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        from mpl_toolkits.mplot3d import Axes3D  # noqa: F401
    except Exception:
        np.save(str(save_path) + ".npy", wires)
        return
    fig = plt.figure()
    ax = fig.add_subplot(projection="3d")
    ax.set_xlim(-3.0, 3.0); ax.set_ylim(-3.0, 3.0); ax.set_zlim(-1.8, 1.8)
    ax.scatter(xs=[0], ys=[0], zs=[0], linewidth=3, c="r")
    for wire in wires:
        x_, z_, y_ = wire.T.astype(float)   # Y and Z flipped intentionally, as in util.py:615
        ax.plot(x_, y_, z_, color="blue", linewidth=0.3)
    plt.savefig(save_path, dpi=dpi)
    plt.close(fig)
