"""Host-side helpers of the hot path, with the reference's semantics (util.py / ops.py of MVTN)."""
import torch


_named_color_cache = {}
_constant_ids = set()


def torch_color(color_type, custom_color=(1.0, 0, 0), max_lightness=False, epsilon=0.00001):
    """util.py:314-334: name -> RGB 3-vector; max_lightness divides by (max + 1e-5) ("white" = 0.99999).
    The reference's `elif color == "custom"` branch can never match a string name (it compares the
    tensor); "custom" therefore raises here instead of returning an unbound variable."""
    names = {"white": (1.0, 1.0, 1.0), "red": (1.0, 0.0, 0.0), "green": (0.0, 1.0, 0.0),
             "blue": (0.0, 0.0, 1.0), "black": (0.0, 0.0, 0.0)}
    if color_type in names:
        hit = _named_color_cache.get((color_type, max_lightness, epsilon))
        if hit is not None:          # named colours are constants: built once (callers never mutate them)
            return hit
        color = torch.tensor(names[color_type])
    elif color_type == "random":
        color = torch.rand(3)
    elif color_type == "custom":
        color = torch.tensor(custom_color, dtype=torch.float32)
    else:
        raise ValueError(f"unknown color '{color_type}'")
    if max_lightness and color_type != "black":
        color = color / (torch.max(color) + epsilon)
    if color_type in names:
        _named_color_cache[(color_type, max_lightness, epsilon)] = color
        _constant_ids.add(id(color))
    return color


def is_cached_constant(t) -> bool:
    """True for the tensors handed out by torch_color's named-colour cache (never mutated, never freed)."""
    return id(t) in _constant_ids


def batch_tensor(tensor, dim=1, squeeze=False):
    """util.py:509-521: fold dimension `dim` into the batch dimension (flat order b*M + m for x.T)."""
    batch_size, dim_size = tensor.shape[0], tensor.shape[dim]
    returned_size = list(tensor.shape)
    returned_size[0] = batch_size * dim_size
    returned_size[dim] = 1
    out = tensor.transpose(0, dim).reshape(returned_size)
    return out.squeeze(dim) if squeeze else out


def unbatch_tensor(tensor, batch_size, dim=1, unsqueeze=False):
    """util.py:524-534: inverse of batch_tensor."""
    nb_chunks = int(tensor.shape[0] / batch_size)
    if unsqueeze:
        tensor = tensor.unsqueeze(dim)
    return torch.cat(torch.chunk(tensor, nb_chunks, dim=0), dim=dim).contiguous()


def check_valid_rotation_matrix(R, tol: float = 1e-6):
    """util.py:403-420 (host-synchronising torch version; the renderer uses the fused device check of
    mvr_look_at_forward instead and keeps this for callers and tests)."""
    N = R.shape[0]
    eye = torch.eye(3, dtype=R.dtype, device=R.device).view(1, 3, 3).expand(N, -1, -1)
    orthogonal = torch.allclose(R.bmm(R.transpose(1, 2)), eye, atol=tol)
    det_R = torch.det(R)
    no_distortion = torch.allclose(det_R, torch.ones_like(det_R))
    return bool(orthogonal and no_distortion)
