"""The regulariser MVTN's training loops apply to the rendered views right behind the renderer -- ops.py:138-178
`regualarize_rendered_views` (run_mvtn.py:186,244; viewGCN/tools/Trainer_mvt.py:104) -- as ONE CUDA pass.

The reference runs dropout2d on the 5-D (B, M, 3, H, W) tensor (feature dropout: a whole view is zeroed or scaled by 1 / (1 - p)),
then, batchwise, RandomHorizontalFlip, ReplicationPad2d(int((1 + crop_ratio) H) - H) and RandomCrop(H): four full-size passes, one
of them over the padded copy.  Here the random decisions are drawn with the very torch calls the reference makes, in its order
(so the same seeds give the same images, bit for bit), and mvr_images_regularize_forward applies them in one read + one write."""
import torch

from . import _lib as L
from .ops import _on, _ptr, _stream


class _Regularize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, scale, flip, sy, sx):
        lib = L.load()
        x = images if images.is_contiguous() else images.contiguous()
        B, M, C, H, W = x.shape
        out = torch.empty_like(x)
        flags = L.IMAGES_BF16 if x.dtype is torch.bfloat16 else 0
        with _on(x.device):
            L.check(lib.mvr_images_regularize_forward(_ptr(x), B * M, C, H, W, _ptr(scale), int(flip), sy, sx, flags, _ptr(out),
                                                      _stream(x.device)), "mvr_images_regularize_forward")
        ctx.cfg = (scale, int(flip), sy, sx, flags)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        scale, flip, sy, sx, flags = ctx.cfg
        dt = torch.bfloat16 if flags else torch.float32
        g = g.to(dt).contiguous()
        B, M, C, H, W = g.shape
        gi = torch.empty_like(g)
        with _on(g.device):
            L.check(lib.mvr_images_regularize_backward(_ptr(g), B * M, C, H, W, _ptr(scale), flip, sy, sx, flags, _ptr(gi),
                                                       _stream(g.device)), "mvr_images_regularize_backward")
        return gi, None, None, None, None


def draw_regularizer(x, dropout_p=0, augment_training=False, crop_ratio=0.3):
    """The random decisions of ops.py:168-176 for the (B, M, C, H, W) tensor x, drawn with the reference's own torch calls in its
    order: (per-view factors (B*M,) float32 or None, flip, shift_y, shift_x), shift = crop offset - pad."""
    B, M, C, H, W = x.shape
    scale = None
    if dropout_p > 0:
        # [torch] feature_dropout: noise = input.new_empty((B, M, 1, 1, 1)).bernoulli_(1 - p).div_(1 - p); input * noise
        if dropout_p >= 1:
            scale = torch.zeros(B * M, dtype=torch.float32, device=x.device)
        else:
            noise = torch.empty((B, M, 1, 1, 1), dtype=x.dtype, device=x.device).bernoulli_(1 - dropout_p).div_(1 - dropout_p)
            scale = noise.reshape(-1).to(torch.float32)
    flip, sy, sx = False, 0, 0
    if augment_training:
        if H != W:
            raise ValueError("augment_training crops H x H out of the padded views (RandomCrop(H), ops.py:145): square images only")
        pad = int((1 + crop_ratio) * H) - H
        if pad < 0 or pad >= H:
            raise ValueError("crop_ratio must lie in [0, 1)")
        flip = bool(torch.rand(1) < 0.5)                        # [torchvision] RandomHorizontalFlip.forward
        if pad > 0:                                             # [torchvision] RandomCrop.get_params (no draw when nothing to choose)
            i = torch.randint(0, 2 * pad + 1, size=(1,)).item()
            j = torch.randint(0, 2 * pad + 1, size=(1,)).item()
            sy, sx = i - pad, j - pad
    return scale, flip, sy, sx


def regularize_rendered_views(rendered_images, dropout_p=0, augment_training=False, crop_ratio=0.3):
    """ops.py:168-176.  rendered_images: (B, M, C, H, W) float32 or bfloat16 on a CUDA device; differentiable.
    dropout_p: probability of dropping a whole view (dropout2d on the 5-D tensor, always in training mode as upstream);
    augment_training: one horizontal-flip decision and one crop offset for the whole batch (ops.py:138-146)."""
    x = rendered_images
    if x.dim() != 5:
        raise ValueError("rendered_images must be (B, M, C, H, W)")
    if not x.is_cuda:
        raise L.MVRError("regularize_rendered_views needs CUDA tensors: mvtn_b200 has no CPU path")
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("rendered_images must be float32 or bfloat16")
    scale, flip, sy, sx = draw_regularizer(x, dropout_p, augment_training, crop_ratio)
    if scale is None and not flip and sy == 0 and sx == 0:
        return x
    return _Regularize.apply(x, scale, flip, sy, sx)


regualarize_rendered_views = regularize_rendered_views      # the reference's spelling (ops.py:168)
