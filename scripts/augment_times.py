"""regularize_rendered_views (one gather kernel) vs the reference composition (dropout2d, flip, ReplicationPad2d, RandomCrop:
torch / torchvision on the GPU) at BASELINE configs[1]'s image tensor (32 x 12 views, 3 x 224 x 224), forward and forward+backward."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mvtn_b200
from oracle import torch_ref as tr      # (a timing script under scripts/: the oracle is the thing compared WITH, not shipped)

dev = torch.device("cuda:0")
B, M, S = 32, 12, 224
def timed(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for dtype in (torch.float32, torch.bfloat16):
    x = torch.rand(B, M, 3, S, S, device=dev).to(dtype).requires_grad_()
    g = torch.randn(B, M, 3, S, S, device=dev).to(dtype)
    for name, f in (("reference composition (torch)", tr.regularize_rendered_views), ("mvtn_b200 (one gather)", mvtn_b200.regularize_rendered_views)):
        with torch.no_grad():
            t_f = timed(lambda: f(x, 0.3, True, 0.3))
        def fb():
            x.grad = None
            f(x, 0.3, True, 0.3).backward(g)
        t_fb = timed(fb)
        nbytes = x.numel() * x.element_size()
        print("%-8s %-32s forward %.3f ms (%.0f GB/s of read+write)   forward+backward %.3f ms" % (str(dtype).split(".")[1], name, t_f, 2 * nbytes / t_f / 1e6, t_fb))
