"""Launch the HBM-bound kernels at bench (C2) and stress (C4) sizes; run under ncu for per-kernel durations."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G  # noqa: E402
from gbnerf_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for Rr, Ss, Nn in ((32768, 64, 64), (32768, 128, 64), (65536, 128, 256), (65536, 384, 256)):
    gg = torch.Generator().manual_seed(1)
    raw_ = torch.randn(Rr, Ss, 4, generator=gg).to(dev).requires_grad_(True)
    z_ = torch.sort(torch.rand(Rr, Ss, generator=gg) * 6.8 + 1.2, -1)[0].to(dev)
    d_ = torch.randn(Rr, 3, generator=gg).to(dev)
    w_ = torch.rand(Rr, Ss, generator=gg).to(dev)
    u_ = torch.rand(Rr, Nn, generator=gg).to(dev)
    noise_ = torch.randn(Rr, Ss, generator=gg).to(dev)
    for it in range(3):
        flush.zero_()
        rgb, disp, acc, w, depth, _ = ops.composite(raw_, z_, d_, noise_ if it == 2 else None, True)
        flush.zero_()
        (rgb.sum() + torch.nan_to_num(disp).sum() * 0.1).backward()
        if Ss <= 128:
            flush.zero_()
            ops.sample_pdf_merge(z_, w_, Nn, None)
            flush.zero_()
            ops.sample_pdf_merge(z_, w_, Nn, u_)
    torch.cuda.synchronize()
print("done")
