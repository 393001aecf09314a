"""Run the fused sample_pdf+merge kernel a few times on one L2-exceeding set of inputs (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gbnerf_b200 import ops
R, S, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rnd = len(sys.argv) > 4 and sys.argv[4] == "rnd"
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
sets = [(torch.sort(torch.rand(R, S, generator=g) * 6.8 + 1.2, -1)[0].to(dev), torch.rand(R, S, generator=g).to(dev),
         torch.rand(R, N, generator=g).to(dev)) for _ in range(6)]
for z, w, u in sets:
    ops.sample_pdf_merge(z, w, N, u if rnd else None)
torch.cuda.synchronize()
print("ok")
