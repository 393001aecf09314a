"""Run the NeRF_TCNN forward kernel a few times (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gbnerf_b200 as G
import bench
R, S = int(sys.argv[1]), int(sys.argv[2])
train = len(sys.argv) > 3 and sys.argv[3] == "train"
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = G.NeRF_TCNN(encoding="hashgrid").to(dev)
with torch.no_grad():
    net.encoder.params.normal_(0, 0.5)
rays2 = bench.synthetic_frame_rays(0)
idx = torch.randint(0, rays2.shape[1], (R,), generator=torch.Generator().manual_seed(1))
o, d = rays2[0, idx].to(dev), rays2[1, idx].to(dev)
vd = d / d.norm(dim=-1, keepdim=True)
z = torch.sort(torch.rand(R, S, device=dev) * 6.8 + 1.2, -1).values
g = torch.randn(R, S, 4, device=dev) * 1e-3
for _ in range(5):
    if train:
        out = net.forward_rays(o, d, vd, z); net.zero_grad(); out.backward(g)
    else:
        with torch.no_grad():
            net.forward_rays(o, d, vd, z)
torch.cuda.synchronize()
print("ok")
