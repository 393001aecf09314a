import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gbnerf_b200 as G
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
R, S = 32768, 128
net = G.NeRF_TCNN(encoding="hashgrid").to(dev)
with torch.no_grad():
    net.encoder.params.normal_(0, 0.5)
o = torch.rand(R, 3, device=dev) * 0.2; d = torch.nn.functional.normalize(torch.randn(R, 3, device=dev), dim=-1)
z = torch.sort(torch.rand(R, S, device=dev) * 6.8 + 1.2, -1).values
g_raw = torch.randn(R, S, 4, device=dev) * 1e-3
for _ in range(2):
    out = net.forward_rays(o, d, d, z); net.zero_grad(); out.backward(g_raw)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = net.forward_rays(o, d, d, z); net.zero_grad(); out.backward(g_raw)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
