"""Launch the fused MLP kernel a few times at a bench-sized chunk (for ncu / timing)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G  # noqa: E402
from gbnerf_b200 import ops  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
S = int(sys.argv[3]) if len(sys.argv) > 3 else 128
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision=prec).to(dev)
c2w = torch.zeros(3, 4); c2w[:, :3] = torch.eye(3); c2w[:, 3] = torch.tensor([0.1, -0.05, 0.2])
o, d = G.get_rays(756, 1008, 815.0, c2w.to(dev))
o, d = o.reshape(-1, 3)[:R].contiguous(), d.reshape(-1, 3)[:R].contiguous()
vd = d / d.norm(dim=-1, keepdim=True)
near, far = torch.full((R, 1), 1.2, device=dev), torch.full((R, 1), 8.0, device=dev)
z = ops.zvals_stratified(near, far, S, True)
packed = net.packed_weights()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    if i == iters - 1:
        e0.record()
    raw, ws = ops.mlp_forward_raw(packed, prec, vd, R, S, rays_o=o, rays_d=d, z=z)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"{prec} R={R} S={S}: {ms:.3f} ms, {R * S * 1186816 / ms / 1e9:.1f} TFLOP/s, err code {ops.mlp_error_code(ws)}")
