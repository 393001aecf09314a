"""One launch each of the training-path kernels at a steady-state size, for `ncu --set full` (profiles/r2_*_ncu_full.md):
stash-writing forward, dgrad, wgrad (2048 rays x 128 samples = 2048 tiles), the sampling kernel (65,536 rays, 64 + 64) and
the segmented compositing backward (32,768 rays x 384 samples)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G
from gbnerf_b200 import ops
from oracle import nerf_oracle as O
dev = torch.device("cuda:0")
R, S = 2048, 128
torch.manual_seed(0)
net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").to(dev)
rays = O.synthetic_rays(R, seed=3).to(dev)
z = O.stratified_z(rays[:, 6:7].cpu(), rays[:, 7:8].cpu(), S, True, torch.rand(R, S)).to(dev)
g_raw = torch.randn(R * S, 4, device=dev)
shapes = [tuple(t.shape) for t in net.param_list()]
for it in range(2):      # first pass: one-time setup + warm caches; ncu profiles the launches it is told to (-s / -c)
    stash = ops._stash(R * S, dev)
    raw, ws = ops.mlp_forward_raw(net.packed_weights(), "bf16", rays[:, 8:11], R, S, rays_o=rays[:, 0:3], rays_d=rays[:, 3:6], z=z, stash=stash)
    grads, ws2, sg = ops.mlp_backward_raw(net.packed_weights_bwd(), g_raw, stash, rays[:, 8:11], R, S, shapes)
    Rs = 65536
    zz = (torch.rand(Rs, 64, device=dev) * 6.8 + 1.2).sort(-1)[0]
    w = torch.rand(Rs, 64, device=dev)
    ops.sample_pdf_merge(zz, w, 64)
    ops.sample_pdf_merge(zz, w, 64, torch.rand(Rs, 64, device=dev))
    Rc = 32768
    rawc, zc, d = torch.randn(Rc, 384, 4, device=dev), (torch.rand(Rc, 384, device=dev) * 6.8 + 1.2).sort(-1)[0], torch.randn(Rc, 3, device=dev)
    ops.composite_backward_raw(rawc, zc, d, torch.randn(Rc, 384, device=dev), True, False, torch.randn(Rc, 3, device=dev), torch.randn(Rc, device=dev), None, None)
    torch.cuda.synchronize()
print("watchdog", ops.mlp_error_code(ws), ops.mlp_error_code(ws2), "ok")
