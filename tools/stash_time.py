"""Device time of the MLP forward without / with the training stash and of dgrad + wgrad, at the fine pass of a 4,096-ray
training batch (4,096 tiles): what writing and re-reading the stash costs on top of the tensor work."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G
from gbnerf_b200 import ops
from oracle import nerf_oracle as O
dev = torch.device("cuda:0")
R, S = 4096, 128
torch.manual_seed(0)
net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").to(dev)
rays = O.synthetic_rays(R, seed=3).to(dev)
z = O.stratified_z(rays[:, 6:7].cpu(), rays[:, 7:8].cpu(), S, True, torch.rand(R, S)).to(dev)
g_raw = torch.randn(R * S, 4, device=dev)
shapes = [tuple(t.shape) for t in net.param_list()]
stash = ops._stash(R * S, dev)
pk, pkb = net.packed_weights(), net.packed_weights_bwd()


def timed(name, fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us")


timed("forward, no stash", lambda: ops.mlp_forward_raw(pk, "bf16", rays[:, 8:11], R, S, rays_o=rays[:, 0:3], rays_d=rays[:, 3:6], z=z))
timed("forward + stash", lambda: ops.mlp_forward_raw(pk, "bf16", rays[:, 8:11], R, S, rays_o=rays[:, 0:3], rays_d=rays[:, 3:6], z=z, stash=stash))
timed("dgrad + wgrad", lambda: ops.mlp_backward_raw(pkb, g_raw, stash, rays[:, 8:11], R, S, shapes))
