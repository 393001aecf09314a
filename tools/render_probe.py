import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G
from gbnerf_b200 import ops
import bench
dev = torch.device("cuda:0")
torch.manual_seed(0)
nets = [G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").to(dev) for _ in range(2)]
e10, _ = G.get_embedder(10, 0); e4, _ = G.get_embedder(4, 0)
kw = dict(network_query_fn=G.NetworkQuery(e10, e4, 65536), perturb=False, N_importance=64, network_fine=nets[1], N_samples=64,
          network_fn=nets[0], use_viewdirs=True, white_bkgd=True, raw_noise_std=0., ndc=False, lindisp=True, near=1.2, far=8.0)
rays = bench.synthetic_frame_rays(0).to(dev)
nchunks = int(sys.argv[1]) if len(sys.argv) > 1 else 6
rays = rays[:, :32768 * nchunks].contiguous()
with torch.no_grad():
    for it in range(3):
        ops.KERNEL_EVENTS = []
        G.render(756, 1008, 815.0, chunk=32768, rays=rays, **kw)
        torch.cuda.synchronize()
        ev, ops.KERNEL_EVENTS = ops.KERNEL_EVENTS, None
        print("iter", it, " ".join(f"{a.elapsed_time(b):.2f}" for n, a, b, p in ev))

def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

R = 32768
o, d = rays[0, :R], rays[1, :R]
vd = d / d.norm(dim=-1, keepdim=True)
near, far = torch.full((R, 1), 1.2, device=dev), torch.full((R, 1), 8.0, device=dev)
batch = torch.cat([o, d, near, far, vd], -1)
z = ops.zvals_stratified(batch[:, 6:7], batch[:, 7:8], 128, True)
pk = nets[1].packed_weights()
print("views of packed batch (pitch 11):", t(lambda: ops.mlp_forward_raw(pk, "bf16", batch[:, 8:11], R, 128, rays_o=batch[:, 0:3], rays_d=batch[:, 3:6], z=z)))
oc, dc, vc = o.contiguous(), d.contiguous(), vd.contiguous()
print("contiguous o,d,vd (pitch 3):     ", t(lambda: ops.mlp_forward_raw(pk, "bf16", vc, R, 128, rays_o=oc, rays_d=dc, z=z)))
rays_mid = rays[:, 32768:32768 * 2]
o2, d2 = rays_mid[0].contiguous(), rays_mid[1].contiguous()
v2 = (d2 / d2.norm(dim=-1, keepdim=True)).contiguous()
print("contiguous, rays from chunk 3:   ", t(lambda: ops.mlp_forward_raw(pk, "bf16", v2, R, 128, rays_o=o2, rays_d=d2, z=z)))
zr = (torch.rand(R, 128, device=dev) * 6.8 + 1.2).sort(-1)[0]
print("contiguous, random sorted z:     ", t(lambda: ops.mlp_forward_raw(pk, "bf16", vc, R, 128, rays_o=oc, rays_d=dc, z=zr)))
with torch.no_grad():
    print("net.forward_rays (autograd fn):  ", t(lambda: nets[1].forward_rays(oc, dc, vc, z)))
