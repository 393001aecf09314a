"""Timings of the ray-setup and depth->normal kernels against the torch expressions they replace (one GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gbnerf_b200 as G
from oracle import nerf_oracle as O

dev = torch.device("cuda:0")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


H, W, f = 756, 1008, 815.0
c2w = O.synthetic_c2w().to(dev)
t_k = timeit(lambda: G.ops.pack_rays(H, W, f, 1.2, 8.0, c2w=c2w, use_viewdirs=True))


def torch_path():
    o, d = G.get_rays(H, W, f, c2w)
    vd = d / torch.norm(d, dim=-1, keepdim=True)
    o, d, vd = o.reshape(-1, 3), d.reshape(-1, 3), vd.reshape(-1, 3)
    return torch.cat([o, d, 1.2 * torch.ones_like(d[:, :1]), 8.0 * torch.ones_like(d[:, :1]), vd], -1)


t_t = timeit(torch_path)
print(f"ray setup {H}x{W}: kernel {t_k:.1f} us ({H * W * 44 / t_k / 1e3:.0f} GB/s written), torch expressions {t_t:.1f} us")

for (h, w) in ((189, 252), (378, 504)):
    depth = (3 + torch.rand(h, w, device=dev)).requires_grad_(True)
    cam = torch.tensor([[0.9 * w, 0, w / 2], [0, 0.9 * w, h / 2], [0, 0, 1.0]], device=dev)

    def ours():
        xyz = G.depth2xyz_torch(depth, cam)
        n = G.depth2normal_geo(xyz.unsqueeze(0).transpose(2, 3).transpose(1, 2), 31)
        n.sum().backward()

    def ref():   # the reference's formulation (unfold + batched inverse + matmuls), run.py:2458-2474, on the GPU
        xyz = G.depth2xyz_torch(depth, cam).unsqueeze(0).permute(0, 3, 1, 2)
        with torch.device(dev):
            n = O.depth2normal_geo(xyz, 31)
        n.sum().backward()

    t_o = timeit(ours, 10)
    try:
        t_r = timeit(ref, 3)
    except RuntimeError as e:
        t_r = float("nan")
    print(f"depth->normals {h}x{w}, k=31, forward+backward: kernels {t_o:.0f} us, reference formulation in torch {t_r:.0f} us")
