"""Throughput of the NeRF_TCNN kernels (BASELINE config 5 shape: rays through the hash-grid model, coarse 64 + fine 128
points per ray)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gbnerf_b200 as G
import bench

R = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
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


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    ms = timeit(lambda: net.forward_rays(o, d, vd, z))
P = R * S
print(f"forward  R={R} S={S}: {ms:.3f} ms  {P / ms / 1e6:.2f} Gpoints/s  gather {P * 512 / ms / 1e6:.1f} GB/s (512 B/point of 4-byte reads)")
g_raw = torch.randn(R, S, 4, device=dev) * 1e-3


def train():
    out = net.forward_rays(o, d, vd, z)
    net.zero_grad(set_to_none=True)
    out.backward(g_raw)


ms2 = timeit(train, 5)
print(f"forward+backward: {ms2:.3f} ms  ({P / ms2 / 1e6:.2f} Gpoints/s); backward alone ~{ms2 - ms:.3f} ms")
