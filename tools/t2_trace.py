"""Timeline of CTA 0 for one steady-state pair iteration of the two-tiles-in-flight MLP kernel (nerf_mlp_t2_kernel).
Per slot: issuer jobs (start, operands ready, weights landed, issued) and epilogue steps (wait start, accumulator ready,
handed over, step done), in SM cycles from the first event.

    python tools/t2_trace.py [iteration]
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G  # noqa: E402
from gbnerf_b200 import _lib, ops  # noqa: E402

it = int(sys.argv[1]) if len(sys.argv) > 1 else 3
R, S = 32768, 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").to(dev)
c2w = torch.zeros(3, 4); c2w[:, :3] = torch.eye(3); c2w[:, 3] = torch.tensor([0.1, -0.05, 0.2])
o, d = G.get_rays(756, 1008, 815.0, c2w.to(dev))
o, d = o.reshape(-1, 3)[:R].contiguous(), d.reshape(-1, 3)[:R].contiguous()
vd = d / d.norm(dim=-1, keepdim=True)
z = ops.zvals_stratified(torch.full((R, 1), 1.2, device=dev), torch.full((R, 1), 8.0, device=dev), S, True)
packed = net.packed_weights()
for _ in range(2):
    ops.mlp_forward_raw(packed, "bf16", vd, R, S, rays_o=o, rays_d=d, z=z)
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
_lib.call("gbn_mlp_set_trace", C.c_void_p(buf.data_ptr()), 0x40000000 | it)
ops.mlp_forward_raw(packed, "bf16", vd, R, S, rays_o=o, rays_d=d, z=z)
torch.cuda.synchronize()
_lib.call("gbn_mlp_set_trace", None, 0)
t = buf.cpu().tolist()
nz = [x for x in t if x]
t0 = min(nz)
rel = lambda x: (x - t0) if x else -1
print(f"pair iteration #{it} of CTA 0; span {max(nz) - t0} cycles")
for s in range(2):
    print(f"slot {s} issuer: job: start | operand wait | weight wait | issue | end")
    for j in range(48):
        a, b, c, e = t[(s * 48 + j) * 4:(s * 48 + j) * 4 + 4]
        if a:
            print(f"  job {j:2d}: {rel(a):7d}  {b - a:6d}  {c - b:6d}  {e - c:5d}  {rel(e):7d}")
for s in range(2):
    print(f"slot {s} epilogue: step: wait start | accumulator ready | handed over | done")
    for si in range(24):
        row = t[512 + (s * 24 + si) * 4:512 + (s * 24 + si) * 4 + 4]
        if row[0]:
            print(f"  step {si:2d}: {rel(row[0]):7d} {rel(row[1]):7d} {rel(row[2]):7d} {rel(row[3]):7d}   "
                  f"(wait {row[1] - row[0]:5d}, to hand-over {row[2] - row[1]:5d}, rest {row[3] - row[2]:5d})")
