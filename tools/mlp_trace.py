"""In-kernel timeline of CTA 0 for one steady-state tile of the fused MLP kernel (gbn_mlp_set_trace).

Trace layout (uint64 clock64 stamps):
  [4j .. 4j+3]            MMA issuer, job j: start, act/enc wait done, weight wait done, issued+committed
  [640 + 2j, +1]          weight producer, job j: before ring-slot wait, after it (TMA issued next)
  [960 + 128*wg + 10u..]  epilogue warpgroup wg, unit u: before acc wait, after, then (ld done, handed over) per block
  [960 + 128*wg + 100..]  unit 10: before wait, after wait, done
  [1216 .. 1219]          encoder row 0: start, computed, ring free, handed over
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G  # noqa: E402
from gbnerf_b200 import ops, _lib  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 3
R, S = 32768, 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision=prec).to(dev)
c2w = torch.zeros(3, 4); c2w[:, :3] = torch.eye(3); c2w[:, 3] = torch.tensor([0.1, -0.05, 0.2])
o, d = G.get_rays(756, 1008, 815.0, c2w.to(dev))
o, d = o.reshape(-1, 3)[:R].contiguous(), d.reshape(-1, 3)[:R].contiguous()
vd = d / d.norm(dim=-1, keepdim=True)
z = ops.zvals_stratified(torch.full((R, 1), 1.2, device=dev), torch.full((R, 1), 8.0, device=dev), S, True)
packed = net.packed_weights()
for _ in range(2):
    ops.mlp_forward_raw(packed, prec, vd, R, S, rays_o=o, rays_d=d, z=z)
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
_lib.call("gbn_mlp_set_trace", C.c_void_p(buf.data_ptr()), tile)
ops.mlp_forward_raw(packed, prec, vd, R, S, rays_o=o, rays_d=d, z=z)
torch.cuda.synchronize()
_lib.call("gbn_mlp_set_trace", None, 0)
t = buf.cpu().tolist()
nz = [x for x in t if x]
t0 = min(nz)
rel = lambda x: (x - t0) if x else -1
print(f"precision {prec}, tile #{tile} of CTA 0; span {max(nz) - t0} cycles")

njobs = max(j for j in range(160) if t[4 * j]) + 1
print("\nMMA issuer per job: start | wait act | wait weights | issue   (cycles since t0; waits are durations)")
prev_end = None
for j in range(njobs):
    a, b, c, e = t[4 * j:4 * j + 4]
    if not a:
        continue
    print(f"  job {j:3d}: start {rel(a):7d}  act-wait {b - a:6d}  w-wait {c - b:6d}  issue {e - c:5d}   end {rel(e):7d}")
print("\nproducer per job: slot wait duration, issue time")
for j in range(njobs):
    a, b = t[640 + 2 * j], t[640 + 2 * j + 1]
    if a:
        print(f"  job {j:3d}: start {rel(a):7d} slot-wait {b - a:6d}")
for wg in range(2):
    base = 960 + 128 * wg
    print(f"\nepilogue wg{wg}: unit: wait-start, acc ready, [ld done, handed over] x blocks")
    for u in range(10):
        row = t[base + 10 * u: base + 10 * u + 10]
        print(f"  unit {u:2d}: " + " ".join(f"{rel(x):7d}" for x in row if x))
    row = t[base + 100: base + 103]
    print("  unit 10: " + " ".join(f"{rel(x):7d}" for x in row if x))
print("\nencoder row 0: start, computed, ring free, handed over:", [rel(x) for x in t[1216:1220]])
