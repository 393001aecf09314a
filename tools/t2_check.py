"""A/B of the two-tiles-in-flight inference MLP kernel (nerf_mlp_t2_kernel) against the one-tile kernel on the same
inputs: output difference, watchdog words, CUDA-event time per launch.  The one-tile kernel is selected in the same
process by handing the library a trace buffer (the dispatch in ts_forward keeps traced launches on it).

    python tools/t2_check.py [R] [S] [iters]
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G  # noqa: E402
from gbnerf_b200 import _lib, ops  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").to(dev)
with torch.no_grad():   # away from the near-zero default biases so that every bias path is exercised
    for p in net.parameters():
        if p.dim() == 1:
            p.add_(0.05 * torch.randn_like(p))
c2w = torch.zeros(3, 4); c2w[:, :3] = torch.eye(3); c2w[:, 3] = torch.tensor([0.1, -0.05, 0.2])
o, d = G.get_rays(756, 1008, 815.0, c2w.to(dev))
o, d = o.reshape(-1, 3)[:R].contiguous(), d.reshape(-1, 3)[:R].contiguous()
vd = d / d.norm(dim=-1, keepdim=True)
z = ops.zvals_stratified(torch.full((R, 1), 1.2, device=dev), torch.full((R, 1), 8.0, device=dev), S, True)
packed = net.packed_weights()
trace = torch.zeros(4096, dtype=torch.int64, device=dev)


def run(one_tile):
    _lib.call("gbn_mlp_set_trace", C.c_void_p(trace.data_ptr()) if one_tile else None, -1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    raw = ws = None
    for i in range(iters):
        if i == 2:
            e0.record()
        raw, ws = ops.mlp_forward_raw(packed, "bf16", vd, R, S, rays_o=o, rays_d=d, z=z)
    e1.record()
    torch.cuda.synchronize()
    _lib.call("gbn_mlp_set_trace", None, 0)
    ms = e0.elapsed_time(e1) / (iters - 2)
    return raw, ms, ops.mlp_error_code(ws)


raw1, ms1, err1 = run(True)
print(f"one tile : {ms1:.3f} ms  {R * S * 1186816 / ms1 / 1e9:7.1f} TFLOP/s  err {err1:#x}", flush=True)
raw2, ms2, err2 = run(False)
print(f"two tiles: {ms2:.3f} ms  {R * S * 1186816 / ms2 / 1e9:7.1f} TFLOP/s  err {err2:#x}", flush=True)
if err2:
    print("watchdog:", _lib.watchdog_report())
diff = (raw1 - raw2).abs()
print(f"max |diff| rgb {diff[..., :3].max().item():.3e}  sigma {diff[..., 3].max().item():.3e}   "
      f"(|raw| max {raw1.abs().max().item():.3e}); nan {torch.isnan(raw2).any().item()}")
raw3, _, _ = run(False)
print("two tiles repeat bit-exactly:", torch.equal(raw2, raw3))
bad = (diff > 2e-2 * (1 + raw1.abs())).any(-1).reshape(-1)
print(f"points off by > 2e-2 relative: {int(bad.sum())} of {bad.numel()}")
if bad.any():
    idx = bad.nonzero().reshape(-1)
    print("first bad points:", idx[:8].tolist(), "tiles:", sorted(set((idx // 128).tolist()))[:16])
    i = int(idx[0])
    print(raw1.reshape(-1, 4)[i].tolist(), raw2.reshape(-1, 4)[i].tolist())
