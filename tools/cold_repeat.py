"""Repeat the stash-writing forward and the dgrad program from a cold L2 and report WHERE outputs differ from the first
run (tile, 16 KB block), if anywhere.  usage: cold_repeat.py [repeats]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G
from gbnerf_b200 import ops
from oracle import nerf_oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
R, S = 1024, 128
P = R * S
torch.manual_seed(11)
net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").cuda()
net.load_state_dict(O.init_params(11))
rays = O.synthetic_rays(R, seed=3).cuda()
z = O.stratified_z(rays[:, 6:7].cpu(), rays[:, 7:8].cpu(), S, True, torch.rand(R, S, generator=torch.Generator().manual_seed(2))).cuda()
g_raw = torch.randn(P, 4, generator=torch.Generator().manual_seed(4)).cuda()
shapes = [tuple(t.shape) for t in net.param_list()]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def once():
    flush.zero_()
    stash = ops._stash(P, rays.device).zero_()
    raw, ws = ops.mlp_forward_raw(net.packed_weights(), "bf16", rays[:, 8:11], R, S, rays_o=rays[:, 0:3], rays_d=rays[:, 3:6], z=z, stash=stash)
    flush.zero_()
    grads, ws2, sg = ops.mlp_backward_raw(net.packed_weights_bwd(), g_raw, stash, rays[:, 8:11], R, S, shapes)
    torch.cuda.synchronize()
    return raw, stash.view(P // 128, 40, -1), sg.view(P // 128, 40, -1)[:, :39], ops.mlp_error_code(ws), ops.mlp_error_code(ws2)

raw0, h0, g0, e1, e2 = once()
print("first run watchdog words", e1, e2, flush=True)
bad = 0
for it in range(N):
    raw, h, g, e1, e2 = once()
    msgs = []
    if not torch.equal(raw, raw0):
        d = (raw != raw0).any(-1).reshape(-1).nonzero().flatten()
        msgs.append(f"raw: {d.numel()} points differ, tiles {sorted(set((d // 128).tolist()))[:6]}")
    for name, a, b in (("H", h, h0), ("G", g, g0)):
        if not torch.equal(a, b):
            blk = (a != b).any(-1).nonzero()
            msgs.append(f"{name} stash: {blk.shape[0]} blocks differ, first (tile, block): {blk[:8].tolist()}")
    if e1 or e2:
        msgs.append(f"watchdog {hex(e1)} {hex(e2)}")
    if msgs:
        bad += 1
        print(f"repeat {it}: " + "; ".join(msgs), flush=True)
print(f"{bad} of {N} repeats differed", flush=True)
