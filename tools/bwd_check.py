"""Stage-by-stage report of the MLP training path (stash / dgrad / wgrad) against fp32 oracle autograd."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import nerf_oracle as O  # noqa: E402
import gbnerf_b200 as G  # noqa: E402
from test_gpu_mlp_backward import decode_stash, rows, oracle_forward_backward, H_FEAT, H_HV, H_ENC, G_HV, G_FEAT, G_L0, G_RAW  # noqa: E402

ops = G.ops
R, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 24)
torch.manual_seed(3)
p = O.init_params(3)
for k in p:
    if k.endswith("weight"):
        p[k] = p[k] * 1.5
net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").cuda()
net.load_state_dict(p)
rays = O.synthetic_rays(R, seed=R)
z = O.stratified_z(rays[:, 6:7], rays[:, 7:8], S, True, torch.rand(R, S, generator=torch.Generator().manual_seed(2)))
P = R * S
ntiles = (P + 127) // 128
pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
emb = torch.cat([O.posenc(pts.reshape(-1, 3), 10), O.posenc(rays[:, None, 8:11].expand(R, S, 3).reshape(-1, 3), 4)], -1)
g_raw = torch.randn(P, 4, generator=torch.Generator().manual_seed(4))
ref = oracle_forward_backward(p, emb, g_raw, emulate_bf16='--fp32' not in sys.argv)
rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-12)).item()

r = rays.cuda()
stash = ops._stash(P, r.device)
stash.zero_()
raw, ws = ops.mlp_forward_raw(net.packed_weights(), "bf16", r[:, 8:11], R, S, rays_o=r[:, 0:3], rays_d=r[:, 3:6], z=z.cuda(), stash=stash)
torch.cuda.synchronize()
print(f"forward: watchdog 0x{ops.mlp_error_code(ws):08x}  raw rel {rel(raw.reshape(P, 4).cpu(), ref['out']):.3e}")
H = decode_stash(stash, ntiles)
print(f"  enc stash rel {rel(rows(H, H_ENC, 1, P)[:, :63], emb[:, :63]):.3e}")
for l in range(8):
    print(f"  h{l} stash rel {rel(rows(H, 4 * l, 4, P), ref['post'][l]):.3e}")
print(f"  feature stash rel {rel(rows(H, H_FEAT, 4, P), ref['feat']):.3e}")
print(f"  hv stash rel {rel(rows(H, H_HV, 2, P), ref['hv']):.3e}")

shapes = [tuple(t.shape) for t in net.param_list()]
grads, ws2, stash_g = ops.mlp_backward_raw(net.packed_weights_bwd(), g_raw.cuda(), stash, r[:, 8:11], R, S, shapes)
torch.cuda.synchronize()
print(f"dgrad watchdog 0x{ops.mlp_error_code(ws2):08x}  wgrad watchdog 0x{int(ws2[256:260].view(torch.int32).item()):08x}")
Gd = decode_stash(stash_g, ntiles)
print(f"  g_raw block rel {rel(rows(Gd, G_RAW, 1, P)[:, :4], g_raw):.3e}")
print(f"  g_hv rel {rel(rows(Gd, G_HV, 2, P), ref['g_hv']):.3e}")
print(f"  g_feat rel {rel(rows(Gd, G_FEAT, 4, P), ref['g_feat']):.3e}")
for l in range(7, -1, -1):
    print(f"  g_{l} rel {rel(rows(Gd, G_L0 + 4 * l, 4, P), ref['g_pre'][l]):.3e}")
for i, name in enumerate(ops.PARAM_ORDER):
    for j, kind in enumerate(("weight", "bias")):
        got, want = grads[2 * i + j].cpu(), ref["grads"][f"{name}.{kind}"]
        print(f"  d{name}.{kind}: rel {rel(got, want):.3e}  |want| {want.norm().item():.3e} |got| {got.norm().item():.3e}")
