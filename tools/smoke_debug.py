"""Debug helper: which render_rays option makes the coarse map differ from the oracle."""
import itertools, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gbnerf_b200 as G
from oracle import nerf_oracle as O

dev = torch.device("cuda:0")
R, S, N = 256, 64, 64
rays = O.synthetic_rays(R, seed=1)
pc, pf = O.init_params(0), O.init_params(None)
nets = []
for p in (pc, pf):
    net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").to(dev)
    net.load_state_dict(dict(p)); nets.append(net)
nq = G.NetworkQuery(G.get_embedder(10, 0)[0], G.get_embedder(4, 0)[0], 65536)
for lindisp, white, grad, nimp in itertools.product((False, True), (False, True), (False, True), (0, N)):
    with torch.set_grad_enabled(grad):
        out = G.render_rays(rays.to(dev), nets[0], nq, S, lindisp=lindisp, perturb=0., N_importance=nimp,
                            network_fine=nets[1], white_bkgd=white, raw_noise_std=0., retraw=True)
    ref = O.render_rays(rays, pc, pf, S, nimp, lindisp=lindisp, white_bkgd=white, retraw=True)
    k = "rgb0" if nimp else "rgb_map"
    e = (out[k].detach().cpu() - ref[k]).abs()
    print(f"lindisp={lindisp} white={white} grad={grad} nimp={nimp}: {k} max err {e.max():.3e} at ray {e.max(1).values.argmax().item()}"
          f" n_bad {(e.max(1).values > 2e-2).sum().item()}")

with torch.no_grad():
    out = G.render_rays(rays.to(dev), nets[0], nq, S, perturb=0., N_importance=0, raw_noise_std=0., retraw=True)
ref = O.render_rays(rays, pc, pf, S, 0, retraw=True)
er = (out["raw"].cpu() - ref["raw"]).abs().amax(-1)
print("raw err per ray (max over samples), worst rays:", er.amax(1).topk(4))
print("ray 160 raw err by sample:", er[160])
print("ray 160:", rays[160])
print("viewdir 160", (rays[160, 3:6] / rays[160, 3:6].norm()))
from gbnerf_b200 import ops
rgb, disp, acc, w, depth = ops.composite(ref["raw"].to(dev), ref["z_vals"].to(dev) if "z_vals" in ref else out["z_vals"], rays[:, 3:6].to(dev).contiguous())[:5]
print("composite-only err:", (rgb.cpu() - ref["rgb_map"]).abs().max())
print("keys", sorted(out.keys()), sorted(ref.keys()))
print("rgb 160 got", out["rgb_map"][160].cpu(), "want", ref["rgb_map"][160])
for k in ("acc_map", "depth_map", "disp_map"):
    if k in out and k in ref: print(k, out[k][160].item(), ref[k][160].item())
zo = out.get("z_vals"); zr = ref.get("z_vals")
if zo is not None and zr is not None: print("z err", (zo.cpu() - zr).abs().max(), (zo.cpu() - zr).abs()[160].max())
r2 = ops.composite(out["raw"], zo if zo is not None else zr.to(dev), rays[:, 3:6].to(dev).contiguous())
print("composite(out raw) 160:", r2[0][160].cpu())
wr = O.composite(out["raw"].cpu(), (zo.cpu() if zo is not None else zr), rays[:, 3:6])
print("oracle composite(out raw) 160:", wr[0][160] if isinstance(wr, tuple) else wr["rgb_map"][160])
print("sigma raw 160 got", out["raw"][160, :8, 3].cpu(), "want", ref["raw"][160, :8, 3])
