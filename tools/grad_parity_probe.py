"""How close is the native bf16 backward to the reference's fp32 gradients (tests/golden/render_train.npz, made from the
unmodified reference)?  Prints per-parameter gradient-norm errors and the relative error of the full gradients the golden
carries; the bounds of tests/test_gpu_mlp_render.py::test_render_train_kwargs_golden are set from this."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G
from oracle import nerf_oracle as O
g = {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(ROOT, "tests/golden/render_train.npz")).items()}
torch.manual_seed(0)
params = (O.init_params(0), O.init_params(None))
nets = []
for p in params:
    n = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").cuda()
    n.load_state_dict(p); nets.append(n)
e10, _ = G.get_embedder(10, 0); e4, _ = G.get_embedder(4, 0)
nq = G.NetworkQuery(e10, e4, 65536)
rays = g["rays"].cuda()
rnd = {k: g[k].cuda() for k in ("t_rand", "noise0", "u", "noise1")}
ret = G.render_rays(rays, nets[0], nq, 64, retraw=True, lindisp=True, perturb=1.0, N_importance=64, network_fine=nets[1],
                    white_bkgd=True, raw_noise_std=1.0, _randoms=rnd)
loss = G.img2mse(ret["rgb_map"], g["target_rgb"].cuda()) + G.img2mse(ret["rgb0"], g["target_rgb"].cuda()) \
    + 0.1 * G.img2mse(ret["disp_map"], g["target_disp"].cuda())
loss.backward()
print("loss", loss.item(), "golden", g["loss"].item())
print("rgb0 max err", (ret["rgb0"].cpu() - g["rgb0"]).abs().max().item(), "rgb_map", (ret["rgb_map"].cpu() - g["rgb_map"]).abs().max().item())
worst = 0
for tag, net in (("c", nets[0]), ("f", nets[1])):
    for name, p in net.named_parameters():
        want = g[f"gnorm_{tag}_{name}"].item(); got = p.grad.norm().item()
        e = abs(got - want) / (want + 1e-12); worst = max(worst, e)
        full = ""
        if f"grad_{tag}_{name}" in g:
            w = g[f"grad_{tag}_{name}"]; full = f"  full-gradient rel err {((p.grad.cpu() - w).norm() / w.norm()).item():.3e}"
        print(f"{tag} {name:28s} gnorm got {got:.4e} want {want:.4e} rel {e:.3e}{full}")
print("worst gnorm rel err", worst)
