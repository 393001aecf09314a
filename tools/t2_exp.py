"""Timing experiments on the two-tile MLP kernel with the -DGBN_T2_EXP build (csrc/build.py --exp):

    GBNERF_LIB=gb-nerf_b200/libgbnerf_exp.so [GBNERF_T2_DBG_EXTRA=n] python tools/t2_exp.py [R] [S]

Sweeps GBNERF_T2_TURN_BACK (half-jobs before the end of an MMA group at which the turn passes to the other issuer; 2 is
the shipped rule) and the ablations of GBNERF_T2_ABL in ONE process (the exp build reads both per launch); GBNERF_T2_DBG_EXTRA
is fixed per process (it changes the job table).  Ablated runs compute wrong results by design: time only.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G  # noqa: E402
from gbnerf_b200 import ops  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").to(dev)
c2w = torch.zeros(3, 4); c2w[:, :3] = torch.eye(3); c2w[:, 3] = torch.tensor([0.1, -0.05, 0.2])
o, d = G.get_rays(756, 1008, 815.0, c2w.to(dev))
o, d = o.reshape(-1, 3)[:R].contiguous(), d.reshape(-1, 3)[:R].contiguous()
vd = d / d.norm(dim=-1, keepdim=True)
z = ops.zvals_stratified(torch.full((R, 1), 1.2, device=dev), torch.full((R, 1), 8.0, device=dev), S, True)
packed = net.packed_weights()
extra = int(os.environ.get("GBNERF_T2_DBG_EXTRA", "0"))
pairs_per_cta = (R * S / 128 / 2) / 148


def timed(iters=6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ws = None
    for i in range(iters):
        if i == 2:
            e0.record()
        raw, ws = ops.mlp_forward_raw(packed, "bf16", vd, R, S, rays_o=o, rays_d=d, z=z)
    e1.record()
    torch.cuda.synchronize()
    w32 = ws[:128].view(torch.int32).cpu()
    cyc = (int(w32[16]) & 0xffffffff) | (int(w32[17]) << 32)
    return e0.elapsed_time(e1) / (iters - 2), int(w32[0]), cyc, int(w32[18])


def run(label, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    ms, err, cyc, npairs = timed()
    us_pair = ms * 1e3 / pairs_per_cta
    print(f"{label:34s} {ms:8.3f} ms  {R * S * 1186816 / ms / 1e9:7.1f} TFLOP/s (nominal net)  {us_pair:7.3f} us per tile pair  "
          f"CTA 0: {cyc / max(npairs, 1):9.1f} cycles per pair ({npairs} pairs, {cyc / (ms * 1e3):6.1f} MHz)  err {err:#x}", flush=True)
    for k in env:
        os.environ.pop(k, None)
    return ms


print(f"R={R} S={S} extra layers={extra} lib={os.environ.get('GBNERF_LIB', 'product')}", flush=True)
run("warm-up")
base = run("shipped (turn_back 2)")
if os.environ.get("T2_EXP_QUICK"):
    run("shipped again")
elif os.environ.get("GBNERF_T2_MODE") == "2":
    for D in (10, 2):
        run(f"staggered, D = {D}", GBNERF_T2_STAGGER=D)
        run(f"staggered, D = {D}, weights probed first", GBNERF_T2_STAGGER=D, GBNERF_T2_ABL=32)
        run(f"staggered, D = {D}, no rgb, no sin/cos", GBNERF_T2_STAGGER=D, GBNERF_T2_ABL=3)
elif extra == 0:
    for tb in (1, 3):
        run(f"turn_back {tb}", GBNERF_T2_TURN_BACK=tb)
    run("no rgb head (abl 1)", GBNERF_T2_ABL=1)
    run("no sin/cos (abl 2)", GBNERF_T2_ABL=2)
    run("OUT: nothing after the ld (abl 8)", GBNERF_T2_ABL=8)
    run("FLUSH hands over early (abl 16)", GBNERF_T2_ABL=16)
    run("abl 24", GBNERF_T2_ABL=24)
    run("shipped again")
