"""Time the 4096-ray training step (BASELINE configs[2]) and its MLP kernels on one GPU."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G  # noqa: E402
from gbnerf_b200 import ops  # noqa: E402
import bench  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
nets = [G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").to(dev)
        for _ in range(2)]
e10, _ = G.get_embedder(10, 0)
e4, _ = G.get_embedder(4, 0)
kw = dict(network_query_fn=G.NetworkQuery(e10, e4, 65536), perturb=1.0, N_importance=64, network_fine=nets[1], N_samples=64,
          network_fn=nets[0], use_viewdirs=True, white_bkgd=True, raw_noise_std=1.0, ndc=False, lindisp=True, near=1.2, far=8.0)
RANK = int(os.environ.get('EMUL_RANK', 0))   # reproduce the data bench.py gives to that rank
rays2 = bench.synthetic_frame_rays(RANK)
idx = torch.randint(0, rays2.shape[1], (R,), generator=torch.Generator().manual_seed(1 + RANK))
rays = rays2[:, idx].contiguous().to(dev)
g = torch.Generator().manual_seed(2 + RANK)
tgt, tgd = torch.rand(R, 3, generator=g).to(dev), torch.rand(R, generator=g).to(dev)
params = [p for n in nets for p in n.parameters()]
opt = (torch.optim.Adam if os.environ.get('STOCK_ADAM') else G.FusedAdam)(params, lr=3e-3)


def step(with_opt=True):
    rgb, disp, acc, depth, ex = G.render(756, 1008, 815.0, chunk=32768, rays=rays, **kw)
    loss = G.img2mse(rgb, tgt) + G.img2mse(ex["rgb0"], tgt) + 0.1 * G.img2mse(disp, tgd)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    if with_opt:
        opt.step()
    return loss


def _report_and_die(exc):
    from gbnerf_b200 import _lib
    import traceback
    print("FAILED:", str(exc).splitlines()[0], flush=True)
    print("   at:", " <- ".join(f"{f.name}:{f.lineno}" for f in reversed(traceback.extract_tb(exc.__traceback__)[-6:])), flush=True)
    print("watchdog report:", _lib.watchdog_report(), flush=True)
    os._exit(3)


try:
    for _ in range(3):
        l = step()
    torch.cuda.synchronize()
except Exception as exc:  # a dead context: the barrier watchdog's host-side record says which wait expired
    _report_and_die(exc)
print("loss", l.item(), "watchdogs", ops.mlp_error_code(nets[1].last_workspace), ops.mlp_error_code(nets[1].last_workspace_bwd))
for with_opt in (False, True):
    ops.KERNEL_EVENTS = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    try:
        for it in range(iters):
            l = step(with_opt)
            if os.environ.get("STEP_SYNC"):
                torch.cuda.synchronize()
                print(f"  [{'on' if with_opt else 'off'} {it}] loss {l.item():.5f}", flush=True)
        e1.record()
        torch.cuda.synchronize()
    except Exception as exc:
        _report_and_die(exc)
    ms = e0.elapsed_time(e1) / iters
    ev, ops.KERNEL_EVENTS = ops.KERNEL_EVENTS, None
    print(f"train step R={R} (optimizer {'on' if with_opt else 'off'}): {ms:.3f} ms/step -> {R / ms * 1e3:.0f} rays/s")
    for name in ("mlp", "mlp_dgrad", "mlp_wgrad"):
        t = sum(a.elapsed_time(b) for n, a, b, _ in ev if n == name) / iters
        pts = sum(p for n, _, _, p in ev if n == name) / iters
        print(f"   {name}: {t:.3f} ms/step  ({pts * 1186816 * (1 if name != 'mlp_wgrad' else 1) / t / 1e9:.0f} TFLOP/s-equivalent of one forward)")
