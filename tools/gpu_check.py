"""First-contact GPU diagnostic: every kernel once against the oracle, then kernel timings at bench sizes.
Run on the B200 box: python tools/gpu_check.py [--skip-mlp]"""
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as O  # noqa: E402
import gbnerf_b200 as G  # noqa: E402
from gbnerf_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)


def stage(name):
    def deco(fn):
        t0 = time.time()
        try:
            fn()
            torch.cuda.synchronize()
            print(f"[ok]   {name} ({time.time() - t0:.2f}s)", flush=True)
        except Exception:  # noqa: BLE001
            print(f"[FAIL] {name}\n{traceback.format_exc()}", flush=True)
        return fn
    return deco


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


g = torch.Generator().manual_seed(0)
R, S = 257, 64
raw = torch.randn(R, S, 4, generator=g)
z = torch.sort(torch.rand(R, S, generator=g) * 6.8 + 1.2, -1)[0]
d = torch.randn(R, 3, generator=g)


@stage("composite forward")
def _():
    want = O.composite(raw, z, d, None, True)
    rgb, disp, acc, w, depth, _a = ops.composite(raw.cuda(), z.cuda(), d.cuda(), None, True)
    for k, v in (("rgb", rgb), ("disp", disp), ("acc", acc), ("weights", w), ("depth", depth)):
        print(f"       {k}: max rel err {((v.cpu() - want[k]).abs() / (want[k].abs() + 1e-6)).max().item():.2e}")


@stage("sample_pdf + merge")
def _():
    w = torch.rand(R, S, generator=g)
    zm = .5 * (z[:, 1:] + z[:, :-1])
    want = O.sample_pdf(zm, w[:, 1:-1], 64)
    merged, std, smp = ops.sample_pdf_merge(z.cuda(), w.cuda(), 64, None, want_samples=True)
    print(f"       samples max abs err {(smp.cpu() - want).abs().max().item():.2e}; merged sorted "
          f"{bool((merged[:, 1:] >= merged[:, :-1]).all().item())}")


params = O.init_params(0)
rays = O.synthetic_rays(300, seed=9)
zz = O.stratified_z(rays[:, 6:7], rays[:, 7:8], 64, True)
pts = rays[:, None, 0:3] + rays[:, None, 3:6] * zz[:, :, None]
want_raw = O.run_network(params, pts, rays[:, 8:11])
_, hidden, feat, hv = O.mlp_forward(params, torch.cat([O.posenc(pts.reshape(-1, 3), 10),
                                                      O.posenc(rays[:, None, 8:11].expand(300, 64, 3).reshape(-1, 3), 4)], -1),
                                    return_hidden=True)

if "--skip-mlp" not in sys.argv:
    for prec in ("bf16", "tf32"):
        @stage(f"mlp forward {prec} (300 rays x 64)")
        def _():
            net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True,
                         precision=prec).to(dev)
            net.load_state_dict(params)
            r = rays.cuda()
            with torch.no_grad():
                got = net.forward_rays(r[:, 0:3], r[:, 3:6], r[:, 8:11], zz.cuda())
            torch.cuda.synchronize()
            code = ops.mlp_error_code(net.last_workspace)
            err = (got.cpu() - want_raw).abs()
            print(f"       watchdog code 0x{code:08x}; max abs err rgb {err[..., :3].max().item():.3e} sigma "
                  f"{err[..., 3].max().item():.3e}; |want| max {want_raw.abs().max().item():.3f}")
            print("       got[0,0:2] ", got[0, 0:2].cpu().tolist())
            print("       want[0,0:2]", want_raw[0, 0:2].tolist())
            print("       got[299,63]", got[299, 63].cpu().tolist(), "want", want_raw[299, 63].tolist())

    for prec in ("bf16", "tf32"):
        @stage(f"mlp timing {prec}")
        def _():
            net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True,
                         precision=prec).to(dev)
            net.load_state_dict(params)
            for Rr, Ss in ((32768, 64), (32768, 128), (4096, 128)):
                r = O.synthetic_rays(Rr, seed=1).cuda()
                zc = ops.zvals_stratified(r[:, 6:7], r[:, 7:8], Ss, True)
                packed = net.packed_weights()
                fn = lambda: ops.mlp_forward_raw(packed, prec, r[:, 8:11], Rr, Ss, rays_o=r[:, 0:3], rays_d=r[:, 3:6], z=zc)
                ms = ev_time(fn)
                tf = Rr * Ss * 1186816 / ms / 1e9
                print(f"       R={Rr} S={Ss}: {ms:.3f} ms/launch  {tf:.1f} TFLOP/s algorithmic "
                      f"({Rr * Ss / ms / 1e3:.1f} Mpts/s)")


@stage("bandwidth kernels timing")
def _():
    for Rr, Ss, Nn in ((32768, 64, 64), (65536, 128, 256)):
        gg = torch.Generator().manual_seed(1)
        raw_ = torch.randn(Rr, Ss, 4, generator=gg).cuda()
        z_ = torch.sort(torch.rand(Rr, Ss, generator=gg) * 6.8 + 1.2, -1)[0].cuda()
        d_ = torch.randn(Rr, 3, generator=gg).cuda()
        w_ = torch.rand(Rr, Ss, generator=gg).cuda()
        u_ = torch.rand(Rr, Nn, generator=gg).cuda()
        ms = ev_time(lambda: ops.composite(raw_, z_, d_, None, True), 20)
        print(f"       composite fwd R={Rr} S={Ss}: {ms * 1e3:.1f} us  {(24 * Ss + 36) * Rr / ms / 1e6:.0f} GB/s algorithmic")
        ms = ev_time(lambda: ops.sample_pdf_merge(z_, w_, Nn, None), 20)
        print(f"       sample+merge det R={Rr} S={Ss} N={Nn}: {ms * 1e3:.1f} us  "
              f"{(8 * Ss + 4 * (Ss + Nn) + 4) * Rr / ms / 1e6:.0f} GB/s algorithmic")
        ms = ev_time(lambda: ops.sample_pdf_merge(z_, w_, Nn, u_), 20)
        print(f"       sample+merge rnd R={Rr} S={Ss} N={Nn}: {ms * 1e3:.1f} us  "
              f"{(8 * Ss + 4 * Nn + 4 * (Ss + Nn) + 4) * Rr / ms / 1e6:.0f} GB/s algorithmic")
        zm = .5 * (z_[:, 1:] + z_[:, :-1]).contiguous()
        wi = w_[:, 1:-1].contiguous()
        ms = ev_time(lambda: ops.sample_pdf(zm, wi, Nn, u_), 20)
        print(f"       sample_pdf rnd R={Rr} B={Ss - 1} N={Nn}: {ms * 1e3:.1f} us  "
              f"{(4 * (Ss - 1) + 4 * (Ss - 2) + 8 * Nn) * Rr / ms / 1e6:.0f} GB/s algorithmic")
    # plain copy for reference
    a = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    b = torch.empty_like(a)
    ms = ev_time(lambda: b.copy_(a), 10)
    print(f"       torch copy 1 GiB: {2 * a.numel() * 4 / ms / 1e6:.0f} GB/s")
