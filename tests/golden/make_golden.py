"""Generate golden vectors from the UNMODIFIED reference (container-only).

    python tests/golden/make_golden.py

Imports the reference hot path from /root/reference through oracle/ref_loader.py,
runs it on seeded synthetic inputs and writes small ``.npz`` fixtures next to this
file.  The fixtures travel to the GPU box; the reference does not.  Model weights
are not stored (2 x 595,844 floats): they are re-created with
``torch.manual_seed(0)`` in the reference's construction order and pinned here by
checksum, so a torch whose RNG/init differs makes the tests fail loudly.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402


def npz(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: tuple(v.shape) for k, v in out.items()})


def main():
    torch.set_num_threads(8)
    ns = ref_loader.load("cpu")
    h = ns["helpers"]

    # ---- positional encoding --------------------------------------------- #
    g = torch.Generator().manual_seed(10)
    x = (torch.rand(50, 3, generator=g) * 2 - 1) * 9.0
    e10, d10 = h.get_embedder(10, 0)
    e4, d4 = h.get_embedder(4, 0)
    assert (d10, d4) == (63, 27)
    npz("posenc.npz", x=x, enc10=e10(x), enc4=e4(x))

    # ---- searchsorted semantics (ties / ends), helpers:333 ----------------- #
    cdf = torch.tensor([[0., .25, .25, .5, 1.], [0., 0., 0., 1., 1.], [0., .1, .2, .3, 1.]])
    u = torch.tensor([[0., .25, .4999, .5, 1., 1.5], [0., 1e-9, .5, 1., 2., -1.], [.1, .2, .3, .05, .99, 1.]])
    npz("searchsorted.npz", cdf=cdf, u=u, inds=torch.searchsorted(cdf, u, right=True))

    # ---- sample_pdf -------------------------------------------------------- #
    g = torch.Generator().manual_seed(3)
    R, S, N = 33, 64, 64
    z = torch.sort(torch.rand(R, S, generator=g) * 6.8 + 1.2, -1)[0]
    bins = .5 * (z[:, 1:] + z[:, :-1])
    w = torch.rand(R, S - 2, generator=g)
    w[0] = 0.                      # flat pdf
    w[1, :] = 0.; w[1, 17] = 5.    # one spike
    w[2, 10:40] = 0.               # plateau in the cdf
    det = h.sample_pdf(bins, w, N, det=True)
    torch.manual_seed(77)
    rnd = h.sample_pdf(bins, w, N, det=False)
    torch.manual_seed(77)
    u_rnd = torch.rand(R, N)
    npz("sample_pdf.npz", bins=bins, weights=w, z=z, det=det, rnd=rnd, u=u_rnd,
        merged_det=torch.sort(torch.cat([z, det], -1), -1)[0],
        merged_rnd=torch.sort(torch.cat([z, rnd], -1), -1)[0])

    # ---- raw2outputs --------------------------------------------------------- #
    g = torch.Generator().manual_seed(0)
    R, S = 37, 64
    raw = torch.randn(R, S, 4, generator=g)
    z = torch.sort(torch.rand(R, S, generator=g) * 6.8 + 1.2, -1)[0]
    d = torch.randn(R, 3, generator=g)
    raw[0, :, 3] = -10.            # empty ray -> acc 0, disp NaN
    raw[1, :, 3] = 50.             # opaque at the first sample
    raw[2, :-1, 3] = -1.; raw[2, -1, 3] = 0.3   # only the 1e10 interval contributes
    noise = torch.randn(R, S, generator=g)
    out = {}
    for wb in (False, True):
        r = h.raw2outputs(raw, z, d, 0., wb, need_alpha=True)
        for k, v in zip(("rgb", "disp", "acc", "weights", "depth", "alpha"), r):
            out[f"{k}_wb{int(wb)}"] = v
    r = h.raw2outputs(raw + torch.cat([torch.zeros(R, S, 3), noise[..., None]], -1), z, d, 0., True, need_alpha=True)
    for k, v in zip(("rgb", "disp", "acc", "weights", "depth", "alpha"), r):
        out[f"{k}_noise"] = v
    # gradients of a fixed scalar functional, for the backward kernel
    raw_g = raw.clone().requires_grad_(True)
    g_rgb = torch.randn(R, 3, generator=g); g_disp = torch.randn(R, generator=g) * .1
    g_acc = torch.randn(R, generator=g); g_depth = torch.randn(R, generator=g)
    for wb in (False, True):
        for dw in (False, True):
            sel = torch.arange(R) >= 1    # drop the NaN-disp ray from the functional
            rr = h.raw2outputs(raw_g, z, d, 0., wb, detach_weights=dw)
            f = (rr[0][sel] * g_rgb[sel]).sum() + (rr[1][sel] * g_disp[sel]).sum() \
                + (rr[2][sel] * g_acc[sel]).sum() + (rr[4][sel] * g_depth[sel]).sum()
            (gr,) = torch.autograd.grad(f, raw_g)
            out[f"graw_wb{int(wb)}_dw{int(dw)}"] = gr
    npz("raw2outputs.npz", raw=raw, z=z, d=d, noise=noise, g_rgb=g_rgb, g_disp=g_disp,
        g_acc=g_acc, g_depth=g_depth, **out)

    # ---- the MLP and the full render ----------------------------------------- #
    tmp = tempfile.mkdtemp()
    args = ref_loader.default_args(tmp)
    torch.manual_seed(0)
    kw_train, kw_test, start, grad_vars, optim = ns["create_nerf"](args)
    for kw in (kw_train, kw_test):
        kw.update(near=O.NEAR, far=O.FAR)
    coarse = O.strip_module_prefix(kw_train["network_fn"].state_dict())
    fine = O.strip_module_prefix(kw_train["network_fine"].state_dict())
    mine_c = O.init_params(0)
    mine_f = O.init_params(None)
    for k in coarse:
        assert torch.equal(coarse[k], mine_c[k]) and torch.equal(fine[k], mine_f[k]), k
    csum = lambda sd: np.array([float(sum(v.double().sum() for v in sd.values())),
                                float(sum((v.double() ** 2).sum() for v in sd.values()))])
    g = torch.Generator().manual_seed(5)
    emb = torch.cat([e10((torch.rand(300, 3, generator=g) * 2 - 1) * 4), e4(torch.nn.functional.normalize(torch.randn(300, 3, generator=g), dim=-1))], -1)
    with torch.no_grad():
        y_c = kw_test["network_fn"](emb)
        y_f = kw_test["network_fine"](emb)
    npz("mlp.npz", emb=emb, out_coarse=y_c, out_fine=y_f, csum_coarse=csum(coarse), csum_fine=csum(fine))

    rays = O.synthetic_rays(48, seed=1)
    o, d = rays[:, 0:3], rays[:, 3:6]
    with torch.no_grad():
        rgb, disp, acc, depth, ex = ns["render"](O.H_FULL, O.W_FULL, O.FOCAL, chunk=32, rays=torch.stack([o, d]),
                                                  retraw=True, need_alpha=True, **kw_test)
    npz("render_test.npz", rays=rays, rgb_map=rgb, disp_map=disp, acc_map=acc, depth_map=depth,
        **{k: v for k, v in ex.items()})

    # train kwargs: replay the RNG stream (t_rand -> noise0 -> u -> noise1), one chunk
    R, S, N = 48, 64, 64
    torch.manual_seed(123)
    rgb, disp, acc, depth, ex = ns["render"](O.H_FULL, O.W_FULL, O.FOCAL, chunk=32768, rays=torch.stack([o, d]),
                                              retraw=True, **kw_train)
    tgt = torch.rand(R, 3, generator=torch.Generator().manual_seed(2))
    tgd = torch.rand(R, generator=torch.Generator().manual_seed(2))
    loss = h.img2mse(rgb, tgt) + h.img2mse(ex["rgb0"], tgt) + 0.1 * h.img2mse(disp, tgd)
    optim.zero_grad()
    loss.backward()
    torch.manual_seed(123)
    t_rand = torch.rand(R, S); noise0 = torch.randn(R, S) * 1.0
    u = torch.rand(R, N); noise1 = torch.randn(R, S + N) * 1.0
    gsel = {}
    for tag, net in (("c", kw_train["network_fn"]), ("f", kw_train["network_fine"])):
        for k, v in O.strip_module_prefix(dict(net.named_parameters())).items():
            if k in ("pts_linears.0.bias", "pts_linears.5.bias", "alpha_linear.weight", "rgb_linear.weight",
                     "views_linears.0.bias", "feature_linear.bias"):
                gsel[f"grad_{tag}_{k}"] = v.grad
            gsel[f"gnorm_{tag}_{k}"] = v.grad.norm()
    npz("render_train.npz", rays=rays, t_rand=t_rand, noise0=noise0, u=u, noise1=noise1,
        target_rgb=tgt, target_disp=tgd, loss=loss, rgb_map=rgb, disp_map=disp, acc_map=acc, depth_map=depth,
        **{k: v for k, v in ex.items()}, **gsel)


if __name__ == "__main__":
    main()
