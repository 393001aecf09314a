"""CPU: the oracle restatement against the golden vectors made from the reference itself."""
import numpy as np
import torch

from oracle import nerf_oracle as O


def close(a, b, rtol=1e-6, atol=1e-7):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol, equal_nan=True)


def test_posenc(golden):
    g = golden("posenc.npz")
    assert torch.equal(O.posenc(g["x"], 10), g["enc10"])
    assert torch.equal(O.posenc(g["x"], 4), g["enc4"])
    assert O.posenc_dim(10) == 63 and O.posenc_dim(4) == 27


def test_upper_bound_ties_and_ends(golden):
    g = golden("searchsorted.npz")
    assert torch.equal(O.upper_bound(g["cdf"], g["u"]), g["inds"])
    # the survey's probe
    assert O.upper_bound(g["cdf"][:1], g["u"][:1]).tolist() == [[1, 3, 3, 4, 5, 5]]


def test_sample_pdf(golden):
    g = golden("sample_pdf.npz")
    assert torch.equal(O.sample_pdf(g["bins"], g["weights"], 64), g["det"])
    assert torch.equal(O.sample_pdf(g["bins"], g["weights"], 64, g["u"]), g["rnd"])
    assert torch.equal(O.merge_sorted(g["z"], g["det"]), g["merged_det"])
    assert torch.equal(O.merge_sorted(g["z"], g["rnd"]), g["merged_rnd"])
    # deterministic ends: u=0 hits the first bin exactly; u=1 the last one up to the round-off of cdf[-1]
    assert torch.equal(g["det"][:, 0], g["bins"][:, 0])
    close(g["det"][:, -1], g["bins"][:, -1], rtol=1e-5, atol=0)


def test_composite(golden):
    g = golden("raw2outputs.npz")
    for tag, wb, noise in (("wb0", False, None), ("wb1", True, None), ("noise", True, g["noise"])):
        r = O.composite(g["raw"], g["z"], g["d"], noise, wb)
        for k in ("rgb", "disp", "acc", "weights", "depth", "alpha"):
            assert torch.equal(torch.nan_to_num(r[k], nan=-7.), torch.nan_to_num(g[f"{k}_{tag}"], nan=-7.)), (tag, k)
    # documented edge semantics
    assert torch.isnan(g["disp_wb0"][0]) and g["acc_wb0"][0] == 0 and g["depth_wb0"][0] == 0
    assert float(g["weights_wb0"][1, 0]) > 0.99 and float(g["weights_wb0"][1, 2:].sum()) < 1e-3


def test_composite_grad(golden):
    g = golden("raw2outputs.npz")
    sel = torch.arange(g["raw"].shape[0]) >= 1
    for wb in (False, True):
        for dw in (False, True):
            raw = g["raw"].clone().requires_grad_(True)
            r = O.composite(raw, g["z"], g["d"], None, wb, dw)
            f = (r["rgb"][sel] * g["g_rgb"][sel]).sum() + (r["disp"][sel] * g["g_disp"][sel]).sum() \
                + (r["acc"][sel] * g["g_acc"][sel]).sum() + (r["depth"][sel] * g["g_depth"][sel]).sum()
            (gr,) = torch.autograd.grad(f, raw)
            close(gr, g[f"graw_wb{int(wb)}_dw{int(dw)}"], rtol=1e-5, atol=1e-7)


def _params(g):
    pc, pf = O.init_params(0), O.init_params(None)
    cs = lambda sd: np.array([float(sum(v.double().sum() for v in sd.values())),
                              float(sum((v.double() ** 2).sum() for v in sd.values()))])
    np.testing.assert_allclose(cs(pc), g["csum_coarse"].numpy(), rtol=0, atol=0)
    np.testing.assert_allclose(cs(pf), g["csum_fine"].numpy(), rtol=0, atol=0)
    return pc, pf


def test_mlp(golden):
    g = golden("mlp.npz")
    pc, pf = _params(g)
    close(O.mlp_forward(pc, g["emb"]), g["out_coarse"], rtol=1e-5, atol=1e-6)
    close(O.mlp_forward(pf, g["emb"]), g["out_fine"], rtol=1e-5, atol=1e-6)
    assert sum(v.numel() for v in pc.values()) == 595844


def test_render_test_kwargs(golden):
    g = golden("render_test.npz")
    pc, pf = _params(golden("mlp.npz"))
    with torch.no_grad():
        r = O.render(g["rays"], chunk=32, p_coarse=pc, p_fine=pf, lindisp=True, white_bkgd=True,
                     retraw=True, need_alpha=True)
    for k in ("rgb_map", "disp_map", "acc_map", "depth_map", "weights", "z_vals", "raw", "alpha", "alpha0",
              "rgb0", "disp0", "acc0", "z_std"):
        close(r[k], g[k], rtol=2e-5, atol=2e-6)


def test_render_train_kwargs_and_grads(golden):
    g = golden("render_train.npz")
    pc, pf = _params(golden("mlp.npz"))
    for p in list(pc.values()) + list(pf.values()):
        p.requires_grad_(True)
    r = O.render(g["rays"], p_coarse=pc, p_fine=pf, lindisp=True, white_bkgd=True, retraw=True,
                 t_rand=g["t_rand"], noise0=g["noise0"], u=g["u"], noise1=g["noise1"])
    for k in ("rgb_map", "disp_map", "acc_map", "depth_map", "weights", "z_vals", "raw", "rgb0", "disp0", "acc0", "z_std"):
        close(r[k].detach(), g[k], rtol=2e-5, atol=2e-6)
    loss = O.reference_loss(r, g["target_rgb"], g["target_disp"])
    close(loss.detach(), g["loss"], rtol=1e-5, atol=1e-7)
    loss.backward()
    for tag, sd in (("c", pc), ("f", pf)):
        for k, v in sd.items():
            close(v.grad.norm(), g[f"gnorm_{tag}_{k}"], rtol=2e-4, atol=1e-8)
            if f"grad_{tag}_{k}" in g:
                close(v.grad, g[f"grad_{tag}_{k}"], rtol=1e-3, atol=1e-7)
