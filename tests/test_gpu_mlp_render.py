"""GPU parity: the tcgen05 MLP kernel and the full render_rays / render path through the reference-shaped API
against the CPU oracle and the golden vectors made from the reference.

Tolerances (BASELINE.json north_star): MLP outputs within 1e-3 (tf32) or 2e-2 (bf16) absolute.
"""
import types

import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu

TOL = {"bf16": 2e-2, "tf32": 1e-3}
GNORM_TOL, GRAD_TOL = 0.02, 0.12     # native bf16 backward vs the reference's fp32 gradients (measured by tools/grad_parity_probe.py: norms within 0.7 %, full gradients within 7.8 %)


@pytest.fixture(scope="module")
def G():
    import gbnerf_b200
    return gbnerf_b200


def make_net(G, params, precision):
    net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True,
                 precision=precision).cuda()
    net.load_state_dict(params)
    return net


def check_clean(G, net):
    assert G.ops.mlp_error_code(net.last_workspace) == 0, "MLP kernel watchdog fired"


@pytest.fixture(scope="module")
def params():
    torch.manual_seed(0)
    return O.init_params(0), O.init_params(None)


def test_weight_checksums(golden, params):
    g = golden("mlp.npz")
    cs = lambda sd: torch.tensor([float(sum(v.double().sum() for v in sd.values())),
                                  float(sum((v.double() ** 2).sum() for v in sd.values()))], dtype=torch.float64)
    torch.testing.assert_close(cs(params[0]), g["csum_coarse"], rtol=1e-12, atol=0)
    torch.testing.assert_close(cs(params[1]), g["csum_fine"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_mlp_embedded_golden(G, golden, params, precision):
    """NeRF.forward's own contract: [P,90] embedded rows -> [P,4], against the reference's outputs."""
    g = golden("mlp.npz")
    for p, key in ((params[0], "out_coarse"), (params[1], "out_fine")):
        net = make_net(G, p, precision)
        with torch.no_grad():
            y = net(g["emb"].cuda())
        check_clean(G, net)
        assert y.shape == (300, 4)
        err = (y.cpu() - g[key]).abs().max().item()
        assert err < TOL[precision], (precision, key, err)


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
@pytest.mark.parametrize("R,S", [(1, 1), (3, 64), (130, 64), (257, 128), (2048, 192)])
def test_mlp_rays_vs_oracle(G, params, precision, R, S):
    """Fused point generation + encoding + MLP; ragged tiles (P not a multiple of 128), many tiles per CTA."""
    rays = O.synthetic_rays(R, seed=R + S)
    z = O.stratified_z(rays[:, 6:7], rays[:, 7:8], S, True, torch.rand(R, S, generator=torch.Generator().manual_seed(1)))
    pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
    want = O.run_network(params[0], pts, rays[:, 8:11])
    net = make_net(G, params[0], precision)
    r = rays.cuda()
    with torch.no_grad():
        got = net.forward_rays(r[:, 0:3], r[:, 3:6], r[:, 8:11], z.cuda())
        check_clean(G, net)
        got_pts = net.forward_points(pts.cuda(), r[:, 8:11])
        check_clean(G, net)
    assert got.shape == (R, S, 4)
    err = (got.cpu() - want).abs().max().item()
    assert err < TOL[precision], (precision, err)
    assert (got_pts.cpu() - want).abs().max().item() < TOL[precision]


def test_mlp_large_weights_relative(G):
    """Scaled-up weights so activations are O(1..10): checks the kernel against fp32 relatively, not just at
    the tiny magnitudes of default init."""
    torch.manual_seed(5)
    p = O.init_params(5)
    for k in p:
        if k.endswith("weight"):
            p[k] = p[k] * 1.7
    rays = O.synthetic_rays(300, seed=9)
    z = O.stratified_z(rays[:, 6:7], rays[:, 7:8], 64, True)
    pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
    want = O.run_network(p, pts, rays[:, 8:11])
    scale = want.abs().max().item()
    for precision, rel in (("bf16", 3e-2), ("tf32", 2e-3)):
        net = make_net(G, p, precision)
        r = rays.cuda()
        with torch.no_grad():
            got = net.forward_rays(r[:, 0:3], r[:, 3:6], r[:, 8:11], z.cuda())
        check_clean(G, net)
        assert (got.cpu() - want).abs().max().item() < rel * scale


def test_repack_after_update(G, params):
    net = make_net(G, params[0], "tf32")
    rays = O.synthetic_rays(64, seed=2).cuda()
    z = G.ops.zvals_stratified(rays[:, 6:7], rays[:, 7:8], 64, True)
    with torch.no_grad():
        a = net.forward_rays(rays[:, 0:3], rays[:, 3:6], rays[:, 8:11], z)
        net.rgb_linear.bias.add_(0.5)
        b = net.forward_rays(rays[:, 0:3], rays[:, 3:6], rays[:, 8:11], z)
    torch.testing.assert_close(b[..., :3], a[..., :3] + 0.5, rtol=0, atol=1e-5)
    torch.testing.assert_close(b[..., 3], a[..., 3], rtol=0, atol=0)


# ---- full path ------------------------------------------------------------------------------------------------ #
def build_path(G, params, precision):
    nets = [make_net(G, p, precision) for p in params]
    e10, _ = G.get_embedder(10, 0)
    e4, _ = G.get_embedder(4, 0)
    return nets, G.NetworkQuery(e10, e4, 65536)


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_render_test_kwargs_golden(G, golden, params, precision):
    """render() with the reference's test kwargs (perturb=0, raw_noise_std=0) on 48 rays, chunk=32 -> 2 chunks."""
    g = golden("render_test.npz")
    nets, nq = build_path(G, params, precision)
    rays = g["rays"].cuda()
    out = G.render(O.H_FULL, O.W_FULL, O.FOCAL, chunk=32, rays=torch.stack([rays[:, 0:3], rays[:, 3:6]]),
                   near=O.NEAR, far=O.FAR, use_viewdirs=True, ndc=False, retraw=True, need_alpha=True,
                   network_query_fn=nq, perturb=False, N_importance=64, network_fine=nets[1], N_samples=64,
                   network_fn=nets[0], white_bkgd=True, raw_noise_std=0., lindisp=True)
    rgb, disp, acc, depth, ex = out
    assert set(ex) == {"weights", "z_vals", "raw", "alpha", "alpha0", "rgb0", "disp0", "acc0", "z_std"}
    tol = 2.5 * TOL[precision]
    # coarse outputs share every input with the reference -> MLP tolerance
    assert (ex["rgb0"].cpu() - g["rgb0"]).abs().max().item() < TOL[precision]
    assert (ex["acc0"].cpu() - g["acc0"]).abs().max().item() < TOL[precision]
    assert (ex["alpha0"].cpu() - g["alpha0"]).abs().max().item() < TOL[precision]
    # fine outputs are evaluated at depths drawn from the (reduced-precision) coarse weights
    assert (rgb.cpu() - g["rgb_map"]).abs().max().item() < tol
    assert (acc.cpu() - g["acc_map"]).abs().max().item() < tol
    # sample_pdf is discontinuous where a bin's cdf step crosses the reference's 1e-5 denominator guard
    # (helpers:344-345), so single samples may jump by a bin under any rounding change; the bulk must agree
    dz = (ex["z_vals"].cpu() - g["z_vals"]).abs()
    # (bf16 coarse weights move ~1e-4, which shifts the inverse-CDF samples by a few 1e-3 on this flat random-init
    # density; measured fractions: tf32 0.6 %, bf16 8.8 %)
    assert (dz > 5e-3).float().mean().item() < (0.15 if precision == "bf16" else 0.02) and dz.median().item() < 1e-4
    assert rgb.shape == (48, 3) and ex["weights"].shape == (48, 128) and ex["raw"].shape == (48, 128, 4)


def test_render_train_kwargs_golden(G, golden, params):
    """Train kwargs with the reference's RNG stream replayed (t_rand -> noise0 -> u -> noise1): forward, loss and the
    parameter gradients of the NATIVE backward (tcgen05 dgrad + wgrad, bf16) against the unmodified reference's fp32
    autograd (tests/golden/render_train.npz): every parameter's gradient norm and the full gradients the golden carries.
    Bounds: what bf16 storage of activations / gradients gives on this batch with head-room (tools/grad_parity_probe.py
    prints the measured figures; a corrupted tile - the round-1 race - moves them by far more)."""
    g = golden("render_train.npz")
    nets, nq = build_path(G, params, "bf16")
    rays = g["rays"].cuda()
    rnd = {k: g[k].cuda() for k in ("t_rand", "noise0", "u", "noise1")}
    ret = G.render_rays(rays, nets[0], nq, 64, retraw=True, lindisp=True, perturb=1.0, N_importance=64,
                        network_fine=nets[1], white_bkgd=True, raw_noise_std=1.0, _randoms=rnd)
    loss = G.img2mse(ret["rgb_map"], g["target_rgb"].cuda()) + G.img2mse(ret["rgb0"], g["target_rgb"].cuda()) \
        + 0.1 * G.img2mse(ret["disp_map"], g["target_disp"].cuda())
    loss.backward()
    assert G.ops.mlp_error_code(nets[1].last_workspace_bwd) == 0 and G.ops.mlp_error_code(nets[0].last_workspace_bwd) == 0
    assert (ret["rgb0"].cpu() - g["rgb0"]).abs().max().item() < TOL["bf16"]
    assert (ret["rgb_map"].cpu() - g["rgb_map"]).abs().max().item() < 2.5 * TOL["bf16"]
    assert abs(loss.item() - g["loss"].item()) < 5e-3 * max(1.0, abs(g["loss"].item()))
    for tag, net in (("c", nets[0]), ("f", nets[1])):
        for name, p in net.named_parameters():
            want = g[f"gnorm_{tag}_{name}"].item()
            got = p.grad.norm().item()
            assert abs(got - want) <= GNORM_TOL * want + 1e-6, (tag, name, got, want)
            if f"grad_{tag}_{name}" in g.keys():
                w = g[f"grad_{tag}_{name}"]
                rel = ((p.grad.cpu() - w).norm() / w.norm()).item()
                assert rel < GRAD_TOL, (tag, name, rel)


def test_tf32_modules_are_inference_only(G, params):
    """No cuBLAS / eager fallback behind the native backward: a tf32 module renders, asking it for gradients raises."""
    nets, nq = build_path(G, params, "tf32")
    rays = O.synthetic_rays(16, seed=1).cuda()
    ret = G.render_rays(rays, nets[0], nq, 64, lindisp=True, perturb=0., N_importance=64, network_fine=nets[1],
                        white_bkgd=True, raw_noise_std=0.)
    with pytest.raises(NotImplementedError, match="inference-only"):
        ret["rgb_map"].sum().backward()
    with pytest.raises(NotImplementedError, match="not implemented"):
        nets_b, nq_b = build_path(G, params, "bf16")
        z = torch.rand(16, 8, device="cuda").sort(-1)[0].requires_grad_(True)
        nets_b[0].forward_rays(rays[:, 0:3], rays[:, 3:6], rays[:, 8:11], z)


def test_embedded_form_uses_the_native_backward(G, params):
    """NeRF.forward(x) on pre-embedded rows (the reference module's own signature): gradients come from the same tcgen05
    dgrad + wgrad kernels and agree with fp32 autograd of the oracle network within bf16 tolerance; a second backward
    through the same call is refused with a clear message."""
    net = make_net(G, params[0], "bf16")
    g = torch.Generator().manual_seed(3)
    pts = torch.rand(300, 3, generator=g) * 4 - 2
    dirs = torch.nn.functional.normalize(torch.randn(300, 3, generator=g), dim=-1)
    emb = torch.cat([O.posenc(pts, 10), O.posenc(dirs, 4)], -1)
    out = net(emb.cuda())
    go = torch.randn(300, 4, generator=g)
    out.backward(go.cuda(), retain_graph=True)
    prm = {k: v.clone().requires_grad_(True) for k, v in params[0].items()}
    want = O.mlp_forward(prm, emb)
    want.backward(go)
    assert (out.detach().cpu() - want.detach()).abs().max().item() < TOL["bf16"]
    for name, p in net.named_parameters():
        rel = ((p.grad.cpu() - prm[name].grad).norm() / (prm[name].grad.norm() + 1e-12)).item()
        # white-noise output gradients through bf16 activations: the error grows towards the first layers (measured
        # 11 % at pts_linears.0); tests/test_gpu_mlp_backward.py pins every stage against a bf16-aware oracle at 4 %
        assert rel < 0.2, (name, rel)
    with pytest.raises(RuntimeError, match="twice"):
        out.backward(go.cuda())


def test_render_rays_pieces_consistent(G, params):
    """Stage-by-stage: feeding the oracle the CUDA path's own intermediate tensors reproduces every later stage
    to the per-stage tolerance (so end-to-end drift is only the documented MLP precision)."""
    nets, nq = build_path(G, params, "tf32")
    rays = O.synthetic_rays(200, seed=12)
    ret = G.render_rays(rays.cuda(), nets[0], nq, 64, retraw=True, lindisp=True, perturb=0., N_importance=64,
                        network_fine=nets[1], white_bkgd=True, raw_noise_std=0.)
    z = ret["z_vals"].detach().cpu()
    pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
    raw = O.run_network(params[1], pts, rays[:, 8:11])
    assert (ret["raw"].detach().cpu() - raw).abs().max().item() < TOL["tf32"]
    c = O.composite(ret["raw"].detach().cpu(), z, rays[:, 3:6], None, True)
    torch.testing.assert_close(ret["rgb_map"].detach().cpu(), c["rgb"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ret["weights"].detach().cpu(), c["weights"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ret["depth_map"].detach().cpu(), c["depth"], rtol=1e-5, atol=1e-6)
    assert torch.equal(torch.sort(z, -1)[0], z)


def test_create_nerf_api(G, tmp_path):
    """create_nerf(args) mirrors run.py:2003-2128: kwargs keys, parameter names/shapes, Adam, checkpoint reload."""
    args = types.SimpleNamespace(
        multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_samples=64, N_importance=64, netdepth=8,
        netdepth_fine=8, netwidth=256, netwidth_fine=256, alpha_model_path=None, no_coarse=False, netchunk=65536,
        lrate=3e-3, basedir=str(tmp_path), expname="exp", ft_path=None, no_reload=False, perturb=1.0,
        white_bkgd=True, raw_noise_std=1.0, dataset_type="llff", no_ndc=True, lindisp=True, sigma_loss=False)
    (tmp_path / "exp").mkdir()
    torch.manual_seed(0)
    kw_train, kw_test, start, grad_vars, optim = G.create_nerf(args)
    assert start == 0 and len(grad_vars) == 48 and sum(p.numel() for p in grad_vars) == 2 * 595844
    assert set(kw_train) == {"network_query_fn", "perturb", "N_importance", "network_fine", "N_samples", "network_fn",
                             "use_viewdirs", "white_bkgd", "raw_noise_std", "ndc", "lindisp"}
    assert kw_test["perturb"] is False and kw_test["raw_noise_std"] == 0.
    sd = kw_train["network_fn"].state_dict()
    assert all(k.startswith("module.") for k in sd)
    want = O.init_params(0)
    for k, v in want.items():
        assert torch.equal(sd["module." + k].cpu(), v), k
    # one optimisation step on a few rays, save, reload through create_nerf
    rays = O.synthetic_rays(64, seed=5).cuda()
    kw = dict(kw_train, near=O.NEAR, far=O.FAR)
    rgb, disp, acc, depth, ex = G.render(O.H_FULL, O.W_FULL, O.FOCAL, chunk=32768,
                                         rays=torch.stack([rays[:, 0:3], rays[:, 3:6]]), **kw)
    loss = G.img2mse(rgb, torch.zeros_like(rgb)) + G.img2mse(ex["rgb0"], torch.zeros_like(rgb))
    optim.zero_grad()
    loss.backward()
    optim.step()
    torch.save({"global_step": 7, "network_fn_state_dict": kw_train["network_fn"].state_dict(),
                "network_fine_state_dict": kw_train["network_fine"].state_dict(),
                "optimizer_state_dict": optim.state_dict()}, str(tmp_path / "exp" / "000007.tar"))
    kw2, _, start2, _, _ = G.create_nerf(args)
    assert start2 == 7
    for a, b in zip(kw_train["network_fn"].parameters(), kw2["network_fn"].parameters()):
        assert torch.equal(a, b)
    rgb2, *_ = G.render(O.H_FULL, O.W_FULL, O.FOCAL, chunk=32768, rays=torch.stack([rays[:, 0:3], rays[:, 3:6]]),
                        **dict(kw_test, near=O.NEAR, far=O.FAR))
    assert torch.isfinite(rgb2).all()


def test_sigma_loss_vs_reference_formula(G, params):
    """SigmaLoss (DS_NeRF/loss.py:15-44): extra march near -> depth, same random tensors injected on both sides."""
    import torch.nn.functional as F
    nets, nq = build_path(G, params, "tf32")
    R, S = 150, 64
    rays = O.synthetic_rays(R, seed=21)
    g = torch.Generator().manual_seed(3)
    depths = 2.0 + 3.0 * torch.rand(R, generator=g)
    t_rand = torch.rand(R, S, generator=g)
    noise = torch.randn(R, S, generator=g)
    sl = G.SigmaLoss(S, 1.0, 1.0)
    r = rays.cuda()
    with torch.no_grad():
        got = sl.calculate_loss(r[:, 0:3], r[:, 3:6], r[:, 8:11], r[:, 6:7], r[:, 7:8], depths.cuda(), nq, nets[1],
                                _randoms={"t_rand": t_rand.cuda(), "noise": noise.cuda()})
    # reference formula on the oracle
    z = O.stratified_z(rays[:, 6:7], depths[:, None], S, False, t_rand)
    pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
    raw = O.run_network(params[1], pts, rays[:, 8:11])
    sigma = F.relu(raw[..., 3] + noise)
    want = -torch.exp(sigma[:, -1]) / (torch.sum(torch.exp(sigma), 1) + 1)
    assert got.shape == (R,)
    torch.testing.assert_close(got.cpu(), want, rtol=2e-3, atol=2e-4)
    # and through render_rays when the batch carries depths (run.py:2372-2375)
    batch = torch.cat([rays[:, :8], depths[:, None], rays[:, 8:11]], -1).cuda()
    ret = G.render_rays(batch, nets[0], nq, 64, lindisp=True, perturb=0., N_importance=64, network_fine=nets[1],
                        white_bkgd=True, raw_noise_std=0., sigma_loss=G.SigmaLoss(S, 0., 0.))
    assert ret["sigma_loss"].shape == (R,) and torch.isfinite(ret["sigma_loss"]).all()


# ---- API behaviour of the drop-in functions (run.py:1624-1748, 2235-2381) ----------------------------------------- #
def test_render_from_c2w_and_options(G, params):
    """render(c2w=...) generates rays like get_rays (helpers:251-262); N_importance=0, lindisp/white_bkgd off,
    need_alpha, detach_weights, patch and c2w_staticcam all follow the reference's branches."""
    nets, nq = build_path(G, params, "tf32")
    c2w = O.synthetic_c2w().cuda()
    H, W, focal = 12, 16, 14.0
    kw = dict(network_query_fn=nq, perturb=0., N_importance=64, network_fine=nets[1], N_samples=64, network_fn=nets[0],
              white_bkgd=True, raw_noise_std=0., lindisp=True)
    with torch.no_grad():
        rgb, disp, acc, depth, ex = G.render(H, W, focal, chunk=64, c2w=c2w[:3, :4], near=O.NEAR, far=O.FAR, use_viewdirs=True,
                                             ndc=False, **kw)
        assert rgb.shape == (H, W, 3) and disp.shape == (H, W) and ex["weights"].shape == (H, W, 128)
        # same thing through explicit rays
        o, d = O.get_rays(H, W, focal, O.synthetic_c2w())
        rgb2, *_ = G.render(H, W, focal, chunk=1000, rays=torch.stack([o.reshape(-1, 3), d.reshape(-1, 3)]).cuda(), near=O.NEAR,
                            far=O.FAR, use_viewdirs=True, ndc=False, **kw)
        torch.testing.assert_close(rgb.reshape(-1, 3), rgb2, rtol=1e-4, atol=1e-5)
        # oracle on the same rays (coarse-only, linear depth sampling, black background)
        rays = O.pack_rays(o, d, O.NEAR, O.FAR)
        kw0 = dict(kw, N_importance=0, white_bkgd=False, lindisp=False)
        r0, d0, a0, z0, e0 = G.render(H, W, focal, chunk=50, c2w=c2w[:3, :4], near=O.NEAR, far=O.FAR, use_viewdirs=True, ndc=False,
                                      retraw=True, **kw0)
        want = O.render(rays, chunk=64, p_coarse=params[0], p_fine=None, n_samples=64, n_importance=0, lindisp=False,
                        white_bkgd=False, retraw=True)
        assert set(e0) == {"weights", "z_vals", "raw"}
        assert (r0.reshape(-1, 3).cpu() - want["rgb_map"]).abs().max().item() < TOL["tf32"]
        assert (e0["z_vals"].reshape(-1, 64).cpu() - want["z_vals"]).abs().max().item() < 1e-5
        # a patch of the frame == the same pixels of the full frame
        p_rgb, *_ = G.render(H, W, focal, chunk=64, c2w=c2w[:3, :4], near=O.NEAR, far=O.FAR, use_viewdirs=True, ndc=False,
                             patch=(2, 3, 5, 6), **kw)
        torch.testing.assert_close(p_rgb, rgb[2:7, 3:9], rtol=1e-4, atol=1e-5)
        # need_alpha adds alpha / alpha0; with N_importance == 0 the reference raises (undefined alpha0, run.py:2365)
        *_, ea = G.render(H, W, focal, chunk=64, c2w=c2w[:3, :4], near=O.NEAR, far=O.FAR, use_viewdirs=True, ndc=False,
                          need_alpha=True, **kw)
        assert ea["alpha"].shape == (H, W, 128) and ea["alpha0"].shape == (H, W, 64)
    with pytest.raises(NameError):
        G.render(H, W, focal, chunk=64, c2w=c2w[:3, :4], near=O.NEAR, far=O.FAR, use_viewdirs=True, ndc=False, need_alpha=True,
                 **dict(kw, N_importance=0))


def test_ndc_and_foreign_network_fallback(G, params):
    """ndc=True runs ndc_rays (helpers:285-302) before packing; run_network keeps the reference's generic route
    (embed, concatenate, netchunk slices) for a network that is not the package's NeRF."""
    nets, nq = build_path(G, params, "tf32")
    H, W, focal = 8, 10, 9.0
    c2w = O.synthetic_c2w()
    o, d = O.get_rays(H, W, focal, c2w)
    on, dn = O.ndc_rays(H, W, focal, 1., o, d)
    got_o, got_d = G.ndc_rays(H, W, focal, 1., o.cuda(), d.cuda())
    torch.testing.assert_close(got_o.cpu(), on, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got_d.cpu(), dn, rtol=1e-5, atol=1e-6)
    with torch.no_grad():
        rgb, *_ = G.render(H, W, focal, chunk=64, c2w=c2w.cuda()[:3, :4], near=0., far=1., use_viewdirs=True, ndc=True,
                           network_query_fn=nq, perturb=0., N_importance=0, network_fine=None, N_samples=32, network_fn=nets[0],
                           white_bkgd=False, raw_noise_std=0.)
    assert rgb.shape == (H, W, 3) and torch.isfinite(rgb).all()

    class Foreign(torch.nn.Module):           # anything callable on [P, 90] -> [P, 4]
        def __init__(self, p):
            super().__init__()
            self.p = {k: v.cuda() for k, v in p.items()}

        def forward(self, x):
            return O.mlp_forward(self.p, x)

    rays = O.synthetic_rays(20, seed=3)
    z = O.stratified_z(rays[:, 6:7], rays[:, 7:8], 16, True)
    pts = (rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]).cuda()
    with torch.no_grad():
        raw_f = G.run_network(pts, rays[:, 8:11].cuda(), Foreign(params[0]), nq.embed_fn, nq.embeddirs_fn, netchunk=100)
        raw_n = G.run_network(pts, rays[:, 8:11].cuda(), nets[0], nq.embed_fn, nq.embeddirs_fn)
    assert raw_f.shape == (20, 16, 4)
    assert (raw_f - raw_n).abs().max().item() < TOL["tf32"]
    assert G.batchify(lambda t: t * 2, None)(torch.ones(3)).sum().item() == 6


def test_high_sample_config_against_oracle(G, params):
    """BASELINE configs[3] (N_samples=128 + N_importance=256: 384-sample fine pass, compositing K=12 -> generic kernel,
    sampling 4x8 register kernel) on 300 rays, test kwargs, against the oracle on the same inputs."""
    nets, nq = build_path(G, params, "tf32")
    rays = O.synthetic_rays(300, seed=11)
    want = O.render(rays, chunk=128, p_coarse=params[0], p_fine=params[1], n_samples=128, n_importance=256, lindisp=True,
                    white_bkgd=True, retraw=True)
    rc = rays.cuda()
    with torch.no_grad():
        rgb, disp, acc, depth, ex = G.render(O.H_FULL, O.W_FULL, O.FOCAL, chunk=128, rays=torch.stack([rc[:, 0:3], rc[:, 3:6]]),
                                             near=O.NEAR, far=O.FAR, use_viewdirs=True, ndc=False, retraw=True,
                                             network_query_fn=nq, perturb=False, N_importance=256, network_fine=nets[1],
                                             N_samples=128, network_fn=nets[0], white_bkgd=True, raw_noise_std=0., lindisp=True)
    assert ex["z_vals"].shape == (300, 384) and ex["raw"].shape == (300, 384, 4) and ex["weights"].shape == (300, 384)
    z = ex["z_vals"].cpu()
    assert (z[:, 1:] >= z[:, :-1]).all()
    assert (ex["rgb0"].cpu() - want["rgb0"]).abs().max().item() < TOL["tf32"]
    # rays whose far-plane sigma sits within the MLP tolerance of zero are ill-conditioned in the reference itself (the
    # 1e10 last interval makes alpha a step function of sign(sigma)): compare the others
    ok = want["raw"][:, -1, 3].abs() > 2 * TOL["tf32"]
    assert ok.float().mean() > 0.8
    assert (rgb.cpu() - want["rgb_map"])[ok].abs().max().item() < 2.5 * TOL["tf32"]
    dz = (z - want["z_vals"]).abs()
    assert dz.median().item() < 1e-4 and (dz > 5e-3).float().mean().item() < 0.02


def test_pytest_determinism_hook(G, params):
    """render_rays(pytest=True) - the reference's determinism hook (run.py:2310-2313, helpers:321-329, 380-383): every random
    tensor is replaced by numpy draws after np.random.seed(0) at each site (uniform even for the noise).  The CUDA path
    with the hook on must equal the oracle fed exactly those tensors (tests/test_oracle_vs_reference.py pins that form
    against the unmodified reference), and two calls must agree bit for bit."""
    import numpy as np
    nets, nq = build_path(G, params, "tf32")
    R = 41
    rays = O.synthetic_rays(R, seed=21)
    def draw(*shape):
        np.random.seed(0)
        return torch.Tensor(np.random.rand(*shape))
    rnd = dict(t_rand=draw(R, 64), noise0=draw(R, 64), u=draw(R, 64), noise1=draw(R, 128))
    kw = dict(lindisp=True, perturb=1.0, N_importance=64, network_fine=nets[1], white_bkgd=True, raw_noise_std=1.0, retraw=True)
    with torch.no_grad():
        a = G.render_rays(rays.cuda(), nets[0], nq, 64, pytest=True, **kw)
        b = G.render_rays(rays.cuda(), nets[0], nq, 64, pytest=True, **kw)
    want = O.render_rays(rays, params[0], params[1], 64, 64, lindisp=True, white_bkgd=True, retraw=True, **rnd)
    for k in ("rgb_map", "z_vals", "weights", "raw"):
        assert torch.equal(a[k], b[k]), k
    assert (a["rgb0"].cpu() - want["rgb0"]).abs().max().item() < TOL["tf32"]
    ok = want["raw"][:, -1, 3].abs() > 2 * TOL["tf32"]       # far-plane alpha is a step function of sign(sigma)
    assert (a["rgb_map"].cpu() - want["rgb_map"])[ok].abs().max().item() < 2.5 * TOL["tf32"]
    dz = (a["z_vals"].cpu() - want["z_vals"]).abs()
    assert dz.median().item() < 1e-5


def test_create_nerf_rejects_geometries_the_kernels_do_not_serve(G, tmp_path):
    base = dict(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_samples=64, N_importance=64, netdepth=8,
                netdepth_fine=8, netwidth=256, netwidth_fine=256, alpha_model_path=None, no_coarse=False, netchunk=65536,
                lrate=3e-3, basedir=str(tmp_path), expname="exp", ft_path=None, no_reload=True, perturb=1.0,
                white_bkgd=True, raw_noise_std=1.0, dataset_type="llff", no_ndc=True, lindisp=True, sigma_loss=False)
    (tmp_path / "exp").mkdir()
    for bad in (dict(netwidth=128), dict(netdepth_fine=4), dict(multires=6), dict(use_viewdirs=False)):
        with pytest.raises(NotImplementedError, match="shipped network"):
            G.create_nerf(types.SimpleNamespace(**dict(base, **bad)))
