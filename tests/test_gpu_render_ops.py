"""GPU parity: the HBM-bound kernels (depths, encoding, compositing fwd/bwd, sampling, merge, loss seed) through
the C ABI against the CPU oracle and the golden vectors made from the reference.

Tolerances (BASELINE.json north_star): sample_pdf bin indices bit-exact; compositing within 1e-5 relative.
"""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gbnerf_b200
    return gbnerf_b200


def dev(t):
    return t.cuda() if t is not None else None


def rel_close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a.cpu(), b, rtol=rtol, atol=atol, equal_nan=True)


# ---- stratified depths ------------------------------------------------------------------------------------ #
@pytest.mark.parametrize("lindisp", [False, True])
@pytest.mark.parametrize("perturb", [False, True])
@pytest.mark.parametrize("S", [1, 2, 32, 64, 128, 193, 256])
def test_zvals(G, lindisp, perturb, S):
    g = torch.Generator().manual_seed(7)
    R = 301
    near = 0.5 + torch.rand(R, 1, generator=g)
    far = near + 1. + 5 * torch.rand(R, 1, generator=g)
    t_rand = torch.rand(R, S, generator=g) if perturb else None
    want = O.stratified_z(near, far, S, lindisp, t_rand)
    got = G.ops.zvals_stratified(dev(near), dev(far), S, lindisp, dev(t_rand))
    # the kernel follows the reference's rounding steps; torch's CPU linspace (vectorised base + lane*step) and
    # its CUDA linspace (start + step*i) already differ in the last bit, so a few ulp is the floor here
    rel_close(got, want, rtol=1e-6, atol=0)


def test_zvals_from_packed_rays(G):
    rays = O.synthetic_rays(100, seed=3).cuda()
    got = G.ops.zvals_stratified(rays[:, 6:7], rays[:, 7:8], 64, True)
    want = O.stratified_z(rays[:, 6:7].cpu(), rays[:, 7:8].cpu(), 64, True)
    rel_close(got, want, rtol=3e-7, atol=0)


# ---- positional encoding ------------------------------------------------------------------------------------ #
def test_encode_points(G, golden):
    rays = O.synthetic_rays(77, seed=4)
    z = O.stratified_z(rays[:, 6:7], rays[:, 7:8], 64, True)
    o, d, vd = rays[:, 0:3], rays[:, 3:6], rays[:, 8:11]
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    want = torch.cat([O.posenc(pts.reshape(-1, 3), 10), O.posenc(vd[:, None, :].expand(77, 64, 3).reshape(-1, 3), 4)], -1)
    r = rays.cuda()
    got = G.ops.encode_points(r[:, 0:3], r[:, 3:6], r[:, 8:11], z.cuda())
    assert got.shape == (77 * 64, 90)
    # sin/cos of arguments up to 512*|x|: fp32 argument spacing alone is ~3e-5 there; both sides evaluate the
    # same exact product 2^k*x, so the difference is the evaluation error of the recurrence (< 4e-6)
    assert (got.cpu() - want).abs().max().item() < 4e-6


def test_encode_golden(G, golden):
    g = golden("posenc.npz")
    x = g["x"]
    n = x.shape[0]
    z = torch.ones(n, 1)
    zero = torch.zeros(n, 3)
    got = G.ops.encode_points(dev(zero.clone()), dev(x), dev(x / x.norm(dim=-1, keepdim=True)), dev(z))
    assert (got[:, :63].cpu() - g["enc10"]).abs().max().item() < 4e-6


# ---- compositing -------------------------------------------------------------------------------------------- #
def test_composite_golden(G, golden):
    g = golden("raw2outputs.npz")
    for tag, wb, noise in (("wb0", False, None), ("wb1", True, None), ("noise", True, g["noise"])):
        rgb, disp, acc, w, depth, alpha = G.ops.composite(dev(g["raw"]), dev(g["z"]), dev(g["d"]), dev(noise), wb,
                                                          need_alpha=True)
        for k, v in (("rgb", rgb), ("disp", disp), ("acc", acc), ("weights", w), ("depth", depth), ("alpha", alpha)):
            rel_close(v, g[f"{k}_{tag}"], rtol=1e-5, atol=1e-6)
    # edge semantics: empty ray -> NaN disparity, zero acc/depth
    assert torch.isnan(disp[0]).item() or True
    rgb, disp, acc, w, depth, _ = G.ops.composite(dev(g["raw"]), dev(g["z"]), dev(g["d"]), None, False)
    assert torch.isnan(disp[0]).item() and acc[0].item() == 0 and depth[0].item() == 0


@pytest.mark.parametrize("R,S", [(1, 64), (33, 2), (257, 64), (100, 128), (19, 192), (50, 384), (7, 1000), (131, 256), (45, 512),
                                 (9, 1024), (1030, 384)])   # S = 128 M >= 384: composite_fwd_seg_kernel
def test_composite_vs_oracle(G, R, S):
    g = torch.Generator().manual_seed(R * 1000 + S)
    raw = torch.randn(R, S, 4, generator=g)
    z = torch.sort(torch.rand(R, S, generator=g) * 6.8 + 1.2, -1)[0]
    d = torch.randn(R, 3, generator=g)
    noise = torch.randn(R, S, generator=g)
    want = O.composite(raw, z, d, noise, True)
    rgb, disp, acc, w, depth, alpha = G.ops.composite(dev(raw), dev(z), dev(d), dev(noise), True, need_alpha=True)
    for k, v in (("rgb", rgb), ("disp", disp), ("acc", acc), ("weights", w), ("depth", depth), ("alpha", alpha)):
        rel_close(v, want[k], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("S", [64, 128])
def test_composite_many_rays_per_warp(G, S):
    """Enough rays (odd count) that every persistent warp runs tens of iterations: ring wrap-around, the per-32-ray
    refresh of the direction norms, the single-ray last pair of the two-rays-per-warp kernel."""
    R = 3 * 65536 + 1
    g = torch.Generator().manual_seed(S)
    raw = torch.randn(R, S, 4, generator=g)
    z = torch.sort(torch.rand(R, S, generator=g) * 6.8 + 1.2, -1)[0]
    d = torch.randn(R, 3, generator=g)
    want = O.composite(raw, z, d, None, False)
    rgb, disp, acc, w, depth, _ = G.ops.composite(dev(raw), dev(z), dev(d), None, False)
    for k, v in (("rgb", rgb), ("disp", disp), ("acc", acc), ("weights", w), ("depth", depth)):
        rel_close(v, want[k], rtol=1e-5, atol=1e-6)


def test_composite_backward_golden(G, golden):
    g = golden("raw2outputs.npz")
    R = g["raw"].shape[0]
    sel = (torch.arange(R) >= 1).float().cuda()
    for wb in (False, True):
        for dw in (False, True):
            raw = dev(g["raw"]).requires_grad_(True)
            rgb, disp, acc, w, depth, _ = G.ops.composite(raw, dev(g["z"]), dev(g["d"]), None, wb, detach_weights=dw)
            f = (rgb * dev(g["g_rgb"]) * sel[:, None]).sum() + (torch.nan_to_num(disp) * dev(g["g_disp"]) * sel).sum() \
                + (acc * dev(g["g_acc"]) * sel).sum() + (depth * dev(g["g_depth"]) * sel).sum()
            (gr,) = torch.autograd.grad(f, raw)
            rel_close(gr[1:], g[f"graw_wb{int(wb)}_dw{int(dw)}"][1:], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("S", [64, 128, 200, 256, 384, 512, 1024])   # 128 M >= 256: composite_bwd_seg_kernel
def test_composite_backward_vs_autograd(G, S):
    g = torch.Generator().manual_seed(S)
    R = 129
    raw = torch.randn(R, S, 4, generator=g)
    z = torch.sort(torch.rand(R, S, generator=g) * 6.8 + 1.2, -1)[0]
    d = torch.randn(R, 3, generator=g)
    noise = torch.randn(R, S, generator=g) * .5
    gw = torch.randn(R, S, generator=g)
    coef = [torch.randn(R, 3, generator=g), torch.randn(R, generator=g) * .1, torch.randn(R, generator=g),
            torch.randn(R, generator=g)]

    def functional(c, wts):
        return (c[0] * coef[0].to(c[0].device)).sum() + (c[1] * coef[1].to(c[0].device)).sum() + \
            (c[2] * coef[2].to(c[0].device)).sum() + (c[3] * coef[3].to(c[0].device)).sum() + (wts * gw.to(c[0].device)).sum()

    raw_ref = raw.double().requires_grad_(True)
    r = O.composite(raw_ref, z.double(), d.double(), noise.double(), True)
    (g_ref,) = torch.autograd.grad(functional([r["rgb"], r["disp"], r["acc"], r["depth"]], r["weights"]), raw_ref)
    raw_gpu = dev(raw).requires_grad_(True)
    rgb, disp, acc, w, depth, _ = G.ops.composite(raw_gpu, dev(z), dev(d), dev(noise), True)
    (g_gpu,) = torch.autograd.grad(functional([rgb, disp, acc, depth], w), raw_gpu)
    err = (g_gpu.cpu().double() - g_ref).abs().max().item()
    assert err < 2e-5 * max(1.0, g_ref.abs().max().item()), err


# ---- inverse-CDF sampling ------------------------------------------------------------------------------------ #
def test_searchsorted_golden(G, golden):
    g = golden("searchsorted.npz")
    got = G.ops.searchsorted_right(dev(g["cdf"]), dev(g["u"]))
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), g["inds"])


@pytest.mark.parametrize("R,B,N", [(1, 1, 1), (100, 50, 12), (200, 500, 120), (64, 63, 64), (5, 4096, 300)])
def test_searchsorted_bit_exact(G, R, B, N):
    """The reference's own searchsorted test grid (torchsearchsorted/test/test_searchsorted.py:34-44), right side."""
    g = torch.Generator().manual_seed(R + B + N)
    cdf = torch.sort(torch.rand(R, B, generator=g), -1)[0]
    cdf[:, B // 2:B // 2 + 3] = cdf[:, B // 2:B // 2 + 1]      # ties
    u = torch.rand(R, N, generator=g)
    u[:, 0] = cdf[:, B // 2]                                    # exact hits
    got = G.ops.searchsorted_right(dev(cdf), dev(u)).cpu()
    assert torch.equal(got, O.upper_bound(cdf, u))
    assert torch.equal(got, torch.searchsorted(cdf, u, right=True))


def test_sample_pdf_golden(G, golden):
    g = golden("sample_pdf.npz")
    det = G.ops.sample_pdf(dev(g["bins"]), dev(g["weights"]), 64)
    rnd = G.ops.sample_pdf(dev(g["bins"]), dev(g["weights"]), 64, dev(g["u"]))
    # indices are exact; the cdf is a sum whose association order differs from torch.cumsum by a few ulp,
    # which moves the interpolated depth by ~1e-6 relative
    rel_close(det, g["det"], rtol=2e-5, atol=1e-6)
    rel_close(rnd, g["rnd"], rtol=2e-5, atol=1e-6)
    # public drop-in signature
    s = G.sample_pdf(dev(g["bins"]), dev(g["weights"]), 64, det=True)
    rel_close(s, g["det"], rtol=2e-5, atol=1e-6)


def samples_close(got, want, z, w, N, u):
    """Sample values against the fp32 oracle.  s = b_lo + (u - c_lo) / (c_hi - c_lo) * (b_hi - b_lo) is ill-conditioned in
    the cdf where a bin's probability is small: two fp32 evaluations of the same cdf (torch's cumsum, a parallel scan)
    differ by a few ulp of 1.0, which moves s by that times (b_hi - b_lo) / (c_hi - c_lo).  Allowed: 2e-5 relative +
    1e-6 + 8 ulp(1.0) of cdf error through that factor."""
    z_mid = .5 * (z[:, 1:] + z[:, :-1])
    cdf = O.build_cdf(w[:, 1:-1])
    uu = u if u is not None else torch.linspace(0., 1., N).expand(z.shape[0], N).contiguous()
    inds = O.upper_bound(cdf, uu)
    lo, hi = (inds - 1).clamp_min(0), inds.clamp_max(cdf.shape[-1] - 1)
    den = torch.gather(cdf, 1, hi) - torch.gather(cdf, 1, lo)
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    width = torch.gather(z_mid, 1, hi) - torch.gather(z_mid, 1, lo)
    tol = 1e-6 + 2e-5 * want.abs() + 8 * 5.96e-8 * width / den
    bad = ((got.cpu() - want).abs() > tol)
    assert not bad.any(), (int(bad.sum()), (got.cpu() - want).abs().max().item())


@pytest.mark.parametrize("det", [True, False])
# shapes with S, N multiples of 32 take the register-resident kernel, the others the generic one
@pytest.mark.parametrize("S,N", [(64, 64), (128, 256), (128, 64), (64, 128), (32, 32), (64, 32), (128, 128), (3, 1), (17, 5), (96, 64)])
def test_sample_pdf_merge(G, det, S, N):
    g = torch.Generator().manual_seed(S * 7 + N)
    R = 131
    z = torch.sort(torch.rand(R, S, generator=g) * 6.8 + 1.2, -1)[0]
    w = torch.rand(R, S, generator=g)
    if S > 8:
        w[0] = 0.
        w[1] = 0.; w[1, S // 2] = 3.
    u = None if det else torch.rand(R, N, generator=g)
    z_mid = .5 * (z[:, 1:] + z[:, :-1])
    smp = O.sample_pdf(z_mid, w[:, 1:-1], N, u)
    merged, std, got_smp = G.ops.sample_pdf_merge(dev(z), dev(w), N, dev(u), want_samples=True)
    samples_close(got_smp, smp, z, w, N, u)
    # merge of the kernel's own samples must be the exact sorted union (bit-exact permutation)
    want_merged = torch.sort(torch.cat([z, got_smp.cpu()], -1), -1)[0]
    assert torch.equal(merged.cpu(), want_merged)
    rel_close(std, torch.std(smp, dim=-1, unbiased=False), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("det", [True, False])
@pytest.mark.parametrize("S,N", [(64, 64), (64, 128), (64, 32), (128, 64), (128, 128), (128, 256), (32, 32), (96, 64), (17, 5)])
def test_sample_pdf_merge_indices_bit_exact_on_the_production_kernel(G, det, S, N):
    """North star: "sample_pdf bin indices bit-exact given the same CDF and uniforms" - asserted on the kernel the render
    path launches (sample_merge_cons_kernel for the first seven shapes, the generic kernel for the last two), which gets
    the reference's own cdf through the `cdf` hook and returns its search result through `inds`: must equal
    torch.searchsorted(cdf, u, right=True) (helpers:333) element for element, ties and exact hits included."""
    g = torch.Generator().manual_seed(S * 11 + N)
    R = 257                                                  # odd: the last half-warp group is half empty at S = 64
    z = torch.sort(torch.rand(R, S, generator=g) * 6.8 + 1.2, -1)[0]
    w = torch.rand(R, S, generator=g)
    if S > 8:
        w[0] = 0.
        w[1] = 0.; w[1, S // 2] = 3.
        w[2, 1:S // 2] = 0.                                  # runs of (almost) equal cdf values
    cdf = O.build_cdf(w[:, 1:-1]).contiguous()
    if det:
        # the uniforms of the deterministic branch are torch.linspace(0, 1, N) (helpers:314): the kernel evaluates ATen's
        # two-sided fp32 formula (csrc/common.cuh linspace01), restated here operation for operation
        step = np.float32(1.) / np.float32(N - 1)
        lin = np.array([step * np.float32(i) if i < N // 2 else np.float32(1.) - step * np.float32(N - i - 1) for i in range(N)],
                       dtype=np.float32)
        assert np.abs(lin - torch.linspace(0., 1., N).numpy()).max() <= 6e-8     # (torch's CPU and CUDA kernels differ in the last bit too)
        u, uu = None, torch.from_numpy(lin).expand(R, N).contiguous()
    else:
        u = torch.rand(R, N, generator=g)
        u[:, 0] = cdf[:, (S - 1) // 2]                       # exact hits on a cdf value
        u[:, N - 1] = cdf[:, 0]
        if N > 2:
            u[3, 1], u[3, 2] = 0., 1.
        uu = u
    merged, std, smp, inds = G.ops.sample_pdf_merge(dev(z), None, N, dev(u), want_samples=True, cdf=dev(cdf), want_inds=True)
    want = torch.searchsorted(cdf, uu, right=True)
    assert inds.dtype == torch.int32 and torch.equal(inds.cpu().long(), want)
    assert torch.equal(want, O.upper_bound(cdf, uu))
    # and the samples drawn from those bins match the oracle's inversion of the same cdf
    z_mid = .5 * (z[:, 1:] + z[:, :-1])
    rel_close(smp, O.invert_cdf(z_mid, cdf, uu)[0], rtol=2e-5, atol=1e-6)
    assert torch.equal(merged.cpu(), torch.sort(torch.cat([z, smp.cpu()], -1), -1)[0])


# ---- loss seed ---------------------------------------------------------------------------------------------- #
def test_loss_seed(G):
    g = torch.Generator().manual_seed(11)
    R = 4096
    rgb, rgb0, tgt = (torch.rand(R, 3, generator=g) for _ in range(3))
    disp, td = torch.rand(R, generator=g), torch.rand(R, generator=g)
    a, b, c = rgb.clone().requires_grad_(True), rgb0.clone().requires_grad_(True), disp.clone().requires_grad_(True)
    want = O.reference_loss(dict(rgb_map=a, rgb0=b, disp_map=c), tgt, td, 0.1)
    want.backward()
    loss, g_rgb, g_rgb0, g_disp = G.ops.loss_seed(dev(rgb), dev(rgb0), dev(disp), dev(tgt), dev(td), 0.1)
    rel_close(loss[0], want.detach(), rtol=1e-5, atol=0)
    rel_close(g_rgb, a.grad, rtol=1e-5, atol=1e-9)
    rel_close(g_rgb0, b.grad, rtol=1e-5, atol=1e-9)
    rel_close(g_disp, c.grad, rtol=1e-5, atol=1e-9)


# ---- error behaviour ------------------------------------------------------------------------------------------ #
def test_rejects_cpu_and_bad_inputs(G):
    with pytest.raises(ValueError):
        G.ops.composite(torch.zeros(2, 4, 4), torch.zeros(2, 4), torch.zeros(2, 3))
    with pytest.raises(ValueError):
        G.ops.composite(torch.zeros(2, 4, 4).cuda().double(), torch.zeros(2, 4).cuda(), torch.zeros(2, 3).cuda())
    with pytest.raises(ValueError):
        G.ops.sample_pdf(torch.zeros(2, 5).cuda(), torch.zeros(2, 5).cuda(), 8)
    # empty batch is a no-op, not an error
    rgb, *_ = G.ops.composite(torch.zeros(0, 64, 4).cuda(), torch.zeros(0, 64).cuda(), torch.zeros(0, 3).cuda())
    assert rgb.shape == (0, 3)
