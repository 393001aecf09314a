"""ray setup of render() (run.py:1700-1736, get_rays / ndc_rays) as one kernel, against the oracle's
restatement of those lines (which tests/test_oracle_vs_reference.py pins to the reference)."""
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu

# fp32 elementwise work, same operation order as the reference; torch leaves the order of its three-term sums (the
# rotation, the norm) unspecified, so a couple of ulp are allowed; NDC subtracts nearly equal quotients -> atol.
TOL = dict(rtol=4e-7, atol=2e-7)


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gbnerf_b200
    return gbnerf_b200


def pose(seed):
    g = torch.Generator().manual_seed(seed)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    c2w = torch.zeros(3, 5)              # LLFF poses carry a fifth (h, w, f) column: only [:3,:4] may be read
    c2w[:, :3], c2w[:, 3], c2w[:, 4] = q, torch.randn(3, generator=g), 777.0
    return c2w


def oracle_batch(H, W, focal, near, far, c2w=None, rays=None, static=None, depths=None, patch=None, use_viewdirs=True, ndc=False):
    if c2w is not None:
        o, d = O.get_rays(H, W, focal, c2w[:3, :4])
        if patch is not None:
            i, j, a, b = patch
            o, d = o[i:i + a, j:j + b], d[i:i + a, j:j + b]
    else:
        o, d = rays
    vd = d
    if use_viewdirs and static is not None:
        o, d = O.get_rays(H, W, focal, static[:3, :4])
    vd = (vd / torch.norm(vd, dim=-1, keepdim=True)).reshape(-1, 3)
    if ndc:
        o, d = O.ndc_rays(H, W, focal, 1., o, d)
    cols = [o.reshape(-1, 3), d.reshape(-1, 3), near * torch.ones(vd.shape[0], 1), far * torch.ones(vd.shape[0], 1)]
    if depths is not None:
        cols.append(depths.reshape(-1, 1))
    if use_viewdirs:
        cols.append(vd)
    return torch.cat(cols, -1)


@pytest.mark.parametrize("H,W,focal", [(37, 53, 41.5), (756, 1008, 815.0)])
@pytest.mark.parametrize("ndc", [False, True])
def test_camera_rays(G, H, W, focal, ndc):
    c2w = pose(H)
    want = oracle_batch(H, W, focal, 1.2, 8.0, c2w=c2w, ndc=ndc)
    n0 = G._lib.kernel_launches()
    got = G.ops.pack_rays(H, W, focal, 1.2, 8.0, c2w=c2w.cuda(), use_viewdirs=True, ndc=ndc)
    assert G._lib.kernel_launches() - n0 == 1
    assert got.shape == (H * W, 11)
    torch.testing.assert_close(got.cpu(), want, **TOL)
    if not ndc:   # origins are copies, near/far constants: exact
        assert torch.equal(got[:, 0:3].cpu(), want[:, 0:3]) and torch.equal(got[:, 6:8].cpu(), want[:, 6:8])
        exact = (got.cpu() == want).float().mean().item()
        assert exact > 0.9, f"only {exact:.3f} of the values are bit-identical to torch"


def test_patch_static_camera_and_no_viewdirs(G):
    H, W, focal = 40, 64, 50.0
    c2w, st = pose(1), pose(2)
    patch = (5, 9, 17, 30)
    want = oracle_batch(H, W, focal, 0.5, 3.0, c2w=c2w, patch=patch)
    got = G.ops.pack_rays(H, W, focal, 0.5, 3.0, c2w=c2w.cuda(), patch=patch, use_viewdirs=True)
    torch.testing.assert_close(got.cpu(), want, **TOL)
    # (the reference cannot combine a patch with c2w_staticcam: its static rays stay full-frame and the cat fails)
    want = oracle_batch(H, W, focal, 0.5, 3.0, c2w=c2w, static=st)
    got = G.ops.pack_rays(H, W, focal, 0.5, 3.0, c2w=c2w.cuda(), c2w_staticcam=st.cuda(), use_viewdirs=True)
    torch.testing.assert_close(got.cpu(), want, **TOL)
    assert not torch.equal(got[:, 3:6], got[:, 8:11] * got[:, 3:6].norm(dim=-1, keepdim=True))   # view dirs: other camera
    want = oracle_batch(H, W, focal, 0.5, 3.0, c2w=c2w, patch=(30, 50, 20, 20), use_viewdirs=False)   # clamps to 10 x 14
    got = G.ops.pack_rays(H, W, focal, 0.5, 3.0, c2w=c2w.cuda(), patch=(30, 50, 20, 20))
    assert got.shape == (140, 8)
    torch.testing.assert_close(got.cpu(), want, **TOL)


@pytest.mark.parametrize("ndc", [False, True])
def test_ray_batch_with_depths(G, ndc):
    g = torch.Generator().manual_seed(5)
    R = 1000
    o = torch.randn(R, 3, generator=g) * 0.3
    d = torch.randn(R, 3, generator=g)
    d[:, 2] = -d[:, 2].abs() - 0.2
    dep = torch.rand(R, generator=g)
    want = oracle_batch(60, 80, 70.0, 0.0, 1.0, rays=(o, d), depths=dep, ndc=ndc)
    both = torch.stack([o, d]).cuda()                        # the [2,R,3] form train() passes (run.py:1366)
    got = G.ops.pack_rays(60, 80, 70.0, 0.0, 1.0, rays_o=both[0], rays_d=both[1], depths=dep.cuda(), use_viewdirs=True, ndc=ndc)
    assert got.shape == (R, 12)
    torch.testing.assert_close(got.cpu(), want, **TOL)
    packed = torch.cat([o, d, torch.zeros(R, 2)], -1).cuda()  # column views of a wider tensor: pitches, not copies
    got2 = G.ops.pack_rays(60, 80, 70.0, 0.0, 1.0, rays_o=packed[:, 0:3], rays_d=packed[:, 3:6], depths=dep.cuda(),
                           use_viewdirs=True, ndc=ndc)
    assert torch.equal(got, got2)
    assert G.ops.pack_rays(60, 80, 70.0, 0.0, 1.0, rays_o=both[0, :0], rays_d=both[1, :0]).shape == (0, 8)


def test_argument_errors(G):
    c2w = pose(0).cuda()
    with pytest.raises(ValueError):
        G.ops.pack_rays(10, 10, 5.0, 0., 1., c2w=c2w.cpu())
    with pytest.raises(ValueError):
        G.ops.pack_rays(10, 10, 5.0, 0., 1., c2w=c2w[:, :3])
    with pytest.raises(ValueError):
        G.ops.pack_rays(10, 10, 5.0, 0., 1., rays_o=torch.zeros(4, 3).cuda(), rays_d=torch.zeros(5, 3).cuda())
    with pytest.raises(ValueError):
        G.ops.pack_rays(10, 10, 5.0, 0., 1., c2w=c2w, depths=torch.zeros(7).cuda())


def test_render_from_pose_equals_render_from_rays(G, monkeypatch):
    """render(c2w=...) through the fused setup == render(rays=get_rays(...)) through the torch expressions."""
    torch.manual_seed(0)
    net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True).cuda()
    nq = G.NetworkQuery(G.get_embedder(10, 0)[0], G.get_embedder(4, 0)[0], 65536)
    kw = dict(network_query_fn=nq, perturb=0., N_importance=8, network_fine=net, N_samples=8, network_fn=net,
              white_bkgd=True, raw_noise_std=0., lindisp=True)
    c2w = pose(3).cuda()
    H, W, focal = 12, 20, 15.0
    with torch.no_grad():
        a = G.render(H, W, focal, c2w=c2w[:3, :4], ndc=False, near=1.2, far=8.0, use_viewdirs=True, **kw)
        monkeypatch.setattr(G.run, "_pack_rays_fused", lambda *args, **k: None)
        b = G.render(H, W, focal, c2w=c2w[:3, :4], ndc=False, near=1.2, far=8.0, use_viewdirs=True, **kw)
    assert a[0].shape == (H, W, 3) and a[1].shape == (H, W)
    for x, y in zip(a[:4], b[:4]):
        torch.testing.assert_close(x, y, rtol=2e-2, atol=5e-3, equal_nan=True)   # last-bit ray differences move bf16 roundings in the MLP
