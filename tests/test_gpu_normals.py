"""depth -> normal map (run.py:2443-2474) as one kernel, forward and backward, against the oracle's
restatement (pinned to the reference by tests/test_oracle_vs_reference.py::test_depth_to_normals)."""
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gbnerf_b200
    return gbnerf_b200


def surface(H, W, seed):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    depth = 3.0 + 0.7 * xx - 0.4 * yy + 0.3 * torch.sin(5 * xx) * torch.cos(3 * yy) + 0.02 * torch.rand(H, W, generator=g)
    cam = torch.tensor([[0.9 * W, 0, W / 2.0], [0, 0.9 * W, H / 2.0], [0, 0, 1.0]])
    return depth, cam


@pytest.mark.parametrize("H,W,k", [(20, 27, 7), (47, 64, 31), (16, 16, 1), (33, 17, 15)])
def test_forward(G, H, W, k):
    depth, cam = surface(H, W, H)
    xyz = O.depth2xyz(depth, cam)
    got_xyz = G.depth2xyz_torch(depth.cuda(), cam.cuda())
    torch.testing.assert_close(got_xyz.cpu(), xyz, rtol=1e-6, atol=1e-6)
    pts = xyz.unsqueeze(0).permute(0, 3, 1, 2).contiguous()
    if k == 1:
        n = G.depth2normal_geo(pts.cuda(), k)      # M = p p^T is singular: garbage or inf/nan, as torch.linalg.inv's domain
        assert n.shape == (1, 3, H, W)
        return
    # the oracle inverts in fp32 like the reference; in double for the comparison (the kernel solves in double)
    want = O.depth2normal_geo(pts.double(), k).float()
    n0 = G._lib.kernel_launches()
    got = G.depth2normal_geo(pts.cuda(), k)
    assert G._lib.kernel_launches() - n0 == 1
    # fp32 moment sums over up to 961 points of magnitude ~10 feeding an ill-conditioned 3x3 system (condition number
    # ~1e3-1e4 for these nearly planar windows): 1e-3 relative to the normal's magnitude
    scale = want.abs().amax(dim=1, keepdim=True)
    assert ((got.cpu() - want).abs() <= 2e-3 * scale + 1e-6).all(), ((got.cpu() - want).abs() / scale).max()
    batch = torch.cat([pts, pts.flip(-1)], 0).cuda()
    both = G.depth2normal_geo(batch, k)
    torch.testing.assert_close(both[0], got[0], rtol=0, atol=0)


@pytest.mark.parametrize("H,W,k", [(20, 27, 7), (40, 36, 31)])
def test_backward(G, H, W, k):
    depth, cam = surface(H, W, 3 + H)
    tgt = torch.randn(1, 3, H, W, generator=torch.Generator().manual_seed(2))
    d_ref = depth.double().requires_grad_(True)
    pts = O.depth2xyz(d_ref, cam.double()).unsqueeze(0).permute(0, 3, 1, 2)
    loss_ref = ((O.depth2normal_geo(pts, k) + 1) / 2 * tgt.double()).sum()
    loss_ref.backward()
    d = depth.cuda().requires_grad_(True)
    xyz = G.depth2xyz_torch(d, cam.cuda())
    n = G.depth2normal_geo(xyz.unsqueeze(0).transpose(2, 3).transpose(1, 2), k)      # the reference's call, run.py:1441-1442
    loss = ((n + 1) / 2 * tgt.cuda()).sum()
    loss.backward()
    rel = (d.grad.cpu().double() - d_ref.grad).norm() / d_ref.grad.norm()
    assert rel < 5e-3, rel
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item()) + 1e-3


def test_errors(G):
    with pytest.raises(ValueError):
        G.depth2normal_geo(torch.zeros(1, 3, 8, 8), 3)               # CPU tensor
    with pytest.raises(ValueError):
        G.depth2normal_geo(torch.zeros(1, 3, 8, 8).cuda(), 4)        # even window
    with pytest.raises(ValueError):
        G.depth2normal_geo(torch.zeros(1, 2, 8, 8).cuda(), 3)
