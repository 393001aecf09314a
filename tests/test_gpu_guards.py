"""Output buffers of the per-ray kernels are carved out of a sentinel-filled arena with guard bands on both sides and
passed straight through the C ABI: no kernel may write a byte outside what it was given (odd ray counts, single-ray
last pairs, tiles cut by the end of the batch).  compute-sanitizer is not available on the GPU pool, so this is the
bounds check."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
SENT = 12345.0
GUARD = 4096     # floats on each side


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gbnerf_b200
    return gbnerf_b200


class Arena:
    def __init__(self):
        self.parts = []

    def out(self, *shape):
        n = 1
        for s in shape:
            n *= s
        n_al = (n + 3) // 4 * 4                       # keep 16-byte alignment of the payload
        buf = torch.full((GUARD + n_al + GUARD,), SENT, device="cuda", dtype=torch.float32)
        self.parts.append((buf, n))
        return buf[GUARD:GUARD + n].view(*shape)

    def check(self):
        torch.cuda.synchronize()
        for buf, n in self.parts:
            assert (buf[:GUARD] == SENT).all() and (buf[GUARD + n:] == SENT).all(), "write outside the output buffer"
            assert not (buf[GUARD:GUARD + n] == SENT).any(), "output not fully written"


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("R,S", [(1, 64), (2, 64), (77, 64), (4737 * 2 + 1, 64), (77, 128), (31, 192), (9, 40),
                                 (1, 384), (613, 384), (37, 256), (19, 512), (5, 1024)])   # 128 M: the segment-walk kernels
@pytest.mark.parametrize("noise", [False, True])
def test_composite_forward_and_backward_stay_inside(G, R, S, noise):
    g = torch.Generator().manual_seed(R + S)
    raw = torch.randn(R, S, 4, generator=g).cuda()
    z = torch.sort(torch.rand(R, S, generator=g) * 6 + 1, -1)[0].cuda()
    d = torch.randn(R, 3, generator=g).cuda()
    nz = torch.randn(R, S, generator=g).cuda() if noise else None
    ar = Arena()
    rgb, disp, acc, depth, w, alpha = ar.out(R, 3), ar.out(R), ar.out(R), ar.out(R), ar.out(R, S), ar.out(R, S)
    G._lib.call("gbn_composite_forward", ptr(raw), ptr(z), ptr(d), 3, ptr(nz), R, S, 1, ptr(rgb), ptr(disp), ptr(acc), ptr(depth),
                ptr(w), ptr(alpha), stream())
    ar.check()
    ar2 = Arena()
    g_raw = ar2.out(R, S, 4)
    grads = [torch.randn(R, 3, generator=g).cuda(), torch.randn(R, generator=g).cuda(), torch.randn(R, generator=g).cuda(),
             torch.randn(R, generator=g).cuda(), torch.randn(R, S, generator=g).cuda()]
    G._lib.call("gbn_composite_backward", ptr(raw), ptr(z), ptr(d), 3, ptr(nz), R, S, 1, 0, *[ptr(t) for t in grads], ptr(g_raw),
                stream())
    ar2.check()


@pytest.mark.parametrize("R,S,N", [(1, 64, 64), (133, 64, 64), (50, 128, 256), (41, 64, 32), (23, 37, 11), (3, 32, 32), (7, 32, 32),
                                   (1185 * 2 + 1, 64, 128), (77, 128, 64), (77, 128, 128)])
@pytest.mark.parametrize("det", [True, False])
def test_sample_merge_stays_inside(G, R, S, N, det):
    g = torch.Generator().manual_seed(R + N)
    z = torch.sort(torch.rand(R, S, generator=g) * 6 + 1, -1)[0].cuda()
    w = torch.rand(R, S, generator=g).cuda()
    u = None if det else torch.rand(R, N, generator=g).cuda()
    ar = Arena()
    smp, merged, std = ar.out(R, N), ar.out(R, S + N), ar.out(R)
    G._lib.call("gbn_sample_pdf_merge", ptr(z), ptr(w), ptr(u), R, S, N, ptr(smp), ptr(merged), ptr(std), stream())
    ar.check()
    # the extended entry point: int32 indices out (carved from the same kind of arena, viewed as int32)
    ar2 = Arena()
    inds_f, merged2, std2 = ar2.out(R, N), ar2.out(R, S + N), ar2.out(R)
    G._lib.call("gbn_sample_pdf_merge_ex", ptr(z), ptr(w), ptr(u), None, R, S, N, None, ptr(merged2), ptr(std2), ptr(inds_f), stream())
    ar2.check()
    assert torch.equal(merged2, merged)
    inds = inds_f.view(torch.int32)
    assert int(inds.min()) >= 0 and int(inds.max()) <= S - 1


@pytest.mark.parametrize("R,S", [(1, 64), (301, 64), (5, 128), (77, 32), (3, 256), (41, 50)])
@pytest.mark.parametrize("perturb", [False, True])
def test_zvals_stays_inside(G, R, S, perturb):
    g = torch.Generator().manual_seed(R + S)
    rays = torch.rand(R, 11, generator=g).cuda()
    rays[:, 6], rays[:, 7] = 1.2, 8.0
    t = torch.rand(R, S, generator=g).cuda() if perturb else None
    ar = Arena()
    z = ar.out(R, S)
    G._lib.call("gbn_zvals_stratified", ptr(rays[:, 6:7]), ptr(rays[:, 7:8]), 11, R, S, 1, ptr(t), ptr(z), stream())
    ar.check()
    assert (z[:, 1:] >= z[:, :-1]).all() and float(z.min()) >= 1.2 - 1e-5 and float(z.max()) <= 8.0 + 1e-5


@pytest.mark.parametrize("H,W", [(7, 9), (33, 50)])
def test_ray_setup_and_normals_stay_inside(G, H, W):
    c2w = torch.eye(4)[:3].contiguous().cuda()
    ar = Arena()
    rays = ar.out(H * W, 11)
    G._lib.call("gbn_pack_rays", ptr(c2w), 4, None, 0, None, 0, None, 0, None, H, W, 20.0, 0, 0, H, W, 1, 0, 1.0, 5.0, H * W,
                ptr(rays), stream())
    ar.check()
    pts = (torch.randn(2, 3, H, W, generator=torch.Generator().manual_seed(H)) + torch.tensor([0., 0., 5.]).view(1, 3, 1, 1)).cuda()
    ar2 = Arena()
    normals, minv = ar2.out(2, 3, H, W), ar2.out(2, 6, H, W)
    G._lib.call("gbn_normals_forward", ptr(pts), 2, H, W, 7, ptr(normals), ptr(minv), stream())
    ar2.check()
    ar3 = Arena()
    g_pts = ar3.out(2, 3, H, W)
    G._lib.call("gbn_normals_backward", ptr(pts), ptr(normals), ptr(minv), ptr(torch.ones_like(pts)), 2, H, W, 7, ptr(g_pts), stream())
    ar3.check()


@pytest.mark.parametrize("R,S", [(1, 1), (5, 7), (33, 31), (64, 64)])
def test_hash_grid_model_stays_inside(G, R, S):
    net = G.NeRF_TCNN(encoding="hashgrid").cuda()
    table = net.table()
    g = torch.Generator().manual_seed(R)
    rays = torch.randn(R, 9, generator=g).cuda()
    z = torch.sort(torch.rand(R, S, generator=g) * 6 + 1, -1)[0].cuda()
    ar = Arena()
    raw = ar.out(R, S, 4)
    stash = torch.full((GUARD + R * S * 32 + GUARD,), 7.0, device="cuda", dtype=torch.float16)
    G._lib.call("gbn_tcnn_forward", ptr(table), ptr(rays[:, 0:3]), ptr(rays[:, 3:6]), ptr(rays[:, 6:9]), 9, ptr(z), None, R, S,
                ptr(raw), C.c_void_p(stash.data_ptr() + 2 * GUARD), stream())
    ar.check()
    assert (stash[:GUARD] == 7.0).all() and (stash[GUARD + R * S * 32:] == 7.0).all()
    ar2 = Arena()
    g_enc, g_sig, g_col = ar2.out(R * S, 32), ar2.out(3072), ar2.out(7168)
    g_sig.zero_(), g_col.zero_()
    g_grid = torch.zeros(GUARD + 14069664 + GUARD, device="cuda")
    g_grid[:GUARD] = SENT
    g_grid[-GUARD:] = SENT
    G._lib.call("gbn_tcnn_backward", ptr(table), ptr(rays[:, 0:3]), ptr(rays[:, 3:6]), ptr(rays[:, 6:9]), 9, ptr(z), None, R, S,
                C.c_void_p(stash.data_ptr() + 2 * GUARD), ptr(torch.randn(R * S, 4, generator=g).cuda()), 128.0, ptr(g_enc),
                C.c_void_p(g_grid.data_ptr() + 4 * GUARD), ptr(g_sig), ptr(g_col), stream())
    torch.cuda.synchronize()
    assert (g_grid[:GUARD] == SENT).all() and (g_grid[-GUARD:] == SENT).all()
    for buf, n in ar2.parts:
        assert (buf[:GUARD] == SENT).all() and (buf[GUARD + n:] == SENT).all()
