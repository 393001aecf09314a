"""GPU parity of the two-tiles-in-flight MLP forward kernel (nerf_mlp_t2_kernel, csrc/mlp_t2.cuh) against the one-tile
kernel it replaces (same packed weights, same inputs) and against the CPU oracle, at ragged sizes: a single partial tile,
an odd number of tiles (one slot of the last pair idles), pairs spread over many CTAs, and the three input forms of
gbn_mlp_forward (rays + depths, points, embedded rows); and of its stash-writing (training) form: the H stash must be
the one-tile kernel's, byte for byte.

The one-tile kernel is selected in-process by handing the library a trace buffer with tile = -1 (ts_forward keeps such
launches on it).  Tolerance: the two kernels run the same bf16 MMAs on the same operands; only the alpha / rgb heads are
summed in a different order (CUDA cores instead of a 16-column MMA), so they agree to fp32 rounding (1e-6 here);
against the fp32 oracle the bf16 bound of BASELINE.json applies (2e-2).
"""
import ctypes as C

import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gbnerf_b200
    return gbnerf_b200


@pytest.fixture(scope="module")
def net(G):
    torch.manual_seed(3)
    n = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").cuda()
    with torch.no_grad():     # biases away from their near-zero initial values
        for p in n.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    return n


def _rays(R, S, seed):
    g = torch.Generator().manual_seed(seed)
    o = (torch.rand(R, 3, generator=g) - 0.5).cuda()
    d = torch.randn(R, 3, generator=g).cuda()
    vd = d / d.norm(dim=-1, keepdim=True)
    z = (1.0 + 5.0 * torch.sort(torch.rand(R, S, generator=g), dim=-1)[0]).cuda()
    return o, d, vd, z


def _run(G, one_tile, fn):
    from gbnerf_b200 import _lib
    buf = torch.zeros(16, dtype=torch.int64, device="cuda")
    _lib.call("gbn_mlp_set_trace", C.c_void_p(buf.data_ptr()) if one_tile else None, -1)
    try:
        raw, ws = fn()
        assert G.ops.mlp_error_code(ws) == 0, "MLP kernel watchdog fired"
    finally:
        _lib.call("gbn_mlp_set_trace", None, 0)
    return raw


@pytest.mark.parametrize("R,S", [(1, 64), (3, 64), (5, 64), (300, 64), (777, 128), (4096, 64)])
def test_two_tiles_equal_one_tile_and_the_oracle(G, net, R, S):
    o, d, vd, z = _rays(R, S, R)
    packed = net.packed_weights()
    fn = lambda: G.ops.mlp_forward_raw(packed, "bf16", vd, R, S, rays_o=o, rays_d=d, z=z)
    one = _run(G, True, fn)
    two = _run(G, False, fn)
    again = _run(G, False, fn)
    assert torch.equal(two, again), "two-tile kernel does not repeat bit-exactly"
    torch.testing.assert_close(two, one, rtol=0, atol=1e-6)
    if R <= 300:
        pts = (o[:, None, :] + d[:, None, :] * z[..., None]).cpu()
        sd = O.strip_module_prefix({k: v.detach().cpu() for k, v in net.state_dict().items()})
        want = O.run_network(sd, pts, vd.cpu())
        assert (two.cpu() - want).abs().max().item() < 2e-2


def test_points_and_embedded_forms(G, net):
    R, S = 37, 64                    # 2,368 points: 19 tiles, 10 pairs
    o, d, vd, z = _rays(R, S, 11)
    packed = net.packed_weights()
    pts = (o[:, None, :] + d[:, None, :] * z[..., None]).contiguous()
    f_pts = lambda: G.ops.mlp_forward_raw(packed, "bf16", vd, R, S, pts=pts)
    torch.testing.assert_close(_run(G, False, f_pts), _run(G, True, f_pts), rtol=0, atol=1e-6)
    emb = torch.cat([O.posenc(pts.reshape(-1, 3).cpu(), 10),
                     O.posenc(vd.cpu()[:, None, :].expand(R, S, 3).reshape(-1, 3), 4)], dim=-1).cuda()
    f_emb = lambda: G.ops.mlp_forward_embedded_raw(packed, "bf16", emb)
    torch.testing.assert_close(_run(G, False, f_emb), _run(G, True, f_emb), rtol=0, atol=1e-6)


@pytest.mark.parametrize("R,S", [(1, 64), (3, 64), (37, 64), (300, 128), (4096, 64)])
def test_training_stash_is_byte_identical_to_the_one_tile_kernel(G, net, R, S):
    """The stash-writing (training) forward on the two-tile kernel: the H stash that dgrad and wgrad consume must be the
    SAME BYTES the one-tile kernel writes (same MMAs, same bias / ReLU / bf16 rounding, same block layout), for partial
    tiles and odd tile counts too (the idle slot of the last pair must not write); raw differs by the heads' summation
    order only."""
    o, d, vd, z = _rays(R, S, 100 + R)
    packed = net.packed_weights()
    out = {}
    for one_tile in (True, False):
        stash = G.ops._stash(R * S, o.device)
        stash.fill_(0x5a)
        fn = lambda: G.ops.mlp_forward_raw(packed, "bf16", vd, R, S, rays_o=o, rays_d=d, z=z, stash=stash)
        out[one_tile] = (_run(G, one_tile, fn), stash)
    torch.testing.assert_close(out[False][0], out[True][0], rtol=0, atol=1e-6)
    a, b = out[False][1], out[True][1]
    if not torch.equal(a, b):
        bad = (a != b).nonzero().reshape(-1)
        blk = (bad // 16384) % 40
        raise AssertionError(f"{bad.numel()} stash bytes differ; tiles {sorted(set((bad // (40 * 16384)).tolist()))[:8]}, "
                             f"blocks {sorted(set(blk.tolist()))}")
    # the embedded-rows form writes the same stash through the same kernel
    pts = (o[:, None, :] + d[:, None, :] * z[..., None]).contiguous()
    res = {}
    for one_tile in (True, False):
        stash = G.ops._stash(R * S, o.device)
        stash.fill_(0x5a)
        fn = lambda: G.ops.mlp_forward_raw(packed, "bf16", vd, R, S, pts=pts, stash=stash)
        res[one_tile] = (_run(G, one_tile, fn), stash)
    assert torch.equal(res[False][1], res[True][1])
