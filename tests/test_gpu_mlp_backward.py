"""GPU parity of the MLP training path: forward stash, tcgen05 dgrad, tcgen05 wgrad (+ the two small CUDA-core
gradient kernels) against fp32 autograd of the CPU oracle, stage by stage so a failure names the layer."""
import os

import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu

N_BLOCKS = 40
H_FEAT, H_HV, H_ENC = 32, 36, 38
G_HV, G_FEAT, G_L0, G_RAW = 0, 2, 6, 38


@pytest.fixture(scope="module")
def G():
    import gbnerf_b200
    return gbnerf_b200


def decode_stash(stash, ntiles):
    """uint8 stash -> bf16 [ntiles, blocks, 128 points, 64 channels]; a block is [half (64 points)][chunk (8 channels)]
    [64 points][8 channels] (csrc/mlp_layout.h, stash_chunk_off)."""
    x = stash.view(torch.bfloat16).view(ntiles, N_BLOCKS, 2, 8, 64, 8)      # half, chunk, point, channel
    return x.permute(0, 1, 2, 4, 3, 5).reshape(ntiles, N_BLOCKS, 128, 64).float().cpu()


def rows(dec, blk0, nblk, P):
    """blocks blk0..blk0+nblk-1 of every tile -> [P, 64*nblk]"""
    t = dec[:, blk0:blk0 + nblk]                       # [T, nblk, 128, 64]
    return t.permute(0, 2, 1, 3).reshape(-1, 64 * nblk)[:P]


def oracle_forward_backward(p, emb, g_raw, emulate_bf16=False):
    """fp32 autograd of the reference network with every pre-activation gradient retained.

    ``emulate_bf16`` rounds the weights and the stored activations to bf16 at the points where the kernel does
    (straight-through), so the ReLU gates coincide with the kernel's: against plain fp32 the gates of
    near-zero pre-activations flip, which is a property of the precision, not of the backward kernels."""
    q = (lambda t: t + (t.bfloat16().float() - t).detach()) if emulate_bf16 else (lambda t: t)
    prm = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    def lin(name, v):
        w = prm[name + ".weight"]
        return torch.addmm(prm[name + ".bias"], v, q(w).t())
    emb_exact = emb
    emb = q(emb) if emulate_bf16 else emb
    x_pts, x_dir = emb[:, :63], emb[:, 63:]
    h = x_pts
    pre, post = [], []
    for i in range(8):
        a = lin(f"pts_linears.{i}", h)
        a.retain_grad()
        pre.append(a)
        h = q(torch.relu(a))
        post.append(h)
        if i == 4:
            h = torch.cat([x_pts, h], -1)
    sigma = lin("alpha_linear", h)
    feat = q(lin("feature_linear", h))
    feat.retain_grad()
    # (the default bf16 kernels feed the 27 direction columns through the tensor cores as bf16 too; only the
    # shared-memory-operand fallback keeps them as an exact fp32 per-ray bias)
    if emulate_bf16 and os.environ.get("GBNERF_MLP", "") == "ss":
        wv = prm["views_linears.0.weight"]
        av = feat @ q(wv[:, :256]).t() + emb_exact[:, 63:] @ wv[:, 256:].t() + prm["views_linears.0.bias"]
    else:
        av = lin("views_linears.0", torch.cat([feat, x_dir], -1))
    av.retain_grad()
    hv = q(torch.relu(av))
    out = torch.cat([lin("rgb_linear", hv), sigma], -1)
    out.backward(g_raw)
    return dict(out=out.detach(), post=[t.detach() for t in post], feat=feat.detach(), hv=hv.detach(),
                g_pre=[t.grad for t in pre], g_feat=feat.grad, g_hv=av.grad, grads={k: v.grad for k, v in prm.items()})


@pytest.mark.parametrize("R,S", [(64, 24), (37, 5), (700, 64)])   # the last: 350 tiles, several per CTA
def test_training_path_stage_by_stage(G, R, S):
    ops = G.ops
    torch.manual_seed(3)
    p = O.init_params(3)
    for k in p:                       # livelier activations / gradients than default init gives
        if k.endswith("weight"):
            p[k] = p[k] * 1.5
    net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True,
                 precision="bf16").cuda()
    net.load_state_dict(p)
    rays = O.synthetic_rays(R, seed=R)
    z = O.stratified_z(rays[:, 6:7], rays[:, 7:8], S, True, torch.rand(R, S, generator=torch.Generator().manual_seed(2)))
    P = R * S
    ntiles = (P + 127) // 128
    pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
    emb = torch.cat([O.posenc(pts.reshape(-1, 3), 10),
                     O.posenc(rays[:, None, 8:11].expand(R, S, 3).reshape(-1, 3), 4)], -1)
    g_raw = torch.randn(P, 4, generator=torch.Generator().manual_seed(4))
    ref = oracle_forward_backward(p, emb, g_raw, emulate_bf16=True)

    r = rays.cuda()
    stash = ops._stash(P, r.device)
    raw, ws = ops.mlp_forward_raw(net.packed_weights(), "bf16", r[:, 8:11], R, S, rays_o=r[:, 0:3], rays_d=r[:, 3:6],
                                  z=z.cuda(), stash=stash)
    assert ops.mlp_error_code(ws) == 0
    scale = ref["out"].abs().max().item()
    assert (raw.reshape(P, 4).cpu() - ref["out"]).abs().max().item() < 3e-2 * max(1.0, scale)

    # ---- forward stash ------------------------------------------------------------------------------------
    H = decode_stash(stash, ntiles)
    rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-12)).item()
    assert rel(rows(H, H_ENC, 1, P)[:, :63], emb[:, :63]) < 1e-2, "enc stash"
    for l in range(8):
        assert rel(rows(H, 4 * l, 4, P), ref["post"][l]) < 2e-2, f"h{l} stash"
    assert rel(rows(H, H_FEAT, 4, P), ref["feat"]) < 2e-2, "feature stash"
    assert rel(rows(H, H_HV, 2, P), ref["hv"]) < 2e-2, "hv stash"

    # ---- dgrad ---------------------------------------------------------------------------------------------
    shapes = [tuple(t.shape) for t in net.param_list()]
    grads, ws2, stash_g = ops.mlp_backward_raw(net.packed_weights_bwd(), g_raw.cuda(), stash, r[:, 8:11], R, S, shapes)
    torch.cuda.synchronize()
    assert ops.mlp_error_code(ws2) == 0, "dgrad watchdog"
    assert int(ws2[256:260].view(torch.int32).item()) == 0, "wgrad watchdog"
    Gd = decode_stash(stash_g, ntiles)
    assert rel(rows(Gd, G_RAW, 1, P)[:, :4], g_raw) < 1e-2, "g_raw block"
    assert rel(rows(Gd, G_HV, 2, P), ref["g_hv"]) < 3e-2, "g_hv"
    assert rel(rows(Gd, G_FEAT, 4, P), ref["g_feat"]) < 3e-2, "g_feat"
    for l in range(7, -1, -1):
        assert rel(rows(Gd, G_L0 + 4 * l, 4, P), ref["g_pre"][l]) < 4e-2, f"g_{l}"

    # ---- wgrad ---------------------------------------------------------------------------------------------
    names = list(G.ops.PARAM_ORDER)
    for i, name in enumerate(names):
        for j, kind in enumerate(("weight", "bias")):
            got, want = grads[2 * i + j].cpu(), ref["grads"][f"{name}.{kind}"]
            assert rel(got, want) < 4e-2, f"d{name}.{kind}: rel {rel(got, want):.3e}"


def test_autograd_end_to_end(G):
    """loss.backward() through render_rays with the native backward == fp32 oracle autograd (bf16 tolerance)."""
    torch.manual_seed(0)
    pc, pf = O.init_params(0), O.init_params(None)
    nets = []
    for p in (pc, pf):
        n = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True,
                   precision="bf16").cuda()
        n.load_state_dict(p)
        nets.append(n)
    e10, _ = G.get_embedder(10, 0)
    e4, _ = G.get_embedder(4, 0)
    nq = G.NetworkQuery(e10, e4, 65536)
    R, S, N = 96, 64, 64
    rays = O.synthetic_rays(R, seed=8)
    g = torch.Generator().manual_seed(5)
    rnd = dict(t_rand=torch.rand(R, S, generator=g), noise0=torch.randn(R, S, generator=g),
               u=torch.rand(R, N, generator=g), noise1=torch.randn(R, S + N, generator=g))
    tgt, tgd = torch.rand(R, 3, generator=g), torch.rand(R, generator=g)
    ret = G.render_rays(rays.cuda(), nets[0], nq, S, lindisp=True, perturb=1.0, N_importance=N, network_fine=nets[1],
                        white_bkgd=True, raw_noise_std=1.0, _randoms={k: v.cuda() for k, v in rnd.items()})
    loss = G.img2mse(ret["rgb_map"], tgt.cuda()) + G.img2mse(ret["rgb0"], tgt.cuda()) + 0.1 * G.img2mse(ret["disp_map"], tgd.cuda())
    loss.backward()
    assert G.ops.mlp_error_code(nets[1].last_workspace_bwd) == 0

    prm = [{k: v.clone().requires_grad_(True) for k, v in p.items()} for p in (pc, pf)]
    rr = O.render_rays(rays, prm[0], prm[1], S, N, lindisp=True, white_bkgd=True, **rnd)
    lref = O.reference_loss(rr, tgt, tgd, 0.1)
    lref.backward()
    assert abs(loss.item() - lref.item()) < 2e-2 * max(1.0, abs(lref.item()))
    for net, pr in zip(nets, prm):
        for name, q in net.named_parameters():
            want = pr[name].grad
            got = q.grad.cpu()
            relerr = ((got - want).norm() / (want.norm() + 1e-12)).item()
            # bf16 activations / gradients through up to nine layers, ReLU gates of near-zero pre-activations flipping:
            # measured <= 8 % on the full gradients and < 1 % on their norms (tools/grad_parity_probe.py); a corrupted
            # tile (the round-1 gate race) is far outside both
            assert relerr < 0.12, (name, relerr)
            assert abs(got.norm().item() - want.norm().item()) <= 0.02 * want.norm().item() + 1e-9, (name, got.norm().item(), want.norm().item())


def test_training_kernels_repeat_bit_exactly_from_a_cold_cache(G):
    """The stash-writing forward and the dgrad program are deterministic: the same launch repeated with the weight
    image evicted from L2 (slow first weight fills, the condition under which an issuer once passed a ring stage on
    the other issuer's phase) must reproduce its outputs bit for bit, with clean watchdog words.
    Round 1 left this test xfail: about 1 launch in 40 on one box differed in one 64-channel block of g_h7.  Cause
    (round 2, tools/dgrad_hunt.py + the chaos mode of the diagnostic library): the dgrad epilogue handed its gate
    staging buffer back to the bulk-copy producer without a generic->async proxy fence; fixed in csrc/mlp_ts.cu, so the
    test is strict again.  tools/dgrad_hunt.py is the same loop with a report of WHERE and HOW a repeat differs."""
    ops = G.ops
    torch.manual_seed(11)
    R, S = 1024, 128                                  # 1024 tiles: ~7 per CTA, 0.66 GB per stash
    P = R * S
    net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True,
                 precision="bf16").cuda()
    net.load_state_dict(O.init_params(11))
    rays = O.synthetic_rays(R, seed=3).cuda()
    z = O.stratified_z(rays[:, 6:7].cpu(), rays[:, 7:8].cpu(), S, True,
                       torch.rand(R, S, generator=torch.Generator().manual_seed(2))).cuda()
    g_raw = torch.randn(P, 4, generator=torch.Generator().manual_seed(4)).cuda()
    shapes = [tuple(t.shape) for t in net.param_list()]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def once():
        flush.zero_()                                  # evict weights and stash from the 126 MB L2
        stash = ops._stash(P, rays.device).zero_()
        raw, ws = ops.mlp_forward_raw(net.packed_weights(), "bf16", rays[:, 8:11], R, S, rays_o=rays[:, 0:3],
                                      rays_d=rays[:, 3:6], z=z, stash=stash)
        flush.zero_()
        grads, ws2, stash_g = ops.mlp_backward_raw(net.packed_weights_bwd(), g_raw, stash, rays[:, 8:11], R, S, shapes)
        torch.cuda.synchronize()
        assert ops.mlp_error_code(ws) == 0 and ops.mlp_error_code(ws2) == 0
        return raw, stash, stash_g.view(P // 128, N_BLOCKS, -1)[:, :G_RAW + 1], grads   # G block 39 is never written

    raw0, h0, g0, grads0 = once()
    for it in range(40):
        raw, h, g, grads = once()
        assert torch.equal(raw, raw0), f"forward output differs on repeat {it}"
        assert torch.equal(h, h0), f"forward stash differs on repeat {it}"
        assert torch.equal(g, g0), f"dgrad stash differs on repeat {it}"
        for a, b in zip(grads, grads0):               # wgrad sums with atomics: order-dependent rounding only
            assert (a - b).norm().item() <= 1e-3 * (b.norm().item() + 1e-12)
        del raw, h, g, grads
    assert G._lib.watchdog_report() is None

