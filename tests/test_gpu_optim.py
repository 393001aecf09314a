"""the optimizer step (run.py:1529 on the Adam of run.py:2065) as one fused kernel per network.
Checked against torch.optim.Adam itself (the reference's optimizer) on the same gradients."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gbnerf_b200
    return gbnerf_b200


def make_net(G, seed):
    torch.manual_seed(seed)
    return G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True,
                  precision="bf16").cuda()


def set_grads(net_a, net_b, seed, scale=1e-2):
    g = torch.Generator(device="cuda").manual_seed(seed)
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        gr = torch.randn(pa.shape, generator=g, device="cuda") * scale
        pa.grad, pb.grad = gr.clone(), gr.clone()


def test_fused_adam_matches_torch_adam_and_repacks_in_place(G):
    ops = G.ops
    a, b = make_net(G, 0), make_net(G, 0)
    ref = torch.optim.Adam(a.parameters(), lr=3e-3, betas=(0.9, 0.999))
    opt = G.FusedAdam(b.parameters(), lr=3e-3, betas=(0.9, 0.999))
    assert isinstance(opt, torch.optim.Adam)
    in_place = G._lib.load().gbn_mlp_variant() == 1   # env GBNERF_MLP=ss: another image layout, re-packed lazily
    n0 = G._lib.kernel_launches()
    b.packed_weights(), b.packed_weights_bwd()
    n_pack = G._lib.kernel_launches() - n0
    for it in range(4):
        set_grads(a, b, 10 + it)
        if it == 2:   # the learning-rate decay of run.py:1540-1544 assigns param_group['lr']
            for o in (ref, opt):
                for gparam in o.param_groups:
                    gparam["lr"] = 3e-3 * 0.1 ** (it / 250000)
        ref.step()
        fwd_before = b.packed_weights().clone()
        n1 = G._lib.kernel_launches()
        opt.step()
        assert G._lib.kernel_launches() - n1 == 1, "one launch per network"
        for (name, pa), pb in zip(a.named_parameters(), b.parameters()):
            torch.testing.assert_close(pb, pa, rtol=2e-6, atol=1e-8, msg=lambda m: f"step {it} {name}: {m}")
            sa, sb = ref.state[pa], opt.state[pb]
            torch.testing.assert_close(sb["exp_avg"], sa["exp_avg"], rtol=2e-6, atol=1e-12)
            torch.testing.assert_close(sb["exp_avg_sq"], sa["exp_avg_sq"], rtol=2e-6, atol=1e-14)
            assert float(sb["step"]) == float(sa["step"]) == it + 1
        # the weight images were patched in place: byte-identical to a fresh re-pack of the new parameters
        n2 = G._lib.kernel_launches()
        fwd, bwd = b.packed_weights(), b.packed_weights_bwd()
        assert not in_place or G._lib.kernel_launches() == n2, "no re-pack pass after a fused step"
        # (re-packing over a copy: bytes the packer never writes - alignment gaps - keep their old content)
        assert torch.equal(fwd, ops.prepack_weights(b.param_list(), "bf16", out=fwd.clone()))
        assert torch.equal(bwd, ops.prepack_weights(b.param_list(), "bf16_bwd", out=bwd.clone()))
        changed = (fwd != fwd_before).float().mean().item()
        assert changed > 0.2, "the image must actually change with the parameters"
    assert n_pack == 2


def test_state_dict_is_interchangeable_with_torch_adam(G):
    a, b = make_net(G, 1), make_net(G, 1)
    ref = torch.optim.Adam(a.parameters(), lr=1e-3)
    opt = G.FusedAdam(b.parameters(), lr=1e-3)
    for it in range(2):
        set_grads(a, b, 20 + it)
        ref.step(), opt.step()
    sd = copy.deepcopy(opt.state_dict())
    assert sd["state"].keys() == ref.state_dict()["state"].keys()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    # fused -> stock and stock -> fused, then one more step each: still the same trajectory
    c, d = make_net(G, 1), make_net(G, 1)
    c.load_state_dict(b.state_dict()), d.load_state_dict(a.state_dict())
    ref2 = torch.optim.Adam(c.parameters(), lr=1e-3)
    ref2.load_state_dict(sd)
    opt2 = G.FusedAdam(d.parameters(), lr=1e-3)
    opt2.load_state_dict(copy.deepcopy(ref.state_dict()))
    set_grads(c, d, 30)
    ref2.step(), opt2.step()
    for pc, pd in zip(c.parameters(), d.parameters()):
        torch.testing.assert_close(pd, pc, rtol=2e-6, atol=1e-8)


def test_foreign_parameters_take_the_stock_path(G):
    a, b = make_net(G, 2), make_net(G, 2)
    torch.manual_seed(3)
    ea, eb = torch.nn.Linear(5, 7).cuda(), torch.nn.Linear(5, 7).cuda()
    eb.load_state_dict(ea.state_dict())
    ref = torch.optim.Adam(list(a.parameters()) + list(ea.parameters()), lr=2e-3)
    opt = G.FusedAdam(list(b.parameters()) + list(eb.parameters()), lr=2e-3)
    for it in range(2):
        set_grads(a, b, 40 + it)
        set_grads(ea, eb, 50 + it)
        ref.step(), opt.step()
    for pa, pb in zip(list(a.parameters()) + list(ea.parameters()), list(b.parameters()) + list(eb.parameters())):
        torch.testing.assert_close(pb, pa, rtol=2e-6, atol=1e-8)
    assert all(p.grad is not None for p in b.parameters()), "gradients are handed back after the stock pass"
    # a network missing a gradient is left to torch (which skips grad-less tensors)
    next(b.parameters()).grad = None
    opt.step()


def test_create_nerf_trains_with_the_fused_optimizer(G, tmp_path):
    import argparse
    from oracle import nerf_oracle as O
    args = argparse.Namespace(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_samples=16, N_importance=16,
                              netdepth=8, netdepth_fine=8, netwidth=256, netwidth_fine=256, alpha_model_path=None,
                              no_coarse=False, netchunk=65536, lrate=5e-4, basedir=str(tmp_path), expname="e", ft_path=None,
                              no_reload=True, perturb=1.0, white_bkgd=True, raw_noise_std=0.0, dataset_type="llff",
                              no_ndc=True, lindisp=True, sigma_loss=False)
    (tmp_path / "e").mkdir()
    torch.manual_seed(0)
    kw, _, _, grad_vars, optimizer = G.create_nerf(args)
    assert isinstance(optimizer, G.FusedAdam) and len(grad_vars) == 48
    kw.update(near=1.2, far=8.0)
    rays = O.synthetic_rays(256, seed=3).cuda()
    rays2 = torch.stack([rays[:, 0:3], rays[:, 3:6]]).contiguous()
    tgt = torch.full((256, 3), 0.25, device="cuda")
    losses = []
    for it in range(12):
        rgb, disp, acc, depth, ex = G.render(756, 1008, 815.0, chunk=32768, rays=rays2, **kw)
        loss = G.img2mse(rgb, tgt) + G.img2mse(ex["rgb0"], tgt)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        losses.append(loss.item())
    assert losses[-1] < 0.7 * losses[0], losses
