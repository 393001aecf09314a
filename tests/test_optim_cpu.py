"""CPU: FusedAdam keeps torch.optim.Adam's behaviour for tensors it cannot step natively (here: everything)."""
import torch


def test_fused_adam_on_cpu_is_stock_adam():
    import gbnerf_b200 as G
    torch.manual_seed(0)
    a, b = torch.nn.Linear(6, 4), torch.nn.Linear(6, 4)
    b.load_state_dict(a.state_dict())
    net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)  # CPU tensors
    ref = torch.optim.Adam(a.parameters(), lr=1e-2)
    opt = G.FusedAdam(list(b.parameters()) + list(net.parameters()), lr=1e-2)
    x = torch.randn(3, 6)
    for _ in range(3):
        for m, o in ((a, ref), (b, opt)):
            o.zero_grad()
            m(x).square().sum().backward()
            o.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.equal(pa, pb)
    assert opt._plans[1] == []          # nothing on a GPU -> no native plan, and no CUDA call was attempted
    assert set(opt.state_dict()["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
