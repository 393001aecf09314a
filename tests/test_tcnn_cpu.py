"""CPU: host-side logic of the NeRF_TCNN drop-in and its restatement (no GPU, no compute through the C ABI)."""
import torch

from oracle import tcnn_oracle as T


def test_level_table_of_the_restatement():
    table, total = T.level_table()
    assert len(table) == 16 and total == 7034832
    assert [t[1] for t in table[:4]] == [16, 31, 57, 107]           # resolutions ceil(16 * pls^l - 1) + 1
    assert [t[4] for t in table] == [False] * 3 + [True] * 13       # dense while res^3 fits 2^19 entries
    assert table[-1][1] == 2048 * T.BOUND                           # finest level: 2048 cells per unit of bound
    assert all(t[2] % 8 == 0 for t in table)


def test_restatement_properties():
    p = T.init_params(0)
    p["encoder.params"] = torch.randn(T.n_grid_params(), generator=torch.Generator().manual_seed(1)) * 0.5
    x = torch.rand(64, 3, generator=torch.Generator().manual_seed(2))
    enc = T.hash_encode(x, p["encoder.params"])
    assert enc.shape == (64, 32) and torch.equal(enc, enc.half().float())
    # at a lattice point of level 0 (scale 15: x = (k - 0.5) / 15) the level-0 features are that entry itself
    k = torch.tensor([[3, 5, 7]])
    xs = (k.float() - 0.5) / 15.0
    idx = int(k[0, 0] + k[0, 1] * 16 + k[0, 2] * 256)
    want = p["encoder.params"].reshape(-1, 2)[idx].half().float()
    got = T.hash_encode(xs, p["encoder.params"])[0, :2]
    assert torch.allclose(got, want, atol=2e-3)
    sh = T.sh4(torch.nn.functional.normalize(torch.randn(100, 3, generator=torch.Generator().manual_seed(3)), dim=-1))
    assert torch.allclose((sh ** 2).sum(-1), torch.full((100,), 16 / (4 * torch.pi)), atol=2e-2)   # addition theorem
    out = T.forward(p, torch.cat([x * 4 - 2, torch.nn.functional.normalize(torch.randn(64, 3), dim=-1)], -1))
    assert out.shape == (64, 4) and torch.isfinite(out).all()


def test_module_tree_on_cpu():
    import gbnerf_b200 as G
    net = G.NeRF_TCNN(encoding="hashgrid")
    assert [k for k, _ in net.named_parameters()] == ["encoder.params", "sigma_net.params", "encoder_dir.params", "color_net.params"]
    assert net.encoder.params.numel() == T.n_grid_params()
    try:
        net(torch.zeros(4, 6))
    except ValueError as e:        # operators refuse CPU tensors: there is no CPU path
        assert "CUDA" in str(e)
    else:
        raise AssertionError("a CPU tensor went through")
