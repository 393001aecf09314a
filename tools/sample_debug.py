import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G
from oracle import nerf_oracle as O
for det, S, N in ((False, 64, 64), (True, 64, 128), (False, 64, 128), (True, 64, 64)):
    g = torch.Generator().manual_seed(S * 7 + N)
    R = 131
    z = torch.sort(torch.rand(R, S, generator=g) * 6.8 + 1.2, -1)[0]
    w = torch.rand(R, S, generator=g)
    w[0] = 0.; w[1] = 0.; w[1, S // 2] = 3.
    u = None if det else torch.rand(R, N, generator=g)
    z_mid = .5 * (z[:, 1:] + z[:, :-1])
    smp = O.sample_pdf(z_mid, w[:, 1:-1], N, u)
    merged, std, got = G.ops.sample_pdf_merge(z.cuda(), w.cuda(), N, u.cuda() if u is not None else None, want_samples=True)
    got, merged = got.cpu(), merged.cpu()
    err = (got - smp).abs() / (smp.abs() * 2e-5 + 1e-6)
    bad = (err > 1).nonzero()
    print(f"det={det} S={S} N={N}: sample mismatches {bad.shape[0]} of {R * N}; first {bad[:10].tolist()}")
    if bad.shape[0]:
        r, n = bad[0].tolist()
        print("   got", got[r, max(0, n - 2):n + 3].tolist(), "want", smp[r, max(0, n - 2):n + 3].tolist(), "rows with mismatch", sorted(set(bad[:, 0].tolist()))[:20], "cols", sorted(set(bad[:, 1].tolist()))[:40])
    want_merged = torch.sort(torch.cat([z, got], -1), -1)[0]
    mb = (merged != want_merged).nonzero()
    print(f"   merged mismatches {mb.shape[0]}; first {mb[:10].tolist()}")
    if mb.shape[0]:
        r = mb[0, 0].item()
        print("   row", r, "merged", merged[r, :12].tolist(), "\n   want", want_merged[r, :12].tolist())
    sd = torch.std(smp, dim=-1, unbiased=False)
    print("   std max rel err", ((std.cpu() - sd).abs() / sd).max().item())
