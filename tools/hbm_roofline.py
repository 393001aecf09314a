"""Achieved HBM bandwidth of the bandwidth-bound kernels: CUDA-graph replay of back-to-back launches that rotate
over input sets larger than L2 (126 MB), timed with CUDA events.  Algorithmic bytes per SURVEY.md §8d."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G  # noqa: E402
from gbnerf_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def graph_time(fns, reps=5):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(fns))   # us per launch


def report(name, us, nbytes):
    gbs = nbytes / us / 1e3
    print(f"{name:58s} {us:8.1f} us  {nbytes / 1e6:8.1f} MB  {gbs:7.0f} GB/s  {100 * gbs / PEAK:5.1f} % of {PEAK:.0f}")
    return gbs


results = {}
for (R, S, N) in ((32768, 64, 64), (32768, 128, 64), (65536, 128, 256)):
    nsets = max(2, int(400e6 / (R * S * 24)) + 1)
    gg = torch.Generator().manual_seed(1)
    sets = []
    for i in range(nsets):
        sets.append(dict(raw=torch.randn(R, S, 4, generator=gg).to(dev), z=torch.sort(torch.rand(R, S, generator=gg) * 6.8 + 1.2, -1)[0].to(dev),
                         d=torch.randn(R, 3, generator=gg).to(dev), w=torch.rand(R, S, generator=gg).to(dev),
                         u=torch.rand(R, N, generator=gg).to(dev), noise=torch.randn(R, S, generator=gg).to(dev),
                         g=[torch.randn(R, 3, generator=gg).to(dev), torch.randn(R, generator=gg).to(dev) * .1,
                            torch.randn(R, generator=gg).to(dev), torch.randn(R, generator=gg).to(dev)]))
    tag = f"R={R} S={S}"
    with torch.no_grad():
        us = graph_time([lambda s=s: ops.composite(s["raw"], s["z"], s["d"], None, True) for s in sets])
        results[f"composite_fwd {tag}"] = report(f"composite forward {tag}", us, R * (24 * S + 36))
        us = graph_time([lambda s=s: ops.composite(s["raw"], s["z"], s["d"], s["noise"], True) for s in sets])
        report(f"composite forward + noise {tag}", us, R * (28 * S + 36))
        us = graph_time([lambda s=s: ops.composite_backward_raw(s["raw"], s["z"], s["d"], None, True, False, *s["g"]) for s in sets])
        results[f"composite_bwd {tag}"] = report(f"composite backward {tag}", us, R * (40 * S + 60))
        us = graph_time([lambda s=s: ops.sample_pdf_merge(s["z"], s["w"], N, None) for s in sets])
        results[f"sample_merge_det {tag} N={N}"] = report(f"sample_pdf+merge det {tag} N={N}", us, R * (8 * S + 4 * (S + N) + 4))
        us = graph_time([lambda s=s: ops.sample_pdf_merge(s["z"], s["w"], N, s["u"]) for s in sets])
        results[f"sample_merge_rnd {tag} N={N}"] = report(f"sample_pdf+merge random u {tag} N={N}", us, R * (8 * S + 4 * N + 4 * (S + N) + 4))
    del sets
a = torch.empty(1 << 28, dtype=torch.float32, device=dev)
b = torch.empty_like(a)
us = graph_time([lambda: b.copy_(a)])
report("torch copy 1 GiB (reference point)", us, 2 * a.numel() * 4)
print(json.dumps(results))
