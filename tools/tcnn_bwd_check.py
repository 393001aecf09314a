"""Per-matrix gradient error of the NeRF_TCNN backward kernels against fp32 autograd of the restatement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gbnerf_b200 as G
from oracle import tcnn_oracle as T
import test_gpu_tcnn as tt

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
gscale = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
p = tt.lively_params(P)
net = tt.load(G, p)
g = torch.Generator().manual_seed(P + 1)
inp = torch.cat([(torch.rand(P, 3, generator=g) * 2 - 1) * 4, torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1)], -1)
g_raw = torch.randn(P, 4, generator=g) * gscale
_, want = tt.oracle_grads(p, inp, g_raw)
for ls in (128.0, 1024.0, 8192.0):
    net.loss_scale = ls
    net.zero_grad()
    out = net(inp.cuda()); out.backward(g_raw.cuda())
    print("loss_scale", ls)
    for name, mod, shapes in (("sigma_net.params", net.sigma_net, T.SIGMA_SHAPES), ("color_net.params", net.color_net, T.COLOR_SHAPES)):
        got, ref = mod.params.grad.cpu(), want[name]
        off = 0
        for o, i in shapes:
            a, b = got[off:off + o * i].reshape(o, i), ref[off:off + o * i].reshape(o, i)
            print(f"  {name} [{o}x{i}]: rel {((a - b).norm() / b.norm()).item():.3e}  |ref| {b.norm().item():.3e}", end="")
            if o == 64 and i == 32 and name.startswith("color"):
                print(f"   col31 rel {((a[:, 31] - b[:, 31]).norm() / b[:, 31].norm()).item():.3e} cols16-30 rel {((a[:, 16:31] - b[:, 16:31]).norm() / b[:, 16:31].norm()).item():.3e}", end="")
            print()
            off += o * i
    a, b = net.encoder.params.grad.cpu(), want["encoder.params"]
    print(f"  grid: rel {((a - b).norm() / b.norm()).item():.3e}")
