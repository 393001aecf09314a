"""Timeline of CTA 0 for one steady-state tile of the TMEM-operand MLP kernel (gbn_mlp_set_trace).
[4j..4j+3] MMA issuer job j: start, operand waits done, weights landed, issued | [960+128*wg + 4*si ..] epilogue step si:
wait start, accumulator ready, handed over."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G
from gbnerf_b200 import ops, _lib
tile = int(sys.argv[1]) if len(sys.argv) > 1 else 3
R, S = 32768, 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").to(dev)
c2w = torch.zeros(3, 4); c2w[:, :3] = torch.eye(3); c2w[:, 3] = torch.tensor([0.1, -0.05, 0.2])
o, d = G.get_rays(756, 1008, 815.0, c2w.to(dev))
o, d = o.reshape(-1, 3)[:R].contiguous(), d.reshape(-1, 3)[:R].contiguous()
vd = d / d.norm(dim=-1, keepdim=True)
z = ops.zvals_stratified(torch.full((R, 1), 1.2, device=dev), torch.full((R, 1), 8.0, device=dev), S, True)
packed = net.packed_weights()
for _ in range(2):
    ops.mlp_forward_raw(packed, "bf16", vd, R, S, rays_o=o, rays_d=d, z=z)
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
_lib.call("gbn_mlp_set_trace", C.c_void_p(buf.data_ptr()), tile)
ops.mlp_forward_raw(packed, "bf16", vd, R, S, rays_o=o, rays_d=d, z=z)
torch.cuda.synchronize()
_lib.call("gbn_mlp_set_trace", None, 0)
t = buf.cpu().tolist()
nz = [x for x in t if x]
t0 = min(nz)
rel = lambda x: (x - t0) if x else -1
print(f"tile #{tile} of CTA 0; span {max(nz) - t0} cycles")
for j in range(96):
    a, b, c, e = t[4 * j:4 * j + 4]
    if a:
        print(f"  job {j:3d}: start {rel(a):7d}  operand-wait {b - a:6d}  w-wait {c - b:6d}  issue {e - c:5d}  end {rel(e):7d}")
for wg in range(2):
    print(f"epilogue wg{wg}: step: wait-start, acc ready, handed over")
    for si in range(24):
        row = t[960 + 128 * wg + 4 * si: 960 + 128 * wg + 4 * si + 3]
        if row[0]:
            print(f"  step {si:2d}: " + " ".join(f"{rel(x):7d}" for x in row if x))
