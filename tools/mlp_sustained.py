"""Sustained behaviour of the MLP kernel under the power cap: back-to-back launches for ~2 s, throughput per window,
with nvidia-smi clocks / power sampled alongside."""
import os, subprocess, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G
from gbnerf_b200 import ops, _lib
if os.environ.get("AB_LIB"):          # timing A/B against another build of the library (tools only)
    _lib.LIB_PATH = os.path.abspath(os.environ["AB_LIB"])
    _lib.SIGNATURES.pop("gbn_watchdog_report", None)   # older builds do not export it
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
windows = int(sys.argv[2]) if len(sys.argv) > 2 else 20
R, S = 32768, 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision=prec).to(dev)
c2w = torch.zeros(3, 4); c2w[:, :3] = torch.eye(3); c2w[:, 3] = torch.tensor([0.1, -0.05, 0.2])
o, d = G.get_rays(756, 1008, 815.0, c2w.to(dev))
o, d = o.reshape(-1, 3)[:R].contiguous(), d.reshape(-1, 3)[:R].contiguous()
vd = d / d.norm(dim=-1, keepdim=True)
z = ops.zvals_stratified(torch.full((R, 1), 1.2, device=dev), torch.full((R, 1), 8.0, device=dev), S, True)
packed = net.packed_weights()
smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu",
                        "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
lines = []
threading.Thread(target=lambda: [lines.append((time.time(), l.strip())) for l in smi.stdout], daemon=True).start()
for _ in range(3):
    ops.mlp_forward_raw(packed, prec, vd, R, S, rays_o=o, rays_d=d, z=z)
torch.cuda.synchronize()
t_start = time.time()
for w in range(windows):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.mlp_forward_raw(packed, prec, vd, R, S, rays_o=o, rays_d=d, z=z)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    now = time.time()
    recent = [l for (t, l) in lines if t > now - 0.15]
    print(f"t={now - t_start:5.2f}s  {ms:6.3f} ms/launch  {R * S * 1186816 / ms / 1e9:7.1f} TFLOP/s   smi: {recent[-1] if recent else '-'}")
smi.terminate()
