"""The HBM-bound kernels of the render path at full chunk sizes, one after the other: CUDA-event GB/s here, and the
launch list for ncu (cold-cache device time + DRAM bytes per launch):

    python tools/hbm_kernels.py [R]                                              # event timing, rotating inputs > L2
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/hbm_launches.csv python tools/hbm_kernels.py 65536 once

Algorithmic bytes per ray are SURVEY 8d's.  Replaces tools/hbm_roofline.py of round 1 (whose
graph-replay harness read a plain 1 GiB copy at 46 % of the driver's copy figure: a captured `copy_` is a memcpy node).
Every launch goes through the C ABI directly with pre-allocated outputs, NSETS input sets are used round-robin so that
no launch finds its operands in the 126 MB L2, and a plain torch copy of 1 GiB is timed the same way as the yardstick."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G  # noqa: E402
from gbnerf_b200 import _lib  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ONCE = len(sys.argv) > 2 and sys.argv[2] == "once"
dev = torch.device("cuda:0")
PEAK = 6473.9
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    pass
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
gen = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.rand(*s, device=dev, generator=gen)
rndn = lambda *s: torch.randn(*s, device=dev, generator=gen)


def nsets(bytes_per_set):
    return 1 if ONCE else max(2, int(300e6 // bytes_per_set) + 1)


def timed(name, bytes_alg, launch, sets):
    """launch(i) enqueues one launch on input set i.  The `reps` launches are captured into ONE CUDA graph and the replay
    is timed: issued from Python through ctypes a launch costs ~10 us of host time, which a 10-30 us kernel cannot hide
    (late round 2: the small launches of this table were host-bound, not HBM-bound, in the first version)."""
    reps = 1 if ONCE else max(6, 2 * sets)
    for i in range(1 if ONCE else sets):
        launch(i % sets)
    torch.cuda.synchronize()
    if ONCE:
        return
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
            for i in range(reps):
                launch(i % sets)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    gbs = bytes_alg / us / 1e3
    print(f"{name:44s} {us:9.1f} us  {bytes_alg / 1e6:9.1f} MB  {gbs:8.1f} GB/s  {gbs / PEAK:6.1%} of measured copy peak", flush=True)


def bench_copy():
    # an elementwise KERNEL as the yardstick: inside a graph `copy_` of a contiguous buffer becomes a memcpy node, which
    # the copy engines serve at ~3 TB/s (that, not a byte-count slip, was round 1's "46 % copy")
    n = 1 << 28
    a = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
    b = torch.empty(n, dtype=torch.float32, device=dev)
    timed("torch add(x, 1) 1 GiB (read + write)", 8 * n, lambda i: torch.add(a[i % 2], 1.0, out=b), 2)


def bench_composite(S, noise, bwd):
    per_ray_in = S * (20 + (4 if noise else 0)) + 12
    k = nsets(R * per_ray_in)
    raw = [rndn(R, S, 4) for _ in range(k)]
    z = [(rnd(R, S) * 6.8 + 1.2).sort(-1)[0] for _ in range(k)]
    nz = [rndn(R, S) for _ in range(k)] if noise else [None] * k
    d = rndn(R, 3)
    rgb, disp, acc, depth, w = torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.empty(R, device=dev), torch.empty(R, device=dev), torch.empty(R, S, device=dev)
    if not bwd:
        def launch(i):
            _lib.call("gbn_composite_forward", p(raw[i]), p(z[i]), p(d), 3, p(nz[i]), R, S, 1, p(rgb), p(disp), p(acc), p(depth), p(w), None, st())
        timed(f"composite fwd S={S}{' +noise' if noise else ''}", R * (24 * S + 36 + (4 * S if noise else 0)), launch, k)
    else:
        g_rgb, g_disp, g_raw = rndn(R, 3), rndn(R), torch.empty(R, S, 4, device=dev)
        def launch(i):
            _lib.call("gbn_composite_backward", p(raw[i]), p(z[i]), p(d), 3, p(nz[i]), R, S, 1, 0, p(g_rgb), p(g_disp), None, None, None, p(g_raw), st())
        timed(f"composite bwd S={S}{' +noise' if noise else ''}", R * (36 * S + 60 + (4 * S if noise else 0)), launch, k)


def bench_sample(S, N, det):
    k = nsets(R * (8 * S + (0 if det else 4 * N)))
    z = [(rnd(R, S) * 6.8 + 1.2).sort(-1)[0] for _ in range(k)]
    w = [rnd(R, S) for _ in range(k)]
    u = [None if det else rnd(R, N) for _ in range(k)]
    merged, std = torch.empty(R, S + N, device=dev), torch.empty(R, device=dev)
    def launch(i):
        _lib.call("gbn_sample_pdf_merge", p(z[i]), p(w[i]), p(u[i]), R, S, N, None, p(merged), p(std), st())
    timed(f"sample+merge S={S} N={N} {'det' if det else 'random u'}", R * (8 * S + 4 * (S + N) + 4 + (0 if det else 4 * N)), launch, k)


def bench_small():
    rays = rnd(R, 11)
    t = rnd(R, 64)
    zz = torch.empty(R, 64, device=dev)
    timed("zvals S=64 (stratified)", R * (8 + 8 * 64), lambda i: _lib.call("gbn_zvals_stratified", p(rays[:, 6:7]), p(rays[:, 7:8]), 11, R, 64, 1, p(t), p(zz), st()), 1)
    rgb, rgb0, disp, trgb, td = rnd(R, 3), rnd(R, 3), rnd(R), rnd(R, 3), rnd(R)
    g1, g0, gd, loss = torch.empty(R, 3, device=dev), torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.zeros(1, device=dev)
    timed("loss seed", R * 80, lambda i: _lib.call("gbn_loss_seed", p(rgb), p(rgb0), p(disp), p(trgb), p(td), R, R, 0.1, p(g1), p(g0), p(gd), p(loss), st()), 1)


print(f"# tools/hbm_kernels.py R={R} {'(one launch per kernel, for ncu)' if ONCE else '(CUDA events, rotating input sets > L2)'}; measured copy peak {PEAK} GB/s")
if not ONCE:
    bench_copy()
for S in (64, 128, 384):
    bench_composite(S, False, False)
    bench_composite(S, True, False)
for S in (64, 128, 384):
    bench_composite(S, True, True)
for S, N in ((64, 64), (128, 256)):
    bench_sample(S, N, True)
    bench_sample(S, N, False)
bench_small()
