"""CPU-side pieces of bench.py: the reference arm (the unmodified reference staged under oracle/_ref, or the oracle port)
and the JSON contract of `--impl reference`."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_renders_and_agrees_with_the_oracle_port():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import nerf_oracle as O, ref_loader
    dt = bench.cpu_render(64)
    assert dt > 0
    assert bench.reference_kind() in ("reference", "port")
    if ref_loader.staged_available():          # same rays through both: bit-identical (the port is pinned to the reference)
        rays2 = bench.synthetic_frame_rays(0)
        idx = torch.randint(0, bench.H * bench.W, (64,), generator=torch.Generator().manual_seed(1))
        o, d = rays2[0, idx], rays2[1, idx]
        with torch.no_grad():
            rgb, *_ = bench._REF["ns"]["render"](bench.H, bench.W, bench.FOCAL, chunk=bench.CHUNK, rays=torch.stack([o, d]),
                                                 **bench._REF["kw"])
            torch.manual_seed(0)
            pc, pf = O.init_params(0), O.init_params(None)
            want = O.render(O.pack_rays(o, d, bench.NEAR, bench.FAR), chunk=bench.CHUNK, p_coarse=pc, p_fine=pf, n_samples=64,
                            n_importance=64, lindisp=True, white_bkgd=True)["rgb_map"]
        assert torch.equal(rgb, want)


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port")
    # torchrun exports OMP_NUM_THREADS=1: the arm must still use every host core
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["metric"].startswith("rays/sec") and line["higher_is_better"] is True
