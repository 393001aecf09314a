"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this).

CPU/torch restatement of the ``NeRF_TCNN`` model of the reference (DS_NeRF/run_nerf_helpers_tcnn.py:13-117), SURVEY.md
§8f rank 1 / BASELINE config 5.

**Parity unpinned.**  The arithmetic lives in the third-party module ``tinycudann`` (NVlabs/tiny-cuda-nn, torch
bindings), which the reference installs unpinned from git (requirements_df.txt) and which is absent from
/root/reference and from this image; no reference test or golden vector touches it.  What follows restates
tiny-cuda-nn's *published* algorithm (Mueller et al., "Instant Neural Graphics Primitives", and the documented
behaviour of its ``HashGrid`` / ``SphericalHarmonics`` encodings and ``FullyFusedMLP`` network) for exactly the
configuration the reference's call sites build (run_nerf_helpers_tcnn.py:40-50, 52-62, 68-74, 78-88) and the forward of
lines 90-117.  Results can only be checked for self-consistency against the CUDA kernels of this repository.

Conventions restated here:
* HashGrid: level l has scale ``s_l = base * per_level_scale**l - 1`` and resolution ``ceil(s_l) + 1``; a level
  owns ``min(round_up(res**3, 8), 2**log2_hashmap_size)`` entries of ``n_features`` values; position
  ``x*s_l + 0.5`` is split into integer cell and fraction; the 8 cell corners are blended tri-linearly; a corner's
  entry is its dense index ``x + y*res + z*res**2`` when the level is dense, else the spatial hash
  ``x*1 ^ y*2654435761 ^ z*805459861`` (uint32), modulo the level's size.  Output feature ``2*l + f``.
* SphericalHarmonics degree 4 on ``2*d - 1`` (the encoding takes [0,1] inputs): the 16 real SH basis values.
* FullyFusedMLP: bias-free linears, ReLU on hidden layers, none on the output; input and output widths padded to a
  multiple of 16.  The reference's 31-wide colour input (16 SH + 15 geometry features) is padded to 32 by
  tiny-cuda-nn's input stage with a constant ONE (so that column of the first matrix acts as a bias); padded output
  columns are computed and dropped.  Parameters are one flat fp32 vector per network, matrices row-major
  ``[out, in]`` in layer order; compute in fp16 (restated: fp16 weights and activations, fp32 accumulation, where
  tiny-cuda-nn accumulates in fp16 - one more reason results here are "same algorithm", not "same bits").
"""
import math

import numpy as np
import torch

N_LEVELS, N_FEATURES, LOG2_HASHMAP, BASE_RES = 16, 2, 19, 16
BOUND = 100
PER_LEVEL_SCALE = float(np.exp2(np.log2(2048 * BOUND / 16) / (16 - 1)))   # run_nerf_helpers_tcnn.py:38
PRIMES = (1, 2654435761, 805459861)
HIDDEN, GEO_FEAT = 64, 15
SIGMA_SHAPES = ((64, 32), (16, 64))                 # sigma_net: 32 -> 64 -> 1 + 15   (run_nerf_helpers_tcnn.py:52-62)
COLOR_SHAPES = ((64, 32), (64, 64), (16, 64))       # color_net: 16 + 15 (+1 pad) -> 64 -> 64 -> 3 (+13 pad) (:78-88)


def level_table():
    """[(scale, resolution, entries, offset, hashed)] per level and the total entry count."""
    out, off = [], 0
    log2_pls = np.float32(np.log2(np.float32(PER_LEVEL_SCALE)))
    for l in range(N_LEVELS):
        scale = float(np.float32(np.exp2(np.float32(l) * log2_pls)) * np.float32(BASE_RES) - np.float32(1.0))
        res = int(math.ceil(scale)) + 1
        dense = res ** 3
        n = min((min(dense, 2 ** 31 - 1) + 7) // 8 * 8, 1 << LOG2_HASHMAP)
        out.append((scale, res, n, off, dense > n))
        off += n
    return out, off


def n_grid_params():
    return level_table()[1] * N_FEATURES


def n_mlp_params(shapes):
    return sum(o * i for o, i in shapes)


def init_params(seed=0):
    """Parameters with tiny-cuda-nn's documented initialisation (grid U(-1e-4, 1e-4); Xavier-uniform matrices), as
    the four flat fp32 vectors the torch bindings expose (``encoder.params`` ...)."""
    g = torch.Generator().manual_seed(seed)
    p = {"encoder.params": (torch.rand(n_grid_params(), generator=g) * 2 - 1) * 1e-4, "encoder_dir.params": torch.zeros(0)}
    for name, shapes in (("sigma_net.params", SIGMA_SHAPES), ("color_net.params", COLOR_SHAPES)):
        mats = []
        for o, i in shapes:
            lim = math.sqrt(6.0 / (o + i))
            mats.append(((torch.rand(o, i, generator=g) * 2 - 1) * lim).reshape(-1))
        p[name] = torch.cat(mats)
    return p


def _h(t):
    """Round to fp16 and carry on in fp32 (the kernels keep fp16 storage, fp32 accumulation)."""
    return t.half().float()


def hash_encode(x, grid_params):
    """x [P,3] in [0,1] -> [P,32]; grid_params flat fp32 [entries*2] (rounded to fp16 like the device table)."""
    table, _ = level_table()
    grid = _h(grid_params).reshape(-1, N_FEATURES)
    feats = []
    for scale, res, n, off, hashed in table:
        pos = x * np.float32(scale) + np.float32(0.5)
        cell = torch.floor(pos)
        frac = pos - cell
        cell = cell.to(torch.int64)
        acc = torch.zeros(x.shape[0], N_FEATURES, dtype=torch.float32, device=x.device)
        for corner in range(8):
            w = torch.ones(x.shape[0], dtype=torch.float32, device=x.device)
            c = []
            for dim in range(3):
                if corner & (1 << dim):
                    w = w * frac[:, dim]
                    c.append(cell[:, dim] + 1)
                else:
                    w = w * (1 - frac[:, dim])
                    c.append(cell[:, dim])
            if hashed:
                idx = ((c[0] * PRIMES[0]) & 0xFFFFFFFF) ^ ((c[1] * PRIMES[1]) & 0xFFFFFFFF) ^ ((c[2] * PRIMES[2]) & 0xFFFFFFFF)
            else:
                idx = (c[0] + c[1] * res + c[2] * res * res) & 0xFFFFFFFF
            idx = idx % n
            acc = acc + w[:, None] * grid[off + idx]
        feats.append(acc)
    return _h(torch.cat(feats, -1))


def sh4(d):
    """d [P,3] unit-ish directions in [-1,1] -> the 16 real spherical-harmonics basis values of degree 4."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    out = [
        torch.full_like(x, 0.28209479177387814),
        -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
        1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999,
        -1.0925484305920792 * xz, 0.54627421529603959 * x2 - 0.54627421529603959 * y2,
        0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0),
        0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
        0.59004358992664352 * x * (-x2 + 3.0 * y2),
    ]
    return _h(torch.stack(out, -1))


def mlp(x, flat, shapes):
    """Bias-free MLP, ReLU between layers, fp16 weights/activations with fp32 accumulation."""
    off = 0
    for li, (o, i) in enumerate(shapes):
        w = _h(flat[off:off + o * i].reshape(o, i))
        off += o * i
        x = x @ w.t()
        if li < len(shapes) - 1:
            x = torch.relu(x)
        x = _h(x)
    return x


def forward(p, inp):
    """NeRF_TCNN.forward (run_nerf_helpers_tcnn.py:90-117): inp [P,6] = (point, view direction) -> [P,4] = (rgb raw,
    sigma raw)."""
    x, d = inp[:, :3].float(), inp[:, 3:].float()
    x = (x + BOUND) / (2 * BOUND)
    h = mlp(hash_encode(x, p["encoder.params"]), p["sigma_net.params"], SIGMA_SHAPES)
    sigma, geo = h[:, 0], h[:, 1:1 + GEO_FEAT]
    d = (d + 1) / 2                              # the reference maps to [0,1]; the SH encoding maps back to [-1,1]
    sh = sh4(d * 2 - 1)
    cin = torch.cat([sh, geo, torch.ones_like(sigma)[:, None]], -1)
    color = mlp(cin, p["color_net.params"], COLOR_SHAPES)[:, :3]
    return torch.cat([color, sigma[:, None]], -1)
