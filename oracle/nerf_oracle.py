"""CPU oracle: a restatement of the reference DS_NeRF volumetric-rendering path.

TEST INFRASTRUCTURE ONLY.  Importable from ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs — never from the
product package (``gb-nerf_b200``), which has no CPU fallback.

Parity status: PINNED.  Every function here is checked bit-for-bit / to fp32
round-off against the unmodified reference imported from ``/root/reference``
(``tests/test_oracle_vs_reference.py``, container-only) and against the golden
vectors committed under ``tests/golden/`` (made by ``tests/golden/make_golden.py``
from the reference itself).

The reference is PyTorch code, so the restatement is written with torch CPU ops
(float32 by default, float64 on request for derivations).  Randomness is never
drawn here: the three random tensors the reference consumes (``t_rand``
run.py:2307, ``noise`` run_nerf_helpers.py:377, ``u`` run_nerf_helpers.py:318)
are explicit inputs so that the CUDA path can be fed the same numbers.

Reference line citations use ``helpers`` = DS_NeRF/run_nerf_helpers.py.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

# --------------------------------------------------------------------------- #
# positional encoding  (helpers:23-71)
# --------------------------------------------------------------------------- #


def posenc(x: torch.Tensor, n_freqs: int) -> torch.Tensor:
    """[.., 3] -> [.., 3 + 6 L]; channel order x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...

    helpers:40 builds the bands as ``2 ** linspace(0, L-1, L)`` (exact powers of
    two in fp32) and helpers:46 multiplies ``x * freq`` before the periodic fn.
    """
    parts = [x]
    for k in range(n_freqs):
        xs = x * float(2 ** k)
        parts.append(torch.sin(xs))
        parts.append(torch.cos(xs))
    return torch.cat(parts, dim=-1)


def posenc_dim(n_freqs: int) -> int:
    return 3 + 6 * n_freqs


# --------------------------------------------------------------------------- #
# the 8x256 MLP (helpers:75-129); parameters addressed by checkpoint key
# --------------------------------------------------------------------------- #

PARAM_SHAPES = OrderedDict([
    ("pts_linears.0", (256, 63)), ("pts_linears.1", (256, 256)), ("pts_linears.2", (256, 256)),
    ("pts_linears.3", (256, 256)), ("pts_linears.4", (256, 256)), ("pts_linears.5", (256, 319)),
    ("pts_linears.6", (256, 256)), ("pts_linears.7", (256, 256)),
    ("views_linears.0", (128, 283)), ("feature_linear", (256, 256)),
    ("alpha_linear", (1, 256)), ("rgb_linear", (3, 128)),
])
"""Registration order of ``NeRF.__init__`` (helpers:88-104) = RNG consumption order."""


def init_params(generator_seed=None, dtype=torch.float32):
    """Default ``nn.Linear`` initialisation in the reference's construction order.

    nn.Linear.reset_parameters: weight ~ kaiming_uniform(a=sqrt 5) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)),
    bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)); weight drawn before bias.
    Uses real nn.Linear objects so the global RNG stream is consumed exactly as
    ``NeRF(...)`` consumes it (helpers:88-104).
    """
    if generator_seed is not None:
        torch.manual_seed(generator_seed)
    out = OrderedDict()
    for name, (fo, fi) in PARAM_SHAPES.items():
        lin = torch.nn.Linear(fi, fo)
        out[name + ".weight"] = lin.weight.detach().to(dtype).clone()
        out[name + ".bias"] = lin.bias.detach().to(dtype).clone()
    return out


def strip_module_prefix(sd):
    return OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in sd.items())


def mlp_forward(p, embedded: torch.Tensor, return_hidden: bool = False):
    """[P, 90] -> [P, 4] = (r, g, b, sigma_raw).  helpers:106-129.

    Skip concatenation is ``[input_pts, h]`` after layer index 4 (helpers:111-112),
    the view branch input is ``[feature, input_views]`` (helpers:117).
    """
    lin = lambda name, v: torch.addmm(p[name + ".bias"], v, p[name + ".weight"].t())
    x_pts, x_dir = embedded[..., :63], embedded[..., 63:]
    h = x_pts
    hidden = []
    for i in range(8):
        h = torch.relu(lin(f"pts_linears.{i}", h))
        hidden.append(h)
        if i == 4:
            h = torch.cat([x_pts, h], dim=-1)
    sigma = lin("alpha_linear", h)
    feat = lin("feature_linear", h)
    hv = torch.relu(lin("views_linears.0", torch.cat([feat, x_dir], dim=-1)))
    rgb = lin("rgb_linear", hv)
    out = torch.cat([rgb, sigma], dim=-1)
    if return_hidden:
        return out, hidden, feat, hv
    return out


def run_network(p, pts: torch.Tensor, viewdirs: torch.Tensor, chunk: int = 65536):
    """pts [R,S,3], viewdirs [R,3] -> raw [R,S,4].  run.py:1637-1653."""
    R, S, _ = pts.shape
    e_p = posenc(pts.reshape(-1, 3), 10)
    e_d = posenc(viewdirs[:, None, :].expand(R, S, 3).reshape(-1, 3), 4)
    emb = torch.cat([e_p, e_d], dim=-1)
    outs = [mlp_forward(p, emb[i:i + chunk]) for i in range(0, emb.shape[0], chunk)]
    return torch.cat(outs, 0).reshape(R, S, 4)


# --------------------------------------------------------------------------- #
# ray set-up (helpers:251-262, helpers:285-302, run.py:1700-1736)
# --------------------------------------------------------------------------- #


def get_rays(H: int, W: int, focal: float, c2w: torch.Tensor):
    """Pinhole rays; pixel (row j, col i) -> dir ((i-W/2)/f, -(j-H/2)/f, -1) rotated by c2w[:3,:3]."""
    dt = c2w.dtype
    ii = torch.linspace(0, W - 1, W, dtype=dt)[None, :].expand(H, W)
    jj = torch.linspace(0, H - 1, H, dtype=dt)[:, None].expand(H, W)
    cam = torch.stack([(ii - W * .5) / focal, -(jj - H * .5) / focal, -torch.ones_like(ii)], -1)
    d = (cam[..., None, :] * c2w[:3, :3]).sum(-1)
    o = c2w[:3, -1].expand(d.shape)
    return o, d


def ndc_rays(H, W, focal, near, o, d):
    """helpers:285-302."""
    t = -(near + o[..., 2]) / d[..., 2]
    o = o + t[..., None] * d
    ax, ay = -1. / (W / (2. * focal)), -1. / (H / (2. * focal))
    o_n = torch.stack([ax * o[..., 0] / o[..., 2], ay * o[..., 1] / o[..., 2],
                       1. + 2. * near / o[..., 2]], -1)
    d_n = torch.stack([ax * (d[..., 0] / d[..., 2] - o[..., 0] / o[..., 2]),
                       ay * (d[..., 1] / d[..., 2] - o[..., 1] / o[..., 2]),
                       -2. * near / o[..., 2]], -1)
    return o_n, d_n


def pack_rays(o, d, near, far, use_viewdirs=True, depths=None):
    """run.py:1707-1736: [R, 8 (+1) (+3)] = o, d, near, far, (depth), (unit viewdir)."""
    o = o.reshape(-1, 3).float()
    d = d.reshape(-1, 3).float()
    cols = [o, d, near * torch.ones_like(d[:, :1]), far * torch.ones_like(d[:, :1])]
    if depths is not None:
        cols.append(depths.reshape(-1, 1))
    if use_viewdirs:
        cols.append(d / torch.norm(d, dim=-1, keepdim=True))
    return torch.cat(cols, -1)


# --------------------------------------------------------------------------- #
# stratified depths (run.py:2291-2315)
# --------------------------------------------------------------------------- #


def stratified_z(near, far, n_samples: int, lindisp: bool, t_rand=None):
    """near/far [R,1] -> z [R,S].  ``t_rand`` [R,S] in [0,1) enables the perturbation."""
    t = torch.linspace(0., 1., n_samples, dtype=near.dtype)[None, :]
    if lindisp:
        z = 1. / (1. / near * (1. - t) + 1. / far * t)
    else:
        z = near * (1. - t) + far * t
    z = z.expand(near.shape[0], n_samples)
    if t_rand is not None:
        mid = .5 * (z[:, 1:] + z[:, :-1])
        hi = torch.cat([mid, z[:, -1:]], -1)
        lo = torch.cat([z[:, :1], mid], -1)
        z = lo + (hi - lo) * t_rand
    return z


# --------------------------------------------------------------------------- #
# alpha compositing (helpers:352-406)
# --------------------------------------------------------------------------- #


def composite(raw, z, d, noise=None, white_bkgd=False, detach_weights=False):
    """raw [R,S,4], z [R,S], d [R,3] -> dict(rgb, disp, acc, weights, depth, alpha).

    delta_i = (z_{i+1}-z_i)|d|, last = 1e10|d| (helpers:369-372); sigma = relu(raw_3+noise);
    alpha = 1-exp(-sigma delta) (helpers:367,385); T_i = prod_{j<i}(1-alpha_j+1e-10)
    (helpers:387); disp = 1/max(1e-10, depth/acc) -> NaN when acc == 0 (helpers:394).
    """
    gap = z[:, 1:] - z[:, :-1]
    gap = torch.cat([gap, torch.full_like(gap[:, :1], 1e10)], -1)
    gap = gap * torch.norm(d[:, None, :], dim=-1)
    colour = torch.sigmoid(raw[..., :3])
    dens = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1. - torch.exp(-torch.relu(dens) * gap)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * trans
    w_rgb = w.detach() if detach_weights else w
    rgb = (w_rgb[..., None] * colour).sum(-2)
    depth = (w * z).sum(-1)
    acc = w.sum(-1)
    disp = 1. / torch.maximum(torch.full_like(depth, 1e-10), depth / acc)
    if white_bkgd:
        rgb = rgb + (1. - acc[:, None])
    return dict(rgb=rgb, disp=disp, acc=acc, weights=w, depth=depth, alpha=alpha)


# --------------------------------------------------------------------------- #
# inverse-CDF sampling (helpers:306-349) and the sort-merge (run.py:2348)
# --------------------------------------------------------------------------- #


def build_cdf(weights):
    """weights [R,M] -> cdf [R,M+1] with cdf[:,0]=0.  helpers:308-311."""
    w = weights + 1e-5
    pdf = w / w.sum(-1, keepdim=True)
    c = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(c[:, :1]), c], -1)


def upper_bound(cdf, u):
    """#{j : cdf[r,j] <= u[r,n]}  ==  searchsorted(cdf, u, right=True) (helpers:333).

    Written as a comparison count so it does not lean on the library routine it pins.
    """
    out = torch.empty(u.shape, dtype=torch.int64)
    step = max(1, (1 << 24) // max(1, cdf.shape[-1] * u.shape[-1]))
    for s in range(0, cdf.shape[0], step):
        out[s:s + step] = (cdf[s:s + step, None, :] <= u[s:s + step, :, None]).sum(-1)
    return out


def invert_cdf(bins, cdf, u):
    """bins [R,B], cdf [R,B], u [R,N] -> (samples [R,N], inds [R,N] int64).  helpers:331-347."""
    nb = cdf.shape[-1]
    inds = upper_bound(cdf, u)
    lo = (inds - 1).clamp_min(0)
    hi = inds.clamp_max(nb - 1)
    c_lo, c_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    b_lo, b_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    den = c_hi - c_lo
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    frac = (u - c_lo) / den
    return b_lo + frac * (b_hi - b_lo), inds


def sample_pdf(bins, weights, n_samples: int, u=None):
    """``u is None`` is the deterministic branch (helpers:314-316: linspace(0,1,N))."""
    cdf = build_cdf(weights)
    if u is None:
        u = torch.linspace(0., 1., n_samples, dtype=cdf.dtype).expand(cdf.shape[0], n_samples)
    u = u.contiguous()
    return invert_cdf(bins, cdf, u)[0]


def merge_sorted(z, z_new):
    """run.py:2348: ascending sort of the concatenation, values only."""
    return torch.sort(torch.cat([z, z_new], -1), -1)[0]


# --------------------------------------------------------------------------- #
# render_rays / render (run.py:2235-2381, 1656-1748)
# --------------------------------------------------------------------------- #


def render_rays(rays, p_coarse, p_fine, n_samples=64, n_importance=64, lindisp=False,
                white_bkgd=False, t_rand=None, noise0=None, u=None, noise1=None,
                retraw=False, need_alpha=False, detach_weights=False, netchunk=65536):
    """rays [R, 11] (o, d, near, far, viewdir) -> dict with the keys of run.py:2359-2370.

    Random inputs, all optional: ``t_rand`` [R,S] (perturb>0), ``noise0`` [R,S] /
    ``noise1`` [R,S+N] = randn*raw_noise_std (coarse / fine), ``u`` [R,N]
    (random importance samples; None -> deterministic).  The reference ties ``u``
    to ``perturb`` (det = perturb==0, run.py:2345).
    """
    o, d, vd = rays[:, 0:3], rays[:, 3:6], rays[:, -3:]
    near, far = rays[:, 6:7], rays[:, 7:8]
    z = stratified_z(near, far, n_samples, lindisp, t_rand)
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    raw = run_network(p_coarse, pts, vd, netchunk)
    c0 = composite(raw, z, d, noise0, white_bkgd, detach_weights)
    out = c0
    z_samples = None
    if n_importance > 0:
        z_mid = .5 * (z[:, 1:] + z[:, :-1])
        z_samples = sample_pdf(z_mid, c0["weights"][:, 1:-1], n_importance, u).detach()
        z = merge_sorted(z, z_samples)
        pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
        raw = run_network(p_fine if p_fine is not None else p_coarse, pts, vd, netchunk)
        out = composite(raw, z, d, noise1, white_bkgd, detach_weights)
    ret = dict(rgb_map=out["rgb"], disp_map=out["disp"], acc_map=out["acc"], depth_map=out["depth"],
               weights=out["weights"], z_vals=z)
    if retraw:
        ret["raw"] = raw
    if need_alpha:
        ret["alpha"] = out["alpha"]
        ret["alpha0"] = c0["alpha"]
    if n_importance > 0:
        ret.update(rgb0=c0["rgb"], disp0=c0["disp"], acc0=c0["acc"],
                   z_std=torch.std(z_samples, dim=-1, unbiased=False))
    return ret


def render(rays, chunk=32768, **kw):
    """batchify_rays (run.py:1656-1669): chunk over rays; per-chunk random inputs are sliced."""
    per_ray = ("t_rand", "noise0", "u", "noise1")
    outs = {}
    for i in range(0, rays.shape[0], chunk):
        kk = {k: (v[i:i + chunk] if (k in per_ray and v is not None) else v) for k, v in kw.items()}
        r = render_rays(rays[i:i + chunk], **kk)
        for k, v in r.items():
            outs.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in outs.items()}


def reference_loss(ret, target_rgb, target_disp, depth_lambda=0.1):
    """mse(rgb)+mse(rgb0)+depth_lambda*mse(disp) (run.py:1483,1502,1513-1515)."""
    mse = lambda a, b: torch.mean((a - b) ** 2)
    return mse(ret["rgb_map"], target_rgb) + mse(ret["rgb0"], target_rgb) + depth_lambda * mse(ret["disp_map"], target_disp)


# --------------------------------------------------------------------------- #
# depth -> normal map (run.py:2443-2474)
# --------------------------------------------------------------------------- #

def depth2xyz(depth_map, cam, depth_scale=1.0):
    """run.py:2443-2456: back-project a depth map [H,W] with intrinsics cam [3,3] -> xyz [H,W,3]."""
    fx, fy, cx, cy = cam[0, 0], cam[1, 1], cam[0, 2], cam[1, 2]
    hh, ww = torch.meshgrid(torch.arange(depth_map.shape[0], dtype=depth_map.dtype),
                            torch.arange(depth_map.shape[1], dtype=depth_map.dtype), indexing="ij")
    z = depth_map / depth_scale
    return torch.stack([(ww - cx) * z / fx, (hh - cy) * z / fy, z], -1)


def depth2normal_geo(points, k=31):
    """run.py:2458-2474: points [B,3,H,W] -> [B,3,H,W]; per pixel n = (A^T A)^-1 A^T 1 over its zero-padded k x k window
    of points A [k*k, 3] (least-squares plane n.p = 1)."""
    B, C, H, W = points.shape
    cols = torch.nn.functional.unfold(points, (k, k), dilation=1, padding=(k - 1) // 2, stride=1)   # [B, 3*k*k, H*W]
    A = cols.transpose(1, 2).reshape(B, H, W, C, k * k).transpose(-1, -2)                              # [B,H,W,k*k,3]
    At = A.transpose(-1, -2)
    n = torch.matmul(torch.matmul(torch.linalg.inv(torch.matmul(At, A)), At), torch.ones(B, H, W, k * k, 1, dtype=points.dtype))
    return n.squeeze(-1).permute(0, 3, 1, 2)


# --------------------------------------------------------------------------- #
# synthetic workloads of SURVEY §8d (shared by tests and bench.py)
# --------------------------------------------------------------------------- #

H_FULL, W_FULL, FOCAL = 756, 1008, 815.0
NEAR, FAR = 1.2, 8.0


def synthetic_c2w(dtype=torch.float32):
    c2w = torch.zeros(3, 4, dtype=dtype)
    c2w[:, :3] = torch.eye(3, dtype=dtype)
    c2w[:, 3] = torch.tensor([0.1, -0.05, 0.2], dtype=dtype)
    return c2w


def synthetic_rays(n=None, seed=1):
    """LLFF/SPIn-NeRF ``images_4``-shaped rays; ``n`` random pixels (seeded) or the full frame."""
    o, d = get_rays(H_FULL, W_FULL, FOCAL, synthetic_c2w())
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    if n is not None:
        g = torch.Generator().manual_seed(seed)
        idx = torch.randint(0, H_FULL * W_FULL, (n,), generator=g)
        o, d = o[idx], d[idx]
    return pack_rays(o, d, NEAR, FAR)


def mlp_flops_per_point() -> int:
    return 2 * (63 * 256 + 4 * 256 * 256 + 319 * 256 + 2 * 256 * 256 + 256 + 256 * 256 + 283 * 128 + 128 * 3)


assert mlp_flops_per_point() == 1186816 and math.isclose(mlp_flops_per_point() * 192 / 1e6, 227.87, abs_tol=0.01)
