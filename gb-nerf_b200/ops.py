"""Tensor-level operators over the C ABI (include/gbnerf.h) + their autograd wiring.

PyTorch is plumbing here: it owns device memory and the current stream; every operator below enqueues exactly
one (or two) kernels of libgbnerf.so on that stream.  Inputs must be CUDA fp32 tensors — anything else raises
(ValueError for shape/dtype/device problems, in the spirit of the reference extension's AT_ASSERTM checks,
torchsearchsorted/src/cuda/searchsorted_cuda_wrapper.cpp:5-7).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import PRECISION, PACK_BWD_BF16


# bench.py sets this to a list to collect (name, start_event, end_event, points) per MLP launch, recorded on
# the launching stream; None (the default) records nothing
KERNEL_EVENTS = None


class _timed_launch:
    def __init__(self, name, units):
        self.name, self.units = name, units

    def __enter__(self):
        if KERNEL_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream())

    def __exit__(self, *exc):
        if KERNEL_EVENTS is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream())
            KERNEL_EVENTS.append((self.name, self.e0, e1, self.units))
        return False


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_WD = (C.c_uint32 * 4)()


def _raise_if_watchdog_fired():
    """Fail loudly, and without a device synchronisation, if an earlier MLP launch of this process ran into its barrier
    watchdog: the kernels write that fact to host memory (gbn_watchdog_report), and results after it are meaningless."""
    if _lib.load().gbn_watchdog_report(_WD, 4) >= 4 and _WD[3]:
        raise _lib.GbnError(f"an MLP kernel hit its barrier watchdog (wait code {hex(_WD[0])}, CTA {_WD[1]}, thread "
                            f"{_WD[2]}); outputs since then are invalid - _lib.watchdog_report() has the full record")


def _chk(t, name, dim=None, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise ValueError(f"{name} is required")
    if not isinstance(t, torch.Tensor):
        raise ValueError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device (no CPU path exists)")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")
    if dim is not None and t.dim() != dim:
        raise ValueError(f"{name} must have {dim} dims, got shape {tuple(t.shape)}")
    return t


def _dense(t, name, dim=None, allow_none=False):
    t = _chk(t, name, dim, allow_none)
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _ray_views(*cols):
    """[R,k] column views of one packed ray batch (run.py:1726-1736) share a row pitch and are passed as they
    are; anything else (different pitches, broadcast strides, a single ray) is copied to dense [R,k]."""
    cols = [_chk(c, "ray tensor", 2) for c in cols]
    R, k = cols[0].shape
    if any(tuple(c.shape) != (R, k) for c in cols):
        raise ValueError("ray tensors disagree in shape")
    pitch = cols[0].stride(0)
    if R > 1 and pitch >= k and all(c.stride(1) == 1 and c.stride(0) == pitch for c in cols):
        return cols, int(pitch)
    return [c.contiguous() for c in cols], int(k)


# --------------------------------------------------------------------------------------------------------- #
def pack_rays(H, W, focal, near, far, c2w=None, rays_o=None, rays_d=None, c2w_staticcam=None, depths=None, patch=None,
              use_viewdirs=False, ndc=False):
    """The ray batch render() hands to batchify_rays (run.py:1700-1736) in one launch: ``[R, 8 (+1) (+3)]`` =
    o, d, near, far, (depth), (unit viewdir).  Either a pose ``c2w`` [3,>=4] (optionally a ``patch`` =
    (i, j, len1, len2) of the frame and a ``c2w_staticcam``) or ``rays_o``/``rays_d`` [...,3]."""
    if c2w is not None:
        c2w = _chk(c2w, "c2w", 2)
        if c2w.shape[0] < 3 or c2w.shape[1] < 4 or c2w.stride(1) != 1:
            raise ValueError(f"c2w must be a dense [3,>=4] pose, got {tuple(c2w.shape)}")
        if c2w_staticcam is not None:
            c2w_staticcam = _chk(c2w_staticcam, "c2w_staticcam", 2)
            if c2w_staticcam.shape[0] < 3 or c2w_staticcam.shape[1] < 4 or c2w_staticcam.stride(1) != 1:
                raise ValueError("c2w_staticcam must be a dense [3,>=4] pose")
        pi, pj, ph, pw = (0, 0, H, W) if patch is None else [int(v) for v in patch]
        ph, pw = max(0, min(ph, H - pi)), max(0, min(pw, W - pj))    # python slicing clamps (run.py:1703-1704)
        R, o, d, dev = ph * pw, None, None, c2w.device
    else:
        o = _chk(rays_o, "rays_o").reshape(-1, 3)
        d = _chk(rays_d, "rays_d").reshape(-1, 3)
        o = o if o.stride(1) == 1 and o.stride(0) >= 3 else o.contiguous()
        d = d if d.stride(1) == 1 and d.stride(0) >= 3 else d.contiguous()
        if o.shape != d.shape:
            raise ValueError("rays_o and rays_d disagree in shape")
        pi = pj = ph = pw = 0
        R, dev = o.shape[0], o.device
    if depths is not None:
        depths = _chk(depths, "depths").reshape(-1).contiguous()
        if depths.numel() != R:
            raise ValueError("depths must hold one value per ray")
    out = torch.empty(R, 8 + (depths is not None) + (3 if use_viewdirs else 0), device=dev, dtype=torch.float32)
    _lib.call("gbn_pack_rays", _ptr(c2w), c2w.stride(0) if c2w is not None else 0, _ptr(c2w_staticcam),
              c2w_staticcam.stride(0) if c2w_staticcam is not None else 0, _ptr(o), o.stride(0) if o is not None else 0,
              _ptr(d), d.stride(0) if d is not None else 0, _ptr(depths), int(H), int(W), float(focal), pi, pj, ph, pw,
              int(bool(use_viewdirs)), int(bool(ndc)), float(near), float(far), R, _ptr(out), _stream())
    return out


def zvals_stratified(near, far, n_samples, lindisp=False, t_rand=None):
    """near, far: [R] or [R,1] (column views of the ray batch are fine) -> z [R,S].  run.py:2291-2315."""
    near = near.reshape(near.shape[0], -1)[:, :1] if near.dim() != 2 else near[:, :1]
    far = far.reshape(far.shape[0], -1)[:, :1] if far.dim() != 2 else far[:, :1]
    (near, far), pitch = _ray_views(near, far)
    R = near.shape[0]
    t_rand = _dense(t_rand, "t_rand", 2, allow_none=True)
    if t_rand is not None and tuple(t_rand.shape) != (R, n_samples):
        raise ValueError("t_rand must be [R, N_samples]")
    z = torch.empty(R, n_samples, device=near.device, dtype=torch.float32)
    _lib.call("gbn_zvals_stratified", _ptr(near), _ptr(far), pitch, R, int(n_samples), int(bool(lindisp)),
              _ptr(t_rand), _ptr(z), _stream())
    return z


def encode_points(rays_o, rays_d, viewdirs, z):
    """Materialised [R*S, 90] embedding of o + d*z and the ray's view direction (test / measurement only)."""
    (rays_o, rays_d, viewdirs), pitch = _ray_views(rays_o, rays_d, viewdirs)
    z = _dense(z, "z", 2)
    R, S = z.shape
    out = torch.empty(R * S, 90, device=z.device, dtype=torch.float32)
    _lib.call("gbn_encode_points", _ptr(rays_o), _ptr(rays_d), _ptr(viewdirs), pitch, _ptr(z), R, S, _ptr(out),
              _stream())
    return out


# --------------------------------------------------------------------------------------------------------- #
class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, z, rays_d, noise, white_bkgd, detach_weights, need_alpha):
        raw = _dense(raw, "raw", 3)
        z = _dense(z, "z_vals", 2)
        noise = _dense(noise, "noise", 2, allow_none=True)
        (rays_d,), pitch = _ray_views(rays_d)
        R, S = z.shape
        if tuple(raw.shape) != (R, S, 4) or rays_d.shape[0] != R:
            raise ValueError(f"raw {tuple(raw.shape)} / z_vals {tuple(z.shape)} / rays_d {tuple(rays_d.shape)} mismatch")
        dev = raw.device
        rgb = torch.empty(R, 3, device=dev); disp = torch.empty(R, device=dev); acc = torch.empty(R, device=dev)
        depth = torch.empty(R, device=dev); weights = torch.empty(R, S, device=dev)
        alpha = torch.empty(R, S, device=dev) if need_alpha else None
        # algorithmic bytes: 24 S + 36 per ray, + 4 S each for noise in / alpha out
        with _timed_launch(f"composite_fwd_S{S}", R * (24 * S + 36 + 4 * S * ((noise is not None) + bool(need_alpha)))):
            _lib.call("gbn_composite_forward", _ptr(raw), _ptr(z), _ptr(rays_d), pitch, _ptr(noise), R, S,
                      int(bool(white_bkgd)), _ptr(rgb), _ptr(disp), _ptr(acc), _ptr(depth), _ptr(weights), _ptr(alpha),
                      _stream())
        ctx.save_for_backward(raw, z, rays_d, noise if noise is not None else torch.empty(0, device=dev))
        ctx.cfg = (pitch, bool(white_bkgd), bool(detach_weights), noise is not None)
        if alpha is None:
            alpha = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(alpha)
        return rgb, disp, acc, weights, depth, alpha

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_weights, g_depth, _g_alpha):
        raw, z, rays_d, noise = ctx.saved_tensors
        pitch, white, detach_w, has_noise = ctx.cfg
        R, S = z.shape
        g_raw = torch.empty_like(raw)
        c = lambda g: g.contiguous().float() if g is not None else None
        g_rgb, g_disp, g_acc, g_weights, g_depth = map(c, (g_rgb, g_disp, g_acc, g_weights, g_depth))
        # algorithmic bytes: read raw, z (20 S) [+ noise 4 S] [+ g_weights 4 S] + grads, write g_raw 16 S.  (SURVEY §8d
        # counts 40 S + 60 with the weights re-read; this kernel recomputes them instead, so 36 S + 60 is what it must move)
        with _timed_launch(f"composite_bwd_S{S}", R * (36 * S + 60 + 4 * S * (has_noise + (g_weights is not None)))):
            _lib.call("gbn_composite_backward", _ptr(raw), _ptr(z), _ptr(rays_d), pitch,
                      _ptr(noise if has_noise else None), R, S, int(white), int(detach_w), _ptr(g_rgb), _ptr(g_disp),
                      _ptr(g_acc), _ptr(g_depth), _ptr(g_weights), _ptr(g_raw), _stream())
        return g_raw, None, None, None, None, None, None


def composite_backward_raw(raw, z, rays_d, noise, white_bkgd, detach_weights, g_rgb, g_disp, g_acc, g_depth, g_weights=None):
    """Kernel launch only: gradient wrt raw [R,S,4] given the output gradients (no autograd)."""
    (rays_d,), pitch = _ray_views(rays_d)
    R, S = z.shape
    g_raw = torch.empty_like(raw)
    _lib.call("gbn_composite_backward", _ptr(raw), _ptr(z), _ptr(rays_d), pitch, _ptr(noise), R, S, int(bool(white_bkgd)),
              int(bool(detach_weights)), _ptr(g_rgb), _ptr(g_disp), _ptr(g_acc), _ptr(g_depth), _ptr(g_weights), _ptr(g_raw),
              _stream())
    return g_raw


def composite(raw, z_vals, rays_d, noise=None, white_bkgd=False, detach_weights=False, need_alpha=False):
    """raw2outputs on the device: returns (rgb, disp, acc, weights, depth, alpha-or-None)."""
    rgb, disp, acc, weights, depth, alpha = _Composite.apply(raw, z_vals, rays_d, noise, white_bkgd, detach_weights,
                                                             need_alpha)
    return rgb, disp, acc, weights, depth, (alpha if need_alpha else None)


# --------------------------------------------------------------------------------------------------------- #
def searchsorted_right(cdf, u):
    """int64 indices == torch.searchsorted(cdf, u, right=True)."""
    cdf, u = _dense(cdf, "cdf", 2), _dense(u, "u", 2)
    if cdf.shape[0] != u.shape[0]:
        raise ValueError("cdf and u disagree on the number of rows")
    inds = torch.empty(u.shape, device=u.device, dtype=torch.int64)
    _lib.call("gbn_searchsorted_right", _ptr(cdf), _ptr(u), u.shape[0], cdf.shape[1], u.shape[1], _ptr(inds), _stream())
    return inds


def sample_pdf(bins, weights, n_samples, u=None):
    """bins [R,B], weights [R,B-1] -> samples [R,N]; u=None is the deterministic linspace branch."""
    bins, weights = _dense(bins.detach(), "bins", 2), _dense(weights.detach(), "weights", 2)
    u = _dense(u, "u", 2, allow_none=True)
    R, B = bins.shape
    if tuple(weights.shape) != (R, B - 1):
        raise ValueError(f"weights must be [R, B-1] = {(R, B - 1)}, got {tuple(weights.shape)}")
    if u is not None and tuple(u.shape) != (R, n_samples):
        raise ValueError("u must be [R, N_samples]")
    out = torch.empty(R, n_samples, device=bins.device, dtype=torch.float32)
    _lib.call("gbn_sample_pdf", _ptr(bins), _ptr(weights), _ptr(u), R, B, int(n_samples), _ptr(out), _stream())
    return out


def sample_pdf_merge(z_vals, weights, n_importance, u=None, want_samples=False, cdf=None, want_inds=False):
    """Fused run.py:2343-2348 + :2370.  Returns (z_merged [R,S+N], z_std [R], z_samples [R,N] or None); with
    ``want_inds`` a fourth item, the int32 bin indices ``searchsorted(cdf, u, right=True)`` of helpers:333.
    ``cdf`` [R,S-1] replaces the cdf built from ``weights`` (test hook: same cdf + same u -> bit-exact indices)."""
    z_vals = _dense(z_vals.detach(), "z_vals", 2)
    weights = _dense(weights.detach() if weights is not None else None, "weights", 2, allow_none=cdf is not None)
    u = _dense(u, "u", 2, allow_none=True)
    cdf = _dense(cdf, "cdf", 2, allow_none=True)
    R, S = z_vals.shape
    if weights is not None and tuple(weights.shape) != (R, S):
        raise ValueError("weights must match z_vals")
    if cdf is not None and tuple(cdf.shape) != (R, S - 1):
        raise ValueError("cdf must be [R, S-1]")
    N = int(n_importance)
    if u is not None and tuple(u.shape) != (R, N):
        raise ValueError("u must be [R, N_importance]")
    dev = z_vals.device
    merged = torch.empty(R, S + N, device=dev); std = torch.empty(R, device=dev)
    samples = torch.empty(R, N, device=dev) if want_samples else None
    inds = torch.empty(R, N, device=dev, dtype=torch.int32) if want_inds else None
    # algorithmic bytes (fused with the merge): read z 4 S + weights 4 S (+ u 4 N), write z 4 (S + N) + z_std 4
    with _timed_launch(f"sample_merge_S{S}_N{N}{'_det' if u is None else ''}",
                       R * (8 * S + 4 * (S + N) + 4 + (4 * N if u is not None else 0) + (4 * N if want_samples else 0))):
        _lib.call("gbn_sample_pdf_merge_ex", _ptr(z_vals), _ptr(weights), _ptr(u), _ptr(cdf), R, S, N, _ptr(samples),
                  _ptr(merged), _ptr(std), _ptr(inds), _stream())
    if want_inds:
        return merged, std, samples, inds
    return merged, std, samples


# --------------------------------------------------------------------------------------------------------- #
PARAM_ORDER = tuple([f"pts_linears.{i}" for i in range(8)] + ["feature_linear", "alpha_linear", "views_linears.0",
                                                               "rgb_linear"])
PARAM_SHAPES = {"pts_linears.0": (256, 63), "pts_linears.5": (256, 319), "feature_linear": (256, 256),
                "alpha_linear": (1, 256), "views_linears.0": (128, 283), "rgb_linear": (3, 128)}


def prepack_weights(params, precision="bf16", out=None):
    """params: 24 tensors (weight, bias) x PARAM_ORDER -> packed uint8 buffer in the kernel's smem layout."""
    prec = PACK_BWD_BF16 if precision == "bf16_bwd" else PRECISION[precision]
    if len(params) != 24:
        raise ValueError("expected 24 parameter tensors")
    keep = []
    for i, p in enumerate(params):
        p = _chk(p.detach(), f"params[{i}]")
        keep.append(p if p.is_contiguous() else p.contiguous())
    for i, name in enumerate(PARAM_ORDER):
        want = PARAM_SHAPES.get(name, (256, 256))
        if tuple(keep[2 * i].shape) != want or tuple(keep[2 * i + 1].shape) != (want[0],):
            raise ValueError(f"{name}: expected weight {want}, got {tuple(keep[2 * i].shape)}")
    nbytes = _lib.load().gbn_mlp_packed_bytes(prec)
    if out is None:
        out = torch.empty(nbytes, device=keep[0].device, dtype=torch.uint8)
    arr = (C.c_void_p * 24)(*[p.data_ptr() for p in keep])
    _lib.call("gbn_mlp_prepack_weights", arr, _ptr(out), prec, _stream())
    return out


def _workspace(R, device):
    return torch.empty(_lib.load().gbn_mlp_workspace_bytes(int(R)), device=device, dtype=torch.uint8)


def _stash(P, device):
    """Training stash (H or G): gbn_mlp_stash_bytes(P) bytes; torch allocations are 512-byte aligned."""
    return torch.empty(_lib.load().gbn_mlp_stash_bytes(int(P)), device=device, dtype=torch.uint8)


def mlp_forward_raw(packed, precision, viewdirs, R, S, rays_o=None, rays_d=None, z=None, pts=None, stash=None):
    """Kernel launch only (no autograd): raw [R,S,4]."""
    if pts is not None:
        pts = _dense(pts, "pts").reshape(-1, 3)
        (viewdirs,), pitch = _ray_views(viewdirs)
        rays_o = rays_d = None
    else:
        (rays_o, rays_d, viewdirs), pitch = _ray_views(rays_o, rays_d, viewdirs)
        z = _dense(z, "z", 2)
    raw = torch.empty(R, S, 4, device=viewdirs.device, dtype=torch.float32)
    ws = _workspace(R, viewdirs.device)
    _raise_if_watchdog_fired()
    with _timed_launch("mlp", R * S):
        _lib.call("gbn_mlp_forward", _ptr(packed), PRECISION[precision], _ptr(rays_o), _ptr(rays_d), _ptr(viewdirs),
                  pitch, _ptr(z), _ptr(pts), R, S, _ptr(raw), _ptr(ws), _ptr(stash), _stream())
    return raw, ws


def mlp_forward_embedded_raw(packed, precision, emb, stash=None):
    emb = _dense(emb, "embedded", 2)
    if emb.shape[1] != 90:
        raise ValueError(f"embedded input must be [P, 90], got {tuple(emb.shape)}")
    P = emb.shape[0]
    raw = torch.empty(P, 4, device=emb.device, dtype=torch.float32)
    ws = _workspace(P, emb.device)
    _lib.call("gbn_mlp_forward_embedded", _ptr(packed), PRECISION[precision], _ptr(emb), P, _ptr(raw), _ptr(ws),
              _ptr(stash), _stream())
    return raw, ws


def mlp_backward_raw(packed_bwd, g_raw, stash_h, viewdirs, R, S, param_shapes):
    """dgrad + wgrad launches: returns (list of 24 fp32 gradients in nn.Linear layout, workspace, G stash)."""
    g_raw = _dense(g_raw.reshape(R * S, 4), "g_raw", 2)
    (viewdirs,), pitch = _ray_views(viewdirs)
    dev = g_raw.device
    stash_g = _stash(R * S, dev)
    ws = torch.empty(512, device=dev, dtype=torch.uint8)
    _raise_if_watchdog_fired()
    with _timed_launch("mlp_dgrad", R * S):
        _lib.call("gbn_mlp_backward_data", _ptr(packed_bwd), _ptr(g_raw), R * S, _ptr(stash_h), _ptr(stash_g), _ptr(ws),
                  _stream())
    flat = torch.zeros(sum(int(torch.Size(s).numel()) for s in param_shapes), device=dev, dtype=torch.float32)
    grads, off = [], 0
    for shp in param_shapes:
        n = int(torch.Size(shp).numel())
        grads.append(flat[off:off + n].view(shp))
        off += n
    arr = (C.c_void_p * 24)(*[g.data_ptr() for g in grads])
    with _timed_launch("mlp_wgrad", R * S):
        _lib.call("gbn_mlp_backward_weights", _ptr(stash_h), _ptr(stash_g), _ptr(g_raw), _ptr(viewdirs), pitch, R, S, arr,
                  C.c_void_p(ws.data_ptr() + 256), _stream())
    return grads, ws, stash_g


def mlp_error_code(ws):
    """Watchdog word of the last launch that used this workspace (0 = clean); synchronises."""
    return int(ws[:4].view(torch.int32).item())


def torch_posenc(x, n_freqs):
    parts = [x]
    for k in range(n_freqs):
        parts += [torch.sin(x * float(2 ** k)), torch.cos(x * float(2 ** k))]
    return torch.cat(parts, -1)


class _Mlp(torch.autograd.Function):
    """raw = NeRF(embed(o + d z), embed(viewdir)).

    Forward: the fused tcgen05 kernel; when parameter gradients are needed it also writes the activation stash.
    Backward: the tcgen05 dgrad kernel (same skeleton, transposed weights) then the tcgen05 wgrad kernel - for all three
    input forms (rays + depths, explicit points, pre-embedded rows).  There is no other backward: tf32 modules are
    inference-only and raise when a gradient is asked of them (the reference trains in fp32 through autograd of
    run_nerf_helpers.py:106-129; here training is the bf16 path).  Inputs carry no gradient, exactly as in the
    reference's loop where z_samples is detached and rays are data (run.py:2346); asking for one raises instead of
    silently returning zeros.
    """

    @staticmethod
    def forward(ctx, module, mode, grad_mode, a0, a1, a2, a3, *params):
        packed = module.packed_weights()
        # needs_input_grad mirrors requires_grad of the inputs even under torch.no_grad(), and grad mode is always off
        # inside forward(): the caller's grad mode comes in as an argument, or every inference call would write the
        # training stash
        need_grad = grad_mode and any(ctx.needs_input_grad[7:])
        if grad_mode and any(ctx.needs_input_grad[3:7]):
            raise NotImplementedError(
                "gbnerf_b200: gradients with respect to rays / points / depths / embedded inputs are not implemented "
                "(the reference's training loop needs none, run.py:2346); detach the inputs")
        # parameter gradients exist for bf16 modules on the default kernels only; a tf32 module is inference-only: its
        # forward runs as usual (callers often render without torch.no_grad()), asking it for a gradient raises
        ctx.unsupported = need_grad and (module.precision != "bf16" or _lib.load().gbn_mlp_variant() != 1)
        ctx.native = need_grad and not ctx.unsupported
        stash = None
        if mode == "rays":
            rays_o, rays_d, viewdirs, z = a0, a1, a2, a3
            R, S = z.shape
            if ctx.native:
                stash = _stash(R * S, z.device)
            raw, ws = mlp_forward_raw(packed, module.precision, viewdirs, R, S, rays_o=rays_o, rays_d=rays_d, z=z,
                                      stash=stash)
        elif mode == "pts":
            pts, viewdirs = a0, a1
            R, S = pts.shape[0], pts.shape[1]
            if ctx.native:
                stash = _stash(R * S, pts.device)
            raw, ws = mlp_forward_raw(packed, module.precision, viewdirs, R, S, pts=pts.contiguous(), stash=stash)
        else:
            R, S = a0.shape[0], 1
            if ctx.native:
                stash = _stash(R, a0.device)
            raw, ws = mlp_forward_embedded_raw(packed, module.precision, a0, stash=stash)
            viewdirs = a0[:, :3]            # placeholder: the wgrad of the default kernels takes directions from the stash
        module.last_workspace = ws
        ctx.mode, ctx.module = mode, module
        if ctx.native:
            ctx.stash, ctx.RS = stash, (R, S)
            ctx.save_for_backward(viewdirs)
            ctx.shapes = [tuple(p.shape) for p in params]
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        if ctx.unsupported:
            raise NotImplementedError(
                "gbnerf_b200: parameter gradients exist for bf16 modules on the default kernels only - a tf32 module is "
                "inference-only (construct the NeRF with precision='bf16' to train)")
        if not ctx.native:
            return (None,) * 7 + (None,) * 24
        if ctx.stash is None:
            raise RuntimeError("gbnerf_b200: backward through the same MLP call twice is not supported (the activation "
                               "stash is released after the first pass; re-run the forward)")
        (viewdirs,) = ctx.saved_tensors
        R, S = ctx.RS
        grads, ws, _ = mlp_backward_raw(ctx.module.packed_weights_bwd(), g_raw.contiguous(), ctx.stash, viewdirs, R, S,
                                        ctx.shapes)
        ctx.module.last_workspace_bwd = ws
        ctx.stash = None
        return (None, None, None, None, None, None, None, *grads)


def mlp_rays(module, rays_o, rays_d, viewdirs, z):
    return _Mlp.apply(module, "rays", torch.is_grad_enabled(), rays_o, rays_d, viewdirs, z, *module.param_list())


def mlp_points(module, pts, viewdirs):
    return _Mlp.apply(module, "pts", torch.is_grad_enabled(), pts, viewdirs, None, None, *module.param_list())


def mlp_embedded(module, emb):
    return _Mlp.apply(module, "emb", torch.is_grad_enabled(), emb, None, None, None, *module.param_list())


# --------------------------------------------------------------------------------------------------------- #
def loss_seed(rgb, rgb0, disp, target_rgb, target_disp, depth_lambda=0.1, global_rays=None):
    """Fused img2mse terms: returns (loss[1], g_rgb, g_rgb0, g_disp) with means over the GLOBAL ray count."""
    rgb, target_rgb = _dense(rgb, "rgb", 2), _dense(target_rgb, "target_rgb", 2)
    rgb0 = _dense(rgb0, "rgb0", 2, allow_none=True)
    disp = _dense(disp, "disp", 1, allow_none=True)
    target_disp = _dense(target_disp, "target_disp", 1, allow_none=True)
    R = rgb.shape[0]
    dev = rgb.device
    g_rgb = torch.empty_like(rgb)
    g_rgb0 = torch.empty_like(rgb0) if rgb0 is not None else None
    g_disp = torch.empty_like(disp) if disp is not None else None
    loss = torch.zeros(1, device=dev)
    _lib.call("gbn_loss_seed", _ptr(rgb), _ptr(rgb0), _ptr(disp), _ptr(target_rgb), _ptr(target_disp), R,
              int(global_rays or R), float(depth_lambda), _ptr(g_rgb), _ptr(g_rgb0), _ptr(g_disp), _ptr(loss), _stream())
    return loss, g_rgb, g_rgb0, g_disp


# ---- depth -> normal map (csrc/normals.cu) -------------------------------------------------------------- #
class _Normals(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, k):
        points = _chk(points, "points", 4).contiguous()
        B, C, H, W = points.shape
        if C != 3:
            raise ValueError(f"points must be [B,3,H,W], got {tuple(points.shape)}")
        if k < 1 or k > 31 or k % 2 == 0:
            raise ValueError(f"window k must be odd and <= 31, got {k}")
        normals = torch.empty_like(points)
        minv = torch.empty(B, 6, H, W, device=points.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        _lib.call("gbn_normals_forward", _ptr(points), B, H, W, int(k), _ptr(normals), _ptr(minv), _stream())
        if minv is not None:
            ctx.k = int(k)
            ctx.save_for_backward(points, normals, minv)
        return normals

    @staticmethod
    def backward(ctx, g):
        points, normals, minv = ctx.saved_tensors
        B, _, H, W = points.shape
        g = _chk(g.contiguous(), "g_normals", 4)
        out = torch.empty_like(points)
        _lib.call("gbn_normals_backward", _ptr(points), _ptr(normals), _ptr(minv), _ptr(g), B, H, W, ctx.k, _ptr(out), _stream())
        return out, None


def normals_from_points(points, k=31):
    """depth2normal_geo (run.py:2458-2474): points [B,3,H,W] -> [B,3,H,W], differentiable wrt the points."""
    return _Normals.apply(points, k)


# ---- NeRF_TCNN (csrc/tcnn_model.cu) ----------------------------------------------------------------------- #
def tcnn_prepack(grid_params, sigma_params, color_params, out=None):
    lib = _lib.load()
    g, sg, co = (_chk(t.detach(), n, 1).contiguous() for t, n in
                 ((grid_params, "encoder.params"), (sigma_params, "sigma_net.params"), (color_params, "color_net.params")))
    if g.numel() != lib.gbn_tcnn_grid_params() or sg.numel() != 64 * 32 + 16 * 64 or co.numel() != 64 * 32 + 64 * 64 + 16 * 64:
        raise ValueError(f"NeRF_TCNN parameter sizes {g.numel()}, {sg.numel()}, {co.numel()} do not match the reference's model")
    if out is None:
        out = torch.empty(lib.gbn_tcnn_table_bytes(), device=g.device, dtype=torch.uint8)
    _lib.call("gbn_tcnn_prepack", _ptr(g), _ptr(sg), _ptr(co), _ptr(out), _stream())
    return out


def tcnn_forward_raw(table, R, S, rays_o=None, rays_d=None, viewdirs=None, z=None, inputs=None, stash=None):
    dev = table.device
    raw = torch.empty(R, S, 4, device=dev, dtype=torch.float32)
    with _timed_launch("tcnn", R * S):
        if inputs is not None:
            _lib.call("gbn_tcnn_forward", _ptr(table), None, None, None, 0, None, _ptr(inputs), R, S, _ptr(raw), _ptr(stash),
                      _stream())
        else:
            (ro, rd, vd), pitch = _ray_views(rays_o, rays_d, viewdirs)
            _lib.call("gbn_tcnn_forward", _ptr(table), _ptr(ro), _ptr(rd), _ptr(vd), pitch, _ptr(z), None, R, S, _ptr(raw),
                      _ptr(stash), _stream())
    return raw


class _Tcnn(torch.autograd.Function):
    """Autograd node of NeRF_TCNN: inputs carry no gradient (run.py:2346), parameters do."""

    @staticmethod
    def forward(ctx, module, mode, grad_mode, a0, a1, a2, a3, *params):
        table = module.table()
        if mode == "rays":
            z = _chk(a3, "z_vals", 2).contiguous()
            R, S = z.shape
        else:
            a0 = _chk(a0, "input", 2).contiguous()
            if a0.shape[1] != 6:
                raise ValueError(f"NeRF_TCNN takes [P,6] rows (point, direction), got {tuple(a0.shape)}")
            R, S, z = a0.shape[0], 1, None
        need_grad = grad_mode and any(ctx.needs_input_grad[7:])
        if grad_mode and any(ctx.needs_input_grad[3:7]):
            raise NotImplementedError(
                "gbnerf_b200: gradients with respect to rays / depths / input rows of NeRF_TCNN are not implemented "
                "(the reference's training loop needs none, run.py:2346); detach the inputs")
        stash = torch.empty(R * S, 32, device=table.device, dtype=torch.float16) if need_grad else None
        if mode == "rays":
            raw = tcnn_forward_raw(table, R, S, rays_o=a0, rays_d=a1, viewdirs=a2, z=z, stash=stash)
        else:
            raw = tcnn_forward_raw(table, R, S, inputs=a0, stash=stash).reshape(R, 4)
        if need_grad:
            ctx.module, ctx.mode, ctx.shape, ctx.table = module, mode, (R, S), table
            ctx.save_for_backward(a0, a1, a2, z, stash)
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        module, (R, S), table = ctx.module, ctx.shape, ctx.table
        a0, a1, a2, z, stash = ctx.saved_tensors
        g_raw = _chk(g_raw.contiguous(), "g_raw").reshape(R * S, 4)
        enc, sig, col = module.param_list()
        dev = table.device
        g_grid, g_sig, g_col = torch.zeros_like(enc), torch.zeros_like(sig), torch.zeros_like(col)
        g_enc = torch.empty(R * S, 32, device=dev, dtype=torch.float32)
        ls = float(module.loss_scale)
        if ctx.mode == "rays":
            (ro, rd, vd), pitch = _ray_views(a0, a1, a2)
            _lib.call("gbn_tcnn_backward", _ptr(table), _ptr(ro), _ptr(rd), _ptr(vd), pitch, _ptr(z), None, R, S, _ptr(stash),
                      _ptr(g_raw), ls, _ptr(g_enc), _ptr(g_grid), _ptr(g_sig), _ptr(g_col), _stream())
        else:
            _lib.call("gbn_tcnn_backward", _ptr(table), None, None, None, 0, None, _ptr(a0), R, S, _ptr(stash), _ptr(g_raw),
                      ls, _ptr(g_enc), _ptr(g_grid), _ptr(g_sig), _ptr(g_col), _stream())
        return (None,) * 7 + (g_grid, g_sig, g_col)


def tcnn_rays(module, rays_o, rays_d, viewdirs, z_vals):
    return _Tcnn.apply(module, "rays", torch.is_grad_enabled(), rays_o, rays_d, viewdirs, z_vals, *module.param_list())


def tcnn_inputs(module, inputs):
    return _Tcnn.apply(module, "inputs", torch.is_grad_enabled(), inputs, None, None, None, *module.param_list())
