"""Drop-in for the six hot-path functions of the reference ``run.py``: ``batchify``, ``run_network``,
``batchify_rays``, ``render``, ``create_nerf``, ``render_rays`` (run.py:1624-1748, 2003-2128, 2235-2381) —
same names, positional/keyword signatures, returned structures and dictionary keys.  ``install(run_module)``
assigns them over a loaded ``run`` module so ``run.train()`` uses the B200 path with its source unchanged.
"""
import os

import numpy as np
import torch

from . import helpers, ops
from .optim import FusedAdam
from .tcnn import NeRF_TCNN
from .helpers import NeRF, SingleDeviceParallel, get_embedder, get_rays, ndc_rays, unwrap

DEBUG = False


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("gbnerf_b200 needs a CUDA device: the render path has no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


# ---- run.py:1624-1653 ------------------------------------------------------------------------------------ #
def batchify(fn, chunk):
    """run.py:1624-1634.  Kept for callers; the fused MLP kernel needs no point chunking."""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)

    return ret


def run_network(inputs2, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """run.py:1637-1653: points [R,S,3] (+ per-ray viewdirs [R,3]) -> raw [R,S,4].

    With the package's ``NeRF`` the embedding + concat + netchunk loop collapse into one fused kernel launch
    (``embed_fn``/``embeddirs_fn`` only certify the 10/4-band geometry).  A foreign ``fn`` gets the reference
    behaviour: embed, concatenate, apply in ``netchunk`` slices.
    """
    net = unwrap(fn)
    if isinstance(net, NeRF) and viewdirs is not None and getattr(embed_fn, "num_freqs", None) == 10 \
            and getattr(embeddirs_fn, "num_freqs", None) == 4 and inputs2.dim() == 3:
        return net.forward_points(inputs2, viewdirs)
    inputs_flat = torch.reshape(inputs2, [-1, inputs2.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs2.shape)
        embedded = torch.cat([embedded, embeddirs_fn(torch.reshape(input_dirs, [-1, input_dirs.shape[-1]]))], -1)
    outputs_flat = batchify(fn, netchunk)(embedded.contiguous())
    return torch.reshape(outputs_flat, list(inputs2.shape[:-1]) + [outputs_flat.shape[-1]])


def _identity(inp):
    """The embedders of create_nerf_tcnn (run.py:2134, 2139): the model encodes its own inputs."""
    return inp


class NetworkQuery:
    """``network_query_fn`` of create_nerf (run.py:2059-2062) as an object: callable like the reference's
    closure, plus ``fused`` which lets render_rays skip materialising the [R,S,3] point tensor."""

    def __init__(self, embed_fn, embeddirs_fn, netchunk):
        self.embed_fn, self.embeddirs_fn, self.netchunk = embed_fn, embeddirs_fn, netchunk

    def __call__(self, inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, embed_fn=self.embed_fn, embeddirs_fn=self.embeddirs_fn,
                           netchunk=self.netchunk)

    def fused(self, rays_o, rays_d, viewdirs, z_vals, network_fn):
        net = unwrap(network_fn)
        if isinstance(net, NeRF) and viewdirs is not None and getattr(self.embed_fn, "num_freqs", None) == 10 \
                and getattr(self.embeddirs_fn, "num_freqs", None) == 4:
            return net.forward_rays(rays_o, rays_d, viewdirs, z_vals)
        if isinstance(net, NeRF_TCNN) and viewdirs is not None and self.embed_fn is _identity and self.embeddirs_fn is _identity:
            return net.forward_rays(rays_o, rays_d, viewdirs, z_vals)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
        return self(pts, viewdirs, network_fn)


# ---- run.py:1656-1669 ------------------------------------------------------------------------------------ #
def batchify_rays(rays_flat, chunk=1024 * 32, need_alpha=False, detach_weights=False, **kwargs):
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], need_alpha=need_alpha, detach_weights=detach_weights, **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in all_ret.items()}


def _pack_rays_fused(H, W, focal, rays, c2w, ndc, near, far, use_viewdirs, c2w_staticcam, depths, patch):
    """One-launch ray setup (ops.pack_rays) when every input is a plain fp32 CUDA tensor without gradient and
    near/far are numbers; returns (ray batch, shape of rays_d) or None."""
    import numbers
    if not (isinstance(near, numbers.Real) and isinstance(far, numbers.Real)):
        return None
    ok = lambda t: isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and not t.requires_grad
    if depths is not None and not ok(depths):
        return None
    if c2w is not None:
        if not ok(c2w) or c2w.dim() != 2 or (c2w_staticcam is not None and (not ok(c2w_staticcam) or not use_viewdirs)):
            return None
        batch = ops.pack_rays(H, W, focal, near, far, c2w=c2w, c2w_staticcam=c2w_staticcam, depths=depths, patch=patch,
                              use_viewdirs=use_viewdirs, ndc=ndc)
        if patch is None:
            return batch, (H, W, 3)
        i, j, len1, len2 = [int(v) for v in patch]
        return batch, (max(0, min(len1, H - i)), max(0, min(len2, W - j)), 3)
    rays_o, rays_d = rays
    if not (ok(rays_o) and ok(rays_d)) or rays_o.shape != rays_d.shape or rays_d.shape[-1] != 3:
        return None
    batch = ops.pack_rays(H, W, focal, near, far, rays_o=rays_o, rays_d=rays_d, depths=depths, use_viewdirs=use_viewdirs,
                          ndc=ndc)
    return batch, tuple(rays_d.shape)


# ---- run.py:1672-1748 ------------------------------------------------------------------------------------ #
def render(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, depths=None, need_alpha=False, detach_weights=False, patch=None, **kwargs):
    """Same contract as run.py:1672: returns [rgb_map, disp_map, acc_map, depth_map, extras]."""
    packed = _pack_rays_fused(H, W, focal, rays, c2w, ndc, near, far, use_viewdirs, c2w_staticcam, depths, patch)
    if packed is not None:
        rays, sh = packed
    else:   # tensors the kernel does not take (other dtypes, gradients into the rays, tensor-valued near/far)
        if c2w is not None:
            rays_o, rays_d = get_rays(H, W, focal, c2w)
            if patch is not None:
                i, j, len1, len2 = patch
                rays_o = rays_o[i:i + len1, j:j + len2, :]
                rays_d = rays_d[i:i + len1, j:j + len2, :]
        else:
            rays_o, rays_d = rays
        if use_viewdirs:
            viewdirs = rays_d
            if c2w_staticcam is not None:
                rays_o, rays_d = get_rays(H, W, focal, c2w_staticcam)
            viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
            viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
        sh = rays_d.shape
        if ndc:
            rays_o, rays_d = ndc_rays(H, W, focal, 1., rays_o, rays_d)
        rays_o = torch.reshape(rays_o, [-1, 3]).float()
        rays_d = torch.reshape(rays_d, [-1, 3]).float()
        near, far = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
        cols = [rays_o, rays_d, near, far]
        if depths is not None:
            cols.append(depths.reshape(-1, 1))
        if use_viewdirs:
            cols.append(viewdirs)
        rays = torch.cat(cols, -1)
    all_ret = batchify_rays(rays, chunk, need_alpha=need_alpha, detach_weights=detach_weights, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    k_extract = ['rgb_map', 'disp_map', 'acc_map', 'depth_map']
    return [all_ret[k] for k in k_extract] + [{k: all_ret[k] for k in all_ret if k not in k_extract}]


# ---- run.py:2003-2128 ------------------------------------------------------------------------------------ #
def create_nerf(args):
    """Instantiate the coarse/fine MLPs, the query function, Adam and the two kwargs dicts (run.py:2003-2128)."""
    device = _device()
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed)
    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    output_ch = 5 if args.N_importance > 0 else 4
    skips = [4]
    if getattr(args, "alpha_model_path", None) is not None:
        raise NotImplementedError("alpha_model_path / NeRF_RGB is outside the B200 hot path (SURVEY.md §8)")
    precision = getattr(args, "precision", None)
    # the kernels serve the network the reference ships (aconfig_1 + --no_tcnn): say so here, at construction, instead of
    # at the first forward call
    if not (args.netdepth == 8 and args.netwidth == 256 and input_ch == 63 and input_ch_views == 27 and args.use_viewdirs
            and (args.N_importance <= 0 or (args.netdepth_fine == 8 and args.netwidth_fine == 256))):
        raise NotImplementedError(
            "gbnerf_b200.create_nerf: the B200 kernels implement the reference's shipped network only - netdepth 8, "
            "netwidth 256, multires 10, multires_views 4, use_viewdirs, i_embed 0 (got netdepth "
            f"{args.netdepth}/{getattr(args, 'netdepth_fine', None)}, netwidth {args.netwidth}/{getattr(args, 'netwidth_fine', None)}, "
            f"input channels {input_ch}+{input_ch_views}, use_viewdirs {args.use_viewdirs})")
    model = NeRF(D=args.netdepth, W=args.netwidth, input_ch=input_ch, output_ch=output_ch, skips=skips,
                 input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs, precision=precision).to(device)
    model = SingleDeviceParallel(model)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = NeRF(D=args.netdepth_fine, W=args.netwidth_fine, input_ch=input_ch, output_ch=output_ch,
                          skips=skips, input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs,
                          precision=precision).to(device)
        grad_vars += list(model_fine.parameters())
        model_fine = SingleDeviceParallel(model_fine)
    network_query_fn = NetworkQuery(embed_fn, embeddirs_fn, args.netchunk)
    # a torch.optim.Adam whose step() is one fused kernel per network (optim.py); same state_dict layout
    optimizer = FusedAdam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))

    start = 0
    basedir, expname = args.basedir, args.expname
    if args.ft_path is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    else:
        ckpts = [os.path.join(basedir, expname, f) for f in sorted(os.listdir(os.path.join(basedir, expname)))
                 if 'tar' in f]
    print('Found ckpts', ckpts)
    if len(ckpts) > 0 and not args.no_reload:
        ckpt_path = ckpts[-1]
        print('Reloading from', ckpt_path)
        ckpt = torch.load(ckpt_path, map_location=device)
        start = ckpt['global_step']
        optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        model.load_state_dict(ckpt['network_fn_state_dict'])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt['network_fine_state_dict'])

    render_kwargs_train = {
        'network_query_fn': network_query_fn, 'perturb': args.perturb, 'N_importance': args.N_importance,
        'network_fine': model_fine, 'N_samples': args.N_samples, 'network_fn': model,
        'use_viewdirs': args.use_viewdirs, 'white_bkgd': args.white_bkgd, 'raw_noise_std': args.raw_noise_std,
    }
    if args.dataset_type != 'llff' or args.no_ndc:
        print('Not ndc!')
        render_kwargs_train['ndc'] = False
        render_kwargs_train['lindisp'] = args.lindisp
    else:
        render_kwargs_train['ndc'] = True
    render_kwargs_test = {k: render_kwargs_train[k] for k in render_kwargs_train}
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    if getattr(args, "sigma_loss", False):
        from .loss import SigmaLoss
        render_kwargs_train['sigma_loss'] = SigmaLoss(args.N_samples, args.perturb, args.raw_noise_std)
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer


# ---- run.py:2131-2232 ------------------------------------------------------------------------------------ #
def create_nerf_tcnn(args):
    """run.py:2131-2232: the hash-grid models (coarse + fine ``NeRF_TCNN``), identity embedders, Adam, kwargs dicts.
    As in the reference, a checkpoint only restores ``global_step`` (its model/optimizer loads are commented out)."""
    device = _device()
    embed_fn = _identity
    embeddirs_fn = _identity if args.use_viewdirs else None
    grad_vars = []
    model = model_fine = None
    if args.alpha_model_path is None:
        model = NeRF_TCNN(encoding="hashgrid").to(device)
        grad_vars = list(model.parameters())
        model = SingleDeviceParallel(model)
    if args.N_importance > 0:
        if args.alpha_model_path is None:
            model_fine = NeRF_TCNN(encoding="hashgrid").to(device)
            grad_vars += list(model_fine.parameters())
            model_fine = SingleDeviceParallel(model_fine)
    network_query_fn = NetworkQuery(embed_fn, embeddirs_fn, args.netchunk)
    optimizer = FusedAdam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))   # flat vectors: takes the stock Adam path
    start = 0
    basedir, expname = args.basedir, args.expname
    if args.ft_path is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    else:
        ckpts = [os.path.join(basedir, expname, f) for f in sorted(os.listdir(os.path.join(basedir, expname))) if 'tar' in f]
    print('Found ckpts', ckpts)
    if len(ckpts) > 0 and not args.no_reload:
        print('Reloading from', ckpts[-1])
        start = torch.load(ckpts[-1], map_location=device)['global_step']
    render_kwargs_train = {
        'network_query_fn': network_query_fn, 'perturb': args.perturb, 'N_importance': args.N_importance,
        'network_fine': model_fine, 'N_samples': args.N_samples, 'network_fn': model, 'use_viewdirs': args.use_viewdirs,
        'white_bkgd': args.white_bkgd, 'raw_noise_std': args.raw_noise_std,
    }
    if args.dataset_type != 'llff' or args.no_ndc:
        print('Not ndc!')
        render_kwargs_train['ndc'] = False
        render_kwargs_train['lindisp'] = args.lindisp
    else:
        render_kwargs_train['ndc'] = True
    render_kwargs_test = {k: render_kwargs_train[k] for k in render_kwargs_train}
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer


# ---- run.py:2235-2381 ------------------------------------------------------------------------------------ #
def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., pytest=False,
                sigma_loss=None, verbose=False, need_alpha=False, detach_weights=False, _randoms=None):
    """Volumetric rendering of one ray chunk; returns the dict of run.py:2359-2370.

    Kernel sequence (all on the current stream): stratified depths -> fused encode+MLP (coarse) -> composite
    -> fused sample_pdf+merge -> fused encode+MLP (fine) -> composite.

    ``_randoms`` (not in the reference): dict with any of ``t_rand`` [R,S], ``noise0`` [R,S], ``u`` [R,N],
    ``noise1`` [R,S+N] to inject the random tensors (already scaled) instead of drawing them — the reference
    draws them in exactly this order (run.py:2307, helpers:377, helpers:318, helpers:377).
    """
    if network_fn is None:
        raise NotImplementedError("network_fn=None (NeRF_RGB / alpha_model branch) is outside the B200 hot path")
    rnd = _randoms or {}
    ray_batch = ray_batch if ray_batch.is_contiguous() else ray_batch.contiguous()
    N_rays = ray_batch.shape[0]
    dev = ray_batch.device
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 9 else None
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]

    def draw(key, shape, normal=False, scale=1.0):
        if key in rnd and rnd[key] is not None:
            return rnd[key]
        if pytest:
            np.random.seed(0)
            return torch.tensor(np.random.rand(*shape) * scale, dtype=torch.float32, device=dev)
        return (torch.randn(shape, device=dev) * scale) if normal else torch.rand(shape, device=dev)

    def query(z, net):
        if hasattr(network_query_fn, "fused"):
            return network_query_fn.fused(rays_o, rays_d, viewdirs, z, net)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
        return network_query_fn(pts, viewdirs, net)

    t_rand = draw("t_rand", (N_rays, N_samples)) if perturb > 0. else None
    z_vals = ops.zvals_stratified(near, far, N_samples, lindisp, t_rand)

    raw = query(z_vals, network_fn)
    noise = draw("noise0", (N_rays, N_samples), normal=True, scale=raw_noise_std) if raw_noise_std > 0. else None
    rgb_map, disp_map, acc_map, weights, depth_map, alpha = ops.composite(
        raw, z_vals, rays_d, noise, white_bkgd, detach_weights, need_alpha)

    if N_importance > 0:
        rgb_map_0, disp_map_0, acc_map_0, alpha0 = rgb_map, disp_map, acc_map, alpha
        u = draw("u", (N_rays, N_importance)) if perturb != 0. else None
        z_vals, z_std, _ = ops.sample_pdf_merge(z_vals, weights, N_importance, u)
        run_fn = network_fn if network_fine is None else network_fine
        raw = query(z_vals, run_fn)
        noise = draw("noise1", (N_rays, N_samples + N_importance), normal=True, scale=raw_noise_std) \
            if raw_noise_std > 0. else None
        rgb_map, disp_map, acc_map, weights, depth_map, alpha = ops.composite(
            raw, z_vals, rays_d, noise, white_bkgd, detach_weights, need_alpha)

    ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map, 'depth_map': depth_map,
           'weights': weights, 'z_vals': z_vals}
    if retraw:
        ret['raw'] = raw
    if need_alpha:
        ret['alpha'] = alpha
        ret['alpha0'] = alpha0  # NameError when N_importance == 0, as in the reference (run.py:2365)
    if N_importance > 0:
        ret['rgb0'], ret['disp0'], ret['acc0'] = rgb_map_0, disp_map_0, acc_map_0
        ret['z_std'] = z_std
    if sigma_loss is not None and ray_batch.shape[-1] > 11:
        depths = ray_batch[:, 8]
        ret['sigma_loss'] = sigma_loss.calculate_loss(rays_o, rays_d, viewdirs, near.reshape(-1, 1),
                                                      far.reshape(-1, 1), depths, network_query_fn, network_fine)
    if DEBUG:
        for k in ret:
            if torch.isnan(ret[k]).any() or torch.isinf(ret[k]).any():
                print(f"! [Numerical Error] {k} contains nan or inf.")
    return ret


# ---- run.py:2443-2474 ------------------------------------------------------------------------------------ #
def depth2xyz_torch(depth_map, depth_cam_matrix, depth_scale=1.0):
    """run.py:2443-2456: depth [h,w] + intrinsics [3,3] -> xyz [h,w,3], on the depth map's device."""
    dev, dt = depth_map.device, depth_map.dtype
    cam = torch.as_tensor(depth_cam_matrix, dtype=dt, device=dev)
    fx, fy, cx, cy = cam[0, 0], cam[1, 1], cam[0, 2], cam[1, 2]
    h = torch.arange(depth_map.shape[0], device=dev, dtype=dt)[:, None].expand(depth_map.shape)
    w = torch.arange(depth_map.shape[1], device=dev, dtype=dt)[None, :].expand(depth_map.shape)
    z = depth_map / depth_scale
    x = (w - cx) * z / fx
    y = (h - cy) * z / fy
    return torch.cat([x.unsqueeze(-1), y.unsqueeze(-1), z.unsqueeze(-1)], axis=-1)


def depth2normal_geo(depth, k=31):
    """run.py:2458-2474: xyz points [b,3,h,w] -> plane-fit normals [b,3,h,w] over k x k windows; one kernel instead of
    an 11.5 KB/pixel unfold, a batched inverse and two batched matmuls (csrc/normals.cu), differentiable."""
    return ops.normals_from_points(depth, k)


HOT_PATH = ("batchify", "run_network", "batchify_rays", "render", "create_nerf", "create_nerf_tcnn", "render_rays")


def install(run_module):
    """Assign the B200 implementations over a loaded reference ``run`` module (and its star-imported helpers)."""
    for name in HOT_PATH:
        setattr(run_module, name, globals()[name])
    for name in ("get_embedder", "NeRF", "get_rays", "ndc_rays", "sample_pdf", "raw2outputs"):
        setattr(run_module, name, getattr(helpers, name))
    run_module.NeRF_TCNN = NeRF_TCNN
    run_module.depth2xyz_torch, run_module.depth2normal_geo = depth2xyz_torch, depth2normal_geo
    return run_module
