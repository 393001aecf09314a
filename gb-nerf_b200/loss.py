"""Drop-in for ``DS_NeRF/loss.py``: ``SigmaLoss`` (loss.py:8-44) on the B200 path.

The extra ray march from ``near`` to the ray's known depth is the same fused kernel pair as a render pass
(stratified depths -> fused encode + MLP); only the final soft-max-like term is left to torch (R x S elementwise).
As in the reference (run.py:2372-2375) the result is only stored in the render dict; nothing adds it to the loss.
"""
import torch
import torch.nn.functional as F

from . import ops
from .helpers import unwrap, NeRF


class SigmaLoss:
    def __init__(self, N_samples, perturb, raw_noise_std):
        self.N_samples = N_samples
        self.perturb = perturb
        self.raw_noise_std = raw_noise_std

    def calculate_loss(self, rays_o, rays_d, viewdirs, near, far, depths, run_func, network, _randoms=None):
        """loss.py:15-44.  ``_randoms`` (not in the reference): dict with ``t_rand`` [R,S] / ``noise`` [R,S] to inject
        the random tensors for parity tests."""
        rnd = _randoms or {}
        N_rays = rays_o.shape[0]
        dev = rays_o.device
        near = near.reshape(N_rays, 1)
        t_rand = None
        if self.perturb > 0.:
            t_rand = rnd.get("t_rand")
            if t_rand is None:
                t_rand = torch.rand(N_rays, self.N_samples, device=dev)
        # z = near*(1-t) + depth*t with the stratified jitter of run.py:2301-2315: the depth kernel with far := depth
        z_vals = ops.zvals_stratified(near.contiguous(), depths.reshape(N_rays, 1).contiguous(), self.N_samples, False, t_rand)
        net = unwrap(network)
        if isinstance(net, NeRF) and hasattr(run_func, "fused"):
            raw = run_func.fused(rays_o, rays_d, viewdirs, z_vals, network)
        else:
            pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
            raw = run_func(pts, viewdirs, network)
        noise = 0.
        if self.raw_noise_std > 0.:
            noise = rnd.get("noise")
            if noise is None:
                noise = torch.randn(raw[..., 3].shape, device=dev) * self.raw_noise_std
        sigma = F.relu(raw[..., 3] + noise)
        return -torch.exp(sigma[:, -1]) / (torch.sum(torch.exp(sigma), dim=1) + 1)
