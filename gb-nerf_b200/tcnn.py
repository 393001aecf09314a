"""Drop-in for ``DS_NeRF/run_nerf_helpers_tcnn.py``: ``NeRF_TCNN`` (lines 13-117) without tiny-cuda-nn.

The reference builds the model from four tiny-cuda-nn modules (``encoder`` HashGrid, ``sigma_net`` FullyFusedMLP,
``encoder_dir`` SphericalHarmonics, ``color_net`` FullyFusedMLP).  Here the same module tree holds the same flat
fp32 ``params`` vectors (so ``state_dict`` keys/shapes are those of the reference: ``encoder.params``,
``sigma_net.params``, ``encoder_dir.params`` (empty), ``color_net.params``) and the whole forward is one kernel
(csrc/tcnn_model.cu).  tiny-cuda-nn is neither vendored by the reference nor installed in this image, so numerical
parity is against the published algorithm as restated in oracle/tcnn_oracle.py ("parity unpinned").
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops

SIGMA_SHAPES = ((64, 32), (16, 64))
COLOR_SHAPES = ((64, 32), (64, 64), (16, 64))


class _FlatParams(nn.Module):
    """Stands where a ``tcnn.Encoding`` / ``tcnn.Network`` stands in the reference: one flat fp32 ``params``."""

    def __init__(self, init):
        super().__init__()
        self.params = nn.Parameter(init)


def _xavier(shapes):
    mats = []
    for o, i in shapes:
        lim = math.sqrt(6.0 / (o + i))
        mats.append(((torch.rand(o, i) * 2 - 1) * lim).reshape(-1))
    return torch.cat(mats)


class NeRF_TCNN(nn.Module):
    def __init__(self, encoding="HashGrid", encoding_dir="SphericalHarmonics", num_layers=2, hidden_dim=64, geo_feat_dim=15,
                 num_layers_color=3, hidden_dim_color=64, bound=100, **kwargs):
        super().__init__(**kwargs)
        if (str(encoding).lower(), encoding_dir, num_layers, hidden_dim, geo_feat_dim, num_layers_color, hidden_dim_color,
                bound) != ("hashgrid", "SphericalHarmonics", 2, 64, 15, 3, 64, 100):
            raise NotImplementedError("only the configuration the reference instantiates (run.py:2142-2156) is built")
        self.bound, self.num_layers, self.hidden_dim, self.geo_feat_dim = bound, num_layers, hidden_dim, geo_feat_dim
        self.num_layers_color, self.hidden_dim_color = num_layers_color, hidden_dim_color
        self.per_level_scale = float(np.exp2(np.log2(2048 * bound / 16) / (16 - 1)))
        n_grid = 14069664                       # == gbn_tcnn_grid_params(); checked against the library on first use
        self.encoder = _FlatParams((torch.rand(n_grid) * 2 - 1) * 1e-4)   # tiny-cuda-nn: U(-1e-4, 1e-4)
        self.sigma_net = _FlatParams(_xavier(SIGMA_SHAPES))
        self.encoder_dir = _FlatParams(torch.zeros(0))
        self.in_dim_color = 16 + geo_feat_dim
        self.color_net = _FlatParams(_xavier(COLOR_SHAPES))
        self.loss_scale = 128.0                 # tiny-cuda-nn's default for fp16 networks (power of two: exact)
        self._table = None
        self._table_key = None

    def param_list(self):
        return [self.encoder.params, self.sigma_net.params, self.color_net.params]

    def table(self):
        """fp16 device table (hash grid + MLP weights in the kernel's layout), rebuilt when a parameter changed."""
        ps = self.param_list()
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._table_key:
            self._table = ops.tcnn_prepack(*ps, out=self._table)
            self._table_key = key
        return self._table

    def forward(self, input):
        """[P, 6] = (point, view direction) -> [P, 4] = (r, g, b, sigma) raw, as lines 90-117 (values carry the
        fp16 rounding of the reference's modules; the tensor itself is fp32 for the compositing kernels)."""
        return ops.tcnn_inputs(self, input)

    def forward_rays(self, rays_o, rays_d, viewdirs, z_vals):
        """raw [R,S,4] for the points o + d*z (run.py:2317 + run_network with identity embedders, run.py:2134-2139)."""
        return ops.tcnn_rays(self, rays_o, rays_d, viewdirs, z_vals)
