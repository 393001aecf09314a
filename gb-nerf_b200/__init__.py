"""gbnerf_b200 — B200 (sm_100a) implementation of GB-NeRF's DS_NeRF volumetric-rendering hot path.

Layout
  csrc/        hand-written CUDA kernels + the C ABI (include/gbnerf.h) -> libgbnerf.so
  _lib.py      ctypes binding of that ABI (the stub of INTEGRATION.md)
  ops.py       tensor-level operators + autograd wiring over the ABI
  helpers.py   drop-in for DS_NeRF/run_nerf_helpers.py (NeRF, get_embedder, sample_pdf, raw2outputs, get_rays …)
  run.py       drop-in for run.py's create_nerf / render / batchify_rays / render_rays / run_network / batchify
  loss.py      drop-in for DS_NeRF/loss.py (SigmaLoss)
  dist.py      ray sharding across one-process-per-GPU ranks, gradient all-reduce, image gather
  train.py     TrainStep: one training iteration (render + loss + backward + all-reduce + Adam) as one CUDA graph

The package directory is ``gb-nerf_b200`` (the project's name); import it as ``gbnerf_b200`` through the
alias module at the repository root.  There is no CPU path: every operator raises without a CUDA device.
"""
from . import _lib, ops, helpers, run, dist, loss, optim, train  # noqa: F401
from .train import TrainStep  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from . import tcnn  # noqa: F401
from .tcnn import NeRF_TCNN  # noqa: F401
from .loss import SigmaLoss  # noqa: F401
from .helpers import (NeRF, Embedder, get_embedder, get_rays, get_rays_np, ndc_rays, sample_pdf,  # noqa: F401
                      raw2outputs, img2mse, mse2psnr, to8b)
from .run import (batchify, run_network, batchify_rays, render, create_nerf, create_nerf_tcnn, render_rays, install,  # noqa: F401
                  NetworkQuery, depth2xyz_torch, depth2normal_geo)  # noqa: F401

__version__ = "0.1.0"
