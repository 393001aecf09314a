"""Loader for the UNMODIFIED reference hot path (container-only test infrastructure).

This is TEST INFRASTRUCTURE.  It imports the reference's own code from
``/root/reference`` so that (a) ``tests/golden/make_golden.py`` can dump golden
vectors and (b) ``tests/test_oracle_vs_reference.py`` can pin the oracle
restatement (``oracle/nerf_oracle.py``) against the real thing.  Nothing in the
product package, ``bench.py`` or the ``-m gpu`` tests may import this module:
``/root/reference`` does not exist on the GPU box.

How it works:
* ``DS_NeRF/run_nerf_helpers.py`` imports cleanly once ``matplotlib`` is stubbed
  (its only missing, unrelated import).
* ``run.py`` cannot be imported (module-scope ``torch.set_default_device('cuda')``,
  ``tkinter``, ``lpips`` …), so the six hot-path ``FunctionDef`` nodes
  (``batchify, run_network, batchify_rays, render, render_rays, create_nerf``,
  run.py:1624-1748, 2003-2128, 2235-2381) are compiled individually from its
  AST into a namespace that holds the helpers' symbols and the few module
  globals those functions read (``device``, ``DEBUG``, ``device_ids``).
"""
from __future__ import annotations

import ast
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GBNERF_REFERENCE_ROOT", "/root/reference")

_WANTED = ("batchify", "run_network", "batchify_rays", "render", "render_rays", "create_nerf",
           "depth2xyz_torch", "depth2normal_geo")   # run.py:2443-2474


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "run.py"))


def load(device="cpu"):
    """Return a namespace dict with the reference helpers + the six run.py functions."""
    import numpy as np
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    if not available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    helpers = importlib.import_module("DS_NeRF.run_nerf_helpers")
    loss_mod = importlib.import_module("DS_NeRF.loss")

    ns = {k: getattr(helpers, k) for k in dir(helpers) if not k.startswith("__")}
    dev = torch.device(device)
    ns.update(torch=torch, np=np, nn=nn, F=F, os=os, SigmaLoss=loss_mod.SigmaLoss,
              device=dev, DEBUG=False,
              device_ids=[dev.index or 0] if dev.type == "cuda" else [])
    with open(os.path.join(REFERENCE_ROOT, "run.py")) as fh:
        tree = ast.parse(fh.read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in _WANTED:
            exec(compile(ast.Module(body=[node], type_ignores=[]), "run.py", "exec"), ns)
    missing = [w for w in _WANTED if w not in ns]
    if missing:
        raise RuntimeError(f"reference run.py lacks {missing}")
    ns["helpers"] = helpers
    return ns


STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def staged_available() -> bool:
    return os.path.isfile(os.path.join(STAGED_ROOT, "run_hotpath.py"))


def load_staged(device="cpu"):
    """The same namespace as ``load()``, from the copy ``oracle/make_ref.py`` staged under ``oracle/_ref/`` (the
    unmodified reference files; present on the GPU box, where ``/root/reference`` is not).  Used by ``bench.py``'s
    ``--impl reference`` / ``cpu_baseline`` legs (kind "reference")."""
    import numpy as np
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    if not staged_available():
        raise FileNotFoundError(f"no staged reference under {STAGED_ROOT} (python oracle/make_ref.py)")
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if STAGED_ROOT not in sys.path:
        sys.path.insert(0, STAGED_ROOT)
    helpers = importlib.import_module("DS_NeRF.run_nerf_helpers")
    loss_mod = importlib.import_module("DS_NeRF.loss")
    ns = {k: getattr(helpers, k) for k in dir(helpers) if not k.startswith("__")}
    dev = torch.device(device)
    ns.update(torch=torch, np=np, nn=nn, F=F, os=os, SigmaLoss=loss_mod.SigmaLoss, device=dev, DEBUG=False,
              device_ids=[dev.index or 0] if dev.type == "cuda" else [])
    with open(os.path.join(STAGED_ROOT, "run_hotpath.py")) as fh:
        exec(compile(fh.read(), "run_hotpath.py", "exec"), ns)
    ns["helpers"] = helpers
    return ns


def default_args(tmpdir, **over):
    """aconfig_1 hot-path values + --no_tcnn."""
    a = types.SimpleNamespace(
        multires=10, multires_views=4, i_embed=0, use_viewdirs=True,
        N_samples=64, N_importance=64, netdepth=8, netdepth_fine=8,
        netwidth=256, netwidth_fine=256, alpha_model_path=None, no_coarse=False,
        netchunk=65536, lrate=3e-3, basedir=str(tmpdir), expname="exp",
        ft_path=None, no_reload=True, perturb=1.0, white_bkgd=True,
        raw_noise_std=1.0, dataset_type="llff", no_ndc=True, lindisp=True,
        sigma_loss=False)
    for k, v in over.items():
        setattr(a, k, v)
    os.makedirs(os.path.join(a.basedir, a.expname), exist_ok=True)
    return a
