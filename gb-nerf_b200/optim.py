"""Optimizer step of the training loop: ``optimizer.step()`` at run.py:1529 on the
``torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))`` that create_nerf builds (run.py:2065).

``FusedAdam`` *is* a ``torch.optim.Adam`` (same constructor, ``param_groups``, ``state_dict`` layout: ``step``,
``exp_avg``, ``exp_avg_sq`` per parameter, so the ``.tar`` checkpoints of run.py:1552-1559 / 2085-2093 load either
way and the learning-rate decay of run.py:1540-1544, which assigns ``param_group['lr']``, keeps working).  For every
``NeRF`` module whose 24 tensors it owns, ``step()`` is ONE launch of ``gbn_adam_step_repack``: Adam's update in
place, and the new values written straight into the module's bf16 forward / transposed weight images, so neither
the seven ``multi_tensor_apply`` launches of the stock optimizer nor a re-pack pass runs.  Parameters that belong
to no such module take the stock path.
"""
import ctypes as C

import torch

from . import _lib, helpers


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **kw):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("the reference uses plain Adam (run.py:2065): no weight decay, no amsgrad")
        super().__init__(params, lr=lr, betas=betas, eps=eps, **kw)
        self._plans = None
        self._lazy_steps = 0      # steps taken by train.TrainStep's graph (device-side counter) not yet in state['step']

    def _flush_lazy(self):
        """TrainStep advances Adam on the device; fold its step count into the per-tensor ``step`` entries."""
        n, self._lazy_steps = self._lazy_steps, 0
        if n:
            for st in self.state.values():
                if "step" in st:
                    st["step"] += n

    def state_dict(self):
        self._flush_lazy()
        return super().state_dict()

    # -- which NeRF modules are stepped natively ------------------------------------------------------------
    def _build_plans(self):
        plans = []
        for gi, group in enumerate(self.param_groups):
            if group.get("maximize") or group.get("capturable") or group.get("differentiable"):
                continue
            ids = {id(p) for p in group["params"]}
            for mod in list(helpers._NERF_REGISTRY):
                try:
                    ps = mod.param_list()
                except Exception:          # a geometry the kernels do not serve
                    continue
                if all(id(p) in ids for p in ps) and all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in ps):
                    plans.append((gi, mod, ps))
        self._plans = (sum(len(g["params"]) for g in self.param_groups), plans)

    def _state_of(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _native_step(self, group, mod, ps):
        states = [self._state_of(p) for p in ps]
        for st in states:
            st["step"] += 1
        step = int(states[0]["step"].item())
        if any(int(st["step"].item()) != step for st in states[1:]):
            raise RuntimeError("FusedAdam: the tensors of one network are at different step counts")
        fused_pack = mod.precision == "bf16" and _lib.load().gbn_mlp_variant() == 1
        if fused_pack:                      # images exist and are current before they are patched in place
            fwd, bwd = mod.packed_weights(), mod.packed_weights_bwd()
        arr = lambda ts: (C.c_void_p * 24)(*[t.data_ptr() for t in ts])
        grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
        beta1, beta2 = group["betas"]
        _lib.call("gbn_adam_step_repack", arr(ps), arr(grads), arr([s["exp_avg"] for s in states]),
                  arr([s["exp_avg_sq"] for s in states]), float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                  step, fwd.data_ptr() if fused_pack else None, bwd.data_ptr() if fused_pack else None,
                  torch.cuda.current_stream(ps[0].device).cuda_stream)
        if not fused_pack:                  # the kernel wrote through raw pointers: tell the cache the weights moved
            mod._packed_key = mod._packed_bwd_key = None

    @torch.no_grad()
    def step(self, closure=None):
        self._flush_lazy()
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        n = sum(len(g["params"]) for g in self.param_groups)
        if self._plans is None or self._plans[0] != n:
            self._build_plans()
        hidden = []
        for gi, mod, ps in self._plans[1]:
            if any(p.grad is None or p.grad.is_sparse or p.grad.dtype != torch.float32 for p in ps):
                continue                    # torch skips tensors without a gradient; let it
            self._native_step(self.param_groups[gi], mod, ps)
            hidden += [(p, p.grad) for p in ps]
        if len(hidden) < n:                 # whatever is left (foreign parameters) takes the stock path
            for p, _ in hidden:
                p.grad = None
            try:
                super().step()
            finally:
                for p, g in hidden:
                    p.grad = g
        return loss
