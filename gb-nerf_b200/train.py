"""One training iteration of the reference loop as ONE CUDA graph (SURVEY.md §7 "CUDA-graph the step", §8e).

What the reference does per iteration (run.py:1429-1546, second stage): ``render(...)`` of an ``N_rand`` ray batch with
the train kwargs (``perturb=1``, ``raw_noise_std=1``), ``loss = img2mse(rgb, t) + img2mse(rgb0, t) + depth_lambda *
img2mse(disp, t_depth)`` (run.py:1483,1502,1513-1515), ``loss.backward()``, ``optimizer.step()``, and a learning-rate
decay written into ``param_group['lr']``.  Under ``nn.DataParallel`` every 65,536-point MLP call is scattered over the
GPUs (run.py:2020,2056).

``TrainStep`` runs the same arithmetic as the drop-in ``render`` + autograd path of this package — the same kernels,
launched directly instead of through ``torch.autograd`` — on this rank's contiguous block of the batch
(``dist.shard_bounds``), with the losses taken as means over the GLOBAL batch, the two networks' gradients summed over
ranks by NCCL (the fine network's all-reduce is issued as soon as its wgrad is done and overlaps the coarse network's
backward), and Adam + weight re-pack as one launch per network.  Everything from the random draws to the optimizer
step is captured once and replayed per step, so the ~150 launches / allocations / autograd nodes of the eager step
(0.55 ms of a 5 ms step at 4096 rays, most of a 0.7 ms step at 512 rays per GPU) collapse into one graph launch.

Kernel sequence per replay (all in libgbnerf.so unless noted):
  torch RNG x2 (one uniform draw for t_rand and u, one normal draw for the two noise tensors) -> zvals -> MLP fwd+stash (coarse) -> composite -> sample+merge ->
  MLP fwd+stash (fine) -> composite -> loss_seed -> composite bwd (fine) -> dgrad -> wgrad -> [NCCL all-reduce fine]
  -> composite bwd (coarse) -> dgrad -> wgrad -> [NCCL all-reduce coarse] -> adam_tick -> adam+repack x2.
"""
import ctypes as C

import torch
import torch.distributed as tdist

from . import _lib, ops
from .dist import shard_bounds, world
from .helpers import NeRF, unwrap

_p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None


class TrainStep:
    """``step(rays, target_rgb, target_disp)`` = render + loss + backward + all-reduce + Adam for one ray batch.

    ``render_kwargs``: the train dict of ``create_nerf`` (+ ``near``/``far`` as run.py:963-968 adds them); ``optimizer``:
    the ``FusedAdam`` it returned; ``n_rays``: GLOBAL batch size (``N_rand``); each rank passes the whole batch and
    takes its own block, or passes its block already cut (``sharded_input=True``).
    ``graph=False`` runs the same launches eagerly (used by the parity tests with injected random tensors).
    """

    def __init__(self, render_kwargs, optimizer, n_rays, near=None, far=None, depth_lambda=0.1, graph=True,
                 overlap_allreduce=True, device=None):
        kw = dict(render_kwargs)
        self.near = float(kw.pop("near", near) if near is None else near)
        self.far = float(kw.pop("far", far) if far is None else far)
        self.coarse, self.fine = unwrap(kw["network_fn"]), unwrap(kw.get("network_fine"))
        if not isinstance(self.coarse, NeRF) or not isinstance(self.fine, NeRF):
            raise NotImplementedError("TrainStep serves the coarse+fine 8x256 NeRF pair of create_nerf")
        if self.coarse.precision != "bf16" or self.fine.precision != "bf16":
            raise NotImplementedError("the native backward is bf16 (tf32 modules are inference-only)")
        if _lib.load().gbn_mlp_variant() != 1:
            raise NotImplementedError("TrainStep needs the default (TMEM-operand) bf16 kernels")
        self.S, self.N = int(kw["N_samples"]), int(kw["N_importance"])
        if self.N <= 0:
            raise NotImplementedError("TrainStep needs N_importance > 0 (coarse + fine pass)")
        self.lindisp = bool(kw.get("lindisp", False))
        self.perturb = float(kw.get("perturb", 0.))
        self.noise_std = float(kw.get("raw_noise_std", 0.))
        self.white = bool(kw.get("white_bkgd", False))
        if kw.get("ndc", False):
            raise NotImplementedError("TrainStep takes rays in world space (no_ndc, as aconfig_1 sets it)")
        self.opt = optimizer
        self.depth_lambda = float(depth_lambda)
        self.rank, self.world = world()
        self.R_global = int(n_rays)
        self.lo, self.hi = shard_bounds(self.R_global, self.rank, self.world)
        self.R = self.hi - self.lo
        self.dev = device or next(self.coarse.parameters()).device
        self.overlap = bool(overlap_allreduce) and self.world > 1
        self.use_graph = bool(graph)
        self.graph = None
        self.steps_done = 0
        self._alloc()

    # ------------------------------------------------------------------------------------------------------
    def _alloc(self):
        dev, R, S, N = self.dev, self.R, self.S, self.N
        f = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)
        self.rays = f(R, 11)                      # o(3) d(3) near far viewdir(3): the batch of run.py:1726-1736
        self.target_rgb, self.target_disp = f(R, 3), f(R)
        # random tensors: one buffer for the two uniform draws and one for the two normal draws (two RNG launches per step
        # instead of four; the draws are i.i.d., so how they are batched does not change their distribution)
        self._uniform, self._normal = f(R * (S + N)), f(R * (2 * S + N))
        self.t_rand, self.u = self._uniform[:R * S].view(R, S), self._uniform[R * S:].view(R, N)
        self.noise0, self.noise1 = self._normal[:R * S].view(R, S), self._normal[R * S:].view(R, S + N)
        self.z0, self.raw0, self.w0 = f(R, S), f(R, S, 4), f(R, S)
        self.z1, self.raw1, self.w1, self.zstd = f(R, S + N), f(R, S + N, 4), f(R, S + N), f(R)
        self.out0 = [f(R, 3), f(R), f(R), f(R)]   # rgb0, disp0, acc0, depth0
        self.out1 = [f(R, 3), f(R), f(R), f(R)]   # rgb, disp, acc, depth
        self.g_rgb, self.g_rgb0, self.g_disp = f(R, 3), f(R, 3), f(R)
        self.g_raw0, self.g_raw1 = f(R, S, 4), f(R, S + N, 4)
        self.loss = torch.zeros(1, device=dev)
        self.stash_h = [ops._stash(R * S, dev), ops._stash(R * (S + N), dev)]
        self.stash_g = [ops._stash(R * S, dev), ops._stash(R * (S + N), dev)]
        self.ws_f = [ops._workspace(R, dev), ops._workspace(R, dev)]
        self.ws_b = [torch.zeros(512, device=dev, dtype=torch.uint8) for _ in range(2)]
        # one flat gradient buffer per network, aliased by p.grad (so optimizer / checkpoint code sees gradients)
        self.nets = [self.coarse, self.fine]
        self.params = [n.param_list() for n in self.nets]
        sizes = [sum(p.numel() for p in ps) for ps in self.params]
        self.flat_all = torch.zeros(sum(sizes), device=dev)                # one buffer: one memset per step
        self.flat = [self.flat_all[:sizes[0]], self.flat_all[sizes[0]:]]
        self.grads = []
        for ps, flat in zip(self.params, self.flat):
            off, views = 0, []
            for p in ps:
                v = flat[off:off + p.numel()].view_as(p)
                p.grad = v
                views.append(v)
                off += p.numel()
            self.grads.append(views)
        # Adam state lives in the optimizer (same tensors its state_dict() saves); step count + lr on the device
        group = None
        for g in self.opt.param_groups:
            ids = {id(p) for p in g["params"]}
            if all(id(p) in ids for ps in self.params for p in ps):
                group = g
        if group is None:
            raise ValueError("the optimizer does not own the two networks' parameters in one param group")
        self.group = group
        self.state = [[self.opt._state_of(p) for p in ps] for ps in self.params]
        step0 = float(self.state[0][0]["step"])
        self.step_dev = torch.full((1,), step0, device=dev, dtype=torch.float64)
        self._expected_step = step0 + getattr(self.opt, "_lazy_steps", 0)
        self.launches_per_step = None
        self.lr_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.lr_dev = torch.zeros(1, device=dev)
        self.scalars = torch.zeros(2, device=dev)
        self.packed = [(n.packed_weights(), n.packed_weights_bwd()) for n in self.nets]
        arr = lambda ts: (C.c_void_p * 24)(*[t.data_ptr() for t in ts])
        self.ptrs = [dict(p=arr(ps), g=arr(gs), m=arr([s["exp_avg"] for s in st]), v=arr([s["exp_avg_sq"] for s in st]))
                     for ps, gs, st in zip(self.params, self.grads, self.state)]
        self.side = torch.cuda.Stream(device=dev) if self.overlap else None

    # ------------------------------------------------------------------------------------------------------
    def _launches(self, randoms=None):
        """Enqueue one whole step on the current stream (captured into the graph, or run eagerly)."""
        R, S, N, SF = self.R, self.S, self.N, self.S + self.N
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        call = _lib.call
        rays = self.rays
        o, d, near, far, vd = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8], rays[:, 8:11]
        pitch = 11
        rnd = randoms or {}
        # the four random tensors of run.py:2307, helpers:377 (twice) and helpers:318; `randoms` injects them (parity tests)
        t_rand = noise0 = u = noise1 = None
        if rnd:
            if self.perturb > 0.:
                t_rand, u = self.t_rand.copy_(rnd["t_rand"]), self.u.copy_(rnd["u"])
            if self.noise_std > 0.:
                noise0, noise1 = self.noise0.copy_(rnd["noise0"]), self.noise1.copy_(rnd["noise1"])
        else:
            if self.perturb > 0.:
                self._uniform.uniform_()
                t_rand, u = self.t_rand, self.u
            if self.noise_std > 0.:
                self._normal.normal_(0., self.noise_std)
                noise0, noise1 = self.noise0, self.noise1
        self.loss.zero_()
        self.flat_all.zero_()                                                # wgrad accumulates

        # ---- forward --------------------------------------------------------------------------------------
        call("gbn_zvals_stratified", _p(near), _p(far), pitch, R, S, int(self.lindisp), _p(t_rand), _p(self.z0), st)
        call("gbn_mlp_forward", _p(self.packed[0][0]), 0, _p(o), _p(d), _p(vd), pitch, _p(self.z0), None, R, S,
             _p(self.raw0), _p(self.ws_f[0]), _p(self.stash_h[0]), st)
        call("gbn_composite_forward", _p(self.raw0), _p(self.z0), _p(d), pitch, _p(noise0), R, S, int(self.white),
             _p(self.out0[0]), _p(self.out0[1]), _p(self.out0[2]), _p(self.out0[3]), _p(self.w0), None, st)
        call("gbn_sample_pdf_merge", _p(self.z0), _p(self.w0), _p(u), R, S, N, None, _p(self.z1), _p(self.zstd), st)
        call("gbn_mlp_forward", _p(self.packed[1][0]), 0, _p(o), _p(d), _p(vd), pitch, _p(self.z1), None, R, SF,
             _p(self.raw1), _p(self.ws_f[1]), _p(self.stash_h[1]), st)
        call("gbn_composite_forward", _p(self.raw1), _p(self.z1), _p(d), pitch, _p(noise1), R, SF, int(self.white),
             _p(self.out1[0]), _p(self.out1[1]), _p(self.out1[2]), _p(self.out1[3]), _p(self.w1), None, st)
        # ---- loss + its gradients (means over the global batch) -------------------------------------------------
        call("gbn_loss_seed", _p(self.out1[0]), _p(self.out0[0]), _p(self.out1[1]), _p(self.target_rgb), _p(self.target_disp),
             R, self.R_global, self.depth_lambda, _p(self.g_rgb), _p(self.g_rgb0), _p(self.g_disp), _p(self.loss), st)
        # ---- backward: fine network first (its all-reduce then overlaps the coarse network's backward) ----------
        works = {}
        for net in (1, 0):
            Sn = SF if net else S
            raw, z, noise = (self.raw1, self.z1, noise1) if net else (self.raw0, self.z0, noise0)
            g_raw = self.g_raw1 if net else self.g_raw0
            call("gbn_composite_backward", _p(raw), _p(z), _p(d), pitch, _p(noise), R, Sn, int(self.white), 0,
                 _p(self.g_rgb if net else self.g_rgb0), _p(self.g_disp) if net else None, None, None, None, _p(g_raw), st)
            call("gbn_mlp_backward_data", _p(self.packed[net][1]), _p(g_raw), R * Sn, _p(self.stash_h[net]),
                 _p(self.stash_g[net]), _p(self.ws_b[net]), st)
            call("gbn_mlp_backward_weights", _p(self.stash_h[net]), _p(self.stash_g[net]), _p(g_raw), _p(vd), pitch, R, Sn,
                 self.ptrs[net]["g"], C.c_void_p(self.ws_b[net].data_ptr() + 256), st)
            if self.world > 1:
                if self.overlap:
                    works[net] = tdist.all_reduce(self.flat[net], op=tdist.ReduceOp.SUM, async_op=True)
                else:
                    tdist.all_reduce(self.flat[net], op=tdist.ReduceOp.SUM)
        # ---- optimizer: torch.optim.Adam arithmetic + bf16 weight images patched in place, one launch per network ---
        # (fine network first: its all-reduce finished long ago, its Adam launch overlaps the coarse network's all-reduce)
        b1, b2 = self.group["betas"]
        call("gbn_adam_tick", _p(self.step_dev), _p(self.lr_dev), float(b1), float(b2), _p(self.scalars), st)
        for net in (1, 0):
            if net in works:
                works[net].wait()
            q = self.ptrs[net]
            call("gbn_adam_step_repack_dev", q["p"], q["g"], q["m"], q["v"], _p(self.scalars), float(b1), float(b2),
                 float(self.group["eps"]), _p(self.packed[net][0]), _p(self.packed[net][1]), st)

    def _capture(self):
        # warm-up on a side stream (one-time kernel attribute / constant uploads, NCCL communicator, RNG state
        # registration) - with the optimizer state restored afterwards, so that capture starts from the caller's state
        snap = [t.clone() for ps in self.params for t in ps]
        snap_m = [s[k].clone() for st in self.state for s in st for k in ("exp_avg", "exp_avg_sq")]
        step0 = self.step_dev.clone()
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self._launches()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        with torch.no_grad():
            for t, c in zip([t for ps in self.params for t in ps], snap):
                t.copy_(c)
            for t, c in zip([s_[k] for st in self.state for s_ in st for k in ("exp_avg", "exp_avg_sq")], snap_m):
                t.copy_(c)
            self.step_dev.copy_(step0)
        for i, n in enumerate(self.nets):       # the warm-up step patched the weight images: rebuilt from the restored
            fwd, bwd = n.packed_weights(), n.packed_weights_bwd()   # weights (their version changed), same buffers
            assert fwd.data_ptr() == self.packed[i][0].data_ptr() and bwd.data_ptr() == self.packed[i][1].data_ptr()
        torch.cuda.synchronize(self.dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.kernel_launches()
        # thread_local: NCCL's watchdog thread polls events of its own while the step is being captured
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self._launches()
        self.launches_per_step = _lib.kernel_launches() - n0    # kernels of libgbnerf.so inside one replay

    # ------------------------------------------------------------------------------------------------------
    def step(self, rays, target_rgb, target_disp, sharded_input=False, randoms=None):
        """rays: [R,11] packed batch (o, d, near, far, unit viewdir) or a (rays_o, rays_d) pair of [R,3]; targets [R,3]
        and [R].  CUDA tensors or pinned host tensors (copied with ``non_blocking=True``).  Returns the loss tensor
        [1] (device; this rank's share of the global mean - summed over ranks it is the reference's loss)."""
        sl = slice(None) if sharded_input or self.world == 1 else slice(self.lo, self.hi)
        if isinstance(rays, (tuple, list)):
            o, d = rays
            o, d = o.reshape(-1, 3)[sl], d.reshape(-1, 3)[sl]
            self.rays[:, 0:3].copy_(o, non_blocking=True)
            self.rays[:, 3:6].copy_(d, non_blocking=True)
            self.rays[:, 6].fill_(self.near)
            self.rays[:, 7].fill_(self.far)
            dn = self.rays[:, 3:6]
            self.rays[:, 8:11] = dn / torch.norm(dn, dim=-1, keepdim=True)
        else:
            self.rays.copy_(rays[sl], non_blocking=True)
        self.target_rgb.copy_(target_rgb[sl], non_blocking=True)
        self.target_disp.copy_(target_disp.reshape(-1)[sl], non_blocking=True)
        cur = float(self.state[0][0]["step"]) + getattr(self.opt, "_lazy_steps", 0)
        if cur != self._expected_step:           # the optimizer was stepped / reloaded outside: re-seed the device counter
            self.step_dev.fill_(cur)
        self.lr_host[0] = float(self.group["lr"])
        self.lr_dev.copy_(self.lr_host, non_blocking=True)
        if not self.use_graph or randoms is not None:
            self._launches(randoms)
        else:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        self.steps_done += 1
        self._expected_step = cur + 1.0
        self.opt._lazy_steps = getattr(self.opt, "_lazy_steps", 0) + 1    # FusedAdam folds it into state['step'] on demand
        return self.loss

    def outputs(self):
        """The last step's render outputs on this rank's rays (views of static buffers)."""
        return {"rgb_map": self.out1[0], "disp_map": self.out1[1], "acc_map": self.out1[2], "depth_map": self.out1[3],
                "rgb0": self.out0[0], "disp0": self.out0[1], "acc0": self.out0[2], "weights": self.w1, "z_vals": self.z1,
                "z_std": self.zstd}

    def error_codes(self):
        """Watchdog words of the four MLP launches (synchronises): all zero on a clean step."""
        return [ops.mlp_error_code(w) for w in (*self.ws_f, *self.ws_b)]
