"""GPU parity of train.TrainStep (one training iteration as one CUDA graph) against the drop-in autograd path
(render_rays + loss.backward() + FusedAdam.step()), which the other GPU tests pin to the oracle / the reference goldens;
and of the ray-sharded multi-GPU split against one GPU: sharded frame == single-GPU frame bit for bit,
all-reduced gradients of 2 x 2048 rays == gradients of the 4096-ray batch."""
import os
import socket

import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu

S, N = 64, 64


@pytest.fixture(scope="module")
def G():
    import gbnerf_b200
    return gbnerf_b200


def make(G, dev, perturb=1.0, noise=1.0, seed=0):
    torch.manual_seed(seed)
    pc, pf = O.init_params(seed), O.init_params(None)
    nets = []
    for p in (pc, pf):
        n = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True,
                   precision="bf16").to(dev)
        n.load_state_dict(p)
        nets.append(n)
    e10, _ = G.get_embedder(10, 0)
    e4, _ = G.get_embedder(4, 0)
    kw = dict(network_query_fn=G.NetworkQuery(e10, e4, 65536), perturb=perturb, N_importance=N, network_fine=nets[1],
              N_samples=S, network_fn=nets[0], use_viewdirs=True, white_bkgd=True, raw_noise_std=noise, ndc=False,
              lindisp=True, near=1.2, far=8.0)
    opt = G.FusedAdam([p for n in nets for p in n.parameters()], lr=3e-3, betas=(0.9, 0.999))
    return nets, kw, opt


def batch(R, seed=5):
    rays = O.synthetic_rays(R, seed=8)
    g = torch.Generator().manual_seed(seed)
    rnd = dict(t_rand=torch.rand(R, S, generator=g), noise0=torch.randn(R, S, generator=g),
               u=torch.rand(R, N, generator=g), noise1=torch.randn(R, S + N, generator=g))
    return rays, rnd, torch.rand(R, 3, generator=g), torch.rand(R, generator=g)


def autograd_step(G, nets, kw, opt, rays, rnd, tgt, tgd):
    kw = {k: v for k, v in kw.items() if k not in ("near", "far", "ndc", "use_viewdirs")}
    ret = G.render_rays(rays, _randoms=rnd, **kw)
    loss = G.img2mse(ret["rgb_map"], tgt) + G.img2mse(ret["rgb0"], tgt) + 0.1 * G.img2mse(ret["disp_map"], tgd)
    for n in nets:
        for p in n.parameters():
            p.grad = None
    loss.backward()
    grads = [p.grad.clone() for n in nets for p in n.param_list()]
    opt.step()
    return loss.detach(), grads, ret


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def test_eager_train_step_equals_autograd_path(G):
    """Same kernels launched directly: loss, every gradient and the parameters after Adam agree with the autograd path
    (wgrad sums with atomics, so gradients agree to rounding, not bit for bit)."""
    dev = torch.device("cuda:0")
    R = 192
    rays, rnd, tgt, tgd = batch(R)
    rays, tgt, tgd = rays.to(dev), tgt.to(dev), tgd.to(dev)
    rnd = {k: v.to(dev) for k, v in rnd.items()}

    nets_a, kw_a, opt_a = make(G, dev)
    loss_a, grads_a, ret_a = autograd_step(G, nets_a, kw_a, opt_a, rays, rnd, tgt, tgd)

    nets_b, kw_b, opt_b = make(G, dev)
    ts = G.TrainStep(kw_b, opt_b, R, graph=False)
    loss_b = ts.step(rays, tgt, tgd, randoms=rnd).clone()
    assert ts.error_codes() == [0, 0, 0, 0]
    out = ts.outputs()
    assert torch.equal(out["rgb_map"], ret_a["rgb_map"]) and torch.equal(out["disp_map"], ret_a["disp_map"])
    assert torch.equal(out["z_vals"], ret_a["z_vals"]) and torch.equal(out["rgb0"], ret_a["rgb0"])
    assert abs(loss_a.item() - loss_b.item()) < 1e-6 * max(1.0, abs(loss_a.item()))
    grads_b = [g for gs in ts.grads for g in gs]
    for i, (ga, gb) in enumerate(zip(grads_a, grads_b)):
        assert rel(gb, ga) < 1e-4, (i, rel(gb, ga))
    init = (O.init_params(0), O.init_params(None))
    for na, nb, p0 in zip(nets_a, nets_b, init):
        for (name, pa), (_, pb) in zip(na.named_parameters(), nb.named_parameters()):
            moved = (pa.detach().cpu() - p0[name]).norm().item()   # one Adam step moves every weight by ~lr = 3e-3
            assert (pa - pb).norm().item() <= 0.02 * moved, name
        # the kernel patched the bf16 images in place: identical to a fresh re-pack of the new weights
        # (re-packing over a copy: bytes the packer never writes - alignment gaps - keep their old content)
        fwd, bwd = nb.packed_weights(), nb.packed_weights_bwd()
        assert torch.equal(fwd, G.ops.prepack_weights(nb.param_list(), "bf16", out=fwd.clone()))
        assert torch.equal(bwd, G.ops.prepack_weights(nb.param_list(), "bf16_bwd", out=bwd.clone()))
    assert opt_b.state_dict()["state"][0]["step"] == 1


def test_graphed_train_step_replays(G):
    """The captured graph: deterministic kwargs (perturb 0, no noise) make replays comparable with the eager launches;
    two steps from the same start give the same parameters; step count and lr live on the device."""
    dev = torch.device("cuda:0")
    R = 256
    rays, _, tgt, tgd = batch(R)
    rays, tgt, tgd = rays.to(dev), tgt.to(dev), tgd.to(dev)
    res = []
    for graph in (False, True):
        nets, kw, opt = make(G, dev, perturb=0.0, noise=0.0)
        ts = G.TrainStep(kw, opt, R, graph=graph)
        losses = []
        for it in range(3):
            opt.param_groups[0]["lr"] = 3e-3 * (0.5 ** it)           # run.py:1540-1544 assigns a decayed lr every step
            losses.append(ts.step(rays, tgt, tgd).item())
        assert ts.error_codes() == [0, 0, 0, 0]
        assert opt.state_dict()["state"][0]["step"] == 3
        res.append((losses, [p.detach().clone() for n in nets for _, p in n.named_parameters()]))
        if graph:
            assert ts.launches_per_step is not None and ts.launches_per_step >= 14
    init = [v for p in (O.init_params(0), O.init_params(None)) for v in p.values()]
    (l0, p0), (l1, p1) = res
    for a, b in zip(l0, l1):
        assert abs(a - b) < 1e-5 * max(1.0, abs(a)), (l0, l1)
    assert l0[2] != l0[0]
    # wgrad sums with atomics (order-dependent rounding) and Adam divides by sqrt(v): an element whose gradient is
    # rounding noise can move by a full step either way, so the two runs are compared against the size of the update
    for a, b, p0_ in zip(p0, p1, init):
        moved = (a.cpu() - p0_).norm().item()
        assert moved > 0 and (a - b).norm().item() <= 0.02 * moved, ((a - b).norm().item(), moved)


def test_graphed_train_step_draws_fresh_randoms(G):
    """Train kwargs (perturb 1, noise 1): every replay draws new random tensors (torch's graph-safe Philox offsets)."""
    dev = torch.device("cuda:0")
    R = 128
    rays, _, tgt, tgd = batch(R)
    nets, kw, opt = make(G, dev)
    ts = G.TrainStep(kw, opt, R)
    ts.step(rays.to(dev), tgt.to(dev), tgd.to(dev))
    t0, u0 = ts.t_rand.clone(), ts.u.clone()
    ts.step(rays.to(dev), tgt.to(dev), tgd.to(dev))
    assert not torch.equal(t0, ts.t_rand) and not torch.equal(u0, ts.u)
    assert 0.0 <= ts.t_rand.min().item() and ts.t_rand.max().item() < 1.0
    assert abs(ts.noise1.std().item() - 1.0) < 0.05
    assert ts.error_codes() == [0, 0, 0, 0] and torch.isfinite(ts.loss).all()


# ---- two ranks ------------------------------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, q):
    import faulthandler
    import sys
    faulthandler.dump_traceback_later(100, exit=True)      # a hung collective must not hang the GPU box: stacks, then exit
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import gbnerf_b200 as G
    say = lambda m: print(f"[two-rank test, rank {rank}] {m}", file=sys.stderr, flush=True)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev)
    try:
        # ---- inference: ONE frame sharded by contiguous blocks, gathered on rank 0 -------------------------------
        say("process group up")
        R = 3001                                                         # odd: blocks differ by one row
        rays = O.synthetic_rays(R, seed=3).to(dev)
        nets, kw, opt = make(G, dev, perturb=0.0, noise=0.0)
        rkw = {k: v for k, v in kw.items() if k not in ("near", "far", "ndc", "use_viewdirs")}
        with torch.no_grad():
            out, _ = G.dist.render_sharded(lambda r, **k: G.batchify_rays(r, 1024, **k), rays, dst=0, **rkw)
            if rank == 0:
                full = G.batchify_rays(rays, 1024, **rkw)
                for k in ("rgb_map", "disp_map", "acc_map", "depth_map"):
                    assert torch.equal(out[k], full[k], ), f"sharded {k} != single-GPU {k}"
            else:
                assert out is None
        say("sharded frame == single-GPU frame")
        # ---- training: ONE 4096-ray batch, 2048 per rank, gradients all-reduced --------------------------------
        Rb = 4096
        rays, _, tgt, tgd = batch(Rb)
        rays, tgt, tgd = rays.to(dev), tgt.to(dev), tgd.to(dev)
        for graph in (False, True):
            nets, kw, opt = make(G, dev, perturb=0.0, noise=0.0)
            ts = G.TrainStep(kw, opt, Rb, graph=graph)
            assert ts.R == Rb // ws
            say(f"TrainStep(graph={graph}) built")
            loss = ts.step(rays, tgt, tgd).clone()
            torch.cuda.synchronize()
            say(f"TrainStep(graph={graph}) stepped")
            dist.all_reduce(loss)
            assert ts.error_codes() == [0, 0, 0, 0]
            grads = [g.clone() for gs in ts.grads for g in gs]
            params = [p.detach().clone() for n in nets for p in n.param_list()]
            if rank == 0:
                # single-GPU reference of the same batch in this process (world-size-1 TrainStep semantics by hand)
                nets1, kw1, opt1 = make(G, dev, perturb=0.0, noise=0.0)
                l1, g1, _ = autograd_step(G, nets1, kw1, opt1, rays, None, tgt, tgd)
                assert abs(l1.item() - loss.item()) < 1e-5 * max(1.0, abs(l1.item())), (l1.item(), loss.item())
                for i, (a, b) in enumerate(zip(grads, g1)):
                    assert rel(a, b) < 1e-3, (graph, i, rel(a, b))      # other tile composition + atomics order: rounding
                init = [v for p0 in (O.init_params(0), O.init_params(None)) for v in p0.values()]
                named1 = [p for n in nets1 for _, p in n.named_parameters()]
                named = [p.detach() for n in nets for _, p in n.named_parameters()]
                for a, n1, p0 in zip(named, named1, init):       # Adam amplifies rounding noise of near-zero gradients:
                    moved = (n1.detach().cpu() - p0).norm().item()   # compare against the size of the update
                    assert (a - n1).norm().item() <= 0.02 * moved, ((a - n1).norm().item(), moved)
            # replicas stay identical: every rank applied the same summed gradient
            chk = torch.stack([p.double().sum() for p in params])
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), "parameter replicas diverged"
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()[-1500:]))
        q.close(); q.join_thread()
        os._exit(1)        # the peer may sit in a collective this rank will never join: leave without tearing NCCL down
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_match_one_gpu():
    import torch.multiprocessing as mp
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = {}
    try:
        for _ in range(ws):
            k, v = q.get(timeout=130)
            res[k] = v
            if v != "ok":        # the other rank may be waiting in a collective for this one: do not wait for it
                break
    finally:
        for p in procs:
            p.join(60 if len(res) == ws and all(v == "ok" for v in res.values()) else 1)
        for p in procs:
            if p.is_alive():
                p.kill()
    assert res == {0: "ok", 1: "ok"}, res
