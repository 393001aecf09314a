"""CPU, world_size 2 over gloo: ray sharding, the flat gradient bucket all-reduce and the image gather
 — the host logic of the N>1 path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from gbnerf_b200 import dist as D
        R = 1001                                    # odd: blocks differ by one row
        rays = torch.arange(R * 11, dtype=torch.float32).reshape(R, 11)
        mine = D.shard_rays(rays)
        lo, hi = D.shard_bounds(R, rank, ws)
        assert mine.shape[0] == hi - lo and torch.equal(mine, rays[lo:hi])

        # gradient bucket: grads alias one flat buffer; one all-reduce sums them
        torch.manual_seed(0)
        lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
        bucket = D.GradBucket(lin.parameters())
        x = torch.full((4, 5), float(rank + 1))
        lin(x).sum().backward()
        local = [p.grad.clone() for p in lin.parameters()]
        assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in lin.parameters())
        bucket.all_reduce()
        gathered = [torch.zeros_like(torch.cat([g.flatten() for g in local])) for _ in range(ws)]
        dist.all_gather(gathered, torch.cat([g.flatten() for g in local]))
        assert torch.allclose(bucket.flat, sum(gathered))
        lin.zero_grad(set_to_none=True)
        bucket.rebind()
        assert all(p.grad is not None for p in lin.parameters())

        # the reference loop calls optimizer.zero_grad() (grads -> None): autograd then allocates fresh gradients
        # OUTSIDE the bucket; all_reduce() must pick them up instead of reducing a stale buffer
        lin.zero_grad(set_to_none=True)
        lin(x * 3).sum().backward()
        assert any(p.grad.data_ptr() < bucket.flat.data_ptr() or
                   p.grad.data_ptr() >= bucket.flat.data_ptr() + bucket.flat.numel() * 4 for p in lin.parameters())
        local2 = torch.cat([p.grad.flatten().clone() for p in lin.parameters()])
        bucket.all_reduce()
        gathered = [torch.zeros_like(local2) for _ in range(ws)]
        dist.all_gather(gathered, local2)
        assert torch.allclose(bucket.flat, sum(gathered))
        assert all(torch.equal(p.grad.flatten(), bucket.flat[o:o + p.numel()])
                   for p, o in zip(bucket.params, [0, 35, 42, 63]))

        # image gather through render_sharded with a stand-in renderer
        def fake_render(r, **kw):
            return {"rgb_map": r[:, 0:3] * 2, "disp_map": r[:, 3], "acc_map": r[:, 4], "depth_map": r[:, 5]}
        out, ret = D.render_sharded(fake_render, rays, dst=0)
        if rank == 0:
            assert torch.equal(out["rgb_map"], rays[:, 0:3] * 2) and torch.equal(out["depth_map"], rays[:, 5])
        else:
            assert out is None
        everywhere = D.gather_rows(mine, R, dst=None)
        assert torch.equal(everywhere, rays)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(ws))
    for p in procs:
        p.join(30)
    assert res == {0: "ok", 1: "ok"}, res


def test_shard_bounds_cover():
    from gbnerf_b200 import dist as D
    for n in (0, 1, 7, 4096, 762048):
        for ws in (1, 2, 4, 8):
            b = [D.shard_bounds(n, r, ws) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
