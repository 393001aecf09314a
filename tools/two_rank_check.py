"""Standalone run of the two-rank GPU parity check (tests/test_gpu_train_step.py::_worker) with progress on stderr."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch.multiprocessing as mp
import test_gpu_train_step as T
if __name__ == "__main__":
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = T._free_port()
    procs = [ctx.Process(target=T._worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = {}
    try:
        for _ in range(2):
            k, v = q.get(timeout=120); res[k] = v
            if v != "ok": break
    except Exception as e:
        print("queue:", repr(e))
    for p in procs: p.join(20 if all(v == 'ok' for v in res.values()) and len(res) == 2 else 1)
    for p in procs:
        if p.is_alive(): p.kill()
    print(res)
