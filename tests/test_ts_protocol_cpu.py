"""CPU: the barrier protocol of the bf16 TMEM-operand MLP kernels (forward and dgrad programs), replayed by the
discrete-event model of tests/ts_protocol_model.py on the job / step tables the library really uses, under random
latencies including cold-cache weight fills.  No GPU: gbn_debug_ts_plan is host code."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ts_protocol_model as M  # noqa: E402


@pytest.fixture(scope="module")
def plans():
    import __graft_entry__ as ge
    ge._load_builder().build()
    from gbnerf_b200 import _lib
    lib = _lib.load()
    return {0: M.Plan(lib, 0), 1: M.Plan(lib, 1)}


def test_tables_are_the_documented_shape(plans):
    fwd, bwd = plans[0], plans[1]
    assert (len(fwd.jobs), fwd.stages, len(bwd.jobs), bwd.stages) == (42, 4, 37, 4)
    # two issuers, one ring: the jobs whose stage was last used by the other issuer carry the guard flag
    for p in (fwd, bwd):
        n = len(p.jobs)
        for j, jb in enumerate(p.jobs):
            other = p.jobs[(j - p.stages) % n].owner != jb.owner
            assert bool(jb.flags & M.TJ["PREV_OTHER"]) == other, (p.bwd, j)
            nxt = p.jobs[(j + p.stages) % n].owner != jb.owner
            assert bool(jb.ksteps & M.NEXT_OTHER) == nxt, (p.bwd, j)


@pytest.mark.parametrize("bwd", [0, 1])
@pytest.mark.parametrize("cold", [0.0, 0.3, 1.0])
def test_protocol_is_clean_under_random_latencies(plans, bwd, cold):
    for seed in range(12):
        errs = M.simulate(plans[bwd], tiles=4, seed=seed, cold=cold)
        assert errs == [], (bwd, cold, seed, errs[:3])


def test_dgrad_protocol_is_clean_without_the_split_handover(plans):
    for seed in range(8):
        errs = M.simulate(plans[1], tiles=3, seed=100 + seed, cold=0.3, no_split=True)
        assert errs == [], (seed, errs[:3])


def test_model_reproduces_the_ring_alias_when_the_guard_is_off(plans):
    """Validation of the model itself: without ts_wait_progress the dgrad program on a THREE-stage ring (4 jobs per layer;
    what the kernel had until the stash staging area left its shared memory) lets half 1's issuer test a stage on the
    other issuer's phase when weight fills are slow - the launch failure seen on the GPU (profiles/r1_ring_alias.md)."""
    import copy
    three = copy.copy(plans[1])
    three.stages = 3
    hits = 0
    for seed in range(12):
        errs = M.simulate(three, tiles=4, seed=seed, cold=1.0, guard=False)
        hits += any(e.split(": ", 1)[1].startswith(("alias", "rearm")) or "weight stage" in e for e in errs)
    assert hits >= 3, hits


@pytest.mark.parametrize("env", [{"GBNERF_TS_EARLY": "0", "GBNERF_TS_BWD_EARLY": "0"}, {"GBNERF_TS_SPLIT": "0"},
                                 {"GBNERF_TS_BWD_EARLY": "0", "GBNERF_TS_SPLIT": "0"}])
def test_alternative_plans_are_clean_too(plans, env):
    """The plans behind the runtime switches (late acc1 release, no split hand-over) are read once per process: replay
    them in a child process."""
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import ts_protocol_model as M\nfrom gbnerf_b200 import _lib\nlib = _lib.load()\nbad = []\n"
            "for bwd in (0, 1):\n    p = M.Plan(lib, bwd)\n    for seed in range(6):\n"
            "        bad += M.simulate(p, tiles=3, seed=seed, cold=0.4)\nprint('ERRORS', len(bad), bad[:3])\n"
            % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ERRORS 0 " in out.stdout, out.stdout[-2000:]
