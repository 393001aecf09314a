"""CPU: job / step tables and barrier protocol of the two-tiles-in-flight inference MLP kernel (csrc/mlp_t2.cuh), replayed
by tests/t2_protocol_model.py under random latencies.  No GPU: gbn_debug_ts_plan is host code."""
import copy
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import t2_protocol_model as M  # noqa: E402
import ts_protocol_model as TS  # noqa: E402


@pytest.fixture(scope="module")
def plans():
    import __graft_entry__ as ge
    ge._load_builder().build()
    from gbnerf_b200 import _lib
    lib = _lib.load()
    return M.T2Plan(lib), TS.Plan(lib, 0)


def test_tables_cover_the_forward_image(plans):
    t2, fwd = plans
    assert t2.enabled and t2.mode == 1 and t2.stages == 4
    assert (len(t2.jobs), len(t2.steps)) == (39, 19)
    # every 128-row slab of the one-tile forward image is used exactly once; the two 16-row heads are decoded instead
    big = sorted(j.w_off for j in fwd.jobs if j.N == 128)
    assert sorted(j.w_off for j in t2.jobs) == big
    heads = sorted(j.w_off for j in fwd.jobs if j.N == 16)
    assert sorted([*t2.off_alpha, t2.off_rgb]) == heads
    # groups: from a FIRST job to the job that commits the accumulator; 19 per tile, one per epilogue step
    groups, j = 0, 0
    while j < len(t2.jobs):
        g = t2.jobs[j]
        assert g.glen in (1, 2, 3) and g.flags & M.F["FIRST"]
        members = t2.jobs[j:j + g.glen]
        assert all(not (m.flags & M.F["C_ACC"]) for m in members[:-1]) and members[-1].flags & M.F["C_ACC"]
        flags = 0
        for m in members:
            flags |= m.flags
        assert g.gflags == flags
        groups += 1
        j += g.glen
    assert groups == len(t2.steps)
    # steps and MMA groups are one to one: job0 of a step is the first job of its group
    firsts = [i for i, jb in enumerate(t2.jobs) if jb.glen]
    assert [s.job0 for s in t2.steps] == firsts
    assert [s.mode for s in t2.steps] == [M.HOLD, M.FLUSH] * 9 + [M.OUT]
    assert [i for i, s in enumerate(t2.steps) if s.dot] == [14, 15]       # layer 7 feeds alpha_linear


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("cold", [0.0, 0.3, 1.0])
def test_protocol_is_clean_under_random_latencies(plans, mode, cold):
    t2, _ = plans
    for seed in range(10):
        errs = M.simulate(t2, pairs=3, seed=seed, mode=mode, cold=cold)
        assert errs == [], (mode, cold, seed, errs[:3])


@pytest.mark.parametrize("flag,what", [("WAIT_EMPTY", "ACC overwritten"), ("WAIT_A", "reads A holding")])
def test_model_catches_a_dropped_wait(plans, flag, what):
    """The model is only worth something if it notices a broken table: drop one wait flag from one wide-layer job."""
    t2, _ = plans
    caught = 0
    for victim in (6, 8, 22):     # first jobs of groups in layers 2, 2 and 5/6
        jobs = copy.deepcopy(t2.jobs)
        # pick the next job at or after `victim` that carries the flag
        k = next(i for i in range(victim, len(jobs)) if jobs[i].flags & M.F[flag] and not jobs[i].flags & M.F["TILE_FIRST"])
        jobs[k].flags &= ~M.F[flag]
        jobs[k].gflags &= ~M.F[flag]
        for seed in range(6):
            errs = M.simulate(t2, pairs=2, seed=seed, mode=1, cold=0.3, jobs=jobs)
            if any(what in e or "alias" in e or "deadlock" in e for e in errs):
                caught += 1
                break
    assert caught == 3


def test_stash_blocks_match_the_one_tile_forward(plans):
    """The training form of the kernel writes the H stash that dgrad / wgrad read in the one-tile kernel's block numbering
    (csrc/mlp_layout.h): every step's out_blk must be the block the one-tile plan gives the same layer half, and the
    steps together must cover blocks 0 .. 37 exactly once (38 = encoding, 39 = directions: the input warps)."""
    t2, fwd = plans
    want = sorted(s.out_blk for s in fwd.steps if s.out_blk != 0xff)          # one 64-channel block pair per step
    hold = [s.out_blk for s in t2.steps if s.mode == M.HOLD]
    flush = [s.out_blk + 2 for s in t2.steps if s.mode == M.FLUSH]             # FLUSH writes the blocks after its HOLD's
    out = [s.out_blk for s in t2.steps if s.mode == M.OUT]
    assert [s.out_blk for s in t2.steps if s.mode == M.FLUSH] == [b + 2 for b in hold]
    assert sorted(hold + [b + 2 for b in hold] + out) == want
    covered = sorted(b + w for b in hold + [h + 2 for h in hold] + out for w in (0, 1))
    assert covered == list(range(38))
    del flush
