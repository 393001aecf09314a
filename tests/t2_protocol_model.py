"""Discrete-event model of the barrier protocol of nerf_mlp_t2_kernel (gb-nerf_b200/csrc/mlp_t2.cuh: two tiles in flight
per CTA), driven by the job / step tables the library really uses (gbn_debug_ts_plan(2, ...)).  Test infrastructure, like
tests/ts_protocol_model.py, whose event engine and parity-only mbarriers it reuses: it restates which role waits on which
barrier with which parity and who arrives where, replays one CTA under random latencies, and checks from the operand
accesses themselves that

  * a weight stage is refilled only after BOTH issuers' MMAs on its previous slab have completed, and MMAs find the
    slab they expect in their stage,
  * an accumulator is overwritten only after the epilogue has read the previous half out of it,
  * the in-place activation buffer A_s is overwritten only when no MMA that reads it is still pending, and MMAs read the
    layer output they expect,
  * the encoding / direction blocks are rewritten only after their last reader has completed,
  * no parity wait passes for a completion that has not happened (alias), nobody deadlocks, no barrier over-arrives.
"""
import ctypes as C
import struct

from ts_protocol_model import Barrier, Sim

F = dict(WAIT_ENC=1, WAIT_DIR=2, WAIT_A=4, WAIT_EMPTY=8, FIRST=16, A_ENC=32, A_DIR=64, C_ACC=128, C_ENC=256, C_DIR=512,
         TILE_FIRST=1024)
HOLD, FLUSH, OUT = 0, 1, 2


class T2Job:
    def __init__(self, raw):
        (self.w_off, self.bytes16, self.flags, self.a_col, self.ksteps, self.nkb, self.glen, _p,
         self.gflags) = struct.unpack("<IHHHBBBBH", raw)
        self.n_mma = (self.ksteps if self.flags & F["A_ENC"] else 2) if self.flags & (F["A_ENC"] | F["A_DIR"]) else 8


class T2Step:
    def __init__(self, raw):
        # job0: first job of the step's MMA group; out_blk: first H-stash block of the step's output (training forward)
        self.mode, self.relu, self.dot, self.job0, self.bias_off, self.out_blk = struct.unpack("<BBBBHH", raw)


class T2Plan:
    def __init__(self, lib):
        jobs, steps, meta = (C.c_uint8 * (16 * 96))(), (C.c_uint8 * (12 * 24))(), (C.c_int * 10)()
        assert lib.gbn_debug_ts_plan(2, jobs, 96, steps, 24, meta) == 0
        self.jobs = [T2Job(bytes(jobs[16 * i:16 * i + 16])) for i in range(meta[0])]
        self.steps = [T2Step(bytes(steps[8 * i:8 * i + 8])) for i in range(meta[1])]
        self.stages, self.mode = meta[2], meta[3]
        self.off_alpha, self.off_rgb, self.enabled = (meta[4], meta[5]), meta[6], bool(meta[7])


def simulate(plan, pairs=3, seed=0, mode=None, cold=0.2, jobs=None):
    """One CTA working through `pairs` tile pairs.  Returns the list of protocol errors (empty = clean)."""
    mode = plan.mode if mode is None else mode
    jobs = plan.jobs if jobs is None else jobs
    sim = Sim(seed)
    rng = sim.rng
    NST, nj, nsteps = plan.stages, len(jobs), len(plan.steps)
    nlayers = (nsteps - 1) // 2
    ecount = 256 if mode == 1 else 128
    B = lambda name, count: Barrier(sim, name, count)
    w_full = [B(f"w_full[{i}]", 1) for i in range(NST)]
    w_empty = [B(f"w_empty[{i}]", 2) for i in range(NST)]
    acc_full = [B(f"acc_full[{s}]", 1) for s in range(2)]
    acc_empty = [B(f"acc_empty[{s}]", ecount) for s in range(2)]
    a_ready = [B(f"a_ready[{s}]", ecount) for s in range(2)]
    enc_full = [B(f"enc_full[{s}]", 128) for s in range(2)]
    enc_empty = [B(f"enc_empty[{s}]", 1) for s in range(2)]
    dir_full = [B(f"dir_full[{s}]", 128) for s in range(2)]
    dir_empty = [B(f"dir_empty[{s}]", 1) for s in range(2)]
    turn = {"v": 0, "w": []}

    # ---- operand state -------------------------------------------------------------------------------------------------
    stage_slab = [None] * NST               # (it, j) of the slab that has landed
    stage_pending = [0] * NST               # MMA jobs issued on the stage that have not completed
    nwg = 2 if mode == 1 else 1             # epilogue processes that read every accumulator
    acc_need = [0, 0]                       # epilogue processes that have not read the accumulator out yet
    acc_group = [None, None]                # (it, group) accumulating / held in ACC_s
    a_tag = [None, None]                    # (it, layer) whose output is in A_s
    a_pending = [0, 0]                      # MMA jobs reading A_s that have not completed
    blk_tag = {"enc": [None, None], "dir": [None, None]}
    blk_pending = {"enc": [0, 0], "dir": [0, 0]}
    pipe = {"free": 0}
    # layer of every job (A_s must hold the previous layer's output) and group index within the tile
    layer_of, group_of, layer, grp = [], [], 0, 0
    for j, jb in enumerate(jobs):
        layer_of.append(layer)
        group_of.append(grp)
        if jb.flags & F["C_ACC"]:
            grp += 1
            if grp % 2 == 0:
                layer += 1
    groups_per_tile = grp

    def lat(lo, hi):
        return rng.randint(lo, hi)

    def set_turn(v):
        turn["v"] = v
        ws, turn["w"] = turn["w"], []
        for w in ws:
            sim.poll(w)

    # ---- tensor pipe: jobs execute in issue order; commits arrive after completion ---------------------------------------
    def issue(slot, it, j, st):
        jb = jobs[j]
        start = max(sim.now + lat(5, 40), pipe["free"])
        end = start + 64 * jb.n_mma
        pipe["free"] = end
        stage_pending[st] += 1
        reads_a = not (jb.flags & (F["A_ENC"] | F["A_DIR"]))
        blk = "enc" if jb.flags & F["A_ENC"] else ("dir" if jb.flags & F["A_DIR"] else None)
        if reads_a:
            a_pending[slot] += 1
        else:
            blk_pending[blk][slot] += 1

        def begin():
            if stage_slab[st] != (it, j):
                sim.error(f"slot {slot} job {j} (pair {it}) runs on stage {st} holding {stage_slab[st]}")
            if jb.flags & F["FIRST"]:
                if acc_need[slot]:
                    sim.error(f"slot {slot} job {j}: ACC overwritten before group {acc_group[slot]} was read out")
                acc_need[slot] = nwg
                acc_group[slot] = (it, group_of[j])
            elif acc_group[slot] != (it, group_of[j]):
                sim.error(f"slot {slot} job {j}: accumulates onto group {acc_group[slot]}")
            if reads_a and a_tag[slot] != (it, layer_of[j] - 1):
                sim.error(f"slot {slot} job {j} (layer {layer_of[j]}) reads A holding {a_tag[slot]}")
            if blk and blk_tag[blk][slot] != it:
                sim.error(f"slot {slot} job {j} reads the {blk} block of pair {blk_tag[blk][slot]}, wants {it}")

        def done():
            stage_pending[st] -= 1
            if reads_a:
                a_pending[slot] -= 1
            else:
                blk_pending[blk][slot] -= 1
            d = lat(150, 300)
            sim.at(sim.now + d, lambda: w_empty[st].arrive())
            if jb.flags & F["C_ACC"]:
                sim.at(sim.now + d, lambda: acc_full[slot].arrive())
            if jb.flags & F["C_ENC"]:
                sim.at(sim.now + d, lambda: enc_empty[slot].arrive())
            if jb.flags & F["C_DIR"]:
                sim.at(sim.now + d, lambda: dir_empty[slot].arrive())

        sim.at(start, begin)
        sim.at(end, done)
        return end

    # ---- roles -----------------------------------------------------------------------------------------------------------
    def producer():
        s = par = 0
        for it in range(pairs):
            for j in range(nj):
                yield ("bar", w_empty[s], par ^ 1, None, "")
                st, tag = s, (it, j)

                def land(st=st, tag=tag):
                    if stage_pending[st]:
                        sim.error(f"stage {st} refilled with {tag} under {stage_pending[st]} pending MMA job(s)")
                    stage_slab[st] = tag
                    w_full[st].arrive()

                sim.at(sim.now + (lat(1500, 4000) if rng.random() < cold else lat(300, 900)), land)
                yield ("delay", lat(20, 60))
                s += 1
                if s == NST:
                    s, par = 0, par ^ 1

    def waits(slot, it, fl, ph):
        if fl & F["WAIT_ENC"]:
            yield ("bar", enc_full[slot], it & 1, it, "enc")
        if fl & F["WAIT_DIR"]:
            yield ("bar", dir_full[slot], it & 1, it, "dir")
        if fl & F["WAIT_A"]:
            yield ("bar", a_ready[slot], ph["a"] & 1, ph["a"], "a_ready")
            ph["a"] += 1
        if (fl & F["WAIT_EMPTY"]) and not ((fl & F["TILE_FIRST"]) and it == 0):
            yield ("bar", acc_empty[slot], ph["e"] & 1, ph["e"], "acc_empty")
            ph["e"] += 1

    def issuer_free(slot):       # MODE 0: one job at a time, no ordering between the slots
        s = par = 0
        ph = {"a": 0, "e": 0}
        for it in range(pairs):
            for j, jb in enumerate(jobs):
                yield from waits(slot, it, jb.flags, ph)
                yield ("bar", w_full[s], par, None, "")
                issue(slot, it, j, s)
                yield ("delay", lat(100, 500))
                s += 1
                if s == NST:
                    s, par = 0, par ^ 1

    def issuer_groups(slot):     # MODE 1: groups back to back, slot 0 / slot 1 in alternation
        s = par = 0
        ph = {"a": 0, "e": 0}
        g = slot
        for it in range(pairs):
            j = 0
            while j < nj:
                glen, fl = jobs[j].glen, jobs[j].gflags
                if glen == 0:
                    sim.error(f"job {j} starts no group")
                    return
                ss, pp = s, par
                for k in range(glen):
                    yield ("bar", w_full[ss], pp, None, "")
                    ss += 1
                    if ss == NST:
                        ss, pp = 0, pp ^ 1
                yield from waits(slot, it, fl, ph)
                yield ("prog", turn, g)
                ss = s
                for k in range(glen):
                    if k == glen - 1:
                        set_turn(g + 1)
                    issue(slot, it, j + k, ss)
                    yield ("delay", lat(30, 300))
                    ss = (ss + 1) % NST
                for k in range(glen):
                    s += 1
                    if s == NST:
                        s, par = 0, par ^ 1
                j += glen
                g += 2

    def read_acc(slot, it, group):
        if acc_group[slot] != (it, group):
            sim.error(f"epilogue reads ACC of slot {slot} holding {acc_group[slot]}, wants {(it, group)}")

    def write_a(slot, it, layer):
        if a_pending[slot]:
            sim.error(f"A of slot {slot} overwritten (layer {layer}) under {a_pending[slot]} pending MMA job(s)")
        a_tag[slot] = (it, layer)

    def epi_step(slot, it, li, half, ph, nthreads):
        """One HOLD / FLUSH step of `nthreads` epilogue threads on one slot."""
        n = it * (2 * nlayers + 1) + 2 * li + half
        yield ("bar", acc_full[slot], ph[slot] & 1, n, f"acc_full step {2 * li + half}")
        ph[slot] += 1
        yield ("delay", lat(150, 400))                   # tcgen05.ld
        read_acc(slot, it, 2 * li + half)
        acc_need[slot] -= 1
        if half == 0:
            acc_empty[slot].arrive(nthreads)
            yield ("delay", lat(200, 500))               # conversion
        else:
            write_a(slot, it, li)                        # tcgen05.st of both halves starts here
            yield ("delay", lat(300, 700))               # conversions + tcgen05.st + wait::st
            a_ready[slot].arrive(nthreads)

    def epi_out(slot, it, ph, nthreads):
        n = it * (2 * nlayers + 1) + 2 * nlayers
        yield ("bar", acc_full[slot], ph[slot] & 1, n, "acc_full out")
        ph[slot] += 1
        yield ("delay", lat(150, 400))
        read_acc(slot, it, 2 * nlayers)
        acc_need[slot] -= 1
        acc_empty[slot].arrive(nthreads)
        yield ("delay", lat(500, 1500))                  # rgb_linear on the CUDA cores

    def epilogue_slot(slot):     # MODE 0: one warpgroup per slot
        ph = [0, 0]
        for it in range(pairs):
            for li in range(nlayers):
                yield from epi_step(slot, it, li, 0, ph, 128)
                yield from epi_step(slot, it, li, 1, ph, 128)
            yield from epi_out(slot, it, ph, 128)

    def epilogue_shared(wg):     # MODE 1: both warpgroups serve both slots in the order of the alternating issue
        ph = [0, 0]
        for it in range(pairs):
            for li in range(nlayers):
                for half in (0, 1):
                    for slot in (0, 1):
                        yield from epi_step(slot, it, li, half, ph, 128)
            for slot in (0, 1):
                yield from epi_out(slot, it, ph, 128)
            yield ("delay", lat(50, 300))                # named barrier + store (not modelled as a barrier)

    def input_warps():
        for it in range(pairs):
            for blk, full, empty in (("enc", enc_full, enc_empty), ("dir", dir_full, dir_empty)):
                for slot in (0, 1):
                    yield ("delay", lat(300, 1500))
                    if it > 0:
                        yield ("bar", empty[slot], (it - 1) & 1, it - 1, f"{blk}_empty")
                    if blk_pending[blk][slot]:
                        sim.error(f"{blk} block of slot {slot} rewritten under a pending MMA job")
                    blk_tag[blk][slot] = it
                    full[slot].arrive(128)

    sim.spawn("producer", producer())
    sim.spawn("input", input_warps())
    if mode == 1:
        sim.spawn("issuer0", issuer_groups(0))
        sim.spawn("issuer1", issuer_groups(1))
        sim.spawn("epilogue wg0", epilogue_shared(0))
        sim.spawn("epilogue wg1", epilogue_shared(1))
    else:
        sim.spawn("issuer0", issuer_free(0))
        sim.spawn("issuer1", issuer_free(1))
        sim.spawn("epilogue slot0", epilogue_slot(0))
        sim.spawn("epilogue slot1", epilogue_slot(1))
    sim.run()
    return sim.errors
