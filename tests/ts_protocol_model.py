"""Discrete-event model of the barrier protocol of nerf_mlp_ts_kernel (gb-nerf_b200/csrc/mlp_ts.cu), driven by the job and
step tables the kernels really use (gbn_debug_ts_plan).  Test infrastructure: it restates the kernel's control flow
(which role waits on which mbarrier with which parity, who arrives where) and replays it under random latencies, with
mbarriers that expose only the PARITY of their phase, as the hardware does.  It reports

  * alias     - a parity wait passed although the completion it was meant for had not happened yet,
  * rearm     - the weight producer armed a stage whose previous fill was still in flight,
  * overflow  - more arrivals on a barrier than its phase expects,
  * deadlock  - a role never finished,
  * hazards   - a tensor-memory / shared-memory operand read or overwritten out of order (accumulator halves, the
                activation buffers A0/A1 at 16-column granularity, encoding blocks, weight stages, gate staging), from access logs.

The kernel is the product; this model only checks its protocol (it found nothing the GPU did not: it reproduces the
weight-ring alias of DESIGN.md §3.1 when the ring guard is switched off, which is how it is itself validated).
"""
import ctypes as C
import heapq
import random
import struct

TJ = dict(ENC=1, A0=2, A1=4, TILE=8, FIRST=16, C_ENC=32, C_ACC0=64, C_ACC1=128, A_SMEM=256, S_ORD=512, W_ORD=1024,
          A_DIR=2048, W_DIR=4096, C_DIR=8192, EMPTY1=16384, PREV_OTHER=32768)
NEXT_OTHER = 8
EPI_OUT, EPI_MASK = 3, 4
ACC1, A0COL = 128, 256


class Job:
    def __init__(self, raw):
        (self.w_off, self.w_bytes16, self.flags, self.d_col, self.a_col, self.n16, self.nkb, self.ksteps,
         self.wait_buf) = struct.unpack("<IHHHHBBBB", raw)
        self.owner = 1 if self.d_col >= ACC1 else 0          # issuer 0 = warp 1 (acc0), issuer 1 = warp 3 (acc1)
        self.N = self.n16 * 16


class Step:
    def __init__(self, raw):
        (self.acc, self.mode, self.out_buf, self.out_half, self.no_act, self.mask_blk, self.out_blk, _p,
         self.bias_off, _p2) = struct.unpack("<BBBBBBBBHH", raw)


class Plan:
    def __init__(self, lib, bwd):
        jobs, steps, meta = (C.c_uint8 * (16 * 96))(), (C.c_uint8 * (12 * 24))(), (C.c_int * 10)()
        rc = lib.gbn_debug_ts_plan(int(bwd), jobs, 96, steps, 24, meta)
        assert rc == 0
        self.bwd = bool(bwd)
        self.jobs = [Job(bytes(jobs[16 * i:16 * i + 16])) for i in range(meta[0])]
        self.steps = [Step(bytes(steps[12 * i:12 * i + 12])) for i in range(meta[1])]
        self.ready = list(meta[2:6])
        self.empty1 = meta[7]
        self.stages = meta[8]
        self.split = bool(meta[9])


class Barrier:
    def __init__(self, sim, name, count):
        self.sim, self.name, self.count, self.pending, self.phase, self.waiters = sim, name, count, count, 0, []

    def passes(self, parity):                      # mbarrier.try_wait.parity: true once the phase of that parity is over
        return (self.phase & 1) != parity

    def arrive(self, n=1):
        self.pending -= n
        if self.pending < 0:
            self.sim.error(f"overflow: {self.name} got more arrivals than its phase expects")
            self.pending = 0
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count
            ws, self.waiters = self.waiters, []
            for w in ws:
                self.sim.poll(w)


class Sim:
    def __init__(self, seed):
        self.now, self.q, self.n, self.errors, self.rng = 0, [], 0, [], random.Random(seed)
        self.blocked = {}

    def error(self, msg):
        if len(self.errors) < 20:
            self.errors.append(f"t={self.now}: {msg}")

    def at(self, t, fn):
        self.n += 1
        heapq.heappush(self.q, (max(t, self.now), self.n, fn))

    def spawn(self, name, gen):
        self.blocked[name] = None
        self.at(self.now, lambda: self._step(name, gen))

    def poll(self, w):
        name, gen, cond = w
        self.at(self.now, lambda: self._resume(name, gen, cond))

    def _resume(self, name, gen, cond):
        kind = cond[0]
        if kind == "bar":
            _, bar, parity, want, label = cond
            if not bar.passes(parity):
                bar.waiters.append((name, gen, cond))
                return
            if want is not None and bar.phase <= want:
                self.error(f"alias: {name} passed {bar.name} for completion #{want} ({label}) in phase {bar.phase}")
        elif kind == "prog":
            _, box, need = cond
            if box["v"] < need:
                box["w"].append((name, gen, cond))
                return
        self.blocked[name] = None
        self._step(name, gen)

    def _step(self, name, gen):
        try:
            req = next(gen)
        except StopIteration:
            del self.blocked[name]
            return
        if req[0] == "delay":
            self.at(self.now + req[1], lambda: self._step(name, gen))
        else:
            self.blocked[name] = req
            self._resume(name, gen, req)

    def run(self, limit=5_000_000):
        while self.q and self.n < limit:
            t, _, fn = heapq.heappop(self.q)
            self.now = t
            fn()
        for name, cond in self.blocked.items():
            what = cond[1].name if cond and cond[0] == "bar" else str(cond and cond[0])
            self.error(f"deadlock: {name} never finished (last wait: {what})")


def simulate(plan, tiles=4, seed=0, guard=True, cold=0.25, no_split=False):
    """One CTA working through `tiles` tiles.  cold = probability that a weight fill takes a cold-cache latency.
    Returns the list of protocol errors (empty = clean)."""
    sim = Sim(seed)
    rng = sim.rng
    P, NST, BWD = plan, plan.stages, plan.bwd
    kSplit = BWD                                   # compiled into the dgrad program only
    use_split = kSplit and plan.split and not no_split
    B = lambda name, count: Barrier(sim, name, count)
    w_full = [B(f"w_full[{i}]", 1) for i in range(NST)]
    w_empty = [B(f"w_empty[{i}]", 1) for i in range(NST)]
    acc_full = [B("acc_full[0]", 1), B("acc_full[1]", 1)]
    a_ready = [B(f"a_ready[{i >> 1}][{i & 1}]", 256) for i in range(4)]
    a_ready_b = [B(f"a_ready_b[{i}]", 256) for i in range(2)]
    enc_full, enc_empty, tile_done = B("enc_full", 128), B("enc_empty", 2), B("tile_done", 256)
    dir_full, dir_empty, acc1_empty = B("dir_full", 128), B("dir_empty", 1), B("acc1_empty", 256)
    m_full = [B(f"m_full[{i}]", 1) for i in range(2)]
    m_empty = [B(f"m_empty[{i}]", 256) for i in range(2)]
    prog = [{"v": 0, "w": []}, {"v": 0, "w": []}]
    pipe = {"free": 0}
    log = dict(fill=[], wread=[], accw=[], accr=[], ast=[], ard=[], sw=[], sr=[], gfill=[], gread=[])
    stage_busy = [False] * NST
    njobs = len(P.jobs)
    # ordinal of the acc1_empty completion each EMPTY1 job waits for, and commit group of every job / step
    e1_ord, k = {}, 0
    for j, jb in enumerate(P.jobs):
        if jb.flags & TJ["EMPTY1"]:
            e1_ord[j] = k
            k += 1
    grp, g = {}, [0, 0]
    for j, jb in enumerate(P.jobs):
        grp[j] = g[jb.owner]
        if jb.flags & (TJ["C_ACC1"] if jb.owner else TJ["C_ACC0"]):
            g[jb.owner] += 1
    groups_per_tile = list(g)
    sgrp, g = {}, [0, 0]
    for si, st in enumerate(P.steps):
        sgrp[si] = g[st.acc]
        g[st.acc] += 1
    assert g == groups_per_tile, "every committed accumulator group has exactly one epilogue step"

    def wait(bar, parity, want, label=""):
        return ("bar", bar, parity & 1, want, label)

    def producer():
        cnt = 0
        for t in range(tiles):
            for j, jb in enumerate(P.jobs):
                s, n = cnt % NST, cnt // NST
                yield wait(w_empty[s], (n & 1) ^ 1, n - 1 if n > 0 else None, f"release before fill of job {j}")
                if stage_busy[s]:
                    sim.error(f"rearm: fill of job {j} (tile {t}) armed w_full[{s}] while its previous fill was in flight")
                stage_busy[s] = True
                lat = rng.randint(2500, 7000) if rng.random() < cold else rng.randint(250, 900)
                t0 = sim.now

                def land(s=s, t=t, j=j, t0=t0):
                    stage_busy[s] = False
                    log["fill"].append((s, (t, j), t0, sim.now))
                    w_full[s].arrive()
                sim.at(sim.now + lat, land)
                yield ("delay", rng.randint(20, 60))
                cnt += 1

    def issuer(me):
        cnt, last_end = 0, 0
        for t in range(tiles):
            for j, jb in enumerate(P.jobs):
                if jb.owner != me:
                    cnt += 1
                    continue
                f, wb = jb.flags, jb.wait_buf
                yield ("delay", rng.randint(20, 80))
                split = False
                if f & TJ["ENC"]:
                    yield wait(enc_full, t, t, "encoding block")
                if f & TJ["W_DIR"]:
                    yield wait(dir_full, t, t, "direction block")
                if (f & TJ["TILE"]) and t > 0:
                    yield wait(tile_done, t - 1, t - 1, "previous tile drained")
                if f & TJ["A0"]:
                    b = (wb & 1) * 2
                    seq = t * P.ready[b] + ((wb >> 1) & 7)
                    yield wait(a_ready[b], seq, seq, f"input half 0, job {j}")
                if f & TJ["A1"]:
                    b = (wb & 1) * 2 + 1
                    seq = t * P.ready[b] + ((wb >> 4) & 7)
                    yield wait(a_ready[b], seq, seq, f"input half 1, job {j}")
                    if kSplit:
                        if (f & TJ["A_SMEM"]) or jb.nkb != 2 or not use_split:
                            yield wait(a_ready_b[wb & 1], seq, seq, f"second instalment, job {j}")
                        else:
                            split = True
                if f & TJ["EMPTY1"]:
                    seq = t * P.empty1 + ((wb >> 7) & 1)
                    yield wait(acc1_empty, seq, t * P.empty1 + e1_ord[j], f"acc1 drained, job {j}")
                assert not (f & (TJ["S_ORD"] | TJ["W_ORD"])), "issue-order signals are not modelled"
                s, n = cnt % NST, cnt // NST
                if guard and (f & TJ["PREV_OTHER"]) and cnt >= NST:
                    yield ("prog", prog[1 - me], cnt - NST + 1)
                yield wait(w_full[s], n, n, f"fill of job {j}")
                # ---- issue: one MMA per 16-wide K step; a shared tensor pipe executes them in issue order
                a_smem = bool(f & TJ["A_SMEM"])
                ks = jb.ksteps & 7
                order = list(range(ks)) if a_smem else list(range(4 * jb.nkb))
                if split:
                    order = [0, 1, 4, 5, 2, 3, 6, 7]
                dur = max(8, jb.N // 2)
                first_t0 = None
                for i, kk in enumerate(order):
                    if split and i == 4:
                        seq = t * P.ready[(wb & 1) * 2 + 1] + ((wb >> 4) & 7)
                        yield wait(a_ready_b[wb & 1], seq, seq, f"second instalment (mid-issue), job {j}")
                    yield ("delay", rng.randint(30, 70))
                    t0 = max(sim.now, pipe["free"])
                    t1 = t0 + dur
                    pipe["free"] = t1
                    last_end = max(last_end, t1)
                    first_t0 = t0 if first_t0 is None else first_t0
                    log["wread"].append((s, (t, j), t0, t1))
                    log["accw"].append((me, t * groups_per_tile[me] + grp[j], bool(f & TJ["FIRST"]) and i == 0, t0, t1, j))
                    if a_smem:
                        log["sr"].append(("dir" if f & TJ["A_DIR"] else "enc", t, t0, t1, j))
                    else:
                        c = jb.a_col - A0COL + 8 * kk          # 8 columns of one K step
                        exp = None
                        if (c // 128) == (wb & 1):               # the job's version fields refer to the buffer it reads
                            half = (c % 128) // 64
                            exp = t * P.ready[(c // 128) * 2 + half] + ((wb >> (4 if half else 1)) & 7)
                        log["ard"].append((c // 128, (c % 128) // 16, exp, t0, t1, j))
                done = last_end

                def commits(s=s, f=f):
                    w_empty[s].arrive()
                    if f & TJ["C_ENC"]:
                        enc_empty.arrive()
                    if f & TJ["C_DIR"]:
                        dir_empty.arrive()
                    if f & TJ["C_ACC0"]:
                        acc_full[0].arrive()
                    if f & TJ["C_ACC1"]:
                        acc_full[1].arrive()
                sim.at(done + rng.randint(20, 250), commits)
                if guard and (jb.ksteps & NEXT_OTHER):
                    prog[me]["v"] = cnt + 1
                    ws, prog[me]["w"] = prog[me]["w"], []
                    for w in ws:
                        sim.poll(w)
                cnt += 1

    def epilogue(warp):
        wg = warp >> 2
        par, mc = [0, 0], 0
        done_cnt = [0, 0, 0, 0]
        for t in range(tiles):
            for si, st in enumerate(P.steps):
                yield wait(acc_full[st.acc], par[st.acc], None)
                par[st.acc] ^= 1
                gid = t * groups_per_tile[st.acc] + sgrp[si]
                if st.mode == EPI_OUT:
                    if wg == 0:
                        t0 = sim.now
                        yield ("delay", rng.randint(40, 120))
                        log["accr"].append((st.acc, gid, warp, t0, sim.now, si))
                    continue
                if BWD and st.mode == EPI_MASK:
                    b = mc & 1
                    yield wait(m_full[b], mc >> 1, mc >> 1, f"gates of step {si}")
                    t0 = sim.now
                    yield ("delay", rng.randint(30, 90))
                    log["gread"].append((b, mc, t0, sim.now, warp))
                    m_empty[b].arrive(32)
                    mc += 1
                t0 = sim.now
                yield ("delay", rng.randint(120, 400))
                log["accr"].append((st.acc, gid, warp, t0, sim.now, si))
                if st.acc == 1:
                    acc1_empty.arrive(32)
                slot = st.out_buf * 2 + st.out_half
                ver = t * P.ready[slot] + done_cnt[slot]
                for gq in range(2):
                    yield ("delay", rng.randint(150, 700 if BWD else 450))
                    if not st.no_act:
                        t0 = sim.now
                        yield ("delay", rng.randint(30, 90))
                        log["ast"].append((st.out_buf, st.out_half * 4 + wg * 2 + gq, ver, warp, t0, sim.now, si))
                        if kSplit and st.out_half == 1 and gq == 0:
                            a_ready[slot].arrive(32)
                if not st.no_act:
                    (a_ready_b[st.out_buf] if (kSplit and st.out_half == 1) else a_ready[slot]).arrive(32)
                    done_cnt[slot] += 1
            for i in range(4):
                done_cnt[i] = 0
            tile_done.arrive(32)

    def input_warp(w):
        for t in range(tiles):
            yield ("delay", rng.randint(300, 2500))
            if t > 0:
                yield wait(enc_empty, t - 1, t - 1, "encoding block free")
            t0 = sim.now
            yield ("delay", rng.randint(40, 150))
            log["sw"].append(("enc", t, w, t0, sim.now))
            enc_full.arrive(32)
            if not BWD:
                yield ("delay", rng.randint(100, 600))
                if t > 0:
                    yield wait(dir_empty, t - 1, t - 1, "direction block free")
                t0 = sim.now
                yield ("delay", rng.randint(40, 150))
                log["sw"].append(("dir", t, w, t0, sim.now))
                dir_full.arrive(32)

    def gate_producer():
        mc = 0
        for t in range(tiles):
            for st in P.steps:
                if st.mode != EPI_MASK:
                    continue
                b, n = mc & 1, mc >> 1
                yield wait(m_empty[b], (n & 1) ^ 1, n - 1 if n > 0 else None, "gate buffer free")
                lat = rng.randint(400, 3000)
                t0 = sim.now

                def gland(b=b, mc=mc, t0=t0):
                    log["gfill"].append((b, mc, t0, sim.now))
                    m_full[b].arrive()
                sim.at(sim.now + lat, gland)
                yield ("delay", rng.randint(20, 60))
                mc += 1

    sim.spawn("producer", producer())
    sim.spawn("issuer0", issuer(0))
    sim.spawn("issuer1", issuer(1))
    for w in range(8):
        sim.spawn(f"epilogue{w}", epilogue(w))
    for w in range(4):
        sim.spawn(f"input{w}", input_warp(w))
    if BWD:
        sim.spawn("gates", gate_producer())
    sim.run()
    errs = list(sim.errors)
    errs += check_logs(log)
    return errs


def check_logs(log):
    """Operand hazards from the access logs (times are [start, end] of the access in the unit that performs it)."""
    errs = []

    def add(msg):
        if len(errs) < 20:
            errs.append("hazard: " + msg)

    # weight stages: an MMA reads the fill meant for it, and no fill is landing on the stage while it reads
    fills = {}
    for s, job, t0, t1 in log["fill"]:
        fills.setdefault(s, []).append((t1, t0, job))
    for s in fills:
        fills[s].sort()
    for s, job, t0, t1 in log["wread"]:
        landed = [f for f in fills.get(s, []) if f[0] <= t0]
        if not landed or landed[-1][2] != job:
            add(f"job {job} read weight stage {s} holding {landed[-1][2] if landed else None}")
        if any(f[1] < t1 and f[0] > t0 for f in fills.get(s, [])):
            add(f"job {job} read weight stage {s} while a fill was landing on it")
    # accumulators: group g of half h is written (first MMA overwrites) only after group g-1 was read out, and read only
    # after all of its MMAs have finished
    wend, wstart, rend, rstart = {}, {}, {}, {}
    for h, gid, first, t0, t1, j in log["accw"]:
        wend[(h, gid)] = max(wend.get((h, gid), 0), t1)
        wstart[(h, gid)] = min(wstart.get((h, gid), 1 << 60), t0)
    for h, gid, warp, t0, t1, si in log["accr"]:
        rend[(h, gid)] = max(rend.get((h, gid), 0), t1)
        rstart[(h, gid)] = min(rstart.get((h, gid), 1 << 60), t0)
    for (h, gid), t0 in wstart.items():
        if (h, gid - 1) in rend and t0 < rend[(h, gid - 1)]:
            add(f"accumulator {h}: group {gid} overwritten at {t0} before group {gid - 1} was read out ({rend[(h, gid - 1)]})")
        if (h, gid - 1) in wend and (h, gid - 1) not in rend:
            add(f"accumulator {h}: group {gid - 1} never read")
    for (h, gid), t0 in rstart.items():
        if (h, gid) not in wend or t0 < wend[(h, gid)]:
            add(f"accumulator {h}: group {gid} read at {t0} before its MMAs finished ({wend.get((h, gid))})")
    # activation buffers, per 16-column group: a read sees the version it expects (the latest store that completed before
    # it started), and no store overlaps a read in time
    st = {}
    for buf, cg, ver, warp, t0, t1, si in log["ast"]:
        st.setdefault((buf, cg), []).append((t0, t1, ver))
    for buf, cg, exp, t0, t1, j in log["ard"]:
        ss = st.get((buf, cg), [])
        if any(a < t1 and b > t0 for a, b, _ in ss):
            add(f"job {j} read A{buf} columns {16 * cg}.. while they were being stored")
        if exp is not None:
            byver = {}
            for a, b, v in ss:
                byver.setdefault(v, []).append(b)
            full = [v for v, ends in byver.items() if len(ends) == 4 and max(ends) <= t0]   # all four warps of the group
            started = [v for a, b, v in ss if a < t0]
            if not full or max(full) != exp or (started and max(started) != exp):
                add(f"job {j} read A{buf} columns {16 * cg}.. expecting version {exp}, found complete {max(full) if full else None}"
                    f" / started {max(started) if started else None}")
    # gate staging (dgrad): a warp reads the gates loaded for its step, and no load is landing on the buffer meanwhile
    gf = {}
    for b, mc, t0, t1 in log["gfill"]:
        gf.setdefault(b, []).append((t1, t0, mc))
    for b in gf:
        gf[b].sort()
    for b, mc, t0, t1, warp in log["gread"]:
        landed = [f for f in gf.get(b, []) if f[0] <= t0]
        if not landed or landed[-1][2] != mc:
            add(f"epilogue warp {warp} read gate buffer {b} holding step {landed[-1][2] if landed else None}, wanted {mc}")
        if any(f[1] < t1 and f[0] > t0 for f in gf.get(b, [])):
            add(f"epilogue warp {warp} read gate buffer {b} while a load was landing on it")
    # encoding / direction blocks in shared memory
    sw = {}
    for name, t, w, t0, t1 in log["sw"]:
        sw.setdefault(name, []).append((t0, t1, t))
    for name, t, t0, t1, j in log["sr"]:
        ws = sw.get(name, [])
        if any(a < t1 and b > t0 for a, b, _ in ws):
            add(f"job {j} read the {name} block while it was being written")
        done = [v for a, b, v in ws if b <= t0]
        if sum(1 for v in done if v == t) != 4 or (done and max(done) != t):
            add(f"job {j} read the {name} block of tile {max(done) if done else None}, wanted {t}")
    return errs
