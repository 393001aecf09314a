"""Hunt for the rare nondeterminism of the dgrad program: ONE stash-writing forward, then the dgrad
launch alone repeated LAUNCHES times (optionally from a cold L2), every G stash compared bit for bit with the first.
For each launch that differs it reports WHERE (tile, block, rows, channels) and WHAT the wrong block looks like
(gate-like 0 <-> value flips, or changed values; a least-squares fit of the change against the K-chunks of the
product says which operand chunk was stale).

usage: dgrad_hunt.py LAUNCHES [R S] [cold|warm|full]   env switches: see DESIGN 6 (GBNERF_TS_*)
`full` = the sequence of the pytest (tests/test_gpu_mlp_backward.py::test_training_kernels_repeat_...): every repeat
runs flush, a fresh zeroed stash, the stash-writing forward, flush, dgrad AND wgrad.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gbnerf_b200 as G  # noqa: E402
from gbnerf_b200 import _lib, ops  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402  (diagnostic tool: the oracle only supplies inputs)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
R = int(sys.argv[2]) if len(sys.argv) > 3 else 1024
S = int(sys.argv[3]) if len(sys.argv) > 3 else 128
MODE = sys.argv[4] if len(sys.argv) > 4 else "cold"
COLD, FULL = MODE != "warm", MODE == "full"
P = R * S
T = (P + 127) // 128
NB = 40
dev = torch.device("cuda:0")
torch.manual_seed(11)
net = G.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").to(dev)
net.load_state_dict(O.init_params(11))
rays = O.synthetic_rays(R, seed=3).to(dev)
z = O.stratified_z(rays[:, 6:7].cpu(), rays[:, 7:8].cpu(), S, True,
                   torch.rand(R, S, generator=torch.Generator().manual_seed(2))).to(dev)
g_raw = torch.randn(P, 4, generator=torch.Generator().manual_seed(4)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stash_h = ops._stash(P, dev).zero_()
raw, ws = ops.mlp_forward_raw(net.packed_weights(), "bf16", rays[:, 8:11], R, S, rays_o=rays[:, 0:3], rays_d=rays[:, 3:6],
                              z=z, stash=stash_h)
assert ops.mlp_error_code(ws) == 0
packed_bwd = net.packed_weights_bwd()
wsb = torch.zeros(512, device=dev, dtype=torch.uint8)
stream = torch.cuda.current_stream().cuda_stream


def dgrad(out):
    _lib.call("gbn_mlp_backward_data", packed_bwd.data_ptr(), g_raw.data_ptr(), P, stash_h.data_ptr(), out.data_ptr(),
              wsb.data_ptr(), stream)


def unswizzle(blk):
    """uint8 [16384] stash block -> float [128 points, 64 channels] (layout: csrc/mlp_layout.h, stash_chunk_off)"""
    return blk.view(torch.bfloat16).view(2, 8, 64, 8).permute(0, 2, 1, 3).reshape(128, 64).float()


def tile_rows(stash, tile, blk0, nblk):
    v = stash.view(T, NB, 16384)
    return torch.cat([unswizzle(v[tile, blk0 + i]) for i in range(nblk)], 1)


ref = ops._stash(P, dev).zero_()
dgrad(ref)
torch.cuda.synchronize()
print(f"# dgrad_hunt: {N} launches, R={R} S={S} ({T} tiles, {T / 148:.1f} per CTA), {'cold' if COLD else 'warm'} L2, "
      f"env {{{', '.join(f'{k}={v}' for k, v in sorted(os.environ.items()) if k.startswith('GBNERF'))}}}", flush=True)
print("first launch watchdog word", hex(ops.mlp_error_code(wsb)), flush=True)
refv = ref.view(T, NB, 16384)[:, :39]
out = ops._stash(P, dev).zero_()
outv = out.view(T, NB, 16384)[:, :39]
Wf = net.feature_linear.weight.detach().bfloat16().float()      # [256 feature, 256 h7]
wa = net.alpha_linear.weight.detach().bfloat16().float()[0]     # [256]


def forensics(tile, blocks):
    """blocks: differing block ids of this tile (sorted)"""
    top = [b for b in blocks if b >= 34 and b <= 37]
    early = [b for b in blocks if b < 6]
    msg = f"   tile {tile} = CTA {tile % 148} local tile {tile // 148}: blocks {blocks}"
    if early:
        msg += f"  (!! g_hv/g_feature blocks differ: {early})"
    print(msg)
    for b in (top or blocks[-1:]):
        good, bad = unswizzle(refv[tile, b]), unswizzle(outv[tile, b])
        d = good != bad
        rws, chs = d.any(1).nonzero().flatten(), d.any(0).nonzero().flatten()
        z2v = int(((good == 0) & (bad != 0)).sum()); v2z = int(((good != 0) & (bad == 0)).sum())
        vv = int(((good != 0) & (bad != 0) & d).sum())
        print(f"   block {b}: {int(d.sum())} of 8192 elements differ; rows {rws.min().item()}..{rws.max().item()} "
              f"({rws.numel()} rows), channels {chs.min().item()}..{chs.max().item()} ({chs.numel()}); "
              f"0->value {z2v}, value->0 {v2z}, value->value {vv}; max|good| {good.abs().max():.3e} max|bad| {bad.abs().max():.3e}")
        by_warp = [int(d[32 * w:32 * w + 32].sum()) for w in range(4)]
        by_chunk = [int(d[:, 8 * c:8 * c + 8].sum()) for c in range(8)]
        print(f"   differing elements by 32-row group {by_warp}, by 8-channel chunk {by_chunk}")
        if 34 <= b <= 37:
            # g_h7[:, n] = gate_h7 * (g_feature W_f[:, n] + g_sigma w_alpha[n]): which 16-wide K chunk changed?
            n0 = 64 * (b - 34)
            gf = tile_rows(ref, tile, 2, 4)                          # [128, 256] g_feature (bit-stable blocks)
            gs = g_raw[tile * 128:tile * 128 + 128, 3].bfloat16().float()
            gate = (tile_rows(stash_h, tile, 28, 4)[:, n0:n0 + 64] != 0).float()
            terms = [gf[:, 16 * i:16 * i + 16] @ Wf[16 * i:16 * i + 16, n0:n0 + 64] for i in range(16)]
            terms.append(gs[:, None] * wa[None, n0:n0 + 64])
            full = sum(terms)
            print(f"   check: |good - gate*full| max {(good - gate * full).abs().max():.3e} (bf16 rounding expected ~1e-2 rel)")
            m = d & (gate != 0)
            if int(m.sum()) > 20:
                A = torch.stack([t[m] for t in terms], 1)           # [n, 17]
                y = (bad - good)[m][:, None]
                coef = torch.linalg.lstsq(A, y).solution.flatten()
                print("   lstsq coefficients of (bad-good) on the 16 K-chunks + alpha term (-1 = chunk missing/stale):")
                print("   " + " ".join(f"{c:+.2f}" for c in coef.tolist()))
            gate_bad = (bad != 0).float()
            print(f"   gate mismatch (bad zero pattern vs h7 gate): {int((gate_bad != gate).sum())} elements; "
                  f"good vs gate {int(((good != 0).float() != gate).sum())}")
            for other in range(0, 38, 2):
                if other == 28 + 2 * ((b - 34) // 2):
                    continue
                og = (unswizzle(stash_h.view(T, NB, 16384)[tile, other + ((b - 34) & 1)]) != 0).float()
                mism = int((gate_bad != og).sum())
                if mism < 200:
                    print(f"   !! bad block's zero pattern matches H block {other + ((b - 34) & 1)} ({mism} mismatches)")


shapes = [tuple(t.shape) for t in net.param_list()]
h_ref = stash_h.clone() if FULL else None
# gate check of the diagnostic library (csrc/build.py --diag, GBNERF_LIB=...): the dgrad epilogue compares every gate chunk
# it read from the staging buffer with the H stash itself and logs mismatches
dbg = None
if os.environ.get("HUNT_GATECHECK") == "1":
    dbg = torch.zeros(65536 + 16 * 128 * 8, dtype=torch.int64, device=dev)
    _lib.call("gbn_mlp_set_trace", dbg.data_ptr(), 0)


def check_order():
    """GBNERF_TS_FIX bit 4 (16): clocks of every gate fill's issue and of every epilogue warp's m_empty arrival, CTAs 0-15:
    fill mc must be issued after all eight warps arrived for step mc - 2."""
    issue = dbg[32768:32768 + 16 * 128].view(16, 128).cpu()
    arrive = dbg[65536:].view(16, 128, 8).cpu()
    viol, worst = 0, 0
    for cta in range(16):
        for mc in range(2, 119):
            if issue[cta, mc] == 0 or (arrive[cta, mc - 2] == 0).any():
                continue
            late = int(arrive[cta, mc - 2].max() - issue[cta, mc])
            if late > 0:
                viol += 1
                worst = max(worst, late)
                if viol <= 5:
                    print(f"   ORDER VIOLATION cta {cta} fill {mc} issued {late} cycles BEFORE the last arrival of step {mc - 2}: "
                          f"arrivals - issue = {(arrive[cta, mc - 2] - issue[cta, mc]).tolist()}")
    return viol, worst


def dump_gatecheck():
    if dbg is None:
        return
    if int(os.environ.get("GBNERF_TS_FIX", "0"), 0) & 16:
        print("order check (last launch): violations, worst cycles =", check_order())
        return
    n = int(dbg[0].item())
    print(f"gate check: {n} mismatching thread-reads logged")
    rec = dbg[8:8 + 8 * min(n, 2000)].view(-1, 8).cpu().tolist()
    by = {}
    for tile, w1, mc, w3, clk, mfull, mempty, cta in rec:
        si, warp, lane = w1 & 0xff, (w1 >> 8) & 0xff, (w1 >> 16) & 0xff
        bad, eqn, eqp = w3 & 0xff, (w3 >> 8) & 0xff, (w3 >> 16) & 0xff
        k = (si, warp)
        d = by.setdefault(k, dict(n=0, bad=[0] * 8, next=0, prev=0, neither=0, clk=[], tiles=set(), lanes=set()))
        d["n"] += 1
        d["tiles"].add(tile); d["lanes"].add(lane); d["clk"].append(clk)
        for c in range(8):
            if (bad >> c) & 1:
                d["bad"][c] += 1
                if (eqn >> c) & 1: d["next"] += 1
                elif (eqp >> c) & 1: d["prev"] += 1
                else: d["neither"] += 1
    for (si, warp), d in sorted(by.items()):
        print(f"   step {si} warp {warp}: {d['n']} thread-reads in {len(d['tiles'])} tiles, lanes {min(d['lanes'])}..{max(d['lanes'])}, "
              f"bad chunks by read order {d['bad']}, wrong chunk == next fill {d['next']}, == previous fill {d['prev']}, neither {d['neither']}")
    for r in rec[:6]:
        print("   raw:", [hex(x & 0xffffffffffffffff) for x in r])


bad_launches = 0
t0 = time.time()
for it in range(N):
    if FULL:
        flush.zero_()
        stash_h = ops._stash(P, dev).zero_()
        raw2, ws = ops.mlp_forward_raw(net.packed_weights(), "bf16", rays[:, 8:11], R, S, rays_o=rays[:, 0:3],
                                       rays_d=rays[:, 3:6], z=z, stash=stash_h)
        flush.zero_()
        grads, wsb, out = ops.mlp_backward_raw(packed_bwd, g_raw, stash_h, rays[:, 8:11], R, S, shapes)
        outv = out.view(T, NB, 16384)[:, :39]
        if not torch.equal(raw2, raw) or not torch.equal(stash_h, h_ref):
            print(f"launch {it}: FORWARD differs (raw {torch.equal(raw2, raw)}, H stash {torch.equal(stash_h, h_ref)})", flush=True)
    else:
        if COLD:
            flush.zero_()
        dgrad(out)
    same = torch.equal(outv, refv)
    e = ops.mlp_error_code(wsb)
    if not same or e:
        bad_launches += 1
        blk = (outv != refv).any(-1).nonzero().tolist()
        tiles = sorted(set(t for t, _ in blk))
        print(f"launch {it}: {len(blk)} blocks differ in tiles {tiles[:8]}; watchdog {hex(e)}", flush=True)
        if bad_launches <= 6:
            for t in tiles[:3]:
                forensics(t, sorted(b for tt, b in blk if tt == t))
        sys.stdout.flush()
dump_gatecheck()
print(f"RESULT {bad_launches} of {N} launches differed ({time.time() - t0:.1f} s)", flush=True)
