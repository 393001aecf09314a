// bf16 NeRF MLP with the activations resident in TENSOR MEMORY (tcgen05.mma with the A operand in TMEM).
// Same network, same roles and same drop-in entry points as mlp_tc.cu (run_nerf_helpers.py:75-129 fused with
// run.py:2317 / run_nerf_helpers.py:23-53); see mlp_ts_layout.h for why this layout exists and for the job tables.
//
//   warp 0      weight producer: 32 KB slabs ([128 out x 128 in] bf16, K-major 128B-swizzle) L2 -> smem ring, bulk TMA
//   warps 1, 3  MMA issuers (warp 1: jobs into accumulator half 0, warp 3: half 1): per job 8 x tcgen05.mma (M=128,
//               N=128, K=16), A from TMEM (or the shared-memory encoding / direction block), B = weight slab
//   warp 2      TMEM allocator; dgrad: gate producer (bulk-loads the H-stash blocks that gate the next step)
//   warps 4-7   per-tile input block (forward: points + positional + direction encoding, backward: padded g_raw) -> smem
//   warps 8-15  epilogue: accumulator half -> registers (tcgen05.ld) -> +bias/ReLU (or ReLU gate) -> bf16 ->
//               tcgen05.st into the other A buffer (the next layer's operand) [+ training stash: 16-byte stores straight
//               from the registers, a warp covers 512 contiguous bytes per chunk (layout: stash_chunk_off)]
// Hand-overs (mbarriers): acc_full[h] (commit of a half's last MMA -> epilogue), a_ready[buf][h] (epilogue -> issuers:
// input half h of the next layer is in TMEM), acc1_empty (half 1 has been read out, ~600 cycles before a_ready), and in
// the dgrad program a_ready_b[buf] (second 32-channel instalment of input half 1).
// The two issuers share the weight ring; ts_wait_progress() keeps either from testing a stage's full barrier before
// the other's previous fill of that stage has landed (mbarrier waits see phase parity only).
#include <cmath>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "mlp_ts_layout.h"
#include "tc_ptx.cuh"

namespace gbn {

// (a0, a1) + (b0, b1) as one packed fp32x2 instruction (sm_100 FADD2); a* are raw accumulator words
__device__ __forceinline__ void add_f32x2(uint32_t a0, uint32_t a1, float b0, float b1, float& r0, float& r1) {
  unsigned long long x, y;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(a0), "r"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(x) : "l"(x), "l"(y));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(x));
}


using namespace tc;

__constant__ TsJob c_tsjobs[2][kTsMaxJobs];
__constant__ TsStep c_tssteps[2][kTsMaxSteps];
__constant__ TsPackJob c_tspack[2][kTsMaxJobs];

// One issuer's view of one of its jobs, fully resolved on the host (ts_build_issue): see ts_issue_loop.
enum : uint16_t { TI_A_SMEM = 1, TI_A_DIR = 2, TI_FIRST = 4, TI_PREV_OTHER = 8, TI_NEXT_OTHER = 16, TI_SIGNAL_ORDER = 32 };
struct TsIssue {
  uint32_t w[5];        // waits before the job: barrier offset from L::bars (bits 0-15) | odd(completions per tile) << 16 |
                        // (index of the awaited completion within the tile & 1) << 17 | valid << 31
  uint32_t wsplit;      // dgrad: barrier of the second instalment of input half 1, waited for in mid-issue (same encoding)
  uint32_t idesc;       // tcgen05 instruction descriptor
  uint16_t d_col, a_col;
  uint16_t jidx;        // index of the job in the tile's job list: ring slot = tile * njobs + jidx
  uint16_t flags;       // TI_*
  uint8_t ksteps, nkb, n16, pad;
  uint16_t commit[3];   // barriers (offsets from L::bars) committed after the job besides the stage's w_empty; 0xffff = none
  uint16_t pad2;
};
static_assert(sizeof(TsIssue) == 48, "TsIssue layout");
__constant__ TsIssue c_tsissue[2][2][kTsMaxJobs];

#ifdef GBN_TS_DIAG
constexpr bool kDiag = true;
#else
constexpr bool kDiag = false;
#endif
constexpr int kTsThreads = 512;
constexpr int kTsStageBytes = 32768;

// Shared memory: input block | weight ring | [backward: gate staging, 2 x 2 stash blocks] | [forward: bias table] |
// barriers.  (The training stash leaves through registers -> global since late round 2; the 32 KB staging area it used to
// take is gone.)
template <bool BWD>
struct TsSmemT {
  static constexpr int NST = 4;   // (the dgrad program had 3 while a 32 KB stash staging area shared its shared memory)
  static constexpr uint32_t enc = 0;
  static constexpr uint32_t dir = enc + kBlkBytes;                       // forward only: view-direction encoding block
  static constexpr uint32_t ring = dir + (BWD ? 0 : kBlkBytes);
  static constexpr uint32_t mstage = ring + NST * kTsStageBytes;
  static constexpr uint32_t bias = mstage + (BWD ? 4 * kBlkBytes : 0);
  static constexpr uint32_t bars = bias + (BWD ? 0 : kTsBiasFloats * 4);
  static constexpr uint32_t w_full = bars;
  static constexpr uint32_t w_empty = w_full + 8 * NST;
  static constexpr uint32_t acc_full = w_empty + 8 * NST;           // [2]
  static constexpr uint32_t a_ready = acc_full + 16;                // [buf][half]
  static constexpr uint32_t enc_full = a_ready + 32;
  static constexpr uint32_t enc_empty = enc_full + 8;
  static constexpr uint32_t tile_done = enc_empty + 8;
  static constexpr uint32_t order = tile_done + 8;
  static constexpr uint32_t m_full = order + 8;                     // [2]
  static constexpr uint32_t m_empty = m_full + 16;                  // [2]
  static constexpr uint32_t dir_full = m_empty + 16;
  static constexpr uint32_t dir_empty = dir_full + 8;
  static constexpr uint32_t acc1_empty = dir_empty + 8;
  static constexpr uint32_t a_ready_b = acc1_empty + 8;             // [buf]: second 32-channel groups of input half 1
  static constexpr uint32_t tmem_ptr = a_ready_b + 16;
  static constexpr uint32_t abort_flag = tmem_ptr + 4;
  static constexpr uint32_t prog = abort_flag + 4;                  // [2]: ring slots whose fill issuer 0 / 1 has seen land
  static constexpr uint32_t total = prog + 8;
  static constexpr uint32_t alloc = total + 1024;
};
static_assert(TsSmemT<false>::alloc <= 232448 && TsSmemT<true>::alloc <= 232448, "shared memory budget");

struct TsArgs {
  const uint8_t* packed;
  const float* ro; const float* rd; const float* z; const float* pts; const float* emb;
  const float* view_bias;
  const float* vd;                     // forward with stash: view directions [R,3] (pitch `stride`)
  float* raw;                          // forward: out [P,4]; backward: gradient in
  uint8_t* stash_h; uint8_t* stash_g;
  int* err;
  unsigned long long* trace;           // optional clock64 trace of CTA 0 (gbn_mlp_set_trace)
  int trace_tile;
  int64_t stride, P;
  int S, njobs, nsteps;
  int nissue[2];           // records of issuer 0 / 1 in c_tsissue[program]
  int ready_per_tile[4];
  int order_per_tile;
  int empty1_per_tile;
  int no_split;            // GBNERF_TS_SPLIT=0: K-high jobs wait for both instalments of input half 1 up front
  int gate_direct;         // diagnostic (GBNERF_TS_GATE_DIRECT=1): dgrad epilogue reads its ReLU gates from global memory
  unsigned long long* dbg; // diagnostic: gate-check event log (dgrad; the buffer of gbn_mlp_set_trace), see the epilogue
  unsigned chaos_roles;    // diagnostic: which roles get chaos delays (bit 0 weight producer, 1 issuers, 2 gate producer,
                           // 3 input warps, 4 epilogue before acc_full, 5 epilogue before the hand-over arrival)
  unsigned fix;            // diagnostic (GBNERF_TS_FIX): candidate fixes, bit 0: fence.proxy.async before the m_empty
                           // arrival, bit 1: consume the gate registers before that arrival
  unsigned chaos;          // diagnostic (GBNERF_TS_CHAOS=seed): pseudo-random nanosleeps at every hand-over point, to shake
                           // timing-dependent holes of the barrier protocol out on real hardware (tools/dgrad_hunt.py)
};

// Post-mortem of a watchdog expiry, in host-mapped memory so that it survives a dead context
// (gbn_watchdog_report): [0] code of the first wait that expired, [1] CTA, [2] thread, [3] 1,
// [8 + 2i, 9 + 2i] raw mbarrier word i counted backwards from the last barrier of the layout,
// [128 + warp] the last wait that warp left through the abort path (after the abort every wait returns at once, so
// this is where the warp was when the kernel wound down, not where it was parked; first-write-wins is a TODO).
__device__ unsigned int* g_ts_wd_host = nullptr;
static unsigned int* g_ts_wd_host_ptr = nullptr;
constexpr int kWdWords = 256, kWdBarriers = 32;

__device__ __noinline__ void ts_wd_dump(uint32_t abort_addr, int code) {
  unsigned int* h = g_ts_wd_host;
  if (h == nullptr) return;
  volatile unsigned int* v = h;
  v[0] = (unsigned)code; v[1] = blockIdx.x; v[2] = threadIdx.x; v[3] = 1u;
  for (int i = 0; i < kWdBarriers; ++i) {
    unsigned long long w;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(w) : "r"(abort_addr - 4u - 8u * (uint32_t)(i + 1)));
    v[8 + 2 * i] = (unsigned)w; v[9 + 2 * i] = (unsigned)(w >> 32);
  }
  __threadfence_system();
}
__device__ __noinline__ void ts_wd_note(int code) {
  unsigned int* h = g_ts_wd_host;
  if (h == nullptr || (threadIdx.x & 31) != 0) return;
  volatile unsigned int* v = h;
  v[128 + (threadIdx.x >> 5)] = (unsigned)code;
  v[160 + (threadIdx.x >> 5)] = blockIdx.x;
  __threadfence_system();
}

__device__ __forceinline__ void ts_wait(uint32_t bar, uint32_t parity, uint32_t abort_addr, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    uint32_t ab;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(ab) : "r"(abort_addr));
    if (ab) { ts_wd_note(code); return; }
    if (clock64() - t0 > kWatchdogCycles) {
      if (atomicCAS(err, 0, code) == 0) ts_wd_dump(abort_addr, code);
      asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(abort_addr), "r"(1u));
      ts_wd_note(code);
      return;
    }
  }
}

// The two issuing warps share one weight ring, and an mbarrier only exposes the parity of its phase: a warp that waits
// for fill #m of a stage without having seen fill #m-1 (the other warp's job) would pass on the wrong phase if that
// earlier fill is still in flight (a cold weight image: L2 misses of > 1 us).  Each issuer therefore publishes how many
// ring slots it has seen land, and a warp about to wait on a stage last used by the other one first waits for that
// count to cover the previous fill.  (Seen as an intermittent launch failure of the dgrad program, whose 3-stage ring
// puts half 1's K-high job on the stage of half 0's K-low job of the same layer)
__device__ __forceinline__ void ts_wait_progress(uint32_t addr, uint32_t need, uint32_t abort_addr, int* err, int code) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  if (v < need) {
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
      if (v >= need) break;
      uint32_t ab;
      asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(ab) : "r"(abort_addr));
      if (ab) { ts_wd_note(code); return; }
      if (clock64() - t0 > kWatchdogCycles) {
        if (atomicCAS(err, 0, code) == 0) ts_wd_dump(abort_addr, code);
        asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(abort_addr), "r"(1u));
        ts_wd_note(code);
        return;
      }
    }
  }
}

// Chaos mode: a warp-uniform pseudo-random delay of 0 .. ~4 us (most calls: none) keyed on the seed, the CTA, the warp
// and a per-site counter.  Results must not depend on it.
__device__ __forceinline__ void ts_chaos(unsigned seed, unsigned site, bool role_on = true) {
#ifndef GBN_TS_DIAG
  (void)seed; (void)site; (void)role_on;
  return;      // diagnostics are compiled into libgbnerf_diag.so only (csrc/build.py --diag): the product kernels carry none
#else
  if (seed == 0u || !role_on) return;
  unsigned h = seed ^ (blockIdx.x * 0x9e3779b9u) ^ ((threadIdx.x >> 5) * 0x85ebca6bu) ^ (site * 0xc2b2ae35u);
  h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
  if ((h & 7u) == 0u) __nanosleep((h >> 8) & 0xfffu);
#endif
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&w)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]),
      "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void ts_st_global16(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ uint4 ld_smem16(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}


// =============================== MMA issuers ===================================================================
// Two warps: warp 1 issues every job that accumulates into acc0 (WHO = 0), warp 3 those into acc1 (WHO = 1).  The halves
// are independent accumulators, so no ordering is needed between the two instruction streams; what it buys is that
// one warp's barrier waits / bookkeeping overlap the other's MMAs.  Each loop is warp-uniform; one elected lane issues.
//
// Round 2: the issuing thread was the pace-setter of the whole kernel (in-kernel trace: ~1400 cycles from one job to
// the next for 512 cycles of tensor work - the TsJob record was decoded field by field, every wait re-derived its
// barrier and parity from per-tile counters, jobs of the other issuer were walked and skipped).  The host now resolves
// each issuer's own jobs into TsIssue records (ts_build_issue): final barrier offsets with a two-bit parity rule per
// wait, the instruction descriptor, the commit targets - and the loop below only loads a record and executes it.
// The protocol (which barriers, in which order of jobs) is exactly that of the TsJob table; tests/ts_protocol_model.py
// keeps checking that table.
template <bool BWD, int WHO>
__device__ __forceinline__ void ts_issue_loop(const TsArgs& a, uint32_t base, uint32_t abort_addr, uint32_t tmem, int my_tiles,
                                              int lane) {
  using L = TsSmemT<BWD>;
  constexpr int kTsStages = L::NST;
  constexpr int PROG = BWD ? 1 : 0;
  const TsIssue* recs = c_tsissue[PROG][WHO];
  const int nrec = a.nissue[WHO];
  const uint64_t adesc_enc = smem_desc_sw128(base + L::enc);
  const uint64_t adesc_dir = smem_desc_sw128(base + L::dir);
  const uint64_t ring_desc0 = smem_desc_sw128(base + L::ring);
  const uint32_t bars = base + L::bars;
  for (int t = 0; t < my_tiles; ++t) {
    const uint32_t cnt0 = (uint32_t)t * (uint32_t)a.njobs;
    for (int i = 0; i < nrec; ++i) {
      const TsIssue rc = recs[i];
      const uint32_t cnt = cnt0 + rc.jidx;
      unsigned long long* tr = (a.trace && blockIdx.x == 0 && t == a.trace_tile && lane == 0) ? a.trace : nullptr;
      if (tr) tr[4 * rc.jidx] = clock64();
      ts_chaos(a.chaos, 2u * cnt, a.chaos_roles & 2u);
      // The ring guard first (it must precede any test of the stage's full barrier, and up here its shared-memory poll is
      // off the path between the operands' hand-over and the first MMA), then the operands, then the weight stage.
      // Tried in round 2 and slower, on one box against this order (1,059 TFLOP/s): probing all of a job's barriers at
      // once (903), a blocking wait for the weight stage before the operands (1,007), a non-blocking probe of it before
      // the operands (1,041), polling with test_wait instead of try_wait (no difference) - every extra probe of an
      // mbarrier is a ~130-cycle round trip for the issuing warp, whether or not the phase is complete.
      const uint32_t s = cnt % kTsStages, par = (cnt / kTsStages) & 1;
      if ((rc.flags & TI_PREV_OTHER) && cnt >= (uint32_t)kTsStages)
        ts_wait_progress(base + L::prog + (WHO ? 0u : 4u), cnt - kTsStages + 1, abort_addr, a.err, 0x26000000 | rc.jidx);
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const uint32_t w = rc.w[k];
        if (w & 0x80000000u)   // parity of completion t * per_tile + idx: ((t & odd(per_tile)) ^ idx) & 1
          ts_wait(bars + (w & 0xffffu), (((uint32_t)t & (w >> 16)) ^ (w >> 17)) & 1u, abort_addr, a.err,
                  0x20000000 | (k << 20) | rc.jidx);
      }
      ts_chaos(a.chaos, 2u * cnt + 1u, a.chaos_roles & 2u);
      if (tr) tr[4 * rc.jidx + 1] = clock64();
      ts_wait(base + L::w_full + 8 * s, par, abort_addr, a.err, 0x22000000 | rc.jidx);
      if (tr) tr[4 * rc.jidx + 2] = clock64();
      tc_fence_after_sync();
      const uint64_t bd0 = ring_desc0 + (uint64_t)(s * (kTsStageBytes >> 4));
      const uint64_t bd1 = bd0 + (uint64_t)((uint32_t)rc.n16 * 128u);            // second K-block image: N rows x 128 B on
      const uint32_t idesc = rc.idesc;
      const uint32_t d = tmem + rc.d_col;
      const uint32_t a_t = tmem + rc.a_col;
      const uint32_t first = (rc.flags & TI_FIRST) ? 0u : 1u;
      const bool split = (rc.wsplit & 0x80000000u) != 0u;
      if (split) {   // K-high job of the dgrad program: the four MMAs whose K ranges arrived first, then the rest
        if (elect_one()) {
          umma_bf16_ts(d, a_t, bd0, idesc, first);
          umma_bf16_ts(d, a_t + 8, bd0 + 2, idesc, 1u);
          umma_bf16_ts(d, a_t + 32, bd1, idesc, 1u);
          umma_bf16_ts(d, a_t + 40, bd1 + 2, idesc, 1u);
        }
        __syncwarp();
        ts_chaos(a.chaos, 0x40000000u + cnt, a.chaos_roles & 2u);
        ts_wait(bars + (rc.wsplit & 0xffffu), (((uint32_t)t & (rc.wsplit >> 16)) ^ (rc.wsplit >> 17)) & 1u, abort_addr, a.err,
                0x21e00000 | rc.jidx);
        tc_fence_after_sync();
      }
      if (elect_one()) {
        if (split) {
          umma_bf16_ts(d, a_t + 16, bd0 + 4, idesc, 1u);
          umma_bf16_ts(d, a_t + 24, bd0 + 6, idesc, 1u);
          umma_bf16_ts(d, a_t + 48, bd1 + 4, idesc, 1u);
          umma_bf16_ts(d, a_t + 56, bd1 + 6, idesc, 1u);
        } else if (!(rc.flags & TI_A_SMEM) && rc.nkb == 2) {          // the common job: 8 back-to-back MMAs, A from TMEM
          umma_bf16_ts(d, a_t, bd0, idesc, first);
          umma_bf16_ts(d, a_t + 8, bd0 + 2, idesc, 1u);
          umma_bf16_ts(d, a_t + 16, bd0 + 4, idesc, 1u);
          umma_bf16_ts(d, a_t + 24, bd0 + 6, idesc, 1u);
          umma_bf16_ts(d, a_t + 32, bd1, idesc, 1u);
          umma_bf16_ts(d, a_t + 40, bd1 + 2, idesc, 1u);
          umma_bf16_ts(d, a_t + 48, bd1 + 4, idesc, 1u);
          umma_bf16_ts(d, a_t + 56, bd1 + 6, idesc, 1u);
        } else if (rc.flags & TI_A_SMEM) {           // encoding / direction block (4 / 2 steps), padded g_raw block (1 step)
          const uint64_t adesc = (rc.flags & TI_A_DIR) ? adesc_dir : adesc_enc;
          umma_bf16(d, adesc, bd0, idesc, first);
          if (rc.ksteps >= 2) umma_bf16(d, adesc + 2, bd0 + 2, idesc, 1u);
          if (rc.ksteps == 4) {
            umma_bf16(d, adesc + 4, bd0 + 4, idesc, 1u);
            umma_bf16(d, adesc + 6, bd0 + 6, idesc, 1u);
          }
        } else {                                     // single K-block from TMEM (not used by the current plans)
          umma_bf16_ts(d, a_t, bd0, idesc, first);
          umma_bf16_ts(d, a_t + 8, bd0 + 2, idesc, 1u);
          umma_bf16_ts(d, a_t + 16, bd0 + 4, idesc, 1u);
          umma_bf16_ts(d, a_t + 24, bd0 + 6, idesc, 1u);
        }
        umma_commit(base + L::w_empty + 8 * s);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (rc.commit[k] != 0xffffu) umma_commit(bars + rc.commit[k]);
        if (rc.flags & TI_SIGNAL_ORDER) mbar_arrive(base + L::order);
      }
      __syncwarp();
      // this stage's next job is the other issuer's: tell it that the fill just consumed has landed (after the MMAs
      // have been issued, off the critical path; shared memory is coherent within the CTA and both sides use volatile
      // accesses in program order after / before their mbarrier tests)
      if ((rc.flags & TI_NEXT_OTHER) && lane == 0)
        asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(base + L::prog + (WHO ? 4u : 0u)), "r"(cnt + 1) : "memory");
      if (tr) tr[4 * rc.jidx + 3] = clock64();
    }
  }
}

template <bool BWD>
__global__ void __launch_bounds__(kTsThreads, 1) nerf_mlp_ts_kernel(const TsArgs a) {
  using L = TsSmemT<BWD>;
  constexpr int kTsStages = L::NST;
  constexpr int PROG = BWD ? 1 : 0;
  constexpr bool kSplit = BWD;            // two-instalment hand-over of input half 1 (dgrad program only)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (a.P + kTileRows - 1) / kTileRows;
  const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t abort_addr = base + L::abort_flag;
  const TsJob* jobs = c_tsjobs[PROG];

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTsStages; ++i) { mbar_init(base + L::w_full + 8 * i, 1); mbar_init(base + L::w_empty + 8 * i, 1); }
    mbar_init(base + L::acc_full, 1); mbar_init(base + L::acc_full + 8, 1);
    for (int i = 0; i < 4; ++i) mbar_init(base + L::a_ready + 8 * i, 256);
    mbar_init(base + L::enc_full, 128);
    mbar_init(base + L::enc_empty, 2);   // one commit from each of the two MMA warps
    mbar_init(base + L::tile_done, 256);
    mbar_init(base + L::order, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(base + L::m_full + 8 * i, 1); mbar_init(base + L::m_empty + 8 * i, 256); }
    mbar_init(base + L::dir_full, 128);
    mbar_init(base + L::dir_empty, 1);
    mbar_init(base + L::acc1_empty, 256);
    mbar_init(base + L::a_ready_b, 256); mbar_init(base + L::a_ready_b + 8, 256);
    *reinterpret_cast<volatile uint32_t*>(gen + L::abort_flag) = 0;
    *reinterpret_cast<volatile uint32_t*>(gen + L::prog) = 0;
    *reinterpret_cast<volatile uint32_t*>(gen + L::prog + 4) = 0;
    mbar_init_fence();
  }
  if (warp == 2) tmem_alloc(base + L::tmem_ptr, kTmemCols);
  if constexpr (!BWD) {
    const float* gb = reinterpret_cast<const float*>(a.packed + reinterpret_cast<const uint32_t*>(a.packed)[2]);
    float* sb = reinterpret_cast<float*>(gen + L::bias);
    for (int i = threadIdx.x; i < kTsBiasFloats; i += blockDim.x) sb[i] = __ldg(gb + i);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + L::tmem_ptr);

  if (warp == 0) {
    // =============================== weight producer ========================================================
    uint32_t cnt = 0;
    for (int t = 0; t < my_tiles; ++t)
      for (int j = 0; j < a.njobs; ++j) {
        const uint32_t s = cnt % kTsStages, par = (cnt / kTsStages) & 1;
        ts_chaos(a.chaos, cnt, a.chaos_roles & 1u);
        ts_wait(base + L::w_empty + 8 * s, par ^ 1, abort_addr, a.err, 0x10000000 | j);
        const uint32_t bytes = (uint32_t)jobs[j].w_bytes16 * 16;
        const uint8_t* src = a.packed + jobs[j].w_off;
        if (elect_one()) {
          mbar_expect_tx(base + L::w_full + 8 * s, bytes);
          tma_bulk_g2s(base + L::ring + s * kTsStageBytes, src, bytes, base + L::w_full + 8 * s);
        }
        __syncwarp();
        ++cnt;
      }
  } else if (warp == 1) {
    ts_issue_loop<BWD, 0>(a, base, abort_addr, tmem, my_tiles, lane);
  } else if (warp == 3) {
    ts_issue_loop<BWD, 1>(a, base, abort_addr, tmem, my_tiles, lane);
  } else if (warp == 2) {
    // =============================== backward: gate producer ====================================================
    // streams the two H-stash block images whose sign gates the next epilogue step into a double-buffered staging
    // area (one 32 KB bulk copy per step), so the epilogue reads them from shared memory instead of issuing
    // row-strided 16-byte global loads
    if constexpr (BWD) {
      uint32_t mc = 0;
      for (int t = 0; t < (kDiag && a.gate_direct ? 0 : my_tiles); ++t) {
        const int64_t tile = blockIdx.x + (int64_t)t * gridDim.x;
        for (int si = 0; si < a.nsteps; ++si) {
          const TsStep st = c_tssteps[PROG][si];
          if (st.mode != EPI_MASK) continue;
          const uint32_t b = mc & 1, par = (mc >> 1) & 1;
          ts_chaos(a.chaos, mc, a.chaos_roles & 4u);
          ts_wait(base + L::m_empty + 8 * b, par ^ 1, abort_addr, a.err, 0x60000000 | si);
          if (kDiag && (a.fix & 8u)) fence_proxy_async_smem();
#ifdef GBN_TS_DIAG
          if (a.dbg != nullptr && (a.fix & 16u) && blockIdx.x < 16 && mc < 128 && lane == 0)
            a.dbg[32768 + blockIdx.x * 128 + mc] = clock64();          // clock at which fill `mc` is issued
#endif
          if (elect_one()) {
            mbar_expect_tx(base + L::m_full + 8 * b, 2 * kBlkBytes);
            tma_bulk_g2s(base + L::mstage + b * 2 * kBlkBytes,
                         a.stash_h + (size_t)tile * kStashTileBytes + (size_t)st.mask_blk * kBlkBytes, 2 * kBlkBytes,
                         base + L::m_full + 8 * b);
          }
          __syncwarp();
          ++mc;
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // =============================== per-tile input block: thread == row =======================================
    const int row = threadIdx.x - 128;
    const uint32_t row_off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t tile = blockIdx.x + (int64_t)t * gridDim.x;
      const int64_t p = tile * kTileRows + row;
      uint32_t w[32];   // 64 bf16 channels of this row
      if constexpr (BWD) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < a.P) g = ld_stream4(reinterpret_cast<const float4*>(a.raw) + p);
#pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = 0u;
        w[0] = pack_bf16(g.x, g.y);
        w[1] = pack_bf16(g.z, g.w);
      } else {
        float e[64];
        if (p < a.P) {
          if (a.emb != nullptr) {
#pragma unroll
            for (int i = 0; i < 63; ++i) e[i] = __ldg(a.emb + p * GBN_EMB_CH + i);
          } else {
            float x[3];
            if (a.pts != nullptr) {
#pragma unroll
              for (int i = 0; i < 3; ++i) x[i] = __ldg(a.pts + p * 3 + i);
            } else {
              const int64_t r = p / a.S;
              const float zz = __ldg(a.z + p);
#pragma unroll
              for (int i = 0; i < 3; ++i)
                x[i] = __fadd_rn(__ldg(a.ro + r * a.stride + i), __fmul_rn(__ldg(a.rd + r * a.stride + i), zz));
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              float sc[20];
              posenc_axis<10>(x[i], sc);
              e[i] = x[i];
#pragma unroll
              for (int k = 0; k < 10; ++k) { e[3 + 6 * k + i] = sc[2 * k]; e[6 + 6 * k + i] = sc[2 * k + 1]; }
            }
          }
          e[63] = 0.f;
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) e[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = pack_bf16(e[2 * i], e[2 * i + 1]);
      }
      ts_chaos(a.chaos, (unsigned)t, a.chaos_roles & 8u);
      if (t > 0) ts_wait(base + L::enc_empty, (t - 1) & 1, abort_addr, a.err, 0x30000000 | t);
      uint8_t* gblk = nullptr;
      if constexpr (BWD) gblk = a.stash_g + (size_t)tile * kStashTileBytes + (size_t)kGRaw * kBlkBytes;
      else if (a.stash_h != nullptr) gblk = a.stash_h + (size_t)tile * kStashTileBytes + (size_t)kHEnc * kBlkBytes;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t off = ((uint32_t)(c ^ (row & 7)) << 4);
        st_smem16(base + L::enc + row_off + off, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        if (gblk != nullptr)
          ts_st_global16(gblk + stash_chunk_off((uint32_t)row, (uint32_t)c), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(base + L::enc_full);
      if constexpr (!BWD) {
        // view-direction encoding (27 of 64 channels): A operand of the direction columns of views_linears.0 and,
        // in training, the B operand of their wgrad item
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) e[i] = 0.f;
        if (p < a.P) {
          if (a.emb != nullptr) {
#pragma unroll
            for (int i = 0; i < 27; ++i) e[i] = __ldg(a.emb + p * GBN_EMB_CH + GBN_PTS_CH + i);
          } else {
            const int64_t r = p / a.S;
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
              const float x = __ldg(a.vd + r * a.stride + ax);
              float sc[8];
              posenc_axis<4>(x, sc);
              e[ax] = x;
#pragma unroll
              for (int k = 0; k < 4; ++k) { e[3 + 6 * k + ax] = sc[2 * k]; e[6 + 6 * k + ax] = sc[2 * k + 1]; }
            }
          }
        }
        if (t > 0) ts_wait(base + L::dir_empty, (t - 1) & 1, abort_addr, a.err, 0x31000000 | t);
        uint8_t* db = a.stash_h != nullptr ? a.stash_h + (size_t)tile * kStashTileBytes + (size_t)kHDir * kBlkBytes : nullptr;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t q[4] = {0u, 0u, 0u, 0u};
          if (c < 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = pack_bf16(e[8 * c + 2 * i], e[8 * c + 2 * i + 1]);
          }
          const uint32_t off = ((uint32_t)(c ^ (row & 7)) << 4);
          st_smem16(base + L::dir + row_off + off, q[0], q[1], q[2], q[3]);
          if (db != nullptr) ts_st_global16(db + stash_chunk_off((uint32_t)row, (uint32_t)c), q[0], q[1], q[2], q[3]);
        }
        fence_proxy_async_smem();
        mbar_arrive(base + L::dir_full);
      }
    }
  } else if (warp >= 8) {
    // =============================== epilogue: thread == row, warpgroup == 64-channel slice of the half ========
    const int wg = (warp - 8) >> 2;
    const int row = ((warp & 3) << 5) | lane;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
    const uint32_t row_off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    const float* sbias = reinterpret_cast<const float*>(gen + L::bias);
    uint32_t accpar = 0, mc = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t tile = blockIdx.x + (int64_t)t * gridDim.x;
      const int64_t p = tile * kTileRows + row;
      float sigma_acc = 0.f;
      unsigned long long* tr = (a.trace && blockIdx.x == 0 && t == a.trace_tile && (threadIdx.x & 127) == 0)
                                   ? a.trace + 960 + wg * 128 : nullptr;
      for (int si = 0; si < a.nsteps; ++si) {
        const TsStep st = c_tssteps[PROG][si];
        if (tr) tr[si * 4] = clock64();
        ts_chaos(a.chaos, (unsigned)(t * 64 + si), a.chaos_roles & 16u);
        ts_wait(base + L::acc_full + 8 * st.acc, (accpar >> st.acc) & 1, abort_addr, a.err, 0x40000000 | (si << 8) | wg);
        accpar ^= 1u << st.acc;
        if (tr) tr[si * 4 + 1] = clock64();
        tc_fence_after_sync();
        if (st.mode == EPI_OUT) {
          if (wg == 0) {
            uint32_t c[4], sv;
            tmem_ld4(lane_addr + kTsAcc0 + kTsColRgb, c);
            tmem_ld1(lane_addr + kTsAcc0 + kTsColAlpha, sv);
            tmem_ld_wait();
            sigma_acc = __uint_as_float(sv);
            if (p < a.P) {
              float4 o;
              o.x = __uint_as_float(c[0]) + sbias[kBiasRgb + 0];
              o.y = __uint_as_float(c[1]) + sbias[kBiasRgb + 1];
              o.z = __uint_as_float(c[2]) + sbias[kBiasRgb + 2];
              o.w = sigma_acc + sbias[kBiasAlpha];
              st_stream4(reinterpret_cast<float4*>(a.raw) + p, o);
            }
          }
          continue;
        }
        const bool relu = (st.mode == EPI_BIAS_RELU);
        const uint32_t acc_col = (st.acc ? kTsAcc1 : kTsAcc0) + 64u * wg;
        const int ch0 = 128 * st.out_half + 64 * wg;            // first of this thread's 64 channels in the layer
        const uint32_t out_col = (st.out_buf ? kTsA1 : kTsA0) + 64u * st.out_half + 32u * wg;
        uint8_t* gout = nullptr;   // this warpgroup's 16 KB block image of the result in the stash
        if (st.out_blk != 0xff) {
          if constexpr (BWD) gout = a.stash_g + (size_t)tile * kStashTileBytes + (size_t)(st.out_blk + wg) * kBlkBytes;
          else if (a.stash_h != nullptr) gout = a.stash_h + (size_t)tile * kStashTileBytes + (size_t)(st.out_blk + wg) * kBlkBytes;
        }
        // both 32-channel groups in flight at once: two tcgen05.ld, one wait, then the arithmetic, two tcgen05.st
        uint4 hm[2][4];
        if constexpr (BWD) {
          if (kDiag && st.mode == EPI_MASK && a.gate_direct) {
            const uint8_t* hg = a.stash_h + (size_t)tile * kStashTileBytes + (size_t)(st.mask_blk + wg) * kBlkBytes;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              hm[c >> 2][c & 3] = __ldg(reinterpret_cast<const uint4*>(hg + stash_chunk_off((uint32_t)row, (uint32_t)c)));
          } else if (st.mode == EPI_MASK) {
            const uint32_t b = mc & 1;
            ts_wait(base + L::m_full + 8 * b, (mc >> 1) & 1, abort_addr, a.err, 0x41000000 | (si << 8) | wg);
            const uint32_t hb = base + L::mstage + (b * 2 + wg) * kBlkBytes;   // byte image of the stash block
#pragma unroll
            for (int c = 0; c < 8; ++c) hm[c >> 2][c & 3] = ld_smem16(hb + stash_chunk_off((uint32_t)row, (uint32_t)c));
#ifdef GBN_TS_DIAG
            if (a.fix & 2u) {
#pragma unroll
              for (int c = 0; c < 8; ++c) asm volatile("" ::"r"(hm[c >> 2][c & 3].x), "r"(hm[c >> 2][c & 3].w) : "memory");
            }
            if (a.fix & 4u) __nanosleep(200);
            if (a.dbg != nullptr && (a.fix & 16u) && blockIdx.x < 16 && mc < 128 && lane == 0)
              a.dbg[65536 + (blockIdx.x * 128 + mc) * 8 + (warp - 8)] = clock64();   // clock of this warp's m_empty arrival
#endif
            // The staging buffer was written by the async proxy (bulk copy) and has just been read through the generic
            // proxy; the arrival below hands it back to the gate producer, whose next bulk copy overwrites it.  An
            // mbarrier hand-over alone does not order generic-proxy READS before later async-proxy WRITES: without this
            // proxy fence the next fill (two EPI_MASK steps ahead) was observed landing in the last chunks a warp read -
            // the "one 64-channel block of g_h7" nondeterminism of round 1 (reproduced at 0.3-23 % of launches
            // with tools/chaos_hunt*.sh, 0 of 10,500 with the fence; a plain delay before the arrival does not close it).
            if (!kDiag || !(a.fix & 32u)) fence_proxy_async_smem();
            mbar_arrive(base + L::m_empty + 8 * b);
#ifdef GBN_TS_DIAG
            if (a.dbg != nullptr && !(a.fix & 16u)) {
              // gate check: the same 128 bytes straight from the H stash; on a mismatch log where, when, and whether the
              // wrong chunk equals the gates of the NEXT fill of this staging buffer (two EPI_MASK steps ahead) or of the
              // PREVIOUS one.  Record: [tile, si | warp << 8 | lane << 16, mc, chunk mask | next << 8 | prev << 16, clock]
              const uint8_t* hg = a.stash_h + (size_t)tile * kStashTileBytes + (size_t)(st.mask_blk + wg) * kBlkBytes;
              unsigned bad = 0;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 g = __ldg(reinterpret_cast<const uint4*>(hg + stash_chunk_off((uint32_t)row, (uint32_t)c))), h = hm[c >> 2][c & 3];
                if (g.x != h.x || g.y != h.y || g.z != h.z || g.w != h.w) bad |= 1u << c;
              }
              if (bad) {
                int seen = 0, nsi = si, blk_next = -1, blk_prev = -1;
                int64_t tile_next = tile, tile_prev = tile;
                for (int k = 0; k < 2 * a.nsteps && seen < 2; ++k) {      // second EPI_MASK step after this one
                  if (++nsi == a.nsteps) { nsi = 0; tile_next += gridDim.x; }
                  if (c_tssteps[PROG][nsi].mode == EPI_MASK) { ++seen; blk_next = c_tssteps[PROG][nsi].mask_blk; }
                }
                seen = 0; nsi = si;
                for (int k = 0; k < 2 * a.nsteps && seen < 2; ++k) {      // second EPI_MASK step before this one
                  if (--nsi < 0) { nsi = a.nsteps - 1; tile_prev -= gridDim.x; }
                  if (c_tssteps[PROG][nsi].mode == EPI_MASK) { ++seen; blk_prev = c_tssteps[PROG][nsi].mask_blk; }
                }
                unsigned eq_next = 0, eq_prev = 0;
                const int64_t ntiles_all = (a.P + kTileRows - 1) / kTileRows;
                for (int c = 0; c < 8; ++c) {
                  if (!((bad >> c) & 1u)) continue;
                  const uint4 h = hm[c >> 2][c & 3];
                  if (tile_next < ntiles_all) {
                    const uint4 g = __ldg(reinterpret_cast<const uint4*>(a.stash_h + (size_t)tile_next * kStashTileBytes +
                                                                       (size_t)(blk_next + wg) * kBlkBytes + stash_chunk_off((uint32_t)row, (uint32_t)c)));
                    if (g.x == h.x && g.y == h.y && g.z == h.z && g.w == h.w) eq_next |= 1u << c;
                  }
                  if (tile_prev >= 0) {
                    const uint4 g = __ldg(reinterpret_cast<const uint4*>(a.stash_h + (size_t)tile_prev * kStashTileBytes +
                                                                       (size_t)(blk_prev + wg) * kBlkBytes + stash_chunk_off((uint32_t)row, (uint32_t)c)));
                    if (g.x == h.x && g.y == h.y && g.z == h.z && g.w == h.w) eq_prev |= 1u << c;
                  }
                }
                const unsigned long long slot = atomicAdd(a.dbg, 1ull);
                if (slot < 2000) {
                  unsigned long long* r = a.dbg + 8 + slot * 8;
                  r[0] = (unsigned long long)tile; r[1] = (unsigned)si | ((unsigned)warp << 8) | ((unsigned)lane << 16);
                  r[2] = mc; r[3] = bad | (eq_next << 8) | (eq_prev << 16); r[4] = clock64();
                  unsigned long long w0, w1;
                  asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(w0) : "r"(base + L::m_full + 8 * b));
                  asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(w1) : "r"(base + L::m_empty + 8 * b));
                  r[5] = w0; r[6] = w1; r[7] = blockIdx.x;
                }
              }
            }
#endif
            ++mc;
          }
        }
        uint32_t v[2][32];
        tmem_ld32(lane_addr + acc_col, v[0]);
        tmem_ld32(lane_addr + acc_col + 32, v[1]);
        tmem_ld_wait();
        if (st.acc == 1) {   // acc1 is in registers: the next layer's half-1 MMAs may overwrite it
          tc_fence_before_sync();
          mbar_arrive(base + L::acc1_empty);
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float f[32];
          if constexpr (BWD) {
            if (st.mode == EPI_MASK) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint32_t hw[4] = {hm[g][c].x, hm[g][c].y, hm[g][c].z, hm[g][c].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  f[c * 8 + 2 * i] = ((hw[i] & 0x7fffu) != 0u) ? __uint_as_float(v[g][c * 8 + 2 * i]) : 0.f;
                  f[c * 8 + 2 * i + 1] = ((hw[i] & 0x7fff0000u) != 0u) ? __uint_as_float(v[g][c * 8 + 2 * i + 1]) : 0.f;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[g][i]);
            }
          } else {
            const float4* bp = reinterpret_cast<const float4*>(sbias + st.bias_off + ch0 + 32 * g);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {   // packed fp32x2 adds (FADD2): two channels per instruction
              const float4 bb = bp[i >> 2];
              add_f32x2(v[g][i], v[g][i + 1], bb.x, bb.y, f[i], f[i + 1]);
              add_f32x2(v[g][i + 2], v[g][i + 3], bb.z, bb.w, f[i + 2], f[i + 3]);
            }
          }
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = relu ? pack_bf16_relu(f[2 * i], f[2 * i + 1]) : pack_bf16(f[2 * i], f[2 * i + 1]);
          if (!st.no_act) {
            tmem_st16(lane_addr + out_col + 16 * g, w);
            if (kSplit && st.out_half == 1 && g == 0) {   // first instalment of input half 1 (see the issuer's TJ_WAIT_A1)
              tmem_st_wait();
              tc_fence_before_sync();
              mbar_arrive(base + L::a_ready + 8 * (st.out_buf * 2 + 1));
            }
          }
          if (gout != nullptr) {   // training stash: a warp's 32 consecutive points make every chunk 512 contiguous bytes
#pragma unroll
            for (int c = 0; c < 4; ++c)
              ts_st_global16(gout + stash_chunk_off((uint32_t)row, (uint32_t)(g * 4 + c)), w[4 * c], w[4 * c + 1], w[4 * c + 2],
                             w[4 * c + 3]);
          }
        }
        ts_chaos(a.chaos, 0x20000000u + (unsigned)(t * 64 + si), a.chaos_roles & 32u);
        if (!st.no_act) {
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive((kSplit && st.out_half == 1) ? base + L::a_ready_b + 8 * st.out_buf
                                                   : base + L::a_ready + 8 * (st.out_buf * 2 + st.out_half));
        }
        if (tr) tr[si * 4 + 2] = clock64();
      }
      tc_fence_before_sync();
      mbar_arrive(base + L::tile_done);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, kTmemCols);
}

// ---- weight pre-pack for the TS plans: slab = nkb K-block images of [rows x 128 B], K-major, 128B swizzle ------------
struct TsParamPtrs {
  const float* w[GBN_NUM_LINEAR];
  const float* b[GBN_NUM_LINEAR];
};
struct TsPackHeader {
  uint32_t magic, plan, off_bias, off_wdir, off_bdir, total_bytes, njobs, pad;
};

__global__ void __launch_bounds__(256) ts_prepack_kernel(TsParamPtrs pp, uint8_t* __restrict__ out, int njobs, TsPackHeader hdr,
                                                         int prog) {
  if ((int)blockIdx.x < njobs) {
    const TsPackJob q = c_tspack[prog][blockIdx.x];
    const float* W = pp.w[q.layer];
    const int kcols = 64 * q.nkb;
    for (int i = threadIdx.x; i < q.rows * kcols; i += blockDim.x) {
      const int n = q.transpose ? i % q.rows : i / kcols;
      const int k = q.transpose ? i / q.rows : i - n * kcols;
      const int ks = k - (int)q.koff;
      float v = 0.f;
      if (n < q.rows_valid && ks >= 0 && ks < q.cols_valid)
        v = q.transpose ? __ldg(W + (size_t)(q.row0 + ks) * q.ld + q.col0 + n) : __ldg(W + (size_t)(q.row0 + n) * q.ld + q.col0 + ks);
      const int kb = k >> 6, kk = k & 63;
      uint8_t* dst = out + q.w_off + (size_t)kb * q.rows * 128 + sw128_offset((uint32_t)n, (uint32_t)(kk >> 3)) + (kk & 7) * 2;
      *reinterpret_cast<uint16_t*>(dst) = (uint16_t)(pack_bf16(v, 0.f) & 0xffff);
    }
    return;
  }
  if (threadIdx.x == 0) *reinterpret_cast<TsPackHeader*>(out) = hdr;
  float* bias = reinterpret_cast<float*>(out + hdr.off_bias);
  for (int i = threadIdx.x; i < kTsBiasFloats; i += blockDim.x) {
    float v = 0.f;
    if (i < kBiasFeat) v = pp.b[i >> 8][i & 255];
    else if (i < kBiasAlpha) v = pp.b[LIN_FEATURE][i - kBiasFeat];
    else if (i == kBiasAlpha) v = pp.b[LIN_ALPHA][0];
    else if (i >= kBiasRgb && i < kBiasRgb + 3) v = pp.b[LIN_RGB][i - kBiasRgb];
    else if (i >= kTsBiasViews) v = pp.b[LIN_VIEWS][i - kTsBiasViews];
    bias[i] = v;
  }
  float* wdir = reinterpret_cast<float*>(out + hdr.off_wdir);
  for (int i = threadIdx.x; i < 128 * 27; i += blockDim.x) {
    const int j = i / 27, c = i - j * 27;
    wdir[i] = pp.w[LIN_VIEWS][(size_t)j * 283 + 256 + c];
  }
  float* bdir = reinterpret_cast<float*>(out + hdr.off_bdir);
  for (int i = threadIdx.x; i < 128; i += blockDim.x) bdir[i] = pp.b[LIN_VIEWS][i];
}

// =========================================================================================================
// host side (called from the C ABI entry points in mlp_tc.cu / mlp_aux.cu)
// =========================================================================================================
static std::once_flag g_ts_once;
static TsPlan g_ts_plan[2];
static bool g_ts_init[64];
static std::mutex g_ts_mutex;

const TsPlan& ts_plan(int bwd) {
  std::call_once(g_ts_once, [] {
    const char* e = getenv("GBNERF_TS_STAGGER");
    const bool stagger = e && e[0] == '1';
    const char* ee = getenv("GBNERF_TS_EARLY");
    const bool early = !(ee && ee[0] == '0');
    const char* ko = getenv("GBNERF_TS_KHI_ORDER");
    const bool khi = ko && ko[0] == '1';   // measured slower (1030 vs 1052-1064 TFLOP/s): off unless asked for
    g_ts_plan[0] = make_ts_plan(kTsFwd, stagger, early, khi);
    // dgrad: early acc1 release as in the forward program (GBNERF_TS_BWD_EARLY=0 switches it off).  An intermittent
    // launch failure first seen with it was the ring-stage parity alias described at ts_wait_progress(), not this.
    const char* be = getenv("GBNERF_TS_BWD_EARLY");
    g_ts_plan[1] = make_ts_plan(kTsBwd, stagger, !(be && be[0] == '0'), khi);
    // debug: early plan, but the K-low jobs of half 1 whose bit is clear in GBNERF_TS_DBG_EARLY_MASK (hex; bit i = the
    // i-th such job of the tile; GBNERF_TS_BWD_EARLY=2 clears all) also wait for input half 1, i.e. do not overlap
    const char* em = getenv("GBNERF_TS_DBG_EARLY_MASK");
    unsigned long mask = em ? strtoul(em, nullptr, 16) : ~0ul;
    if (be && be[0] == '2') mask = 0;
    int ei = 0;
    for (auto& j : g_ts_plan[1].jobs)
      if (j.flags & TJ_WAIT_EMPTY1) {
        if (!((mask >> ei) & 1)) j.flags |= TJ_WAIT_A1;
        ++ei;
      }
    // test hook for the watchdog path (tests/test_gpu_watchdog.py, opt-in): forward job 3 waits for an issue-order signal
    // that nobody raises, so issuer 0 runs into the bounded wait (code 0x24000003) and the CTA aborts
    const char* hang = getenv("GBNERF_TS_DBG_HANG");
    if (hang && hang[0] == '1') g_ts_plan[0].jobs[3].flags |= TJ_WAIT_ORDER;
    // ring-stage ownership (see ts_wait_progress): mark the jobs whose stage was last / is next used by the other
    // issuer; the job sequence repeats every tile, so the neighbour `stages` slots away wraps around the table
    const char* ng = getenv("GBNERF_TS_DBG_NO_RING_GUARD");   // timing A/B only: without the guard the kernel can fail
    for (int pr = 0; pr < 2 && !(ng && ng[0] == '1'); ++pr) {
      std::vector<TsJob>& jobs = g_ts_plan[pr].jobs;
      const int n = (int)jobs.size(), nst = pr ? TsSmemT<true>::NST : TsSmemT<false>::NST;
      for (int j = 0; j < n; ++j) {
        const bool mine = jobs[j].d_col >= kTsAcc1;
        if ((jobs[((j - nst) % n + n) % n].d_col >= kTsAcc1) != mine) jobs[j].flags |= TJ_PREV_OTHER;
        if ((jobs[(j + nst) % n].d_col >= kTsAcc1) != mine) jobs[j].ksteps |= kTjNextOther;
      }
    }
  });
  return g_ts_plan[bwd ? 1 : 0];
}


// TsJob table -> per-issuer TsIssue records (see ts_issue_loop).  Barrier offsets are relative to L::bars.
static int g_ts_nissue[2][2];
template <bool BWD>
static void ts_build_issue(const TsPlan& p, std::vector<TsIssue> (&out)[2]) {
  using L = TsSmemT<BWD>;
  static const bool no_split = [] { const char* e = getenv("GBNERF_TS_SPLIT"); return e && e[0] == '0'; }();
  auto W = [](uint32_t bar_abs, int per_tile, int idx) {
    return (bar_abs - L::bars) | ((uint32_t)(per_tile & 1) << 16) | ((uint32_t)(idx & 1) << 17) | 0x80000000u;
  };
  for (size_t j = 0; j < p.jobs.size(); ++j) {
    const TsJob& jb = p.jobs[j];
    TsIssue r{};
    int nw = 0;
    auto add = [&](uint32_t w) { if (nw < 5) r.w[nw] = w; ++nw; };
    if (jb.flags & TJ_WAIT_ENC) add(W(L::enc_full, 1, 0));
    if (jb.flags & TJ_WAIT_DIR) add(W(L::dir_full, 1, 0));
    if (jb.flags & TJ_WAIT_TILE) add(W(L::tile_done, 1, 1));     // completion t - 1; tile 0 passes on the fresh barrier
    if (jb.flags & TJ_WAIT_A0) {
      const int b = (jb.wait_buf & 1) * 2;
      add(W(L::a_ready + 8 * b, p.ready_per_tile[b], (jb.wait_buf >> 1) & 7));
    }
    if (jb.flags & TJ_WAIT_A1) {
      const int b = (jb.wait_buf & 1) * 2 + 1;
      const int idx = (jb.wait_buf >> 4) & 7;
      add(W(L::a_ready + 8 * b, p.ready_per_tile[b], idx));
      if (BWD) {   // two-instalment hand-over of input half 1: second barrier up front, or in mid-issue for the common job
        const uint32_t w2 = W(L::a_ready_b + 8 * (jb.wait_buf & 1), p.ready_per_tile[b], idx);
        if ((jb.flags & TJ_A_SMEM) || jb.nkb != 2 || no_split) add(w2); else r.wsplit = w2;
      }
    }
    if (jb.flags & TJ_WAIT_EMPTY1) add(W(L::acc1_empty, p.empty1_per_tile, (jb.wait_buf >> 7) & 1));
    if (jb.flags & TJ_WAIT_ORDER) add(W(L::order, p.order_per_tile, (jb.ksteps >> 4) & 7));
    if (nw > 5) abort();   // a plan that needs more wait slots must widen TsIssue
    r.idesc = make_idesc(1, 128, (uint32_t)jb.n16 * 16);
    r.d_col = jb.d_col; r.a_col = jb.a_col; r.jidx = (uint16_t)j;
    r.flags = (uint16_t)(((jb.flags & TJ_A_SMEM) ? TI_A_SMEM : 0) | ((jb.flags & TJ_A_DIR) ? TI_A_DIR : 0) |
                         ((jb.flags & TJ_FIRST) ? TI_FIRST : 0) | ((jb.flags & TJ_PREV_OTHER) ? TI_PREV_OTHER : 0) |
                         ((jb.ksteps & kTjNextOther) ? TI_NEXT_OTHER : 0) | ((jb.flags & TJ_SIGNAL_ORDER) ? TI_SIGNAL_ORDER : 0));
    r.ksteps = jb.ksteps & 7; r.nkb = jb.nkb; r.n16 = jb.n16;
    int nc = 0;
    r.commit[0] = r.commit[1] = r.commit[2] = 0xffff;
    if (jb.flags & TJ_COMMIT_ENC) r.commit[nc++] = (uint16_t)(L::enc_empty - L::bars);
    if (jb.flags & TJ_COMMIT_DIR) r.commit[nc++] = (uint16_t)(L::dir_empty - L::bars);
    if (jb.flags & TJ_COMMIT_ACC0) r.commit[nc++] = (uint16_t)(L::acc_full - L::bars);
    if (jb.flags & TJ_COMMIT_ACC1) { if (nc >= 3) abort(); r.commit[nc++] = (uint16_t)(L::acc_full + 8 - L::bars); }
    out[jb.d_col >= kTsAcc1 ? 1 : 0].push_back(r);
  }
}

#include "mlp_t2.cuh"   // two tiles in flight per CTA: the inference kernel (same packed image)

static int ts_ensure_device(cudaStream_t stream) {
  int dev = 0;
  GBN_CUDA(cudaGetDevice(&dev));
  GBN_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_ts_mutex);
  if (g_ts_init[dev]) return GBN_OK;
  for (int pr = 0; pr < 2; ++pr) {
    const TsPlan& p = ts_plan(pr);
    GBN_REQUIRE((int)p.jobs.size() <= kTsMaxJobs && (int)p.steps.size() <= kTsMaxSteps, "TS table overflow");
    GBN_CUDA(cudaMemcpyToSymbol(c_tsjobs, p.jobs.data(), p.jobs.size() * sizeof(TsJob), pr * kTsMaxJobs * sizeof(TsJob), cudaMemcpyHostToDevice));
    GBN_CUDA(cudaMemcpyToSymbol(c_tssteps, p.steps.data(), p.steps.size() * sizeof(TsStep),
                                     pr * kTsMaxSteps * sizeof(TsStep), cudaMemcpyHostToDevice));
    GBN_CUDA(cudaMemcpyToSymbol(c_tspack, p.pack.data(), p.pack.size() * sizeof(TsPackJob),
                                     pr * kTsMaxJobs * sizeof(TsPackJob), cudaMemcpyHostToDevice));
  }
  for (int pr = 0; pr < 2; ++pr) {
    std::vector<TsIssue> iss[2];
    if (pr == 0) ts_build_issue<false>(ts_plan(0), iss); else ts_build_issue<true>(ts_plan(1), iss);
    for (int who = 0; who < 2; ++who) {
      GBN_REQUIRE((int)iss[who].size() <= kTsMaxJobs, "TS issue table overflow");
      g_ts_nissue[pr][who] = (int)iss[who].size();
      GBN_CUDA(cudaMemcpyToSymbol(c_tsissue, iss[who].data(), iss[who].size() * sizeof(TsIssue),
                                  (size_t)(pr * 2 + who) * kTsMaxJobs * sizeof(TsIssue), cudaMemcpyHostToDevice));
    }
  }
  if (g_ts_wd_host_ptr == nullptr) {   // watchdog post-mortem record (zero-copy host memory, one per process)
    void* hp = nullptr;
    if (cudaHostAlloc(&hp, kWdWords * sizeof(unsigned int), cudaHostAllocMapped) == cudaSuccess) {
      memset(hp, 0, kWdWords * sizeof(unsigned int));
      g_ts_wd_host_ptr = static_cast<unsigned int*>(hp);
    } else {
      (void)cudaGetLastError();
    }
  }
  if (g_ts_wd_host_ptr != nullptr) {
    unsigned int* dp = nullptr;
    GBN_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dp), g_ts_wd_host_ptr, 0));
    GBN_CUDA(cudaMemcpyToSymbol(g_ts_wd_host, &dp, sizeof(dp), 0, cudaMemcpyHostToDevice));
  }
  GBN_CUDA(cudaFuncSetAttribute(nerf_mlp_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TsSmemT<false>::alloc));
  GBN_CUDA(cudaFuncSetAttribute(nerf_mlp_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TsSmemT<true>::alloc));
  { const int rc2 = t2_upload(); if (rc2 != GBN_OK) return rc2; }
  g_ts_init[dev] = true;
  return GBN_OK;
}

// GBNERF_TS_CHAOS=<seed>: every launch gets a different delay pattern (seed advanced per launch); 0 / unset = off
static unsigned ts_chaos_seed() {
  static const unsigned base = [] { const char* e = getenv("GBNERF_TS_CHAOS"); return e ? (unsigned)strtoul(e, nullptr, 0) : 0u; }();
  static unsigned n = 0;
  return base ? base + 0x9e3779b9u * (++n) : 0u;
}

size_t ts_packed_bytes(int bwd) { return ts_plan(bwd).total_bytes; }

int ts_prepack(const void* const* params, void* packed, int bwd, cudaStream_t st) {
  int rc = ts_ensure_device(st);
  if (rc != GBN_OK) return rc;
  TsParamPtrs pp;
  for (int i = 0; i < GBN_NUM_LINEAR; ++i) {
    pp.w[i] = static_cast<const float*>(params[2 * i]);
    pp.b[i] = static_cast<const float*>(params[2 * i + 1]);
  }
  const TsPlan& p = ts_plan(bwd);
  TsPackHeader hdr{0x4e425473u, (uint32_t)p.id, p.off_bias, p.off_wdir, p.off_bdir, p.total_bytes, (uint32_t)p.jobs.size(), 0};
  const int njobs = (int)p.jobs.size();
  ts_prepack_kernel<<<njobs + 1, 256, 0, st>>>(pp, static_cast<uint8_t*>(packed), njobs, hdr, bwd ? 1 : 0);
  return check_launch("ts_prepack_kernel");
}

// ---- optimizer step fused with the re-pack ---------------------------------------------------
// torch.optim.Adam(betas, eps; no weight decay, no amsgrad) on the 24 nn.Linear tensors of one network, in place,
// one thread per parameter; the same thread then drops the new value (bf16) at its position(s) in the forward and the
// transposed (dgrad) weight images by walking the pack-job table backwards, so no separate re-pack pass runs after
// an optimizer step.  Padding bytes of the images are never touched (they were zeroed by ts_prepack_kernel).
struct TsAdamArgs {
  float* p[2 * GBN_NUM_LINEAR];
  const float* g[2 * GBN_NUM_LINEAR];
  float* m[2 * GBN_NUM_LINEAR];
  float* v[2 * GBN_NUM_LINEAR];
  uint32_t start[2 * GBN_NUM_LINEAR + 1];
  uint16_t ld[GBN_NUM_LINEAR];
  uint8_t* pack[2];
  uint32_t njobs[2];
  uint32_t off_bias[2], off_wdir[2], off_bdir[2];
  float step_size, w1, b2, w2, eps, bc2_sqrt;
  const float* dev_scalars;   // non-NULL: {step_size, bc2_sqrt} are read from device memory (adam_tick_kernel), so that a
                              // captured CUDA graph can be replayed step after step
};

__device__ __forceinline__ void ts_scatter_weight(const TsAdamArgs& a, int layer, int n, int k, float val) {
  const uint16_t bits = (uint16_t)(pack_bf16(val, 0.f) & 0xffff);
#pragma unroll
  for (int prog = 0; prog < 2; ++prog) {
    uint8_t* out = a.pack[prog];
    if (!out) continue;
    for (int j = 0; j < (int)a.njobs[prog]; ++j) {
      const TsPackJob q = c_tspack[prog][j];
      if (q.layer != layer) continue;
      const int nl = q.transpose ? k - (int)q.col0 : n - (int)q.row0;
      const int ks = q.transpose ? n - (int)q.row0 : k - (int)q.col0;
      if (nl < 0 || nl >= (int)q.rows_valid || ks < 0 || ks >= (int)q.cols_valid) continue;
      const int kpos = ks + (int)q.koff, kb = kpos >> 6, kk = kpos & 63;
      uint8_t* dst = out + q.w_off + (size_t)kb * q.rows * 128 + sw128_offset((uint32_t)nl, (uint32_t)(kk >> 3)) + (kk & 7) * 2;
      *reinterpret_cast<uint16_t*>(dst) = bits;
    }
    if (layer == LIN_VIEWS && k >= 256) reinterpret_cast<float*>(out + a.off_wdir[prog])[n * 27 + (k - 256)] = val;
  }
}

__device__ __forceinline__ void ts_scatter_bias(const TsAdamArgs& a, int layer, int i, float val) {
  int at;
  if (layer < 8) at = 256 * layer + i;
  else if (layer == LIN_FEATURE) at = kBiasFeat + i;
  else if (layer == LIN_ALPHA) at = kBiasAlpha;
  else if (layer == LIN_RGB) at = kBiasRgb + i;
  else at = kTsBiasViews + i;
#pragma unroll
  for (int prog = 0; prog < 2; ++prog) {
    uint8_t* out = a.pack[prog];
    if (!out) continue;
    reinterpret_cast<float*>(out + a.off_bias[prog])[at] = val;
    if (layer == LIN_VIEWS) reinterpret_cast<float*>(out + a.off_bdir[prog])[i] = val;
  }
}

// One thread: step count += 1 and the two step-dependent scalars of torch.optim.Adam, in double as torch takes them on
// the host (bias_correction1/2 = 1 - beta^step; step_size = lr / bias_correction1).  state[0] = step (double),
// scalars = {step_size, sqrt(bias_correction2)} (float).
__global__ void adam_tick_kernel(double* state, const float* lr, double beta1, double beta2, float* scalars) {
  const double step = state[0] + 1.0;
  state[0] = step;
  scalars[0] = (float)((double)lr[0] / (1.0 - pow(beta1, step)));
  scalars[1] = (float)sqrt(1.0 - pow(beta2, step));
}

__global__ void __launch_bounds__(256) adam_repack_kernel(const __grid_constant__ TsAdamArgs a) {
  const uint32_t total = a.start[2 * GBN_NUM_LINEAR];
  const float step_size = a.dev_scalars ? __ldg(a.dev_scalars) : a.step_size;
  const float bc2_sqrt = a.dev_scalars ? __ldg(a.dev_scalars + 1) : a.bc2_sqrt;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    int lo = 0, hi = 2 * GBN_NUM_LINEAR;   // tensor ti with start[ti] <= e < start[ti+1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (a.start[mid] <= e) lo = mid; else hi = mid;
    }
    const int ti = lo;
    const uint32_t i = e - a.start[ti];
    const float g = __ldg(a.g[ti] + i);
    float m = a.m[ti][i], v = a.v[ti][i], p = a.p[ti][i];
    m = m + a.w1 * (g - m);                       // exp_avg.lerp_(grad, 1 - beta1)
    v = v * a.b2;                                 // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    v = v + a.w2 * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + a.eps;
    p = p + (-step_size) * (m / denom);         // param.addcdiv_(exp_avg, denom, value = -lr / bias_correction1)
    a.m[ti][i] = m; a.v[ti][i] = v; a.p[ti][i] = p;
    const int layer = ti >> 1;
    if (ti & 1) ts_scatter_bias(a, layer, (int)i, p);
    else {
      const int ld = a.ld[layer];
      const int n = (int)(i / (uint32_t)ld);
      ts_scatter_weight(a, layer, n, (int)i - n * ld, p);
    }
  }
}

int ts_adam_tick(double* state, const float* lr, double beta1, double beta2, float* scalars, cudaStream_t st) {
  adam_tick_kernel<<<1, 1, 0, st>>>(state, lr, beta1, beta2, scalars);
  return check_launch("adam_tick_kernel");
}

int ts_adam_repack(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq, double lr,
                   double beta1, double beta2, double eps, int64_t step, void* packed_fwd, void* packed_bwd, cudaStream_t st,
                   const float* dev_scalars) {
  int rc = ts_ensure_device(st);
  if (rc != GBN_OK) return rc;
  static const int kIn[GBN_NUM_LINEAR] = {63, 256, 256, 256, 256, 319, 256, 256, 256, 256, 283, 128};
  static const int kOut[GBN_NUM_LINEAR] = {256, 256, 256, 256, 256, 256, 256, 256, 256, 1, 128, 3};
  TsAdamArgs a{};
  uint32_t off = 0;
  for (int l = 0; l < GBN_NUM_LINEAR; ++l) {
    a.ld[l] = (uint16_t)kIn[l];
    for (int b = 0; b < 2; ++b) {
      const int ti = 2 * l + b;
      a.p[ti] = static_cast<float*>(params[ti]);
      a.g[ti] = static_cast<const float*>(grads[ti]);
      a.m[ti] = static_cast<float*>(exp_avg[ti]);
      a.v[ti] = static_cast<float*>(exp_avg_sq[ti]);
      a.start[ti] = off;
      off += (uint32_t)(b ? kOut[l] : kOut[l] * kIn[l]);
    }
  }
  a.start[2 * GBN_NUM_LINEAR] = off;   // == GBN_PARAM_COUNT
  const TsPlan& pf = ts_plan(0);
  const TsPlan& pb = ts_plan(1);
  a.pack[0] = static_cast<uint8_t*>(packed_fwd);
  a.pack[1] = static_cast<uint8_t*>(packed_bwd);
  a.njobs[0] = (uint32_t)pf.pack.size();
  a.njobs[1] = (uint32_t)pb.pack.size();
  a.off_bias[0] = pf.off_bias; a.off_wdir[0] = pf.off_wdir; a.off_bdir[0] = pf.off_bdir;
  a.off_bias[1] = pb.off_bias; a.off_wdir[1] = pb.off_wdir; a.off_bdir[1] = pb.off_bdir;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  a.step_size = (float)(lr / bc1);
  a.w1 = (float)(1.0 - beta1); a.b2 = (float)beta2; a.w2 = (float)(1.0 - beta2);
  a.eps = (float)eps; a.bc2_sqrt = (float)sqrt(bc2);
  a.dev_scalars = dev_scalars;
  adam_repack_kernel<<<(off + 255) / 256, 256, 0, st>>>(a);
  return check_launch("adam_repack_kernel");
}

void mlp_get_trace(unsigned long long** buf, int* tile);   // mlp_tc.cu
int launch_view_bias_raw(const float* wdir, const float* bdir, const float* viewdirs, int64_t stride, const float* emb,
                         int64_t n, float* out, cudaStream_t stream);  // mlp_aux.cu

int ts_forward(const void* packed, const float* ro, const float* rd, const float* vd, int64_t stride, const float* z,
               const float* pts, const float* emb, int64_t R, int S, float* raw, void* workspace, void* stash,
               cudaStream_t stream) {
  int rc = ts_ensure_device(stream);
  if (rc != GBN_OK) return rc;
  const TsPlan& p = ts_plan(0);
  int* err = reinterpret_cast<int*>(workspace);
  GBN_CUDA(cudaMemsetAsync(err, 0, 256, stream));
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed);
  {
    unsigned long long* trace = nullptr; int trace_tile = 0;
    mlp_get_trace(&trace, &trace_tile);
    // inference (no stash): two tiles in flight per CTA.  gbn_mlp_set_trace(buf, tile): tile = 0x40000000 | i traces pair
    // iteration i of this kernel; any other tile (a tile index, or -1 for "no trace") keeps the launch on the one-tile
    // kernel, which is how tools/t2_check.py runs both in one process
    const bool t2_trace = trace != nullptr && trace_tile >= 0 && (trace_tile & 0x40000000);
    if (stash == nullptr && (trace == nullptr || t2_trace) && t2_enabled())
      return t2_forward(packed, ro, rd, vd, stride, z, pts, emb, R, S, raw, err, stream, trace, trace_tile & 0xffff);
    // training forward (writes the H stash): the same kernel with the stash stores compiled in
    if (stash != nullptr && trace == nullptr && t2_stash_enabled())
      return t2_forward(packed, ro, rd, vd, stride, z, pts, emb, R, S, raw, err, stream, nullptr, 0, stash);
  }
  TsArgs a{};
  a.packed = pk; a.ro = ro; a.rd = rd; a.z = z; a.pts = pts; a.emb = emb; a.vd = vd; a.raw = raw;
  a.stash_h = static_cast<uint8_t*>(stash); a.err = err; a.stride = stride; a.P = R * S; a.S = S;
  a.njobs = (int)p.jobs.size(); a.nsteps = (int)p.steps.size();
  a.nissue[0] = g_ts_nissue[0][0]; a.nissue[1] = g_ts_nissue[0][1];
  for (int i = 0; i < 4; ++i) a.ready_per_tile[i] = p.ready_per_tile[i];
  a.order_per_tile = p.order_per_tile;
  a.empty1_per_tile = p.empty1_per_tile;
  { static const bool ns = [] { const char* e = getenv("GBNERF_TS_SPLIT"); return e && e[0] == '0'; }(); a.no_split = ns; }
  a.chaos = ts_chaos_seed();
  a.chaos_roles = 0xffu;
  mlp_get_trace(&a.trace, &a.trace_tile);
  const int64_t ntiles = (a.P + kTileRows - 1) / kTileRows;
  const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  nerf_mlp_ts_kernel<false><<<grid, kTsThreads, TsSmemT<false>::alloc, stream>>>(a);
  return check_launch("nerf_mlp_ts_kernel");
}

int ts_backward_data(const void* packed_bwd, const float* g_raw, int64_t P, const void* stash_h, void* stash_g, void* workspace,
                     cudaStream_t stream) {
  int rc = ts_ensure_device(stream);
  if (rc != GBN_OK) return rc;
  const TsPlan& p = ts_plan(1);
  int* err = reinterpret_cast<int*>(workspace);
  GBN_CUDA(cudaMemsetAsync(err, 0, 256, stream));
  TsArgs a{};
  a.packed = reinterpret_cast<const uint8_t*>(packed_bwd);
  a.raw = const_cast<float*>(g_raw);
  a.stash_h = static_cast<uint8_t*>(const_cast<void*>(stash_h));
  a.stash_g = static_cast<uint8_t*>(stash_g);
  a.err = err; a.P = P; a.S = 1;
  a.njobs = (int)p.jobs.size(); a.nsteps = (int)p.steps.size();
  a.nissue[0] = g_ts_nissue[1][0]; a.nissue[1] = g_ts_nissue[1][1];
  for (int i = 0; i < 4; ++i) a.ready_per_tile[i] = p.ready_per_tile[i];
  a.order_per_tile = p.order_per_tile;
  a.empty1_per_tile = p.empty1_per_tile;
  { static const bool ns = [] { const char* e = getenv("GBNERF_TS_SPLIT"); return e && e[0] == '0'; }(); a.no_split = ns; }
  { static const bool gd = [] { const char* e = getenv("GBNERF_TS_GATE_DIRECT"); return e && e[0] == '1'; }(); a.gate_direct = gd; }
  a.chaos = ts_chaos_seed();
  { static const unsigned r = [] { const char* e = getenv("GBNERF_TS_CHAOS_ROLES"); return e ? (unsigned)strtoul(e, nullptr, 0) : 0xffu; }(); a.chaos_roles = r; }
  { static const unsigned f = [] { const char* e = getenv("GBNERF_TS_FIX"); return e ? (unsigned)strtoul(e, nullptr, 0) : 0u; }(); a.fix = f; }
  { int tt = 0; mlp_get_trace(&a.dbg, &tt); }
  const int64_t ntiles = (P + kTileRows - 1) / kTileRows;
  const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  nerf_mlp_ts_kernel<true><<<grid, kTsThreads, TsSmemT<true>::alloc, stream>>>(a);
  return check_launch("nerf_mlp_ts_kernel<bwd>");
}

// host only: the job / step tables the kernels run (for the protocol model of tests/test_ts_protocol_cpu.py)
int ts_debug_plan(int bwd, void* jobs, int max_jobs, void* steps, int max_steps, int* meta) {
  if (bwd == 2) {   // the two-tiles-in-flight inference kernel (mlp_t2.cuh): 16-byte T2Job / 8-byte T2Step records
    if (g_t2.jobs.empty()) g_t2 = t2_build(ts_plan(0));
    const int nj = (int)g_t2.jobs.size(), ns = (int)g_t2.steps.size();
    if (!g_t2.ok || nj > max_jobs || ns > max_steps) return -1;
    memcpy(jobs, g_t2.jobs.data(), nj * sizeof(T2Job));
    memcpy(steps, g_t2.steps.data(), ns * sizeof(T2Step));
    for (int i = 0; i < 10; ++i) meta[i] = 0;
    meta[0] = nj; meta[1] = ns; meta[2] = T2Smem::NST; meta[3] = t2_mode();
    meta[4] = (int)g_t2.off_alpha[0]; meta[5] = (int)g_t2.off_alpha[1]; meta[6] = (int)g_t2.off_rgb;
    meta[7] = t2_enabled() ? 1 : 0;
    return 0;
  }
  const TsPlan& p = ts_plan(bwd);
  const int nj = (int)p.jobs.size(), ns = (int)p.steps.size();
  if (nj > max_jobs || ns > max_steps) return -1;
  memcpy(jobs, p.jobs.data(), nj * sizeof(TsJob));
  memcpy(steps, p.steps.data(), ns * sizeof(TsStep));
  meta[0] = nj; meta[1] = ns;
  for (int i = 0; i < 4; ++i) meta[2 + i] = p.ready_per_tile[i];
  meta[6] = p.order_per_tile; meta[7] = p.empty1_per_tile;
  meta[8] = bwd ? TsSmemT<true>::NST : TsSmemT<false>::NST;
  const char* e = getenv("GBNERF_TS_SPLIT");
  meta[9] = (bwd && !(e && e[0] == '0')) ? 1 : 0;      // two-instalment hand-over of input half 1 in use
  return 0;
}

int ts_watchdog_report(unsigned int* out, int words) {
  if (g_ts_wd_host_ptr == nullptr) return 0;
  const int n = words < kWdWords ? words : kWdWords;
  for (int i = 0; i < n; ++i) out[i] = reinterpret_cast<volatile unsigned int*>(g_ts_wd_host_ptr)[i];
  return n;
}

}  // namespace gbn
