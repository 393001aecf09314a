// Quarter-pipelined bf16 NeRF MLP with the activations resident in TENSOR MEMORY.
// Same network, roles and drop-in entry points as mlp_ts.cu (run_nerf_helpers.py:75-129 fused with run.py:2317 and
// run_nerf_helpers.py:23-53); see mlp_tq_layout.h for the job tables and the reason for the finer grain.
//
//   warp 0      weight producer: 32 KB slabs ([64 out x 256 in] bf16 as four K-block images) L2 -> smem ring, bulk TMA
//   warps 1, 3  MMA issuers, strictly alternating in plan order (warp 1: accumulator quarters 0/2, warp 3: 1/3): while
//               one issues its 16 x tcgen05.mma (M=128, N=64, K=16, A from TMEM) the other performs the barrier waits of
//               its next job; a shared-memory sequence counter hands the issue slot over
//   warp 2      TMEM allocator; backward: gate producer (bulk-loads the H-stash block gating each epilogue step)
//   warps 4-7   per-tile input blocks (forward: point encoding + view-direction encoding, backward: padded g_raw)
//   warps 8-11 / 12-15   epilogue warpgroups 0 / 1: drain accumulator quarters 0,2 / 1,3 (tcgen05.ld) -> bias/ReLU or
//               ReLU gate -> bf16 -> tcgen05.st into the other A buffer [+ stash block via smem staging + bulk store]
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "mlp_tq_layout.h"
#include "tc_ptx.cuh"

namespace gbn {

using namespace tc;

__constant__ TqJob c_tqjobs[2][kTqMaxJobs];
__constant__ TqStep c_tqsteps[2][kTqMaxSteps];
__constant__ TsPackJob c_tqpack[2][kTqMaxJobs];

constexpr int kTqThreads = 512;
constexpr int kTqStageBytes = 32768;

template <bool BWD>
struct TqSmem {
  static constexpr int NST = BWD ? 3 : 4;
  static constexpr uint32_t enc = 0;
  static constexpr uint32_t dir = enc + kBlkBytes;                       // forward only
  static constexpr uint32_t ring = dir + (BWD ? 0 : kBlkBytes);
  static constexpr uint32_t ostage = ring + NST * kTqStageBytes;          // one block image per epilogue warpgroup
  static constexpr uint32_t mstage = ostage + 2 * kBlkBytes;              // backward: [wg][2] block images
  static constexpr uint32_t bias = mstage + (BWD ? 4 * kBlkBytes : 0);
  static constexpr uint32_t bars = bias + (BWD ? 0 : kTqBiasFloats * 4);
  static constexpr uint32_t w_full = bars;
  static constexpr uint32_t w_empty = w_full + 8 * NST;
  static constexpr uint32_t acc_full = w_empty + 8 * NST;                 // [4]
  static constexpr uint32_t ready = acc_full + 32;                        // [buf*4 + quarter]
  static constexpr uint32_t enc_full = ready + 64;
  static constexpr uint32_t enc_empty = enc_full + 8;
  static constexpr uint32_t dir_full = enc_empty + 8;
  static constexpr uint32_t dir_empty = dir_full + 8;
  static constexpr uint32_t tile_done = dir_empty + 8;
  static constexpr uint32_t m_full = tile_done + 8;                       // [wg*2 + buf]
  static constexpr uint32_t m_empty = m_full + 32;
  static constexpr uint32_t issued = m_empty + 32;                        // sequence counter of the issue slot
  static constexpr uint32_t tmem_ptr = issued + 8;
  static constexpr uint32_t abort_flag = tmem_ptr + 4;
  static constexpr uint32_t total = abort_flag + 4;
  static constexpr uint32_t alloc = total + 1024;
};
static_assert(TqSmem<false>::alloc <= 232448 && TqSmem<true>::alloc <= 232448, "shared memory budget");

struct TqArgs {
  const uint8_t* packed;
  const float* ro; const float* rd; const float* z; const float* pts; const float* emb;
  const float* vd;                     // view directions [R,3] (pitch `stride`)
  float* raw;                          // forward: out [P,4]; backward: gradient in
  uint8_t* stash_h; uint8_t* stash_g;
  int* err;
  unsigned long long* trace;
  int trace_tile;
  int64_t stride, P;
  int S, njobs, nsteps;
  int ready_per_tile[8];
  int use_token;
};

__device__ __forceinline__ void tq_wait(uint32_t bar, uint32_t parity, uint32_t abort_addr, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    uint32_t ab;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(ab) : "r"(abort_addr));
    if (ab) return;
    if (clock64() - t0 > kWatchdogCycles) {
      atomicCAS(err, 0, code);
      asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(abort_addr), "r"(1u));
      return;
    }
  }
}

__device__ __forceinline__ void tq_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tq_tmem_st16(uint32_t taddr, const uint32_t* w) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]),
      "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])
      : "memory");
}
__device__ __forceinline__ void tq_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tq_st_global16(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tq_bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tq_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tq_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tq_wg_bar(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory"); }
__device__ __forceinline__ uint4 tq_ld_smem16(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint32_t tq_ld_vol(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

template <bool BWD>
__global__ void __launch_bounds__(kTqThreads, 1) nerf_mlp_tq_kernel(const TqArgs a) {
  using L = TqSmem<BWD>;
  constexpr int NST = L::NST;
  constexpr int PROG = BWD ? 1 : 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (a.P + kTileRows - 1) / kTileRows;
  const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t abort_addr = base + L::abort_flag;
  const TqJob* jobs = c_tqjobs[PROG];

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(base + L::w_full + 8 * i, 1); mbar_init(base + L::w_empty + 8 * i, 1); }
    for (int i = 0; i < 4; ++i) mbar_init(base + L::acc_full + 8 * i, 1);
    for (int i = 0; i < 8; ++i) mbar_init(base + L::ready + 8 * i, 128);
    mbar_init(base + L::enc_full, 128);
    mbar_init(base + L::enc_empty, 4);      // four trailing jobs per tile read the block last (one per quarter)
    mbar_init(base + L::dir_full, 128);
    mbar_init(base + L::dir_empty, 2);      // the two views quarters
    mbar_init(base + L::tile_done, 256);
    for (int i = 0; i < 4; ++i) { mbar_init(base + L::m_full + 8 * i, 1); mbar_init(base + L::m_empty + 8 * i, 128); }
    *reinterpret_cast<volatile uint32_t*>(gen + L::issued) = 0;
    *reinterpret_cast<volatile uint32_t*>(gen + L::abort_flag) = 0;
    mbar_init_fence();
  }
  if (warp == 2) tmem_alloc(base + L::tmem_ptr, kTmemCols);
  if constexpr (!BWD) {
    const float* gb = reinterpret_cast<const float*>(a.packed + reinterpret_cast<const uint32_t*>(a.packed)[2]);
    float* sb = reinterpret_cast<float*>(gen + L::bias);
    for (int i = threadIdx.x; i < kTqBiasFloats; i += blockDim.x) sb[i] = __ldg(gb + i);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + L::tmem_ptr);

  if (warp == 0) {
    // =============================== weight producer ============================================================
    uint32_t cnt = 0;
    for (int t = 0; t < my_tiles; ++t)
      for (int j = 0; j < a.njobs; ++j) {
        const uint32_t s = cnt % NST, par = (cnt / NST) & 1;
        tq_wait(base + L::w_empty + 8 * s, par ^ 1, abort_addr, a.err, 0x10000000 | j);
        const uint32_t bytes = (uint32_t)jobs[j].w_bytes16 * 16;
        const uint8_t* src = a.packed + jobs[j].w_off;
        if (elect_one()) {
          mbar_expect_tx(base + L::w_full + 8 * s, bytes);
          tma_bulk_g2s(base + L::ring + s * kTqStageBytes, src, bytes, base + L::w_full + 8 * s);
        }
        __syncwarp();
        ++cnt;
      }
  } else if (warp == 1 || warp == 3) {
    // =============================== MMA issuers ================================================================
    const int me = (warp == 3) ? 1 : 0;
    uint32_t cnt = 0;
    const uint64_t adesc_enc = smem_desc_sw128(base + L::enc);
    const uint64_t adesc_dir = smem_desc_sw128(base + L::dir);
    TqJob nxt = jobs[0];
    for (int t = 0; t < my_tiles; ++t)
      for (int j = 0; j < a.njobs; ++j, ++cnt) {
        const TqJob jb = nxt;
        nxt = jobs[j + 1 < a.njobs ? j + 1 : 0];
        if ((jb.acc & 1) != me) continue;
        unsigned long long* tr = (a.trace && blockIdx.x == 0 && t == a.trace_tile && lane == 0) ? a.trace : nullptr;
        if (tr) tr[4 * j] = clock64();
        if (jb.flags & (QJ_WAIT_ENC | QJ_WAIT_TILE | QJ_WAIT_DIR)) {
          if (jb.flags & QJ_WAIT_ENC) tq_wait(base + L::enc_full, t & 1, abort_addr, a.err, 0x20000000 | j);
          if (jb.flags & QJ_WAIT_DIR) tq_wait(base + L::dir_full, t & 1, abort_addr, a.err, 0x20800000 | j);
          if ((jb.flags & QJ_WAIT_TILE) && t > 0) tq_wait(base + L::tile_done, (t - 1) & 1, abort_addr, a.err, 0x23000000 | j);
        }
        const uint32_t s = cnt % NST, par = (cnt / NST) & 1;
        tq_wait(base + L::w_full + 8 * s, par, abort_addr, a.err, 0x22000000 | j);
        if (tr) tr[4 * j + 1] = clock64();
        const uint32_t N = (uint32_t)jb.n16 * 16;
        const uint64_t bd0 = smem_desc_sw128(base + L::ring + s * kTqStageBytes);
        const uint32_t idesc = make_idesc(1, 128, N);
        const uint32_t d = tmem + jb.d_col;
        const uint32_t a_t = tmem + jb.a_col;
        const bool a_smem = (jb.flags & (QJ_A_ENC | QJ_A_DIR)) != 0;
        const uint64_t adesc = (jb.flags & QJ_A_DIR) ? adesc_dir : adesc_enc;
        // ---- conditions this job has to observe before its first MMA (K-block 0) + the issue slot ---------------
        int wi = 0;
        while (wi < 6 && jb.waits[wi] != 0xff && (jb.waits[wi] & 3) == 0) {
          const uint32_t bar = (jb.waits[wi] >> 2) & 7;
          const uint32_t seq = (uint32_t)t * a.ready_per_tile[bar] + (jb.waits[wi] >> 5);
          tq_wait(base + L::ready + 8 * bar, seq & 1, abort_addr, a.err, 0x21000000 | (j << 8) | bar);
          ++wi;
        }
        // the issue slot: MMAs enter the tensor pipe in plan order, so accumulator quarters finish one after the other
        // and their drains line up behind them
        if (a.use_token && tq_ld_vol(base + L::issued) != cnt) {
          const long long t0 = clock64();
          while (tq_ld_vol(base + L::issued) != cnt) {
            if (tq_ld_vol(abort_addr)) break;
            if (clock64() - t0 > kWatchdogCycles) {
              atomicCAS(a.err, 0, 0x25000000 | j);
              asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(abort_addr), "r"(1u));
              break;
            }
          }
        }
        if (tr) tr[4 * j + 2] = clock64();
        tc_fence_after_sync();
        const uint32_t first = (jb.flags & QJ_FIRST) ? 0u : 1u;
        const bool more_waits = wi < 6 && jb.waits[wi] != 0xff;
        const uint32_t kbs = N * 8;   // descriptor units between K-block images (N rows x 128 B)
        if (a_smem) {
          if (elect_one()) {
            umma_bf16(d, adesc, bd0, idesc, first);
            if (jb.ksteps >= 2) umma_bf16(d, adesc + 2, bd0 + 2, idesc, 1u);
            if (jb.ksteps == 4) {
              umma_bf16(d, adesc + 4, bd0 + 4, idesc, 1u);
              umma_bf16(d, adesc + 6, bd0 + 6, idesc, 1u);
            }
          }
          __syncwarp();
        } else if (!more_waits) {
          // everything this job reads is already in TMEM: all K-blocks back to back
          if (elect_one()) {
            tq_umma_ts(d, a_t, bd0, idesc, first);
            tq_umma_ts(d, a_t + 8, bd0 + 2, idesc, 1u);
            tq_umma_ts(d, a_t + 16, bd0 + 4, idesc, 1u);
            tq_umma_ts(d, a_t + 24, bd0 + 6, idesc, 1u);
            if (jb.nkb >= 2) {
              const uint64_t b1 = bd0 + kbs;
              tq_umma_ts(d, a_t + 32, b1, idesc, 1u);
              tq_umma_ts(d, a_t + 40, b1 + 2, idesc, 1u);
              tq_umma_ts(d, a_t + 48, b1 + 4, idesc, 1u);
              tq_umma_ts(d, a_t + 56, b1 + 6, idesc, 1u);
            }
            if (jb.nkb == 4) {
              const uint64_t b2 = bd0 + 2 * kbs, b3 = bd0 + 3 * kbs;
              tq_umma_ts(d, a_t + 64, b2, idesc, 1u);
              tq_umma_ts(d, a_t + 72, b2 + 2, idesc, 1u);
              tq_umma_ts(d, a_t + 80, b2 + 4, idesc, 1u);
              tq_umma_ts(d, a_t + 88, b2 + 6, idesc, 1u);
              tq_umma_ts(d, a_t + 96, b3, idesc, 1u);
              tq_umma_ts(d, a_t + 104, b3 + 2, idesc, 1u);
              tq_umma_ts(d, a_t + 112, b3 + 4, idesc, 1u);
              tq_umma_ts(d, a_t + 120, b3 + 6, idesc, 1u);
            }
          }
          __syncwarp();
        } else {
          // K-block by K-block, waiting for the input quarters that are still being drained
          for (int kb = 0; kb < jb.nkb; ++kb) {
            bool waited = false;
            while (wi < 6 && jb.waits[wi] != 0xff && (jb.waits[wi] & 3) == kb) {
              const uint32_t bar = (jb.waits[wi] >> 2) & 7;
              const uint32_t seq = (uint32_t)t * a.ready_per_tile[bar] + (jb.waits[wi] >> 5);
              tq_wait(base + L::ready + 8 * bar, seq & 1, abort_addr, a.err, 0x21000000 | (j << 8) | bar);
              ++wi;
              waited = true;
            }
            if (waited) tc_fence_after_sync();
            const uint64_t bd = bd0 + (uint64_t)(kb * kbs);
            if (elect_one()) {
              tq_umma_ts(d, a_t + 32 * kb, bd, idesc, kb == 0 ? first : 1u);
              tq_umma_ts(d, a_t + 32 * kb + 8, bd + 2, idesc, 1u);
              tq_umma_ts(d, a_t + 32 * kb + 16, bd + 4, idesc, 1u);
              tq_umma_ts(d, a_t + 32 * kb + 24, bd + 6, idesc, 1u);
            }
            __syncwarp();
          }
        }
        if (elect_one()) {
          asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(base + L::issued), "r"(cnt + 1) : "memory");
          umma_commit(base + L::w_empty + 8 * s);
          if (jb.flags & QJ_COMMIT_ENC) umma_commit(base + L::enc_empty);
          if (jb.flags & QJ_COMMIT_DIR) umma_commit(base + L::dir_empty);
          if (jb.flags & QJ_COMMIT_ACC) umma_commit(base + L::acc_full + 8 * jb.acc);
        }
        __syncwarp();
        if (tr) tr[4 * j + 3] = clock64();
      }
  } else if (warp == 2) {
    // =============================== backward: gate producer ====================================================
    if constexpr (BWD) {
      uint32_t mc[2] = {0, 0};
      for (int t = 0; t < my_tiles; ++t) {
        const int64_t tile = blockIdx.x + (int64_t)t * gridDim.x;
        for (int si = 0; si < a.nsteps; ++si) {
          const TqStep st = c_tqsteps[PROG][si];
          if (st.mode != EPI_MASK) continue;
          const int wg = st.acc & 1;
          const uint32_t b = wg * 2 + (mc[wg] & 1), par = (mc[wg] >> 1) & 1;
          tq_wait(base + L::m_empty + 8 * b, par ^ 1, abort_addr, a.err, 0x60000000 | si);
          if (elect_one()) {
            mbar_expect_tx(base + L::m_full + 8 * b, kBlkBytes);
            tma_bulk_g2s(base + L::mstage + b * kBlkBytes,
                         a.stash_h + (size_t)tile * kStashTileBytes + (size_t)st.mask_blk * kBlkBytes, kBlkBytes,
                         base + L::m_full + 8 * b);
          }
          __syncwarp();
          ++mc[wg];
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // =============================== per-tile input blocks: thread == row =======================================
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
      if (t > 0) tq_wait(base + L::enc_empty, (t - 1) & 1, abort_addr, a.err, 0x30000000 | t);
      uint8_t* gblk = nullptr;
      if constexpr (BWD) gblk = a.stash_g + (size_t)tile * kStashTileBytes + (size_t)kGRaw * kBlkBytes + row_off;
      else if (a.stash_h != nullptr) gblk = a.stash_h + (size_t)tile * kStashTileBytes + (size_t)kHEnc * kBlkBytes + row_off;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t off = ((uint32_t)(c ^ (row & 7)) << 4);
        st_smem16(base + L::enc + row_off + off, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        if (gblk != nullptr) tq_st_global16(gblk + off, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(base + L::enc_full);
      if constexpr (!BWD) {
        // view-direction encoding (27 of 64 channels): A operand of the direction columns of views_linears.0
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
        if (t > 0) tq_wait(base + L::dir_empty, (t - 1) & 1, abort_addr, a.err, 0x31000000 | t);
        uint8_t* db = a.stash_h != nullptr ? a.stash_h + (size_t)tile * kStashTileBytes + (size_t)kHDir * kBlkBytes + row_off : nullptr;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t q[4] = {0u, 0u, 0u, 0u};
          if (c < 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = pack_bf16(e[8 * c + 2 * i], e[8 * c + 2 * i + 1]);
          }
          const uint32_t off = ((uint32_t)(c ^ (row & 7)) << 4);
          st_smem16(base + L::dir + row_off + off, q[0], q[1], q[2], q[3]);
          if (db != nullptr) tq_st_global16(db + off, q[0], q[1], q[2], q[3]);
        }
        fence_proxy_async_smem();
        mbar_arrive(base + L::dir_full);
      }
    }
  } else if (warp >= 8) {
    // =============================== epilogue: thread == row, one warpgroup per accumulator quarter ===========
    const int wg = (warp - 8) >> 2;
    const int row = ((warp & 3) << 5) | lane;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
    const uint32_t row_off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    const float* sbias = reinterpret_cast<const float*>(gen + L::bias);
    const bool storer = (threadIdx.x & 127) == 0;
    const uint32_t my_ostage = base + L::ostage + wg * kBlkBytes;
    uint32_t accpar = 0, mc = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t tile = blockIdx.x + (int64_t)t * gridDim.x;
      const int64_t p = tile * kTileRows + row;
      float sigma_acc = 0.f;
      unsigned long long* tr = (a.trace && blockIdx.x == 0 && t == a.trace_tile && (threadIdx.x & 127) == 0) ? a.trace + 960 : nullptr;
      for (int si = 0; si < a.nsteps; ++si) {
        const TqStep st = c_tqsteps[PROG][si];
        if ((st.acc & 1) != wg) continue;
        if (tr) tr[si * 4] = clock64();
        tq_wait(base + L::acc_full + 8 * st.acc, (accpar >> st.acc) & 1, abort_addr, a.err, 0x40000000 | (si << 8) | wg);
        accpar ^= 1u << st.acc;
        if (tr) tr[si * 4 + 1] = clock64();
        tc_fence_after_sync();
        const uint32_t acc_col = 64u * st.acc;
        if (st.mode == EPI_ALPHA) {
          uint32_t sv;
          tmem_ld1(lane_addr + acc_col + kTqColAlpha, sv);
          tmem_ld_wait();
          sigma_acc = __uint_as_float(sv);
          continue;
        }
        if (st.mode == EPI_OUT) {
          uint32_t c[4];
          tmem_ld4(lane_addr + acc_col + kTqColRgb, c);
          tmem_ld_wait();
          if (p < a.P) {
            float4 o;
            o.x = __uint_as_float(c[0]) + sbias[kBiasRgb + 0];
            o.y = __uint_as_float(c[1]) + sbias[kBiasRgb + 1];
            o.z = __uint_as_float(c[2]) + sbias[kBiasRgb + 2];
            o.w = sigma_acc + sbias[kBiasAlpha];
            st_stream4(reinterpret_cast<float4*>(a.raw) + p, o);
          }
          continue;
        }
        const bool relu = (st.mode == EPI_BIAS_RELU);
        const int ch0 = 64 * st.out_q;                               // first of this row's 64 channels in the layer
        const uint32_t out_col = (st.out_buf ? kTqA1 : kTqA0) + 32u * st.out_q;
        uint8_t* gout = nullptr;
        if (st.out_blk != 0xff) {
          if constexpr (BWD) gout = a.stash_g + (size_t)tile * kStashTileBytes + (size_t)st.out_blk * kBlkBytes;
          else if (a.stash_h != nullptr) gout = a.stash_h + (size_t)tile * kStashTileBytes + (size_t)st.out_blk * kBlkBytes;
        }
        if (gout != nullptr) {       // the previous bulk store must have finished reading the staging block
          if (storer) tq_bulk_wait_read0();
          tq_wg_bar(wg);
        }
        uint4 hm[8];
        if constexpr (BWD) {
          if (st.mode == EPI_MASK) {
            const uint32_t b = wg * 2 + (mc & 1);
            tq_wait(base + L::m_full + 8 * b, (mc >> 1) & 1, abort_addr, a.err, 0x41000000 | (si << 8) | wg);
            const uint32_t hb = base + L::mstage + b * kBlkBytes + row_off;
#pragma unroll
            for (int c = 0; c < 8; ++c) hm[c] = tq_ld_smem16(hb + ((uint32_t)(c ^ (row & 7)) << 4));
            mbar_arrive(base + L::m_empty + 8 * b);
            ++mc;
          }
        }
        uint32_t v[2][32];
        tmem_ld32(lane_addr + acc_col, v[0]);
        tmem_ld32(lane_addr + acc_col + 32, v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float f[32];
          if constexpr (BWD) {
            if (st.mode == EPI_MASK) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint32_t hw[4] = {hm[g * 4 + c].x, hm[g * 4 + c].y, hm[g * 4 + c].z, hm[g * 4 + c].w};
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
            for (int i = 0; i < 32; i += 4) {
              const float4 bb = bp[i >> 2];
              f[i] = __uint_as_float(v[g][i]) + bb.x; f[i + 1] = __uint_as_float(v[g][i + 1]) + bb.y;
              f[i + 2] = __uint_as_float(v[g][i + 2]) + bb.z; f[i + 3] = __uint_as_float(v[g][i + 3]) + bb.w;
            }
          }
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = relu ? pack_bf16_relu(f[2 * i], f[2 * i + 1]) : pack_bf16(f[2 * i], f[2 * i + 1]);
          if (!st.no_act) tq_tmem_st16(lane_addr + out_col + 16 * g, w);
          if (gout != nullptr) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              st_smem16(my_ostage + row_off + ((uint32_t)((g * 4 + c) ^ (row & 7)) << 4), w[4 * c], w[4 * c + 1], w[4 * c + 2],
                        w[4 * c + 3]);
          }
        }
        if (gout != nullptr) {       // block image complete -> one 16 KB bulk store
          fence_proxy_async_smem();
          tq_wg_bar(wg);
          if (storer) tq_bulk_s2g(gout, my_ostage, kBlkBytes);
        }
        if (!st.no_act) {
          tq_tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive(base + L::ready + 8 * (st.out_buf * 4 + st.out_q));
        }
        if (tr) tr[si * 4 + 2] = clock64();
      }
      tc_fence_before_sync();
      mbar_arrive(base + L::tile_done);
    }
    if (storer) tq_bulk_wait0();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, kTmemCols);
}

// ---- weight pre-pack: slab = nkb K-block images of [rows x 128 B], K-major, 128B swizzle ----------------------------------
struct TqParamPtrs {
  const float* w[GBN_NUM_LINEAR];
  const float* b[GBN_NUM_LINEAR];
};
struct TqPackHeader {
  uint32_t magic, plan, off_bias, off_wdir, off_bdir, total_bytes, njobs, pad;
};

__global__ void __launch_bounds__(256) tq_prepack_kernel(TqParamPtrs pp, uint8_t* __restrict__ out, int njobs, TqPackHeader hdr,
                                                         int prog) {
  if ((int)blockIdx.x < njobs) {
    const TsPackJob q = c_tqpack[prog][blockIdx.x];
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
  if (threadIdx.x == 0) *reinterpret_cast<TqPackHeader*>(out) = hdr;
  float* bias = reinterpret_cast<float*>(out + hdr.off_bias);
  for (int i = threadIdx.x; i < kTqBiasFloats; i += blockDim.x) {
    float v = 0.f;
    if (i < kBiasFeat) v = pp.b[i >> 8][i & 255];
    else if (i < kBiasAlpha) v = pp.b[LIN_FEATURE][i - kBiasFeat];
    else if (i == kBiasAlpha) v = pp.b[LIN_ALPHA][0];
    else if (i >= kBiasRgb && i < kBiasRgb + 3) v = pp.b[LIN_RGB][i - kBiasRgb];
    else if (i >= kTqBiasViews) v = pp.b[LIN_VIEWS][i - kTqBiasViews];
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
static std::once_flag g_tq_once;
static TqPlan g_tq_plan[2];
static bool g_tq_init[64];
static std::mutex g_tq_mutex;

static int tq_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static const TqPlan& tq_plan(int bwd) {
  std::call_once(g_tq_once, [] {
    g_tq_plan[0] = make_tq_plan(kTqFwd);
    g_tq_plan[1] = make_tq_plan(kTqBwd);
    if (tq_env("GBNERF_TQ_PREWAIT", 0))   // experiment: observe every input quarter before the first MMA of a job
      for (int p = 0; p < 2; ++p)
        for (auto& j : g_tq_plan[p].jobs)
          for (int i = 0; i < 6; ++i)
            if (j.waits[i] != 0xff) j.waits[i] &= ~3;
  });
  return g_tq_plan[bwd ? 1 : 0];
}

static int tq_ensure_device(cudaStream_t stream) {
  int dev = 0;
  GBN_CUDA(cudaGetDevice(&dev));
  GBN_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_tq_mutex);
  if (g_tq_init[dev]) return GBN_OK;
  for (int pr = 0; pr < 2; ++pr) {
    const TqPlan& p = tq_plan(pr);
    GBN_REQUIRE((int)p.jobs.size() <= kTqMaxJobs && (int)p.steps.size() <= kTqMaxSteps, "TQ table overflow");
    GBN_CUDA(cudaMemcpyToSymbolAsync(c_tqjobs, p.jobs.data(), p.jobs.size() * sizeof(TqJob), pr * kTqMaxJobs * sizeof(TqJob),
                                     cudaMemcpyHostToDevice, stream));
    GBN_CUDA(cudaMemcpyToSymbolAsync(c_tqsteps, p.steps.data(), p.steps.size() * sizeof(TqStep),
                                     pr * kTqMaxSteps * sizeof(TqStep), cudaMemcpyHostToDevice, stream));
    GBN_CUDA(cudaMemcpyToSymbolAsync(c_tqpack, p.pack.data(), p.pack.size() * sizeof(TsPackJob),
                                     pr * kTqMaxJobs * sizeof(TsPackJob), cudaMemcpyHostToDevice, stream));
  }
  GBN_CUDA(cudaFuncSetAttribute(nerf_mlp_tq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TqSmem<false>::alloc));
  GBN_CUDA(cudaFuncSetAttribute(nerf_mlp_tq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TqSmem<true>::alloc));
  g_tq_init[dev] = true;
  return GBN_OK;
}

size_t tq_packed_bytes(int bwd) { return tq_plan(bwd).total_bytes; }

int tq_prepack(const void* const* params, void* packed, int bwd, cudaStream_t st) {
  int rc = tq_ensure_device(st);
  if (rc != GBN_OK) return rc;
  TqParamPtrs pp;
  for (int i = 0; i < GBN_NUM_LINEAR; ++i) {
    pp.w[i] = static_cast<const float*>(params[2 * i]);
    pp.b[i] = static_cast<const float*>(params[2 * i + 1]);
  }
  const TqPlan& p = tq_plan(bwd);
  TqPackHeader hdr{0x4e425471u, (uint32_t)p.id, p.off_bias, p.off_wdir, p.off_bdir, p.total_bytes, (uint32_t)p.jobs.size(), 0};
  const int njobs = (int)p.jobs.size();
  tq_prepack_kernel<<<njobs + 1, 256, 0, st>>>(pp, static_cast<uint8_t*>(packed), njobs, hdr, bwd ? 1 : 0);
  return check_launch("tq_prepack_kernel");
}

void mlp_get_trace(unsigned long long** buf, int* tile);   // mlp_tc.cu

static void tq_fill(TqArgs& a, const TqPlan& p) {
  a.njobs = (int)p.jobs.size();
  a.nsteps = (int)p.steps.size();
  for (int i = 0; i < 8; ++i) a.ready_per_tile[i] = p.ready_per_tile[i];
  a.use_token = tq_env("GBNERF_TQ_TOKEN", 1);
  mlp_get_trace(&a.trace, &a.trace_tile);
}

int tq_forward(const void* packed, const float* ro, const float* rd, const float* vd, int64_t stride, const float* z,
               const float* pts, const float* emb, int64_t R, int S, float* raw, void* workspace, void* stash,
               cudaStream_t stream) {
  int rc = tq_ensure_device(stream);
  if (rc != GBN_OK) return rc;
  const TqPlan& p = tq_plan(0);
  int* err = reinterpret_cast<int*>(workspace);
  GBN_CUDA(cudaMemsetAsync(err, 0, 256, stream));
  TqArgs a{};
  a.packed = reinterpret_cast<const uint8_t*>(packed);
  a.ro = ro; a.rd = rd; a.z = z; a.pts = pts; a.emb = emb; a.vd = vd; a.raw = raw;
  a.stash_h = static_cast<uint8_t*>(stash); a.err = err; a.stride = stride; a.P = R * S; a.S = S;
  tq_fill(a, p);
  const int64_t ntiles = (a.P + kTileRows - 1) / kTileRows;
  const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  nerf_mlp_tq_kernel<false><<<grid, kTqThreads, TqSmem<false>::alloc, stream>>>(a);
  return check_launch("nerf_mlp_tq_kernel");
}

int tq_backward_data(const void* packed_bwd, const float* g_raw, int64_t P, const void* stash_h, void* stash_g, void* workspace,
                     cudaStream_t stream) {
  int rc = tq_ensure_device(stream);
  if (rc != GBN_OK) return rc;
  const TqPlan& p = tq_plan(1);
  int* err = reinterpret_cast<int*>(workspace);
  GBN_CUDA(cudaMemsetAsync(err, 0, 256, stream));
  TqArgs a{};
  a.packed = reinterpret_cast<const uint8_t*>(packed_bwd);
  a.raw = const_cast<float*>(g_raw);
  a.stash_h = static_cast<uint8_t*>(const_cast<void*>(stash_h));
  a.stash_g = static_cast<uint8_t*>(stash_g);
  a.err = err; a.P = P; a.S = 1;
  tq_fill(a, p);
  const int64_t ntiles = (P + kTileRows - 1) / kTileRows;
  const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  nerf_mlp_tq_kernel<true><<<grid, kTqThreads, TqSmem<true>::alloc, stream>>>(a);
  return check_launch("nerf_mlp_tq_kernel<bwd>");
}

}  // namespace gbn
