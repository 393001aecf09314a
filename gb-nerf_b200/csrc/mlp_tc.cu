// The 8x256 NeRF MLP (run_nerf_helpers.py:75-129) as ONE persistent, warp-specialised tcgen05 kernel, fused
// with ray-point generation and the sin/cos positional encoding (run.py:2317, run_nerf_helpers.py:23-53,
// run.py:1637-1653).
//
// One CTA per SM, 128 points (rows) per tile, the whole network per tile without touching HBM in between:
//   warp 0      weight producer : streams the pre-packed weight chunks (16 KB, UMMA K-major/128B-swizzle
//                                 images) L2 -> smem ring with bulk TMA (cp.async.bulk + mbarrier tx-count)
//   warp 1      MMA issuer      : one thread issues tcgen05.mma (M=128, N=128/16, K=32 B per step) with the
//                                 activation tile as A (smem), the weight chunk as B (smem), fp32
//                                 accumulators in TMEM; tcgen05.commit releases ring slots / signals layers
//   warp 2      TMEM allocator
//   warps 4-7   encoders        : next tile's points o + d*z, 63-channel encoding -> A operand of layer 0/5
//   warps 8-15  epilogue        : TMEM -> registers (tcgen05.ld) -> +bias, ReLU -> bf16/tf32 -> swizzled smem
//                                 = the next layer's A operand, handed over per 128-byte K-block so the next
//                                 layer's MMAs start while the rest of the tile is still being converted
// Accumulators ping-pong between two 256-column TMEM regions (X, Y), so layer l+1 runs on the tensor pipe
// while layer l's accumulator is drained.  The last epilogue writes (r,g,b,sigma_raw) as one float4 per row.
#include <mutex>

#include "common.cuh"
#include "mlp_layout.h"
#include "tc_ptx.cuh"

namespace gbn {

using namespace tc;

__constant__ MlpJob c_jobs[kNumPlans][kMaxJobs];
__constant__ int c_unit_begin[kNumPlans][kNumUnits + 1];
__constant__ EpiUnit c_epi[kNumPlans][kNumUnits];

// PROG: 0 forward bf16, 1 forward tf32, 2 backward (dgrad) bf16 — the index of the plan in mlp_layout.h
template <int PROG>
struct Cfg;
template <>
struct Cfg<0> {
  static constexpr int ESZ = 2, KB = 64, NBLK = 4, ENCB = 1, NST = 4, STAGE = 32768, NUNITS = 11;
  static constexpr uint32_t FMT = 1;
  static constexpr bool BWD = false, BF16 = true;
};
template <>
struct Cfg<1> {
  static constexpr int ESZ = 4, KB = 32, NBLK = 8, ENCB = 2, NST = 3, STAGE = 16384, NUNITS = 11;
  static constexpr uint32_t FMT = 2;
  static constexpr bool BWD = false, BF16 = false;
};
template <>
struct Cfg<2> {
  static constexpr int ESZ = 2, KB = 64, NBLK = 4, ENCB = 1, NST = 4, STAGE = 32768, NUNITS = 10;
  static constexpr uint32_t FMT = 1;
  static constexpr bool BWD = true, BF16 = true;
};

constexpr int kThreadsMlp = 512;

template <int PROG>
struct Smem {
  using C = Cfg<PROG>;
  static constexpr uint32_t act = 0;
  static constexpr uint32_t enc = act + C::NBLK * kBlkBytes;
  static constexpr uint32_t ring = enc + C::ENCB * kBlkBytes;
  static constexpr uint32_t bias = ring + C::NST * C::STAGE;
  static constexpr uint32_t bars = bias + kBiasFloats * 4;
  // barrier slots (8 B each)
  static constexpr uint32_t w_full = bars;
  static constexpr uint32_t w_empty = w_full + 8 * C::NST;
  static constexpr uint32_t acc_full = w_empty + 8 * C::NST;
  static constexpr uint32_t act_ready = acc_full + 8 * kNumUnits;
  static constexpr uint32_t enc_full = act_ready + 8 * C::NBLK;
  static constexpr uint32_t enc_empty = enc_full + 8;
  static constexpr uint32_t tmem_ptr = enc_empty + 8;
  static constexpr uint32_t abort_flag = tmem_ptr + 4;
  static constexpr uint32_t total = abort_flag + 4;
  static constexpr uint32_t alloc = total + 1024;  // slack for the 1024-byte alignment of the base
};

struct MlpArgs {
  const uint8_t* packed;
  const float* ro; const float* rd;   // ray origins / directions, pitch `stride` floats
  const float* z;                     // [R,S]
  const float* pts;                   // optional dense [P,3]
  const float* emb;                   // optional dense [P,90] (pre-embedded rows)
  const float* view_bias;             // [P / S, 128] fp32
  float* raw;                         // forward: [P,4] out.  backward: [P,4] gradient in (const in effect)
  uint8_t* stash_h;                   // forward: optional H stash out.  backward: H stash in
  uint8_t* stash_g;                   // backward: G stash out
  int* err;
  unsigned long long* trace;          // optional clock64 trace of CTA 0 (gbn_mlp_set_trace), else nullptr
  int trace_tile;
  int64_t stride;
  int64_t P;
  int S;
};

// bounded mbarrier wait that also honours the CTA-wide abort flag (set by whoever times out first)
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity, uint32_t abort_addr, int* err, int code) {
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

__device__ __forceinline__ void st_global16(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// 32 consecutive output channels of one row -> the swizzled K-block in shared memory (bf16: 4 chunks of 16 B at
// chunk0.., tf32: 8 chunks) and, when `gblk` (the BASE of a stash block) is given, the same values to that block in HBM
template <bool BF16>
__device__ __forceinline__ void store_group32(uint32_t row_addr, uint8_t* gblk, int row, int chunk0,
                                              const float (&f)[32], bool relu, bool to_smem = true) {
  if constexpr (BF16) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        w[i] = relu ? pack_bf16_relu(f[c * 8 + 2 * i], f[c * 8 + 2 * i + 1]) : pack_bf16(f[c * 8 + 2 * i], f[c * 8 + 2 * i + 1]);
      const uint32_t off = ((uint32_t)((chunk0 + c) ^ (row & 7)) << 4);
      if (to_smem) st_smem16(row_addr + off, w[0], w[1], w[2], w[3]);
      if (gblk != nullptr) st_global16(gblk + stash_chunk_off((uint32_t)row, (uint32_t)(chunk0 + c)), w[0], w[1], w[2], w[3]);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = to_tf32(relu ? fmaxf(f[c * 4 + i], 0.f) : f[c * 4 + i]);
      st_smem16(row_addr + (((chunk0 + c) ^ (row & 7)) << 4), w[0], w[1], w[2], w[3]);
    }
  }
}

template <int PROG>
__global__ void __launch_bounds__(kThreadsMlp, 1) nerf_mlp_kernel(const MlpArgs a) {
  using C = Cfg<PROG>;
  using L = Smem<PROG>;
  constexpr int LAST = C::NUNITS - 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));  // generic pointer to the aligned base
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (a.P + kTileRows - 1) / kTileRows;
  const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t abort_addr = base + L::abort_flag;
  const MlpJob* jobs = c_jobs[PROG];
  const int* ub = c_unit_begin[PROG];

  // ---- one-time setup ------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NST; ++i) { mbar_init(base + L::w_full + 8 * i, 1); mbar_init(base + L::w_empty + 8 * i, 1); }
    for (int i = 0; i < kNumUnits; ++i) mbar_init(base + L::acc_full + 8 * i, 1);
    for (int i = 0; i < C::NBLK; ++i) mbar_init(base + L::act_ready + 8 * i, C::KB == 64 ? 256 : 128);
    mbar_init(base + L::enc_full, 128);
    mbar_init(base + L::enc_empty, 1);
    *reinterpret_cast<volatile uint32_t*>(gen + L::abort_flag) = 0;
    mbar_init_fence();
  }
  if (warp == 2) tmem_alloc(base + L::tmem_ptr, kTmemCols);
  if constexpr (!C::BWD) {  // biases -> smem
    const float* gb = reinterpret_cast<const float*>(a.packed + reinterpret_cast<const uint32_t*>(a.packed)[2]);
    float* sb = reinterpret_cast<float*>(gen + L::bias);
    for (int i = threadIdx.x; i < kBiasFloats; i += blockDim.x) sb[i] = __ldg(gb + i);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + L::tmem_ptr);

  // Unit order per tile.  Forward: the first layer of tile t+1 is issued before the last (tiny rgb) unit of tile
  // t, so the tensor pipe has work while the final epilogue drains.  Backward: plain order.
  if (warp == 0) {
    // =============================== weight producer (warp-uniform loop, one elected lane issues) =========
    uint32_t cnt = 0;
    auto emit = [&](int u, int t) {
      unsigned long long* tr = (a.trace && blockIdx.x == 0 && t == a.trace_tile && lane == 0) ? a.trace + 640 : nullptr;
      for (int j = ub[u]; j < ub[u + 1]; ++j) {
        const uint32_t s = cnt % C::NST, par = (cnt / C::NST) & 1;
        if (tr) tr[2 * j] = clock64();
        wait_bar(base + L::w_empty + 8 * s, par ^ 1, abort_addr, a.err, 0x10000000 | j);
        if (tr) tr[2 * j + 1] = clock64();
        const uint32_t bytes = (uint32_t)jobs[j].w_bytes16 * 16;
        const uint8_t* src = a.packed + jobs[j].w_off;
        if (elect_one()) {
          mbar_expect_tx(base + L::w_full + 8 * s, bytes);
          tma_bulk_g2s(base + L::ring + s * C::STAGE, src, bytes, base + L::w_full + 8 * s);
        }
        __syncwarp();
        ++cnt;
      }
    };
    if constexpr (!C::BWD) {
      if (my_tiles > 0) emit(0, 0);
      for (int t = 0; t < my_tiles; ++t) {
        for (int u = 1; u < LAST; ++u) emit(u, t);
        if (t + 1 < my_tiles) emit(0, t + 1);
        emit(LAST, t);
      }
    } else {
      for (int t = 0; t < my_tiles; ++t)
        for (int u = 0; u <= LAST; ++u) emit(u, t);
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =============================================================
    // The whole warp walks the job list (warp-uniform control flow, so descriptors live in uniform registers and
    // UTCHMMA issues without a divergence waterfall); one elected lane issues the MMAs and commits.
    uint32_t cnt = 0, act_par = 0;
    auto issue = [&](int u, int t) {
      unsigned long long* tr = (a.trace && blockIdx.x == 0 && t == a.trace_tile && lane == 0) ? a.trace : nullptr;
      for (int j = ub[u]; j < ub[u + 1]; ++j) {
        const MlpJob jb = jobs[j];
        if (tr) tr[4 * j] = clock64();
        if (jb.flags & JF_WAIT_ENC) wait_bar(base + L::enc_full, t & 1, abort_addr, a.err, 0x20000000 | j);
        if (jb.flags & JF_WAIT_ACT) {
          wait_bar(base + L::act_ready + 8 * jb.a_blk, (act_par >> jb.a_blk) & 1, abort_addr, a.err, 0x21000000 | j);
          act_par ^= 1u << jb.a_blk;
        }
        const uint32_t s = cnt % C::NST, par = (cnt / C::NST) & 1;
        if (tr) tr[4 * j + 1] = clock64();
        wait_bar(base + L::w_full + 8 * s, par, abort_addr, a.err, 0x22000000 | j);
        if (tr) tr[4 * j + 2] = clock64();
        tc_fence_after_sync();
        const uint32_t a_addr = (jb.a_blk & kEncBlkFlag) ? base + L::enc + (jb.a_blk & 0x7f) * kBlkBytes
                                                         : base + L::act + jb.a_blk * kBlkBytes;
        const uint64_t adesc = smem_desc_sw128(a_addr);
        const uint64_t bdesc = smem_desc_sw128(base + L::ring + s * C::STAGE);
        const uint32_t idesc = make_idesc(C::FMT, 128, (uint32_t)jb.n16 * 16);
        const uint32_t d = tmem + jb.d_col;
        const uint32_t first = (jb.flags & JF_FIRST) ? 0u : 1u;
        const bool full = jb.ksteps == 4;
        if (elect_one()) {
          if constexpr (C::BF16) {
            umma_bf16(d, adesc, bdesc, idesc, first);
            if (full) {
              umma_bf16(d, adesc + 2, bdesc + 2, idesc, 1u);
              umma_bf16(d, adesc + 4, bdesc + 4, idesc, 1u);
              umma_bf16(d, adesc + 6, bdesc + 6, idesc, 1u);
            }
          } else {
            umma_tf32(d, adesc, bdesc, idesc, first);
            umma_tf32(d, adesc + 2, bdesc + 2, idesc, 1u);
            umma_tf32(d, adesc + 4, bdesc + 4, idesc, 1u);
            umma_tf32(d, adesc + 6, bdesc + 6, idesc, 1u);
          }
          umma_commit(base + L::w_empty + 8 * s);
          if (jb.flags & JF_COMMIT_ENC) umma_commit(base + L::enc_empty);
          if (jb.flags & JF_COMMIT_ACC) umma_commit(base + L::acc_full + 8 * jb.unit);
        }
        __syncwarp();
        if (tr) tr[4 * j + 3] = clock64();
        ++cnt;
      }
    };
    if constexpr (!C::BWD) {
      if (my_tiles > 0) issue(0, 0);
      for (int t = 0; t < my_tiles; ++t) {
        for (int u = 1; u < LAST; ++u) issue(u, t);
        if (t + 1 < my_tiles) issue(0, t + 1);
        issue(LAST, t);
      }
    } else {
      for (int t = 0; t < my_tiles; ++t)
        for (int u = 0; u <= LAST; ++u) issue(u, t);
    }
  } else if (warp >= 4 && warp < 8) {
    // =============================== producers of the per-tile input K-block: thread == row ==================
    // forward: points o + d*z and their 63-channel encoding (A operand of layers 0 and 5)
    // backward: the incoming gradient (g_r, g_g, g_b, g_sigma) padded to a 64-column bf16 block
    const int row = threadIdx.x - 128;
    const uint32_t row_off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t tile = blockIdx.x + (int64_t)t * gridDim.x;
      const int64_t p = tile * kTileRows + row;
      unsigned long long* tr = (a.trace && blockIdx.x == 0 && t == a.trace_tile && row == 0) ? a.trace + 1216 : nullptr;
      if (tr) tr[0] = clock64();
      if constexpr (C::BWD) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < a.P) g = ld_stream4(reinterpret_cast<const float4*>(a.raw) + p);
        if (tr) tr[1] = clock64();
        if (t > 0) wait_bar(base + L::enc_empty, (t - 1) & 1, abort_addr, a.err, 0x30000000 | t);
        if (tr) tr[2] = clock64();
        uint8_t* gblk = a.stash_g + (size_t)tile * kStashTileBytes + (size_t)kGRaw * kBlkBytes;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t w0 = c == 0 ? pack_bf16(g.x, g.y) : 0u, w1 = c == 0 ? pack_bf16(g.z, g.w) : 0u;
          const uint32_t off = ((uint32_t)(c ^ (row & 7)) << 4);
          st_smem16(base + L::enc + row_off + off, w0, w1, 0u, 0u);
          st_global16(gblk + stash_chunk_off((uint32_t)row, (uint32_t)c), w0, w1, 0u, 0u);
        }
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
        if (tr) tr[1] = clock64();
        if (t > 0) wait_bar(base + L::enc_empty, (t - 1) & 1, abort_addr, a.err, 0x30000000 | t);
        if (tr) tr[2] = clock64();
        uint8_t* gblk = (C::BF16 && a.stash_h != nullptr)
                            ? a.stash_h + (size_t)tile * kStashTileBytes + (size_t)kHEnc * kBlkBytes : nullptr;
        // 64 channels: bf16 -> one K-block (8 chunks); tf32 -> two K-blocks (8 chunks each)
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = e[g * 32 + i];
          if constexpr (C::BF16)
            store_group32<true>(base + L::enc + row_off, gblk, row, g * 4, f, false);
          else
            store_group32<false>(base + L::enc + g * kBlkBytes + row_off, nullptr, row, 0, f, false);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(base + L::enc_full);
      if (tr) tr[3] = clock64();
    }
  } else if (warp >= 8) {
    // =============================== epilogue: thread == row ================================================
    const int wg = (warp - 8) >> 2;                 // 0 or 1
    const int row = ((warp & 3) << 5) | lane;       // TMEM lane == tile row; warp%4 selects the lane quadrant
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
    const uint32_t row_off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    const float* sbias = reinterpret_cast<const float*>(gen + L::bias);
    constexpr int GPB = C::KB / 32;                 // 32-column groups per K-block (bf16 2, tf32 1)
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t tile = blockIdx.x + (int64_t)t * gridDim.x;
      const int64_t p = tile * kTileRows + row;
      const uint32_t par = t & 1;
      float sigma_acc = 0.f;
      unsigned long long* tr = (a.trace && blockIdx.x == 0 && t == a.trace_tile && (threadIdx.x & 127) == 0)
                                   ? a.trace + 960 + wg * 128 : nullptr;
      uint8_t* const tile_h = (C::BF16 && a.stash_h != nullptr) ? a.stash_h + (size_t)tile * kStashTileBytes : nullptr;
      uint8_t* const tile_g = (C::BWD) ? a.stash_g + (size_t)tile * kStashTileBytes : nullptr;
      for (int u = 0; u < C::NUNITS; ++u) {
        const EpiUnit eu = c_epi[PROG][u];
        if (eu.mode == EPI_OUT) break;
        if (tr) tr[u * 10] = clock64();
        wait_bar(base + L::acc_full + 8 * u, par, abort_addr, a.err, 0x40000000 | (u << 8) | wg);
        if (tr) tr[u * 10 + 1] = clock64();
        tc_fence_after_sync();
        const uint32_t col0 = (u & 1) ? kColY : kColX;
        const int nb = eu.nb;
        const bool relu = (eu.mode == EPI_BIAS_RELU || eu.mode == EPI_VBIAS_RELU);
        const float* vb = nullptr;
        if (eu.mode == EPI_VBIAS_RELU) {
          const int64_t pr = p < a.P ? p : a.P - 1;
          vb = a.view_bias + (pr / a.S) * 128;
          if (wg == 0) {  // sigma accumulator finished together with the feature layer
            uint32_t sv;
            tmem_ld1(lane_addr + kColAlpha, sv);
            tmem_ld_wait();
            sigma_acc = __uint_as_float(sv);
          }
        }
        // bf16: a K-block is 64 columns = two 32-column groups, one per warpgroup, so block 0 (what the next
        // layer's first MMAs wait for) is ready after half the work; tf32: 32-column blocks alternate
        for (int b = (GPB == 2 ? 0 : wg); b < nb; b += (GPB == 2 ? 1 : 2)) {
          const int g = (GPB == 2) ? wg : 0;
          const int c0 = b * C::KB + g * 32;
          uint4 hm[4];
          if constexpr (C::BWD) {
            if (eu.mode == EPI_MASK) {  // the activation whose sign gates this gradient (64 B of this row)
              const uint8_t* hb = tile_h + (size_t)(eu.mask_blk + b) * kBlkBytes;
#pragma unroll
              for (int c = 0; c < 4; ++c)
                hm[c] = __ldg(reinterpret_cast<const uint4*>(hb + stash_chunk_off((uint32_t)row, (uint32_t)(g * 4 + c))));
            }
          }
          uint32_t v[32];
          tmem_ld32(lane_addr + col0 + c0, v);
          tmem_ld_wait();
          if (tr) tr[u * 10 + 2 + 2 * (GPB == 2 ? b : (b >> 1))] = clock64();
          float f[32];
          if constexpr (C::BWD) {
            if (eu.mode == EPI_MASK) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint32_t w[4] = {hm[c].x, hm[c].y, hm[c].z, hm[c].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  // post-ReLU activations are >= 0: a bf16 half is "on" iff it is non-zero (and not -0)
                  f[c * 8 + 2 * i] = ((w[i] & 0x7fffu) != 0u) ? __uint_as_float(v[c * 8 + 2 * i]) : 0.f;
                  f[c * 8 + 2 * i + 1] = ((w[i] & 0x7fff0000u) != 0u) ? __uint_as_float(v[c * 8 + 2 * i + 1]) : 0.f;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
            }
          } else if (eu.mode == EPI_VBIAS_RELU) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(vb + c0 + i));
              f[i] = __uint_as_float(v[i]) + bb.x; f[i + 1] = __uint_as_float(v[i + 1]) + bb.y;
              f[i + 2] = __uint_as_float(v[i + 2]) + bb.z; f[i + 3] = __uint_as_float(v[i + 3]) + bb.w;
            }
          } else {
            const float4* bp = reinterpret_cast<const float4*>(sbias + eu.bias_off + c0);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bb = bp[i >> 2];
              f[i] = __uint_as_float(v[i]) + bb.x; f[i + 1] = __uint_as_float(v[i + 1]) + bb.y;
              f[i + 2] = __uint_as_float(v[i + 2]) + bb.z; f[i + 3] = __uint_as_float(v[i + 3]) + bb.w;
            }
          }
          uint8_t* gout = nullptr;
          if constexpr (C::BWD) gout = tile_g + (size_t)(eu.out_blk + b) * kBlkBytes;
          else if (tile_h != nullptr) gout = tile_h + (size_t)(eu.out_blk + b) * kBlkBytes;
          store_group32<C::BF16>(base + L::act + b * kBlkBytes + row_off, gout, row, g * 4, f, relu, !eu.no_act);
          tc_fence_before_sync();
          if (!eu.no_act) {
            fence_proxy_async_smem();
            mbar_arrive(base + L::act_ready + 8 * b);
          }
          if (tr) tr[u * 10 + 3 + 2 * (GPB == 2 ? b : (b >> 1))] = clock64();
        }
      }
      if constexpr (!C::BWD) {
        // ---- last unit: rgb accumulator + sigma -> raw[p] ---------------------------------------------------
        if (tr) tr[100] = clock64();
        wait_bar(base + L::acc_full + 8 * LAST, par, abort_addr, a.err, 0x40000000 | (LAST << 8) | wg);
        if (tr) tr[101] = clock64();
        tc_fence_after_sync();
        if (wg == 0) {
          uint32_t c[4];
          tmem_ld4(lane_addr + kColRgb, c);
          tmem_ld_wait();
          if (p < a.P) {
            float4 o;
            o.x = __uint_as_float(c[0]) + sbias[kBiasRgb + 0];
            o.y = __uint_as_float(c[1]) + sbias[kBiasRgb + 1];
            o.z = __uint_as_float(c[2]) + sbias[kBiasRgb + 2];
            o.w = sigma_acc + sbias[kBiasAlpha];
            st_stream4(reinterpret_cast<float4*>(a.raw) + p, o);
          }
        }
        if (tr) tr[102] = clock64();
      }
      tc_fence_before_sync();
    }
  }

  // ---- teardown ------------------------------------------------------------------------------------------
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, kTmemCols);
}

// =========================================================================================================
// host side
// =========================================================================================================
static std::once_flag g_plan_once;
static MlpPlan g_plan[kNumPlans];
static bool g_dev_init[64];
static std::mutex g_dev_mutex;

static unsigned long long* g_trace = nullptr;  // diagnostic only (gbn_mlp_set_trace)
static int g_trace_tile = 0;

static const MlpPlan& plan(int which) {
  std::call_once(g_plan_once, [] {
    for (int i = 0; i < kNumPlans; ++i) g_plan[i] = make_plan(i);
  });
  return g_plan[which];
}

const MlpPlan& mlp_plan(int which) { return plan(which); }
void mlp_get_trace(unsigned long long** buf, int* tile) { *buf = g_trace; *tile = g_trace_tile; }

// per-device one-time setup: job tables -> constant memory, opt-in shared memory size
static int ensure_device(cudaStream_t stream) {
  int dev = 0;
  GBN_CUDA(cudaGetDevice(&dev));
  GBN_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  if (g_dev_init[dev]) return GBN_OK;
  for (int pr = 0; pr < kNumPlans; ++pr) {
    const MlpPlan& p = plan(pr);
    GBN_REQUIRE((int)p.jobs.size() <= kMaxJobs, "job table overflow");
    GBN_CUDA(cudaMemcpyToSymbol(c_jobs, p.jobs.data(), p.jobs.size() * sizeof(MlpJob),
                                     pr * kMaxJobs * sizeof(MlpJob), cudaMemcpyHostToDevice));
    GBN_CUDA(cudaMemcpyToSymbol(c_unit_begin, p.unit_begin, sizeof(p.unit_begin),
                                     pr * sizeof(p.unit_begin), cudaMemcpyHostToDevice));
    GBN_CUDA(cudaMemcpyToSymbol(c_epi, p.epi, sizeof(p.epi), pr * sizeof(p.epi), cudaMemcpyHostToDevice));
  }
  GBN_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<0>::alloc));
  GBN_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<1>::alloc));
  GBN_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<2>::alloc));
  g_dev_init[dev] = true;
  return GBN_OK;
}

int launch_view_bias(const uint8_t* packed, const MlpPlan& p, const float* viewdirs, int64_t stride,
                     const float* emb, int64_t n, float* out, cudaStream_t stream);  // mlp_aux.cu
bool mlp_use_ts();                                                                    // mlp_aux.cu
int mlp_variant();                                                                    // mlp_aux.cu
size_t ts_packed_bytes(int bwd);                                                      // mlp_ts.cu
int ts_forward(const void* packed, const float* ro, const float* rd, const float* vd, int64_t stride, const float* z,
               const float* pts, const float* emb, int64_t R, int S, float* raw, void* workspace, void* stash,
               cudaStream_t stream);
int ts_backward_data(const void* packed_bwd, const float* g_raw, int64_t P, const void* stash_h, void* stash_g, void* workspace,
                     cudaStream_t stream);

static int run_mlp(const void* packed, int precision, const float* ro, const float* rd, const float* vd,
                   int64_t stride, const float* z, const float* pts, const float* emb, int64_t R, int S, float* raw,
                   void* workspace, void* stash, cudaStream_t stream) {
  GBN_REQUIRE(precision == GBN_PRECISION_BF16 || precision == GBN_PRECISION_TF32, "mlp: unknown precision %d", precision);
  GBN_REQUIRE(R >= 0 && S >= 1, "mlp: bad sizes R=%lld S=%d", (long long)R, S);
  if (R == 0) return GBN_OK;
  GBN_REQUIRE(packed && raw && workspace, "mlp: null pointer");
  GBN_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 255) == 0, "mlp: packed weights must be 256-byte aligned");
  GBN_REQUIRE(((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
              "mlp: raw / workspace must be 16-byte aligned");
  GBN_REQUIRE(stash == nullptr || precision == GBN_PRECISION_BF16, "mlp: the training stash exists for bf16 only");
  GBN_REQUIRE((reinterpret_cast<uintptr_t>(stash) & 127) == 0, "mlp: stash must be 128-byte aligned");
  if (precision == GBN_PRECISION_BF16 && mlp_variant() == 1)
    return ts_forward(packed, ro, rd, vd, stride, z, pts, emb, R, S, raw, workspace, stash, stream);
  int rc = ensure_device(stream);
  if (rc != GBN_OK) return rc;
  const MlpPlan& p = plan(precision);
  int* err = reinterpret_cast<int*>(workspace);
  float* vbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 256);
  GBN_CUDA(cudaMemsetAsync(err, 0, 256, stream));
  rc = launch_view_bias(reinterpret_cast<const uint8_t*>(packed), p, vd, stride, emb, R, vbias, stream);
  if (rc != GBN_OK) return rc;
  MlpArgs a{};
  a.packed = reinterpret_cast<const uint8_t*>(packed);
  a.ro = ro; a.rd = rd; a.z = z; a.pts = pts; a.emb = emb;
  a.view_bias = vbias; a.raw = raw; a.err = err;
  a.stash_h = static_cast<uint8_t*>(stash);
  a.trace = g_trace; a.trace_tile = g_trace_tile;
  a.stride = stride; a.P = R * S; a.S = S;
  const int64_t ntiles = (a.P + kTileRows - 1) / kTileRows;
  const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  if (precision == GBN_PRECISION_BF16)
    nerf_mlp_kernel<0><<<grid, kThreadsMlp, Smem<0>::alloc, stream>>>(a);
  else
    nerf_mlp_kernel<1><<<grid, kThreadsMlp, Smem<1>::alloc, stream>>>(a);
  return check_launch("nerf_mlp_kernel");
}

}  // namespace gbn

using namespace gbn;

extern "C" size_t gbn_mlp_packed_bytes(int precision) {
  if (precision < 0 || precision >= kNumPlans) return 0;
  if (precision != GBN_PRECISION_TF32 && mlp_variant() == 1) return ts_packed_bytes(precision == GBN_PACK_BWD_BF16);
  return mlp_plan(precision).total_bytes;
}

extern "C" int gbn_mlp_set_trace(void* buf, int tile) {
  g_trace = static_cast<unsigned long long*>(buf);
  g_trace_tile = tile;
  return GBN_OK;
}

namespace gbn {
int ts_watchdog_report(unsigned int* out, int words);                                             // mlp_ts.cu
int ts_debug_plan(int bwd, void* jobs, int max_jobs, void* steps, int max_steps, int* meta);     // mlp_ts.cu
}
extern "C" int gbn_debug_ts_plan(int bwd, void* jobs, int max_jobs, void* steps, int max_steps, int* meta) {
  GBN_REQUIRE(jobs && steps && meta, "debug_ts_plan: null pointer");
  GBN_REQUIRE(gbn::ts_debug_plan(bwd, jobs, max_jobs, steps, max_steps, meta) == 0, "debug_ts_plan: buffers too small");
  return GBN_OK;
}
extern "C" int gbn_watchdog_report(unsigned int* out, int words) {
  if (out == nullptr || words <= 0) return 0;
  return gbn::ts_watchdog_report(out, words);
}

extern "C" size_t gbn_mlp_workspace_bytes(int64_t R) { return 256 + (size_t)(R < 0 ? 0 : R) * 128 * sizeof(float); }

extern "C" size_t gbn_mlp_stash_bytes(int64_t P) {
  if (P <= 0) return 0;
  return (size_t)((P + kTileRows - 1) / kTileRows) * kStashTileBytes;
}

extern "C" int gbn_mlp_forward(const void* packed, int precision, const float* rays_o, const float* rays_d,
                               const float* viewdirs, int64_t ray_stride, const float* z, const float* pts,
                               int64_t R, int S, float* raw, void* workspace, void* stash, void* stream) {
  GBN_REQUIRE(R == 0 || viewdirs, "mlp_forward: viewdirs is required");
  GBN_REQUIRE(R == 0 || pts || (rays_o && rays_d && z), "mlp_forward: need pts or (rays_o, rays_d, z)");
  return run_mlp(packed, precision, rays_o, rays_d, viewdirs, ray_stride, z, pts, nullptr, R, S, raw, workspace, stash,
                 (cudaStream_t)stream);
}

extern "C" int gbn_mlp_forward_embedded(const void* packed, int precision, const float* emb, int64_t P, float* raw,
                                        void* workspace, void* stash, void* stream) {
  GBN_REQUIRE(P == 0 || emb, "mlp_forward_embedded: null pointer");
  return run_mlp(packed, precision, nullptr, nullptr, nullptr, 0, nullptr, nullptr, emb, P, 1, raw, workspace, stash,
                 (cudaStream_t)stream);
}

extern "C" int gbn_mlp_backward_data(const void* packed_bwd, const float* g_raw, int64_t P, const void* stash_h,
                                     void* stash_g, void* workspace, void* stream) {
  GBN_REQUIRE(P >= 0, "mlp_backward_data: bad size");
  if (P == 0) return GBN_OK;
  GBN_REQUIRE(packed_bwd && g_raw && stash_h && stash_g && workspace, "mlp_backward_data: null pointer");
  GBN_REQUIRE((reinterpret_cast<uintptr_t>(packed_bwd) & 255) == 0, "mlp_backward_data: packed weights must be 256-byte aligned");
  GBN_REQUIRE(((reinterpret_cast<uintptr_t>(g_raw) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0 &&
                  ((reinterpret_cast<uintptr_t>(stash_h) | reinterpret_cast<uintptr_t>(stash_g)) & 127) == 0,
              "mlp_backward_data: misaligned buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (mlp_variant() == 1) return ts_backward_data(packed_bwd, g_raw, P, stash_h, stash_g, workspace, st);
  int rc = ensure_device(st);
  if (rc != GBN_OK) return rc;
  int* err = reinterpret_cast<int*>(workspace);
  GBN_CUDA(cudaMemsetAsync(err, 0, 256, st));
  MlpArgs a{};
  a.packed = reinterpret_cast<const uint8_t*>(packed_bwd);
  a.raw = const_cast<float*>(g_raw);
  a.stash_h = static_cast<uint8_t*>(const_cast<void*>(stash_h));
  a.stash_g = static_cast<uint8_t*>(stash_g);
  a.err = err;
  a.trace = g_trace; a.trace_tile = g_trace_tile;
  a.P = P; a.S = 1;
  const int64_t ntiles = (P + kTileRows - 1) / kTileRows;
  const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  nerf_mlp_kernel<2><<<grid, kThreadsMlp, Smem<2>::alloc, st>>>(a);
  return check_launch("nerf_mlp_kernel<bwd>");
}
