// Parameter gradients of the 8x256 NeRF MLP from the training stashes (H: forward activations, G: pre-activation
// gradients written by the dgrad pass of mlp_tc.cu).  Autograd equivalent: the weight/bias gradients of the twelve
// nn.Linear modules of run_nerf_helpers.py:88-104.
//
//   wgrad_tc_kernel     dW_l[n,k] = sum_p G_l[p,n] H_{l-1}[p,k] for the ten wide layers on tcgen05: both operands are
//                       the stash half blocks themselves ([8-channel chunk][64 points][16 B], mlp_layout.h), consumed as
//                       no-swizzle MN-major UMMA operands (M = out channel, N = in channel, K = points), fp32 accumulators for
//                       a whole 256x256 weight gradient in TMEM (2 x 256 columns), split-K over point tiles across
//                       CTAs, flushed with red.global.add.  Bias gradients (column sums of G) are taken from the
//                       same shared-memory tiles by otherwise idle warps.
//                       rgb_linear / alpha_linear ride along as two more items whose A operand is the padded
//                       g_raw block (M = 128 filled by loading it twice; rows 0..2 / 3 are the gradients).
//   viewdir_grad_kernel the 27 direction columns of views_linears.0 (constant along a ray, folded into a per-ray
//                       bias in the forward): per-ray sums of g_hv, then an outer product with enc(viewdir).
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "mlp_layout.h"
#include "tc_ptx.cuh"

namespace gbn {

using namespace tc;

struct WgItem {
  uint8_t a_blk;      // first G-stash block of the gradient (A operand), 64 output channels per block
  uint8_t m_blocks;   // 2 (M = 128) or 4 (M = 256, two accumulator halves)
  uint8_t b_blk;      // first H-stash block of the layer input (B operand)
  uint8_t n_blocks;   // 1 (N = 64) or 4 (N = 256)
  uint8_t layer;      // index into the 12 linears (order of gbn_mlp_prepack_weights)
  uint8_t do_bias;    // this item also reduces the bias gradient of `layer`
  uint8_t head;       // 0: wide layer | 1: alpha_linear, 2: rgb_linear (A = the padded g_raw block, loaded twice
                      //    to fill M = 128; only accumulator rows 3 / 0..2 are meaningful)
  uint8_t pad;
  uint16_t ld;        // in_features of that linear
  uint16_t col0;      // first weight column this item produces
  uint16_t n_valid;   // columns that exist (63 for the encoding block)
  uint16_t cta_begin, cta_end;   // CTAs [begin, end) of the 148 share this item's point tiles
};

constexpr int kWgItems = 14;
constexpr int kWgThreads = 384;      // warp 0 producer, 1 MMA, 2 TMEM alloc, 4-7 bias sums, 8-11 flush
constexpr int kWgStages = 3;
constexpr int kWgHalf = 8192;        // bytes of the 64-point half of a 16 KB block image
constexpr int kWgStageBytes = 8 * kWgHalf;

__constant__ WgItem c_wg[kWgItems];

struct WgArgs {
  const uint8_t* stash_h;
  const uint8_t* stash_g;
  float* w[GBN_NUM_LINEAR];
  float* b[GBN_NUM_LINEAR];
  int64_t ntiles;
  int* err;
};

struct WgSmem {
  static constexpr uint32_t ring = 0;
  static constexpr uint32_t full = ring + kWgStages * kWgStageBytes;
  static constexpr uint32_t empty = full + 8 * kWgStages;
  static constexpr uint32_t done = empty + 8 * kWgStages;
  static constexpr uint32_t tmem_ptr = done + 8;
  static constexpr uint32_t abort_flag = tmem_ptr + 4;
  static constexpr uint32_t total = abort_flag + 4;
  static constexpr uint32_t alloc = total + 1024;
};

// MN-major operand without swizzle (cute INTERLEAVE: ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)), T = 8 bf16 = 16 B): a core
// matrix is 8 K rows (points) x 16 B of MN (8 channels) = 128 contiguous bytes; core matrices are LBO apart along K
// (128 B: the next 8 points of the same chunk) and SBO apart along MN (1 KB: the next 8-channel chunk, uniform across the
// half blocks of a stage because a half block is exactly 8 chunks)
__device__ __forceinline__ uint64_t smem_desc_mn_interleave(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
constexpr uint32_t kWgLbo = 128, kWgSbo = 1024;

__device__ __forceinline__ void wg_wait(uint32_t bar, uint32_t parity, uint32_t abort_addr, int* err, int code) {
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

__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// four adjacent floats in one reduction (16-byte aligned address): a quarter of the instructions and L2 transactions of
// the flush, which is a fixed ~50 us tail of every wgrad launch
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const WgArgs a) {
  using L = WgSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t abort_addr = base + L::abort_flag;

  // which item, which tiles
  int it = 0;
  for (int i = 0; i < kWgItems; ++i)
    if ((int)blockIdx.x >= c_wg[i].cta_begin && (int)blockIdx.x < c_wg[i].cta_end) it = i;
  const WgItem w = c_wg[it];
  const int ncta = w.cta_end - w.cta_begin, rank = (int)blockIdx.x - w.cta_begin;
  const int64_t t_begin = a.ntiles * rank / ncta, t_end = a.ntiles * (rank + 1) / ncta;
  const int64_t nstages = (t_end - t_begin) * 2;   // 64-point half tiles
  const int nhalves = w.m_blocks / 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(base + L::full + 8 * i, 1);
      mbar_init(base + L::empty + 8 * i, 1 + 4);   // MMA commit + the four bias-sum warps
    }
    mbar_init(base + L::done, 1);
    *reinterpret_cast<volatile uint32_t*>(gen + L::abort_flag) = 0;
    mbar_init_fence();
  }
  if (warp == 2) tmem_alloc(base + L::tmem_ptr, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + L::tmem_ptr);

  if (warp == 0) {
    // ---- producer: per stage (m_blocks + n_blocks) bulk copies of the 8 KB half-block images -----------------
    for (int64_t i = 0; i < nstages; ++i) {
      const uint32_t s = (uint32_t)(i % kWgStages), par = (uint32_t)((i / kWgStages) & 1);
      wg_wait(base + L::empty + 8 * s, par ^ 1, abort_addr, a.err, 0x50000000 | (int)s);
      const int64_t tile = t_begin + (i >> 1);
      const uint32_t half_off = (uint32_t)(i & 1) * kWgHalf;
      const uint8_t* g = a.stash_g + (size_t)tile * kStashTileBytes + (size_t)w.a_blk * kBlkBytes + half_off;
      const uint8_t* h = a.stash_h + (size_t)tile * kStashTileBytes + (size_t)w.b_blk * kBlkBytes + half_off;
      if (elect_one()) {
        mbar_expect_tx(base + L::full + 8 * s, (uint32_t)(w.m_blocks + w.n_blocks) * kWgHalf);
        for (int b = 0; b < w.m_blocks; ++b)
          tma_bulk_g2s(base + L::ring + s * kWgStageBytes + b * kWgHalf, g + (w.head ? 0 : (size_t)b * kBlkBytes),
                       kWgHalf, base + L::full + 8 * s);
        for (int b = 0; b < w.n_blocks; ++b)
          tma_bulk_g2s(base + L::ring + s * kWgStageBytes + (4 + b) * kWgHalf, h + (size_t)b * kBlkBytes, kWgHalf,
                       base + L::full + 8 * s);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer: D[half][n_out, k_in] += A^T B over 64 points per stage (4 K-steps of 16) ----------------
    const uint32_t idesc = make_idesc(1, 128, (uint32_t)w.n_blocks * 64) | (1u << 15) | (1u << 16);
    for (int64_t i = 0; i < nstages; ++i) {
      const uint32_t s = (uint32_t)(i % kWgStages), par = (uint32_t)((i / kWgStages) & 1);
      wg_wait(base + L::full + 8 * s, par, abort_addr, a.err, 0x51000000 | (int)s);
      tc_fence_after_sync();
      const uint32_t sa = base + L::ring + s * kWgStageBytes, sb = sa + 4 * kWgHalf;
      if (elect_one()) {
        for (int hf = 0; hf < nhalves; ++hf) {
          const uint64_t adesc = smem_desc_mn_interleave(sa + hf * 2 * kWgHalf, kWgLbo, kWgSbo);
          const uint64_t bdesc = smem_desc_mn_interleave(sb, kWgLbo, kWgSbo);
          const uint32_t d = tmem + hf * 256;
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 16 points = two 8-point core-matrix rows of 128 B -> +16 in the address field
            umma_bf16(d, adesc + 16 * k, bdesc + 16 * k, idesc, (i == 0 && k == 0) ? 0u : 1u);
        }
        umma_commit(base + L::empty + 8 * s);
        if (i == nstages - 1) umma_commit(base + L::done);
      }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- bias gradient: column sums of the G tile.  Warp q owns the 8-channel chunks q, q + 4, ... of the tile, a lane
    // the points lane and lane + 32 of the stage: one conflict-free 16-byte read per chunk and point, eight running sums
    // per chunk in registers, and one shuffle reduction per channel at the very end
    const int q = warp - 4;
    constexpr int kMaxChunks = 8;                                     // (4 blocks x 8 chunks) / 4 warps
    const int nchunks = !w.do_bias ? 0 : (w.head ? (q == 0 ? 1 : 0) : w.m_blocks * 2);   // chunks of this warp
    float acc[kMaxChunks][8];
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
    for (int64_t i = 0; i < nstages; ++i) {
      const uint32_t s = (uint32_t)(i % kWgStages), par = (uint32_t)((i / kWgStages) & 1);
      wg_wait(base + L::full + 8 * s, par, abort_addr, a.err, 0x52000000 | (int)s);
      const uint8_t* tile = gen + L::ring + s * kWgStageBytes;
#pragma unroll
      for (int c = 0; c < kMaxChunks; ++c) {
        if (c < nchunks) {
          const uint8_t* cp = tile + (size_t)(q + 4 * c) * 1024;       // chunk (q + 4c): block (q + 4c) / 8, chunk % 8 - contiguous
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint4 v = *reinterpret_cast<const uint4*>(cp + (lane + 32 * h) * 16);
            const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc[c][2 * e] += __uint_as_float(wv[e] << 16);
              acc[c][2 * e + 1] += __uint_as_float(wv[e] & 0xffff0000u);
            }
          }
        }
      }
      // generic-proxy reads of a bulk-copied stage, then the stage goes back to the bulk-copy producer: proxy fence
      // first (same rule as the gate staging of mlp_ts.cu; an mbarrier hand-over does not order them by itself)
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(base + L::empty + 8 * s);
    }
    if (nstages > 0) {
#pragma unroll
      for (int c = 0; c < kMaxChunks; ++c) {
        if (c < nchunks) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float v = acc[c][e];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[c][e] = v;
          }
          if (lane < 8) {
            float mine = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) mine = lane == e ? acc[c][e] : mine;
            const int ch = (q + 4 * c) * 8 + lane;                     // channel within the item's M rows
            if (w.head == 0) red_add(a.b[w.layer] + ch, mine);
            else if (w.head == 1) { if (ch == 3) red_add(a.b[w.layer], mine); }      // g_raw channels: (r, g, b, sigma)
            else if (ch < 3) red_add(a.b[w.layer] + ch, mine);
          }
        }
      }
    }
  } else if (warp >= 8) {
    // ---- flush: accumulator rows -> red.global.add into the nn.Linear-layout gradient ------------------------
    if (nstages > 0) {
      wg_wait(base + L::done, 0, abort_addr, a.err, 0x53000000);
      tc_fence_after_sync();
      const int q = warp & 3;
      const uint32_t lane_addr = tmem + ((uint32_t)(q << 5) << 16);
      for (int hf = 0; hf < nhalves; ++hf) {
        int n_out = hf * 128 + q * 32 + lane;
        bool keep = true;                                            // (tcgen05.ld below is warp-wide: no early out)
        if (w.head == 1) { keep = (n_out == 3); n_out = 0; }         // alpha_linear.weight is [1, 256]
        if (w.head == 2) { keep = (n_out < 3); n_out = keep ? n_out : 0; }   // rgb_linear.weight is [3, 128]
        float* dst = a.w[w.layer] + (size_t)n_out * w.ld + w.col0;
        const bool vec4 = ((w.ld | w.col0) & 3) == 0 && (w.n_valid & 3) == 0 && (reinterpret_cast<uintptr_t>(a.w[w.layer]) & 15) == 0;
        for (int c0 = 0; c0 < w.n_blocks * 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(lane_addr + hf * 256 + c0, v);
          tmem_ld_wait();
          if (vec4) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              if (keep && c0 + i < w.n_valid)
                red_add4(dst + c0 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (keep && c0 + i < w.n_valid) red_add(dst + c0 + i, __uint_as_float(v[i]));
          }
        }
      }
      tc_fence_before_sync();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, kTmemCols);
}

__device__ __forceinline__ float stash_bf16(const uint8_t* tile, int blk, int row, int col) {
  const uint32_t off = (uint32_t)blk * kBlkBytes + stash_chunk_off((uint32_t)row, (uint32_t)(col >> 3)) + (uint32_t)(col & 7) * 2u;
  return __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(tile + off)) << 16);
}

// =========================================================================================================
// direction columns of views_linears.0: dW_v[j, 256 + c] = sum_r (sum_s g_hv[r,s,j]) enc4(viewdir_r)[c]
// =========================================================================================================
constexpr int kVdRays = 8;
__global__ void __launch_bounds__(128) viewdir_grad_kernel(const uint8_t* __restrict__ stash_g, const float* __restrict__ vd,
                                                           int64_t stride, int64_t R, int S, float* __restrict__ dw_views) {
  __shared__ float enc[kVdRays][28];
  const int j = threadIdx.x;
  float acc[27];
#pragma unroll
  for (int c = 0; c < 27; ++c) acc[c] = 0.f;
  const int64_t ngroups = (R + kVdRays - 1) / kVdRays;
  for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    __syncthreads();
    const int64_t r0 = grp * kVdRays;
    if (j < kVdRays * 3) {
      const int rr = j / 3, ax = j - rr * 3;
      const float x = (r0 + rr < R) ? __ldg(vd + (r0 + rr) * stride + ax) : 0.f;
      float sc[8];
      posenc_axis<4>(x, sc);
      enc[rr][ax] = x;
#pragma unroll
      for (int k = 0; k < 4; ++k) { enc[rr][3 + 6 * k + ax] = sc[2 * k]; enc[rr][6 + 6 * k + ax] = sc[2 * k + 1]; }
    }
    __syncthreads();
    for (int rr = 0; rr < kVdRays && r0 + rr < R; ++rr) {
      float gs = 0.f;
      const int64_t p0 = (r0 + rr) * S;
      for (int s = 0; s < S; ++s) {
        const int64_t p = p0 + s;
        gs += stash_bf16(stash_g + (size_t)(p >> 7) * kStashTileBytes, kGHv + (j >> 6), (int)(p & 127), j & 63);
      }
#pragma unroll
      for (int c = 0; c < 27; ++c) acc[c] = fmaf(gs, enc[rr][c], acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 27; ++c) red_add(dw_views + (size_t)j * 283 + 256 + c, acc[c]);
}

// =========================================================================================================
// host
// =========================================================================================================
static bool g_wg_init[64];
static std::mutex g_wg_mutex;

bool mlp_use_ts();  // mlp_aux.cu

static void make_items(WgItem* items, bool with_dir) {
  int n = 0;
  auto add = [&](int a_blk, int m_blocks, int b_blk, int n_blocks, int layer, int do_bias, int ld, int col0, int n_valid,
                 int head = 0) {
    WgItem w{};
    w.head = (uint8_t)head;
    w.a_blk = (uint8_t)a_blk; w.m_blocks = (uint8_t)m_blocks; w.b_blk = (uint8_t)b_blk; w.n_blocks = (uint8_t)n_blocks;
    w.layer = (uint8_t)layer; w.do_bias = (uint8_t)do_bias; w.ld = (uint16_t)ld; w.col0 = (uint16_t)col0;
    w.n_valid = (uint16_t)n_valid;
    items[n++] = w;
  };
  add(kGLayer0, 4, kHEnc, 1, 0, 1, 63, 0, 63);                                           // pts_linears.0
  for (int l = 1; l <= 7; ++l)                                                           // pts_linears.1-7
    add(kGLayer0 + 4 * l, 4, 4 * (l - 1), 4, l, 1, l == 5 ? 319 : 256, l == 5 ? 63 : 0, 256);
  add(kGLayer0 + 4 * 5, 4, kHEnc, 1, 5, 0, 319, 0, 63);                                  // skip columns of layer 5
  add(kGFeat, 4, 4 * 7, 4, LIN_FEATURE, 1, 256, 0, 256);                                 // feature_linear
  add(kGHv, 2, kHFeat, 4, LIN_VIEWS, 1, 283, 0, 256);                                    // views_linears.0[:, :256]
  add(kGRaw, 2, 4 * 7, 4, LIN_ALPHA, 1, 256, 0, 256, 1);                                 // alpha_linear
  add(kGRaw, 2, kHHv, 2, LIN_RGB, 1, 128, 0, 128, 2);                                    // rgb_linear
  add(kGHv, 2, kHDir, 1, LIN_VIEWS, 0, 283, 256, 27);                                    // views_linears.0[:, 256:283]
  // CTAs in proportion to the bytes each item streams per tile
  const int nitems = with_dir ? kWgItems : kWgItems - 1;   // the direction item needs the forward's dir stash block
  int cost[kWgItems], total = 0;
  for (int i = 0; i < nitems; ++i) { cost[i] = items[i].m_blocks + items[i].n_blocks; total += cost[i]; }
  // Every CTA of an item streams the same share of its tiles, and the kernel ends with its slowest CTA: minimise the
  // largest bytes-per-CTA over the items (greedy: the next CTA goes to the item whose CTAs carry the most).  Round 1
  // rounded proportional shares down and handed the remainder to the widest items, which left the direction item at
  // 0.75 units per CTA against a mean of 0.63 - the whole launch ran 19 % behind its mean (ncu: 79 % of the copy peak).
  int share[kWgItems];
  (void)total;
  for (int i = 0; i < nitems; ++i) share[i] = 1;
  for (int given = nitems; given < kNumSMs; ++given) {
    int best = 0;
    for (int i = 1; i < nitems; ++i)
      if (cost[i] * share[best] > cost[best] * share[i]) best = i;     // cost[i] / share[i] > cost[best] / share[best]
    ++share[best];
  }
  int c = 0;
  for (int i = 0; i < nitems; ++i) { items[i].cta_begin = (uint16_t)c; c += share[i]; items[i].cta_end = (uint16_t)c; }
  for (int i = nitems; i < kWgItems; ++i) { items[i].cta_begin = items[i].cta_end = 0xffff; }
}

}  // namespace gbn

using namespace gbn;

extern "C" size_t gbn_mlp_wgrad_workspace_bytes(int64_t R) { (void)R; return 256; }

extern "C" int gbn_mlp_backward_weights(const void* stash_h, const void* stash_g, const float* g_raw,
                                        const float* viewdirs, int64_t ray_stride, int64_t R, int S,
                                        void* const* grads, void* workspace, void* stream) {
  GBN_REQUIRE(R >= 0 && S >= 1, "mlp_backward_weights: bad sizes");
  if (R == 0) return GBN_OK;
  GBN_REQUIRE(stash_h && stash_g && g_raw && viewdirs && grads && workspace, "mlp_backward_weights: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  GBN_CUDA(cudaGetDevice(&dev));
  GBN_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  {
    std::lock_guard<std::mutex> lk(g_wg_mutex);
    if (!g_wg_init[dev]) {
      WgItem items[kWgItems];
      make_items(items, mlp_use_ts());
      GBN_CUDA(cudaMemcpyToSymbol(c_wg, items, sizeof(items), 0, cudaMemcpyHostToDevice));
      GBN_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgSmem::alloc));
      g_wg_init[dev] = true;
    }
  }
  WgArgs a{};
  a.stash_h = static_cast<const uint8_t*>(stash_h);
  a.stash_g = static_cast<const uint8_t*>(stash_g);
  for (int i = 0; i < GBN_NUM_LINEAR; ++i) {
    GBN_REQUIRE(grads[2 * i] && grads[2 * i + 1], "mlp_backward_weights: grads[%d] is null", 2 * i);
    a.w[i] = static_cast<float*>(grads[2 * i]);
    a.b[i] = static_cast<float*>(grads[2 * i + 1]);
  }
  const int64_t P = R * S;
  a.ntiles = (P + kTileRows - 1) / kTileRows;
  a.err = static_cast<int*>(workspace);
  GBN_CUDA(cudaMemsetAsync(a.err, 0, 256, st));
  wgrad_tc_kernel<<<kNumSMs, kWgThreads, WgSmem::alloc, st>>>(a);
  int rc = check_launch("wgrad_tc_kernel");
  if (rc != GBN_OK) return rc;
  if (mlp_use_ts()) return GBN_OK;   // direction columns came from the tensor-core item
  const int64_t groups = (R + kVdRays - 1) / kVdRays;
  const int vgrid = (int)(groups < 2 * kNumSMs ? groups : 2 * kNumSMs);
  viewdir_grad_kernel<<<vgrid, 128, 0, st>>>(a.stash_g, viewdirs, ray_stride, R, S, a.w[LIN_VIEWS]);
  return check_launch("viewdir_grad_kernel");
}
