// Layout shared by the weight pre-packer and the tcgen05 MLP kernel: the per-tile list of tensor-core jobs
// (one K-block of one layer against an N-row slab of its weight matrix: all 256 rows in bf16, 128-row halves in
// tf32 where a 256-row slab would not leave room for a ring) and the packed weight image.
//
// Network (run_nerf_helpers.py:75-129, D=8 W=256 skips=[4] use_viewdirs):
//   unit 0..7  pts_linears.i   (unit 5 reads [pts-encoding | h], the skip concat of helpers:111-112)
//   unit 8     feature_linear (no activation) + alpha_linear (N padded 1 -> 16)
//   unit 9     views_linears.0, feature part only (K = 256); the direction part W[:,256:283] * enc(dir) + b is
//              constant along a ray and is folded into a per-ray fp32 bias by view_bias_kernel
//   unit 10    rgb_linear (N padded 3 -> 16)
#pragma once

#include <stdint.h>

#include <vector>

namespace gbn {

constexpr int kTileRows = 128;        // points per CTA tile == UMMA M == TMEM lanes
constexpr int kBlkBytes = 128 * 128;  // one [128 rows x 128 B] swizzled K-block
constexpr int kNumUnits = 11;         // forward; the backward (dgrad) program has 10
constexpr int kMaxJobs = 160;
constexpr int kNumPlans = 3;          // 0 forward bf16, 1 forward tf32, 2 backward (dgrad) bf16
constexpr int kPlanBwd = 2;

// Training stash (bf16 only): per 128-point tile, 16 KB blocks of [128 points x 64 channels].  Inside a block:
//   [half (64 points)][chunk (8 channels = 16 B)][64 points][16 B]     (stash_chunk_off below)
// so that (i) a warp of the epilogue (32 consecutive points, thread == point) writes each chunk as 512 contiguous
// bytes straight from its registers - round 1/2 staged a 128-byte-swizzled row-major image in shared memory and bulk-stored
// it, which cost the forward 35 % - (ii) a 64-point half block is 8 contiguous KB for wgrad's bulk copies and
// IS the no-swizzle MN-major UMMA operand (core matrix = 8 points x 8 channels = 128 contiguous bytes, 128 B between
// core matrices along the points, 1 KB along the channels), and (iii) dgrad's gate reads are conflict-free.
//   H (written by the forward): h_l block b -> 4*l + b (l = 0..7) | feature -> 32..35 | hv -> 36,37 | enc -> 38 |
//                               per-point view-direction encoding (27 of 64 channels) -> 39
//   G (written by dgrad): g_hv -> 0,1 | g_feature -> 2..5 | g_l block b -> 6 + 4*l + b | g_raw (padded) -> 38
constexpr int kStashBlocks = 40;
// byte offset inside a stash block of the 16-byte chunk c (channels 8c .. 8c+7) of point r
__host__ __device__ constexpr uint32_t stash_chunk_off(uint32_t r, uint32_t c) {
  return (r >> 6) * 8192u + c * 1024u + (r & 63u) * 16u;
}
constexpr size_t kStashTileBytes = (size_t)kStashBlocks * kBlkBytes;
constexpr int kHFeat = 32, kHHv = 36, kHEnc = 38, kHDir = 39;
constexpr int kGHv = 0, kGFeat = 2, kGLayer0 = 6, kGRaw = 38;

enum : uint8_t { EPI_BIAS_RELU = 0, EPI_BIAS = 1, EPI_VBIAS_RELU = 2, EPI_OUT = 3, EPI_MASK = 4, EPI_PLAIN = 5 };
struct EpiUnit {        // what the epilogue warps do with a unit's accumulator
  uint8_t mode, nb;     // nb = 64-column (bf16) / 32-column (tf32) K-blocks produced
  uint8_t mask_blk;     // EPI_MASK: H-stash block of the activation whose sign gates the gradient
  uint8_t out_blk;      // stash block the result goes to (when a stash pointer is given), 0xff = none
  uint16_t bias_off;    // float offset into the bias block
  uint8_t no_act;       // result is not an MMA operand of a later unit: no shared-memory store, no hand-over
  uint8_t pad;
};

constexpr uint32_t kColX = 0, kColY = 256, kColAlpha = 256 + 128, kColRgb = 256 + 144;
constexpr uint32_t kTmemCols = 512;

enum : uint8_t { JF_WAIT_ACT = 1, JF_WAIT_ENC = 2, JF_FIRST = 4, JF_COMMIT_ACC = 8, JF_COMMIT_ENC = 16 };
constexpr uint8_t kEncBlkFlag = 0x80;

struct MlpJob {         // consumed by the TMA producer and the MMA issuer
  uint32_t w_off;       // byte offset of the weight chunk in the packed buffer (16-byte aligned)
  uint16_t w_bytes16;   // chunk bytes / 16  (N rows x 128 B)
  uint8_t a_blk;        // activation K-block index, or kEncBlkFlag | encoding K-block index
  uint8_t flags;
  uint16_t d_col;       // TMEM column of the accumulator
  uint8_t n16;          // N >> 4
  uint8_t unit;
  uint8_t ksteps;       // 32-byte MMA K-steps taken from the K-block (4 = all of it)
  uint8_t pad[3];
};
static_assert(sizeof(MlpJob) == 16, "MlpJob must stay 16 bytes");

struct PackJob {        // consumed by the pre-pack kernel: which slice of which nn.Linear weight fills a chunk
  uint32_t w_off;
  uint16_t layer;       // index into the 12 linears, order of gbn_mlp_prepack_weights
  uint16_t ld;          // in_features of that linear (row pitch)
  uint16_t row0, rows_valid, rows;   // rows = N of the chunk; rows >= rows_valid are zero
  uint16_t col0, cols_valid;         // K-block covers cols [col0, col0 + kb); only cols_valid of them exist
  uint8_t koff;                      // transposed slabs: slab column k reads source index k - koff
  uint8_t transpose;                 // 0: slab(n,k) = W[row0+n][col0+k]   1: slab(n,k) = W[row0+k-koff][col0+n]
};
static_assert(sizeof(PackJob) == 20, "PackJob layout");

// indices into the params array of gbn_mlp_prepack_weights (weight,bias pairs)
enum { LIN_PTS0 = 0, LIN_FEATURE = 8, LIN_ALPHA = 9, LIN_VIEWS = 10, LIN_RGB = 11 };

// bias block (fp32) inside the packed buffer / shared memory
constexpr int kBiasPts = 0;          // 8 x 256
constexpr int kBiasFeat = 2048;      // 256
constexpr int kBiasAlpha = 2304;     // 1 (+3 pad)
constexpr int kBiasRgb = 2308;       // 3 (+1 pad)
constexpr int kBiasFloats = 2312;

struct MlpPlan {
  int precision;
  int esz, kb, nblk, encb, nj;       // element bytes, K-block elements, act K-blocks, enc K-blocks, N per job
  std::vector<MlpJob> jobs;
  std::vector<PackJob> pack;
  int nunits;
  EpiUnit epi[kNumUnits];
  int unit_begin[kNumUnits + 1];
  uint32_t off_bias, off_wdir, off_bdir, total_bytes;
};

inline MlpPlan make_bwd_plan();

inline MlpPlan make_plan(int precision) {
  if (precision == kPlanBwd) return make_bwd_plan();
  MlpPlan p;
  p.precision = precision;
  p.nunits = kNumUnits;
  p.esz = precision == 0 ? 2 : 4;
  p.kb = 128 / p.esz;
  p.nblk = 256 / p.kb;
  p.encb = 64 / p.kb;
  p.nj = precision == 0 ? 256 : 128;
  const int nh = 256 / p.nj;
  uint32_t off = 256;  // header
  auto add = [&](int unit, int layer, int ld, int row0, int rows_valid, int rows, int col0, int cols_valid,
                 uint8_t a_blk, uint8_t flags, uint32_t d_col) {
    MlpJob j{};
    j.w_off = off;
    j.w_bytes16 = (uint16_t)(rows * 128 / 16);
    j.a_blk = a_blk;
    j.flags = flags;
    j.d_col = (uint16_t)d_col;
    j.n16 = (uint8_t)(rows / 16);
    j.unit = (uint8_t)unit;
    j.ksteps = 4;
    p.jobs.push_back(j);
    PackJob q{};
    q.w_off = off;
    q.layer = (uint16_t)layer;
    q.ld = (uint16_t)ld;
    q.row0 = (uint16_t)row0;
    q.rows_valid = (uint16_t)rows_valid;
    q.rows = (uint16_t)rows;
    q.col0 = (uint16_t)col0;
    q.cols_valid = (uint16_t)cols_valid;
    p.pack.push_back(q);
    off += (uint32_t)rows * 128;
  };
  auto clampc = [&](int col0, int ld) { int c = ld - col0; return c < 0 ? 0 : (c > p.kb ? p.kb : c); };
  for (int u = 0; u < kNumUnits; ++u) {
    p.unit_begin[u] = (int)p.jobs.size();
    const uint32_t dX = (u % 2 == 0) ? kColX : kColY;
    if (u == 0) {
      for (int e = 0; e < p.encb; ++e)
        for (int h = 0; h < nh; ++h)
          add(u, 0, 63, h * p.nj, p.nj, p.nj, e * p.kb, clampc(e * p.kb, 63), kEncBlkFlag | e,
              (uint8_t)((e == 0 && h == 0 ? JF_WAIT_ENC : 0) | (e == 0 ? JF_FIRST : 0)), dX + h * p.nj);
    } else if (u <= 7) {
      const int ld = (u == 5) ? 319 : 256, hoff = (u == 5) ? 63 : 0;
      if (u == 5)
        for (int e = 0; e < p.encb; ++e)
          for (int h = 0; h < nh; ++h)
            add(u, u, ld, h * p.nj, p.nj, p.nj, e * p.kb, clampc(e * p.kb, 63), kEncBlkFlag | e,
                (uint8_t)((e == 0 ? JF_FIRST : 0) | (e == p.encb - 1 && h == nh - 1 ? JF_COMMIT_ENC : 0)),
                dX + h * p.nj);
      for (int k = 0; k < p.nblk; ++k)
        for (int h = 0; h < nh; ++h)
          add(u, u, ld, h * p.nj, p.nj, p.nj, hoff + k * p.kb, p.kb, (uint8_t)k,
              (uint8_t)((h == 0 ? JF_WAIT_ACT : 0) | ((k == 0 && u != 5) ? JF_FIRST : 0)), dX + h * p.nj);
    } else if (u == 8) {
      for (int k = 0; k < p.nblk; ++k)
        for (int h = 0; h < nh; ++h)
          add(u, LIN_FEATURE, 256, h * p.nj, p.nj, p.nj, k * p.kb, p.kb, (uint8_t)k,
              (uint8_t)((h == 0 ? JF_WAIT_ACT : 0) | (k == 0 ? JF_FIRST : 0)), kColX + h * p.nj);
      for (int k = 0; k < p.nblk; ++k)
        add(u, LIN_ALPHA, 256, 0, 1, 16, k * p.kb, p.kb, (uint8_t)k, (uint8_t)(k == 0 ? JF_FIRST : 0), kColAlpha);
    } else if (u == 9) {
      for (int k = 0; k < p.nblk; ++k)
        add(u, LIN_VIEWS, 283, 0, 128, 128, k * p.kb, p.kb, (uint8_t)k,
            (uint8_t)(JF_WAIT_ACT | (k == 0 ? JF_FIRST : 0)), kColY);
    } else {
      for (int k = 0; k < 128 / p.kb; ++k)
        add(u, LIN_RGB, 128, 0, 3, 16, k * p.kb, p.kb, (uint8_t)k,
            (uint8_t)(JF_WAIT_ACT | (k == 0 ? JF_FIRST : 0)), kColRgb);
    }
    p.jobs.back().flags |= JF_COMMIT_ACC;
    EpiUnit e{};
    e.nb = (uint8_t)p.nblk;
    e.mask_blk = 0xff;
    if (u <= 7) { e.mode = EPI_BIAS_RELU; e.bias_off = (uint16_t)(u * 256); e.out_blk = (uint8_t)(4 * u); }
    else if (u == 8) { e.mode = EPI_BIAS; e.bias_off = kBiasFeat; e.out_blk = kHFeat; }
    else if (u == 9) { e.mode = EPI_VBIAS_RELU; e.nb = (uint8_t)(128 / p.kb); e.out_blk = kHHv; }
    else { e.mode = EPI_OUT; e.nb = 0; e.out_blk = 0xff; }
    p.epi[u] = e;
  }
  p.unit_begin[kNumUnits] = (int)p.jobs.size();
  p.off_bias = off;
  off += kBiasFloats * 4;
  p.off_wdir = off;  // views_linears.0.weight[:, 256:283] as fp32 [128][27]
  off += 128 * 27 * 4;
  p.off_bdir = off;  // views_linears.0.bias fp32 [128]
  off += 128 * 4;
  p.total_bytes = (off + 255) & ~255u;
  return p;
}

// Backward (dgrad) program, bf16: g_{l-1} = (g_l . W_l) * [h_{l-1} > 0], walked from the heads down to layer 1.
// Weight slabs are the TRANSPOSED linears (N = input index, K = output index).  Unit order:
//   0  g_hv   = g_rgb . W_rgb                      (A = padded g_raw block, K = 16)        gate hv
//   1  g_feat = g_hv . W_views[:, :256]            (K = 128)                               no gate
//   2  g_h7   = g_feat . W_feature + g_sigma w_a   (K = 256, + one K = 16 step on g_raw)   gate h7
//   3..9      g_{l-1} = g_l . W_l for l = 7..1     (K = 256; l = 5 skips its 63 encoding columns)
inline MlpPlan make_bwd_plan() {
  MlpPlan p;
  p.precision = kPlanBwd;
  p.esz = 2; p.kb = 64; p.nblk = 4; p.encb = 1; p.nj = 256;
  p.nunits = 10;
  uint32_t off = 256;
  auto add = [&](int unit, int layer, int ld, int row0, int rows_valid, int rows, int col0, int cols_valid, int koff,
                 uint8_t a_blk, uint8_t flags, uint32_t d_col, int ksteps) {
    MlpJob j{};
    j.w_off = off; j.w_bytes16 = (uint16_t)(rows * 128 / 16); j.a_blk = a_blk; j.flags = flags;
    j.d_col = (uint16_t)d_col; j.n16 = (uint8_t)(rows / 16); j.unit = (uint8_t)unit; j.ksteps = (uint8_t)ksteps;
    p.jobs.push_back(j);
    PackJob q{};
    q.w_off = off; q.layer = (uint16_t)layer; q.ld = (uint16_t)ld; q.row0 = (uint16_t)row0;
    q.rows_valid = (uint16_t)rows_valid; q.rows = (uint16_t)rows; q.col0 = (uint16_t)col0;
    q.cols_valid = (uint16_t)cols_valid; q.koff = (uint8_t)koff; q.transpose = 1;
    p.pack.push_back(q);
    off += (uint32_t)rows * 128;
  };
  for (int u = 0; u < p.nunits; ++u) {
    p.unit_begin[u] = (int)p.jobs.size();
    const uint32_t d = (u % 2 == 0) ? kColX : kColY;
    EpiUnit e{};
    e.nb = 4; e.mode = EPI_MASK; e.mask_blk = 0xff; e.out_blk = 0xff;
    if (u == 0) {
      add(u, LIN_RGB, 128, 0, 128, 128, 0, 3, 0, kEncBlkFlag, JF_WAIT_ENC | JF_FIRST, d, 1);
      e.nb = 2; e.mask_blk = kHHv; e.out_blk = kGHv;
    } else if (u == 1) {
      for (int k = 0; k < 2; ++k)
        add(u, LIN_VIEWS, 283, 64 * k, 256, 256, 0, 64, 0, (uint8_t)k, (uint8_t)(JF_WAIT_ACT | (k == 0 ? JF_FIRST : 0)), d, 4);
      e.mode = EPI_PLAIN; e.out_blk = kGFeat;
    } else if (u == 2) {
      for (int k = 0; k < 4; ++k)
        add(u, LIN_FEATURE, 256, 64 * k, 256, 256, 0, 64, 0, (uint8_t)k, (uint8_t)(JF_WAIT_ACT | (k == 0 ? JF_FIRST : 0)), d, 4);
      add(u, LIN_ALPHA, 256, 0, 256, 256, 0, 1, 3, kEncBlkFlag, JF_COMMIT_ENC, d, 1);
      e.mask_blk = 4 * 7; e.out_blk = (uint8_t)(kGLayer0 + 4 * 7);
    } else {
      const int l = 10 - u;  // 7..1
      const int ld = (l == 5) ? 319 : 256, c0 = (l == 5) ? 63 : 0;
      for (int k = 0; k < 4; ++k)
        add(u, l, ld, 64 * k, 256, 256, c0, 64, 0, (uint8_t)k, (uint8_t)(JF_WAIT_ACT | (k == 0 ? JF_FIRST : 0)), d, 4);
      e.mask_blk = (uint8_t)(4 * (l - 1)); e.out_blk = (uint8_t)(kGLayer0 + 4 * (l - 1));
      e.no_act = (l == 1);   // g_0 only goes to the stash (layer 0 has no data gradient to propagate)
    }
    p.jobs.back().flags |= JF_COMMIT_ACC;
    p.epi[u] = e;
  }
  for (int u = p.nunits; u <= kNumUnits; ++u) p.unit_begin[u] = (int)p.jobs.size();
  p.off_bias = off;
  off += kBiasFloats * 4;
  p.off_wdir = off; off += 128 * 27 * 4;
  p.off_bdir = off; off += 128 * 4;
  p.total_bytes = (off + 255) & ~255u;
  return p;
}

}  // namespace gbn
