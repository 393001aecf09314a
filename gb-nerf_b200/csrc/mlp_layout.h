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
constexpr int kNumUnits = 11;
constexpr int kMaxJobs = 160;

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
  uint32_t pad;
};
static_assert(sizeof(MlpJob) == 16, "MlpJob must stay 16 bytes");

struct PackJob {        // consumed by the pre-pack kernel: which slice of which nn.Linear weight fills a chunk
  uint32_t w_off;
  uint16_t layer;       // index into the 12 linears, order of gbn_mlp_prepack_weights
  uint16_t ld;          // in_features of that linear (row pitch)
  uint16_t row0, rows_valid, rows;   // rows = N of the chunk; rows >= rows_valid are zero
  uint16_t col0, cols_valid;         // K-block covers cols [col0, col0 + kb); only cols_valid of them exist
  uint16_t pad;
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
  int unit_begin[kNumUnits + 1];
  uint32_t off_bias, off_wdir, off_bdir, total_bytes;
};

inline MlpPlan make_plan(int precision) {
  MlpPlan p;
  p.precision = precision;
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

}  // namespace gbn
