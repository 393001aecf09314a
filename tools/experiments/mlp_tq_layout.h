// Job tables of the quarter-pipelined TMEM-operand bf16 MLP kernels (mlp_tq.cu).
//
// Same idea as mlp_ts_layout.h (activations live in tensor memory as the next layer's A operand), finer grain:
// a layer's 256 outputs are four 64-column accumulators.  tcgen05.ld drains ~64 B/clk/SM, so emptying a 128x256 fp32
// layer takes as long as computing it; with quarters the drains of one layer run back to back underneath the MMAs of
// the following quarters, and the next layer starts on K-block k as soon as quarter k is back in TMEM as bf16.
//
// TMEM (512 columns): acc q at [64q, 64q+64), q = 0..3 | A0 [256,384) | A1 [384,512)   (A buffer: 128 rows x 256 bf16)
#pragma once

#include <stdint.h>

#include <vector>

#include "mlp_layout.h"
#include "mlp_ts_layout.h"   // TsPackJob

namespace gbn {

constexpr int kTqFwd = 5, kTqBwd = 6;          // plan ids
constexpr int kTqMaxJobs = 64, kTqMaxSteps = 48;
constexpr uint32_t kTqA0 = 256, kTqA1 = 384;
constexpr uint32_t kTqColAlpha = 0, kTqColRgb = 16;   // inside acc 0 after the feature quarter 0 has been drained
constexpr int kTqBiasViews = kBiasFloats;             // b_views appended to the bias block (128 floats)
constexpr int kTqBiasFloats = kBiasFloats + 128;

enum : uint16_t {
  QJ_WAIT_ENC = 1, QJ_WAIT_TILE = 2, QJ_FIRST = 4, QJ_COMMIT_ENC = 8, QJ_COMMIT_ACC = 16, QJ_A_ENC = 32, QJ_A_DIR = 64,
  QJ_WAIT_DIR = 128, QJ_COMMIT_DIR = 256
};

struct TqJob {           // 24 bytes
  uint32_t w_off;        // weight slab offset in the packed buffer
  uint16_t w_bytes16;    // slab bytes / 16
  uint16_t flags;
  uint16_t d_col;        // accumulator column
  uint16_t a_col;        // TMEM column of the first K-block of A (TMEM-operand jobs)
  uint8_t n16;           // N >> 4
  uint8_t nkb;           // K-blocks (64 K each): 1, 2 or 4
  uint8_t ksteps;        // 16-wide MMA steps per K-block (4; 2 for the direction block; 1 for the padded g_raw block)
  uint8_t acc;           // accumulator quarter 0..3: selects the issuing warp (acc & 1) and the acc_full barrier
  // barrier waits this job performs itself, in order: bits 0-1 = before which K-block, bits 2-4 = ready barrier
  // (buffer*4 + quarter), bits 5-7 = which completion of that barrier within the tile (parity source); 0xff = none.
  // Redundant waits (already observed by the same issuing warp earlier in the tile) are removed when the plan is built.
  uint8_t waits[6];
  uint8_t pad[2];
};
static_assert(sizeof(TqJob) == 24, "TqJob layout");

enum : uint8_t { EPI_ALPHA = 6 };   // continues the EPI_* list of mlp_layout.h
struct TqStep {          // one accumulator quarter handled by one epilogue warpgroup (acc & 1)
  uint8_t acc;           // 0..3
  uint8_t mode;          // EPI_*
  uint8_t out_buf;       // A buffer the bf16 result goes to
  uint8_t out_q;         // quarter of that buffer -> ready[out_buf*4 + out_q]
  uint8_t no_act;        // result only goes to the stash
  uint8_t mask_blk;      // EPI_MASK: H-stash block of the gating activation
  uint8_t out_blk;       // stash block of the result (0xff: none)
  uint8_t pad;
  uint16_t bias_off;     // float offset in the bias block
  uint16_t pad2;
};
static_assert(sizeof(TqStep) == 12, "TqStep layout");

struct TqPlan {
  int id;
  int ready_per_tile[8];
  std::vector<TqJob> jobs;
  std::vector<TqStep> steps;
  std::vector<TsPackJob> pack;
  uint32_t off_bias, off_wdir, off_bdir, total_bytes;
};

inline TqPlan make_tq_plan(int id) {
  TqPlan p;
  p.id = id;
  uint32_t off = 256;
  const bool tr = (id == kTqBwd);
  int done[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // completions of ready[buf*4+q] emitted so far, in program order
  int seen[2][8];                           // per issuing warp: completions already observed in this tile
  for (int w = 0; w < 2; ++w) for (int i = 0; i < 8; ++i) seen[w][i] = 0;
  const uint32_t abuf[2] = {kTqA0, kTqA1};

  struct Need { int pos, bar; };            // "before K-block pos, ready barrier bar must have completed"
  // forward slab(n, k)  = W[row0 + n][col0 + k - koff];  backward slab(n, k) = W[row0 + k - koff][col0 + n]
  auto job = [&](int layer, int ld, int row0, int rows_valid, int rows, int col0, int cols_valid, int koff, int nkb, int ksteps,
                 int flags, int acc, uint32_t d_col, uint32_t a_col, std::vector<Need> needs) {
    TqJob j{};
    j.w_off = off; j.w_bytes16 = (uint16_t)(nkb * rows * 8); j.flags = (uint16_t)flags; j.d_col = (uint16_t)d_col;
    j.a_col = (uint16_t)a_col; j.n16 = (uint8_t)(rows / 16); j.nkb = (uint8_t)nkb; j.ksteps = (uint8_t)ksteps; j.acc = (uint8_t)acc;
    for (int i = 0; i < 6; ++i) j.waits[i] = 0xff;
    int nw = 0;
    const int w = acc & 1;
    for (const Need& nd : needs) {
      const int seq = done[nd.bar];            // the latest completion of that barrier so far (1-based)
      if (seq <= seen[w][nd.bar]) continue;    // this warp has already observed it
      seen[w][nd.bar] = seq;
      if (nw < 6) j.waits[nw++] = (uint8_t)((nd.pos & 3) | ((nd.bar & 7) << 2) | (((seq - 1) & 7) << 5));
    }
    p.jobs.push_back(j);
    TsPackJob q{};
    q.w_off = off; q.layer = (uint16_t)layer; q.ld = (uint16_t)ld; q.row0 = (uint16_t)row0; q.rows_valid = (uint16_t)rows_valid;
    q.rows = (uint16_t)rows; q.col0 = (uint16_t)col0; q.cols_valid = (uint16_t)cols_valid; q.nkb = (uint8_t)nkb;
    q.koff = (uint8_t)koff; q.transpose = tr ? 1 : 0;
    p.pack.push_back(q);
    off += (uint32_t)nkb * rows * 128;
  };
  auto step = [&](int acc, int mode, int out_buf, int out_q, int no_act, int mask_blk, int out_blk, int bias_off) {
    TqStep s{};
    s.acc = (uint8_t)acc; s.mode = (uint8_t)mode; s.out_buf = (uint8_t)out_buf; s.out_q = (uint8_t)out_q;
    s.no_act = (uint8_t)no_act; s.mask_blk = (uint8_t)mask_blk; s.out_blk = (uint8_t)out_blk; s.bias_off = (uint16_t)bias_off;
    if (!no_act && mode != EPI_OUT && mode != EPI_ALPHA) ++done[out_buf * 4 + out_q];
    p.steps.push_back(s);
  };
  // the K-blocks of a job reading all 256 channels of A buffer `in` into accumulator quarter q: K-block kb needs input
  // quarter kb back in TMEM, and nothing may be written before quarter q (which emptied this accumulator) is
  auto needs_wide = [&](int in, int q, int nkb) {
    std::vector<Need> v;
    for (int kb = 0; kb < nkb; ++kb) v.push_back({kb, in * 4 + (kb > q ? kb : q)});
    return v;
  };
  // 256 -> 256 layer, input in A buffer `in`: one job per output quarter; `extra` = trailing shared-memory-operand job
  auto wide = [&](int layer, int ld, int col0, int in, int extra_flags, int extra_ksteps, int x_layer, int x_ld, int x_koff,
                  int x_cols) {
    for (int q = 0; q < 4; ++q) {
      const int fl = QJ_FIRST | (extra_flags ? 0 : QJ_COMMIT_ACC);
      if (!tr) job(layer, ld, 64 * q, 64, 64, col0, 256, 0, 4, 4, fl, q, 64 * q, abuf[in], needs_wide(in, q, 4));
      else job(layer, ld, 0, 64, 64, col0 + 64 * q, 256, 0, 4, 4, fl, q, 64 * q, abuf[in], needs_wide(in, q, 4));
      if (extra_flags) {
        if (!tr) job(x_layer, x_ld, 64 * q, 64, 64, 0, x_cols, x_koff, 1, extra_ksteps, extra_flags | QJ_COMMIT_ACC, q, 64 * q, 0, {});
        else job(x_layer, x_ld, 0, 64, 64, 64 * q, x_cols, x_koff, 1, extra_ksteps, extra_flags | QJ_COMMIT_ACC, q, 64 * q, 0, {});
      }
    }
  };

  if (id == kTqFwd) {
    // layer 0: A = the encoding block in shared memory (K = 64, 63 valid)
    for (int q = 0; q < 4; ++q)
      job(0, 63, 64 * q, 64, 64, 0, 63, 0, 1, 4, QJ_A_ENC | QJ_FIRST | QJ_WAIT_ENC | QJ_WAIT_TILE | QJ_COMMIT_ACC, q, 64 * q, 0, {});
    for (int q = 0; q < 4; ++q) step(q, EPI_BIAS_RELU, 0, q, 0, 0xff, q, 0);
    for (int l = 1; l <= 7; ++l) {
      const int in = (l - 1) & 1;
      if (l != 5) wide(l, 256, 0, in, 0, 0, 0, 0, 0, 0);
      else wide(l, 319, 63, in, QJ_A_ENC | QJ_COMMIT_ENC, 4, l, 319, 0, 63);   // + the 63 skip columns on the encoding block
      for (int q = 0; q < 4; ++q) step(q, EPI_BIAS_RELU, l & 1, q, 0, 0xff, 4 * l + q, 256 * l);
    }
    // feature_linear: h7 (buffer 1) -> feature (buffer 0), no activation
    wide(LIN_FEATURE, 256, 0, 1, 0, 0, 0, 0, 0, 0);
    for (int q = 0; q < 4; ++q) step(q, EPI_BIAS, 0, q, 0, 0xff, kHFeat + q, kBiasFeat);
    // alpha_linear on h7 (still intact in buffer 1) into 16 columns of acc 0 once the feature quarter 0 has left it
    job(LIN_ALPHA, 256, 0, 1, 16, 0, 256, 0, 4, 4, QJ_FIRST | QJ_COMMIT_ACC, 0, kTqColAlpha, abuf[1], {{0, 0 * 4 + 0}});
    step(0, EPI_ALPHA, 0, 0, 1, 0xff, 0xff, 0);
    // views_linears.0: two 64-wide quarters into acc 2 / acc 3 (feature K = 256) + the direction block (K = 27 -> 32)
    for (int v = 0; v < 2; ++v) {
      std::vector<Need> nd;
      for (int kb = 0; kb < 4; ++kb) nd.push_back({kb, 0 * 4 + (kb > 2 + v ? kb : 2 + v)});
      job(LIN_VIEWS, 283, 64 * v, 64, 64, 0, 256, 0, 4, 4, QJ_FIRST, 2 + v, 64 * (2 + v), abuf[0], nd);
      job(LIN_VIEWS, 283, 64 * v, 64, 64, 256, 27, 0, 1, 2, QJ_A_DIR | QJ_WAIT_DIR | QJ_COMMIT_DIR | QJ_COMMIT_ACC, 2 + v,
          64 * (2 + v), 0, {});
    }
    for (int v = 0; v < 2; ++v) step(2 + v, EPI_BIAS_RELU, 1, v, 0, 0xff, kHHv + v, kTqBiasViews);
    // rgb_linear on hv (buffer 1, K = 128) into acc 0 next to alpha
    job(LIN_RGB, 128, 0, 3, 16, 0, 128, 0, 2, 4, QJ_FIRST | QJ_COMMIT_ACC, 0, kTqColRgb, abuf[1], {{0, 1 * 4 + 0}, {1, 1 * 4 + 1}});
    step(0, EPI_OUT, 0, 0, 1, 0xff, 0xff, 0);
  } else {
    // g_hv = g_rgb . W_rgb, gated by hv: A = padded g_raw block in shared memory, one 16-wide K step; two quarters
    for (int v = 0; v < 2; ++v)
      job(LIN_RGB, 128, 0, 64, 64, 64 * v, 3, 0, 1, 1, QJ_A_ENC | QJ_WAIT_ENC | QJ_WAIT_TILE | QJ_FIRST | QJ_COMMIT_ACC, v, 64 * v, 0, {});
    for (int v = 0; v < 2; ++v) step(v, EPI_MASK, 0, v, 0, kHHv + v, kGHv + v, 0);
    // g_feature = g_hv . W_views[:, :256]: K = 128 (buffer 0, quarters 0-1), no gate.  acc 2/3 are free since the
    // previous tile; acc 0/1 are free once g_hv quarter 0/1 has been drained
    for (int q = 0; q < 4; ++q) {
      std::vector<Need> nd;
      nd.push_back({0, 0 * 4 + (q == 1 ? 1 : 0)});
      nd.push_back({1, 0 * 4 + 1});
      job(LIN_VIEWS, 283, 0, 64, 64, 64 * q, 128, 0, 2, 4, QJ_FIRST | QJ_COMMIT_ACC | (q >= 2 ? QJ_WAIT_TILE : 0), q, 64 * q, abuf[0], nd);
    }
    for (int q = 0; q < 4; ++q) step(q, EPI_PLAIN, 1, q, 0, 0xff, kGFeat + q, 0);
    // g_h7 = g_feature . W_feature + g_sigma w_alpha, gated by h7
    wide(LIN_FEATURE, 256, 0, 1, QJ_A_ENC | QJ_COMMIT_ENC | QJ_WAIT_ENC, 1, LIN_ALPHA, 256, 3, 1);
    for (int q = 0; q < 4; ++q) step(q, EPI_MASK, 0, q, 0, 4 * 7 + q, kGLayer0 + 4 * 7 + q, 0);
    for (int l = 7; l >= 1; --l) {
      const int in = (7 - l) & 1;
      wide(l, l == 5 ? 319 : 256, l == 5 ? 63 : 0, in, 0, 0, 0, 0, 0, 0);
      for (int q = 0; q < 4; ++q) step(q, EPI_MASK, in ^ 1, q, l == 1, 4 * (l - 1) + q, kGLayer0 + 4 * (l - 1) + q, 0);
    }
  }
  for (int i = 0; i < 8; ++i) p.ready_per_tile[i] = done[i];
  p.off_bias = off;
  off += kTqBiasFloats * 4;
  p.off_wdir = off; off += 128 * 27 * 4;
  p.off_bdir = off; off += 128 * 4;
  p.total_bytes = (off + 255) & ~255u;
  return p;
}

}  // namespace gbn
