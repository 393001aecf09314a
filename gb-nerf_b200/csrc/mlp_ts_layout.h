// Job tables of the TMEM-operand ("TS") bf16 MLP kernels (mlp_ts.cu): forward and backward-data programs.
//
// Why a second layout: with activations in shared memory every layer moves 64 KB (A reads) + 128 KB (B reads) +
// 128 KB (weight fills) + 64 KB (epilogue stores) through the 128 B/clk shared-memory port = 3072 cycles for 2048
// cycles of MMA.  Here the activations never leave tensor memory: the epilogue converts an accumulator half to bf16
// and writes it back with tcgen05.st as the NEXT layer's A operand (tcgen05.mma with A in TMEM), so shared memory
// carries the weight stream only.
//
// TMEM (512 columns): acc0 [0,128) | acc1 [128,256) | A0 [256,384) | A1 [384,512).  A buffer = 128 rows x 256 bf16
// (two K elements per 32-bit column).  A layer's 256 outputs are produced as two 128-wide halves (acc0, acc1); the
// half finished first is drained and converted while the MMAs of the other half run, and the next layer starts on
// K-blocks 0-1 as soon as half 0 is back in TMEM.
#pragma once

#include <stdint.h>

#include <vector>

#include "mlp_layout.h"

namespace gbn {

constexpr int kTsFwd = 3, kTsBwd = 4;          // plan ids (0..2 are the shared-memory-operand plans)
constexpr int kTsMaxJobs = 96, kTsMaxSteps = 24;
constexpr uint32_t kTsAcc0 = 0, kTsAcc1 = 128, kTsA0 = 256, kTsA1 = 384;
constexpr uint32_t kTsColAlpha = 0, kTsColRgb = 16;   // inside acc0, used after it has been drained

enum : uint16_t {
  TJ_WAIT_ENC = 1, TJ_WAIT_A0 = 2, TJ_WAIT_A1 = 4, TJ_WAIT_TILE = 8, TJ_FIRST = 16, TJ_COMMIT_ENC = 32,
  TJ_COMMIT_ACC0 = 64, TJ_COMMIT_ACC1 = 128, TJ_A_SMEM = 256,
  // stagger of the two issuing warps inside a wide layer: the acc1 warp starts only after the acc0 warp has ISSUED
  // its half, so half 0 finishes first and is drained/converted while half 1's MMAs run
  TJ_SIGNAL_ORDER = 512, TJ_WAIT_ORDER = 1024,
  // the view-direction encoding block in shared memory as A operand (direction columns of views_linears.0)
  TJ_A_DIR = 2048, TJ_WAIT_DIR = 4096, TJ_COMMIT_DIR = 8192,
  // acc1 has been read out by the epilogue (signalled right after its tcgen05.ld, ~600 cycles before the converted
  // activations are handed over): lets the first MMAs of the next layer's half 1 overlap the rest of that epilogue step
  TJ_WAIT_EMPTY1 = 16384,
  // the weight-ring stage of this job was last used by a job of the OTHER issuing warp: wait until that warp has seen
  // its fill land before testing the stage's full barrier (set by ts_plan() for the kernel's stage count)
  TJ_PREV_OTHER = 32768
};
constexpr uint8_t kTjNextOther = 8;   // TsJob::ksteps bit 3: the stage's next job is the other issuer's (publish progress)
constexpr int kTsBiasViews = kBiasFloats;            // b_views appended to the bias block (128 floats)
constexpr int kTsBiasFloats = kBiasFloats + 128;

struct TsJob {
  uint32_t w_off;       // weight slab offset in the packed buffer
  uint16_t w_bytes16;   // slab bytes / 16 = nkb * N * 8
  uint16_t flags;
  uint16_t d_col;       // accumulator column
  uint16_t a_col;       // TMEM column of the first K-block of A (ignored for TJ_A_SMEM)
  uint8_t n16;          // N >> 4
  uint8_t nkb;          // K-blocks (64 K each) in this job: 1 or 2
  uint8_t ksteps;       // bits 0-2: 16-wide MMA steps per K-block (4, or 1 for the padded g_raw block); bit 3: kTjNextOther;
                        // bits 4-6: index within the tile of the issue-order signal this job raises / waits for
  uint8_t wait_buf;     // bit 0: A buffer whose ready barriers TJ_WAIT_A0/A1 refer to; bits 1-3 / 4-6: which completion
                        // of that barrier within the tile the job waits for (half 0 / half 1), so that each of the
                        // two issuing warps derives the phase parity without having seen the other's waits
};
static_assert(sizeof(TsJob) == 16, "TsJob must stay 16 bytes");

struct TsStep {          // one accumulator half handled by the epilogue warps
  uint8_t acc;           // 0/1: which accumulator (and acc_full barrier)
  uint8_t mode;          // EPI_* of mlp_layout.h
  uint8_t out_buf;       // A buffer the bf16 result is stored to
  uint8_t out_half;      // half (128 channels) of that buffer -> a_ready[out_buf][out_half]
  uint8_t no_act;        // result only goes to the stash
  uint8_t mask_blk;      // EPI_MASK: first H-stash block of the gating activation
  uint8_t out_blk;       // first stash block of the result (0xff: none)
  uint8_t pad;
  uint16_t bias_off;     // float offset in the bias block
  uint16_t pad2;
};
static_assert(sizeof(TsStep) == 12, "TsStep layout");

struct TsPackJob {       // which slice of which nn.Linear fills a slab (rows x nkb K-blocks of 64)
  uint32_t w_off;
  uint16_t layer, ld;
  uint16_t row0, rows_valid, rows;
  uint16_t col0, cols_valid;   // slab column k (0 .. 64*nkb) exists iff 0 <= k - koff < cols_valid
  uint8_t nkb, koff, transpose, pad;
};
static_assert(sizeof(TsPackJob) == 24, "TsPackJob layout");

struct TsPlan {
  int id;
  int ready_per_tile[4];   // completions of a_ready[buf][half] per tile
  int order_per_tile;      // issue-order signals per tile
  int empty1_per_tile;     // acc1-drained signals per tile (forward program)
  std::vector<TsJob> jobs;
  std::vector<TsStep> steps;
  std::vector<TsPackJob> pack;
  uint32_t off_bias, off_wdir, off_bdir, total_bytes;
};

inline TsPlan make_ts_plan(int id, bool stagger = false, bool early_empty = true, bool khi_order = false) {
  TsPlan p;
  p.id = id;
  uint32_t off = 256;
  const bool tr = (id == kTsBwd);
  int done[4] = {0, 0, 0, 0};   // completions of a_ready[buf*2+half] emitted so far (steps are added in program order)
  int order = 0;                // issue-order signals emitted so far
  int empty1 = 0;               // acc1-drained signals emitted so far (one per epilogue step on acc1)
  // forward slab(n, k)  = W[row0 + n][col0 + k - koff]
  // backward slab(n, k) = W[row0 + k - koff][col0 + n]      (transposed: N = layer input, K = layer output)
  auto job = [&](int layer, int ld, int row0, int rows_valid, int rows, int col0, int cols_valid, int koff, int nkb,
                 int ksteps, int flags, uint32_t d_col, uint32_t a_col, int wait_buf) {
    TsJob j{};
    j.w_off = off; j.w_bytes16 = (uint16_t)(nkb * rows * 8); j.flags = (uint16_t)flags; j.d_col = (uint16_t)d_col;
    j.a_col = (uint16_t)a_col; j.n16 = (uint8_t)(rows / 16); j.nkb = (uint8_t)nkb; j.ksteps = (uint8_t)ksteps;
    const int s0 = done[wait_buf * 2] - 1, s1 = done[wait_buf * 2 + 1] - 1;
    j.wait_buf = (uint8_t)(wait_buf | ((s0 < 0 ? 0 : s0) << 1) | ((s1 < 0 ? 0 : s1) << 4) |
                           ((flags & TJ_WAIT_EMPTY1) ? (((empty1 - 1) & 1) << 7) : 0));
    p.jobs.push_back(j);
    TsPackJob q{};
    q.w_off = off; q.layer = (uint16_t)layer; q.ld = (uint16_t)ld; q.row0 = (uint16_t)row0; q.rows_valid = (uint16_t)rows_valid;
    q.rows = (uint16_t)rows; q.col0 = (uint16_t)col0; q.cols_valid = (uint16_t)cols_valid; q.nkb = (uint8_t)nkb;
    q.koff = (uint8_t)koff; q.transpose = tr ? 1 : 0;
    p.pack.push_back(q);
    off += (uint32_t)nkb * rows * 128;
  };
  auto step = [&](int acc, int mode, int out_buf, int out_half, int no_act, int mask_blk, int out_blk, int bias_off) {
    TsStep s{};
    s.acc = (uint8_t)acc; s.mode = (uint8_t)mode; s.out_buf = (uint8_t)out_buf; s.out_half = (uint8_t)out_half;
    s.no_act = (uint8_t)no_act; s.mask_blk = (uint8_t)mask_blk; s.out_blk = (uint8_t)out_blk; s.bias_off = (uint16_t)bias_off;
    if (!no_act && mode != EPI_OUT) ++done[out_buf * 2 + out_half];
    if (acc == 1 && mode != EPI_OUT) ++empty1;
    p.steps.push_back(s);
  };
  const uint32_t abuf[2] = {kTsA0, kTsA1}, accc[2] = {kTsAcc0, kTsAcc1};
  const int commit[2] = {TJ_COMMIT_ACC0, TJ_COMMIT_ACC1};
  // 256 -> 256 layer with its input in A buffer `in`: per output half two jobs (K-block pairs); `extra` = a trailing
  // shared-memory-operand job per half (the skip encoding / the g_sigma term), which then carries the commit
  auto wide = [&](int layer, int ld, int col0, int in, bool extra) {
    for (int h = 0; h < 2; ++h)
      for (int kp = 0; kp < 2; ++kp) {
        int fl = 0;
        // acc0 jobs and acc1 jobs are issued by two different warps: each waits for what IT needs.  Half 0 starts on
        // K-blocks 0-1 as soon as input half 0 is back (that also certifies acc0 drained) and needs input half 1
        // for K-blocks 2-3; half 1 needs input half 0 (operand) and input half 1 (= acc1 drained) before its first MMA
        if (h == 0) fl |= (kp == 0) ? TJ_WAIT_A0 : TJ_WAIT_A1;
        else if (!early_empty) fl |= (kp == 0) ? (TJ_WAIT_A0 | TJ_WAIT_A1) : 0;
        else fl |= (kp == 0) ? (TJ_WAIT_A0 | TJ_WAIT_EMPTY1) : TJ_WAIT_A1;   // acc1 is free before input half 1 is back
        if (kp == 0) fl |= TJ_FIRST;
        if (kp == 1 && !extra) fl |= commit[h];
        if (!tr) job(layer, ld, 128 * h, 128, 128, col0 + 128 * kp, 128, 0, 2, 4, fl, accc[h], abuf[in] + 64 * kp, in);
        else job(layer, ld, 128 * kp, 128, 128, col0 + 128 * h, 128, 0, 2, 4, fl, accc[h], abuf[in] + 64 * kp, in);
        // Both K-high jobs become ready at the same moment (input half 1 handed over).  Half 0's gates the next
        // epilogue step, half 1's has a whole epilogue step of slack: half 1's issuer lets half 0's go first, so that
        // it gets the tensor pipe to itself instead of sharing it.  (Experiment, GBNERF_TS_KHI_ORDER=1: measured 2-3 % slower.)
        if (!tr && khi_order && !stagger && kp == 1) {
          TsJob& jb = p.jobs.back();
          jb.flags |= (h == 0) ? TJ_SIGNAL_ORDER : TJ_WAIT_ORDER;
          jb.ksteps |= (uint8_t)((order & 7) << 4);
          if (h == 1) ++order;
        }
      }
  };

  if (id == kTsFwd) {
    for (int h = 0; h < 2; ++h)   // layer 0: A = the encoding block in shared memory (K = 64, 63 valid)
      job(0, 63, 128 * h, 128, 128, 0, 63, 0, 1, 4,
          TJ_A_SMEM | TJ_FIRST | TJ_WAIT_ENC | TJ_WAIT_TILE | commit[h], accc[h], 0, 0);
    for (int h = 0; h < 2; ++h) step(h, EPI_BIAS_RELU, 0, h, 0, 0xff, 2 * h, 0);
    for (int l = 1; l <= 7; ++l) {
      const int in = (l - 1) & 1;
      if (l != 5) {
        wide(l, 256, 0, in, false);
      } else {
        // reorder so that each half ends with its encoding job: emit the wide jobs, then splice
        const size_t first = p.jobs.size();
        wide(l, 319, 63, in, true);
        // jobs now: h0kp0, h0kp1, h1kp0, h1kp1 ; append enc jobs and rotate them into place
        std::vector<TsJob> jj(p.jobs.begin() + first, p.jobs.end());
        std::vector<TsPackJob> pp(p.pack.begin() + first, p.pack.end());
        p.jobs.resize(first); p.pack.resize(first);
        off = jj[0].w_off;
        for (int h = 0; h < 2; ++h) {
          for (int kp = 0; kp < 2; ++kp) {
            TsJob j = jj[2 * h + kp]; TsPackJob q = pp[2 * h + kp];
            j.w_off = off; q.w_off = off;
            p.jobs.push_back(j); p.pack.push_back(q);
            off += (uint32_t)j.nkb * (j.n16 * 16) * 128;
          }
          job(l, 319, 128 * h, 128, 128, 0, 63, 0, 1, 4, TJ_A_SMEM | commit[h] | TJ_COMMIT_ENC, accc[h], 0, 0);
        }
      }
      for (int h = 0; h < 2; ++h) step(h, EPI_BIAS_RELU, l & 1, h, 0, 0xff, 4 * l + 2 * h, 256 * l);
    }
    // feature_linear: h7 (buffer 1) -> feature (buffer 0), no activation
    wide(LIN_FEATURE, 256, 0, 1, false);
    for (int h = 0; h < 2; ++h) step(h, EPI_BIAS, 0, h, 0, 0xff, kHFeat + 2 * h, kBiasFeat);
    // alpha_linear on h7 (still intact in buffer 1) into 16 spare columns of acc0, once acc0 has been drained
    for (int kp = 0; kp < 2; ++kp)
      job(LIN_ALPHA, 256, 0, 1, 16, 128 * kp, 128, 0, 2, 4, kp == 0 ? (TJ_WAIT_A0 | TJ_FIRST) : 0, kTsAcc0 + kTsColAlpha,
          abuf[1] + 64 * kp, 0);
    // views_linears.0: the feature (K = 256) from TMEM + the 27 direction columns on the direction block in smem
    for (int kp = 0; kp < 2; ++kp)
      job(LIN_VIEWS, 283, 0, 128, 128, 128 * kp, 128, 0, 2, 4,
          kp == 0 ? (TJ_WAIT_A0 | (early_empty ? TJ_WAIT_EMPTY1 : TJ_WAIT_A1) | TJ_FIRST) : (early_empty ? TJ_WAIT_A1 : 0), kTsAcc1,
          abuf[0] + 64 * kp, 0);
    job(LIN_VIEWS, 283, 0, 128, 128, 256, 27, 0, 1, 2, TJ_A_SMEM | TJ_A_DIR | TJ_WAIT_DIR | TJ_COMMIT_DIR | TJ_COMMIT_ACC1, kTsAcc1,
        0, 0);
    step(1, EPI_BIAS_RELU, 1, 0, 0, 0xff, kHHv, kTsBiasViews);
    // rgb_linear on hv (buffer 1, K = 128)
    job(LIN_RGB, 128, 0, 3, 16, 0, 128, 0, 2, 4, TJ_WAIT_A0 | TJ_FIRST | TJ_COMMIT_ACC0, kTsAcc0 + kTsColRgb, abuf[1], 1);
    step(0, EPI_OUT, 0, 0, 1, 0xff, 0xff, 0);
  } else {
    // g_hv = g_rgb . W_rgb, gated by hv : A = padded g_raw block in shared memory, one 16-wide K step
    job(LIN_RGB, 128, 0, 128, 128, 0, 3, 0, 1, 1, TJ_A_SMEM | TJ_WAIT_ENC | TJ_WAIT_TILE | TJ_FIRST | TJ_COMMIT_ACC0, kTsAcc0, 0, 0);
    step(0, EPI_MASK, 0, 0, 0, kHHv, kGHv, 0);
    // g_feature = g_hv . W_views[:, :256] : K = 128 (buffer 0, half 0), no gate
    for (int h = 0; h < 2; ++h)
      job(LIN_VIEWS, 283, 0, 128, 128, 128 * h, 128, 0, 2, 4, TJ_WAIT_A0 | (h == 1 ? TJ_WAIT_TILE : 0) | TJ_FIRST | commit[h],
          accc[h], abuf[0], 0);
    for (int h = 0; h < 2; ++h) step(h, EPI_PLAIN, 1, h, 0, 0xff, kGFeat + 2 * h, 0);
    // g_h7 = g_feature . W_feature + g_sigma w_alpha, gated by h7
    {
      const size_t first = p.jobs.size();
      wide(LIN_FEATURE, 256, 0, 1, true);
      std::vector<TsJob> jj(p.jobs.begin() + first, p.jobs.end());
      std::vector<TsPackJob> pp(p.pack.begin() + first, p.pack.end());
      p.jobs.resize(first); p.pack.resize(first);
      off = jj[0].w_off;
      for (int h = 0; h < 2; ++h) {
        for (int kp = 0; kp < 2; ++kp) {
          TsJob j = jj[2 * h + kp]; TsPackJob q = pp[2 * h + kp];
          j.w_off = off; q.w_off = off;
          p.jobs.push_back(j); p.pack.push_back(q);
          off += (uint32_t)j.nkb * (j.n16 * 16) * 128;
        }
        job(LIN_ALPHA, 256, 0, 128, 128, 128 * h, 1, 3, 1, 1, TJ_A_SMEM | commit[h] | TJ_COMMIT_ENC | (h == 1 ? TJ_WAIT_ENC : 0),
            accc[h], 0, 0);
      }
    }
    for (int h = 0; h < 2; ++h) step(h, EPI_MASK, 0, h, 0, 4 * 7 + 2 * h, kGLayer0 + 4 * 7 + 2 * h, 0);
    for (int l = 7; l >= 1; --l) {
      const int in = (7 - l) & 1;
      wide(l, l == 5 ? 319 : 256, l == 5 ? 63 : 0, in, false);
      for (int h = 0; h < 2; ++h)
        step(h, EPI_MASK, in ^ 1, h, l == 1, 4 * (l - 1) + 2 * h, kGLayer0 + 4 * (l - 1) + 2 * h, 0);
    }
  }
  // issue-order stagger: in every run  [acc0 jobs ...][acc1 jobs ...]  of a unit with >= 2 jobs per half, the last acc0
  // job raises the signal and the first acc1 job waits for it
  for (size_t i = 0; stagger && i + 1 < p.jobs.size(); ++i) {
    const bool a0 = p.jobs[i].d_col < kTsAcc1, b1 = p.jobs[i + 1].d_col >= kTsAcc1;
    if (a0 && b1 && i >= 1 && p.jobs[i - 1].d_col < kTsAcc1 && (p.jobs[i + 1].flags & TJ_FIRST) &&
        i + 2 < p.jobs.size() && p.jobs[i + 2].d_col >= kTsAcc1 && !(p.jobs[i + 2].flags & TJ_FIRST)) {
      p.jobs[i].flags |= TJ_SIGNAL_ORDER;
      p.jobs[i + 1].flags |= TJ_WAIT_ORDER;
      p.jobs[i].ksteps |= (uint8_t)(order << 4);
      p.jobs[i + 1].ksteps |= (uint8_t)(order << 4);
      ++order;
    }
  }
  p.order_per_tile = order;
  p.empty1_per_tile = empty1;
  for (int i = 0; i < 4; ++i) p.ready_per_tile[i] = done[i];
  p.off_bias = off;
  off += kTsBiasFloats * 4;
  p.off_wdir = off; off += 128 * 27 * 4;
  p.off_bdir = off; off += 128 * 4;
  p.total_bytes = (off + 255) & ~255u;
  return p;
}

}  // namespace gbn
