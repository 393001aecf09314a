// Two tiles in flight per CTA ("T2"): the forward of the bf16 TMEM-operand MLP (included by mlp_ts.cu, inside namespace
// gbn; same packed weight image, same inputs and outputs as nerf_mlp_ts_kernel<fwd>).  nerf_mlp_t2_kernel<1> serves
// inference, nerf_mlp_t2_kernel<1, true> the training forward: the same kernel with the H-stash stores compiled in
// (byte-identical to the one-tile kernel's stash, tests/test_gpu_mlp_t2.py); dgrad stays the one-tile program.
//
// Why: one 128-point tile per SM leaves the tensor pipe idle while the epilogue converts an accumulator half and the
// issuer waits for the hand-over - a ~2,700-cycle dependency chain per layer against 2,048 cycles of tensor work
// (DESIGN 3.1).  A second tile fills those gaps, but the four-region layout (acc0 | acc1 | A0 | A1) already takes all
// 512 TMEM columns for ONE tile.  Here a tile owns 256 columns:
//     slot s:  A_s = [256 s, 256 s + 128)   activations, 128 rows x 256 bf16, overwritten IN PLACE
//              ACC_s = [256 s + 128, 256 s + 256)   one 128-channel accumulator half
// A layer runs as: MMAs of output half 0 -> epilogue reads ACC, converts, KEEPS the 128 bf16 channels in registers ->
// MMAs of half 1 (ACC again) -> epilogue reads ACC, and - every MMA that reads A_s has completed by then - stores
// both halves over A_s.  Each slot has its own issuing warp and its own epilogue warpgroup, so the two tiles drift into
// anti-phase by themselves: one slot's MMAs run while the other's epilogue works.
//
//   warp 0      weight producer: every slab is fetched ONCE for both slots (half the L2 -> shared-memory stream)
//   warp 1, 3   MMA issuer of slot 0 / slot 1; both walk the same slab sequence, w_empty counts two commits
//   warp 2      TMEM allocator
//   warps 4-7   per-tile input blocks (positional + direction encoding) of both slots
//   warps 8-11  epilogue of slot 0 (thread == row, 128 channels per step);  warps 12-15: slot 1
// alpha_linear (256 -> 1) and rgb_linear (128 -> 3) run on the CUDA cores inside the epilogue that produces their
// input (same bf16 operands and fp32 accumulation as the tiny MMAs of the one-tile kernel; 640 FMAs per point), which
// removes their accumulator columns and two hand-overs per tile.

constexpr int kT2MaxJobs = 48, kT2MaxSteps = 24;
// -DGBN_T2_EXP (csrc/build.py --exp -> libgbnerf_exp.so, selected with GBNERF_LIB): timing experiments of DESIGN 3.1.1 -
// GBNERF_T2_TURN_BACK (where inside a group the turn passes), GBNERF_T2_ABL (ablations that break the results: 1 = no rgb
// head, 2 = no sin/cos in the input warps, 8 = OUT does nothing after reading the accumulator, 16 = FLUSH hands over
// before it converts half 1), GBNERF_T2_DBG_EXTRA (repeat a wide layer n times).  The product build
// carries none of it.
#ifdef GBN_T2_EXP
constexpr bool kT2Exp = true;
#else
constexpr bool kT2Exp = false;
#endif
enum : uint16_t {
  T2_WAIT_ENC = 1, T2_WAIT_DIR = 2, T2_WAIT_A = 4, T2_WAIT_EMPTY = 8, T2_FIRST = 16, T2_A_ENC = 32, T2_A_DIR = 64,
  T2_COMMIT_ACC = 128, T2_COMMIT_ENC = 256, T2_COMMIT_DIR = 512, T2_TILE_FIRST = 1024
};
struct T2Job {
  uint32_t w_off;       // slab offset in the packed image (the kTsFwd image, unchanged)
  uint16_t bytes16;     // slab bytes / 16
  uint16_t flags;
  uint16_t a_col;       // column of the first K-block of A inside the slot's A region
  uint8_t ksteps;       // shared-memory A operand: 16-wide K steps (4 encoding, 2 direction)
  uint8_t nkb;
  uint8_t glen;         // first job of an MMA group (the jobs of one accumulator half): jobs in the group, else 0
  uint8_t pad;
  uint16_t gflags;      // first job of a group: the wait flags of all its jobs
};
static_assert(sizeof(T2Job) == 16, "T2Job layout");
enum : uint8_t { T2_HOLD = 0, T2_FLUSH = 1, T2_OUT = 2 };
struct T2Step {
  uint8_t mode, relu, dot, job0;  // dot: this layer's output feeds alpha_linear (sigma accumulates in the epilogue);
  uint16_t bias_off, out_blk;     // job0: first job of the step's MMA group (= jobs of the tile before this group);
};                                // out_blk: first H-stash block of the step's output (training forward, mlp_layout.h)
static_assert(sizeof(T2Step) == 8, "T2Step layout");
__constant__ T2Job c_t2jobs[kT2MaxJobs];
__constant__ T2Step c_t2steps[kT2MaxSteps];

struct T2Smem {
  static constexpr int NST = 4;
  static constexpr uint32_t enc = 0;                                   // [2] encoding block per slot
  static constexpr uint32_t dir = enc + 2 * kBlkBytes;                 // [2] direction block per slot
  static constexpr uint32_t ring = dir + 2 * kBlkBytes;
  static constexpr uint32_t bias = ring + NST * kTsStageBytes;
  static constexpr uint32_t walpha = bias + kTsBiasFloats * 4;         // 256 floats
  static constexpr uint32_t wrgb = walpha + 256 * 4;                   // 128 x float4 (r, g, b, 0)
  static constexpr uint32_t red = wrgb + 128 * 16;                     // [tile parity][slot][row] float4: (r, g, b, sigma) partial sums of wg 1
  static constexpr uint32_t bars = red + 4 * 128 * 16;                 // (modes 0 / 1 use the first half)
  static constexpr uint32_t w_full = bars;
  static constexpr uint32_t w_empty = w_full + 8 * NST;
  static constexpr uint32_t acc_full = w_empty + 8 * NST;              // [slot]
  static constexpr uint32_t acc_empty = acc_full + 16;
  static constexpr uint32_t a_ready = acc_empty + 16;
  static constexpr uint32_t enc_full = a_ready + 16;
  static constexpr uint32_t enc_empty = enc_full + 16;
  static constexpr uint32_t dir_full = enc_empty + 16;
  static constexpr uint32_t dir_empty = dir_full + 16;
  static constexpr uint32_t lead = dir_empty + 16;                     // u32 pad | u32: MMA groups whose issue has been released (turn)
  static constexpr uint32_t tmem_ptr = lead + 8;
  static constexpr uint32_t abort_flag = tmem_ptr + 4;
  static constexpr uint32_t total = abort_flag + 4;
  static constexpr uint32_t alloc = total + 1024;
};
static_assert(T2Smem::alloc <= 232448, "shared memory budget");
static_assert((T2Smem::bars & 7) == 0 && (T2Smem::wrgb & 15) == 0 && (T2Smem::bias & 15) == 0, "alignment");

struct T2Args {
  const uint8_t* packed;
  const float* ro; const float* rd; const float* z; const float* pts; const float* emb; const float* vd;
  float* raw;
  int* err;
  int64_t stride, P;
  int S, njobs, nsteps;
  uint32_t off_alpha[2], off_rgb;
  unsigned long long* trace;   // optional clock64 timeline of CTA 0, pair iteration trace_it (tools/t2_trace.py):
  int mode;                    // kernel variant (GBNERF_T2_MODE): see the MODE template parameter
  int trace_it;                // [(slot*48 + job)*4 ..] issuer: start, operands ready, weights landed, issued |
                               // [512 + (slot*24 + step)*4 ..] epilogue: wait start, accumulator ready, handed over, step done
  int turn_back, abl;          // GBN_T2_EXP builds only
  int stagger;                 // MODE 2: MMA groups by which slot 1 runs behind slot 0
  uint8_t* stash;              // STASH kernels: the H stash of the training forward (kStashTileBytes per tile)
  int64_t ntiles;
};

// the 8 x 16-byte chunks a thread holds of one stash block (64 channels of its point): w = 32 bf16x2 words, or a part
template <int NCHUNK>
__device__ __forceinline__ void t2_stash_chunks(uint8_t* blk_row, int chunk0, const uint32_t* w) {
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c)
    ts_st_global16(blk_row + (uint32_t)(chunk0 + c) * 1024u, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

__device__ __forceinline__ float t2_bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float t2_bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// 32 accumulator columns + bias (-> ReLU) -> 16 bf16x2 words
template <bool RELU>
__device__ __forceinline__ void t2_convert(const uint32_t (&v)[32], const float* bias32, uint32_t* w) {
  const float4* bp = reinterpret_cast<const float4*>(bias32);
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 bb = bp[i >> 2];
    float f0, f1, f2, f3;
    add_f32x2(v[i], v[i + 1], bb.x, bb.y, f0, f1);
    add_f32x2(v[i + 2], v[i + 3], bb.z, bb.w, f2, f3);
    w[i >> 1] = RELU ? pack_bf16_relu(f0, f1) : pack_bf16(f0, f1);
    w[(i >> 1) + 1] = RELU ? pack_bf16_relu(f2, f3) : pack_bf16(f2, f3);
  }
}
// sigma += <32 bf16 channels, 32 fp32 weights>
__device__ __forceinline__ float t2_dot32(const uint32_t* w, const float* wa32, float acc) {
  const float4* ap = reinterpret_cast<const float4*>(wa32);
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float4 a = ap[i >> 1];
    acc = fmaf(t2_bf16_lo(w[i]), a.x, acc);
    acc = fmaf(t2_bf16_hi(w[i]), a.y, acc);
    acc = fmaf(t2_bf16_lo(w[i + 1]), a.z, acc);
    acc = fmaf(t2_bf16_hi(w[i + 1]), a.w, acc);
  }
  return acc;
}
// rgb += hv[32 channels] . W_rgb^T   (wrgb: float4 (r, g, b, 0) per input channel)
__device__ __forceinline__ void t2_rgb32(const uint32_t* w, const float4* wr, float& r, float& g, float& b) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float lo = t2_bf16_lo(w[i]), hi = t2_bf16_hi(w[i]);
    const float4 a = wr[2 * i], c = wr[2 * i + 1];
    r = fmaf(lo, a.x, r); g = fmaf(lo, a.y, g); b = fmaf(lo, a.z, b);
    r = fmaf(hi, c.x, r); g = fmaf(hi, c.y, g); b = fmaf(hi, c.z, b);
  }
}

template <int SLOT>
__device__ __noinline__ void t2_issue_loop(int* err, int njobs, uint32_t base, uint32_t abort_addr, uint32_t tmem, int my_pairs,
                                           unsigned long long* trace, int trace_it) {
  using L = T2Smem;
  const uint64_t adesc_enc = smem_desc_sw128(base + L::enc + SLOT * kBlkBytes);
  const uint64_t adesc_dir = smem_desc_sw128(base + L::dir + SLOT * kBlkBytes);
  const uint64_t ring_desc0 = smem_desc_sw128(base + L::ring);
  const uint32_t idesc = make_idesc(1, 128, 128);
  const uint32_t d = tmem + 256u * SLOT + 128u, a_base = tmem + 256u * SLOT;
  const uint32_t b_acc_full = base + L::acc_full + 8 * SLOT, b_acc_empty = base + L::acc_empty + 8 * SLOT;
  const uint32_t b_a_ready = base + L::a_ready + 8 * SLOT;
  const uint32_t b_enc_full = base + L::enc_full + 8 * SLOT, b_enc_empty = base + L::enc_empty + 8 * SLOT;
  const uint32_t b_dir_full = base + L::dir_full + 8 * SLOT, b_dir_empty = base + L::dir_empty + 8 * SLOT;
  uint32_t s = 0, par = 0, ph_a = 0, ph_e = 0;
  for (int it = 0; it < my_pairs; ++it) {
    for (int j = 0; j < njobs; ++j) {
      const T2Job rc = c_t2jobs[j];
      const uint32_t fl = rc.flags;
      unsigned long long* tr = (kDiag && trace && blockIdx.x == 0 && it == trace_it && (threadIdx.x & 31) == 0) ? trace + (SLOT * 48 + j) * 4 : nullptr;
      if (tr) tr[0] = clock64();
      if (fl & T2_WAIT_ENC) ts_wait(b_enc_full, (uint32_t)it & 1u, abort_addr, err, 0x72000000 | (SLOT << 16) | j);
      if (fl & T2_WAIT_DIR) ts_wait(b_dir_full, (uint32_t)it & 1u, abort_addr, err, 0x73000000 | (SLOT << 16) | j);
      if (fl & T2_WAIT_A) { ts_wait(b_a_ready, ph_a, abort_addr, err, 0x74000000 | (SLOT << 16) | j); ph_a ^= 1u; }
      if ((fl & T2_WAIT_EMPTY) && !((fl & T2_TILE_FIRST) && it == 0)) {
        ts_wait(b_acc_empty, ph_e, abort_addr, err, 0x75000000 | (SLOT << 16) | j);
        ph_e ^= 1u;
      }
      if (tr) tr[1] = clock64();
      ts_wait(base + L::w_full + 8 * s, par, abort_addr, err, 0x76000000 | (SLOT << 16) | j);
      if (tr) tr[2] = clock64();
      tc_fence_after_sync();
      const uint64_t bd0 = ring_desc0 + (uint64_t)(s * (kTsStageBytes >> 4));
      const uint64_t bd1 = bd0 + 1024u;        // second K-block image: 128 rows x 128 B further on
      const uint32_t first = (fl & T2_FIRST) ? 0u : 1u;
      if (elect_one()) {
        if (!(fl & (T2_A_ENC | T2_A_DIR))) {   // 8 MMAs, A = two K-blocks (64 K each) of the slot's activations in TMEM
          const uint32_t a_t = a_base + rc.a_col;
          umma_bf16_ts(d, a_t, bd0, idesc, first);
          umma_bf16_ts(d, a_t + 8, bd0 + 2, idesc, 1u);
          umma_bf16_ts(d, a_t + 16, bd0 + 4, idesc, 1u);
          umma_bf16_ts(d, a_t + 24, bd0 + 6, idesc, 1u);
          umma_bf16_ts(d, a_t + 32, bd1, idesc, 1u);
          umma_bf16_ts(d, a_t + 40, bd1 + 2, idesc, 1u);
          umma_bf16_ts(d, a_t + 48, bd1 + 4, idesc, 1u);
          umma_bf16_ts(d, a_t + 56, bd1 + 6, idesc, 1u);
        } else {
          const uint64_t adesc = (fl & T2_A_DIR) ? adesc_dir : adesc_enc;
          umma_bf16(d, adesc, bd0, idesc, first);
          umma_bf16(d, adesc + 2, bd0 + 2, idesc, 1u);
          if (rc.ksteps == 4) {
            umma_bf16(d, adesc + 4, bd0 + 4, idesc, 1u);
            umma_bf16(d, adesc + 6, bd0 + 6, idesc, 1u);
          }
        }
        umma_commit(base + L::w_empty + 8 * s);
        if (fl & T2_COMMIT_ACC) umma_commit(b_acc_full);
        if (fl & T2_COMMIT_ENC) umma_commit(b_enc_empty);
        if (fl & T2_COMMIT_DIR) umma_commit(b_dir_empty);
      }
      __syncwarp();
      if (tr) tr[3] = clock64();
      if (++s == (uint32_t)L::NST) { s = 0; par ^= 1u; }
    }
  }
}

// Mode 2: a group's MMAs (16 for a wide layer half) are issued back to back after ALL of the group's waits - operands,
// the turn, every weight stage of the group - and the turn passes to the other issuer just before the group's last job
// is issued, so that the other issuer's wake-up and probes run in the shadow of those MMAs.  Global group order:
// slot 0 group 0, slot 1 group 0, slot 0 group 1, ...
template <int SLOT>
__device__ __noinline__ void t2_issue_groups(int* err, int njobs, uint32_t base, uint32_t abort_addr, uint32_t tmem, int my_pairs,
                                             unsigned long long* trace, int trace_it, int turn_back) {
  using L = T2Smem;
  const uint64_t adesc_enc = smem_desc_sw128(base + L::enc + SLOT * kBlkBytes);
  const uint64_t adesc_dir = smem_desc_sw128(base + L::dir + SLOT * kBlkBytes);
  const uint64_t ring_desc0 = smem_desc_sw128(base + L::ring);
  const uint32_t idesc = make_idesc(1, 128, 128);
  const uint32_t d = tmem + 256u * SLOT + 128u, a_base = tmem + 256u * SLOT;
  const uint32_t b_acc_full = base + L::acc_full + 8 * SLOT, b_acc_empty = base + L::acc_empty + 8 * SLOT;
  const uint32_t b_a_ready = base + L::a_ready + 8 * SLOT;
  const uint32_t b_enc_full = base + L::enc_full + 8 * SLOT, b_enc_empty = base + L::enc_empty + 8 * SLOT;
  const uint32_t b_dir_full = base + L::dir_full + 8 * SLOT, b_dir_empty = base + L::dir_empty + 8 * SLOT;
  const uint32_t turn_addr = base + L::lead + 4;
  uint32_t s = 0, par = 0, ph_a = 0, ph_e = 0, g = (uint32_t)SLOT;
  for (int it = 0; it < my_pairs; ++it) {
    int j = 0;
    while (j < njobs) {
      const T2Job r0 = c_t2jobs[j];
      const uint32_t fl = r0.gflags;
      const int glen = r0.glen;
      unsigned long long* tr = (kDiag && trace && blockIdx.x == 0 && it == trace_it && (threadIdx.x & 31) == 0) ? trace + (SLOT * 48 + j) * 4 : nullptr;
      if (tr) tr[0] = clock64();
      // The group's weight stages first: they landed long ago (the ring runs a group ahead), and probing them here takes
      // their round trips off the path between the operands' hand-over and the first MMA.
      {
        uint32_t ss = s, pp = par;
        for (int k = 0; k < glen; ++k) {
          ts_wait(base + L::w_full + 8 * ss, pp, abort_addr, err, 0x76000000 | (SLOT << 16) | (j + k));
          if (++ss == (uint32_t)L::NST) { ss = 0; pp ^= 1u; }
        }
      }
      if (fl & T2_WAIT_ENC) ts_wait(b_enc_full, (uint32_t)it & 1u, abort_addr, err, 0x72000000 | (SLOT << 16) | j);
      if (fl & T2_WAIT_DIR) ts_wait(b_dir_full, (uint32_t)it & 1u, abort_addr, err, 0x73000000 | (SLOT << 16) | j);
      if (fl & T2_WAIT_A) { ts_wait(b_a_ready, ph_a, abort_addr, err, 0x74000000 | (SLOT << 16) | j); ph_a ^= 1u; }
      if ((fl & T2_WAIT_EMPTY) && !((fl & T2_TILE_FIRST) && it == 0)) {
        ts_wait(b_acc_empty, ph_e, abort_addr, err, 0x75000000 | (SLOT << 16) | j);
        ph_e ^= 1u;
      }
      if (tr) tr[1] = clock64();
      ts_wait_progress(turn_addr, g, abort_addr, err, 0x77000000 | (SLOT << 16) | j);
      if (tr) tr[2] = clock64();
      tc_fence_after_sync();
      if (elect_one()) {
        uint32_t ss = s;
        const int store_at = kT2Exp ? (2 * glen - turn_back > 0 ? 2 * glen - turn_back : 0) : 2 * (glen - 1);   // in half-jobs
        for (int k = 0; k < glen; ++k) {
          const T2Job rc = c_t2jobs[j + k];   // (fetching the group's records ahead of the waits: +3.7 % cycles)
          const uint32_t f = rc.flags;
          if (2 * k == store_at) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(turn_addr), "r"(g + 1u) : "memory");
          const uint64_t bd0 = ring_desc0 + (uint64_t)(ss * (kTsStageBytes >> 4));
          const uint64_t bd1 = bd0 + 1024u;
          const uint32_t first = (f & T2_FIRST) ? 0u : 1u;
          if (!(f & (T2_A_ENC | T2_A_DIR))) {
            const uint32_t a_t = a_base + rc.a_col;
            umma_bf16_ts(d, a_t, bd0, idesc, first);
            umma_bf16_ts(d, a_t + 8, bd0 + 2, idesc, 1u);
            umma_bf16_ts(d, a_t + 16, bd0 + 4, idesc, 1u);
            umma_bf16_ts(d, a_t + 24, bd0 + 6, idesc, 1u);
            if (kT2Exp && 2 * k + 1 == store_at) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(turn_addr), "r"(g + 1u) : "memory");
            umma_bf16_ts(d, a_t + 32, bd1, idesc, 1u);
            umma_bf16_ts(d, a_t + 40, bd1 + 2, idesc, 1u);
            umma_bf16_ts(d, a_t + 48, bd1 + 4, idesc, 1u);
            umma_bf16_ts(d, a_t + 56, bd1 + 6, idesc, 1u);
          } else {
            const uint64_t adesc = (f & T2_A_DIR) ? adesc_dir : adesc_enc;
            umma_bf16(d, adesc, bd0, idesc, first);
            umma_bf16(d, adesc + 2, bd0 + 2, idesc, 1u);
            if (kT2Exp && 2 * k + 1 == store_at) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(turn_addr), "r"(g + 1u) : "memory");
            if (rc.ksteps == 4) {
              umma_bf16(d, adesc + 4, bd0 + 4, idesc, 1u);
              umma_bf16(d, adesc + 6, bd0 + 6, idesc, 1u);
            }
          }
          umma_commit(base + L::w_empty + 8 * ss);
          if (f & T2_COMMIT_ACC) umma_commit(b_acc_full);
          if (f & T2_COMMIT_ENC) umma_commit(b_enc_empty);
          if (f & T2_COMMIT_DIR) umma_commit(b_dir_empty);
          if (++ss == (uint32_t)L::NST) ss = 0;
        }
      }
      __syncwarp();
      if (tr) tr[3] = clock64();
      for (int k = 0; k < glen; ++k)
        if (++s == (uint32_t)L::NST) { s = 0; par ^= 1u; }
      j += glen;
      g += 2u;
    }
  }
}

// ---- MODE 2, "staggered": slot 1 runs D MMA groups (about half a network) behind slot 0 ------------------------------
// With both slots in lock-step (MODE 1) they cross the tile boundary together: the view layer, OUT and layer 0 are short
// MMA groups (18, 4 and 4 MMAs) that cannot hide an epilogue step, so the tensor pipe idles for ~16,000 of a pair's
// ~59,000 cycles there (tools/t2_exp.py: 4,840 cycles per wide layer of both slots = 85 % busy, the rest is the boundary).
// Staggered, one slot crosses the boundary while the other is in its wide layers.  The merged order of MMA groups -
// issue turns, weight ring and epilogue service all follow it - is, for i = 0, 1, 2, ...:
//     slot 0's group i (if it has one)      then      slot 1's group i - D (if i >= D)
// Each slot fetches its own weights now (one consumer per ring stage, w_empty counts one commit): the ring position of a
// group's first job is the number of jobs the producer emitted before it, which both sides compute from the same rule.
// An issuer probes its weight stages AFTER it has the turn: every earlier group in the merged order has then finished
// its own probes, so the previous fill of each stage has landed and the parity test cannot pass on a stale phase (the
// two-consumer ring alias of DESIGN 3.1).
struct T2Cnt { int q, r; };   // a group count n = q * G + r (q tiles, r groups)
__device__ __forceinline__ void t2_cnt_inc(T2Cnt& c, int G) { if (++c.r == G) { c.r = 0; ++c.q; } }

template <int SLOT>
__device__ __noinline__ void t2_issue_stag(int* err, int njobs, int G, int D, uint32_t base, uint32_t abort_addr, uint32_t tmem,
                                           int my_pairs, int abl) {
  using L = T2Smem;
  const uint64_t adesc_enc = smem_desc_sw128(base + L::enc + SLOT * kBlkBytes);
  const uint64_t adesc_dir = smem_desc_sw128(base + L::dir + SLOT * kBlkBytes);
  const uint64_t ring_desc0 = smem_desc_sw128(base + L::ring);
  const uint32_t idesc = make_idesc(1, 128, 128);
  const uint32_t d = tmem + 256u * SLOT + 128u, a_base = tmem + 256u * SLOT;
  const uint32_t b_acc_full = base + L::acc_full + 8 * SLOT, b_acc_empty = base + L::acc_empty + 8 * SLOT;
  const uint32_t b_a_ready = base + L::a_ready + 8 * SLOT;
  const uint32_t b_enc_full = base + L::enc_full + 8 * SLOT, b_enc_empty = base + L::enc_empty + 8 * SLOT;
  const uint32_t b_dir_full = base + L::dir_full + 8 * SLOT, b_dir_empty = base + L::dir_empty + 8 * SLOT;
  const uint32_t turn_addr = base + L::lead + 4;
  const int N0 = my_pairs * G;          // groups per slot
  T2Cnt own{0, 0};
  int m = SLOT == 0 ? 0 : (D + 1 < N0 ? D + 1 : N0);   // the other slot's groups that precede this one in the merged order
  T2Cnt oth{m / G, m % G};
  uint32_t ph_a = 0, ph_e = 0;
  for (int n = 0; n < N0; ++n) {
    const int tile = own.q;
    const int j = c_t2steps[own.r].job0;
    const T2Job r0 = c_t2jobs[j];
    const uint32_t fl = r0.gflags;
    const int glen = r0.glen;
    const uint32_t pos = (uint32_t)(own.q + oth.q) * (uint32_t)njobs + (uint32_t)j + (uint32_t)c_t2steps[oth.r].job0;
    const uint32_t t = SLOT == 0 ? 2u * (uint32_t)n : 2u * (uint32_t)(n + D) + 1u;
    const uint32_t next = SLOT == 0 ? (n >= D ? t + 1u : t + 2u) : (n + D + 1 < N0 ? t + 1u : t + 2u);
    if (fl & T2_WAIT_ENC) ts_wait(b_enc_full, (uint32_t)tile & 1u, abort_addr, err, 0x72000000 | (SLOT << 16) | j);
    if (fl & T2_WAIT_DIR) ts_wait(b_dir_full, (uint32_t)tile & 1u, abort_addr, err, 0x73000000 | (SLOT << 16) | j);
    if (fl & T2_WAIT_A) { ts_wait(b_a_ready, ph_a, abort_addr, err, 0x74000000 | (SLOT << 16) | j); ph_a ^= 1u; }
    if ((fl & T2_WAIT_EMPTY) && !((fl & T2_TILE_FIRST) && tile == 0)) {
      ts_wait(b_acc_empty, ph_e, abort_addr, err, 0x75000000 | (SLOT << 16) | j);
      ph_e ^= 1u;
    }
    if (kT2Exp && (abl & 32)) {   // timing only: weight stages probed before the turn (not safe against the ring alias)
      for (int k = 0; k < glen; ++k) {
        const uint32_t pk = pos + (uint32_t)k;
        ts_wait(base + L::w_full + 8 * (pk & (L::NST - 1)), (pk / L::NST) & 1u, abort_addr, err, 0x76000000 | (SLOT << 16) | (j + k));
      }
    }
    ts_wait_progress(turn_addr, t, abort_addr, err, 0x77000000 | (SLOT << 16) | j);
    if (!(kT2Exp && (abl & 32))) {
    for (int k = 0; k < glen; ++k) {
      const uint32_t pk = pos + (uint32_t)k;
      ts_wait(base + L::w_full + 8 * (pk & (L::NST - 1)), (pk / L::NST) & 1u, abort_addr, err, 0x76000000 | (SLOT << 16) | (j + k));
    }
    }
    tc_fence_after_sync();
    if (elect_one()) {
      for (int k = 0; k < glen; ++k) {
        const T2Job rc = c_t2jobs[j + k];
        const uint32_t f = rc.flags;
        const uint32_t ss = (pos + (uint32_t)k) & (L::NST - 1);
        if (k == glen - 1) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(turn_addr), "r"(next) : "memory");
        const uint64_t bd0 = ring_desc0 + (uint64_t)(ss * (kTsStageBytes >> 4));
        const uint64_t bd1 = bd0 + 1024u;
        const uint32_t first = (f & T2_FIRST) ? 0u : 1u;
        if (!(f & (T2_A_ENC | T2_A_DIR))) {
          const uint32_t a_t = a_base + rc.a_col;
          umma_bf16_ts(d, a_t, bd0, idesc, first);
          umma_bf16_ts(d, a_t + 8, bd0 + 2, idesc, 1u);
          umma_bf16_ts(d, a_t + 16, bd0 + 4, idesc, 1u);
          umma_bf16_ts(d, a_t + 24, bd0 + 6, idesc, 1u);
          umma_bf16_ts(d, a_t + 32, bd1, idesc, 1u);
          umma_bf16_ts(d, a_t + 40, bd1 + 2, idesc, 1u);
          umma_bf16_ts(d, a_t + 48, bd1 + 4, idesc, 1u);
          umma_bf16_ts(d, a_t + 56, bd1 + 6, idesc, 1u);
        } else {
          const uint64_t adesc = (f & T2_A_DIR) ? adesc_dir : adesc_enc;
          umma_bf16(d, adesc, bd0, idesc, first);
          umma_bf16(d, adesc + 2, bd0 + 2, idesc, 1u);
          if (rc.ksteps == 4) {
            umma_bf16(d, adesc + 4, bd0 + 4, idesc, 1u);
            umma_bf16(d, adesc + 6, bd0 + 6, idesc, 1u);
          }
        }
        umma_commit(base + L::w_empty + 8 * ss);
        if (f & T2_COMMIT_ACC) umma_commit(b_acc_full);
        if (f & T2_COMMIT_ENC) umma_commit(b_enc_empty);
        if (f & T2_COMMIT_DIR) umma_commit(b_dir_empty);
      }
    }
    __syncwarp();
    t2_cnt_inc(own, G);
    if (SLOT == 0) { if (n >= D) t2_cnt_inc(oth, G); }
    else if (m < N0) { ++m; t2_cnt_inc(oth, G); }
  }
}

// One epilogue step of one slot in MODE 2 (all eight epilogue warps, thread == (row, 64-channel slice)).  `held` is the
// slot's output half 0 between its HOLD and its FLUSH; the other slot's held half may be live at any step, so FLUSH keeps
// one 32-column group in flight.
struct T2Epi {
  uint32_t base, abort_addr, lane_addr;
  int wg, row;
  const float* sbias; const float* walpha; const float4* wrgb; float4* red;
};
template <int SL>
__device__ __forceinline__ void t2_epi_step(const T2Args& a, const T2Epi& e, int gi, int it, uint32_t (&held)[32], float& sig,
                                            uint32_t& ph) {
  using L = T2Smem;
  const T2Step st = c_t2steps[gi];
  const float* bias = e.sbias + st.bias_off + 64 * e.wg;
  const float* wal = e.walpha + 64 * e.wg;
  const uint32_t acc = e.lane_addr + 256u * SL + 128u + 64u * e.wg;
  const uint32_t abuf = e.lane_addr + 256u * SL + 32u * e.wg;
  ts_wait(e.base + L::acc_full + 8 * SL, ph, e.abort_addr, a.err, 0x7a000000 | (SL << 16) | gi);
  ph ^= 1u;
  tc_fence_after_sync();
  if (st.mode == T2_HOLD) {
    if (gi == 0) sig = 0.f;
    uint32_t v[32];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      tmem_ld32(acc + 32 * c, v);
      tmem_ld_wait();
      if (c == 1) {                    // ACC is in registers: the MMAs of half 1 may overwrite it
        tc_fence_before_sync();
        mbar_arrive(e.base + L::acc_empty + 8 * SL);
      }
      if (st.relu) t2_convert<true>(v, bias + 32 * c, &held[16 * c]);
      else t2_convert<false>(v, bias + 32 * c, &held[16 * c]);
    }
    if (st.dot) {
      sig = t2_dot32(&held[0], wal, sig);
      sig = t2_dot32(&held[16], wal + 32, sig);
    }
  } else if (st.mode == T2_FLUSH) {
    uint32_t v[32];
    tmem_ld32(acc, v);               // issued ahead of the stores of the held half (a tcgen05.st ahead of the ld delays it)
    {
      uint32_t w[16];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = held[16 * c + i];
        tmem_st16(abuf + 16 * c, w);
      }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t w[16];
      if (c > 0) tmem_ld32(acc + 32, v);
      tmem_ld_wait();
      if (st.relu) t2_convert<true>(v, bias + 128 + 32 * c, w); else t2_convert<false>(v, bias + 128 + 32 * c, w);
      tmem_st16(abuf + 64 + 16 * c, w);
      if (st.dot) sig = t2_dot32(w, wal + 128 + 32 * c, sig);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    mbar_arrive(e.base + L::a_ready + 8 * SL);   // next layer's input is in A (and ACC has been read out)
  } else {
    // OUT: views_linears.0 output (this thread's 64 of 128 channels) -> ReLU -> partial rgb_linear; the two slices of a
    // row meet in shared memory (double-buffered by tile parity: the next write of a buffer is two barriers away)
    uint32_t v[32], w[32];
    tmem_ld32(acc, v);
    tmem_ld_wait();
    t2_convert<true>(v, bias, &w[0]);
    tmem_ld32(acc + 32, v);
    tmem_ld_wait();
    tc_fence_before_sync();
    mbar_arrive(e.base + L::acc_empty + 8 * SL);   // the next tile's layer 0 may start on ACC
    t2_convert<true>(v, bias + 32, &w[16]);
    float cr = 0.f, cg = 0.f, cb = 0.f, cr2 = 0.f, cg2 = 0.f, cb2 = 0.f;
    t2_rgb32(&w[0], e.wrgb + 64 * e.wg, cr, cg, cb);
    t2_rgb32(&w[16], e.wrgb + 64 * e.wg + 32, cr2, cg2, cb2);
    float4* redp = e.red + ((it & 1) * 2 + SL) * 128;
    if (e.wg == 1) redp[e.row] = make_float4(cr + cr2, cg + cg2, cb + cb2, sig);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (e.wg == 0) {
      const int64_t p = (2 * ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) + SL) * kTileRows + e.row;
      if (p < a.P) {
        const float4 q = redp[e.row];
        float4 o;
        o.x = cr + cr2 + q.x + e.sbias[kBiasRgb + 0];
        o.y = cg + cg2 + q.y + e.sbias[kBiasRgb + 1];
        o.z = cb + cb2 + q.z + e.sbias[kBiasRgb + 2];
        o.w = sig + q.w + e.sbias[kBiasAlpha];
        st_stream4(reinterpret_cast<float4*>(a.raw) + p, o);
      }
    }
  }
}

// MODE 0: free-running issuers, one epilogue warpgroup per slot.  MODE 1: MMA groups issued in alternation (slot 0,
// slot 1, ...), all eight epilogue warps serve whichever slot's accumulator comes next in that fixed order.  MODE 2: as
// MODE 1 with slot 1 staggered by a.stagger groups (above).
template <int MODE, bool STASH = false>
__global__ void __launch_bounds__(kTsThreads, 1) nerf_mlp_t2_kernel(const T2Args a) {
  static_assert(!STASH || MODE == 1, "the stash-writing form exists for MODE 1 only");
  using L = T2Smem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (a.P + kTileRows - 1) / kTileRows;
  const int64_t npairs = (ntiles + 1) >> 1;
  const int my_pairs = (int)((npairs - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t abort_addr = base + L::abort_flag;

  if (threadIdx.x == 0) {
    for (int i = 0; i < L::NST; ++i) { mbar_init(base + L::w_full + 8 * i, 1); mbar_init(base + L::w_empty + 8 * i, MODE == 2 ? 1 : 2); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(base + L::acc_full + 8 * s, 1);
      mbar_init(base + L::acc_empty + 8 * s, MODE != 0 ? 256 : 128);
      mbar_init(base + L::a_ready + 8 * s, MODE != 0 ? 256 : 128);
      mbar_init(base + L::enc_full + 8 * s, 128);
      mbar_init(base + L::enc_empty + 8 * s, 1);
      mbar_init(base + L::dir_full + 8 * s, 128);
      mbar_init(base + L::dir_empty + 8 * s, 1);
    }
    *reinterpret_cast<volatile uint32_t*>(gen + L::abort_flag) = 0;
    *reinterpret_cast<volatile uint32_t*>(gen + L::lead) = 0;
    *reinterpret_cast<volatile uint32_t*>(gen + L::lead + 4) = 0;
    mbar_init_fence();
  }
  if (warp == 2) tmem_alloc(base + L::tmem_ptr, kTmemCols);
  {
    const float* gb = reinterpret_cast<const float*>(a.packed + reinterpret_cast<const uint32_t*>(a.packed)[2]);
    float* sb = reinterpret_cast<float*>(gen + L::bias);
    for (int i = threadIdx.x; i < kTsBiasFloats; i += blockDim.x) sb[i] = __ldg(gb + i);
    // alpha_linear.weight[0, :] and rgb_linear.weight[0:3, :] decoded from their bf16 slabs (row n of a 16-row
    // K-block image: sw128_offset(n, chunk))
    float* wa = reinterpret_cast<float*>(gen + L::walpha);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
      const int kp = i >> 7, kb = (i >> 6) & 1, kk = i & 63;
      const uint16_t bits = __ldg(reinterpret_cast<const uint16_t*>(a.packed + a.off_alpha[kp] + kb * 2048 + kk * 2));
      wa[i] = __uint_as_float((uint32_t)bits << 16);
    }
    float* wr = reinterpret_cast<float*>(gen + L::wrgb);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
      const int k = i >> 2, n = i & 3, kb = k >> 6, kk = k & 63;
      float v = 0.f;
      if (n < 3) {
        const uint16_t bits = __ldg(reinterpret_cast<const uint16_t*>(
            a.packed + a.off_rgb + kb * 2048 + sw128_offset((uint32_t)n, (uint32_t)(kk >> 3)) + (kk & 7) * 2));
        v = __uint_as_float((uint32_t)bits << 16);
      }
      wr[i] = v;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + L::tmem_ptr);
  long long exp_t0 = 0;
  if (kT2Exp) exp_t0 = clock64();

  if (warp == 0) {
    // =============================== weight producer: one fill per slab, consumed by both issuers ===================
    uint32_t s = 0, par = 0;
    if constexpr (MODE == 2) {
      // merged order: slot 0's group i, then slot 1's group i - D
      const int G = a.nsteps, D = a.stagger, N0 = my_pairs * G;
      int r0 = 0, r1 = 0;
      for (int i = 0; i < N0 + D; ++i) {
#pragma unroll 1
        for (int sl = 0; sl < 2; ++sl) {
          if (sl == 0 ? (i >= N0) : (i < D)) continue;
          const int j0 = c_t2steps[sl ? r1 : r0].job0;
          const int glen = c_t2jobs[j0].glen;
          for (int j = j0; j < j0 + glen; ++j) {
            ts_wait(base + L::w_empty + 8 * s, par ^ 1u, abort_addr, a.err, 0x71000000 | (sl << 16) | j);
            const uint32_t bytes = (uint32_t)c_t2jobs[j].bytes16 * 16;
            const uint8_t* src = a.packed + c_t2jobs[j].w_off;
            if (elect_one()) {
              mbar_expect_tx(base + L::w_full + 8 * s, bytes);
              tma_bulk_g2s(base + L::ring + s * kTsStageBytes, src, bytes, base + L::w_full + 8 * s);
            }
            __syncwarp();
            if (++s == (uint32_t)L::NST) { s = 0; par ^= 1u; }
          }
        }
        if (i < N0 && ++r0 == G) r0 = 0;
        if (i >= D && ++r1 == G) r1 = 0;
      }
    } else
    for (int it = 0; it < my_pairs; ++it)
      for (int j = 0; j < a.njobs; ++j) {
        ts_wait(base + L::w_empty + 8 * s, par ^ 1u, abort_addr, a.err, 0x71000000 | j);
        const uint32_t bytes = (uint32_t)c_t2jobs[j].bytes16 * 16;
        const uint8_t* src = a.packed + c_t2jobs[j].w_off;
        if (elect_one()) {
          mbar_expect_tx(base + L::w_full + 8 * s, bytes);
          tma_bulk_g2s(base + L::ring + s * kTsStageBytes, src, bytes, base + L::w_full + 8 * s);
        }
        __syncwarp();
        if (++s == (uint32_t)L::NST) { s = 0; par ^= 1u; }
      }
  } else if (warp == 1) {
    if constexpr (MODE == 2) t2_issue_stag<0>(a.err, a.njobs, a.nsteps, a.stagger, base, abort_addr, tmem, my_pairs, a.abl);
    else if constexpr (MODE == 1) t2_issue_groups<0>(a.err, a.njobs, base, abort_addr, tmem, my_pairs, a.trace, a.trace_it, a.turn_back);
    else t2_issue_loop<0>(a.err, a.njobs, base, abort_addr, tmem, my_pairs, a.trace, a.trace_it);
  } else if (warp == 3) {
    if constexpr (MODE == 2) t2_issue_stag<1>(a.err, a.njobs, a.nsteps, a.stagger, base, abort_addr, tmem, my_pairs, a.abl);
    else if constexpr (MODE == 1) t2_issue_groups<1>(a.err, a.njobs, base, abort_addr, tmem, my_pairs, a.trace, a.trace_it, a.turn_back);
    else t2_issue_loop<1>(a.err, a.njobs, base, abort_addr, tmem, my_pairs, a.trace, a.trace_it);
  } else if (warp >= 4 && warp < 8) {
    // =============================== per-tile input blocks of both slots: thread == row =============================
    const int row = threadIdx.x - 128;
    const uint32_t row_off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    for (int it = 0; it < my_pairs; ++it) {
      const int64_t tile0 = 2 * ((int64_t)blockIdx.x + (int64_t)it * gridDim.x);
#pragma unroll 1
      for (int sl = 0; sl < 2; ++sl) {
        const int64_t p = (tile0 + sl) * kTileRows + row;
        uint32_t w[32];
        {
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
                if (kT2Exp && (a.abl & 2)) {
#pragma unroll
                  for (int k = 0; k < 20; ++k) sc[k] = x[i] * (float)(k + 1);
                } else
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
        if (it > 0) ts_wait(base + L::enc_empty + 8 * sl, (uint32_t)(it - 1) & 1u, abort_addr, a.err, 0x78000000 | (sl << 16) | it);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t off = ((uint32_t)(c ^ (row & 7)) << 4);
          st_smem16(base + L::enc + sl * kBlkBytes + row_off + off, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        }
        if constexpr (STASH) {
          if (tile0 + sl < a.ntiles)
            t2_stash_chunks<8>(a.stash + (size_t)(tile0 + sl) * kStashTileBytes + (size_t)kHEnc * kBlkBytes +
                                   stash_chunk_off((uint32_t)row, 0u), 0, w);
        }
        fence_proxy_async_smem();
        mbar_arrive(base + L::enc_full + 8 * sl);
      }
#pragma unroll 1
      for (int sl = 0; sl < 2; ++sl) {
        const int64_t p = (tile0 + sl) * kTileRows + row;
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
        if (it > 0) ts_wait(base + L::dir_empty + 8 * sl, (uint32_t)(it - 1) & 1u, abort_addr, a.err, 0x79000000 | (sl << 16) | it);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t q[4] = {0u, 0u, 0u, 0u};
          if (c < 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = pack_bf16(e[8 * c + 2 * i], e[8 * c + 2 * i + 1]);
          }
          const uint32_t off = ((uint32_t)(c ^ (row & 7)) << 4);
          st_smem16(base + L::dir + sl * kBlkBytes + row_off + off, q[0], q[1], q[2], q[3]);
          if constexpr (STASH) {
            if (tile0 + sl < a.ntiles)
              t2_stash_chunks<1>(a.stash + (size_t)(tile0 + sl) * kStashTileBytes + (size_t)kHDir * kBlkBytes +
                                     stash_chunk_off((uint32_t)row, 0u), c, q);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(base + L::dir_full + 8 * sl);
      }
    }
  } else if (warp >= 8) {
   if constexpr (MODE == 0) {
    // =============================== epilogue of one slot: thread == row, 128 channels per step ======================
    const int slot = (warp - 8) >> 2;
    const int row = ((warp & 3) << 5) | lane;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
    const uint32_t acc = lane_addr + 256u * slot + 128u, abuf = lane_addr + 256u * slot;
    const uint32_t b_acc_full = base + L::acc_full + 8 * slot, b_acc_empty = base + L::acc_empty + 8 * slot;
    const uint32_t b_a_ready = base + L::a_ready + 8 * slot;
    const float* sbias = reinterpret_cast<const float*>(gen + L::bias);
    const float* walpha = reinterpret_cast<const float*>(gen + L::walpha);
    const float4* wrgb = reinterpret_cast<const float4*>(gen + L::wrgb);
    uint32_t ph = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const int64_t tile = 2 * ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) + slot;
      const int64_t p = tile * kTileRows + row;
      float sigma = 0.f;
      const bool trace_on = kDiag && a.trace && blockIdx.x == 0 && it == a.trace_it && (threadIdx.x & 127) == 0;
      // Steps come in pairs (one layer): HOLD converts output half 0 and keeps it in registers, FLUSH converts half 1 and
      // stores both over A_s.  `held` lives inside the pair so that its 64 registers are free again for FLUSH's loads.
      const int nlayers = (a.nsteps - 1) >> 1;
      for (int li = 0; li < nlayers; ++li) {
        const T2Step st = c_t2steps[2 * li];
        const float* bias = sbias + st.bias_off;
        unsigned long long* tr = trace_on ? a.trace + 512 + (slot * 24 + 2 * li) * 4 : nullptr;
        uint32_t held[64];   // output half 0 of this layer (bf16x2), waiting for half 1
        uint32_t v[32];
        // ---- HOLD: one 32-column group in flight at a time (64 held + 32 in flight + conversion temporaries is what
        // 128 registers per thread allow)
        if (tr) tr[0] = clock64();
        ts_wait(b_acc_full, ph, abort_addr, a.err, 0x7a000000 | (slot << 16) | (2 * li));
        ph ^= 1u;
        if (tr) tr[1] = clock64();
        tc_fence_after_sync();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld32(acc + 32 * c, v);
          tmem_ld_wait();
          if (c == 3) {                      // ACC_s is in registers: the MMAs of half 1 may overwrite it
            tc_fence_before_sync();
            mbar_arrive(b_acc_empty);
            if (tr) tr[2] = clock64();
          }
          if (st.relu) t2_convert<true>(v, bias + 32 * c, &held[16 * c]);
          else t2_convert<false>(v, bias + 32 * c, &held[16 * c]);
        }
        if (st.dot) {
#pragma unroll
          for (int c = 0; c < 4; ++c) sigma = t2_dot32(&held[16 * c], walpha + 32 * c, sigma);
        }
        if (tr) tr[3] = clock64();
        // ---- FLUSH: every MMA of this layer has completed (acc_full follows the last one): A_s is overwritten in place
        if (tr) tr[4] = clock64();
        ts_wait(b_acc_full, ph, abort_addr, a.err, 0x7a000000 | (slot << 16) | (2 * li + 1));
        ph ^= 1u;
        if (tr) tr[5] = clock64();
        tc_fence_after_sync();
        {
          uint32_t w[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = held[16 * c + i];
            tmem_st16(abuf + 16 * c, w);
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w[16];
          tmem_ld32(acc + 32 * c, v);
          tmem_ld_wait();
          if (st.relu) t2_convert<true>(v, bias + 128 + 32 * c, w); else t2_convert<false>(v, bias + 128 + 32 * c, w);
          tmem_st16(abuf + 64 + 16 * c, w);
          if (st.dot) sigma = t2_dot32(w, walpha + 128 + 32 * c, sigma);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive(b_a_ready);              // next layer's input is in A_s (and ACC_s has been read out)
        if (tr) { tr[6] = clock64(); tr[7] = tr[6]; }
      }
      {
        // ---- OUT: views_linears.0 output (128 channels) -> ReLU -> rgb_linear on the CUDA cores -> (r, g, b, sigma)
        const T2Step st = c_t2steps[a.nsteps - 1];
        const float* bias = sbias + st.bias_off;
        unsigned long long* tr = trace_on ? a.trace + 512 + (slot * 24 + a.nsteps - 1) * 4 : nullptr;
        if (tr) tr[0] = clock64();
        ts_wait(b_acc_full, ph, abort_addr, a.err, 0x7a000000 | (slot << 16) | (a.nsteps - 1));
        ph ^= 1u;
        if (tr) tr[1] = clock64();
        tc_fence_after_sync();
        float cr = 0.f, cg = 0.f, cb = 0.f;
        uint32_t v[32], v2[32];
        tmem_ld32(acc, v);
        tmem_ld32(acc + 32, v2);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w[16];
          tmem_ld_wait();
          if ((c & 1) == 0) {
            t2_convert<true>(v, bias + 32 * c, w);
            if (c == 0) tmem_ld32(acc + 64, v);
          } else {
            t2_convert<true>(v2, bias + 32 * c, w);
            if (c == 1) tmem_ld32(acc + 96, v2);
          }
          if (c == 3) {                      // the next tile's layer 0 may start on ACC_s
            tc_fence_before_sync();
            mbar_arrive(b_acc_empty);
            if (tr) tr[2] = clock64();
          }
          t2_rgb32(w, wrgb + 32 * c, cr, cg, cb);
        }
        if (p < a.P) {
          float4 o;
          o.x = cr + sbias[kBiasRgb + 0];
          o.y = cg + sbias[kBiasRgb + 1];
          o.z = cb + sbias[kBiasRgb + 2];
          o.w = sigma + sbias[kBiasAlpha];
          st_stream4(reinterpret_cast<float4*>(a.raw) + p, o);
        }
        if (tr) tr[3] = clock64();
      }
    }
   } else if constexpr (MODE == 2) {
    // =============================== epilogue, both slots in the merged (staggered) order ============================
    T2Epi e;
    e.base = base; e.abort_addr = abort_addr;
    e.lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
    e.wg = (warp - 8) >> 2; e.row = ((warp & 3) << 5) | lane;
    e.sbias = reinterpret_cast<const float*>(gen + L::bias);
    e.walpha = reinterpret_cast<const float*>(gen + L::walpha);
    e.wrgb = reinterpret_cast<const float4*>(gen + L::wrgb);
    e.red = reinterpret_cast<float4*>(gen + L::red);
    const int G = a.nsteps, D = a.stagger, N0 = my_pairs * G;
    uint32_t heldA[32], heldB[32];
    float sigA = 0.f, sigB = 0.f;
    uint32_t phA = 0, phB = 0;
    T2Cnt cA{0, 0}, cB{0, 0};
    for (int i = 0; i < N0 + D; ++i) {
      if (i < N0) { t2_epi_step<0>(a, e, cA.r, cA.q, heldA, sigA, phA); t2_cnt_inc(cA, G); }
      if (i >= D) { t2_epi_step<1>(a, e, cB.r, cB.q, heldB, sigB, phB); t2_cnt_inc(cB, G); }
    }
   } else {
    // =============================== epilogue, both slots: thread == (row, 64-channel slice) =========================
    // acc_full events arrive in the fixed order of the alternating issue: per layer  slot 0 half 0, slot 1 half 0,
    // slot 0 half 1, slot 1 half 1.  All eight warps take each of them in turn: a step is two 32-column groups per
    // thread instead of four, so a slot's hand-over comes back in about the time the other slot's MMA group runs.
    const int wg = (warp - 8) >> 2;
    const int row = ((warp & 3) << 5) | lane;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
    const float* sbias = reinterpret_cast<const float*>(gen + L::bias);
    const float* walpha = reinterpret_cast<const float*>(gen + L::walpha);
    const float4* wrgb = reinterpret_cast<const float4*>(gen + L::wrgb);
    float4* red = reinterpret_cast<float4*>(gen + L::red);
    const uint32_t srow = stash_chunk_off((uint32_t)row, 0u);   // this point's place inside a stash block
    (void)srow;
    uint32_t ph0 = 0, ph1 = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const int64_t tile0 = 2 * ((int64_t)blockIdx.x + (int64_t)it * gridDim.x);
      float sig0 = 0.f, sig1 = 0.f;
      const bool trace_on = kDiag && a.trace && blockIdx.x == 0 && it == a.trace_it && (threadIdx.x & 255) == 0;
      const int nlayers = (a.nsteps - 1) >> 1;
      for (int li = 0; li < nlayers; ++li) {
        const T2Step st = c_t2steps[2 * li];
        const float* bias = sbias + st.bias_off + 64 * wg;
        const float* wal = walpha + 64 * wg;
        uint32_t heldA[32], heldB[32];   // this thread's 64 channels of output half 0, slot 0 / slot 1
        unsigned long long* tr = trace_on ? a.trace + 512 + (2 * li) * 4 : nullptr;
        // ---- half 0 of both slots: convert and keep
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          uint32_t* held = sl ? heldB : heldA;
          const uint32_t acc = lane_addr + 256u * sl + 128u + 64u * wg;
          unsigned long long* t4 = tr ? tr + sl * 24 * 4 : nullptr;
          if (t4) t4[0] = clock64();
          ts_wait(base + L::acc_full + 8 * sl, sl ? ph1 : ph0, abort_addr, a.err, 0x7a000000 | (sl << 16) | (2 * li));
          if (sl) ph1 ^= 1u; else ph0 ^= 1u;
          if (t4) t4[1] = clock64();
          tc_fence_after_sync();
          uint32_t v[32];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            tmem_ld32(acc + 32 * c, v);
            tmem_ld_wait();
            if (c == 1) {                    // ACC is in registers: the MMAs of half 1 may overwrite it
              tc_fence_before_sync();
              mbar_arrive(base + L::acc_empty + 8 * sl);
              if (t4) t4[2] = clock64();
            }
            if (st.relu) t2_convert<true>(v, bias + 32 * c, &held[16 * c]);
            else t2_convert<false>(v, bias + 32 * c, &held[16 * c]);
            if constexpr (STASH) {
              if (tile0 + sl < a.ntiles)
                t2_stash_chunks<4>(a.stash + (size_t)(tile0 + sl) * kStashTileBytes + (size_t)(st.out_blk + wg) * kBlkBytes + srow,
                                   4 * c, &held[16 * c]);
            }
          }
          if (st.dot) {
            float sg = sl ? sig1 : sig0;
            sg = t2_dot32(&held[0], wal, sg);
            sg = t2_dot32(&held[16], wal + 32, sg);
            if (sl) sig1 = sg; else sig0 = sg;
          }
          if (t4) t4[3] = clock64();
        }
        // ---- half 1 of both slots: every MMA of the layer has completed, A is overwritten in place
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const uint32_t* held = sl ? heldB : heldA;
          const uint32_t acc = lane_addr + 256u * sl + 128u + 64u * wg;
          const uint32_t abuf = lane_addr + 256u * sl + 32u * wg;
          unsigned long long* t4 = tr ? tr + sl * 24 * 4 + 4 : nullptr;
          if (t4) t4[0] = clock64();
          ts_wait(base + L::acc_full + 8 * sl, sl ? ph1 : ph0, abort_addr, a.err, 0x7a000000 | (sl << 16) | (2 * li + 1));
          if (sl) ph1 ^= 1u; else ph0 ^= 1u;
          if (t4) t4[1] = clock64();
          tc_fence_after_sync();
          // Loads are issued BEFORE the stores of the held half (measured: a tcgen05.st ahead of the tcgen05.ld delays it
          // by more than the store takes).  Slot 1 is the last user of a held half in the layer, so both of its column
          // groups fit in flight; slot 0 shares the registers with slot 1's held half and takes one group at a time.
          uint32_t v[32], v2[32];
          tmem_ld32(acc, v);
          if (sl == 1) tmem_ld32(acc + 32, v2);
          {
            uint32_t w[16];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
              for (int i = 0; i < 16; ++i) w[i] = held[16 * c + i];
              tmem_st16(abuf + 16 * c, w);
            }
          }
          if (kT2Exp && (a.abl & 16)) {   // hand over as soon as the held half is stored and the first loads are back
            tmem_ld_wait();
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(base + L::a_ready + 8 * sl);
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t w[16];
            if (c > 0 && sl == 0) tmem_ld32(acc + 32, v2);
            if (c == 0 || sl == 0) tmem_ld_wait();
            if (c == 0) { if (st.relu) t2_convert<true>(v, bias + 128, w); else t2_convert<false>(v, bias + 128, w); }
            else { if (st.relu) t2_convert<true>(v2, bias + 160, w); else t2_convert<false>(v2, bias + 160, w); }
            tmem_st16(abuf + 64 + 16 * c, w);
            if constexpr (STASH) {
              if (tile0 + sl < a.ntiles)
                t2_stash_chunks<4>(a.stash + (size_t)(tile0 + sl) * kStashTileBytes + (size_t)(st.out_blk + 2 + wg) * kBlkBytes + srow,
                                   4 * c, w);
            }
            if (st.dot) {
              if (sl) sig1 = t2_dot32(w, wal + 128 + 32 * c, sig1); else sig0 = t2_dot32(w, wal + 128 + 32 * c, sig0);
            }
          }
          tmem_st_wait();
          tc_fence_before_sync();
          if (!(kT2Exp && (a.abl & 16)))
            mbar_arrive(base + L::a_ready + 8 * sl);   // next layer's input is in A (and ACC has been read out)
          if (t4) { t4[2] = clock64(); t4[3] = t4[2]; }
        }
      }
      // ---- views_linears.0 output (this thread's 64 of 128 channels) -> ReLU -> partial rgb_linear; the two slices of
      // a row meet in shared memory
      {
        const T2Step st = c_t2steps[a.nsteps - 1];
        const float* bias = sbias + st.bias_off + 64 * wg;
        float4 part[2];
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const uint32_t acc = lane_addr + 256u * sl + 128u + 64u * wg;
          unsigned long long* t4 = trace_on ? a.trace + 512 + (sl * 24 + a.nsteps - 1) * 4 : nullptr;
          if (t4) t4[0] = clock64();
          ts_wait(base + L::acc_full + 8 * sl, sl ? ph1 : ph0, abort_addr, a.err, 0x7a000000 | (sl << 16) | (a.nsteps - 1));
          if (sl) ph1 ^= 1u; else ph0 ^= 1u;
          if (t4) t4[1] = clock64();
          tc_fence_after_sync();
          uint32_t v[32], w[32];
          tmem_ld32(acc, v);
          tmem_ld_wait();
          t2_convert<true>(v, bias, &w[0]);
          tmem_ld32(acc + 32, v);
          tmem_ld_wait();
          tc_fence_before_sync();
          mbar_arrive(base + L::acc_empty + 8 * sl);   // the next tile's layer 0 may start on ACC
          if (t4) t4[2] = clock64();
          if (!(kT2Exp && (a.abl & 8))) t2_convert<true>(v, bias + 32, &w[16]);
          if constexpr (STASH) {
            if (tile0 + sl < a.ntiles)
              t2_stash_chunks<8>(a.stash + (size_t)(tile0 + sl) * kStashTileBytes + (size_t)(st.out_blk + wg) * kBlkBytes + srow, 0, w);
          }
          float cr = 0.f, cg = 0.f, cb = 0.f, cr2 = 0.f, cg2 = 0.f, cb2 = 0.f;
          if (kT2Exp && (a.abl & 9)) {
            cr = __uint_as_float(w[0]); cr2 = __uint_as_float(w[31]);
          } else {
            t2_rgb32(&w[0], wrgb + 64 * wg, cr, cg, cb);
            t2_rgb32(&w[16], wrgb + 64 * wg + 32, cr2, cg2, cb2);
          }
          part[sl] = make_float4(cr + cr2, cg + cg2, cb + cb2, sl ? sig1 : sig0);
          if (t4) t4[3] = clock64();
        }
        if (wg == 1) { red[row] = part[0]; red[128 + row] = part[1]; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (wg == 0) {
#pragma unroll
          for (int sl = 0; sl < 2; ++sl) {
            const int64_t p = (tile0 + sl) * kTileRows + row;
            if (p < a.P) {
              const float4 q = red[sl * 128 + row];
              float4 o;
              o.x = part[sl].x + q.x + sbias[kBiasRgb + 0];
              o.y = part[sl].y + q.y + sbias[kBiasRgb + 1];
              o.z = part[sl].z + q.z + sbias[kBiasRgb + 2];
              o.w = part[sl].w + q.w + sbias[kBiasAlpha];
              st_stream4(reinterpret_cast<float4*>(a.raw) + p, o);
            }
          }
        }
      }
    }
   }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, kTmemCols);
  if (kT2Exp && blockIdx.x == 0 && threadIdx.x == 0) {   // SM cycles of CTA 0 and its pair count: byte 64 of the workspace
    *reinterpret_cast<long long*>(a.err + 16) = clock64() - exp_t0;
    a.err[18] = my_pairs;
  }
}

// ---- host: job / step tables from the forward plan's slab list -------------------------------------------------------
struct T2Tables {
  std::vector<T2Job> jobs;
  std::vector<T2Step> steps;
  uint32_t off_alpha[2], off_rgb;
  bool ok;
};

static T2Tables t2_build(const TsPlan& p) {
  T2Tables t{};
  t.ok = true;
  auto slab = [&](int layer, int row0, int col0, int nkb) -> const TsPackJob* {
    for (const TsPackJob& q : p.pack)
      if (q.layer == layer && q.row0 == row0 && q.col0 == col0 && q.nkb == nkb) return &q;
    t.ok = false;
    return &p.pack[0];
  };
  auto job = [&](const TsPackJob* q, int flags, int a_col, int ksteps) {
    T2Job j{};
    j.w_off = q->w_off; j.bytes16 = (uint16_t)(q->nkb * q->rows * 8); j.flags = (uint16_t)flags; j.a_col = (uint16_t)a_col;
    j.ksteps = (uint8_t)ksteps; j.nkb = q->nkb;
    if (q->rows != 128) t.ok = false;
    t.jobs.push_back(j);
  };
  auto step = [&](int mode, int relu, int dot, int bias_off, int out_blk) {
    T2Step s{};
    s.mode = (uint8_t)mode; s.relu = (uint8_t)relu; s.dot = (uint8_t)dot; s.bias_off = (uint16_t)bias_off;
    s.out_blk = (uint16_t)out_blk;
    t.steps.push_back(s);
  };
  // layer 0: A = the slot's encoding block
  job(slab(0, 0, 0, 1), T2_WAIT_ENC | T2_TILE_FIRST | T2_WAIT_EMPTY | T2_FIRST | T2_A_ENC | T2_COMMIT_ACC, 0, 4);
  job(slab(0, 128, 0, 1), T2_WAIT_EMPTY | T2_FIRST | T2_A_ENC | T2_COMMIT_ACC, 0, 4);
  step(T2_HOLD, 1, 0, 0, 0); step(T2_FLUSH, 1, 0, 0, 2);
  auto wide = [&](int layer, int c0, bool skip) {
    for (int h = 0; h < 2; ++h) {
      job(slab(layer, 128 * h, c0, 2), (h == 0 ? T2_WAIT_A : T2_WAIT_EMPTY) | T2_FIRST, 0, 4);
      job(slab(layer, 128 * h, c0 + 128, 2), skip ? 0 : T2_COMMIT_ACC, 64, 4);
      if (skip) job(slab(layer, 128 * h, 0, 1), T2_A_ENC | T2_COMMIT_ACC | (h == 1 ? T2_COMMIT_ENC : 0), 0, 4);
    }
  };
  for (int l = 1; l <= 7; ++l) {
    wide(l, l == 5 ? 63 : 0, l == 5);
    step(T2_HOLD, 1, l == 7, 256 * l, 4 * l); step(T2_FLUSH, 1, l == 7, 256 * l, 4 * l + 2);
    if (kT2Exp && l == 2) {   // timing only: layer 2 again, n times (steady-state period of a wide layer = d time / d n)
      const char* e = getenv("GBNERF_T2_DBG_EXTRA");
      for (int n = e ? atoi(e) : 0; n > 0; --n) { wide(2, 0, false); step(T2_HOLD, 1, 0, 512, 8); step(T2_FLUSH, 1, 0, 512, 10); }
    }
  }
  wide(LIN_FEATURE, 0, false);
  step(T2_HOLD, 0, 0, kBiasFeat, kHFeat); step(T2_FLUSH, 0, 0, kBiasFeat, kHFeat + 2);
  job(slab(LIN_VIEWS, 0, 0, 2), T2_WAIT_A | T2_FIRST, 0, 4);
  job(slab(LIN_VIEWS, 0, 128, 2), 0, 64, 4);
  job(slab(LIN_VIEWS, 0, 256, 1), T2_WAIT_DIR | T2_A_DIR | T2_COMMIT_ACC | T2_COMMIT_DIR, 0, 2);
  step(T2_OUT, 1, 0, kTsBiasViews, kHHv);
  // alpha / rgb slabs (16 rows): decoded to fp32 in the kernel prologue
  t.off_alpha[0] = t.off_alpha[1] = t.off_rgb = 0;
  for (const TsPackJob& q : p.pack) {
    if (q.layer == LIN_ALPHA && q.rows == 16 && q.nkb == 2) t.off_alpha[q.col0 >= 128 ? 1 : 0] = q.w_off;
    if (q.layer == LIN_RGB && q.rows == 16 && q.nkb == 2) t.off_rgb = q.w_off;
  }
  if (!t.off_alpha[0] || !t.off_alpha[1] || !t.off_rgb) t.ok = false;
  for (size_t i = 0; i < t.jobs.size();) {   // groups: from a T2_FIRST job up to and including the job that commits the accumulator
    size_t e = i;
    uint16_t gf = 0;
    for (;; ++e) {
      if (e >= t.jobs.size()) { t.ok = false; break; }
      gf |= t.jobs[e].flags;
      if (t.jobs[e].flags & T2_COMMIT_ACC) break;
    }
    if (!t.ok || !(t.jobs[i].flags & T2_FIRST)) { t.ok = false; break; }
    t.jobs[i].glen = (uint8_t)(e - i + 1);
    t.jobs[i].gflags = gf;
    i = e + 1;
  }
  {   // job0 of every step's group (steps and groups are one to one, in order)
    size_t g = 0;
    for (size_t i = 0; i < t.jobs.size(); ++i)
      if (t.jobs[i].glen) {
        if (g < t.steps.size()) t.steps[g].job0 = (uint8_t)i;
        ++g;
      }
    if (g != t.steps.size()) t.ok = false;
  }
  if ((int)t.jobs.size() > kT2MaxJobs || (int)t.steps.size() > kT2MaxSteps) t.ok = false;
  return t;
}

static T2Tables g_t2;

// GBNERF_MLP_T2=0 keeps inference on the one-tile kernel
static bool t2_enabled() {
  static const bool on = [] { const char* e = getenv("GBNERF_MLP_T2"); return !(e && e[0] == '0'); }();
  return on;
}

// GBNERF_T2_STASH=0 keeps the stash-writing (training) forward on the one-tile kernel
static bool t2_stash_enabled() {
  static const bool on = [] { const char* e = getenv("GBNERF_T2_STASH"); return !(e && e[0] == '0'); }();
  return on && t2_enabled();
}

// GBNERF_T2_MODE: 1 (default) = alternating MMA groups + shared epilogue warps, 0 = free-running issuers; in the exp
// build (-DGBN_T2_EXP) also 2 = as 1 with slot 1 staggered by GBNERF_T2_STAGGER groups (default 10)
static int t2_mode() {
  static const int m = [] { const char* e = getenv("GBNERF_T2_MODE"); return (e && e[0] == '0') ? 0 : (kT2Exp && e && e[0] == '2') ? 2 : 1; }();
  return m;
}
static int t2_stagger(int nsteps) {
  static const int d0 = [] { const char* e = getenv("GBNERF_T2_STAGGER"); return e ? atoi(e) : 10; }();
  int d = d0;
  if (kT2Exp) { const char* e = getenv("GBNERF_T2_STAGGER"); if (e) d = atoi(e); }   // per launch: tools/t2_exp.py sweeps it
  return d < 1 ? 1 : d >= nsteps ? nsteps - 1 : d;
}

static int t2_upload() {   // called from ts_ensure_device (per device)
  if (g_t2.jobs.empty()) g_t2 = t2_build(ts_plan(0));
  GBN_REQUIRE(g_t2.ok, "T2 job table does not match the forward plan");
  GBN_CUDA(cudaMemcpyToSymbol(c_t2jobs, g_t2.jobs.data(), g_t2.jobs.size() * sizeof(T2Job), 0, cudaMemcpyHostToDevice));
  GBN_CUDA(cudaMemcpyToSymbol(c_t2steps, g_t2.steps.data(), g_t2.steps.size() * sizeof(T2Step), 0, cudaMemcpyHostToDevice));
  GBN_CUDA(cudaFuncSetAttribute(nerf_mlp_t2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2Smem::alloc));
  GBN_CUDA(cudaFuncSetAttribute(nerf_mlp_t2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2Smem::alloc));
  GBN_CUDA(cudaFuncSetAttribute(nerf_mlp_t2_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2Smem::alloc));
  if constexpr (kT2Exp)
    GBN_CUDA(cudaFuncSetAttribute(nerf_mlp_t2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2Smem::alloc));
  return GBN_OK;
}

static int t2_forward(const void* packed, const float* ro, const float* rd, const float* vd, int64_t stride, const float* z,
                      const float* pts, const float* emb, int64_t R, int S, float* raw, int* err, cudaStream_t stream,
                      unsigned long long* trace, int trace_it, void* stash = nullptr) {
  T2Args a{};
  a.packed = reinterpret_cast<const uint8_t*>(packed);
  a.ro = ro; a.rd = rd; a.z = z; a.pts = pts; a.emb = emb; a.vd = vd; a.raw = raw; a.err = err;
  a.stride = stride; a.P = R * S; a.S = S;
  a.njobs = (int)g_t2.jobs.size(); a.nsteps = (int)g_t2.steps.size();
  a.trace = trace; a.trace_it = trace_it;
  a.mode = t2_mode();
  if (kT2Exp) {
    const char* e = getenv("GBNERF_T2_TURN_BACK");   // read per launch: tools/t2_exp.py sweeps them in one process
    a.turn_back = e ? atoi(e) : 2;
    e = getenv("GBNERF_T2_ABL");
    a.abl = e ? atoi(e) : 0;
  }
  a.off_alpha[0] = g_t2.off_alpha[0]; a.off_alpha[1] = g_t2.off_alpha[1]; a.off_rgb = g_t2.off_rgb;
  const int64_t ntiles = (a.P + kTileRows - 1) / kTileRows;
  const int64_t npairs = (ntiles + 1) / 2;
  const int grid = (int)(npairs < kNumSMs ? npairs : kNumSMs);
  a.stagger = t2_stagger(a.nsteps);
  a.stash = static_cast<uint8_t*>(stash); a.ntiles = ntiles;
  if (stash != nullptr) {   // training forward: MODE 1 with the H stash written from the epilogue's registers
    nerf_mlp_t2_kernel<1, true><<<grid, kTsThreads, T2Smem::alloc, stream>>>(a);
    return check_launch("nerf_mlp_t2_kernel<stash>");
  }
  if constexpr (kT2Exp) {   // the staggered mode is an experiment (measured slower): exp build only
    if (a.mode == 2) {
      nerf_mlp_t2_kernel<2><<<grid, kTsThreads, T2Smem::alloc, stream>>>(a);
      return check_launch("nerf_mlp_t2_kernel");
    }
  }
  if (a.mode != 0) nerf_mlp_t2_kernel<1><<<grid, kTsThreads, T2Smem::alloc, stream>>>(a);
  else nerf_mlp_t2_kernel<0><<<grid, kTsThreads, T2Smem::alloc, stream>>>(a);
  return check_launch("nerf_mlp_t2_kernel");
}
