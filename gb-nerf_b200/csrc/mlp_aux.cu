// Companions of the tcgen05 MLP kernel: the weight pre-packer (nn.Linear fp32 -> UMMA shared-memory images)
// and the per-ray view-direction bias of views_linears.0 (run_nerf_helpers.py:117-121).
#include <stdlib.h>

#ifndef GBN_DEFAULT_MLP_VARIANT
#define GBN_DEFAULT_MLP_VARIANT 1
#endif

#include <mutex>

#include "common.cuh"
#include "mlp_layout.h"
#include "tc_ptx.cuh"

namespace gbn {

const MlpPlan& mlp_plan(int precision);  // mlp_tc.cu

__constant__ PackJob c_pack[kNumPlans][kMaxJobs];

struct ParamPtrs {
  const float* w[GBN_NUM_LINEAR];
  const float* b[GBN_NUM_LINEAR];
};

struct PackHeader {
  uint32_t magic, precision, off_bias, off_wdir, off_bdir, total_bytes, njobs, pad;
};

// one block per weight chunk: [rows x KB] slice of a linear's weight -> K-major, 128-byte-swizzled image
template <int PREC>
__global__ void __launch_bounds__(256) prepack_kernel(ParamPtrs pp, uint8_t* __restrict__ out, int njobs,
                                                      PackHeader hdr, int plan) {
  constexpr int ESZ = PREC == GBN_PRECISION_BF16 ? 2 : 4;
  constexpr int KB = 128 / ESZ;
  if ((int)blockIdx.x < njobs) {
    const PackJob q = c_pack[plan][blockIdx.x];
    const float* W = pp.w[q.layer];
    for (int i = threadIdx.x; i < q.rows * KB; i += blockDim.x) {
      // transposed slabs read W with n as the fast index: let consecutive threads walk n there
      const int n = q.transpose ? i % q.rows : i / KB;
      const int k = q.transpose ? i / q.rows : i - n * KB;
      float v = 0.f;
      if (!q.transpose) {
        if (n < q.rows_valid && k < q.cols_valid) v = __ldg(W + (size_t)(q.row0 + n) * q.ld + q.col0 + k);
      } else {
        const int ks = k - (int)q.koff;
        if (n < q.rows_valid && ks >= 0 && ks < q.cols_valid) v = __ldg(W + (size_t)(q.row0 + ks) * q.ld + q.col0 + n);
      }
      const uint32_t byte = (uint32_t)k * ESZ;
      uint8_t* dst = out + q.w_off + tc::sw128_offset((uint32_t)n, byte >> 4) + (byte & 15);
      if constexpr (PREC == GBN_PRECISION_BF16) {
        *reinterpret_cast<uint16_t*>(dst) = (uint16_t)(tc::pack_bf16(v, 0.f) & 0xffff);
      } else {
        *reinterpret_cast<uint32_t*>(dst) = tc::to_tf32(v);
      }
    }
    return;
  }
  // trailing block: header, biases, direction part of views_linears.0
  if (threadIdx.x == 0) *reinterpret_cast<PackHeader*>(out) = hdr;
  float* bias = reinterpret_cast<float*>(out + hdr.off_bias);
  for (int i = threadIdx.x; i < kBiasFloats; i += blockDim.x) {
    float v = 0.f;
    if (i < kBiasFeat) v = pp.b[i >> 8][i & 255];
    else if (i < kBiasAlpha) v = pp.b[LIN_FEATURE][i - kBiasFeat];
    else if (i == kBiasAlpha) v = pp.b[LIN_ALPHA][0];
    else if (i >= kBiasRgb && i < kBiasRgb + 3) v = pp.b[LIN_RGB][i - kBiasRgb];
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

// out[r, j] = b_views[j] + sum_c W_views[j, 256 + c] * enc4(viewdir_r)[c]   (fp32, exact direction term)
constexpr int kVbRays = 16;
__global__ void __launch_bounds__(128) view_bias_kernel(const float* __restrict__ wdir, const float* __restrict__ bdir,
                                                        const float* __restrict__ vd, int64_t stride,
                                                        const float* __restrict__ emb, int64_t n,
                                                        float* __restrict__ out) {
  __shared__ float w[27][128];
  __shared__ float enc[kVbRays][28];
  for (int i = threadIdx.x; i < 128 * 27; i += blockDim.x) w[i % 27][i / 27] = __ldg(wdir + i);
  const float bj = __ldg(bdir + threadIdx.x);
  const int64_t ngroups = (n + kVbRays - 1) / kVbRays;
  for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
    __syncthreads();
    const int64_t r0 = g * kVbRays;
    if (emb != nullptr) {
      for (int i = threadIdx.x; i < kVbRays * 27; i += blockDim.x) {
        const int rr = i / 27, c = i - rr * 27;
        enc[rr][c] = (r0 + rr < n) ? __ldg(emb + (r0 + rr) * GBN_EMB_CH + GBN_PTS_CH + c) : 0.f;
      }
    } else if (threadIdx.x < kVbRays * 3) {
      const int rr = threadIdx.x / 3, ax = threadIdx.x - rr * 3;
      const float x = (r0 + rr < n) ? __ldg(vd + (r0 + rr) * stride + ax) : 0.f;
      float sc[8];
      posenc_axis<4>(x, sc);
      enc[rr][ax] = x;
#pragma unroll
      for (int k = 0; k < 4; ++k) { enc[rr][3 + 6 * k + ax] = sc[2 * k]; enc[rr][6 + 6 * k + ax] = sc[2 * k + 1]; }
    }
    __syncthreads();
    float acc[kVbRays];
#pragma unroll
    for (int rr = 0; rr < kVbRays; ++rr) acc[rr] = bj;
#pragma unroll
    for (int c = 0; c < 27; ++c) {
      const float wc = w[c][threadIdx.x];
#pragma unroll
      for (int rr = 0; rr < kVbRays; ++rr) acc[rr] = fmaf(wc, enc[rr][c], acc[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < kVbRays; ++rr)
      if (r0 + rr < n) out[(r0 + rr) * 128 + threadIdx.x] = acc[rr];
  }
}

int launch_view_bias_raw(const float* wdir, const float* bdir, const float* viewdirs, int64_t stride, const float* emb,
                         int64_t n, float* out, cudaStream_t stream) {
  const int64_t groups = (n + kVbRays - 1) / kVbRays;
  const int grid = (int)(groups < kNumSMs * 8 ? groups : kNumSMs * 8);
  view_bias_kernel<<<grid, 128, 0, stream>>>(wdir, bdir, viewdirs, stride, emb, n, out);
  return check_launch("view_bias_kernel");
}

int launch_view_bias(const uint8_t* packed, const MlpPlan& p, const float* viewdirs, int64_t stride,
                     const float* emb, int64_t n, float* out, cudaStream_t stream) {
  return launch_view_bias_raw(reinterpret_cast<const float*>(packed + p.off_wdir),
                              reinterpret_cast<const float*>(packed + p.off_bdir), viewdirs, stride, emb, n, out, stream);
}

// TMEM-operand bf16 kernels (mlp_ts.cu); GBNERF_MLP_SS=1 selects the shared-memory-operand kernels instead
size_t ts_packed_bytes(int bwd);
int ts_prepack(const void* const* params, void* packed, int bwd, cudaStream_t st);
int ts_adam_repack(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq, double lr,
                   double beta1, double beta2, double eps, int64_t step, void* packed_fwd, void* packed_bwd, cudaStream_t st,
                   const float* dev_scalars = nullptr);
int ts_adam_tick(double* state, const float* lr, double beta1, double beta2, float* scalars, cudaStream_t st);
// bf16 kernel family: 0 = shared-memory operands (mlp_tc.cu), 1 = TMEM operands (mlp_ts.cu, default).
// GBNERF_MLP=ss|ts overrides the default (GBNERF_MLP_SS=1 == ss).
int mlp_variant() {
  static const int v = [] {
    const char* e = getenv("GBNERF_MLP");
    if (e && e[0] == 's') return 0;
    if (e && e[0] == 't' && e[1] == 's') return 1;
    const char* s = getenv("GBNERF_MLP_SS");
    if (s && s[0] == '1') return 0;
    return GBN_DEFAULT_MLP_VARIANT;
  }();
  return v;
}
bool mlp_use_ts() { return mlp_variant() != 0; }

static bool g_pack_init[64];
static std::mutex g_pack_mutex;

}  // namespace gbn

using namespace gbn;

extern "C" int gbn_mlp_prepack_weights(const void* const* params, void* packed, int precision, void* stream) {
  GBN_REQUIRE(precision >= 0 && precision < kNumPlans, "prepack: unknown precision/plan %d", precision);
  GBN_REQUIRE(params && packed, "prepack: null pointer");
  GBN_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 255) == 0, "prepack: packed buffer must be 256-byte aligned");
  for (int i = 0; i < 2 * GBN_NUM_LINEAR; ++i) GBN_REQUIRE(params[i], "prepack: params[%d] is null", i);
  if (precision != GBN_PRECISION_TF32 && mlp_variant() == 1)
    return ts_prepack(params, packed, precision == GBN_PACK_BWD_BF16, (cudaStream_t)stream);
  ParamPtrs pp;
  for (int i = 0; i < GBN_NUM_LINEAR; ++i) {
    GBN_REQUIRE(params[2 * i] && params[2 * i + 1], "prepack: params[%d] is null", 2 * i);
    pp.w[i] = static_cast<const float*>(params[2 * i]);
    pp.b[i] = static_cast<const float*>(params[2 * i + 1]);
  }
  const MlpPlan& p = mlp_plan(precision);
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  GBN_CUDA(cudaGetDevice(&dev));
  GBN_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  {
    std::lock_guard<std::mutex> lk(g_pack_mutex);
    if (!g_pack_init[dev]) {
      for (int pr = 0; pr < kNumPlans; ++pr) {
        const MlpPlan& q = mlp_plan(pr);
        GBN_CUDA(cudaMemcpyToSymbol(c_pack, q.pack.data(), q.pack.size() * sizeof(PackJob),
                                         pr * kMaxJobs * sizeof(PackJob), cudaMemcpyHostToDevice));
      }
      g_pack_init[dev] = true;
    }
  }
  PackHeader hdr{0x4e42476bu, (uint32_t)precision, p.off_bias, p.off_wdir, p.off_bdir, p.total_bytes,
                 (uint32_t)p.jobs.size(), 0};
  const int njobs = (int)p.jobs.size();
  if (precision != GBN_PRECISION_TF32)
    prepack_kernel<GBN_PRECISION_BF16><<<njobs + 1, 256, 0, st>>>(pp, static_cast<uint8_t*>(packed), njobs, hdr, precision);
  else
    prepack_kernel<GBN_PRECISION_TF32><<<njobs + 1, 256, 0, st>>>(pp, static_cast<uint8_t*>(packed), njobs, hdr, precision);
  return check_launch("prepack_kernel");
}

extern "C" int gbn_mlp_variant(void) { return mlp_variant(); }

extern "C" int gbn_adam_step_repack(void* const* params, const void* const* grads, void* const* exp_avg,
                                    void* const* exp_avg_sq, double lr, double beta1, double beta2, double eps, int64_t step,
                                    void* packed_fwd, void* packed_bwd, void* stream) {
  GBN_REQUIRE(params && grads && exp_avg && exp_avg_sq, "adam_step_repack: null pointer table");
  for (int i = 0; i < 2 * GBN_NUM_LINEAR; ++i)
    GBN_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i], "adam_step_repack: tensor %d has a null pointer", i);
  GBN_REQUIRE(step >= 1, "adam_step_repack: step counts from 1 (got %lld)", (long long)step);
  GBN_REQUIRE(lr >= 0 && beta1 >= 0 && beta1 < 1 && beta2 >= 0 && beta2 < 1 && eps >= 0, "adam_step_repack: bad hyper-parameters");
  GBN_REQUIRE((packed_fwd == nullptr && packed_bwd == nullptr) || mlp_variant() == 1,
              "adam_step_repack: in-place re-pack exists for the default bf16 weight images only (pass NULL and re-pack)");
  return ts_adam_repack(params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, packed_fwd, packed_bwd,
                        (cudaStream_t)stream);
}

extern "C" int gbn_adam_tick(double* step_state, const float* lr, double beta1, double beta2, float* scalars, void* stream) {
  GBN_REQUIRE(step_state && lr && scalars, "adam_tick: null pointer");
  GBN_REQUIRE(beta1 >= 0 && beta1 < 1 && beta2 >= 0 && beta2 < 1, "adam_tick: bad betas");
  return ts_adam_tick(step_state, lr, beta1, beta2, scalars, (cudaStream_t)stream);
}

extern "C" int gbn_adam_step_repack_dev(void* const* params, const void* const* grads, void* const* exp_avg,
                                        void* const* exp_avg_sq, const float* scalars, double beta1, double beta2, double eps,
                                        void* packed_fwd, void* packed_bwd, void* stream) {
  GBN_REQUIRE(params && grads && exp_avg && exp_avg_sq && scalars, "adam_step_repack_dev: null pointer");
  for (int i = 0; i < 2 * GBN_NUM_LINEAR; ++i)
    GBN_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i], "adam_step_repack_dev: tensor %d has a null pointer", i);
  GBN_REQUIRE(beta1 >= 0 && beta1 < 1 && beta2 >= 0 && beta2 < 1 && eps >= 0, "adam_step_repack_dev: bad hyper-parameters");
  GBN_REQUIRE((packed_fwd == nullptr && packed_bwd == nullptr) || mlp_variant() == 1,
              "adam_step_repack_dev: in-place re-pack exists for the default bf16 weight images only");
  return ts_adam_repack(params, grads, exp_avg, exp_avg_sq, 0.0, beta1, beta2, eps, 1, packed_fwd, packed_bwd,
                        (cudaStream_t)stream, scalars);
}
