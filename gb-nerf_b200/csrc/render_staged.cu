// Staged (cp.async pipelined) versions of the HBM-bound per-ray kernels.
//
// One warp per ray, but the ray's operands do not go through registers on their way in: every warp keeps a private
// D-deep ring of whole rays in shared memory, filled with cp.async (SASS LDGSTS, 16 B per lane for raw) D-1 rays
// ahead of the one being composited.  Bytes in flight per SM are then set by the ring (2 CTAs x 8 warps x D rays,
// ~160 KB), not by how many warps happen to be stalled on a load, and there is no block-level synchronisation at
// all (cp.async.wait_group + __syncwarp).  Each lane owns K CONSECUTIVE samples of its ray, so one warp shuffle
// scan covers the whole ray and weights leave as vector stores.
// (A bulk-TMA ring was measured first: UBLKCP keeps too few DRAM requests in flight per SM for a cold stream.)
// Reference: raw2outputs, run_nerf_helpers.py:352-406.
#include "common.cuh"

namespace gbn {

constexpr int kStWarps = 8;
constexpr int kStThreads = kStWarps * 32;
constexpr int kStagedSmemBudget = 50 * 1024;            // per CTA -> four CTAs (32 warps) per SM

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// 1 / (1 + e^-x): two MUFU ops, ~2 ulp (the compositing tolerance is 1e-5 relative)
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_approx(1.f + __expf(-x)); }

template <int K, int D, bool NOISE>
__global__ void __launch_bounds__(kStThreads) composite_fwd_staged_kernel(
    const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ noise,
    const float* __restrict__ d, int64_t stride, int64_t R, int S, int white,
    float* __restrict__ rgb, float* __restrict__ disp, float* __restrict__ acc, float* __restrict__ depth,
    float* __restrict__ weights, float* __restrict__ alpha_out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ray_bytes = (uint32_t)S * (NOISE ? 24u : 20u);       // raw | z | noise of one ray
  uint8_t* const ring = smem + (size_t)warp * D * ray_bytes;
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
  const int64_t nwarps = (int64_t)gridDim.x * kStWarps;
  const int64_t w0 = (int64_t)blockIdx.x * kStWarps + warp;
  const int64_t my_rays = w0 < R ? (R - w0 + nwarps - 1) / nwarps : 0;

  auto issue = [&](int64_t i) {   // ray i of this warp -> ring slot i % D
    const int64_t ray = w0 + i * nwarps;
    const uint32_t dst = ring_u32 + (uint32_t)(i % D) * ray_bytes;
    const float4* rsrc = reinterpret_cast<const float4*>(raw) + ray * S;
    for (int s = lane; s < S; s += 32) {
      cp_async16(dst + s * 16, rsrc + s);
      cp_async4(dst + S * 16 + s * 4, z + ray * S + s);
      if (NOISE) cp_async4(dst + S * 20 + s * 4, noise + ray * S + s);
    }
  };
#pragma unroll
  for (int j = 0; j < D - 1; ++j) {
    if (j < my_rays) issue(j);
    cp_async_commit();
  }

  // the ray direction is the one operand that does not come through the ring (strided view of the ray batch):
  // lane j fetches |d| of ray i0 + j for 32 rays at a time, so its latency is paid once per 32 rays
  float dn_lane = 0.f;
  for (int64_t i = 0; i < my_rays; ++i) {
    const int64_t ray = w0 + i * nwarps;
    if (i + D - 1 < my_rays) issue(i + D - 1);
    cp_async_commit();
    if ((i & 31) == 0) {
      const int64_t ii = i + lane;
      if (ii < my_rays) {
        const int64_t r = w0 + ii * nwarps;
        const float dx = __ldg(d + r * stride), dy = __ldg(d + r * stride + 1), dz = __ldg(d + r * stride + 2);
        dn_lane = sqrtf(dx * dx + dy * dy + dz * dz);
      }
    }
    const float dnorm = __shfl_sync(kFullMask, dn_lane, (int)(i & 31));
    cp_async_wait<D - 1>();
    __syncwarp();
    const uint8_t* st = ring + (size_t)(i % D) * ray_bytes;
    const float4* sraw = reinterpret_cast<const float4*>(st);
    const float* sz = reinterpret_cast<const float*>(st + S * 16);
    const float* sn = reinterpret_cast<const float*>(st + S * 20);
    float4 rw[K];
    float zz[K], nz[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {   // lanes past the end re-read the last sample (branch-free); masked below
      const int sidx = min(lane * K + k, S - 1);
      rw[k] = sraw[sidx];
      zz[k] = sz[sidx];
      nz[k] = NOISE ? sn[sidx] : 0.f;
    }
    __syncwarp();   // slot may be refilled by the next iteration's issue

    const float znext_lane = __shfl_down_sync(kFullMask, zz[0], 1);
    float a[K], tl[K];
    float prod = 1.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int sidx = lane * K + k;
      const float zn = (k + 1 < K) ? zz[k + 1 < K ? k + 1 : k] : znext_lane;
      const float dl = (sidx == S - 1) ? 1e10f : (zn - zz[k]);
      const float sigma = fmaxf(rw[k].w + nz[k], 0.f);
      a[k] = (sidx < S) ? 1.f - __expf(-sigma * (dl * dnorm)) : 0.f;
      tl[k] = prod;                                   // exclusive product inside the lane
      prod *= (1.f - a[k]) + 1e-10f;
    }
    const float incl = warp_scan_mul(prod, lane);
    float excl = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) excl = 1.f;

    float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
    float w[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      w[k] = a[k] * (excl * tl[k]);
      sr = fmaf(w[k], fast_sigmoid(rw[k].x), sr);
      sg = fmaf(w[k], fast_sigmoid(rw[k].y), sg);
      sb = fmaf(w[k], fast_sigmoid(rw[k].z), sb);
      sd = fmaf(w[k], zz[k], sd);
      sa += w[k];
    }
    // weights (and alpha): K consecutive floats per lane
    float* wrow = weights + ray * S;
    float* arow = alpha_out ? alpha_out + ray * S : nullptr;
    if (K == 4 && (S & 3) == 0) {
      if (lane * 4 < S) {
        st_stream4(reinterpret_cast<float4*>(wrow) + lane, make_float4(w[0], w[1], w[2], w[3]));
        if (arow) st_stream4(reinterpret_cast<float4*>(arow) + lane, make_float4(a[0], a[1], a[2], a[3]));
      }
    } else if (K == 2 && (S & 1) == 0) {
      if (lane * 2 < S) {
        *reinterpret_cast<float2*>(wrow + lane * 2) = make_float2(w[0], w[1]);
        if (arow) *reinterpret_cast<float2*>(arow + lane * 2) = make_float2(a[0], a[1]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (lane * K + k < S) { st_stream(wrow + lane * K + k, w[k]); if (arow) st_stream(arow + lane * K + k, a[k]); }
    }
    sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd); sa = warp_sum(sa);
    if (lane == 0) {
      const float bg = white ? (1.f - sa) : 0.f;
      rgb[ray * 3 + 0] = sr + bg;
      rgb[ray * 3 + 1] = sg + bg;
      rgb[ray * 3 + 2] = sb + bg;
      const float q = __fdividef(sd, sa);   // NaN when acc == 0, as torch.max propagates the NaN of 0/0
      disp[ray] = __fdividef(1.f, (q != q) ? q : fmaxf(1e-10f, q));
      acc[ray] = sa;
      depth[ray] = sd;
    }
  }
}

// Launches the staged kernel when the shapes allow (S <= 256, 16-byte aligned raw); returns how many rays it took
// (0 = the caller uses the generic kernel).
int64_t launch_composite_fwd_staged(const float* raw, const float* z, const float* rays_d, int64_t ray_stride,
                                    const float* noise, int64_t R, int S, int white, float* rgb, float* disp, float* acc,
                                    float* depth, float* weights, float* alpha, cudaStream_t stream, int* rc) {
  *rc = GBN_OK;
  const int K = (S + 31) / 32;
  const uintptr_t al = reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(weights) | reinterpret_cast<uintptr_t>(alpha);
  if (R < 1 || S < 2 || K > 8 || K == 5 || K == 7 || (al & 15) != 0) return 0;
  const size_t ray_bytes = (size_t)S * (noise ? 24 : 20);
  const int dmax = (int)(kStagedSmemBudget / (kStWarps * ray_bytes));
  const int D = dmax >= 8 ? 8 : (dmax >= 4 ? 4 : 2);
  if (dmax < 2) return 0;
  const size_t smem = (size_t)kStWarps * D * ray_bytes;
  const int64_t blocks = (R + kStWarps - 1) / kStWarps, cap = (int64_t)kNumSMs * 2;
  const int grid = (int)(blocks < cap ? blocks : cap);
#define GBN_ST_LAUNCH(KK, DD, NN)                                                                                   \
  do {                                                                                                              \
    static bool attr = false;                                                                                       \
    if (!attr) {                                                                                                    \
      cudaFuncSetAttribute(composite_fwd_staged_kernel<KK, DD, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                           kStagedSmemBudget);                                                                      \
      attr = true;                                                                                                  \
    }                                                                                                               \
    composite_fwd_staged_kernel<KK, DD, NN><<<grid, kStThreads, smem, stream>>>(                                    \
        raw, z, noise, rays_d, ray_stride, R, S, white, rgb, disp, acc, depth, weights, alpha);                     \
  } while (0)
#define GBN_ST_D(KK)                                                                                                \
  do {                                                                                                              \
    if (noise) { if (D == 8) GBN_ST_LAUNCH(KK, 8, true); else if (D == 4) GBN_ST_LAUNCH(KK, 4, true); else GBN_ST_LAUNCH(KK, 2, true); } \
    else { if (D == 8) GBN_ST_LAUNCH(KK, 8, false); else if (D == 4) GBN_ST_LAUNCH(KK, 4, false); else GBN_ST_LAUNCH(KK, 2, false); }    \
  } while (0)
  switch (K) {
    case 1: GBN_ST_D(1); break;
    case 2: GBN_ST_D(2); break;
    case 3: GBN_ST_D(3); break;
    case 4: GBN_ST_D(4); break;
    case 6: GBN_ST_D(6); break;
    default: GBN_ST_D(8); break;
  }
#undef GBN_ST_D
#undef GBN_ST_LAUNCH
  *rc = check_launch("composite_fwd_staged_kernel");
  return R;
}

}  // namespace gbn
