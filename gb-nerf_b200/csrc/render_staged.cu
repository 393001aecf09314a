// Staged (cp.async pipelined) versions of the HBM-bound per-ray kernels.
//
// One warp per ray, but the ray's operands do not go through registers on their way in: every warp keeps a private
// D-deep ring of whole rays in shared memory, filled with cp.async (SASS LDGSTS, 16 B per lane for raw) D-1 rays
// ahead of the one being composited.  Bytes in flight per SM are then set by the ring (4 CTAs x 8 warps x D rays,
// ~200 KB), not by how many warps happen to be stalled on a load, and there is no block-level synchronisation at
// all (cp.async.wait_group + __syncwarp).  Each lane owns K CONSECUTIVE samples of its ray, so one warp shuffle
// scan covers the whole ray and weights leave as vector stores.
// (A bulk-TMA ring was measured first: UBLKCP keeps too few DRAM requests in flight per SM for a cold stream.)
// Reference: raw2outputs, run_nerf_helpers.py:352-406.
#include <stdlib.h>

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

// five per-lane partial sums -> totals, with 10 shuffles instead of 25: the butterfly halves the number of live values
// at each of its first three steps (which value a lane carries is decided by its lane bits 16, 8, 4), two more steps
// finish the sum.  Afterwards lane 0 holds sum(a), lane 4: sum(b), lane 8: sum(c), lane 12: sum(d), lane 16: sum(e).
__device__ __forceinline__ float warp_sum5(float a, float b, float c, float d, float e, int lane) {
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
  // step 16: lower half keeps (a, b, c, d), upper half keeps (e, 0, 0, 0)
  const float xa = __shfl_xor_sync(kFullMask, u16 ? a : e, 16);
  const float xb = __shfl_xor_sync(kFullMask, b, 16), xc = __shfl_xor_sync(kFullMask, c, 16), xd = __shfl_xor_sync(kFullMask, d, 16);
  float v0 = (u16 ? e : a) + xa, v1 = u16 ? 0.f : b + xb, v2 = u16 ? 0.f : c + xc, v3 = u16 ? 0.f : d + xd;
  // step 8: bit 8 clear keeps (v0, v1), set keeps (v2, v3)
  const float y0 = __shfl_xor_sync(kFullMask, u8 ? v0 : v2, 8), y1 = __shfl_xor_sync(kFullMask, u8 ? v1 : v3, 8);
  v0 = (u8 ? v2 : v0) + y0; v1 = (u8 ? v3 : v1) + y1;
  // step 4: bit 4 clear keeps v0, set keeps v1
  const float w = __shfl_xor_sync(kFullMask, u4 ? v0 : v1, 4);
  v0 = (u4 ? v1 : v0) + w;
  v0 += __shfl_xor_sync(kFullMask, v0, 2);
  v0 += __shfl_xor_sync(kFullMask, v0, 1);
  return v0;
}

template <int K, int D, bool NOISE, bool FULL>
__global__ void __launch_bounds__(kStThreads) composite_fwd_staged_kernel(
    const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ noise,
    const float* __restrict__ d, int64_t stride, int64_t R, int S_rt, int white,
    float* __restrict__ rgb, float* __restrict__ disp, float* __restrict__ acc, float* __restrict__ depth,
    float* __restrict__ weights, float* __restrict__ alpha_out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = FULL ? 32 * K : S_rt;                                  // FULL: every lane owns exactly K samples
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
    if (FULL && (K % 4 == 0 || K == 2)) {   // rows of z / noise are multiples of 16 bytes: vector copies
#pragma unroll
      for (int k = 0; k < K; ++k) cp_async16(dst + (k * 32 + lane) * 16, rsrc + k * 32 + lane);
#pragma unroll
      for (int k = 0; k < (K + 3) / 4; ++k)
        if (K % 4 == 0 || lane < 16) {
          cp_async16(dst + S * 16 + (k * 32 + lane) * 16, reinterpret_cast<const float4*>(z + ray * S) + k * 32 + lane);
          if (NOISE) cp_async16(dst + S * 20 + (k * 32 + lane) * 16, reinterpret_cast<const float4*>(noise + ray * S) + k * 32 + lane);
        }
    } else {
      for (int s = lane; s < S; s += 32) {
        cp_async16(dst + s * 16, rsrc + s);
        cp_async4(dst + S * 16 + s * 4, z + ray * S + s);
        if (NOISE) cp_async4(dst + S * 20 + s * 4, noise + ray * S + s);
      }
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
    if (FULL && K % 4 == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) rw[k] = sraw[lane * K + k];
#pragma unroll
      for (int k = 0; k < K; k += 4) {
        const float4 zq = reinterpret_cast<const float4*>(sz)[(lane * K + k) >> 2];
        zz[k] = zq.x; zz[k + 1] = zq.y; zz[k + 2] = zq.z; zz[k + 3] = zq.w;
        if (NOISE) {
          const float4 nq = reinterpret_cast<const float4*>(sn)[(lane * K + k) >> 2];
          nz[k] = nq.x; nz[k + 1] = nq.y; nz[k + 2] = nq.z; nz[k + 3] = nq.w;
        } else nz[k] = nz[k + 1] = nz[k + 2] = nz[k + 3] = 0.f;
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) {   // lanes past the end re-read the last sample (branch-free); masked below
        const int sidx = FULL ? lane * K + k : min(lane * K + k, S - 1);
        rw[k] = sraw[sidx];
        zz[k] = sz[sidx];
        nz[k] = NOISE ? sn[sidx] : 0.f;
      }
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
      a[k] = (FULL || sidx < S) ? 1.f - __expf(-sigma * (dl * dnorm)) : 0.f;
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
    // totals: lane 0 <- acc, lane 4 <- red, lane 8 <- green, lane 12 <- blue, lane 16 <- depth (warp_sum5's layout)
    const float tot = warp_sum5(sa, sr, sg, sb, sd, lane);
    const float tacc = __shfl_sync(kFullMask, tot, 0);
    if ((lane & 3) == 0 && lane <= 16) {
      const float bg = white ? (1.f - tacc) : 0.f;
      if (lane == 0) acc[ray] = tot;
      else if (lane == 4) rgb[ray * 3 + 0] = tot + bg;
      else if (lane == 8) rgb[ray * 3 + 1] = tot + bg;
      else if (lane == 12) rgb[ray * 3 + 2] = tot + bg;
      else {
        const float q = __fdividef(tot, tacc);   // NaN when acc == 0, as torch.max propagates the NaN of 0/0
        disp[ray] = __fdividef(1.f, (q != q) ? q : fmaxf(1e-10f, q));
        depth[ray] = tot;
      }
    }
  }
}

// ---- S = 64 (the coarse pass): two rays per warp -------------------------------------------------------------------
// With 64 samples a lane would own only two, and the per-ray fixed cost (scan, reductions, outputs: ~150 of the 316
// warp instructions per ray) dominates.  Two CONSECUTIVE rays are therefore staged as one 128-sample pseudo-ray (their
// raw / z / noise rows are contiguous in memory), lanes 0-15 composite the first and lanes 16-31 the second with
// 16-lane segmented scans and reductions: the warp does the work of one 128-sample ray for two 64-sample rays.
template <int D, bool NOISE>
__global__ void __launch_bounds__(kStThreads) composite_fwd_pair64_kernel(
    const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ noise,
    const float* __restrict__ d, int64_t stride, int64_t R, int white,
    float* __restrict__ rgb, float* __restrict__ disp, float* __restrict__ acc, float* __restrict__ depth,
    float* __restrict__ weights, float* __restrict__ alpha_out) {
  constexpr int S = 64, P = 128, K = 4;                          // P: samples of the pseudo-ray
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 4, l16 = lane & 15;
  constexpr uint32_t pair_bytes = (uint32_t)P * (NOISE ? 24u : 20u);
  uint8_t* const ring = smem + (size_t)warp * D * pair_bytes;
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
  const int64_t npairs = (R + 1) >> 1;
  const int64_t nwarps = (int64_t)gridDim.x * kStWarps;
  const int64_t w0 = (int64_t)blockIdx.x * kStWarps + warp;
  const int64_t my_pairs = w0 < npairs ? (npairs - w0 + nwarps - 1) / nwarps : 0;

  auto issue = [&](int64_t i) {
    const int64_t pr = w0 + i * nwarps;
    const bool both = 2 * pr + 1 < R;                            // the last pair of an odd batch has one ray
    const uint32_t dst = ring_u32 + (uint32_t)(i % D) * pair_bytes;
    const float4* rsrc = reinterpret_cast<const float4*>(raw) + pr * P;
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (both || k < 2) cp_async16(dst + (k * 32 + lane) * 16, rsrc + k * 32 + lane);
    if (both || lane < 16) {
      cp_async16(dst + P * 16 + lane * 16, reinterpret_cast<const float4*>(z + pr * P) + lane);
      if (NOISE) cp_async16(dst + P * 20 + lane * 16, reinterpret_cast<const float4*>(noise + pr * P) + lane);
    }
  };
#pragma unroll
  for (int j = 0; j < D - 1; ++j) {
    if (j < my_pairs) issue(j);
    cp_async_commit();
  }
  float dn_lane = 0.f;   // |d| of ray 2 * (pair i0 + lane / 2) + (lane & 1), refreshed every 16 pairs
  for (int64_t i = 0; i < my_pairs; ++i) {
    const int64_t pr = w0 + i * nwarps;
    const int64_t ray = 2 * pr + sub;
    const bool valid = ray < R;
    if (i + D - 1 < my_pairs) issue(i + D - 1);
    cp_async_commit();
    if ((i & 15) == 0) {
      const int64_t ii = i + (lane >> 1);
      const int64_t r = 2 * (w0 + ii * nwarps) + (lane & 1);
      dn_lane = 0.f;
      if (ii < my_pairs && r < R) {
        const float dx = __ldg(d + r * stride), dy = __ldg(d + r * stride + 1), dz = __ldg(d + r * stride + 2);
        dn_lane = sqrtf(dx * dx + dy * dy + dz * dz);
      }
    }
    const float dnorm = __shfl_sync(kFullMask, dn_lane, 2 * (int)(i & 15) + sub);
    cp_async_wait<D - 1>();
    __syncwarp();
    const uint8_t* st = ring + (size_t)(i % D) * pair_bytes;
    float4 rw[K];
#pragma unroll
    for (int k = 0; k < K; ++k) rw[k] = reinterpret_cast<const float4*>(st)[lane * K + k];
    const float4 zq = reinterpret_cast<const float4*>(st + P * 16)[lane];
    float4 nq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (NOISE) nq = reinterpret_cast<const float4*>(st + P * 20)[lane];
    __syncwarp();   // slot may be refilled by the next iteration's issue
    const float zz[K] = {zq.x, zq.y, zq.z, zq.w}, nz[K] = {nq.x, nq.y, nq.z, nq.w};
    const float znext_lane = __shfl_down_sync(kFullMask, zz[0], 1);
    float a[K], tl[K];
    float prod = 1.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float zn = (k + 1 < K) ? zz[k + 1 < K ? k + 1 : k] : znext_lane;
      const float dl = (k == K - 1 && l16 == 15) ? 1e10f : (zn - zz[k]);      // the ray's last sample (helpers:369-372)
      const float sigma = fmaxf(rw[k].w + nz[k], 0.f);
      a[k] = valid ? 1.f - __expf(-sigma * (dl * dnorm)) : 0.f;
      tl[k] = prod;
      prod *= (1.f - a[k]) + 1e-10f;
    }
    float incl = prod;   // inclusive product scan inside each 16-lane half
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const float t = __shfl_up_sync(kFullMask, incl, o);
      if (l16 >= o) incl *= t;
    }
    float excl = __shfl_up_sync(kFullMask, incl, 1);
    if (l16 == 0) excl = 1.f;
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
    if (valid) {
      st_stream4(reinterpret_cast<float4*>(weights + pr * P) + lane, make_float4(w[0], w[1], w[2], w[3]));
      if (alpha_out) st_stream4(reinterpret_cast<float4*>(alpha_out + pr * P) + lane, make_float4(a[0], a[1], a[2], a[3]));
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {   // xor offsets < 16 stay inside the half
      sr += __shfl_xor_sync(kFullMask, sr, o); sg += __shfl_xor_sync(kFullMask, sg, o); sb += __shfl_xor_sync(kFullMask, sb, o);
      sd += __shfl_xor_sync(kFullMask, sd, o); sa += __shfl_xor_sync(kFullMask, sa, o);
    }
    if (l16 == 0 && valid) {
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

// ---- compositing backward, staged ---------------------------------------------------------------------------------
// Same staging and ownership as the forward (a lane owns K consecutive samples of its ray, operands arrive through a
// per-warp cp.async ring), so the transmittance is one forward product scan and the "sum of G_k w_k over later samples"
// one reverse sum scan per ray; the generic kernel (lane-strided samples) needs one pair of scans per 32-sample chunk
// and ran 1017 warp instructions per 128-sample ray.  Gradient formulas: SURVEY §8a row 13 (checked against autograd
// of the reference raw2outputs).  The forward quantities are recomputed, nothing is stashed.
template <int K, int D, bool NOISE, bool GW>
__global__ void __launch_bounds__(kStThreads, K <= 2 ? 4 : 3) composite_bwd_staged_kernel(
    const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ noise,
    const float* __restrict__ d, int64_t stride, int64_t R, int white, int detach_w,
    const float* __restrict__ g_rgb, const float* __restrict__ g_disp, const float* __restrict__ g_acc,
    const float* __restrict__ g_depth, const float* __restrict__ g_w, float* __restrict__ g_raw) {
  constexpr int S = 32 * K;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t ray_bytes = (uint32_t)S * (20u + (NOISE ? 4u : 0u) + (GW ? 4u : 0u));   // raw | z | noise | g_w
  constexpr uint32_t off_z = S * 16, off_n = S * 20, off_g = S * (NOISE ? 24 : 20);
  uint8_t* const ring = smem + (size_t)warp * D * ray_bytes;
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
  const int64_t nwarps = (int64_t)gridDim.x * kStWarps;
  const int64_t w0 = (int64_t)blockIdx.x * kStWarps + warp;
  const int64_t my_rays = w0 < R ? (R - w0 + nwarps - 1) / nwarps : 0;

  auto issue = [&](int64_t i) {
    const int64_t ray = w0 + i * nwarps;
    const uint32_t dst = ring_u32 + (uint32_t)(i % D) * ray_bytes;
    const float4* rsrc = reinterpret_cast<const float4*>(raw) + ray * S;
#pragma unroll
    for (int k = 0; k < K; ++k) cp_async16(dst + (k * 32 + lane) * 16, rsrc + k * 32 + lane);
    if (K % 4 == 0 || lane < 8 * K) {      // S floats = S/4 sixteen-byte pieces
#pragma unroll
      for (int k = 0; k < (K + 3) / 4; ++k) {
        cp_async16(dst + off_z + (k * 32 + lane) * 16, reinterpret_cast<const float4*>(z + ray * S) + k * 32 + lane);
        if (NOISE) cp_async16(dst + off_n + (k * 32 + lane) * 16, reinterpret_cast<const float4*>(noise + ray * S) + k * 32 + lane);
        if (GW) cp_async16(dst + off_g + (k * 32 + lane) * 16, reinterpret_cast<const float4*>(g_w + ray * S) + k * 32 + lane);
      }
    }
  };
#pragma unroll
  for (int j = 0; j < D - 1; ++j) {
    if (j < my_rays) issue(j);
    cp_async_commit();
  }
  // per-ray scalars (|d| and the six incoming gradients) are fetched for 32 rays at a time, one ray per lane
  float dn_l = 0.f, gr_l = 0.f, gg_l = 0.f, gb_l = 0.f, gdisp_l = 0.f, gacc_l = 0.f, gdep_l = 0.f;
  for (int64_t i = 0; i < my_rays; ++i) {
    const int64_t ray = w0 + i * nwarps;
    if (i + D - 1 < my_rays) issue(i + D - 1);
    cp_async_commit();
    if ((i & 31) == 0) {
      const int64_t ii = i + lane;
      if (ii < my_rays) {
        const int64_t r = w0 + ii * nwarps;
        const float dx = __ldg(d + r * stride), dy = __ldg(d + r * stride + 1), dz = __ldg(d + r * stride + 2);
        dn_l = sqrtf(dx * dx + dy * dy + dz * dz);
        gr_l = g_rgb ? __ldg(g_rgb + r * 3) : 0.f; gg_l = g_rgb ? __ldg(g_rgb + r * 3 + 1) : 0.f; gb_l = g_rgb ? __ldg(g_rgb + r * 3 + 2) : 0.f;
        gdisp_l = g_disp ? __ldg(g_disp + r) : 0.f; gacc_l = g_acc ? __ldg(g_acc + r) : 0.f; gdep_l = g_depth ? __ldg(g_depth + r) : 0.f;
      }
    }
    const int src = (int)(i & 31);
    const float dnorm = __shfl_sync(kFullMask, dn_l, src);
    const float gr = __shfl_sync(kFullMask, gr_l, src), gg = __shfl_sync(kFullMask, gg_l, src), gb = __shfl_sync(kFullMask, gb_l, src);
    const float gdisp = __shfl_sync(kFullMask, gdisp_l, src);
    float gacc = __shfl_sync(kFullMask, gacc_l, src), gdep = __shfl_sync(kFullMask, gdep_l, src);
    cp_async_wait<D - 1>();
    __syncwarp();
    const uint8_t* st = ring + (size_t)(i % D) * ray_bytes;
    float4 rw[K];
    float zz[K], nz[K], gw[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      rw[k] = reinterpret_cast<const float4*>(st)[lane * K + k];
      zz[k] = reinterpret_cast<const float*>(st + off_z)[lane * K + k];
      nz[k] = NOISE ? reinterpret_cast<const float*>(st + off_n)[lane * K + k] : 0.f;
      gw[k] = GW ? reinterpret_cast<const float*>(st + off_g)[lane * K + k] : 0.f;
    }
    __syncwarp();   // slot may be refilled by the next iteration's issue
    // ---- forward again: alpha, e = exp(-sigma delta), delta, transmittance T -------------------------------------------
    const float znext_lane = __shfl_down_sync(kFullMask, zz[0], 1);
    float a[K], e[K], dlt[K], T[K];
    float prod = 1.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float zn = (k + 1 < K) ? zz[k + 1 < K ? k + 1 : k] : znext_lane;
      const float dl = (lane == 31 && k == K - 1) ? 1e10f : (zn - zz[k]);
      dlt[k] = dl * dnorm;
      const float sigma = fmaxf(rw[k].w + nz[k], 0.f);
      e[k] = __expf(-sigma * dlt[k]);
      a[k] = 1.f - e[k];
      T[k] = prod;
      prod *= (1.f - a[k]) + 1e-10f;
    }
    const float incl = warp_scan_mul(prod, lane);
    float excl = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) excl = 1.f;
    float sd = 0.f, sa = 0.f;
    float w[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      T[k] *= excl;
      w[k] = a[k] * T[k];
      sd = fmaf(w[k], zz[k], sd);
      sa += w[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sd += __shfl_xor_sync(kFullMask, sd, o); sa += __shfl_xor_sync(kFullMask, sa, o); }
    // disp = 1 / max(1e-10, depth / acc): above the clamp disp = acc / depth
    if (gdisp != 0.f) {
      const float q = sd / sa;
      if (!(q <= 1e-10f)) {   // also taken for NaN, which then propagates as in autograd
        gdep += -gdisp * sa / (sd * sd);
        gacc += gdisp / sd;
      }
    }
    if (white) gacc -= gr + gg + gb;
    // ---- G_k, suffix sums of G_k w_k, gradients --------------------------------------------------------------------
    constexpr bool kKeepColours = K <= 2;   // K = 4: recomputing three sigmoids is cheaper than 12 more live registers
    float G[K], suf[K], c3[kKeepColours ? K : 1][3];
    float run = 0.f;
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
      const float cr = fast_sigmoid(rw[k].x), cg = fast_sigmoid(rw[k].y), cb = fast_sigmoid(rw[k].z);
      if (kKeepColours) { c3[k][0] = cr; c3[k][1] = cg; c3[k][2] = cb; }
      G[k] = gdep * zz[k] + gacc + gw[k];
      if (!detach_w) G[k] += gr * cr + gg * cg + gb * cb;
      suf[k] = run;                 // later samples inside this lane
      run = fmaf(G[k], w[k], run);
    }
    const float rincl = warp_rscan_add(run, lane);
    float later = __shfl_down_sync(kFullMask, rincl, 1);   // lanes above this one
    if (lane == 31) later = 0.f;
    float4* out = reinterpret_cast<float4*>(g_raw) + ray * S + lane * K;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float f = (1.f - a[k]) + 1e-10f;
      const float dalpha = G[k] * T[k] - __fdividef(suf[k] + later, f);
      const float pre = rw[k].w + nz[k];
      const float cr = kKeepColours ? c3[k][0] : fast_sigmoid(rw[k].x), cg = kKeepColours ? c3[k][1] : fast_sigmoid(rw[k].y),
                  cb = kKeepColours ? c3[k][2] : fast_sigmoid(rw[k].z);
      float4 o;
      o.x = w[k] * gr * cr * (1.f - cr);
      o.y = w[k] * gg * cg * (1.f - cg);
      o.z = w[k] * gb * cb * (1.f - cb);
      o.w = (pre > 0.f) ? dalpha * dlt[k] * e[k] : 0.f;
      st_stream4(out + k, o);
    }
  }
}

// ---- long rays: S = 128 M (M = 2 .. 8), e.g. the fine pass of BASELINE configs[3] (128 + 256 = 384 samples) --------------
// A lane that owned S / 32 consecutive samples would need ~110 registers at S = 384.  Here the ring slot still holds the
// WHOLE ray (cp.async, D rays deep per warp), but the warp walks it in M segments of 128 samples with the K = 4
// register footprint: inside a segment a lane owns 4 consecutive samples, one shuffle product scan covers the segment,
// and the transmittance at the segment's start is a carried scalar.  The backward walks the segments forward once
// (transmittance carries, the lane's exclusive in-segment products, depth and acc totals) and then in reverse with a
// carried suffix sum: per 128 samples the same one product scan + one reverse sum scan as composite_bwd_staged_kernel.
constexpr int kSegWarps = 4;
constexpr int kSegThreads = kSegWarps * 32;

template <int M, int D, bool NOISE>
__global__ void __launch_bounds__(kSegThreads) composite_fwd_seg_kernel(
    const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ noise,
    const float* __restrict__ d, int64_t stride, int64_t R, int white,
    float* __restrict__ rgb, float* __restrict__ disp, float* __restrict__ acc, float* __restrict__ depth,
    float* __restrict__ weights, float* __restrict__ alpha_out) {
  constexpr int S = 128 * M;
  constexpr uint32_t ray_bytes = (uint32_t)S * (NOISE ? 24u : 20u);       // raw | z | noise of one ray
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* const ring = smem + (size_t)warp * D * ray_bytes;
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
  const int64_t nwarps = (int64_t)gridDim.x * kSegWarps;
  const int64_t w0 = (int64_t)blockIdx.x * kSegWarps + warp;
  const int64_t my_rays = w0 < R ? (R - w0 + nwarps - 1) / nwarps : 0;

  auto issue = [&](int64_t i) {
    const int64_t ray = w0 + i * nwarps;
    const uint32_t dst = ring_u32 + (uint32_t)(i % D) * ray_bytes;
    const float4* rsrc = reinterpret_cast<const float4*>(raw) + ray * S;
#pragma unroll
    for (int k = 0; k < 4 * M; ++k) cp_async16(dst + (k * 32 + lane) * 16, rsrc + k * 32 + lane);
#pragma unroll
    for (int k = 0; k < M; ++k) {
      cp_async16(dst + S * 16 + (k * 32 + lane) * 16, reinterpret_cast<const float4*>(z + ray * S) + k * 32 + lane);
      if (NOISE) cp_async16(dst + S * 20 + (k * 32 + lane) * 16, reinterpret_cast<const float4*>(noise + ray * S) + k * 32 + lane);
    }
  };
#pragma unroll
  for (int j = 0; j < D - 1; ++j) {
    if (j < my_rays) issue(j);
    cp_async_commit();
  }
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
    float carry = 1.f;                                   // transmittance at the start of the segment
    float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
    float* wrow = weights + ray * S;
    float* arow = alpha_out ? alpha_out + ray * S : nullptr;
#pragma unroll
    for (int seg = 0; seg < M; ++seg) {
      const int base = seg * 128 + lane * 4;
      float4 rw[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) rw[k] = sraw[base + k];
      const float4 zq = reinterpret_cast<const float4*>(sz)[base >> 2];
      float4 nq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (NOISE) nq = reinterpret_cast<const float4*>(sn)[base >> 2];
      const float zz[4] = {zq.x, zq.y, zq.z, zq.w}, nz[4] = {nq.x, nq.y, nq.z, nq.w};
      float znext = __shfl_down_sync(kFullMask, zz[0], 1);
      if (lane == 31 && seg + 1 < M) znext = sz[base + 4];          // first depth of the next segment
      float a[4], tl[4];
      float prod = 1.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float zn = k < 3 ? zz[k < 3 ? k + 1 : k] : znext;
        const float dl = (seg == M - 1 && lane == 31 && k == 3) ? 1e10f : (zn - zz[k]);
        const float sigma = fmaxf(rw[k].w + nz[k], 0.f);
        a[k] = 1.f - __expf(-sigma * (dl * dnorm));
        tl[k] = prod;
        prod *= (1.f - a[k]) + 1e-10f;
      }
      const float incl = warp_scan_mul(prod, lane);
      float excl = __shfl_up_sync(kFullMask, incl, 1);
      if (lane == 0) excl = 1.f;
      excl *= carry;
      float w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        w[k] = a[k] * (excl * tl[k]);
        sr = fmaf(w[k], fast_sigmoid(rw[k].x), sr);
        sg = fmaf(w[k], fast_sigmoid(rw[k].y), sg);
        sb = fmaf(w[k], fast_sigmoid(rw[k].z), sb);
        sd = fmaf(w[k], zz[k], sd);
        sa += w[k];
      }
      st_stream4(reinterpret_cast<float4*>(wrow + base), make_float4(w[0], w[1], w[2], w[3]));
      if (arow) st_stream4(reinterpret_cast<float4*>(arow + base), make_float4(a[0], a[1], a[2], a[3]));
      carry *= __shfl_sync(kFullMask, incl, 31);
    }
    __syncwarp();   // slot may be refilled by the next iteration's issue
    const float tot = warp_sum5(sa, sr, sg, sb, sd, lane);
    const float tacc = __shfl_sync(kFullMask, tot, 0);
    if ((lane & 3) == 0 && lane <= 16) {
      const float bg = white ? (1.f - tacc) : 0.f;
      if (lane == 0) acc[ray] = tot;
      else if (lane == 4) rgb[ray * 3 + 0] = tot + bg;
      else if (lane == 8) rgb[ray * 3 + 1] = tot + bg;
      else if (lane == 12) rgb[ray * 3 + 2] = tot + bg;
      else {
        const float q = __fdividef(tot, tacc);   // NaN when acc == 0, as torch.max propagates the NaN of 0/0
        disp[ray] = __fdividef(1.f, (q != q) ? q : fmaxf(1e-10f, q));
        depth[ray] = tot;
      }
    }
  }
}

// The raw [S, 4] block of a ray is stored TRANSPOSED within each 128-sample segment (slot 32 c + l for sample 4 l + c): a
// lane owns four consecutive samples, and with the row-major image its 16-byte reads are 64 bytes apart from lane to
// lane - a 4-way bank conflict on every float4 and 16-way on the scalar .w reads (shared-memory pipe 68 % busy in ncu,
// 33 % with this placement).  The bulk-copy side stays conflict-free and coalesced.  The kernel itself gained only 2 %:
// at 384 samples it holds 12 warps per SM (two whole rays per warp in shared memory) and is bound by the latency of its
// two passes over the ray, not by a pipe.
template <int M, int D, bool NOISE, bool GW>
__global__ void __launch_bounds__(kSegThreads) composite_bwd_seg_kernel(
    const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ noise,
    const float* __restrict__ d, int64_t stride, int64_t R, int white, int detach_w,
    const float* __restrict__ g_rgb, const float* __restrict__ g_disp, const float* __restrict__ g_acc,
    const float* __restrict__ g_depth, const float* __restrict__ g_w, float* __restrict__ g_raw) {
  constexpr int S = 128 * M;
  constexpr uint32_t ray_bytes = (uint32_t)S * (20u + (NOISE ? 4u : 0u) + (GW ? 4u : 0u));   // raw | z | noise | g_w
  constexpr uint32_t off_z = S * 16, off_n = S * 20, off_g = S * (NOISE ? 24 : 20);
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* const ring = smem + (size_t)warp * D * ray_bytes;
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
  const int64_t nwarps = (int64_t)gridDim.x * kSegWarps;
  const int64_t w0 = (int64_t)blockIdx.x * kSegWarps + warp;
  const int64_t my_rays = w0 < R ? (R - w0 + nwarps - 1) / nwarps : 0;

  auto issue = [&](int64_t i) {
    const int64_t ray = w0 + i * nwarps;
    const uint32_t dst = ring_u32 + (uint32_t)(i % D) * ray_bytes;
    const float4* rsrc = reinterpret_cast<const float4*>(raw) + ray * S;
#pragma unroll
    for (int k = 0; k < 4 * M; ++k)   // transposed within each 128-sample segment: sample 4 l + c of a segment -> slot 32 c + l
      cp_async16(dst + ((k >> 2) * 128 + (lane & 3) * 32 + (k & 3) * 8 + (lane >> 2)) * 16, rsrc + k * 32 + lane);
#pragma unroll
    for (int k = 0; k < M; ++k) {
      cp_async16(dst + off_z + (k * 32 + lane) * 16, reinterpret_cast<const float4*>(z + ray * S) + k * 32 + lane);
      if (NOISE) cp_async16(dst + off_n + (k * 32 + lane) * 16, reinterpret_cast<const float4*>(noise + ray * S) + k * 32 + lane);
      if (GW) cp_async16(dst + off_g + (k * 32 + lane) * 16, reinterpret_cast<const float4*>(g_w + ray * S) + k * 32 + lane);
    }
  };
#pragma unroll
  for (int j = 0; j < D - 1; ++j) {
    if (j < my_rays) issue(j);
    cp_async_commit();
  }
  float dn_l = 0.f, gr_l = 0.f, gg_l = 0.f, gb_l = 0.f, gdisp_l = 0.f, gacc_l = 0.f, gdep_l = 0.f;
  for (int64_t i = 0; i < my_rays; ++i) {
    const int64_t ray = w0 + i * nwarps;
    if (i + D - 1 < my_rays) issue(i + D - 1);
    cp_async_commit();
    if ((i & 31) == 0) {
      const int64_t ii = i + lane;
      if (ii < my_rays) {
        const int64_t r = w0 + ii * nwarps;
        const float dx = __ldg(d + r * stride), dy = __ldg(d + r * stride + 1), dz = __ldg(d + r * stride + 2);
        dn_l = sqrtf(dx * dx + dy * dy + dz * dz);
        gr_l = g_rgb ? __ldg(g_rgb + r * 3) : 0.f; gg_l = g_rgb ? __ldg(g_rgb + r * 3 + 1) : 0.f; gb_l = g_rgb ? __ldg(g_rgb + r * 3 + 2) : 0.f;
        gdisp_l = g_disp ? __ldg(g_disp + r) : 0.f; gacc_l = g_acc ? __ldg(g_acc + r) : 0.f; gdep_l = g_depth ? __ldg(g_depth + r) : 0.f;
      }
    }
    const int src = (int)(i & 31);
    const float dnorm = __shfl_sync(kFullMask, dn_l, src);
    const float gr = __shfl_sync(kFullMask, gr_l, src), gg = __shfl_sync(kFullMask, gg_l, src), gb = __shfl_sync(kFullMask, gb_l, src);
    const float gdisp = __shfl_sync(kFullMask, gdisp_l, src);
    float gacc = __shfl_sync(kFullMask, gacc_l, src), gdep = __shfl_sync(kFullMask, gdep_l, src);
    cp_async_wait<D - 1>();
    __syncwarp();
    const uint8_t* st = ring + (size_t)(i % D) * ray_bytes;
    const float4* sraw = reinterpret_cast<const float4*>(st);
    const float* sz = reinterpret_cast<const float*>(st + off_z);
    const float* sn = reinterpret_cast<const float*>(st + off_n);
    const float* sgw = reinterpret_cast<const float*>(st + off_g);

    // what one segment needs recomputed in both passes: alpha, e, delta, the lane's in-lane exclusive products
    auto segment = [&](int seg, float (&a)[4], float (&e)[4], float (&dlt)[4], float (&tl)[4], float (&zz)[4], float (&pre)[4]) -> float {
      const int base = seg * 128 + lane * 4;
      const float4 zq = reinterpret_cast<const float4*>(sz)[base >> 2];
      float4 nq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (NOISE) nq = reinterpret_cast<const float4*>(sn)[base >> 2];
      zz[0] = zq.x; zz[1] = zq.y; zz[2] = zq.z; zz[3] = zq.w;
      const float nz[4] = {nq.x, nq.y, nq.z, nq.w};
      float znext = __shfl_down_sync(kFullMask, zz[0], 1);
      if (lane == 31 && seg + 1 < M) znext = sz[base + 4];
      float prod = 1.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float zn = k < 3 ? zz[k < 3 ? k + 1 : k] : znext;
        const float dl = (seg == M - 1 && lane == 31 && k == 3) ? 1e10f : (zn - zz[k]);
        dlt[k] = dl * dnorm;
        pre[k] = sraw[seg * 128 + k * 32 + lane].w + nz[k];
        e[k] = __expf(-fmaxf(pre[k], 0.f) * dlt[k]);
        a[k] = 1.f - e[k];
        tl[k] = prod;
        prod *= (1.f - a[k]) + 1e-10f;
      }
      return prod;
    };

    // ---- pass 1, forward over the segments: exclusive transmittance of every lane's first sample, depth and acc totals
    float t0[M];                                         // transmittance in front of this lane's 4 samples of segment seg
    float carry = 1.f, sd = 0.f, sa = 0.f;
#pragma unroll
    for (int seg = 0; seg < M; ++seg) {
      float a[4], e[4], dlt[4], tl[4], zz[4], pre[4];
      const float prod = segment(seg, a, e, dlt, tl, zz, pre);
      const float incl = warp_scan_mul(prod, lane);
      float excl = __shfl_up_sync(kFullMask, incl, 1);
      if (lane == 0) excl = 1.f;
      t0[seg] = excl * carry;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float w = a[k] * (t0[seg] * tl[k]);
        sd = fmaf(w, zz[k], sd);
        sa += w;
      }
      carry *= __shfl_sync(kFullMask, incl, 31);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sd += __shfl_xor_sync(kFullMask, sd, o); sa += __shfl_xor_sync(kFullMask, sa, o); }
    if (gdisp != 0.f) {   // disp = 1 / max(1e-10, depth / acc): above the clamp disp = acc / depth
      const float q = sd / sa;
      if (!(q <= 1e-10f)) {   // also taken for NaN, which then propagates as in autograd
        gdep += -gdisp * sa / (sd * sd);
        gacc += gdisp / sd;
      }
    }
    if (white) gacc -= gr + gg + gb;
    // ---- pass 2, segments in reverse: G_k, suffix sums of G_k w_k, gradients -------------------------------------------
    float later_segs = 0.f;                              // sum of G w over all later segments
#pragma unroll
    for (int seg = M - 1; seg >= 0; --seg) {
      const int base = seg * 128 + lane * 4;
      float a[4], e[4], dlt[4], tl[4], zz[4], pre[4];
      segment(seg, a, e, dlt, tl, zz, pre);
      float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (GW) gq = reinterpret_cast<const float4*>(sgw)[base >> 2];
      const float gw[4] = {gq.x, gq.y, gq.z, gq.w};
      float G[4], suf[4], T[4], w[4];
      float run = 0.f;
#pragma unroll
      for (int k = 3; k >= 0; --k) {
        const float4 rw = sraw[seg * 128 + k * 32 + lane];
        T[k] = t0[seg] * tl[k];
        w[k] = a[k] * T[k];
        G[k] = gdep * zz[k] + gacc + gw[k];
        if (!detach_w) G[k] += gr * fast_sigmoid(rw.x) + gg * fast_sigmoid(rw.y) + gb * fast_sigmoid(rw.z);
        suf[k] = run;
        run = fmaf(G[k], w[k], run);
      }
      const float rincl = warp_rscan_add(run, lane);
      float later = __shfl_down_sync(kFullMask, rincl, 1);
      if (lane == 31) later = 0.f;
      later += later_segs;
      float4* out = reinterpret_cast<float4*>(g_raw) + ray * S + base;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 rw = sraw[seg * 128 + k * 32 + lane];
        const float f = (1.f - a[k]) + 1e-10f;
        const float dalpha = G[k] * T[k] - __fdividef(suf[k] + later, f);
        const float cr = fast_sigmoid(rw.x), cg = fast_sigmoid(rw.y), cb = fast_sigmoid(rw.z);
        float4 o;
        o.x = w[k] * gr * cr * (1.f - cr);
        o.y = w[k] * gg * cg * (1.f - cg);
        o.z = w[k] * gb * cb * (1.f - cb);
        o.w = (pre[k] > 0.f) ? dalpha * dlt[k] * e[k] : 0.f;
        st_stream4(out + k, o);
      }
      later_segs += __shfl_sync(kFullMask, rincl, 0);
    }
    __syncwarp();   // slot may be refilled by the next iteration's issue
  }
}

// S = 128 M for M in 2..8 (forward: M >= 3, shorter rays take composite_fwd_staged_kernel); returns 1 when launched
template <int M>
static int launch_fwd_seg(const float* raw, const float* z, const float* rays_d, int64_t ray_stride, const float* noise, int64_t R,
                          int white, float* rgb, float* disp, float* acc, float* depth, float* weights, float* alpha,
                          cudaStream_t stream) {
  constexpr int D = 2;
  const size_t smem = (size_t)kSegWarps * D * 128 * M * (noise ? 24 : 20);
  const int per_sm = (int)((227 * 1024) / (smem + 1024)) < 1 ? 1 : (int)((227 * 1024) / (smem + 1024));   // 1 KB per CTA is reserved
  const int64_t blocks = (R + kSegWarps - 1) / kSegWarps, cap = (int64_t)kNumSMs * (per_sm > 8 ? 8 : per_sm);
  const int grid = (int)(blocks < cap ? blocks : cap);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(composite_fwd_seg_kernel<M, D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kSegWarps * D * 128 * M * 24));
    cudaFuncSetAttribute(composite_fwd_seg_kernel<M, D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kSegWarps * D * 128 * M * 20));
    attr = true;
  }
  if (noise) composite_fwd_seg_kernel<M, D, true><<<grid, kSegThreads, smem, stream>>>(raw, z, noise, rays_d, ray_stride, R, white, rgb, disp, acc, depth, weights, alpha);
  else composite_fwd_seg_kernel<M, D, false><<<grid, kSegThreads, smem, stream>>>(raw, z, noise, rays_d, ray_stride, R, white, rgb, disp, acc, depth, weights, alpha);
  return check_launch("composite_fwd_seg_kernel");
}

template <int M, bool NOISE, bool GW>
static int launch_bwd_seg2(const float* raw, const float* z, const float* rays_d, int64_t ray_stride, const float* noise, int64_t R,
                           int white, int detach_w, const float* g_rgb, const float* g_disp, const float* g_acc,
                           const float* g_depth, const float* g_w, float* g_raw, cudaStream_t stream) {
  constexpr int D = 2;
  constexpr size_t smem = (size_t)kSegWarps * D * 128 * M * (20 + (NOISE ? 4 : 0) + (GW ? 4 : 0));
  constexpr int per_sm = (int)((227 * 1024) / (smem + 1024)) < 1 ? 1 : (int)((227 * 1024) / (smem + 1024));   // 1 KB per CTA is reserved
  const int64_t blocks = (R + kSegWarps - 1) / kSegWarps, cap = (int64_t)kNumSMs * (per_sm > 8 ? 8 : per_sm);
  const int grid = (int)(blocks < cap ? blocks : cap);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(composite_bwd_seg_kernel<M, D, NOISE, GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  composite_bwd_seg_kernel<M, D, NOISE, GW><<<grid, kSegThreads, smem, stream>>>(raw, z, noise, rays_d, ray_stride, R, white, detach_w,
                                                                               g_rgb, g_disp, g_acc, g_depth, g_w, g_raw);
  return check_launch("composite_bwd_seg_kernel");
}
template <int M>
static int launch_bwd_seg(const float* raw, const float* z, const float* rays_d, int64_t ray_stride, const float* noise, int64_t R,
                          int white, int detach_w, const float* g_rgb, const float* g_disp, const float* g_acc,
                          const float* g_depth, const float* g_w, float* g_raw, cudaStream_t stream) {
#define GBN_SEG_ARGS raw, z, rays_d, ray_stride, noise, R, white, detach_w, g_rgb, g_disp, g_acc, g_depth, g_w, g_raw, stream
  if (noise) return g_w ? launch_bwd_seg2<M, true, true>(GBN_SEG_ARGS) : launch_bwd_seg2<M, true, false>(GBN_SEG_ARGS);
  return g_w ? launch_bwd_seg2<M, false, true>(GBN_SEG_ARGS) : launch_bwd_seg2<M, false, false>(GBN_SEG_ARGS);
#undef GBN_SEG_ARGS
}

// returns 1 when it launched (shape S = 64 or 128, 16-byte aligned operands), 0 when the caller must use the generic kernel
int launch_composite_bwd_staged(const float* raw, const float* z, const float* rays_d, int64_t ray_stride, const float* noise,
                                int64_t R, int S, int white, int detach_w, const float* g_rgb, const float* g_disp,
                                const float* g_acc, const float* g_depth, const float* g_w, float* g_raw, cudaStream_t stream,
                                int* rc) {
  *rc = GBN_OK;
  static const bool off = [] { const char* e = getenv("GBNERF_COMP_BWD_GENERIC"); return e && e[0] == '1'; }();
  const uintptr_t al = reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(noise) |
                       reinterpret_cast<uintptr_t>(g_w) | reinterpret_cast<uintptr_t>(g_raw);
  if (off || R < 1 || (al & 15) != 0) return 0;
  if (S >= 256 && S <= 1024 && S % 128 == 0) {   // long rays: segment walk (composite_bwd_seg_kernel)
#define GBN_SEG_CALL(MM) *rc = launch_bwd_seg<MM>(raw, z, rays_d, ray_stride, noise, R, white, detach_w, g_rgb, g_disp, g_acc, g_depth, g_w, g_raw, stream)
    switch (S / 128) {
      case 2: GBN_SEG_CALL(2); break; case 3: GBN_SEG_CALL(3); break; case 4: GBN_SEG_CALL(4); break; case 5: GBN_SEG_CALL(5); break;
      case 6: GBN_SEG_CALL(6); break; case 7: GBN_SEG_CALL(7); break; default: GBN_SEG_CALL(8); break;
    }
#undef GBN_SEG_CALL
    return 1;
  }
  if (S != 64 && S != 128) return 0;
  // persistent grid = what is resident: 4 CTAs per SM at K = 2, 3 at K = 4 (80 registers; __launch_bounds__ above) - a
  // fourth CTA per SM would run as a second, mostly empty wave
  const int64_t blocks = (R + kStWarps - 1) / kStWarps, cap = (int64_t)kNumSMs * (S == 64 ? 4 : 3);
  const int grid = (int)(blocks < cap ? blocks : cap);
#define GBN_BW_LAUNCH(KK, NN, GG)                                                                                       \
  do {                                                                                                                  \
    constexpr int DD = 2;                                                                                               \
    constexpr size_t smem = (size_t)kStWarps * DD * 32 * KK * (20 + (NN ? 4 : 0) + (GG ? 4 : 0));                       \
    static bool attr = false;                                                                                           \
    if (!attr) {                                                                                                        \
      cudaFuncSetAttribute(composite_bwd_staged_kernel<KK, DD, NN, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      attr = true;                                                                                                      \
    }                                                                                                                   \
    composite_bwd_staged_kernel<KK, DD, NN, GG><<<grid, kStThreads, smem, stream>>>(                                    \
        raw, z, noise, rays_d, ray_stride, R, white, detach_w, g_rgb, g_disp, g_acc, g_depth, g_w, g_raw);              \
  } while (0)
#define GBN_BW_K(KK)                                                                                                    \
  do {                                                                                                                  \
    if (noise) { if (g_w) GBN_BW_LAUNCH(KK, true, true); else GBN_BW_LAUNCH(KK, true, false); }                         \
    else { if (g_w) GBN_BW_LAUNCH(KK, false, true); else GBN_BW_LAUNCH(KK, false, false); }                             \
  } while (0)
  if (S == 64) GBN_BW_K(2); else GBN_BW_K(4);
#undef GBN_BW_K
#undef GBN_BW_LAUNCH
  *rc = check_launch("composite_bwd_staged_kernel");
  return 1;
}

// Launches the staged kernel when the shapes allow (S <= 256, 16-byte aligned raw); returns how many rays it took
// (0 = the caller uses the generic kernel).
int64_t launch_composite_fwd_staged(const float* raw, const float* z, const float* rays_d, int64_t ray_stride,
                                    const float* noise, int64_t R, int S, int white, float* rgb, float* disp, float* acc,
                                    float* depth, float* weights, float* alpha, cudaStream_t stream, int* rc) {
  *rc = GBN_OK;
  const int K = (S + 31) / 32;
  const uintptr_t al = reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(weights) | reinterpret_cast<uintptr_t>(alpha);
  const bool full = S == 32 * K && ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(noise)) & 15) == 0;
  if (R < 1 || S < 2 || (al & 15) != 0) return 0;
  if (S >= 384 && S <= 1024 && S % 128 == 0 && full) {   // long rays: segment walk (composite_fwd_seg_kernel)
#define GBN_SEG_CALL(MM) *rc = launch_fwd_seg<MM>(raw, z, rays_d, ray_stride, noise, R, white, rgb, disp, acc, depth, weights, alpha, stream)
    switch (S / 128) {
      case 3: GBN_SEG_CALL(3); break; case 4: GBN_SEG_CALL(4); break; case 5: GBN_SEG_CALL(5); break;
      case 6: GBN_SEG_CALL(6); break; case 7: GBN_SEG_CALL(7); break; default: GBN_SEG_CALL(8); break;
    }
#undef GBN_SEG_CALL
    return R;
  }
  if (K > 8 || K == 5 || K == 7) return 0;
  static const bool no_pairs = [] { const char* e = getenv("GBNERF_COMP_NOPAIR"); return e && e[0] == '1'; }();
  if (S == 64 && full && !no_pairs) {   // two rays per warp (composite_fwd_pair64_kernel)
    constexpr int D = 2;
    const size_t smem = (size_t)kStWarps * D * 128 * (noise ? 24 : 20);
    static const int ctas = [] { const char* e = getenv("GBNERF_COMP_CTAS"); const int v = e ? atoi(e) : 4; return v >= 1 && v <= 4 ? v : 4; }();
    const int64_t blocks = ((R + 1) / 2 + kStWarps - 1) / kStWarps, cap = (int64_t)kNumSMs * ctas;
    const int grid = (int)(blocks < cap ? blocks : cap);
    static bool attr = false;   // immutable kernel attribute, set once
    if (!attr) {
      cudaFuncSetAttribute(composite_fwd_pair64_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStagedSmemBudget);
      cudaFuncSetAttribute(composite_fwd_pair64_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStagedSmemBudget);
      attr = true;
    }
    if (noise) composite_fwd_pair64_kernel<D, true><<<grid, kStThreads, smem, stream>>>(raw, z, noise, rays_d, ray_stride, R, white, rgb, disp, acc, depth, weights, alpha);
    else composite_fwd_pair64_kernel<D, false><<<grid, kStThreads, smem, stream>>>(raw, z, noise, rays_d, ray_stride, R, white, rgb, disp, acc, depth, weights, alpha);
    *rc = check_launch("composite_fwd_pair64_kernel");
    return R;
  }
  const size_t ray_bytes = (size_t)S * (noise ? 24 : 20);
  const int dmax = (int)(kStagedSmemBudget / (kStWarps * ray_bytes));
  const int D = dmax >= 8 ? 8 : (dmax >= 4 ? 4 : 2);
  if (dmax < 2) return 0;
  const size_t smem = (size_t)kStWarps * D * ray_bytes;
  static const int ctas_per_sm = [] { const char* e = getenv("GBNERF_COMP_CTAS"); const int v = e ? atoi(e) : 4; return v >= 1 && v <= 4 ? v : 4; }();
  const int64_t blocks = (R + kStWarps - 1) / kStWarps, cap = (int64_t)kNumSMs * ctas_per_sm;
  const int grid = (int)(blocks < cap ? blocks : cap);
#define GBN_ST_LAUNCH2(KK, DD, NN, FF)                                                                              \
  do {                                                                                                              \
    static bool attr = false;                                                                                       \
    if (!attr) {                                                                                                    \
      cudaFuncSetAttribute(composite_fwd_staged_kernel<KK, DD, NN, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           kStagedSmemBudget);                                                                      \
      attr = true;                                                                                                  \
    }                                                                                                               \
    composite_fwd_staged_kernel<KK, DD, NN, FF><<<grid, kStThreads, smem, stream>>>(                                \
        raw, z, noise, rays_d, ray_stride, R, S, white, rgb, disp, acc, depth, weights, alpha);                     \
  } while (0)
#define GBN_ST_LAUNCH(KK, DD, NN)                                                                                   \
  do {                                                                                                              \
    if (full) GBN_ST_LAUNCH2(KK, DD, NN, true); else GBN_ST_LAUNCH2(KK, DD, NN, false);                             \
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
