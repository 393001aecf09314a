// HBM-bound stages of the DS_NeRF render path: stratified depths, standalone positional encoding,
// alpha compositing (forward + backward), inverse-CDF sampling + sort-merge, loss seed.
//
// All of them are one-warp-per-ray (or one-thread-per-element) streaming kernels: every byte is read once
// with coalesced (16-byte where the layout allows) accesses, scans run on warp shuffles, nothing is staged
// through HBM.  Grids are persistent: kNumSMs x resident CTAs, looping over rays.
#include <stdlib.h>

#include "common.cuh"

namespace gbn {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

static inline int persistent_grid(int64_t warps_needed, int ctas_per_sm) {
  int64_t blocks = (warps_needed + kWarpsPerBlock - 1) / kWarpsPerBlock;
  int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

// =========================================================================================================
// stratified depths — run.py:2291-2315
// =========================================================================================================
__device__ __forceinline__ float z_at(float near, float far, int s, int S, int lindisp) {
  const float t = linspace01(s, S);
  const float omt = __fsub_rn(1.f, t);
  if (lindisp) {
    // 1 / (1/near * (1-t) + 1/far * t), evaluated with the reference's rounding steps
    const float a = __fmul_rn(__fdiv_rn(1.f, near), omt);
    const float b = __fmul_rn(__fdiv_rn(1.f, far), t);
    return __fdiv_rn(1.f, __fadd_rn(a, b));
  }
  return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
}

__global__ void __launch_bounds__(256) zvals_kernel(const float* __restrict__ near, const float* __restrict__ far,
                                                    int64_t stride, int64_t R, int S, int lindisp,
                                                    const float* __restrict__ t_rand, float* __restrict__ z) {
  const int64_t total = R * S;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / S;
    const int s = (int)(i - r * S);
    const float n = __ldg(near + r * stride), f = __ldg(far + r * stride);
    float zc = z_at(n, f, s, S, lindisp);
    if (t_rand != nullptr) {
      const float zp = s > 0 ? z_at(n, f, s - 1, S, lindisp) : zc;
      const float zn = s < S - 1 ? z_at(n, f, s + 1, S, lindisp) : zc;
      const float lo = s > 0 ? __fmul_rn(.5f, __fadd_rn(zc, zp)) : zc;
      const float hi = s < S - 1 ? __fmul_rn(.5f, __fadd_rn(zn, zc)) : zc;
      zc = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), ld_stream(t_rand + i)));
    }
    z[i] = zc;
  }
}

// Same arithmetic for S in {32, 64, 128, 256} (the render path): a thread owns FOUR consecutive samples of a ray, so
// linspace(0, 1, S) is evaluated once per thread (six values: its four samples and their two neighbours), the int64
// division by S is gone, the jitter is read and the depths are written as 16-byte streaming accesses, and two rays per
// thread are in flight at a time.  (The generic kernel ran at 12 % of the HBM copy peak; a first fixed-S version with one
// 4-byte element per thread and the neighbours exchanged through shared memory at 20 %: ~1 MB in flight per wave.)
template <int S>
__global__ void __launch_bounds__(256) zvals_fixed_kernel(const float* __restrict__ near, const float* __restrict__ far,
                                                          int64_t stride, int64_t R, int lindisp,
                                                          const float* __restrict__ t_rand, float* __restrict__ z) {
  constexpr int TPR = S / 4, RPB = 256 / TPR;          // threads per ray, rays per block iteration
  const int q = threadIdx.x % TPR, rl = threadIdx.x / TPR, s0 = 4 * q;
  float t[6], omt[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int s = min(max(s0 - 1 + k, 0), S - 1);
    t[k] = linspace01(s, S);
    omt[k] = __fsub_rn(1.f, t[k]);
  }
  auto depths = [&](float n, float f, float (&zc)[6]) {
    if (lindisp) {
      const float in = __fdiv_rn(1.f, n), inf = __fdiv_rn(1.f, f);
#pragma unroll
      for (int k = 0; k < 6; ++k) zc[k] = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(in, omt[k]), __fmul_rn(inf, t[k])));
    } else {
#pragma unroll
      for (int k = 0; k < 6; ++k) zc[k] = __fadd_rn(__fmul_rn(n, omt[k]), __fmul_rn(f, t[k]));
    }
  };
  auto jitter = [&](const float (&zc)[6], float4 u) {
    float o[4];
    const float uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int s = s0 + k;
      const float c = zc[k + 1];
      const float lo = s > 0 ? __fmul_rn(.5f, __fadd_rn(c, zc[k])) : c;
      const float hi = s < S - 1 ? __fmul_rn(.5f, __fadd_rn(zc[k + 2], c)) : c;
      o[k] = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), uu[k]));
    }
    return make_float4(o[0], o[1], o[2], o[3]);
  };
  const int64_t step = (int64_t)gridDim.x * RPB;
  for (int64_t r0 = (int64_t)blockIdx.x * RPB + rl; r0 < R; r0 += 2 * step) {
    const int64_t r1 = r0 + step;
    const bool two = r1 < R;
    float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0;
    if (t_rand != nullptr) {
      u0 = ld_stream4(reinterpret_cast<const float4*>(t_rand + r0 * S) + q);
      if (two) u1 = ld_stream4(reinterpret_cast<const float4*>(t_rand + r1 * S) + q);
    }
    float zc[6];
    depths(__ldg(near + r0 * stride), __ldg(far + r0 * stride), zc);
    st_stream4(reinterpret_cast<float4*>(z + r0 * S) + q,
               t_rand != nullptr ? jitter(zc, u0) : make_float4(zc[1], zc[2], zc[3], zc[4]));
    if (two) {
      depths(__ldg(near + r1 * stride), __ldg(far + r1 * stride), zc);
      st_stream4(reinterpret_cast<float4*>(z + r1 * S) + q,
                 t_rand != nullptr ? jitter(zc, u1) : make_float4(zc[1], zc[2], zc[3], zc[4]));
    }
  }
}

// =========================================================================================================
// standalone positional encoding — [R*S, 90] fp32 (measurement / test entry; the MLP kernel fuses this)
// =========================================================================================================
constexpr int kEncPts = 64;  // points per tile; tile = 64 x 90 floats staged in smem for coalesced stores

__global__ void __launch_bounds__(256) encode_kernel(const float* __restrict__ ro, const float* __restrict__ rd,
                                                     const float* __restrict__ vd, int64_t stride,
                                                     const float* __restrict__ z, int64_t R, int S,
                                                     float* __restrict__ out) {
  __shared__ __align__(16) float tile[kEncPts * GBN_EMB_CH];
  const int64_t P = R * S;
  const int64_t ntiles = (P + kEncPts - 1) / kEncPts;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t p0 = t * kEncPts;
    // 64 points x 4 jobs: job 0..2 = point axis (21 values), job 3 = the three direction axes (27 values)
    {
      const int pl = threadIdx.x & (kEncPts - 1), job = threadIdx.x / kEncPts;
      const int64_t p = p0 + pl;
      if (p < P) {
        const int64_t r = p / S;
        float* row = tile + pl * GBN_EMB_CH;
        if (job < 3) {
          const float x = __fadd_rn(__ldg(ro + r * stride + job), __fmul_rn(__ldg(rd + r * stride + job), __ldg(z + p)));
          float e[20];
          posenc_axis<10>(x, e);
          row[job] = x;
#pragma unroll
          for (int k = 0; k < 10; ++k) {
            row[3 + 6 * k + job] = e[2 * k];
            row[6 + 6 * k + job] = e[2 * k + 1];
          }
        } else {
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const float x = __ldg(vd + r * stride + a);
            float e[8];
            posenc_axis<4>(x, e);
            row[63 + a] = x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              row[66 + 6 * k + a] = e[2 * k];
              row[69 + 6 * k + a] = e[2 * k + 1];
            }
          }
        }
      }
    }
    __syncthreads();
    const int64_t remain = P - p0;
    const int npts = remain < kEncPts ? (int)remain : kEncPts;
    const int nvec = npts * GBN_EMB_CH / 4;  // 90*npts floats; tile base is 16B aligned (64*360 B per tile)
    float4* dst = reinterpret_cast<float4*>(out + p0 * GBN_EMB_CH);
    const float4* src = reinterpret_cast<const float4*>(tile);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) st_stream4(dst + i, src[i]);
    for (int i = nvec * 4 + threadIdx.x; i < npts * GBN_EMB_CH; i += blockDim.x) out[p0 * GBN_EMB_CH + i] = tile[i];
    __syncthreads();
  }
}

// =========================================================================================================
// alpha compositing — raw2outputs, run_nerf_helpers.py:352-406.  One warp per ray, lane-strided samples.
// =========================================================================================================
struct RaySample {
  float4 raw;
  float z, nz;
};

template <int NCH>
__device__ __forceinline__ void load_ray(RaySample (&sm)[NCH], const float* __restrict__ raw,
                                         const float* __restrict__ z, const float* __restrict__ noise, int64_t ray,
                                         int S, int lane) {
  const float4* rawp = reinterpret_cast<const float4*>(raw) + ray * S;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int s = c * 32 + lane;
    if (s < S) {
      sm[c].raw = ld_stream4(rawp + s);
      sm[c].z = ld_stream(z + ray * S + s);
      sm[c].nz = noise ? ld_stream(noise + ray * S + s) : 0.f;
    } else {
      sm[c].raw = make_float4(0.f, 0.f, 0.f, 0.f);
      sm[c].z = 0.f;
      sm[c].nz = 0.f;
    }
  }
}

__device__ __forceinline__ float ray_norm(const float* __restrict__ d, int64_t stride, int64_t ray) {
  const float x = __ldg(d + ray * stride), y = __ldg(d + ray * stride + 1), w = __ldg(d + ray * stride + 2);
  return sqrtf(x * x + y * y + w * w);
}

// per-chunk forward quantities: alpha, transmittance T (exclusive product), e = exp(-sigma*delta), delta
template <int NCH>
__device__ __forceinline__ void march(const RaySample (&sm)[NCH], float dnorm, int S, int lane, float (&alpha)[NCH],
                                      float (&T)[NCH], float (&e)[NCH], float (&delta)[NCH]) {
  float carry = 1.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int s = c * 32 + lane;
    float zn = __shfl_down_sync(kFullMask, sm[c].z, 1);
    const float z0next = (c + 1 < NCH) ? __shfl_sync(kFullMask, sm[c + 1 < NCH ? c + 1 : c].z, 0) : 0.f;
    if (lane == 31) zn = z0next;
    const float dl = (s == S - 1) ? 1e10f : __fsub_rn(zn, sm[c].z);
    delta[c] = __fmul_rn(dl, dnorm);
    const float sigma = fmaxf(__fadd_rn(sm[c].raw.w, sm[c].nz), 0.f);
    e[c] = (s < S) ? expf(-__fmul_rn(sigma, delta[c])) : 1.f;
    alpha[c] = __fsub_rn(1.f, e[c]);
    const float f = __fadd_rn(__fsub_rn(1.f, alpha[c]), 1e-10f);
    const float incl = warp_scan_mul(f, lane);
    float excl = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) excl = 1.f;
    T[c] = carry * excl;
    carry *= __shfl_sync(kFullMask, incl, 31);
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }

template <int NCH>
__global__ void __launch_bounds__(kThreads) composite_fwd_kernel(
    const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ d, int64_t stride,
    const float* __restrict__ noise, int64_t R, int S, int white, float* __restrict__ rgb, float* __restrict__ disp,
    float* __restrict__ acc, float* __restrict__ depth, float* __restrict__ weights, float* __restrict__ alpha_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t ray = warp0; ray < R; ray += nwarps) {
    RaySample sm[NCH];
    load_ray<NCH>(sm, raw, z, noise, ray, S, lane);
    const float dnorm = ray_norm(d, stride, ray);
    float a[NCH], T[NCH], e[NCH], dl[NCH];
    march<NCH>(sm, dnorm, S, lane, a, T, e, dl);
    float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int s = c * 32 + lane;
      if (s < S) {
        const float w = a[c] * T[c];
        sr += w * sigmoidf_(sm[c].raw.x);
        sg += w * sigmoidf_(sm[c].raw.y);
        sb += w * sigmoidf_(sm[c].raw.z);
        sd += w * sm[c].z;
        sa += w;
        st_stream(weights + ray * S + s, w);
        if (alpha_out) st_stream(alpha_out + ray * S + s, a[c]);
      }
    }
    sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd); sa = warp_sum(sa);
    if (lane == 0) {
      const float bg = white ? (1.f - sa) : 0.f;
      rgb[ray * 3 + 0] = sr + bg;
      rgb[ray * 3 + 1] = sg + bg;
      rgb[ray * 3 + 2] = sb + bg;
      // 1 / max(1e-10, depth/acc): NaN when acc == 0, as torch.max propagates the NaN of 0/0
      const float q = __fdiv_rn(sd, sa);
      disp[ray] = __fdiv_rn(1.f, (q != q) ? q : fmaxf(1e-10f, q));
      acc[ray] = sa;
      depth[ray] = sd;
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(kThreads) composite_bwd_kernel(
    const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ d, int64_t stride,
    const float* __restrict__ noise, int64_t R, int S, int white, int detach_w, const float* __restrict__ g_rgb,
    const float* __restrict__ g_disp, const float* __restrict__ g_acc, const float* __restrict__ g_depth,
    const float* __restrict__ g_w, float* __restrict__ g_raw) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t ray = warp0; ray < R; ray += nwarps) {
    RaySample sm[NCH];
    load_ray<NCH>(sm, raw, z, noise, ray, S, lane);
    float gw[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int s = c * 32 + lane;
      gw[c] = (g_w != nullptr && s < S) ? ld_stream(g_w + ray * S + s) : 0.f;
    }
    const float dnorm = ray_norm(d, stride, ray);
    const float gr = g_rgb ? __ldg(g_rgb + ray * 3) : 0.f, gg = g_rgb ? __ldg(g_rgb + ray * 3 + 1) : 0.f,
                gb = g_rgb ? __ldg(g_rgb + ray * 3 + 2) : 0.f;
    const float gdisp = g_disp ? __ldg(g_disp + ray) : 0.f;
    float gacc = g_acc ? __ldg(g_acc + ray) : 0.f;
    float gdep = g_depth ? __ldg(g_depth + ray) : 0.f;

    float a[NCH], T[NCH], e[NCH], dl[NCH];
    march<NCH>(sm, dnorm, S, lane, a, T, e, dl);
    float sd = 0.f, sa = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const float w = a[c] * T[c];  // a == 0 past S
      sd += w * sm[c].z;
      sa += w;
    }
    sd = warp_sum(sd);
    sa = warp_sum(sa);
    // disp = 1/max(1e-10, depth/acc): above the clamp disp = acc/depth
    if (gdisp != 0.f) {
      const float q = sd / sa;
      if (!(q <= 1e-10f)) {  // also taken for NaN, which then propagates as in autograd
        gdep += -gdisp * sa / (sd * sd);
        gacc += gdisp / sd;
      }
    }
    if (white) gacc -= gr + gg + gb;

    float carry = 0.f;  // sum of G_k w_k over later chunks
#pragma unroll
    for (int c = NCH - 1; c >= 0; --c) {
      const int s = c * 32 + lane;
      const float cr = sigmoidf_(sm[c].raw.x), cg = sigmoidf_(sm[c].raw.y), cb = sigmoidf_(sm[c].raw.z);
      const float w = a[c] * T[c];
      float G = gdep * sm[c].z + gacc + gw[c];
      if (!detach_w) G += gr * cr + gg * cg + gb * cb;
      const float Gw = (s < S) ? G * w : 0.f;
      const float incl = warp_rscan_add(Gw, lane);
      float excl = __shfl_down_sync(kFullMask, incl, 1);
      if (lane == 31) excl = 0.f;
      const float suffix = carry + excl;
      carry += __shfl_sync(kFullMask, incl, 0);
      if (s < S) {
        const float f = __fadd_rn(__fsub_rn(1.f, a[c]), 1e-10f);
        const float dalpha = G * T[c] - suffix / f;
        const float pre = __fadd_rn(sm[c].raw.w, sm[c].nz);
        const float dsig = (pre > 0.f) ? dalpha * dl[c] * e[c] : 0.f;
        float4 o;
        o.x = w * gr * cr * (1.f - cr);
        o.y = w * gg * cg * (1.f - cg);
        o.z = w * gb * cb * (1.f - cb);
        o.w = dsig;
        st_stream4(reinterpret_cast<float4*>(g_raw) + ray * S + s, o);
      }
    }
  }
}

// =========================================================================================================
// inverse-CDF sampling — sample_pdf, run_nerf_helpers.py:306-349; merge — run.py:2348
// =========================================================================================================

// #{j < n : a[j] <= v}  (torch.searchsorted(..., right=True)); a ascending in shared memory
__device__ __forceinline__ int upper_bound_s(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// #{j < n : a[j] < v}
__device__ __forceinline__ int lower_bound_s(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// warp builds cdf[0..M] from M weights: cdf[0]=0, cdf[j+1] = cumsum((w+1e-5)/sum)[j]
__device__ __forceinline__ void warp_build_cdf(const float* __restrict__ w, int M, float* cdf, int lane) {
  float tot = 0.f;
  for (int j = lane; j < M; j += 32) {
    const float v = __fadd_rn(ld_stream(w + j), 1e-5f);
    cdf[j + 1] = v;  // parked; normalised in place below by the same lane
    tot += v;
  }
  tot = warp_sum(tot);
  float carry = 0.f;
  if (lane == 0) cdf[0] = 0.f;
  for (int j0 = 0; j0 < M; j0 += 32) {
    const int j = j0 + lane;
    const float p = (j < M) ? __fdiv_rn(cdf[j + 1], tot) : 0.f;
    const float incl = warp_scan_add(p, lane);
    if (j < M) cdf[j + 1] = carry + incl;
    carry += __shfl_sync(kFullMask, incl, 31);
  }
}

__device__ __forceinline__ float invert_one(const float* cdf, const float* bins, int B, float u, int* lo_out) {
  const int ind = upper_bound_s(cdf, B, u);
  const int lo = ind - 1 < 0 ? 0 : ind - 1;
  const int hi = ind > B - 1 ? B - 1 : ind;
  const float c0 = cdf[lo], c1 = cdf[hi];
  float den = __fsub_rn(c1, c0);
  if (den < 1e-5f) den = 1.f;
  const float t = __fdiv_rn(__fsub_rn(u, c0), den);
  const float b0 = bins[lo], b1 = bins[hi];
  *lo_out = lo;
  return __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
}

__global__ void __launch_bounds__(kThreads) searchsorted_kernel(const float* __restrict__ cdf,
                                                                const float* __restrict__ u, int64_t R, int B, int N,
                                                                int64_t* __restrict__ inds) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* c = smem + wib * B;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + wib; ray < R; ray += nwarps) {
    for (int j = lane; j < B; j += 32) c[j] = ld_stream(cdf + ray * B + j);
    __syncwarp();
    for (int n = lane; n < N; n += 32) inds[ray * N + n] = upper_bound_s(c, B, ld_stream(u + ray * N + n));
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kThreads) sample_pdf_kernel(const float* __restrict__ bins,
                                                              const float* __restrict__ weights,
                                                              const float* __restrict__ u, int64_t R, int B, int N,
                                                              float* __restrict__ samples) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* cdf = smem + wib * 2 * B;
  float* bn = cdf + B;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + wib; ray < R; ray += nwarps) {
    for (int j = lane; j < B; j += 32) bn[j] = ld_stream(bins + ray * B + j);
    warp_build_cdf(weights + ray * (B - 1), B - 1, cdf, lane);
    __syncwarp();
    for (int n = lane; n < N; n += 32) {
      const float uu = u ? ld_stream(u + ray * N + n) : linspace01(n, N);
      int lo;
      st_stream(samples + ray * N + n, invert_one(cdf, bn, B, uu, &lo));
    }
    __syncwarp();
  }
}

// in-place ascending bitonic sort of a[0..n) (n a power of two) by one warp
__device__ __forceinline__ void warp_bitonic_sort(float* a, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n; i += 32) {
        const int p = i ^ j;
        if (p > i) {
          const float x = a[i], y = a[p];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[p] = x; }
        }
      }
      __syncwarp();
    }
  }
}

// smem per warp: zv[S] | cdf[S] (B=S-1 used) | bins[S] | smp[Npad] | merged[S+N] | hist[S+1] (ints)
//
// Merge.  Both lists ascending.  A sample drawn from bin lo lies in [z_mid[lo], z_mid[lo+1]], i.e. strictly
// between z[lo] and z[lo+2], so its count of coarse depths below-or-equal is lo+1+(z[lo+1] <= s): one
// compare, no search.  The coarse depths' ranks follow from a histogram of those counts and one prefix sum:
// #{s < z_i} = #{j : c_j <= i}.  Random u (training) gives unsorted samples: they are bitonic-sorted first
// and ranked by binary search instead.
__global__ void __launch_bounds__(kThreads) sample_merge_kernel(const float* __restrict__ z_vals,
                                                                const float* __restrict__ weights,
                                                                const float* __restrict__ u, const float* __restrict__ cdf_in,
                                                                int64_t R, int S, int N,
                                                                int Npad, float* __restrict__ z_samples,
                                                                float* __restrict__ z_merged,
                                                                float* __restrict__ z_std, int* __restrict__ inds_out) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int per_warp = 5 * S + Npad + N + 1;
  float* zv = smem + (size_t)wib * per_warp;
  float* cdf = zv + S;
  float* bn = cdf + S;
  float* smp = bn + S;
  float* mrg = smp + Npad;
  int* hist = reinterpret_cast<int*>(mrg + S + N);
  const int B = S - 1;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + wib; ray < R; ray += nwarps) {
    for (int j = lane; j < S; j += 32) zv[j] = ld_stream(z_vals + ray * S + j);
    if (u == nullptr) for (int j = lane; j <= S; j += 32) hist[j] = 0;
    __syncwarp();
    for (int j = lane; j < B; j += 32) bn[j] = __fmul_rn(.5f, __fadd_rn(zv[j + 1], zv[j]));
    if (cdf_in != nullptr) { for (int j = lane; j < B; j += 32) cdf[j] = __ldg(cdf_in + ray * B + j); }
    else warp_build_cdf(weights + ray * S + 1, S - 2, cdf, lane);
    __syncwarp();
    float sum = 0.f;
    for (int n = lane; n < Npad; n += 32) {
      float v = __int_as_float(0x7f800000);
      if (n < N) {
        const float uu = u ? ld_stream(u + ray * N + n) : linspace01(n, N);
        int lo;
        v = invert_one(cdf, bn, B, uu, &lo);
        sum += v;
        if (z_samples) st_stream(z_samples + ray * N + n, v);
        if (inds_out) inds_out[ray * N + n] = upper_bound_s(cdf, B, uu);
        if (u == nullptr) {
          const int c = lo + 1 + (zv[lo + 1] <= v ? 1 : 0);
          mrg[n + c] = v;
          atomicAdd(&hist[c], 1);
        }
      }
      smp[n] = v;
    }
    __syncwarp();
    if (z_std) {
      const float mean = warp_sum(sum) / (float)N;
      float var = 0.f;
      for (int n = lane; n < N; n += 32) { const float dv = smp[n] - mean; var += dv * dv; }
      var = warp_sum(var);
      if (lane == 0) z_std[ray] = sqrtf(var / (float)N);
    }
    if (u == nullptr) {
      int carry = 0;
      for (int i0 = 0; i0 < S; i0 += 32) {
        const int i = i0 + lane;
        int v = (i < S) ? hist[i] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(kFullMask, v, o);
          if (lane >= o) v += t;
        }
        if (i < S) mrg[i + carry + v] = zv[i];
        carry += __shfl_sync(kFullMask, v, 31);
      }
    } else {
      warp_bitonic_sort(smp, Npad, lane);
      for (int i = lane; i < S; i += 32) mrg[i + lower_bound_s(smp, N, zv[i])] = zv[i];
      for (int j = lane; j < N; j += 32) mrg[j + upper_bound_s(zv, S, smp[j])] = smp[j];
    }
    __syncwarp();
    for (int i = lane; i < S + N; i += 32) st_stream(z_merged + ray * (S + N) + i, mrg[i]);
    __syncwarp();
  }
}

template <int STEPS, bool UPPER, int STRIDE = 1>   // #{j < n : a[j] <= v} (UPPER) or #{j < n : a[j] < v}, n <= 2^STEPS
__device__ __forceinline__ int count_below(const float* a, int n, float v) {
  int lo = 0;
#pragma unroll
  for (int s = STEPS - 1; s >= 0; --s) {
    const int probe = lo + (1 << s);
    const float x = a[STRIDE * (min(probe, n) - 1)];
    const bool take = probe <= n && (UPPER ? x <= v : x < v);
    lo = take ? probe : lo;
  }
  return lo;
}


// ---- sample + merge with CONSECUTIVE ownership (the render path's shapes) ---------------------------------------------
// S = 4 LANES coarse depths and N = KN LANES new samples per ray, LANES = 8, 16 (four / two rays per warp) or 32.  A lane owns
// four consecutive depths / weights / cdf entries (one 16-byte load each) and KN consecutive samples, so
//   * the cdf is an in-lane prefix + ONE LANES-wide shuffle scan of (w + 1e-5), scaled by one reciprocal of the total
//     (the generic kernels divide every bin by the total and scan 32-strided chunks);
//   * bins, the rank histogram of the deterministic merge and the merged row are vector accesses;
//   * with 64 coarse samples the two half-warps share every shuffle / scan instruction (cf. composite_fwd_pair64_kernel);
//   * the interpolation uses one fast reciprocal per sample (the reference's IEEE division costs ~10 instructions; the
//     value tolerance of the path is 2e-5 relative, the bin INDEX is exact either way: it comes from the search alone).
// Searches are the same fixed-depth upper-bound walks as in sample_merge_fast_kernel; `inds_out` returns their result
// (= torch.searchsorted(cdf, u, right=True), helpers:333) and `cdf_in` replaces the cdf built from the weights, which
// together let the test assert bit-exact indices on this kernel given the same cdf and uniforms.
// ~290 warp instructions and ~75 shared-memory wavefronts per 64 + 64 ray (465 instructions for the round-1 kernel); the
// kernel is bound by the shared-memory pipe and the issue slots together (profiles/r2_sample_cons_ncu.md).
template <int LANES, int KN>
__device__ __forceinline__ void seg_bitonic_sort(float (&v)[KN], int l) {
  constexpr int N = LANES * KN;
#pragma unroll
  for (int kk = 2; kk <= N; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      if (j < KN) {                          // partner element lives in the same lane
#pragma unroll
        for (int k = 0; k < KN; ++k) {
          const int k2 = k ^ j;
          if (k2 > k) {
            const bool up = ((l * KN + k) & kk) == 0;
            const float x = v[k], y = v[k2];
            const bool sw = (x > y) == up;
            v[k] = sw ? y : x;
            v[k2] = sw ? x : y;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < KN; ++k) {
          const float y = __shfl_xor_sync(kFullMask, v[k], j / KN);
          const bool lower = (l & (j / KN)) == 0;
          const bool up = ((l * KN + k) & kk) == 0;
          v[k] = (lower == up) ? fminf(v[k], y) : fmaxf(v[k], y);
        }
      }
    }
  }
}

template <int LANES, int KN, bool DET>
__global__ void __launch_bounds__(kThreads, 4) sample_merge_cons_kernel(const float* __restrict__ z_vals,
                                                                     const float* __restrict__ weights,
                                                                     const float* __restrict__ u,
                                                                     const float* __restrict__ cdf_in, int64_t R,
                                                                     float* __restrict__ z_samples,
                                                                     float* __restrict__ z_merged, float* __restrict__ z_std,
                                                                     int* __restrict__ inds_out) {
  constexpr int S = 4 * LANES, N = KN * LANES, B = S - 1, RPW = 32 / LANES;
  constexpr int LOGB = S == 32 ? 5 : (S == 64 ? 6 : 7);                        // B = 2^LOGB - 1
  constexpr int LOGS1 = LOGB + 1;                                              // counts 0 .. S
  constexpr int LOGN1 = N <= 32 ? 6 : (N <= 64 ? 7 : (N <= 128 ? 8 : 9));      // counts 0 .. N
  constexpr int kPerRay = 5 * S + 2 * N + 4;                                   // cdf | bins | z | samples | merged | hist
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int sub = lane / LANES, l = lane % LANES;
  float* const base = smem + ((size_t)wib * RPW + sub) * kPerRay;
  // deterministic path: (cdf[i], bin midpoint[i]) pairs, one 8-byte gather per end of a sample's bin; the random-u path
  // searches the cdf (unit stride keeps its probes on fewer banks) and keeps the two arrays apart
  float2* const cb = reinterpret_cast<float2*>(base);
  float* const cdf = base;
  float* const bn = base + S;
  float* const zv = base + 2 * S;
  float* const smp = zv + S;
  float* const mrg = smp + N;
  int* const hist = reinterpret_cast<int*>(mrg + S + N);                       // S + 4 counters
  const int64_t ngroups = (R + RPW - 1) / RPW;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  // operands of the NEXT group are fetched into registers while the current one is processed: the loop body is a long
  // dependent chain (scan -> search -> merge), and without the prefetch every iteration also paid a full DRAM round trip
  float4 zq_n = make_float4(0.f, 0.f, 0.f, 0.f), wq_n = zq_n;
  float u_n[DET ? 1 : KN];
  auto fetch = [&](int64_t g) {
    const int64_t r0 = g * RPW + sub;
    const int64_t r1 = r0 < R ? r0 : R - 1;                                    // the odd last half-warp re-does the last ray
    zq_n = ld_stream4(reinterpret_cast<const float4*>(z_vals + r1 * S + l * 4));
    if (cdf_in == nullptr) wq_n = ld_stream4(reinterpret_cast<const float4*>(weights + r1 * S + l * 4));
    if (!DET) {
      if (KN == 2) {
        const float2 q = *reinterpret_cast<const float2*>(u + r1 * N + l * KN);
        u_n[0] = q.x; u_n[KN > 1 ? 1 : 0] = q.y;
      } else {
#pragma unroll
        for (int k = 0; k < KN; k += 4) {
          const float4 q = ld_stream4(reinterpret_cast<const float4*>(u + r1 * N + l * KN + k));
          u_n[k] = q.x; u_n[k + 1 < KN ? k + 1 : k] = q.y; u_n[k + 2 < KN ? k + 2 : k] = q.z; u_n[k + 3 < KN ? k + 3 : k] = q.w;
        }
      }
    }
  };
  // deterministic samples: u = linspace(0, 1, N) is the same for every ray - this lane's KN values, once
  float ud[KN];
#pragma unroll
  for (int k = 0; k < KN; ++k) ud[k] = linspace01(l * KN + k, N);
  int* const fh = reinterpret_cast<int*>(smp);                                 // DET: histogram of first-sample indices (below)
  const int64_t grp0 = (int64_t)blockIdx.x * kWarpsPerBlock + wib;
  if (grp0 < ngroups) fetch(grp0);
  for (int64_t grp = grp0; grp < ngroups; grp += nwarps) {
    const int64_t ray = grp * RPW + sub;
    const bool valid = ray < R;
    const int64_t rr = valid ? ray : R - 1;
    const float4 zq = zq_n, wq = wq_n;
    float uu[KN];
    if (!DET) {
#pragma unroll
      for (int k = 0; k < KN; ++k) uu[k] = u_n[k];
    }
    if (grp + nwarps < ngroups) fetch(grp + nwarps);
    const float zz[4] = {zq.x, zq.y, zq.z, zq.w};
    *reinterpret_cast<float4*>(zv + l * 4) = zq;
    float c[4];
    if (cdf_in != nullptr) {                                                   // externally supplied cdf [R, S - 1] (test hook)
#pragma unroll
      for (int k = 0; k < 4; ++k) c[k] = (l * 4 + k < B) ? __ldg(cdf_in + rr * B + l * 4 + k) : 1.f;
    } else {
      const float ww[4] = {wq.x, wq.y, wq.z, wq.w};
      float run = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {                                            // q_i = w_i + 1e-5 for 1 <= i <= S - 2 (helpers:308)
        const int i = l * 4 + k;
        run += (i >= 1 && i <= S - 2) ? __fadd_rn(ww[k], 1e-5f) : 0.f;
        c[k] = run;
      }
      float incl = run;
#pragma unroll
      for (int o = 1; o < LANES; o <<= 1) {
        const float t = __shfl_up_sync(kFullMask, incl, o, LANES);
        if (l >= o) incl += t;
      }
      const float rtot = __frcp_rn(__shfl_sync(kFullMask, incl, LANES - 1, LANES));
      const float excl = incl - run;
#pragma unroll
      for (int k = 0; k < 4; ++k) c[k] = (excl + c[k]) * rtot;                 // cdf[i] = sum_{1 <= m <= i} q_m / sum q
    }
    const float znext = __shfl_down_sync(kFullMask, zz[0], 1, LANES);
    if constexpr (DET) {
      *reinterpret_cast<float4*>(cb + l * 4) = make_float4(c[0], .5f * (zz[1] + zz[0]), c[1], .5f * (zz[2] + zz[1]));
      *reinterpret_cast<float4*>(cb + l * 4 + 2) = make_float4(c[2], .5f * (zz[3] + zz[2]), c[3], .5f * (znext + zz[3]));
    } else {
      *reinterpret_cast<float4*>(cdf + l * 4) = make_float4(c[0], c[1], c[2], c[3]);
      *reinterpret_cast<float4*>(bn + l * 4) = make_float4(.5f * (zz[1] + zz[0]), .5f * (zz[2] + zz[1]), .5f * (zz[3] + zz[2]),
                                                           .5f * (znext + zz[3]));
    }
    if (DET) {
      *reinterpret_cast<int4*>(hist + l * 4) = make_int4(0, 0, 0, 0);
      if (l == 0) *reinterpret_cast<int4*>(hist + S) = make_int4(0, 0, 0, 0);
      if constexpr (KN % 4 == 0) {
#pragma unroll
        for (int k = 0; k < KN; k += 4) *reinterpret_cast<int4*>(fh + l * KN + k) = make_int4(0, 0, 0, 0);
      } else {
#pragma unroll
        for (int k = 0; k < KN; k += 2) *reinterpret_cast<int2*>(fh + l * KN + k) = make_int2(0, 0);
      }
    }
    __syncwarp();
    // ---- KN consecutive samples per lane --------------------------------------------------------------------------
    float sv[KN];
    int inds[KN];
    if (DET) {
      // searchsorted(cdf, u, right=True)[j] = #{i : cdf[i] <= u_j}.  The u_j are ascending and known in closed form, so
      // the search is turned around: f_i = #{j : u_j < cdf[i]} (the first sample at or above cdf[i]) follows from one
      // multiplication and two exact comparisons against linspace values, cdf[i] <= u_j  <=>  f_i <= j, and the index of
      // sample j is a prefix sum over a histogram of the f_i - the cdf entries stay in registers and 6 dependent,
      // bank-conflicting shared-memory probes per sample (70 % of this kernel's shared-memory wavefronts) disappear.
#pragma unroll
      for (int k = 0; k < KN; ++k) uu[k] = ud[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (l * 4 + k < B) {
          // j0 = trunc(cdf * (N - 1)): u_j < cdf for every j < j0 (margin 1 / (N - 1) against rounding errors of 1e-7), and
          // u_{j0 + 2} > cdf for the same reason, so f is j0 plus the outcome of two exact comparisons; linspace01() is
          // evaluated branch-free (both sides of its two-sided formula, then a select)
          const float ck = c[k];
          int j0 = (int)(ck * (float)(N - 1));
          j0 = j0 < 0 ? 0 : (j0 > N ? N : j0);
          constexpr float kStep = 1.f / (float)(N - 1);
          const float a0 = (float)j0, b0 = (float)(N - 1 - j0);
          const float ua = j0 < N / 2 ? __fmul_rn(kStep, a0) : __fsub_rn(1.f, __fmul_rn(kStep, b0));
          const float ub = j0 + 1 < N / 2 ? __fmul_rn(kStep, a0 + 1.f) : __fsub_rn(1.f, __fmul_rn(kStep, b0 - 1.f));
          const int f = j0 + ((j0 < N && ua < ck) ? 1 + ((j0 + 1 < N && ub < ck) ? 1 : 0) : 0);
          if (f < N) atomicAdd(&fh[f], 1);
        }
      }
      __syncwarp();
      int run = 0;
      if constexpr (KN % 4 == 0) {
#pragma unroll
        for (int k = 0; k < KN; k += 4) {
          const int4 q = *reinterpret_cast<const int4*>(fh + l * KN + k);
          inds[k] = run + q.x; inds[k + 1] = inds[k] + q.y; inds[k + 2] = inds[k + 1] + q.z; inds[k + 3] = inds[k + 2] + q.w;
          run = inds[k + 3];
        }
      } else {
#pragma unroll
        for (int k = 0; k < KN; k += 2) {
          const int2 q = *reinterpret_cast<const int2*>(fh + l * KN + k);
          inds[k] = run + q.x; inds[k + 1] = inds[k] + q.y;
          run = inds[k + 1];
        }
      }
      int incl = run;
#pragma unroll
      for (int o = 1; o < LANES; o <<= 1) {
        const int t = __shfl_up_sync(kFullMask, incl, o, LANES);
        if (l >= o) incl += t;
      }
#pragma unroll
      for (int k = 0; k < KN; ++k) inds[k] += incl - run;
    } else {
#pragma unroll
      for (int k = 0; k < KN; ++k) inds[k] = count_below<LOGB, true>(cdf, B, uu[k]);
    }
    float sum = 0.f;
    int los[KN];
#pragma unroll
    for (int k = 0; k < KN; ++k) {
      const int ind = inds[k];
      const int lo = ind - 1 < 0 ? 0 : ind - 1;
      const int hi = ind > B - 1 ? B - 1 : ind;
      float2 p0, p1;
      if constexpr (DET) { p0 = cb[lo]; p1 = cb[hi]; }
      else { p0 = make_float2(cdf[lo], bn[lo]); p1 = make_float2(cdf[hi], bn[hi]); }
      float den = p1.x - p0.x;
      if (den < 1e-5f) den = 1.f;
      const float t = __fdividef(uu[k] - p0.x, den);
      sv[k] = fmaf(t, p1.y - p0.y, p0.y);
      sum += sv[k];
      los[k] = lo;
    }
    if (inds_out != nullptr) {                                                 // test hook: uniform branch, off in the render path
#pragma unroll
      for (int k = 0; k < KN; ++k)
        if (valid) inds_out[ray * N + l * KN + k] = inds[k];
    }
    if (z_samples != nullptr && valid) {
      if constexpr (KN % 4 == 0) {
        if ((reinterpret_cast<uintptr_t>(z_samples) & 15) == 0) {
#pragma unroll
          for (int k = 0; k < KN; k += 4)
            st_stream4(reinterpret_cast<float4*>(z_samples + ray * N + l * KN + k), make_float4(sv[k], sv[k + 1], sv[k + 2], sv[k + 3]));
        } else {
#pragma unroll
          for (int k = 0; k < KN; ++k) st_stream(z_samples + ray * N + l * KN + k, sv[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < KN; ++k) st_stream(z_samples + ray * N + l * KN + k, sv[k]);
      }
    }
    if (z_std != nullptr) {                                                    // std(unbiased=False), run.py:2370
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(kFullMask, sum, o);
      const float mean = sum * (1.f / (float)N);
      float var = 0.f;
#pragma unroll
      for (int k = 0; k < KN; ++k) { const float dv = sv[k] - mean; var = fmaf(dv, dv, var); }
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) var += __shfl_xor_sync(kFullMask, var, o);
      if (l == 0 && valid) z_std[ray] = sqrtf(var * (1.f / (float)N));
    }
    // ---- merge with the coarse depths -------------------------------------------------------------------------------
    if (DET) {
      // a sample from bin lo lies between z[lo] and z[lo + 2]: rank among the coarse depths = lo + 1 + (z[lo + 1] <= s);
      // the samples are ascending (u is), so sample n lands at n + rank; coarse depth i at i + #{samples ranked <= i}
#pragma unroll
      for (int k = 0; k < KN; ++k) {
        const int cnt = los[k] + 1 + (zv[los[k] + 1] <= sv[k] ? 1 : 0);
        mrg[l * KN + k + cnt] = sv[k];
        atomicAdd(&hist[cnt], 1);
      }
      __syncwarp();
      const int4 hq = *reinterpret_cast<const int4*>(hist + l * 4);
      const int h1 = hq.x, h2 = h1 + hq.y, h3 = h2 + hq.z, h4 = h3 + hq.w;
      int incl = h4;
#pragma unroll
      for (int o = 1; o < LANES; o <<= 1) {
        const int t = __shfl_up_sync(kFullMask, incl, o, LANES);
        if (l >= o) incl += t;
      }
      const int excl = incl - h4;
      mrg[l * 4 + 0 + excl + h1] = zz[0];
      mrg[l * 4 + 1 + excl + h2] = zz[1];
      mrg[l * 4 + 2 + excl + h3] = zz[2];
      mrg[l * 4 + 3 + excl + h4] = zz[3];
    } else {
      seg_bitonic_sort<LANES, KN>(sv, l);
#pragma unroll
      for (int k = 0; k < KN; ++k) smp[l * KN + k] = sv[k];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) mrg[l * 4 + k + count_below<LOGN1, false>(smp, N, zz[k])] = zz[k];
#pragma unroll
      for (int k = 0; k < KN; ++k) mrg[l * KN + k + count_below<LOGS1, true>(zv, S, sv[k])] = sv[k];
    }
    __syncwarp();
    // merged rows of the warp's RPW consecutive rays are contiguous in memory: 16-byte stores
    {
      const float* wm = smem + (size_t)wib * RPW * kPerRay;
      const int64_t first = grp * RPW;
      const int nvalid = (int)((R - first) < RPW ? (R - first) : RPW);
      float4* out = reinterpret_cast<float4*>(z_merged + first * (S + N));
      constexpr int per_ray4 = (S + N) / 4;
      if (nvalid == RPW && (RPW * per_ray4) % 32 == 0) {                       // whole groups: fixed trip count
#pragma unroll
        for (int it = 0; it < (RPW * per_ray4) / 32; ++it) {
          const int idx = it * 32 + lane;
          const int r = idx / per_ray4, o = idx - r * per_ray4;
          st_stream4(out + idx, *reinterpret_cast<const float4*>(wm + r * kPerRay + 3 * S + N + 4 * o));
        }
      } else {
        for (int idx = lane; idx < nvalid * per_ray4; idx += 32) {
          const int r = idx / per_ray4, o = idx - r * per_ray4;
          st_stream4(out + idx, *reinterpret_cast<const float4*>(wm + (size_t)r * kPerRay + 3 * S + N + 4 * o));
        }
      }
    }
    __syncwarp();
  }
}

template <int LANES, int KN>
static int launch_sample_merge_cons(const float* z_vals, const float* weights, const float* u, const float* cdf_in, int64_t R,
                                    float* z_samples, float* z_merged, float* z_std, int* inds_out, cudaStream_t stream) {
  constexpr int S = 4 * LANES, N = KN * LANES, RPW = 32 / LANES;
  constexpr size_t smem = (size_t)kWarpsPerBlock * RPW * (5 * S + 2 * N + 4) * sizeof(float);
  static_assert(smem <= 48 * 1024, "sample_merge_cons: shared memory above the default limit");
  // resident CTAs per SM: 4 by registers (__launch_bounds__(256, 4), <= 64 registers), fewer if shared memory says so; a
  // persistent grid larger than what is resident runs its surplus CTAs as a second, half-empty wave
  const int per_sm = (int)((200 * 1024) / smem);
  // (5 or 6 CTAs per SM at 48 / 40 registers: no faster - the kernel is bound by shared-memory wavefronts and issue slots)
  const int grid = persistent_grid((R + RPW - 1) / RPW, per_sm > 4 ? 4 : per_sm);
  if (u) sample_merge_cons_kernel<LANES, KN, false><<<grid, kThreads, smem, stream>>>(z_vals, weights, u, cdf_in, R, z_samples, z_merged, z_std, inds_out);
  else sample_merge_cons_kernel<LANES, KN, true><<<grid, kThreads, smem, stream>>>(z_vals, weights, u, cdf_in, R, z_samples, z_merged, z_std, inds_out);
  return check_launch("sample_merge_cons_kernel");
}

// =========================================================================================================
// loss seed — img2mse terms, run.py:1483,1502,1513-1515
// =========================================================================================================
__global__ void __launch_bounds__(256) loss_seed_kernel(const float* __restrict__ rgb, const float* __restrict__ rgb0,
                                                        const float* __restrict__ disp, const float* __restrict__ trgb,
                                                        const float* __restrict__ tdisp, int64_t R, float inv3R,
                                                        float invR, float lambda, float* __restrict__ g_rgb,
                                                        float* __restrict__ g_rgb0, float* __restrict__ g_disp,
                                                        float* __restrict__ loss) {
  float part = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < 3 * R; i += (int64_t)gridDim.x * blockDim.x) {
    const float t = trgb[i];
    const float e1 = rgb[i] - t, e0 = rgb0 ? rgb0[i] - t : 0.f;
    g_rgb[i] = 2.f * e1 * inv3R;
    if (rgb0) g_rgb0[i] = 2.f * e0 * inv3R;
    part += (e1 * e1 + e0 * e0) * inv3R;
    if (i < R && disp != nullptr) {
      const float ed = disp[i] - tdisp[i];
      g_disp[i] = 2.f * lambda * ed * invR;
      part += lambda * ed * ed * invR;
    }
  }
  part = warp_sum(part);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(loss, s);
  }
}



int64_t launch_composite_fwd_staged(const float* raw, const float* z, const float* rays_d, int64_t ray_stride,
                                    const float* noise, int64_t R, int S, int white, float* rgb, float* disp, float* acc,
                                    float* depth, float* weights, float* alpha, cudaStream_t stream, int* rc);  // render_staged.cu
int launch_composite_bwd_staged(const float* raw, const float* z, const float* rays_d, int64_t ray_stride, const float* noise,
                                int64_t R, int S, int white, int detach_w, const float* g_rgb, const float* g_disp,
                                const float* g_acc, const float* g_depth, const float* g_w, float* g_raw, cudaStream_t stream,
                                int* rc);  // render_staged.cu
}  // namespace gbn

// =========================================================================================================
// C ABI
// =========================================================================================================
using namespace gbn;

extern "C" int gbn_zvals_stratified(const float* near, const float* far, int64_t ray_stride, int64_t R, int S,
                                    int lindisp, const float* t_rand, float* z, void* stream) {
  if (R == 0) return GBN_OK;
  GBN_REQUIRE(near && far && z, "zvals_stratified: null pointer");
  GBN_REQUIRE(R >= 0 && S >= 1 && ray_stride >= 1, "zvals_stratified: bad sizes R=%lld S=%d", (long long)R, S);
  const int64_t blocks = (R * S + 255) / 256;
  const int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
  const bool vec = ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(t_rand)) & 15) == 0;
#define GBN_ZV(SS) if (S == SS && vec) { const int64_t bl = (R * (SS / 4) + 255) / 256; const int g = (int)(bl < kNumSMs * 8 ? bl : kNumSMs * 8); \
    zvals_fixed_kernel<SS><<<g, 256, 0, (cudaStream_t)stream>>>(near, far, ray_stride, R, lindisp, t_rand, z); return check_launch("zvals_fixed_kernel"); }
  GBN_ZV(64) GBN_ZV(128) GBN_ZV(32) GBN_ZV(256)
#undef GBN_ZV
  zvals_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(near, far, ray_stride, R, S, lindisp, t_rand, z);
  return check_launch("zvals_kernel");
}

extern "C" int gbn_encode_points(const float* rays_o, const float* rays_d, const float* viewdirs, int64_t ray_stride,
                                 const float* z, int64_t R, int S, float* out, void* stream) {
  if (R == 0) return GBN_OK;
  GBN_REQUIRE(rays_o && rays_d && viewdirs && z && out, "encode_points: null pointer");
  GBN_REQUIRE(R >= 0 && S >= 1, "encode_points: bad sizes");
  GBN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "encode_points: out must be 16-byte aligned");
  const int64_t ntiles = (R * S + kEncPts - 1) / kEncPts;
  const int grid = (int)(ntiles < kNumSMs * 8 ? ntiles : kNumSMs * 8);
  encode_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, viewdirs, ray_stride, z, R, S, out);
  return check_launch("encode_kernel");
}

#define GBN_DISPATCH_NCH(S, CALL)                                                  \
  do {                                                                             \
    const int nch__ = ((S) + 31) / 32;                                             \
    if (nch__ <= 1) { CALL(1); } else if (nch__ <= 2) { CALL(2); }                 \
    else if (nch__ <= 3) { CALL(3); } else if (nch__ <= 4) { CALL(4); }            \
    else if (nch__ <= 6) { CALL(6); } else if (nch__ <= 8) { CALL(8); }            \
    else if (nch__ <= 12) { CALL(12); } else if (nch__ <= 16) { CALL(16); }        \
    else { CALL(32); }                                                             \
  } while (0)

extern "C" int gbn_composite_forward(const float* raw, const float* z, const float* rays_d, int64_t ray_stride,
                                     const float* noise, int64_t R, int S, int white_bkgd, float* rgb, float* disp,
                                     float* acc, float* depth, float* weights, float* alpha, void* stream) {
  if (R == 0) return GBN_OK;
  GBN_REQUIRE(raw && z && rays_d && rgb && disp && acc && depth && weights, "composite_forward: null pointer");
  GBN_REQUIRE(R >= 0 && S >= 1 && S <= 1024, "composite_forward: S=%d outside [1,1024]", S);
  GBN_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "composite_forward: raw must be 16-byte aligned");
  // bulk-copy pipelined kernel for the leading groups of 8 rays, generic warp-per-ray kernel for the rest
  int rc = GBN_OK;
  const int64_t done = launch_composite_fwd_staged(raw, z, rays_d, ray_stride, noise, R, S, white_bkgd, rgb, disp, acc,
                                                   depth, weights, alpha, (cudaStream_t)stream, &rc);
  if (rc != GBN_OK) return rc;
  if (done == R) return GBN_OK;
  raw += done * S * 4; z += done * S; rays_d += done * ray_stride;
  if (noise) noise += done * S;
  rgb += done * 3; disp += done; acc += done; depth += done; weights += done * S;
  if (alpha) alpha += done * S;
  R -= done;
  const int grid = persistent_grid(R, 8);
#define CALL(N) composite_fwd_kernel<N><<<grid, kThreads, 0, (cudaStream_t)stream>>>( \
      raw, z, rays_d, ray_stride, noise, R, S, white_bkgd, rgb, disp, acc, depth, weights, alpha)
  GBN_DISPATCH_NCH(S, CALL);
#undef CALL
  return check_launch("composite_fwd_kernel");
}

extern "C" int gbn_composite_backward(const float* raw, const float* z, const float* rays_d, int64_t ray_stride,
                                      const float* noise, int64_t R, int S, int white_bkgd, int detach_weights,
                                      const float* g_rgb, const float* g_disp, const float* g_acc,
                                      const float* g_depth, const float* g_weights, float* g_raw, void* stream) {
  if (R == 0) return GBN_OK;
  GBN_REQUIRE(raw && z && rays_d && g_raw, "composite_backward: null pointer");
  GBN_REQUIRE(R >= 0 && S >= 1 && S <= 1024, "composite_backward: S=%d outside [1,1024]", S);
  GBN_REQUIRE(((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(g_raw)) & 15) == 0,
              "composite_backward: raw/g_raw must be 16-byte aligned");
  {
    int rc = GBN_OK;
    if (launch_composite_bwd_staged(raw, z, rays_d, ray_stride, noise, R, S, white_bkgd, detach_weights, g_rgb, g_disp, g_acc,
                                    g_depth, g_weights, g_raw, (cudaStream_t)stream, &rc))
      return rc;
  }
  const int grid = persistent_grid(R, 6);
#define CALL(N) composite_bwd_kernel<N><<<grid, kThreads, 0, (cudaStream_t)stream>>>(                   \
      raw, z, rays_d, ray_stride, noise, R, S, white_bkgd, detach_weights, g_rgb, g_disp, g_acc, g_depth, \
      g_weights, g_raw)
  GBN_DISPATCH_NCH(S, CALL);
#undef CALL
  return check_launch("composite_bwd_kernel");
}

extern "C" int gbn_searchsorted_right(const float* cdf, const float* u, int64_t R, int B, int N, int64_t* inds,
                                      void* stream) {
  if (R == 0) return GBN_OK;
  GBN_REQUIRE(cdf && u && inds, "searchsorted_right: null pointer");
  GBN_REQUIRE(R >= 0 && B >= 1 && N >= 1 && B <= 4096, "searchsorted_right: bad sizes B=%d N=%d", B, N);
  const size_t smem = (size_t)kWarpsPerBlock * B * sizeof(float);
  static bool attr_set = false;  // immutable kernel attribute, set once
  if (!attr_set) {
    GBN_CUDA(cudaFuncSetAttribute(searchsorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  searchsorted_kernel<<<persistent_grid(R, 4), kThreads, smem, (cudaStream_t)stream>>>(cdf, u, R, B, N, inds);
  return check_launch("searchsorted_kernel");
}

extern "C" int gbn_sample_pdf(const float* bins, const float* weights, const float* u, int64_t R, int B, int N,
                              float* samples, void* stream) {
  if (R == 0) return GBN_OK;
  GBN_REQUIRE(bins && weights && samples, "sample_pdf: null pointer");
  GBN_REQUIRE(R >= 0 && B >= 2 && N >= 1 && B <= 2048, "sample_pdf: bad sizes B=%d N=%d", B, N);
  const size_t smem = (size_t)kWarpsPerBlock * 2 * B * sizeof(float);
  static bool attr_set = false;  // immutable kernel attribute, set once
  if (!attr_set) {
    GBN_CUDA(cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  sample_pdf_kernel<<<persistent_grid(R, 4), kThreads, smem, (cudaStream_t)stream>>>(bins, weights, u, R, B, N, samples);
  return check_launch("sample_pdf_kernel");
}

extern "C" int gbn_sample_pdf_merge_ex(const float* z_vals, const float* weights, const float* u, const float* cdf_in,
                                       int64_t R, int S, int N, float* z_samples, float* z_merged, float* z_std,
                                       int* inds_out, void* stream) {
  if (R == 0) return GBN_OK;
  GBN_REQUIRE(z_vals && (weights || cdf_in) && z_merged, "sample_pdf_merge: null pointer");
  GBN_REQUIRE(R >= 0 && S >= 3 && N >= 1, "sample_pdf_merge: need S>=3, N>=1 (S=%d N=%d)", S, N);
  {
    static const bool generic_only = [] { const char* e = getenv("GBNERF_SAMPLE_GENERIC"); return e && e[0] == '1'; }();
    cudaStream_t st = (cudaStream_t)stream;
    const uintptr_t al = reinterpret_cast<uintptr_t>(z_vals) | reinterpret_cast<uintptr_t>(weights) | reinterpret_cast<uintptr_t>(u) |
                         reinterpret_cast<uintptr_t>(z_merged);
    // the render path's shapes: S = 4 LANES, N = KN LANES (rows are then multiples of 16 bytes: vector accesses)
#define GBN_SM_CONS(LANES, KN) \
  if (!generic_only && (al & 15) == 0 && S == 4 * LANES && N == KN * LANES) \
    return launch_sample_merge_cons<LANES, KN>(z_vals, weights, u, cdf_in, R, z_samples, z_merged, z_std, inds_out, st)
    GBN_SM_CONS(16, 4); GBN_SM_CONS(16, 8); GBN_SM_CONS(16, 2); GBN_SM_CONS(32, 2); GBN_SM_CONS(32, 4); GBN_SM_CONS(32, 8);
    GBN_SM_CONS(8, 4);
#undef GBN_SM_CONS
  }
  int npad = 1;
  while (npad < N) npad <<= 1;
  const size_t smem = (size_t)kWarpsPerBlock * (5 * (size_t)S + npad + N + 1) * sizeof(float);
  GBN_REQUIRE(smem <= 200 * 1024, "sample_pdf_merge: S=%d N=%d needs %zu B of shared memory", S, N, smem);
  static bool attr_set = false;  // immutable kernel attribute, set once
  if (!attr_set) {
    GBN_CUDA(cudaFuncSetAttribute(sample_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int per_sm = smem > 0 ? (int)((220 * 1024) / smem) : 4;
  sample_merge_kernel<<<persistent_grid(R, per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm)), kThreads, smem,
                        (cudaStream_t)stream>>>(z_vals, weights, u, cdf_in, R, S, N, npad, z_samples, z_merged, z_std, inds_out);
  return check_launch("sample_merge_kernel");
}

extern "C" int gbn_sample_pdf_merge(const float* z_vals, const float* weights, const float* u, int64_t R, int S,
                                    int N, float* z_samples, float* z_merged, float* z_std, void* stream) {
  return gbn_sample_pdf_merge_ex(z_vals, weights, u, nullptr, R, S, N, z_samples, z_merged, z_std, nullptr, stream);
}

extern "C" int gbn_loss_seed(const float* rgb, const float* rgb0, const float* disp, const float* target_rgb,
                             const float* target_disp, int64_t R, int64_t R_global, float depth_lambda, float* g_rgb,
                             float* g_rgb0, float* g_disp, float* loss, void* stream) {
  if (R == 0) return GBN_OK;
  GBN_REQUIRE(rgb && target_rgb && g_rgb && loss, "loss_seed: null pointer");
  GBN_REQUIRE((rgb0 == nullptr) == (g_rgb0 == nullptr), "loss_seed: rgb0 and g_rgb0 go together");
  GBN_REQUIRE((disp == nullptr) || (target_disp && g_disp), "loss_seed: disp needs target_disp and g_disp");
  GBN_REQUIRE(R >= 0 && R_global >= R && R_global > 0, "loss_seed: bad sizes");
  const int64_t blocks = (3 * R + 255) / 256;
  const int grid = (int)(blocks < kNumSMs * 4 ? blocks : kNumSMs * 4);
  loss_seed_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rgb, rgb0, disp, target_rgb, target_disp, R,
                                                           1.f / (3.f * (float)R_global), 1.f / (float)R_global,
                                                           depth_lambda, g_rgb, g_rgb0, g_disp, loss);
  return check_launch("loss_seed_kernel");
}
