// Shared host/device helpers for libgbnerf.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gbnerf.h"

namespace gbn {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr unsigned kFullMask = 0xffffffffu;

// ---- error reporting (thread-local, never throws) -------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define GBN_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::gbn::set_error(__VA_ARGS__);      \
      return GBN_EINVAL;                  \
    }                                     \
  } while (0)

#define GBN_CUDA(call)                                      \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return ::gbn::cuda_fail(e__, #call); \
  } while (0)

void count_launch();   // abi.cu: one more kernel of this library went onto a stream (gbn_kernel_launches)

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return cuda_fail(e, what);
  }
  count_launch();
  return GBN_OK;
}

// ---- warp primitives ------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// inclusive prefix product / sum over the 32 lanes
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// inclusive SUFFIX sum: lane i gets sum_{k >= i} v_k
__device__ __forceinline__ float warp_rscan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(kFullMask, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

// torch.linspace(0, 1, n)[i] in fp32 (ATen RangeFactories: symmetric two-sided formula)
__device__ __forceinline__ float linspace01(int i, int n) {
  if (n <= 1) return 0.f;
  const float step = __fdiv_rn(1.f, (float)(n - 1));
  return (i < n / 2) ? __fmul_rn(step, (float)i) : __fsub_rn(1.f, __fmul_rn(step, (float)(n - i - 1)));
}

// streaming 16-byte accesses that do not pollute L1
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---- positional encoding of one coordinate --------------------------------------------------------------
// out[2k] = sin(2^k x), out[2k+1] = cos(2^k x), k < L.  The argument 2^k x is exact in fp32, the reference
// evaluates sin/cos of it directly (run_nerf_helpers.py:46).  Here: an accurate sincosf every 4th octave and
// the double-angle recurrence in between, so the error stays below 16 x 1 ulp-of-one (~1e-6) at a third of
// the cost of L sincosf calls.
template <int L>
__device__ __forceinline__ void posenc_axis(float x, float* out) {
#pragma unroll
  for (int k = 0; k < L; ++k) {
    if ((k & 3) == 0) {
      sincosf(x * (float)(1 << k), &out[2 * k], &out[2 * k + 1]);
    } else {
      const float s = out[2 * k - 2], c = out[2 * k - 1];
      out[2 * k] = 2.f * s * c;
      out[2 * k + 1] = fmaf(-2.f * s, s, 1.f);
    }
  }
}

}  // namespace gbn
