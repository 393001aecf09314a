// Ray setup of render() in one kernel: get_rays (run_nerf_helpers.py:251-262) for a camera pose or
// a caller-supplied ray batch, view-direction normalisation (run.py:1711-1718), the optional NDC warp
// (run_nerf_helpers.py:285-302, called with near = 1 at run.py:1723) and the packing of the [R, 8 (+1) (+3)] ray batch
// (run.py:1726-1736).  The reference spends ~15 elementwise/reduce/cat launches and half a dozen [H*W,3]
// temporaries on this; here one thread computes one ray and a CTA writes its rows as one contiguous block.
//
// Arithmetic follows the reference operation by operation with explicitly rounded (non-fused) multiplies and adds,
// so results agree with torch to the last bit except where torch's own reduction order is unspecified (the
// three-term sums of the rotation and of the norm; tests allow 2 ulp there).
#include "common.cuh"

namespace gbn {
namespace {

constexpr int kRayThreads = 256;

struct RaySetupArgs {
  const float* c2w;        // [3, >=4] row-major, pitch c2w_ld floats (NULL: rays_o / rays_d are given)
  const float* c2w_static; // optional second pose: origins/directions come from it, view directions from c2w
  int c2w_ld, c2w_static_ld;
  const float* rays_o;     // [R,3] pitch o_ld
  const float* rays_d;     // [R,3] pitch d_ld
  int64_t o_ld, d_ld;
  const float* depths;     // [R] or NULL
  int H, W;
  float focal, half_w, half_h;
  int pi, pj, len2;        // patch origin (row, col) and patch width; full frame: 0, 0, W
  int use_viewdirs, ndc;
  float ndc_ax, ndc_ay;    // -1 / (W / (2 focal)), -1 / (H / (2 focal)) rounded from double like torch's scalars
  float near, far;
  int64_t R;
  int row;                 // floats per output row
  float* out;
};

__device__ __forceinline__ void camera_ray(const float* __restrict__ c2w, int ld, float cx, float cy, float o[3], float d[3]) {
  // dirs = [(i - W/2)/f, -(j - H/2)/f, -1];  rays_d[k] = sum_c dirs[c] * c2w[k][c]
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float p0 = __fmul_rn(cx, __ldg(c2w + k * ld + 0));
    const float p1 = __fmul_rn(cy, __ldg(c2w + k * ld + 1));
    const float p2 = __fmul_rn(-1.f, __ldg(c2w + k * ld + 2));
    d[k] = __fadd_rn(__fadd_rn(p0, p1), p2);
    o[k] = __ldg(c2w + k * ld + 3);
  }
}

__global__ void __launch_bounds__(kRayThreads) ray_setup_kernel(const RaySetupArgs a) {
  extern __shared__ float s_rows[];   // kRayThreads rows of a.row floats
  const int64_t base = (int64_t)blockIdx.x * kRayThreads;
  const int64_t r = base + threadIdx.x;
  if (r < a.R) {
    float o[3], d[3], v[3];
    if (a.c2w) {
      const int pr = (int)(r / a.len2), pc = (int)(r - (int64_t)pr * a.len2);
      const float cx = __fdiv_rn(__fsub_rn((float)(a.pj + pc), a.half_w), a.focal);
      const float cy = -__fdiv_rn(__fsub_rn((float)(a.pi + pr), a.half_h), a.focal);
      camera_ray(a.c2w, a.c2w_ld, cx, cy, o, d);
      v[0] = d[0]; v[1] = d[1]; v[2] = d[2];
      if (a.c2w_static) camera_ray(a.c2w_static, a.c2w_static_ld, cx, cy, o, d);
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        o[k] = __ldg(a.rays_o + r * a.o_ld + k);
        d[k] = __ldg(a.rays_d + r * a.d_ld + k);
        v[k] = d[k];
      }
    }
    if (a.use_viewdirs) {   // viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
      const float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
      v[0] = __fdiv_rn(v[0], n); v[1] = __fdiv_rn(v[1], n); v[2] = __fdiv_rn(v[2], n);
    }
    if (a.ndc) {            // ndc_rays(H, W, focal, near = 1., rays_o, rays_d)
      const float t = __fdiv_rn(-__fadd_rn(1.f, o[2]), d[2]);
#pragma unroll
      for (int k = 0; k < 3; ++k) o[k] = __fadd_rn(o[k], __fmul_rn(t, d[k]));
      const float rz = __frcp_rn(o[2]);               // python-scalar / tensor is reciprocal(tensor) * scalar in torch
      const float ox = __fdiv_rn(o[0], o[2]), oy = __fdiv_rn(o[1], o[2]);
      const float n0 = __fdiv_rn(__fmul_rn(a.ndc_ax, o[0]), o[2]);
      const float n1 = __fdiv_rn(__fmul_rn(a.ndc_ay, o[1]), o[2]);
      const float n2 = __fadd_rn(1.f, __fmul_rn(rz, 2.f));
      const float e0 = __fmul_rn(a.ndc_ax, __fsub_rn(__fdiv_rn(d[0], d[2]), ox));
      const float e1 = __fmul_rn(a.ndc_ay, __fsub_rn(__fdiv_rn(d[1], d[2]), oy));
      const float e2 = __fmul_rn(rz, -2.f);
      o[0] = n0; o[1] = n1; o[2] = n2; d[0] = e0; d[1] = e1; d[2] = e2;
    }
    float* row = s_rows + (size_t)threadIdx.x * a.row;
    row[0] = o[0]; row[1] = o[1]; row[2] = o[2];
    row[3] = d[0]; row[4] = d[1]; row[5] = d[2];
    row[6] = a.near; row[7] = a.far;
    int c = 8;
    if (a.depths) row[c++] = __ldg(a.depths + r);
    if (a.use_viewdirs) { row[c] = v[0]; row[c + 1] = v[1]; row[c + 2] = v[2]; }
  }
  __syncthreads();
  const int64_t rows_here = a.R - base < kRayThreads ? a.R - base : kRayThreads;
  const int n = (int)rows_here * a.row;
  float* dst = a.out + base * a.row;
  for (int i = threadIdx.x; i < n; i += kRayThreads) dst[i] = s_rows[i];   // consecutive rows are contiguous in the batch
}

}  // namespace
}  // namespace gbn

using namespace gbn;

extern "C" int gbn_pack_rays(const float* c2w, int c2w_ld, const float* c2w_static, int c2w_static_ld, const float* rays_o,
                             int64_t o_ld, const float* rays_d, int64_t d_ld, const float* depths, int H, int W, double focal,
                             int patch_i, int patch_j, int patch_h, int patch_w, int use_viewdirs, int ndc, float near,
                             float far, int64_t R, float* rays_out, void* stream) {
  if (R == 0) return GBN_OK;
  GBN_REQUIRE(rays_out, "pack_rays: null output");
  GBN_REQUIRE(R > 0, "pack_rays: negative ray count");
  GBN_REQUIRE((c2w != nullptr) != (rays_o != nullptr && rays_d != nullptr), "pack_rays: give a camera pose or a ray batch, not both");
  GBN_REQUIRE(c2w || (rays_o && rays_d), "pack_rays: rays_o and rays_d go together");
  GBN_REQUIRE(!c2w_static || (c2w && use_viewdirs), "pack_rays: a static camera needs a pose and use_viewdirs (run.py:1713)");
  GBN_REQUIRE(H > 0 && W > 0 && focal != 0, "pack_rays: bad intrinsics");
  RaySetupArgs a{};
  a.c2w = c2w; a.c2w_static = c2w_static; a.c2w_ld = c2w_ld; a.c2w_static_ld = c2w_static_ld;
  a.rays_o = rays_o; a.rays_d = rays_d; a.o_ld = o_ld; a.d_ld = d_ld; a.depths = depths;
  a.H = H; a.W = W; a.focal = (float)focal; a.half_w = (float)(W * .5); a.half_h = (float)(H * .5);
  if (c2w) {
    GBN_REQUIRE(c2w_ld >= 4 && (!c2w_static || c2w_static_ld >= 4), "pack_rays: a pose is [3,4] (pitch >= 4)");
    GBN_REQUIRE(patch_i >= 0 && patch_j >= 0 && patch_h > 0 && patch_w > 0 && patch_i + patch_h <= H && patch_j + patch_w <= W,
                "pack_rays: patch outside the frame");
    GBN_REQUIRE((int64_t)patch_h * patch_w == R, "pack_rays: R must equal the patch size");
  } else {
    GBN_REQUIRE(o_ld >= 3 && d_ld >= 3, "pack_rays: ray pitch < 3");
  }
  a.pi = patch_i; a.pj = patch_j; a.len2 = patch_w > 0 ? patch_w : W;
  a.use_viewdirs = use_viewdirs; a.ndc = ndc;
  a.ndc_ax = (float)(-1. / (W / (2. * focal))); a.ndc_ay = (float)(-1. / (H / (2. * focal)));
  a.near = near; a.far = far; a.R = R;
  a.row = 8 + (depths ? 1 : 0) + (use_viewdirs ? 3 : 0);
  a.out = rays_out;
  const int64_t blocks = (R + kRayThreads - 1) / kRayThreads;
  GBN_REQUIRE(blocks < (int64_t)1 << 31, "pack_rays: too many rays for one launch");
  ray_setup_kernel<<<(unsigned)blocks, kRayThreads, kRayThreads * a.row * sizeof(float), (cudaStream_t)stream>>>(a);
  return check_launch("ray_setup_kernel");
}
