// Depth -> normal maps: depth2normal_geo, run.py:2458-2474, called every training iteration on the
// down-sampled render (run.py:1440-1443).  The reference unfolds a k x k (k = 31) zero-padded window of xyz points per
// pixel into a [k*k, 3] matrix A (11.5 KB per pixel in HBM), forms A^T A, inverts it and multiplies by A^T 1: the
// least-squares plane n.p = 1 through the window.  That is  n = M^-1 s  with  M = sum p p^T (3x3, symmetric) and
// s = sum p  over the window, i.e. a 9-channel box filter followed by a 3x3 solve per pixel.  Here: one kernel, a CTA
// stages the 9 moment channels of its (16 + k - 1)^2 halo in shared memory, sums separably (rows, then columns) and
// solves in double (the 3x3 systems of nearly planar patches are ill-conditioned in fp32).
// Backward (the normal-guidance loss needs d/d depth): with a = M^-1 g,  dL/ds = a,  dL/dM = -a n^T, so
//   dL/dp = (sum_q C_q) p + sum_q a_q,  C_q = -(a_q n_q^T + n_q a_q^T),  q over the window of p
// - the same 9-channel box filter over (a, C), then a 3x3 product per pixel.  M^-1 (6 floats) is kept by the forward.
#include "common.cuh"

namespace gbn {
namespace {

constexpr int kNT = 16;            // output tile edge
constexpr int kNMaxK = 31;
constexpr int kNThreads = kNT * kNT;

struct NormArgs {
  const float* points;   // [B,3,H,W]
  const float* normals;  // backward: [B,3,H,W]
  const float* minv;     // backward: [B,6,H,W] (xx, xy, xz, yy, yz, zz of M^-1)
  const float* g;        // backward: dL/dn [B,3,H,W]
  float* out;            // forward: normals; backward: dL/dpoints
  float* minv_out;       // forward: [B,6,H,W] or NULL
  int B, H, W, k;
};

// 9 channel values of source pixel (b, y, x): forward = moments of the point, backward = (a, C) of the pixel
template <bool BWD>
__device__ __forceinline__ void source9(const NormArgs& a, int b, int y, int x, float (&v)[9]) {
  const size_t hw = (size_t)a.H * a.W, at = (size_t)y * a.W + x;
  if (!BWD) {
    const float* p = a.points + (size_t)b * 3 * hw + at;
    const float px = __ldg(p), py = __ldg(p + hw), pz = __ldg(p + 2 * hw);
    v[0] = px * px; v[1] = px * py; v[2] = px * pz; v[3] = py * py; v[4] = py * pz; v[5] = pz * pz;
    v[6] = px; v[7] = py; v[8] = pz;
  } else {
    const float* mi = a.minv + (size_t)b * 6 * hw + at;
    const float* n = a.normals + (size_t)b * 3 * hw + at;
    const float* g = a.g + (size_t)b * 3 * hw + at;
    const float ixx = __ldg(mi), ixy = __ldg(mi + hw), ixz = __ldg(mi + 2 * hw), iyy = __ldg(mi + 3 * hw), iyz = __ldg(mi + 4 * hw),
                izz = __ldg(mi + 5 * hw);
    const float gx = __ldg(g), gy = __ldg(g + hw), gz = __ldg(g + 2 * hw);
    const float nx = __ldg(n), ny = __ldg(n + hw), nz = __ldg(n + 2 * hw);
    const float ax = ixx * gx + ixy * gy + ixz * gz, ay = ixy * gx + iyy * gy + iyz * gz, az = ixz * gx + iyz * gy + izz * gz;
    v[0] = -2.f * ax * nx; v[1] = -(ax * ny + ay * nx); v[2] = -(ax * nz + az * nx);
    v[3] = -2.f * ay * ny; v[4] = -(ay * nz + az * ny); v[5] = -2.f * az * nz;
    v[6] = ax; v[7] = ay; v[8] = az;
  }
}

template <bool BWD>
__global__ void __launch_bounds__(kNThreads) normals_kernel(const NormArgs a) {
  extern __shared__ float sm[];
  const int r = a.k >> 1, E = kNT + a.k - 1;          // halo radius, staged tile edge
  float* src = sm;                                      // [9][E][E]
  float* rows = sm + 9 * E * E;                         // [9][E][kNT]: horizontal sums
  const int b = blockIdx.z, y0 = blockIdx.y * kNT, x0 = blockIdx.x * kNT;
  for (int i = threadIdx.x; i < E * E; i += kNThreads) {
    const int ly = i / E, lx = i - ly * E, y = y0 + ly - r, x = x0 + lx - r;
    float v[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // zero padding (unfold's padding, run.py:2462)
    if (y >= 0 && y < a.H && x >= 0 && x < a.W) source9<BWD>(a, b, y, x, v);
#pragma unroll
    for (int c = 0; c < 9; ++c) src[(c * E + ly) * E + lx] = v[c];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < E * kNT; i += kNThreads) {   // row sums: E rows x 16 output columns
    const int ly = i / kNT, ox = i - ly * kNT;
    float s[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < a.k; ++t) {
#pragma unroll
      for (int c = 0; c < 9; ++c) s[c] += src[(c * E + ly) * E + ox + t];
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) rows[(c * E + ly) * kNT + ox] = s[c];
  }
  __syncthreads();
  const int oy = threadIdx.x / kNT, ox = threadIdx.x - oy * kNT, y = y0 + oy, x = x0 + ox;
  if (y >= a.H || x >= a.W) return;
  float s[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int t = 0; t < a.k; ++t) {
#pragma unroll
    for (int c = 0; c < 9; ++c) s[c] += rows[(c * E + oy + t) * kNT + ox];
  }
  const size_t hw = (size_t)a.H * a.W, at = (size_t)y * a.W + x;
  if (!BWD) {
    // n = M^-1 s by the adjugate, in double
    const double xx = s[0], xy = s[1], xz = s[2], yy = s[3], yz = s[4], zz = s[5];
    const double c00 = yy * zz - yz * yz, c01 = xz * yz - xy * zz, c02 = xy * yz - xz * yy;
    const double c11 = xx * zz - xz * xz, c12 = xy * xz - xx * yz, c22 = xx * yy - xy * xy;
    const double det = xx * c00 + xy * c01 + xz * c02, id = 1.0 / det;
    const double ixx = c00 * id, ixy = c01 * id, ixz = c02 * id, iyy = c11 * id, iyz = c12 * id, izz = c22 * id;
    float* o = a.out + (size_t)b * 3 * hw + at;
    o[0] = (float)(ixx * s[6] + ixy * s[7] + ixz * s[8]);
    o[hw] = (float)(ixy * s[6] + iyy * s[7] + iyz * s[8]);
    o[2 * hw] = (float)(ixz * s[6] + iyz * s[7] + izz * s[8]);
    if (a.minv_out) {
      float* m = a.minv_out + (size_t)b * 6 * hw + at;
      m[0] = (float)ixx; m[hw] = (float)ixy; m[2 * hw] = (float)ixz; m[3 * hw] = (float)iyy; m[4 * hw] = (float)iyz; m[5 * hw] = (float)izz;
    }
  } else {
    const float* p = a.points + (size_t)b * 3 * hw + at;
    const float px = __ldg(p), py = __ldg(p + hw), pz = __ldg(p + 2 * hw);
    float* o = a.out + (size_t)b * 3 * hw + at;
    o[0] = s[0] * px + s[1] * py + s[2] * pz + s[6];
    o[hw] = s[1] * px + s[3] * py + s[4] * pz + s[7];
    o[2 * hw] = s[2] * px + s[4] * py + s[5] * pz + s[8];
  }
}

template <bool BWD>
int launch_normals(const NormArgs& a, cudaStream_t st) {
  const int E = kNT + a.k - 1;
  const size_t smem = (size_t)(9 * E * E + 9 * E * kNT) * sizeof(float);
  static bool attr = false;   // immutable kernel attribute, set once per instantiation
  if (!attr) {
    const int Emax = kNT + kNMaxK - 1;
    GBN_CUDA(cudaFuncSetAttribute(normals_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)((9 * Emax * Emax + 9 * Emax * kNT) * sizeof(float))));
    attr = true;
  }
  dim3 grid((a.W + kNT - 1) / kNT, (a.H + kNT - 1) / kNT, a.B);
  normals_kernel<BWD><<<grid, kNThreads, smem, st>>>(a);
  return check_launch(BWD ? "normals_kernel<bwd>" : "normals_kernel<fwd>");
}

}  // namespace
}  // namespace gbn

using namespace gbn;

extern "C" int gbn_normals_forward(const float* points, int B, int H, int W, int k, float* normals, float* minv, void* stream) {
  if (B == 0 || H == 0 || W == 0) return GBN_OK;
  GBN_REQUIRE(points && normals, "normals_forward: null pointer");
  GBN_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535, "normals_forward: bad sizes");
  GBN_REQUIRE(k >= 1 && k <= kNMaxK && (k & 1), "normals_forward: window k must be odd and <= %d (got %d)", kNMaxK, k);
  NormArgs a{};
  a.points = points; a.out = normals; a.minv_out = minv; a.B = B; a.H = H; a.W = W; a.k = k;
  return launch_normals<false>(a, (cudaStream_t)stream);
}

extern "C" int gbn_normals_backward(const float* points, const float* normals, const float* minv, const float* g_normals, int B,
                                    int H, int W, int k, float* g_points, void* stream) {
  if (B == 0 || H == 0 || W == 0) return GBN_OK;
  GBN_REQUIRE(points && normals && minv && g_normals && g_points, "normals_backward: null pointer");
  GBN_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535, "normals_backward: bad sizes");
  GBN_REQUIRE(k >= 1 && k <= kNMaxK && (k & 1), "normals_backward: window k must be odd and <= %d (got %d)", kNMaxK, k);
  NormArgs a{};
  a.points = points; a.normals = normals; a.minv = minv; a.g = g_normals; a.out = g_points; a.B = B; a.H = H; a.W = W; a.k = k;
  return launch_normals<true>(a, (cudaStream_t)stream);
}
