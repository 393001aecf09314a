// Diagnostic: one tcgen05.mma with the A operand in TMEM (written with tcgen05.st, row == lane, two consecutive bf16
// K-elements per 32-bit column) against a K-major 128B-swizzled B tile in shared memory.  D[128,128] = A[128,64] B^T.
#include "common.cuh"   // compiled with -I gb-nerf_b200/csrc by csrc/build.py --exp / --diag
#include "tc_ptx.cuh"

namespace gbn {
using namespace tc;

// F16: fp16 operands and an fp16 accumulator (idesc c_format = a_format = b_format = 0); D then receives the RAW
// 32-bit TMEM cells of columns [0,128) so that the host can establish how 16-bit accumulators are laid out.
template <bool F16>
__global__ void __launch_bounds__(128, 1) ts_probe_kernel(const uint16_t* __restrict__ A, const uint8_t* __restrict__ Bimg,
                                                          float* __restrict__ D, int a_col, int col_per_kstep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + 16384, tptr = base + 16384 + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc(tptr, 512);
  for (int i = threadIdx.x; i < 16384 / 16; i += 128)
    reinterpret_cast<uint4*>(gen)[i] = reinterpret_cast<const uint4*>(Bimg)[i];
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 16384 + 8);
  const int row = threadIdx.x;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp << 5) << 16);
  uint32_t w[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) w[i] = (uint32_t)A[row * 64 + 2 * i] | ((uint32_t)A[row * 64 + 2 * i + 1] << 16);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(lane_addr + a_col),
      "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]),
      "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]), "r"(w[20]),
      "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]), "r"(w[29]), "r"(w[30]),
      "r"(w[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 0) {
    const uint64_t bdesc = smem_desc_sw128(base);
    const uint32_t idesc = F16 ? (((128u >> 3) << 17) | ((128u >> 4) << 24)) : make_idesc(1, 128, 128);
    if (elect_one()) {
      for (int k = 0; k < 4; ++k) {
        const uint32_t a_t = tmem + a_col + k * col_per_kstep;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
            "r"(a_t), "l"(bdesc + 2 * k), "r"(idesc), "r"(k == 0 ? 0u : 1u)
            : "memory");
      }
      umma_commit(bar);
    }
    __syncwarp();
  }
  while (!mbar_try_wait(bar, 0)) {}
  tc_fence_after_sync();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(lane_addr + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[row * 128 + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace gbn

extern "C" int gbn_debug_ts_mma(const void* A, const void* Bimg, float* D, int a_col, int col_per_kstep, void* stream) {
  using namespace gbn;
  GBN_REQUIRE(A && Bimg && D, "debug_ts_mma: null pointer");
  static bool attr = false;
  if (!attr) { GBN_CUDA(cudaFuncSetAttribute(ts_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480)); attr = true; }
  ts_probe_kernel<false><<<1, 128, 16384 + 64 + 1024, (cudaStream_t)stream>>>(static_cast<const uint16_t*>(A),
                                                                               static_cast<const uint8_t*>(Bimg), D, a_col, col_per_kstep);
  return check_launch("ts_probe_kernel");
}

extern "C" int gbn_debug_ts_mma_f16(const void* A, const void* Bimg, void* D_raw, int a_col, int col_per_kstep, void* stream) {
  using namespace gbn;
  GBN_REQUIRE(A && Bimg && D_raw, "debug_ts_mma_f16: null pointer");
  static bool attr = false;
  if (!attr) { GBN_CUDA(cudaFuncSetAttribute(ts_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480)); attr = true; }
  ts_probe_kernel<true><<<1, 128, 16384 + 64 + 1024, (cudaStream_t)stream>>>(static_cast<const uint16_t*>(A),
                                                                              static_cast<const uint8_t*>(Bimg),
                                                                              static_cast<float*>(D_raw), a_col, col_per_kstep);
  return check_launch("ts_probe_kernel<f16>");
}
