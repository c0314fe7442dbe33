// Measured tensor-core peaks for the roofline denominators (SURVEY.md 8d asks for a measured kind::tf32 figure instead of
// "bf16 / 2").  One CTA per SM issues back-to-back dense tcgen05.mma instructions (M = 128, N = 256, SS form, both
// operands resident in shared memory, two alternating TMEM accumulators) with nothing else going on: the rate of the
// tensor pipe alone for cta_group::1.  The data is irrelevant (zeros), only the issue rate is timed -- by the caller,
// with CUDA events around the launch.
//   kind 0: kind::tf32 (K = 8 per instruction)      kind 1: kind::f16 with bf16 operands (K = 16 per instruction)
#include "common.cuh"
#include "tc.cuh"

using namespace tc;

namespace {

constexpr int kPeakN = 256;

template <int KIND>
__global__ void __launch_bounds__(128, 1) k_tc_peak(int n_mma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  // one K step of operands: A [128 rows x 32 B], B [256 rows x 32 B], K-major no-swizzle core matrices (8 rows x 16 B)
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * 32;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + kPeakN) * 32 / 4; i += 128) ((uint32_t*)smem)[i] = 0u;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(&tslot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (warp == 0) {
    // K-major, no swizzle: LBO = 128 B between the two 16-byte K chunks of a step, SBO = 256 B between 8-row groups
    const uint64_t da = smem_desc(smem_u32(sA), 128, 256, LAYOUT_NONE);
    const uint64_t db = smem_desc(smem_u32(sB), 128, 256, LAYOUT_NONE);
    const uint32_t idesc = KIND == 0 ? idesc_tf32(128, kPeakN, 0, 0) : idesc_bf16(128, kPeakN, 0, 0);
    if (elect_one()) {
      for (int i = 0; i < n_mma; i++) {
        const uint32_t d = tbase + (uint32_t)((i & 1) * kPeakN);
        if (KIND == 0) {
          mma_tf32_ss(d, da, db, idesc, i > 1 ? 1u : 0u);
        } else {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
              ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(i > 1 ? 1u : 0u)
              : "memory");
        }
      }
      mma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace

// Launches the probe on every SM; *flops receives the floating-point operations the launch performs (2 M N K per MMA).
extern "C" int b2no_tc_peak_probe(int kind, int n_mma, double* flops, void* stream) {
  if (n_mma < 2 || (kind != 0 && kind != 1)) return B2NO_E_ARG;
  if (!b2no_tc_available()) return B2NO_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = b2no_sm_count();
  const int smem = (128 + kPeakN) * 32 + 1024;
  if (kind == 0) k_tc_peak<0><<<grid, 128, smem, st>>>(n_mma);
  else k_tc_peak<1><<<grid, 128, smem, st>>>(n_mma);
  B2NO_LAUNCH_CHECK();
  if (flops) *flops = 2.0 * 128.0 * kPeakN * (kind == 0 ? 8.0 : 16.0) * (double)n_mma * (double)grid;
  return 0;
}
