// Stand-alone probe of the tcgen05 building blocks the b2no kernels rely on (descriptor layouts, operand
// majors, TMEM A operand, tf32 truncation, 3xTF32 accuracy).  Build: see tools/build_probe.sh; run on a B200.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../pde_policylearning_b200/csrc/tc.cuh"

#define CK(x)                                                                           \
  do {                                                                                  \
    cudaError_t e = (x);                                                                \
    if (e != cudaSuccess) {                                                             \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);    \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

using namespace tc;

// -------------------------------------------------------------------------------------------------
// T1: SS, A [128 x K] and B [N x K] both K-major no-swizzle.  D = A B^T
// -------------------------------------------------------------------------------------------------
template <int N, int K>
__global__ void __launch_bounds__(128) k_t1(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 4;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i % K;
    *(float*)(sA + kmajor_off(r, k, K)) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    *(float*)(sB + kmajor_off(r, k, K)) = B[i];
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(&tslot, 32 > N ? 32 : N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(128, N, 0, 0);
    const uint32_t sbo = (K / 4) * 128;
    for (int k = 0; k < K / 8; k++) {
      const uint64_t da = smem_desc(smem_u32(sA) + k * 256, 128, sbo, LAYOUT_NONE);
      const uint64_t db = smem_desc(smem_u32(sB) + k * 256, 128, sbo, LAYOUT_NONE);
      mma_tf32_ss(tbase, da, db, idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[16];
  for (int c = 0; c < N; c += 16) {
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; j++) D[(size_t)tid * N + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 32 > N ? 32 : N);
}

// -------------------------------------------------------------------------------------------------
// T2 / T5: A = X^T, X [C x 128] pixel-contiguous, loaded by TMA with 128B swizzle as an MN-major
// operand; B = W [N x C] K-major no-swizzle.  D[p][o] = sum_c X[c][p] W[o][c].
// mode 0: 1xTF32 (raw X, raw W);  mode 1: 3xTF32 (hi/lo)
// -------------------------------------------------------------------------------------------------
template <int N, int C>
__global__ void __launch_bounds__(128) k_t2(const __grid_constant__ CUtensorMap tmx, const float* __restrict__ W,
                                            float* __restrict__ D, int mode, float* __restrict__ dump) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;                        // 4 boxes of [C][32 px] = C*128 B each
  uint8_t* sXlo = sX + 4 * C * 128;
  uint8_t* sWh = sXlo + 4 * C * 128;
  uint8_t* sWl = sWh + N * C * 4;
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < N * C; i += 128) {
    const int r = i / C, k = i % C;
    const float w = W[i];
    const float hi = mode ? tf32_rna(w) : w;
    *(float*)(sWh + kmajor_off(r, k, C)) = hi;
    *(float*)(sWl + kmajor_off(r, k, C)) = tf32_rna(w - hi);
  }
  if (tid == 0) { mbar_init(&bar_tma, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tslot, 32 > N ? 32 : N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_tma, 4 * C * 128);
    for (int j = 0; j < 4; j++) tma_load_3d(sX + j * C * 128, &tmx, &bar_tma, 32 * j, 0, 0);
  }
  mbar_wait(&bar_tma, 0);
  if (dump) for (int i = tid; i < C * 128; i += 128) dump[i] = ((const float*)sX)[i];
  // lo tile: elementwise, same (swizzled) positions
  for (int i = tid; i < C * 128; i += 128) {
    const float x = ((const float*)sX)[i];
    ((float*)sXlo)[i] = tf32_lo(x);
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(128, N, 1, 0);
    const uint32_t sbo_b = (C / 4) * 128;
    int first = 1;
    for (int pass = 0; pass < (mode ? 3 : 1); pass++) {
      const uint8_t* xa = pass == 1 ? sXlo : sX;
      const uint8_t* wb = pass == 2 ? sWl : sWh;
      for (int k = 0; k < C / 8; k++) {
        // MN-major SW128_32B: atom = 4 channel rows x 128 B (512 B apart = SBO); MN groups of 32 px are C*128 B apart (LBO)
        const uint64_t da = smem_desc(smem_u32(xa) + k * 1024, C * 128, 512, LAYOUT_SW128_32B);
        const uint64_t db = smem_desc(smem_u32(wb) + k * 256, 128, sbo_b, LAYOUT_NONE);
        mma_tf32_ss(tbase, da, db, idesc, first ? 0 : 1);
        first = 0;
      }
    }
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  float v[16];
  for (int c = 0; c < N; c += 16) {
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; j++) D[(size_t)tid * N + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 32 > N ? 32 : N);
}

// -------------------------------------------------------------------------------------------------
// T3: TS.  A [128 x K] in TMEM (written with tcgen05.st, thread = row), B [N x K] K-major no-swizzle
// T4: same A, but B in the K-major 128B-swizzle layout written by threads (bsw = 1)
// -------------------------------------------------------------------------------------------------
template <int N, int K>
__global__ void __launch_bounds__(128) k_t3(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int bsw) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    uint32_t off;
    if (bsw) off = (k / 32) * (N * 128) + n * 128 + ((((k % 32) / 4) ^ (n & 7)) * 16) + (k & 3) * 4;
    else off = kmajor_off(n, k, K);
    *(float*)(sB + off) = B[i];
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(&tslot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  // A -> TMEM columns [32, 32+K)
  for (int c = 0; c < K; c += 16) {
    float v[16];
    for (int j = 0; j < 16; j++) v[j] = A[(size_t)tid * K + c + j];
    tmem_st16(tbase + lane_base + 32 + c, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(128, N, 0, 0);
    for (int k = 0; k < K / 8; k++) {
      uint64_t db;
      if (bsw) db = smem_desc(smem_u32(sB) + (k / 4) * (N * 128) + (k % 4) * 32, 16, 1024, LAYOUT_SW128);
      else db = smem_desc(smem_u32(sB) + k * 256, 128, (K / 4) * 128, LAYOUT_NONE);
      mma_tf32_ts(tbase, tbase + 32 + 8 * k, db, idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[16];
  for (int c = 0; c < N; c += 16) {
    tmem_ld16(tbase + lane_base + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; j++) D[(size_t)tid * N + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

// -------------------------------------------------------------------------------------------------
// T6 / T7: MN-major NO-swizzle ("interleave") operands.  Element (mn, k) of an operand lives at
//     (mn / 4) * mn_stride + (k / 8) * k_stride + (k % 8) * 16 + (mn % 4) * 4            (bytes)
// i.e. 16-byte groups of 4 MN-contiguous elements, 8 consecutive k at a 16-byte pitch.  `swap` selects which of the
// two strides goes into the descriptor's LBO field.  amn / bmn: operand is MN-major (else K-major no-swizzle).
// -------------------------------------------------------------------------------------------------
template <int N, int K>
__global__ void __launch_bounds__(128) k_t6(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                            int amn, int bmn, int swap, int a_mn_stride, int a_k_stride, int b_mn_stride,
                                            int b_k_stride) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sA = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = sA + 128 * K * 4;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i % K;
    const uint32_t off = amn ? (uint32_t)((r / 4) * a_mn_stride + (k / 8) * a_k_stride + (k % 8) * 16 + (r % 4) * 4) : kmajor_off(r, k, K);
    *(float*)(sA + off) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    const uint32_t off = bmn ? (uint32_t)((r / 4) * b_mn_stride + (k / 8) * b_k_stride + (k % 8) * 16 + (r % 4) * 4) : kmajor_off(r, k, K);
    *(float*)(sB + off) = B[i];
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(&tslot, 32 > N ? 32 : N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(128, N, amn, bmn);
    const uint32_t sbo = (K / 4) * 128;
    for (int k = 0; k < K / 8; k++) {
      uint64_t da, db;
      if (amn) da = swap ? smem_desc(smem_u32(sA) + k * a_k_stride, a_mn_stride, a_k_stride, LAYOUT_NONE)
                         : smem_desc(smem_u32(sA) + k * a_k_stride, a_k_stride, a_mn_stride, LAYOUT_NONE);
      else da = smem_desc(smem_u32(sA) + k * 256, 128, sbo, LAYOUT_NONE);
      if (bmn) db = swap ? smem_desc(smem_u32(sB) + k * b_k_stride, b_mn_stride, b_k_stride, LAYOUT_NONE)
                         : smem_desc(smem_u32(sB) + k * b_k_stride, b_k_stride, b_mn_stride, LAYOUT_NONE);
      else db = smem_desc(smem_u32(sB) + k * 256, 128, sbo, LAYOUT_NONE);
      mma_tf32_ss(tbase, da, db, idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[16];
  for (int c = 0; c < N; c += 16) {
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; j++) D[(size_t)tid * N + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 32 > N ? 32 : N);
}

// -------------------------------------------------------------------------------------------------
// T7: B = X^T as an MN-major operand (the transposed head backward needs it): X [C x NPX] pixel-contiguous, loaded by
// TMA with the 128B / 32B-atom swizzle; A = W [128 x C] K-major no-swizzle.  D[m][p] = sum_c W[m][c] X[c][p], 3xTF32.
// -------------------------------------------------------------------------------------------------
template <int NPX, int C>
__global__ void __launch_bounds__(128) k_t7(const __grid_constant__ CUtensorMap tmx, const float* __restrict__ W,
                                            float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int NB = NPX / 32;
  uint8_t* sX = smem;                        // NB boxes of [C][32 px] = C*128 B each
  uint8_t* sXlo = sX + NB * C * 128;
  uint8_t* sWh = sXlo + NB * C * 128;
  uint8_t* sWl = sWh + 128 * C * 4;
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * C; i += 128) {
    const int r = i / C, k = i % C;
    const float w = W[i];
    const float hi = tf32_rna(w);
    *(float*)(sWh + kmajor_off(r, k, C)) = hi;
    *(float*)(sWl + kmajor_off(r, k, C)) = tf32_rna(w - hi);
  }
  if (tid == 0) { mbar_init(&bar_tma, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tslot, NPX < 32 ? 32 : NPX);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_tma, NB * C * 128);
    for (int j = 0; j < NB; j++) tma_load_3d(sX + j * C * 128, &tmx, &bar_tma, 32 * j, 0, 0);
  }
  mbar_wait(&bar_tma, 0);
  for (int i = tid; i < NB * C * 32; i += 128) {
    const float x = ((const float*)sX)[i];
    ((float*)sXlo)[i] = tf32_lo(x);
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(128, NPX, 0, 1);
    const uint32_t sbo_a = (C / 4) * 128;
    int first = 1;
    for (int pass = 0; pass < 3; pass++) {
      const uint8_t* wa = pass == 1 ? sWl : sWh;
      const uint8_t* xb = pass == 2 ? sXlo : sX;
      for (int k = 0; k < C / 8; k++) {
        const uint64_t da = smem_desc(smem_u32(wa) + k * 256, 128, sbo_a, LAYOUT_NONE);
        // MN-major SW128_32B: 8 channel rows per K step (1024 B); groups of 32 px are C*128 B apart (LBO); SBO = 512 as for the A side
        const uint64_t db = smem_desc(smem_u32(xb) + k * 1024, C * 128, 512, LAYOUT_SW128_32B);
        mma_tf32_ss(tbase, da, db, idesc, first ? 0 : 1);
        first = 0;
      }
    }
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  float v[16];
  for (int c = 0; c < NPX; c += 16) {
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; j++) D[(size_t)tid * NPX + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, NPX < 32 ? 32 : NPX);
}

// -------------------------------------------------------------------------------------------------
static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rna_tf32(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x;
}
static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

struct Err { double trunc, rna, exact; };
// D[m][n] = sum_k A(m,k) B(n,k)
template <class FA, class FB>
static Err compare(const std::vector<float>& D, int M, int N, int K, FA a, FB b) {
  Err e{0, 0, 0};
  double nt = 0, nr = 0, ne = 0, den = 0;
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N; n++) {
      double st = 0, sr = 0, se = 0;
      for (int k = 0; k < K; k++) {
        st += (double)trunc_tf32(a(m, k)) * trunc_tf32(b(n, k));
        sr += (double)rna_tf32(a(m, k)) * rna_tf32(b(n, k));
        se += (double)a(m, k) * b(n, k);
      }
      const double d = D[(size_t)m * N + n];
      nt += (d - st) * (d - st); nr += (d - sr) * (d - sr); ne += (d - se) * (d - se); den += se * se;
    }
  e.trunc = sqrt(nt / den); e.rna = sqrt(nr / den); e.exact = sqrt(ne / den);
  return e;
}

int main() {
  srand(1);
  // ---------------- T1 ----------------
  {
    constexpr int N = 32, K = 32;
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    for (auto& v : A) v = frand();
    for (auto& v : B) v = frand();
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size() * 4));
    const int smem = (128 + N) * K * 4;
    k_t1<N, K><<<1, 128, smem>>>(dA, dB, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    Err e = compare(D, 128, N, K, [&](int m, int k) { return A[m * K + k]; }, [&](int n, int k) { return B[n * K + k]; });
    printf("T1 SS K-major/no-swizzle   N=%d K=%d : rel err vs trunc %.3e  vs rna %.3e  vs exact %.3e\n", N, K, e.trunc, e.rna, e.exact);
  }
  // ---------------- T2 / T5 ----------------
  {
    constexpr int N = 32, C = 32, P = 128;
    std::vector<float> X(C * P), W(N * C), D(128 * N);
    for (auto& v : X) v = frand();
    for (auto& v : W) v = frand();
    float *dX, *dW, *dD;
    CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap tm;
    uint64_t dims[3] = {P, C, 1}, str[3] = {4, (uint64_t)P * 4, (uint64_t)P * C * 4};
    uint32_t box[3] = {32, C, 1};
    int rc = make_tmap_f32(&tm, dX, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) { printf("tensor map failed %d\n", rc); return 1; }
    const int smem = 8 * C * 128 + 2 * N * C * 4 + 1024;
    CK(cudaFuncSetAttribute(k_t2<N, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    float* dDump; CK(cudaMalloc(&dDump, C * 128 * 4));
    for (int mode = 0; mode < 2; mode++) {
      CK(cudaMemset(dD, 0, D.size() * 4));
      k_t2<N, C><<<1, 128, smem>>>(tm, dW, dD, mode, dDump);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      Err e = compare(D, 128, N, C, [&](int m, int k) { return X[k * P + m]; }, [&](int n, int k) { return W[n * C + k]; });
      if (mode == 0) {
        std::vector<float> T(C * 128);
        CK(cudaMemcpy(T.data(), dDump, T.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int c = 0; c < C; c++) for (int p = 0; p < P; p++) {
          const int off = (p / 32) * C * 32 + c * 32 + ((((p % 32) / 8) ^ (c & 3)) * 8) + (p % 8);
          if (T[off] != X[c * P + p]) bad++;
        }
        printf("TMA 128B_ATOM_32B tile layout check: %d mismatches of %d\n", bad, C * P);
      }
      printf("T%d TMA SW128 MN-major A, %s : rel err vs trunc %.3e  vs rna %.3e  vs exact %.3e\n", mode ? 5 : 2,
             mode ? "3xTF32" : "1xTF32", e.trunc, e.rna, e.exact);
    }
  }
  // ---------------- T3 / T4 ----------------
  {
    constexpr int N = 16, K = 128;
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    for (auto& v : A) v = frand();
    for (auto& v : B) v = frand();
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    const int smem = N * K * 4 + 1024;
    for (int bsw = 0; bsw < 2; bsw++) {
      CK(cudaMemset(dD, 0, D.size() * 4));
      k_t3<N, K><<<1, 128, smem>>>(dA, dB, dD, bsw);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      Err e = compare(D, 128, N, K, [&](int m, int k) { return A[m * K + k]; }, [&](int n, int k) { return B[n * K + k]; });
      printf("T%d TS (A in TMEM), B %s : rel err vs trunc %.3e  vs rna %.3e  vs exact %.3e\n", bsw ? 4 : 3,
             bsw ? "K-major SW128 " : "K-major no-swz", e.trunc, e.rna, e.exact);
    }
  }


  // ---------------- T7: MN-major B operand from a TMA 128B_ATOM_32B tile ----------------
  {
    constexpr int NPX = 64, C = 32;
    std::vector<float> X(C * NPX), W(128 * C), D(128 * NPX);
    for (auto& v : X) v = frand();
    for (auto& v : W) v = frand();
    float *dX, *dW, *dD;
    CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size() * 4));
    CUtensorMap tm;
    uint64_t dims[3] = {NPX, C, 1}, str[3] = {4, (uint64_t)NPX * 4, (uint64_t)NPX * C * 4};
    uint32_t box[3] = {32, C, 1};
    int rc = make_tmap_f32(&tm, dX, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) { printf("tensor map failed %d\n", rc); return 1; }
    const int smem = 2 * (NPX / 32) * C * 128 + 2 * 128 * C * 4 + 1024;
    CK(cudaFuncSetAttribute(k_t7<NPX, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_t7<NPX, C><<<1, 128, smem>>>(tm, dW, dD);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    Err e = compare(D, 128, NPX, C, [&](int m, int k) { return W[m * C + k]; }, [&](int n, int k) { return X[k * NPX + n]; });
    printf("T7 MN-major B (TMA SW128_32B), 3xTF32 : rel err vs exact %.3e  (expect ~4e-7; ~1 means the layout is not accepted)\n", e.exact);
  }
  // ---------------- T6 / T7: MN-major no-swizzle operands ----------------
  {
    constexpr int N = 32, K = 128;
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    for (auto& v : A) v = frand();
    for (auto& v : B) v = frand();
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    const int smem = (128 + N) * K * 4 + 1024;
    CK(cudaFuncSetAttribute(k_t6<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    struct Case { const char* name; int amn, bmn, a_mn, a_k, b_mn, b_k; };
    // a: F-tile style (k groups dense at 128 B, MN groups at (K/8)*128);  b: K-major image reinterpreted (MN groups at 128 B, k groups at (MN/4)*128)
    const Case cases[] = {
        {"sanity: both K-major ", 0, 0, 0, 0, 0, 0},
        {"A MN-major (k-dense) ", 1, 0, (K / 8) * 128, 128, 0, 0},
        {"A MN-major (mn-dense)", 1, 0, 128, (128 / 4) * 128, 0, 0},
        {"B MN-major (k-dense) ", 0, 1, 0, 0, (K / 8) * 128, 128},
        {"B MN-major (mn-dense)", 0, 1, 0, 0, 128, (N / 4) * 128},
        {"A+B MN-major         ", 1, 1, (K / 8) * 128, 128, 128, (N / 4) * 128},
    };
    for (const Case& c : cases)
      for (int swap = 0; swap < 2; swap++) {
        CK(cudaMemset(dD, 0, D.size() * 4));
        k_t6<N, K><<<1, 128, smem>>>(dA, dB, dD, c.amn, c.bmn, swap, c.a_mn, c.a_k, c.b_mn, c.b_k);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        Err e = compare(D, 128, N, K, [&](int m, int k) { return A[m * K + k]; }, [&](int n, int k) { return B[n * K + k]; });
        printf("   D[0..3] = %g %g %g %g   D[5*N+7] = %g\n", D[0], D[1], D[2], D[3], D[5 * N + 7]);
        printf("T6 %s no-swizzle, %s : rel err vs trunc %.3e  vs exact %.3e\n", c.name,
               swap ? "LBO = MN-group stride, SBO = k-group stride" : "LBO = k-group stride, SBO = MN-group stride", e.trunc, e.exact);
      }
  }
  printf("probe done\n");
  return 0;
}
