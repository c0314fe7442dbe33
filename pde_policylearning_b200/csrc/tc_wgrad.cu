// 1x1-convolution weight gradient on tcgen05:  dW[o,i] = sum_{b,p} G[b,o,p] X[b,i,p],  db[o] = sum_{b,p} G[b,o,p]
// (backward of fno_block.py:131 skips, tfno.py:11-38 lifting/projection, rno.py:224-228 / pinobserver.py:223 convs).
//
// The contraction runs over pixels, and both tensors are pixel-contiguous, so both operands are K-major exactly
// as they sit in HBM: TMA boxes of [32 px x channels] with the 128-byte swizzle are valid UMMA tiles with no
// transposition.  A = G tile (M = output channels, padded to 128 rows by reading on into the following shared
// memory -- those accumulator rows are never read), B = X tile plus a constant block of ones rows (so the same
// MMAs also produce db), D = [Co x (Ci + 16)] fp32 in TMEM, accumulated over every tile a CTA owns and written
// once at the end as a per-CTA partial; k_pw_wgrad_reduce (pointwise.cu) sums the partials deterministically.
// 3xTF32: raw tiles are the hi parts (the tensor core ignores the low 13 mantissa bits), lo tiles are produced
// elementwise in shared memory by the converter warps.
#include <string.h>

#include "common.cuh"
#include "tc.cuh"

using namespace tc;

namespace {

constexpr int kThreadsWg = 192;  // warp 0 TMA, warp 1 MMA, warps 2-5 converter + final read-out

struct WgTc {
  int Co, Cop, Ci, Cip, TP, nb, S, mblocks;
  int tiles_per_img;
  long tiles, tiles_per_cta;
  float* partial;
};

struct WgLayout { uint32_t gbytes, xbytes, glo, x, xlo, stage_bytes, bars, total; };

__host__ __device__ inline WgLayout wg_layout(const WgTc& p) {
  WgLayout L;
  L.gbytes = (uint32_t)p.Cop * p.TP * 4;
  L.xbytes = (uint32_t)(p.Cip + 16) * p.TP * 4;
  L.glo = L.gbytes; L.x = 2 * L.gbytes; L.xlo = L.x + L.xbytes;
  L.stage_bytes = 2 * (L.gbytes + L.xbytes);
  L.bars = L.stage_bytes * p.S + 16384;          // 16 KB tail: the M=128 A reads run past short G tiles
  L.total = L.bars + 8 * (3 * p.S + 1) + 16 + 1024;
  return L;
}

__global__ void __launch_bounds__(kThreadsWg, 1)
k_wgrad_tc(const __grid_constant__ CUtensorMap tmg, const __grid_constant__ CUtensorMap tmx, const WgTc p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const WgLayout L = wg_layout(p);
  uint64_t* full = (uint64_t*)(smem + L.bars);
  uint64_t* cvt = full + p.S;
  uint64_t* empty = cvt + p.S;
  uint64_t* done = empty + p.S;
  uint32_t* tslot = (uint32_t*)(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NW = p.Cip + 16;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(p.mblocks * NW)) ncols <<= 1;

  // zero everything once (tail, ones blocks), then write the ones rows: row Cip of every X box, hi copy only
  for (uint32_t i = tid; i < L.bars / 16; i += kThreadsWg) ((float4*)smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int i = tid; i < p.S * p.nb * 32; i += kThreadsWg) {
    const int s = i / (p.nb * 32), r = i - s * (p.nb * 32);
    const int j = r >> 5, e = r & 31;
    ((float*)(smem + (size_t)s * L.stage_bytes + L.x + (size_t)j * NW * 128 + (size_t)p.Cip * 128))[e] = 1.0f;
  }
  if (tid == 0) {
    for (int s = 0; s < p.S; s++) { mbar_init(&full[s], 1); mbar_init(&cvt[s], 128); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tslot, ncols);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmg); tma_prefetch_desc(&tmx); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tslot;
  const long t_first = (long)blockIdx.x * p.tiles_per_cta;
  const long t_end = t_first + p.tiles_per_cta < p.tiles ? t_first + p.tiles_per_cta : p.tiles;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(p.Cop + p.Cip) * p.TP * 4;
      int it = 0;
      for (long tile = t_first; tile < t_end; tile++, it++) {
        const int s = it % p.S;
        const uint32_t ph = (uint32_t)(it / p.S) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], bytes);
        uint8_t* st = smem + (size_t)s * L.stage_bytes;
        const int b = (int)(tile / p.tiles_per_img);
        const int p0 = (int)(tile - (long)b * p.tiles_per_img) * p.TP;
        for (int j = 0; j < p.nb; j++) {
          tma_load_3d(st + (size_t)j * p.Cop * 128, &tmg, &full[s], p0 + 32 * j, 0, b);
          tma_load_3d(st + L.x + (size_t)j * NW * 128, &tmx, &full[s], p0 + 32 * j, 0, b);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_tf32(128, NW, 0, 0);
    const uint32_t sbase = smem_u32(smem);
    int it = 0;
    uint32_t acc = 0;
    for (long tile = t_first; tile < t_end; tile++, it++) {
      const int s = it % p.S;
      const uint32_t ph = (uint32_t)(it / p.S) & 1u;
      mbar_wait(&full[s], ph);
      mbar_wait(&cvt[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = sbase + (uint32_t)s * L.stage_bytes;
        for (int mb = 0; mb < p.mblocks; mb++) {
          const uint32_t d = tbase + (uint32_t)(mb * NW);
          uint32_t a2 = acc;
          for (int pass = 0; pass < 3; pass++) {
            const uint32_t ga = st + (pass == 1 ? L.glo : 0) + (uint32_t)mb * 16384;
            const uint32_t xa = st + (pass == 2 ? L.xlo : L.x);
            for (int j = 0; j < p.nb; j++) {
              const uint64_t dg = smem_desc(ga + (uint32_t)j * p.Cop * 128, 16, 1024, LAYOUT_SW128);
              const uint64_t dx = smem_desc(xa + (uint32_t)j * NW * 128, 16, 1024, LAYOUT_SW128);
#pragma unroll
              for (int ks = 0; ks < 4; ks++) {
                mma_tf32_ss(d, dg + (uint64_t)(ks * 2), dx + (uint64_t)(ks * 2), idesc, a2);
                a2 = 1;
              }
            }
          }
        }
        mma_commit(&empty[s]);
      }
      acc = 1;
      __syncwarp();
    }
    if (elect_one()) mma_commit(done);
    __syncwarp();
  } else {
    // converter: lo tiles, elementwise (layout-agnostic); ones rows give lo = 0
    const int ct = tid - 64;
    int it = 0;
    for (long tile = t_first; tile < t_end; tile++, it++) {
      const int s = it % p.S;
      const uint32_t ph = (uint32_t)(it / p.S) & 1u;
      mbar_wait(&full[s], ph);
      uint8_t* st = smem + (size_t)s * L.stage_bytes;
      const float4* src = (const float4*)st;
      float4* dst = (float4*)(st + L.glo);
      for (uint32_t i = ct; i < L.gbytes / 16; i += 128) {
        const float4 x = src[i];
        dst[i] = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
      }
      src = (const float4*)(st + L.x);
      dst = (float4*)(st + L.xlo);
      for (uint32_t i = ct; i < L.xbytes / 16; i += 128) {
        const float4 x = src[i];
        dst[i] = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
      }
      fence_proxy_async();
      mbar_arrive(&cvt[s]);
    }
    // read-out: thread = accumulator row
    mbar_wait(done, 0);
    tc_fence_after();
    const int quad = warp & 3;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    float* pout = p.partial + (size_t)blockIdx.x * ((size_t)p.Co * p.Ci + p.Co);
    const bool any = t_end > t_first;
    for (int mb = 0; mb < p.mblocks; mb++) {
      const int o = mb * 128 + quad * 32 + lane;
      for (int c0 = 0; c0 < NW; c0 += 16) {
        float v[16];
        tmem_ld16(tbase + lane_base + (uint32_t)(mb * NW + c0), v);
        tmem_ld_wait();
        if (o < p.Co) {
#pragma unroll
          for (int j = 0; j < 16; j++) {
            const int c = c0 + j;
            const float val = any ? v[j] : 0.f;
            if (c < p.Ci) pout[(size_t)o * p.Ci + c] = val;
            else if (c == p.Cip) pout[(size_t)p.Co * p.Ci + o] = val;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, ncols);
}

}  // namespace

void b2no_tc_count_launch();

// Fills `partial` with *nblk per-CTA partials laid out [Co*Ci | Co]; returns 0 on success, 1 if not eligible.
int b2no_tc_wgrad(const float* g, const float* x, float* partial, int max_blocks, int batch, int ci, int co, long pixels,
                  int* nblk, cudaStream_t st) {
  if (!b2no_tc_available()) return 1;
  if (pixels % 128 != 0 || ci > 240 || co > 512 || ci < 1 || co < 1) return 1;
  if (((uintptr_t)g | (uintptr_t)x) & 15) return 1;
  WgTc p;
  memset(&p, 0, sizeof(p));
  p.Co = co; p.Cop = b2no_round_up(co, 8); p.Ci = ci; p.Cip = b2no_round_up(ci, 16);
  p.mblocks = (p.Cop + 127) / 128;
  if (p.mblocks * (p.Cip + 16) > 512) return 1;
  int dev = 0, max_smem = 0;
  B2NO_CHECK_CUDA(cudaGetDevice(&dev));
  B2NO_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  WgLayout L;
  bool ok = false;
  for (int min_s = 3; min_s >= 2 && !ok; min_s--) {       // prefer a pipeline at least 3 deep
    for (p.TP = 128; p.TP >= 32 && !ok; p.TP >>= 1) {
      for (p.S = 4; p.S >= min_s; p.S--) {
        L = wg_layout(p);
        if ((int)L.total <= max_smem) { ok = true; break; }
      }
      if (ok) break;
    }
  }
  if (!ok) return 1;
  p.nb = p.TP / 32;
  p.tiles_per_img = (int)(pixels / p.TP);
  p.tiles = (long)batch * p.tiles_per_img;
  long grid = b2no_sm_count();
  if (grid > max_blocks) grid = max_blocks;
  if (grid > p.tiles) grid = p.tiles;
  p.tiles_per_cta = (p.tiles + grid - 1) / grid;
  grid = (p.tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  p.partial = partial;
  CUtensorMap tmg, tmx;
  {
    uint64_t dims[3] = {(uint64_t)pixels, (uint64_t)co, (uint64_t)batch};
    uint64_t str[3] = {4, (uint64_t)pixels * 4, (uint64_t)pixels * 4 * co};
    uint32_t box[3] = {32, (uint32_t)p.Cop, 1};
    if (make_tmap_f32(&tmg, g, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    uint64_t dims2[3] = {(uint64_t)pixels, (uint64_t)ci, (uint64_t)batch};
    uint64_t str2[3] = {4, (uint64_t)pixels * 4, (uint64_t)pixels * 4 * ci};
    uint32_t box2[3] = {32, (uint32_t)p.Cip, 1};
    if (make_tmap_f32(&tmx, x, 3, dims2, str2, box2, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  B2NO_CHECK_CUDA(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  k_wgrad_tc<<<(unsigned)grid, kThreadsWg, L.total, st>>>(tmg, tmx, p);
  B2NO_LAUNCH_CHECK();
  b2no_tc_count_launch();
  *nblk = (int)grid;
  return 0;
}
