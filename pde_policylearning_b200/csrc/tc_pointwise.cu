// Fused pixel-tile kernel on the 5th-gen tensor cores (tcgen05 / TMEM), fed by TMA.
//
//   out[b, o, p] = act( sum_i W1[o,i] X1[b,i,p] + sum_i W2[o,i] X2[b,i,p]            1x1 convolutions
//                       + sum_q T[p, q] A'[b, row(p), q, o]                            last inverse-DFT stage
//                       + bias[o] + add[b,o,p] ) * mul[b,o,p] * dact'(dz[b,o,p])
//
// One tile = 128 consecutive pixels of one sample = the M dimension of every MMA (TMEM lane = pixel), the
// output channels are N (TMEM columns), the input channels / spectral index q are K.  All three sums
// accumulate into the SAME TMEM tile, so the FNO layer (spectral_convolution.py:342-345 + fno_block.py:131-150),
// the RNO FourierLayer2d (rno.py:224-228) and their dx adjoints are one pass over HBM: read x once, write y once.
//
// fp32 parity on tf32 tensor cores: every product is issued three times (3xTF32): x_hi*w_hi + x_lo*w_hi +
// x_hi*w_lo, where x_hi is the raw fp32 tile exactly as TMA delivered it (the tensor core ignores the low 13
// mantissa bits -- measured, tools/tc_probe.cu) and x_lo = x - trunc(x) is produced by an elementwise pass over
// the tile in shared memory.  Measured error of the scheme: 4e-7 relative (probe T5).
//
// Warp roles (448 threads, persistent CTAs, static round-robin tile schedule):
//   warp 0      TMA producer: X tiles as MN-major operands (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B boxes of
//               32 px x C rows), A' rows with cp.async.bulk
//   warp 1      MMA issuer (one lane): tcgen05.mma.kind::tf32, tcgen05.commit -> mbarriers
//   warps 2-5   converter: lo tiles
//   warps 6-13  two epilogue groups alternating over the two TMEM accumulators: tcgen05.ld -> bias/act ->
//               coalesced global stores (a warp writes 128 B per output channel)
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc.cuh"

using namespace tc;

namespace {

constexpr int kThreads = 448;

struct PwTc {
  int B, Co, Np, C1, C1p, C2, C2p, Ks, Qp, R, S;
  int tiles_per_img;
  long tiles, P;
  const float* w1; int w1_t;
  const float* w2; int w2_t;
  const float* ahi; const float* alo; const float* timg;
  const float* bias; const float* add; const float* mul; const float* dz;
  float* preact; float* y;
  int act, dact;
};

struct PwLayout {
  uint32_t w1h, w1l, w2h, w2l, th, tl, stages, stage_bytes, x1lo, x2, x2lo, ah, al, bars, total;
};

__host__ __device__ inline PwLayout pw_layout(const PwTc& p) {
  PwLayout L;
  const uint32_t wb1 = (uint32_t)p.Np * p.C1p * 4, wb2 = (uint32_t)p.Np * p.C2p * 4, tb = 128u * p.Ks * 4;
  const uint32_t xb1 = (uint32_t)p.C1p * 512, xb2 = (uint32_t)p.C2p * 512, ab = (uint32_t)p.Ks * p.Np * 4;
  uint32_t o = 0;
  L.w1h = o; o += wb1; L.w1l = o; o += wb1;
  L.w2h = o; o += wb2; L.w2l = o; o += wb2;
  L.th = o; o += tb; L.tl = o; o += tb;
  o = (o + 1023u) & ~1023u;
  L.stages = o;
  L.x1lo = xb1; L.x2 = 2 * xb1; L.x2lo = 2 * xb1 + xb2; L.ah = 2 * xb1 + 2 * xb2; L.al = L.ah + ab;
  L.stage_bytes = (L.al + ab + 1023u) & ~1023u;
  o += L.stage_bytes * p.S;
  L.bars = o;
  o += 8 * (3 * p.S + 4) + 16;
  L.total = o + 1024;  // slack for the manual 1024-byte alignment of the dynamic window
  return L;
}

template <int MODE>
__device__ __forceinline__ float epi_value(float z, float mulv, float dzv, int act, int dact) {
  // MODE 1: no activation, no dact;  2: GELU;  3: multiply by GELU'(dz);  0: generic
  if (MODE == 1) return z * mulv;
  if (MODE == 2) return b2no_act(z, B2NO_ACT_GELU) * mulv;
  if (MODE == 3) return z * mulv * b2no_act_grad(dzv, B2NO_ACT_GELU);
  float v = b2no_act(z, act) * mulv;
  if (dact) v *= b2no_act_grad(dzv, dact);
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
k_pw_tc(const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2, const PwTc p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const PwLayout L = pw_layout(p);
  uint64_t* full = (uint64_t*)(smem + L.bars);
  uint64_t* cvt = full + p.S;
  uint64_t* empty = cvt + p.S;
  uint64_t* acc_full = empty + p.S;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tslot = (uint32_t*)(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t ncols = 32;
  while (ncols < 2u * p.Np) ncols <<= 1;

  // ---- one-time setup: weights (hi/lo, K-major core-matrix layout), spectral T image, barriers, TMEM ----
  for (int i = tid; i < p.Np * p.C1p; i += kThreads) {
    const int n = i / p.C1p, k = i - n * p.C1p;
    float w = 0.f;
    if (n < p.Co && k < p.C1) w = p.w1_t ? p.w1[(size_t)k * p.Co + n] : p.w1[(size_t)n * p.C1 + k];
    const float hi = tf32_rna(w);
    *(float*)(smem + L.w1h + kmajor_off(n, k, p.C1p)) = hi;
    *(float*)(smem + L.w1l + kmajor_off(n, k, p.C1p)) = tf32_rna(w - hi);
  }
  for (int i = tid; i < p.Np * p.C2p; i += kThreads) {
    const int n = i / p.C2p, k = i - n * p.C2p;
    float w = 0.f;
    if (n < p.Co && k < p.C2) w = p.w2_t ? p.w2[(size_t)k * p.Co + n] : p.w2[(size_t)n * p.C2 + k];
    const float hi = tf32_rna(w);
    *(float*)(smem + L.w2h + kmajor_off(n, k, p.C2p)) = hi;
    *(float*)(smem + L.w2l + kmajor_off(n, k, p.C2p)) = tf32_rna(w - hi);
  }
  for (int i = tid; i < 2 * 128 * p.Ks; i += kThreads) ((float*)(smem + L.th))[i] = p.timg[i];
  if (tid == 0) {
    for (int s = 0; s < p.S; s++) { mbar_init(&full[s], 1); mbar_init(&cvt[s], 128); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tslot, ncols);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm1); if (p.C2p) tma_prefetch_desc(&tm2); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tslot;

  const uint32_t xb1 = (uint32_t)p.C1p * 512, xb2 = (uint32_t)p.C2p * 512, ab = (uint32_t)p.Ks * p.Np * 4;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t bytes = xb1 + xb2 + 2 * ab;
      int it = 0;
      for (long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, it++) {
        const int s = it % p.S;
        const uint32_t ph = (uint32_t)(it / p.S) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], bytes);
        uint8_t* st = smem + L.stages + (size_t)s * L.stage_bytes;
        const int b = (int)(tile / p.tiles_per_img);
        const int t_in_img = (int)(tile - (long)b * p.tiles_per_img);
        const int p0 = t_in_img * 128;
        for (int j = 0; j < 4; j++) tma_load_3d(st + j * (xb1 / 4), &tm1, &full[s], p0 + 32 * j, 0, b);
        if (p.C2p)
          for (int j = 0; j < 4; j++) tma_load_3d(st + L.x2 + j * (xb2 / 4), &tm2, &full[s], p0 + 32 * j, 0, b);
        if (p.Ks) {
          const size_t row0 = ((size_t)b * p.tiles_per_img + t_in_img) * p.R;
          const size_t off = row0 * (size_t)p.Qp * p.Np;
          bulk_load(st + L.ah, p.ahi + off, ab, &full[s]);
          bulk_load(st + L.al, p.alo + off, ab, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t id_x = idesc_tf32(128, p.Np, 1, 0), id_t = idesc_tf32(128, p.Np, 0, 0);
      const uint32_t sbase = smem_u32(smem);
      const uint64_t d_w1h = smem_desc(sbase + L.w1h, 128, (p.C1p / 4) * 128, LAYOUT_NONE);
      const uint64_t d_w1l = smem_desc(sbase + L.w1l, 128, (p.C1p / 4) * 128, LAYOUT_NONE);
      const uint64_t d_w2h = smem_desc(sbase + L.w2h, 128, (p.C2p / 4) * 128, LAYOUT_NONE);
      const uint64_t d_w2l = smem_desc(sbase + L.w2l, 128, (p.C2p / 4) * 128, LAYOUT_NONE);
      const uint64_t d_th = smem_desc(sbase + L.th, 128, (p.Ks / 4) * 128, LAYOUT_NONE);
      const uint64_t d_tl = smem_desc(sbase + L.tl, 128, (p.Ks / 4) * 128, LAYOUT_NONE);
      const uint32_t lbo_a = (uint32_t)(p.Np / 8) * 128;
      int it = 0;
      for (long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, it++) {
        const int s = it % p.S;
        const uint32_t ph = (uint32_t)(it / p.S) & 1u;
        const int a = it & 1;
        const uint32_t aph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&full[s], ph);
        mbar_wait(&cvt[s], ph);
        mbar_wait(&acc_empty[a], aph ^ 1u);
        tc_fence_after();
        const uint32_t st = sbase + L.stages + (uint32_t)s * L.stage_bytes;
        const uint32_t d = tbase + (uint32_t)a * p.Np;
        uint32_t acc = 0;
        // MN-major X operand: atom = 4 channel rows x 128 B (SBO 512); 32-px groups C*128 B apart (LBO)
        const uint64_t d_x1 = smem_desc(st, p.C1p * 128, 512, LAYOUT_SW128_32B);
        const uint64_t d_x1lo = smem_desc(st + L.x1lo, p.C1p * 128, 512, LAYOUT_SW128_32B);
        for (int pass = 0; pass < 3; pass++) {
          const uint64_t dx = pass == 1 ? d_x1lo : d_x1;
          const uint64_t dw = pass == 2 ? d_w1l : d_w1h;
          for (int k = 0; k < p.C1p / 8; k++) {
            mma_tf32_ss(d, dx + (uint64_t)(k * 64), dw + (uint64_t)(k * 16), id_x, acc);
            acc = 1;
          }
        }
        if (p.C2p) {
          const uint64_t d_x2 = smem_desc(st + L.x2, p.C2p * 128, 512, LAYOUT_SW128_32B);
          const uint64_t d_x2lo = smem_desc(st + L.x2lo, p.C2p * 128, 512, LAYOUT_SW128_32B);
          for (int pass = 0; pass < 3; pass++) {
            const uint64_t dx = pass == 1 ? d_x2lo : d_x2;
            const uint64_t dw = pass == 2 ? d_w2l : d_w2h;
            for (int k = 0; k < p.C2p / 8; k++) mma_tf32_ss(d, dx + (uint64_t)(k * 64), dw + (uint64_t)(k * 16), id_x, 1);
          }
        }
        if (p.Ks) {
          // B operand A'[n = channel][k = (row, q)]: K-chunks outermost (LBO = Np/8 * 128), channel groups 128 B apart
          const uint64_t d_ah = smem_desc(st + L.ah, lbo_a, 128, LAYOUT_NONE);
          const uint64_t d_al = smem_desc(st + L.al, lbo_a, 128, LAYOUT_NONE);
          for (int pass = 0; pass < 3; pass++) {
            const uint64_t dt = pass == 1 ? d_tl : d_th;
            const uint64_t da = pass == 2 ? d_al : d_ah;
            for (int k = 0; k < p.Ks / 8; k++)
              mma_tf32_ss(d, dt + (uint64_t)(k * 16), da + (uint64_t)(k * (2 * lbo_a / 16)), id_t, 1);
          }
        }
        mma_commit(&empty[s]);
        mma_commit(&acc_full[a]);
      }
    }
  } else if (warp < 6) {
    // ===================== converter: lo tiles =====================
    const int ct = tid - 64;
    int it = 0;
    for (long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, it++) {
      const int s = it % p.S;
      const uint32_t ph = (uint32_t)(it / p.S) & 1u;
      mbar_wait(&full[s], ph);
      uint8_t* st = smem + L.stages + (size_t)s * L.stage_bytes;
      const float4* src = (const float4*)st;
      float4* dst = (float4*)(st + L.x1lo);
      for (int i = ct; i < p.C1p * 32; i += 128) {
        const float4 x = src[i];
        dst[i] = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
      }
      if (p.C2p) {
        src = (const float4*)(st + L.x2);
        dst = (float4*)(st + L.x2lo);
        for (int i = ct; i < p.C2p * 32; i += 128) {
          const float4 x = src[i];
          dst[i] = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
        }
      }
      fence_proxy_async();
      mbar_arrive(&cvt[s]);
    }
  } else {
    // ===================== epilogue (two groups, one per accumulator) =====================
    const int g = (warp - 6) >> 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int t = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    for (int it = g;; it += 2) {
      const long tile = (long)blockIdx.x + (long)it * gridDim.x;
      if (tile >= p.tiles) break;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      const int b = (int)(tile / p.tiles_per_img);
      const long px = (tile - (long)b * p.tiles_per_img) * 128 + t;
      const size_t base = (size_t)b * p.Co * p.P + px;
      mbar_wait(&acc_full[g], aph);
      tc_fence_after();
      for (int c0 = 0; c0 < p.Co; c0 += 16) {
        float v[16], av[16], mv[16], dv[16];
        tmem_ld16(tbase + lane_base + (uint32_t)(g * p.Np + c0), v);
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const bool ok = c0 + j < p.Co;
          const size_t idx = base + (size_t)(c0 + j) * p.P;
          av[j] = (p.add && ok) ? __ldg(p.add + idx) : 0.f;
          mv[j] = (p.mul && ok) ? __ldg(p.mul + idx) : 1.f;
          dv[j] = (p.dz && ok) ? __ldg(p.dz + idx) : 0.f;
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j++) {
          if (c0 + j < p.Co) {
            const size_t idx = base + (size_t)(c0 + j) * p.P;
            const float z = v[j] + (p.bias ? __ldg(p.bias + c0 + j) : 0.f) + av[j];
            if (p.preact) p.preact[idx] = z;
            p.y[idx] = epi_value<MODE>(z, mv[j], dv[j], p.act, p.dact);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[g]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, ncols);
}

// A'[b][row][q/4][n/8][n%8][q%4] (hi and lo) from the kept spectrum: the second-to-last inverse stage.
//   S[b,o,h,ky] = sum_kx M[kx][h] * spec[b][o][kx][ky];   q = 2 ky -> Re S, 2 ky + 1 -> Im S
__global__ void __launch_bounds__(256)
k_inv_h(const float2* __restrict__ spec, const float2* __restrict__ M, float* __restrict__ ahi, float* __restrict__ alo,
        int B, int Co, int Np, int Kx, int H, int Ky, int Qp) {
  const int ng = Np >> 3, nq = Qp >> 2;
  const long total = (long)B * H * nq * ng * 16;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int l0 = (int)(idx & 1), o8 = (int)((idx >> 1) & 7);
    long r = idx >> 4;
    const int og = (int)(r % ng); r /= ng;
    const int kq = (int)(r % nq); r /= nq;
    const int h = (int)(r % H);
    const int b = (int)(r / H);
    const int ky = kq * 2 + l0, o = og * 8 + o8;
    float sr = 0.f, si = 0.f;
    if (o < Co && ky < Ky) {
      const float2* sp = spec + (((size_t)b * Co + o) * Kx) * Ky + ky;
      for (int kx = 0; kx < Kx; kx++) {
        const float2 m = __ldg(M + (size_t)kx * H + h);
        const float2 v = __ldg(sp + (size_t)kx * Ky);
        sr = fmaf(m.x, v.x, fmaf(-m.y, v.y, sr));
        si = fmaf(m.x, v.y, fmaf(m.y, v.x, si));
      }
    }
    const size_t off = ((size_t)b * H + h) * Qp * Np + (size_t)kq * ng * 32 + og * 32 + o8 * 4 + l0 * 2;
    const float hr = tf32_rna(sr), hi = tf32_rna(si);
    *reinterpret_cast<float2*>(ahi + off) = make_float2(hr, hi);
    *reinterpret_cast<float2*>(alo + off) = make_float2(tf32_rna(sr - hr), tf32_rna(si - hi));
  }
}

int g_tc_state = -1;  // -1 unknown, 0 off, 1 on
long g_tc_launches = 0;

}  // namespace

bool b2no_tc_available() {
  if (g_tc_state < 0) {
    int on = 1;
    const char* e = getenv("B2NO_DISABLE_TC");
    if (e && e[0] && e[0] != '0') on = 0;
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        major != 10)
      on = 0;
    if (on && !encode_fn()) on = 0;
    g_tc_state = on;
  }
  return g_tc_state == 1;
}

extern "C" int64_t b2no_tensor_core_launches(void) { return g_tc_launches; }

extern "C" int b2no_set_tensor_core_mode(int on) {
  g_tc_state = -1;
  if (!on) { g_tc_state = 0; return 0; }
  return b2no_tc_available() ? 1 : 0;
}

// Returns 0 when the tile kernel ran, 1 when the shape is not eligible (caller falls back to the CUDA-core
// kernel), otherwise an error code.
int b2no_tc_pointwise(const b2no_plan* plan, int which, const float* spec, float* y, float* work, int batch, int channels,
                      long pixels, const b2no_epilogue* e, cudaStream_t st) {
  if (!b2no_tc_available()) return 1;
  if (!e || !e->pw_w || !e->pw_x || e->pw_ci < 1) return 1;          // needs at least one 1x1 operand (the TMA-fed A tile)
  if (pixels % 128 != 0 || channels > 256 || e->pw_ci > 128 || e->pw2_ci > 128) return 1;
  if (((uintptr_t)e->pw_x | (uintptr_t)y) & 15) return 1;
  const bool has2 = e->pw2_w && e->pw2_x && e->pw2_ci > 0;
  if (has2 && ((uintptr_t)e->pw2_x & 15)) return 1;
  PwTc p;
  memset(&p, 0, sizeof(p));
  p.B = batch; p.Co = channels; p.Np = b2no_round_up(channels, 16); p.P = pixels;
  p.C1 = e->pw_ci; p.C1p = b2no_round_up(e->pw_ci, 8);
  p.C2 = has2 ? e->pw2_ci : 0; p.C2p = has2 ? b2no_round_up(e->pw2_ci, 8) : 0;
  p.tiles_per_img = (int)(pixels / 128);
  p.tiles = (long)batch * p.tiles_per_img;
  p.w1 = e->pw_w; p.w1_t = e->pw_transposed; p.w2 = has2 ? e->pw2_w : nullptr; p.w2_t = e->pw2_transposed;
  p.bias = e->bias; p.add = e->add; p.mul = e->mul; p.dz = e->dact_z; p.preact = e->preact; p.y = y;
  p.act = e->act; p.dact = e->dact_z ? e->dact : 0;
  if (spec) {
    if (!plan || plan->g.ndim != 2 || !work) return 1;
    const b2no_tc_tables& tt = plan->tc[which];
    if (!tt.timg) return 1;
    p.Ks = tt.Ks; p.Qp = tt.Qp; p.R = tt.R; p.timg = tt.timg;
    const int32_t* n = which == 0 ? plan->g.nout : plan->g.nin;
    if ((long)n[0] * n[1] != pixels) return B2NO_E_ARG;
    const size_t afl = (size_t)batch * n[0] * p.Qp * p.Np;
    float* ahi = work;
    float* alo = work + afl;
    p.ahi = ahi; p.alo = alo;
  }
  // shared-memory budget -> number of stages
  int dev = 0, max_smem = 0;
  B2NO_CHECK_CUDA(cudaGetDevice(&dev));
  B2NO_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  PwLayout L;
  for (p.S = 4; p.S >= 2; p.S--) {
    L = pw_layout(p);
    if ((int)L.total <= max_smem) break;
  }
  if (p.S < 2) return 1;

  CUtensorMap tm1, tm2;
  {
    uint64_t dims[3] = {(uint64_t)pixels, (uint64_t)p.C1, (uint64_t)batch};
    uint64_t str[3] = {4, (uint64_t)pixels * 4, (uint64_t)pixels * 4 * p.C1};
    uint32_t box[3] = {32, (uint32_t)p.C1p, 1};
    if (make_tmap_f32(&tm1, e->pw_x, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
    tm2 = tm1;
    if (has2) {
      uint64_t dims2[3] = {(uint64_t)pixels, (uint64_t)p.C2, (uint64_t)batch};
      uint64_t str2[3] = {4, (uint64_t)pixels * 4, (uint64_t)pixels * 4 * p.C2};
      uint32_t box2[3] = {32, (uint32_t)p.C2p, 1};
      if (make_tmap_f32(&tm2, e->pw2_x, 3, dims2, str2, box2, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
    }
  }
  if (spec) {
    const int32_t* n = which == 0 ? plan->g.nout : plan->g.nin;
    const float2* M = which == 0 ? plan->m_inv[0] : plan->m_adjfwd[0];
    const long total = (long)batch * n[0] * (p.Qp / 4) * (p.Np / 8) * 16;
    long blocks = (total + 255) / 256;
    const long cap = (long)b2no_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k_inv_h<<<(unsigned)blocks, 256, 0, st>>>((const float2*)spec, M, (float*)p.ahi, (float*)p.alo, batch, channels, p.Np,
                                              plan->K[0], n[0], plan->K[1], p.Qp);
    B2NO_LAUNCH_CHECK();
  }
  int mode = 0;
  if (!p.dact && p.act == B2NO_ACT_NONE) mode = 1;
  else if (!p.dact && p.act == B2NO_ACT_GELU) mode = 2;
  else if (p.dact == B2NO_ACT_GELU && p.act == B2NO_ACT_NONE) mode = 3;
  long grid = p.tiles < b2no_sm_count() ? p.tiles : b2no_sm_count();
#define LAUNCH(M)                                                                                                  \
  do {                                                                                                             \
    B2NO_CHECK_CUDA(cudaFuncSetAttribute(k_pw_tc<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));  \
    k_pw_tc<M><<<(unsigned)grid, kThreads, L.total, st>>>(tm1, tm2, p);                                           \
  } while (0)
  switch (mode) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 3: LAUNCH(3); break;
    default: LAUNCH(0); break;
  }
#undef LAUNCH
  B2NO_LAUNCH_CHECK();
  g_tc_launches++;
  return 0;
}
