// PINO PDE-residual + initial-condition loss of the channel-flow observer, fused.
//
// Reference: FDM_NS_vorticity (libs/envs/diff_control_env.py:5-41) and Channelflow_PINO_loss (:44-60):
//     Du = w_t + u . grad w - nu lap w           spectral derivatives over (x, y) for every time slice, central difference in t
//     loss_f = LpLoss.rel(Du, forcing)           loss_ic = LpLoss.rel(w[..., 0], u0)          (libs/pino_utils/losses.py:182-194)
// The reference takes a full fft2 of every slice, builds five spectra (stream-function velocities, vorticity gradient,
// Laplacian), runs five irfft2 and a dozen elementwise passes over (B, N, N, T) tensors.  Here ONE CTA owns one (sample,
// time slice) plane and keeps everything between the input plane and the residual in shared memory / registers:
//
//     A  = P Cy^T - i P Sy^T                (real -> half spectrum along y)          P = w[b, :, :, t]
//     W  = (Cx - i Sx) A                    (complex, along x)                       N x (N/2 + 1) kept columns
//     for each of the five multipliers m_j(kx, ky):   G_j = (Cx + i Sx)^T (m_j W) / N,   field_j = Re-part C2R of G_j / N
//     adv = ux wx + uy wy - nu wlap  (register tiles),   Du = (w[t+1] - w[t-1]) / (2 dt) + adv
//     sum (Du - f)^2, sum f^2, and for t = 0 sum (w - u0)^2, sum u0^2 -> warp-shuffle + block reduction -> partial[b][t][4]
//
// as mode-restricted DFT contractions (the form every transform of this library has): small real matrix products on
// shared-memory operands with 4x4 / 4x3 register tiles.  Quirks kept: the signed wavenumber of index N/2 is -N/2 on both axes
// (:15-19); lap[0,0] = 1 for EVERY use, including -lap * w_h (:21-29); the imaginary parts of the ky = 0 and Nyquist columns
// are dropped by the C2R transform.  The backward kernel is the hand-derived adjoint of the same chain (same 36 products,
// transposed), one CTA per plane again, no atomics: each plane's gradient is produced by exactly one CTA.
#include "common.cuh"

namespace {

constexpr int kPinoThreads = 256;
constexpr int kPinoMaxN = 64;

struct PinoDims {
  int B, N, T, H, Hp, ldn, ldh;     // H = N/2 + 1 kept columns, Hp = H rounded up to 3; leading dimensions (odd: conflict-free)
};

__host__ __device__ inline PinoDims pino_dims(int B, int N, int T) {
  PinoDims d;
  d.B = B; d.N = N; d.T = T; d.H = N / 2 + 1; d.Hp = ((d.H + 2) / 3) * 3;
  d.ldn = N + 1; d.ldh = d.Hp + 1 + ((d.Hp + 1) % 2 == 0 ? 1 : 0);
  return d;
}

__host__ __device__ inline size_t pino_smem_floats(const PinoDims& d) {
  // tables C, S [N][ldn]; plane [N][ldn]; three complex N x Hp buffers (2 floats planes each)
  return (size_t)3 * d.N * d.ldn + (size_t)6 * d.N * d.ldh + 64;
}

__device__ __forceinline__ float signed_k(int i, int N) { return (float)(i < N / 2 ? i : i - N); }

// out tile (rows tm + i * MT, cols tp * RP + j) += sum_k a(row, k) * b(k, col)
template <int RM, int RP, class FA, class FB>
__device__ __forceinline__ void mm_tile(float (&acc)[RM][RP], int tm, int tp, int MT, int K, FA a, FB b) {
#pragma unroll 4
  for (int k = 0; k < K; k++) {
    float av[RM], bv[RP];
#pragma unroll
    for (int i = 0; i < RM; i++) av[i] = a(tm + i * MT, k);
#pragma unroll
    for (int j = 0; j < RP; j++) bv[j] = b(k, tp * RP + j);
#pragma unroll
    for (int i = 0; i < RM; i++)
#pragma unroll
      for (int j = 0; j < RP; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// block-wide sum of NV values per thread -> result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = warp_sum_f(v[i]);
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; i++) scratch[warp * NV + i] = v[i];
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) {
      float s = 0.f;
      for (int w = 0; w < kPinoThreads / 32; w++) s += scratch[w * NV + i];
      v[i] = s;
    }
  }
  __syncthreads();
}

// multipliers: (sr, si) = m_j(kx, ky) (wr, wi), j = 0 ux, 1 wx, 2 uy, 3 wy, 4 wlap   (diff_control_env.py:21-29)
__device__ __forceinline__ void pino_mult(int j, float kx, float ky, float lap, float wr, float wi, float* sr, float* si) {
  switch (j) {
    case 0: { const float c = ky / lap; *sr = -c * wi; *si = c * wr; break; }        // ux_h = i ky w_h / lap
    case 1: *sr = -kx * wi; *si = kx * wr; break;                                    // wx_h = i kx w_h
    case 2: { const float c = kx / lap; *sr = c * wi; *si = -c * wr; break; }        // uy_h = -i kx w_h / lap
    case 3: *sr = -ky * wi; *si = ky * wr; break;                                    // wy_h = i ky w_h
    default: *sr = -lap * wr; *si = -lap * wi; break;                                // wlap_h = -lap w_h
  }
}
// adjoint of the same real-linear map: (dwr, dwi) += m_j^T (dsr, dsi)
__device__ __forceinline__ void pino_mult_adj(int j, float kx, float ky, float lap, float dsr, float dsi, float* dwr, float* dwi) {
  switch (j) {
    case 0: { const float c = ky / lap; *dwr += c * dsi; *dwi -= c * dsr; break; }
    case 1: *dwr += kx * dsi; *dwi -= kx * dsr; break;
    case 2: { const float c = kx / lap; *dwr -= c * dsi; *dwi += c * dsr; break; }
    case 3: *dwr += ky * dsi; *dwi -= ky * dsr; break;
    default: *dwr -= lap * dsr; *dwi -= lap * dsi; break;
  }
}

struct PinoSmem {
  float* C; float* S; float* P;
  float* Ar; float* Ai; float* Wr; float* Wi; float* Gr; float* Gi;
  float* red;
};

__device__ __forceinline__ PinoSmem pino_carve(float* smem, const PinoDims& d) {
  PinoSmem s;
  const size_t nn = (size_t)d.N * d.ldn, nh = (size_t)d.N * d.ldh;
  s.C = smem; s.S = s.C + nn; s.P = s.S + nn;
  s.Ar = s.P + nn; s.Ai = s.Ar + nh; s.Wr = s.Ai + nh; s.Wi = s.Wr + nh; s.Gr = s.Wi + nh; s.Gi = s.Gr + nh;
  s.red = s.Gi + nh;
  return s;
}

__device__ __forceinline__ void pino_tables(const PinoSmem& s, const PinoDims& d) {
  for (int i = threadIdx.x; i < d.N * d.N; i += kPinoThreads) {
    const int k = i / d.N, n = i - k * d.N;
    float sn, cs;
    sincospif(2.0f * (float)((k * n) % d.N) / (float)d.N, &sn, &cs);
    s.C[k * d.ldn + n] = cs;
    s.S[k * d.ldn + n] = sn;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
// w (B, N, N, T) t fastest; u0 (B, N, N); forcing (N, N); nu (B,)
// du_p (B, T, N, N): residual planes (planes 0 and T-1 are not written); fields (4, B, T, N, N): ux, wx, uy, wy for the backward
// partial (B, T, 4): sum (Du - f)^2, sum f^2, sum (w0 - u0)^2, sum u0^2
__global__ void __launch_bounds__(kPinoThreads)
k_pino_residual_fwd(const float* __restrict__ w, const float* __restrict__ u0, const float* __restrict__ forcing,
                    const float* __restrict__ nu, float inv_2dt, float* __restrict__ du_p, float* __restrict__ fields,
                    float* __restrict__ partial, PinoDims d) {
  extern __shared__ float smem_f[];
  const PinoSmem s = pino_carve(smem_f, d);
  const int N = d.N, T = d.T, H = d.H, ldn = d.ldn, ldh = d.ldh;
  const int b = blockIdx.x / T, t = blockIdx.x - b * T;
  const int tid = threadIdx.x;
  pino_tables(s, d);
  const float* wb = w + (size_t)b * N * N * T;
  for (int i = tid; i < N * N; i += kPinoThreads) {
    const int x = i / N, y = i - x * N;
    s.P[x * ldn + y] = __ldg(wb + (size_t)i * T + t);
  }
  // zero the pad columns of the N x Hp buffers once (they are read as B-operand columns >= H)
  for (int i = tid; i < 6 * N * ldh; i += kPinoThreads) s.Ar[i] = 0.f;
  __syncthreads();

  const int MT = N / 4;
  const int PTh = d.Hp / 3;                 // column tiles of the N x Hp products (4 x 3 register tiles)
  const int PTn = N / 4;                    // column tiles of the N x N products (4 x 4 register tiles)
  // ---- A = P (Cy - i Sy)^T ----
  for (int tile = tid; tile < MT * PTh; tile += kPinoThreads) {
    const int tm = tile % MT, tp = tile / MT;
    float ar[4][3] = {}, ai[4][3] = {};
    mm_tile<4, 3>(ar, tm, tp, MT, N, [&](int m, int k) { return s.P[m * ldn + k]; },
                  [&](int k, int p) { return p < H ? s.C[p * ldn + k] : 0.f; });
    mm_tile<4, 3>(ai, tm, tp, MT, N, [&](int m, int k) { return s.P[m * ldn + k]; },
                  [&](int k, int p) { return p < H ? s.S[p * ldn + k] : 0.f; });
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int m = tm + i * MT, p = tp * 3 + j;
        s.Ar[m * ldh + p] = ar[i][j];
        s.Ai[m * ldh + p] = -ai[i][j];
      }
  }
  __syncthreads();
  // ---- W = (Cx - i Sx) A ----
  for (int tile = tid; tile < MT * PTh; tile += kPinoThreads) {
    const int tm = tile % MT, tp = tile / MT;
    float wr[4][3] = {}, wi[4][3] = {}, t1[4][3] = {}, t2[4][3] = {};
    auto cx = [&](int m, int k) { return s.C[m * ldn + k]; };
    auto sx = [&](int m, int k) { return s.S[m * ldn + k]; };
    auto ar = [&](int k, int p) { return s.Ar[k * ldh + p]; };
    auto ai = [&](int k, int p) { return s.Ai[k * ldh + p]; };
    mm_tile<4, 3>(wr, tm, tp, MT, N, cx, ar);
    mm_tile<4, 3>(t1, tm, tp, MT, N, sx, ai);
    mm_tile<4, 3>(wi, tm, tp, MT, N, cx, ai);
    mm_tile<4, 3>(t2, tm, tp, MT, N, sx, ar);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int m = tm + i * MT, p = tp * 3 + j;
        s.Wr[m * ldh + p] = wr[i][j] + t1[i][j];
        s.Wi[m * ldh + p] = wi[i][j] - t2[i][j];
      }
  }
  __syncthreads();

  // ---- five fields; the N x N results stay in register tiles (4 x 4 per thread) ----
  const float invN = 1.0f / (float)N;
  float adv[4][4] = {}, keep[4][4];
  const bool has_tile = tid < MT * PTn;             // N = 64: every thread owns one 4 x 4 tile of the plane
  const int tmn = tid % MT, tpn = tid / MT;
  const float nub = __ldg(nu + b);
  float* fl = fields + ((size_t)b * T + t) * N * N;
  const size_t fstride = (size_t)d.B * T * N * N;
  for (int j = 0; j < 5; j++) {
    // S = m_j W  (into the A buffers), pad columns stay zero
    for (int i = tid; i < N * H; i += kPinoThreads) {
      const int kxi = i / H, kyi = i - kxi * H;
      const float kx = signed_k(kxi, N), ky = signed_k(kyi, N);
      float lap = kx * kx + ky * ky;
      if (kxi == 0 && kyi == 0) lap = 1.0f;
      float sr, si;
      pino_mult(j, kx, ky, lap, s.Wr[kxi * ldh + kyi], s.Wi[kxi * ldh + kyi], &sr, &si);
      s.Ar[kxi * ldh + kyi] = sr;
      s.Ai[kxi * ldh + kyi] = si;
    }
    __syncthreads();
    // G[x][ky] = c(ky) / N^2 * sum_kx (Cx + i Sx)[kx][x] S[kx][ky]      (the 1/N of both inverse stages and the C2R weights)
    for (int tile = tid; tile < MT * PTh; tile += kPinoThreads) {
      const int tm = tile % MT, tp = tile / MT;
      float g1[4][3] = {}, g2[4][3] = {}, g3[4][3] = {}, g4[4][3] = {};
      auto cxt = [&](int m, int k) { return s.C[k * ldn + m]; };
      auto sxt = [&](int m, int k) { return s.S[k * ldn + m]; };
      auto sr = [&](int k, int p) { return s.Ar[k * ldh + p]; };
      auto si = [&](int k, int p) { return s.Ai[k * ldh + p]; };
      mm_tile<4, 3>(g1, tm, tp, MT, N, cxt, sr);
      mm_tile<4, 3>(g2, tm, tp, MT, N, sxt, si);
      mm_tile<4, 3>(g3, tm, tp, MT, N, cxt, si);
      mm_tile<4, 3>(g4, tm, tp, MT, N, sxt, sr);
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jj = 0; jj < 3; jj++) {
          const int m = tm + i * MT, p = tp * 3 + jj;
          const float c = (p == 0 || p == N / 2) ? 1.0f : 2.0f;
          const float sc = (p < H) ? c * invN * invN : 0.f;
          s.Gr[m * ldh + p] = (g1[i][jj] - g2[i][jj]) * sc;
          s.Gi[m * ldh + p] = (g3[i][jj] + g4[i][jj]) * sc;
        }
    }
    __syncthreads();
    // field[x][y] = sum_ky Gr[x][ky] C[ky][y] - Gi[x][ky] S[ky][y]
    if (has_tile) {
      float f1[4][4] = {}, f2[4][4] = {};
      mm_tile<4, 4>(f1, tmn, tpn, MT, H, [&](int m, int k) { return s.Gr[m * ldh + k]; },
                    [&](int k, int p) { return s.C[k * ldn + p]; });
      mm_tile<4, 4>(f2, tmn, tpn, MT, H, [&](int m, int k) { return s.Gi[m * ldh + k]; },
                    [&](int k, int p) { return s.S[k * ldn + p]; });
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          const float v = f1[i][jj] - f2[i][jj];
          const int x = tmn + i * MT, y = tpn * 4 + jj;
          if (j < 4) fl[(size_t)j * fstride + (size_t)x * N + y] = v;
          if (j == 0 || j == 2) keep[i][jj] = v;                    // ux, uy wait for their partner
          else if (j == 1 || j == 3) adv[i][jj] = fmaf(keep[i][jj], v, adv[i][jj]);
          else adv[i][jj] = fmaf(-nub, v, adv[i][jj]);
        }
    }
    __syncthreads();
  }
  // ---- residual + reductions ----
  float sums[4] = {0.f, 0.f, 0.f, 0.f};
  if (has_tile) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const int x = tmn + i * MT, y = tpn * 4 + jj;
        const size_t pix = (size_t)x * N + y;
        if (t >= 1 && t <= T - 2) {
          const float dv = (__ldg(wb + pix * T + t + 1) - __ldg(wb + pix * T + t - 1)) * inv_2dt + adv[i][jj];
          du_p[((size_t)b * T + t) * N * N + pix] = dv;
          const float f = __ldg(forcing + pix);
          sums[0] = fmaf(dv - f, dv - f, sums[0]);
          sums[1] = fmaf(f, f, sums[1]);
        }
        if (t == 0) {
          const float a = s.P[x * ldn + y], u = __ldg(u0 + (size_t)b * N * N + pix);
          sums[2] = fmaf(a - u, a - u, sums[2]);
          sums[3] = fmaf(u, u, sums[3]);
        }
      }
  }
  block_sum<4>(sums, s.red);
  if (tid == 0) {
    float* pp = partial + ((size_t)b * T + t) * 4;
    pp[0] = sums[0]; pp[1] = sums[1]; pp[2] = sums[2]; pp[3] = sums[3];
  }
}

// loss[0] = loss_ic, loss[1] = loss_f (means over the batch of sqrt(num) / sqrt(den)); coef (B, 2): 1 / (B ||d|| ||y||).
// One thread per sample sums its T partials in a fixed order; thread 0 then adds the per-sample ratios in sample order
// (deterministic, no atomics).
__global__ void k_pino_finish(const float* __restrict__ partial, float* __restrict__ loss, float* __restrict__ coef, int B, int T) {
  extern __shared__ float ratios[];      // [B][2]
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float nf = 0.f, df = 0.f, ni = 0.f, di = 0.f;
    for (int t = 0; t < T; t++) {
      const float* pp = partial + ((size_t)b * T + t) * 4;
      nf += pp[0]; df += pp[1]; ni += pp[2]; di += pp[3];
    }
    ratios[b * 2 + 0] = sqrtf(ni) / sqrtf(di);
    ratios[b * 2 + 1] = sqrtf(nf) / sqrtf(df);
    coef[b * 2 + 0] = 1.0f / ((float)B * sqrtf(ni) * sqrtf(di));
    coef[b * 2 + 1] = 1.0f / ((float)B * sqrtf(nf) * sqrtf(df));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int b = 0; b < B; b++) { a += ratios[b * 2]; c += ratios[b * 2 + 1]; }
    loss[0] = a / (float)B;
    loss[1] = c / (float)B;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward: dw = g_ic d loss_ic / dw + g_f d loss_f / dw          gup[0] = upstream gradient of loss_ic, gup[1] of loss_f
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPinoThreads)
k_pino_residual_bwd(const float* __restrict__ w, const float* __restrict__ u0, const float* __restrict__ forcing,
                    const float* __restrict__ nu, float inv_2dt, const float* __restrict__ du_p,
                    const float* __restrict__ fields, const float* __restrict__ coef, const float* __restrict__ gup,
                    float* __restrict__ dw, PinoDims d) {
  extern __shared__ float smem_f[];
  const PinoSmem s = pino_carve(smem_f, d);
  const int N = d.N, T = d.T, H = d.H, ldn = d.ldn, ldh = d.ldh;
  const int b = blockIdx.x / T, t = blockIdx.x - b * T;
  const int tid = threadIdx.x;
  const int MT = N / 4, PTh = d.Hp / 3, PTn = N / 4;
  const bool has_tile = tid < MT * PTn;
  const int tmn = tid % MT, tpn = tid / MT;
  const float gf = __ldg(gup + 1) * __ldg(coef + b * 2 + 1);       // d loss_f / d Du = gf (Du - f)
  const float gi = __ldg(gup + 0) * __ldg(coef + b * 2 + 0);
  const float nub = __ldg(nu + b);
  const bool interior = t >= 1 && t <= T - 2;
  float dp[4][4] = {};                                             // this thread's tile of d loss / d w[b, :, :, t]

  if (interior) {
    pino_tables(s, d);
    // dW accumulators (Wr / Wi buffers) start at zero; pad columns of every N x Hp buffer too
    for (int i = tid; i < 6 * N * ldh; i += kPinoThreads) s.Ar[i] = 0.f;
    __syncthreads();
    const float* dup = du_p + ((size_t)b * T + t) * N * N;
    const float* fl = fields + ((size_t)b * T + t) * N * N;
    const size_t fstride = (size_t)d.B * T * N * N;
    const float invN = 1.0f / (float)N;
    for (int j = 0; j < 5; j++) {
      // h_j = g * partner_j:  ux <-> wx, uy <-> wy, wlap <-> -nu
      for (int i = tid; i < N * N; i += kPinoThreads) {
        const int x = i / N, y = i - x * N;
        const float g = gf * (__ldg(dup + i) - __ldg(forcing + i));
        float pr;
        if (j == 0) pr = __ldg(fl + 1 * fstride + i);
        else if (j == 1) pr = __ldg(fl + 0 * fstride + i);
        else if (j == 2) pr = __ldg(fl + 3 * fstride + i);
        else if (j == 3) pr = __ldg(fl + 2 * fstride + i);
        else pr = -nub;
        s.P[x * ldn + y] = g * pr;
      }
      __syncthreads();
      // dG[x][ky] = c(ky) / N^2 * (sum_y h C[ky][y],  -sum_y h S[ky][y])
      for (int tile = tid; tile < MT * PTh; tile += kPinoThreads) {
        const int tm = tile % MT, tp = tile / MT;
        float a1[4][3] = {}, a2[4][3] = {};
        auto hp = [&](int m, int k) { return s.P[m * ldn + k]; };
        mm_tile<4, 3>(a1, tm, tp, MT, N, hp, [&](int k, int p) { return p < H ? s.C[p * ldn + k] : 0.f; });
        mm_tile<4, 3>(a2, tm, tp, MT, N, hp, [&](int k, int p) { return p < H ? s.S[p * ldn + k] : 0.f; });
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int jj = 0; jj < 3; jj++) {
            const int m = tm + i * MT, p = tp * 3 + jj;
            const float c = (p == 0 || p == N / 2) ? 1.0f : 2.0f;
            const float sc = (p < H) ? c * invN * invN : 0.f;
            s.Gr[m * ldh + p] = a1[i][jj] * sc;
            s.Gi[m * ldh + p] = -a2[i][jj] * sc;
          }
      }
      __syncthreads();
      // dS[kx][ky] = sum_x (Cx[kx][x] dGr + Sx[kx][x] dGi,  -Sx[kx][x] dGr + Cx[kx][x] dGi);  dW += m_j^T dS
      for (int tile = tid; tile < MT * PTh; tile += kPinoThreads) {
        const int tm = tile % MT, tp = tile / MT;
        float q1[4][3] = {}, q2[4][3] = {}, q3[4][3] = {}, q4[4][3] = {};
        auto cx = [&](int m, int k) { return s.C[m * ldn + k]; };
        auto sx = [&](int m, int k) { return s.S[m * ldn + k]; };
        auto gr = [&](int k, int p) { return s.Gr[k * ldh + p]; };
        auto gi2 = [&](int k, int p) { return s.Gi[k * ldh + p]; };
        mm_tile<4, 3>(q1, tm, tp, MT, N, cx, gr);
        mm_tile<4, 3>(q2, tm, tp, MT, N, sx, gi2);
        mm_tile<4, 3>(q3, tm, tp, MT, N, sx, gr);
        mm_tile<4, 3>(q4, tm, tp, MT, N, cx, gi2);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int jj = 0; jj < 3; jj++) {
            const int kxi = tm + i * MT, kyi = tp * 3 + jj;
            if (kyi < H) {
              const float kx = signed_k(kxi, N), ky = signed_k(kyi, N);
              float lap = kx * kx + ky * ky;
              if (kxi == 0 && kyi == 0) lap = 1.0f;
              float dwr = s.Wr[kxi * ldh + kyi], dwi = s.Wi[kxi * ldh + kyi];
              pino_mult_adj(j, kx, ky, lap, q1[i][jj] + q2[i][jj], q4[i][jj] - q3[i][jj], &dwr, &dwi);
              s.Wr[kxi * ldh + kyi] = dwr;          // each (kx, ky) entry belongs to exactly one thread
              s.Wi[kxi * ldh + kyi] = dwi;
            }
          }
      }
      __syncthreads();
    }
    // dA[x][ky] = sum_kx (Cx[kx][x] dWr - Sx[kx][x] dWi,  Sx[kx][x] dWr + Cx[kx][x] dWi)
    for (int tile = tid; tile < MT * PTh; tile += kPinoThreads) {
      const int tm = tile % MT, tp = tile / MT;
      float q1[4][3] = {}, q2[4][3] = {}, q3[4][3] = {}, q4[4][3] = {};
      auto cxt = [&](int m, int k) { return s.C[k * ldn + m]; };
      auto sxt = [&](int m, int k) { return s.S[k * ldn + m]; };
      auto wr = [&](int k, int p) { return s.Wr[k * ldh + p]; };
      auto wi = [&](int k, int p) { return s.Wi[k * ldh + p]; };
      mm_tile<4, 3>(q1, tm, tp, MT, N, cxt, wr);
      mm_tile<4, 3>(q2, tm, tp, MT, N, sxt, wi);
      mm_tile<4, 3>(q3, tm, tp, MT, N, sxt, wr);
      mm_tile<4, 3>(q4, tm, tp, MT, N, cxt, wi);
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jj = 0; jj < 3; jj++) {
          const int m = tm + i * MT, p = tp * 3 + jj;
          s.Ar[m * ldh + p] = (p < H) ? q1[i][jj] - q2[i][jj] : 0.f;
          s.Ai[m * ldh + p] = (p < H) ? q3[i][jj] + q4[i][jj] : 0.f;
        }
    }
    __syncthreads();
    // dP[x][y] = sum_ky dAr[x][ky] C[ky][y] - dAi[x][ky] S[ky][y]
    if (has_tile) {
      float f1[4][4] = {}, f2[4][4] = {};
      mm_tile<4, 4>(f1, tmn, tpn, MT, H, [&](int m, int k) { return s.Ar[m * ldh + k]; },
                    [&](int k, int p) { return s.C[k * ldn + p]; });
      mm_tile<4, 4>(f2, tmn, tpn, MT, H, [&](int m, int k) { return s.Ai[m * ldh + k]; },
                    [&](int k, int p) { return s.S[k * ldn + p]; });
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jj = 0; jj < 4; jj++) dp[i][jj] = f1[i][jj] - f2[i][jj];
    }
  }
  // ---- central difference in t (slice t-1 reads w[t], with +; slice t+1 with -) and the initial-condition term ----
  if (has_tile) {
    float* dwb = dw + (size_t)b * N * N * T;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const int x = tmn + i * MT, y = tpn * 4 + jj;
        const size_t pix = (size_t)x * N + y;
        float v = dp[i][jj];
        const float f = __ldg(forcing + pix);
        if (t - 1 >= 1 && t - 1 <= T - 2) v += gf * (__ldg(du_p + ((size_t)b * T + t - 1) * N * N + pix) - f) * inv_2dt;
        if (t + 1 >= 1 && t + 1 <= T - 2) v -= gf * (__ldg(du_p + ((size_t)b * T + t + 1) * N * N + pix) - f) * inv_2dt;
        if (t == 0) v += gi * (__ldg(w + ((size_t)b * N * N + pix) * T) - __ldg(u0 + (size_t)b * N * N + pix));
        dwb[pix * T + t] = v;
      }
  }
}

}  // namespace

static int pino_check(int B, int N, int T) {
  if (B < 1 || T < 3) return B2NO_E_ARG;
  if (B > 4096) return B2NO_E_UNSUPPORTED;
  if (N < 8 || N > kPinoMaxN || N % 4 != 0) return B2NO_E_UNSUPPORTED;     // register tiles: N / 4 x N / 4 <= 256 threads
  return 0;
}

extern "C" int64_t b2no_pino_residual_scratch_floats(int B, int N, int T, int which) {
  // which 0: du_p, 1: fields, 2: partial
  if (which == 0) return (int64_t)B * T * N * N;
  if (which == 1) return (int64_t)4 * B * T * N * N;
  return (int64_t)B * T * 4;
}

extern "C" int b2no_pino_residual_fwd(const float* w, const float* u0, const float* forcing, const float* nu, float t_interval,
                                      float* du_p, float* fields, float* partial, float* loss, float* coef, int B, int N, int T,
                                      void* stream) {
  if (!w || !u0 || !forcing || !nu || !du_p || !fields || !partial || !loss || !coef) return B2NO_E_ARG;
  int rc = pino_check(B, N, T);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const PinoDims d = pino_dims(B, N, T);
  const size_t smem = pino_smem_floats(d) * sizeof(float);
  B2NO_CHECK_CUDA(cudaFuncSetAttribute(k_pino_residual_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float inv_2dt = (float)(T - 1) / (2.0f * t_interval);
  k_pino_residual_fwd<<<B * T, kPinoThreads, smem, st>>>(w, u0, forcing, nu, inv_2dt, du_p, fields, partial, d);
  B2NO_LAUNCH_CHECK();
  k_pino_finish<<<1, 128, (size_t)B * 2 * sizeof(float), st>>>(partial, loss, coef, B, T);
  B2NO_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2no_pino_residual_bwd(const float* w, const float* u0, const float* forcing, const float* nu, float t_interval,
                                      const float* du_p, const float* fields, const float* coef, const float* gup, float* dw,
                                      int B, int N, int T, void* stream) {
  if (!w || !u0 || !forcing || !nu || !du_p || !fields || !coef || !gup || !dw) return B2NO_E_ARG;
  int rc = pino_check(B, N, T);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const PinoDims d = pino_dims(B, N, T);
  const size_t smem = pino_smem_floats(d) * sizeof(float);
  B2NO_CHECK_CUDA(cudaFuncSetAttribute(k_pino_residual_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float inv_2dt = (float)(T - 1) / (2.0f * t_interval);
  k_pino_residual_bwd<<<B * T, kPinoThreads, smem, st>>>(w, u0, forcing, nu, inv_2dt, du_p, fields, coef, gup, dw, d);
  B2NO_LAUNCH_CHECK();
  return 0;
}
