// Per-mode complex channel mixing and its weight gradient on the 5th-gen tensor cores (tcgen05 / TMEM).
//
//   mix      Yh[b,o,k]  (+)= sum_i Xh[b,i,k]  W[i,o,k]            the einsum of spectral_convolution.py:31-36, rno.py:51-58,71-74,
//   mix^H    gXh[b,i,k] (+)= sum_o gYh[b,o,k] conj(W[i,o,k])       basics.py:21-24 and its input adjoint
//   dW       dW[i,o,k]  (+)= sum_b conj(Xh[b,i,k]) gYh[b,o,k]
//
// For one kept mode k this is a dense complex product over the channels, batched over the samples -- the one stage of the
// spectral convolution SURVEY.md App. B marks as real tensor-core work once the batch is large (RNO: 288 modes x
// [B = 256] x [34 x 34]).  In real-block form a complex (1 x Cq)(Cq x Cp) product is a real (1 x 2Cq)(2Cq x 2Cp) one:
//
//     [Yr Yi] = [Xr Xi] [ Wr  Wi ]          row index of the weight block = (q, re|im), column = (p, re|im), both
//                       [-Wi  Wr ]          interleaved exactly like the complex64 spectra in HBM
//
// so one 128-sample tile of one mode is ONE accumulator [128 x 2Cp] in TMEM:
//   k_mix_tc   M = 128 samples (TMEM lane = sample), N = 2Cp, K = 2Cq.  A = the samples' spectrum values of this mode,
//              gathered from the (batch, channel, mode) layout by the thread that owns the lane and written to TMEM as
//              hi | lo (TS form, no shared-memory round trip); B = the mode's weight block, built once per mode in shared
//              memory (K-major core matrices, hi and lo images).  3xTF32: hi*hi + lo*hi + hi*lo.
//   k_dw_tc    M = 2Ci rows (i, re|im), N = 2Co, K = samples, accumulated in TMEM over the whole batch in chunks of 64
//              samples; both operands go through shared memory (SS form, K-major with the K chunks padded to 144 B so
//              that the per-sample scalar stores are conflict-free); the four real sums of a complex product land in
//              adjacent TMEM lanes (re / im rows of the same i) and are combined with one shuffle at read-out.
// A CTA owns one mode at a time (grid-stride over the modes); the spectra are a few tens of MB and L2-resident between
// the transform kernels and these, so the gathers are L2 traffic, not HBM.
#include <string.h>

#include "common.cuh"
#include "tc.cuh"

using namespace tc;

void b2no_tc_count_launch();

namespace {

constexpr int kMixThreads = 256;

struct MixTc {
  int B, Cq, Cp, Kt, Kp, Np, conjt, accumulate, npass;
  int layout;        // spectra: 0 (batch, channel, mode), 1 (mode, batch, channel)
  const float2* in;
  float2* out;
  b2no_weights w;
  ModeMap mm;
};

// ---------------------------------------------------------------------------------------------------------------------
// k_mix_tc
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMixThreads, 2)
k_mix_tc(const MixTc p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kp = p.Kp, Np = p.Np;
  const uint32_t wbytes = (uint32_t)Np * Kp * 4;
  uint8_t* s_bh = smem;               // weight block, hi image: [Np rows x Kp] K-major core matrices
  uint8_t* s_bl = smem + wbytes;      // lo image
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(2 * Kp + Np)) ncols <<= 1;

  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  // pad rows / columns of the weight block stay zero for the whole kernel
  for (uint32_t i = tid; i < 2 * wbytes / 16; i += kMixThreads) ((float4*)smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 0) tmem_alloc(&tslot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  const uint32_t t_ah = tbase, t_al = tbase + Kp, t_d = tbase + 2 * Kp;
  const int quad = warp & 3, grp = warp >> 2;       // TMEM lane quadrant of this warp; two warps share a quadrant
  const int row = quad * 32 + lane;
  const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
  const uint32_t idesc = idesc_tf32(128, Np, 0, 0);
  const uint32_t sbase = smem_u32(smem);
  const uint64_t d_bh = smem_desc(sbase, 128, (uint32_t)(Kp / 4) * 128, LAYOUT_NONE);
  const uint64_t d_bl = smem_desc(sbase + wbytes, 128, (uint32_t)(Kp / 4) * 128, LAYOUT_NONE);
  const int ntiles = (p.B + 127) / 128;
  uint32_t phase = 0;
  // work items = (mode, 128-sample tile), mode-major; every CTA owns a contiguous run, so a mode's weight block is staged
  // once per CTA that touches it.  (One CTA per mode left a single CTA per SM with every gather latency exposed.)
  const long items = (long)p.Kt * ntiles;
  const long per_cta = (items + gridDim.x - 1) / gridDim.x;
  const long it0 = (long)blockIdx.x * per_cta;
  const long it1 = it0 + per_cta < items ? it0 + per_cta : items;
  int staged = -1;

  for (long item = it0; item < it1; item++) {
    const int k = (int)(item / ntiles), tile = (int)(item - (long)k * ntiles);
    if (k != staged) {
      // ---- the mode's weight block -> shared memory (hi / lo), real-block form ----
      staged = k;
      int corner;
      long woff;
      decode_mode(p.mm, p.w, k, &corner, &woff);
      const float2* wb = (const float2*)p.w.corner[corner] + woff;
      const long sq = p.conjt ? p.w.stride_o : p.w.stride_i;     // stride of the contraction channel q
      const long sp = p.conjt ? p.w.stride_i : p.w.stride_o;     // stride of the output channel p
      for (int e = tid; e < p.Cq * p.Cp; e += kMixThreads) {
        const int q = e / p.Cp, pp = e - q * p.Cp;
        const float2 w = __ldg(wb + (long)q * sq + (long)pp * sp);
        const float wi = p.conjt ? -w.y : w.y;
        // rows (2pp, 2pp+1) = (re, im) of the output channel; columns (2q, 2q+1) = (re, im) of the contraction channel
        const float r0a = w.x, r0b = -wi, r1a = wi, r1b = w.x;
        const float h0a = tf32_rna(r0a), h0b = tf32_rna(r0b), h1a = tf32_rna(r1a), h1b = tf32_rna(r1b);
        const uint32_t o0 = kmajor_off(2 * pp, 2 * q, Kp), o1 = kmajor_off(2 * pp + 1, 2 * q, Kp);
        *(float2*)(s_bh + o0) = make_float2(h0a, h0b);
        *(float2*)(s_bh + o1) = make_float2(h1a, h1b);
        *(float2*)(s_bl + o0) = make_float2(tf32_rna(r0a - h0a), tf32_rna(r0b - h0b));
        *(float2*)(s_bl + o1) = make_float2(tf32_rna(r1a - h1a), tf32_rna(r1b - h1b));
      }
      fence_proxy_async();
    }
    {
      const int b = tile * 128 + row;
      // ---- A operand: this lane's sample, all contraction channels of mode k -> TMEM (hi | lo); the two warps of a
      //      quadrant take alternate 8-channel chunks; up to three chunks (24 gathers) are in flight per thread ----
      {
        const int bb = b < p.B ? b : 0;
        // layout 0: channel stride Kt (8-byte gathers, one 32-byte sector each); layout 1: the sample's Cq values of this
        // mode are contiguous (the sectors a thread touches are fully used, by itself, a few loads later)
        const float2* src = p.layout ? p.in + ((size_t)k * p.B + bb) * p.Cq : p.in + ((size_t)bb * p.Cq) * p.Kt + k;
        const size_t cs = p.layout ? 1 : (size_t)p.Kt;
        const bool vec = p.layout && !(p.Cq & 1);      // mode-major, even channel count: two channels per 16-byte load
        for (int cb = grp * 8; cb * 2 < Kp; cb += 48) {
          float2 v[3][8];
          if (vec) {
#pragma unroll
            for (int u = 0; u < 3; u++)
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                const int c = cb + 16 * u + j;
                const float4 t = (b < p.B && c < p.Cq) ? __ldg(reinterpret_cast<const float4*>(src + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[u][j] = make_float2(t.x, t.y);
                v[u][j + 1] = make_float2(t.z, t.w);
              }
          } else {
#pragma unroll
            for (int u = 0; u < 3; u++)
#pragma unroll
              for (int j = 0; j < 8; j++) {
                const int c = cb + 16 * u + j;
                v[u][j] = (b < p.B && c < p.Cq) ? __ldg(src + (size_t)c * cs) : make_float2(0.f, 0.f);
              }
          }
#pragma unroll
          for (int u = 0; u < 3; u++) {
            const int c0 = cb + 16 * u;
            if (c0 * 2 < Kp) {
              float hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 8; j++) { hi[2 * j] = v[u][j].x; hi[2 * j + 1] = v[u][j].y; }
#pragma unroll
              for (int j = 0; j < 16; j++) lo[j] = tf32_lo(hi[j]);
              tmem_st16(t_ah + lane_base + (uint32_t)(2 * c0), hi);
              tmem_st16(t_al + lane_base + (uint32_t)(2 * c0), lo);
            }
          }
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncthreads();       // A complete (and the weight block, when it was restaged); the previous item's epilogue is done
      tc_fence_after();
      if (warp == 0) {
        if (elect_one()) {
          uint32_t acc = 0;
          _Pragma("unroll") for (int pass = 0; pass < 3; pass++) {
            if (pass >= p.npass) break;   // compile-time trip count (descriptors stay folded); 1-pass mode leaves early
            const uint32_t a = pass == 1 ? t_al : t_ah;
            const uint64_t db = pass == 2 ? d_bl : d_bh;
            for (int ks = 0; ks < Kp / 8; ks++) {
              mma_tf32_ts(t_d, a + 8 * ks, db + (uint64_t)(ks * 16), idesc, acc);
              acc = 1;
            }
          }
          mma_commit(&bar);
        }
        __syncwarp();
      }
      mbar_wait(&bar, phase);
      phase ^= 1u;
      tc_fence_after();
      // ---- epilogue: 8 accumulator columns = 4 complex outputs per step; the two warps of a quadrant alternate ----
      for (int c0 = grp * 8; c0 < Np; c0 += 16) {
        float v[8];
        tmem_ld8(t_d + lane_base + (uint32_t)c0, v);
        tmem_ld_wait();
        if (b < p.B && p.layout && !(p.Cp & 1)) {
          // mode-major, even channel count: the 4 outputs are two 16-byte stores into the sample's contiguous row
#pragma unroll
          for (int j = 0; j < 4; j += 2) {
            const int o = (c0 >> 1) + j;
            if (o < p.Cp) {
              float4* dst = reinterpret_cast<float4*>(p.out + ((size_t)k * p.B + b) * p.Cp + o);
              float4 r = make_float4(v[2 * j], v[2 * j + 1], v[2 * j + 2], v[2 * j + 3]);
              if (p.accumulate) { const float4 old = *dst; r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w; }
              *dst = r;
            }
          }
        } else if (b < p.B) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int o = (c0 >> 1) + j;
            if (o < p.Cp) {
              float2* dst = p.layout ? p.out + ((size_t)k * p.B + b) * p.Cp + o : p.out + ((size_t)b * p.Cp + o) * p.Kt + k;
              float2 r = make_float2(v[2 * j], v[2 * j + 1]);
              if (p.accumulate) { const float2 old = *dst; r.x += old.x; r.y += old.y; }
              *dst = r;
            }
          }
        }
      }
      tc_fence_before();
      __syncthreads();       // the item's MMAs are complete (bar) and every thread is past its accumulator reads: the weight
                             // block and the A columns may be rewritten
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, ncols);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_dw_tc
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDwKB = 64;                        // samples per shared-memory chunk (K of one round of MMAs)
constexpr uint32_t kDwLbo = 144;                 // bytes between K-adjacent core matrices (padded: conflict-free scalar stores)
constexpr uint32_t kDwSbo = (kDwKB / 4) * kDwLbo;  // bytes between 8-row groups

struct DwTc {
  int B, Ci, Co, Kt, Np, Ga, accumulate, npass, layout;      // Ga = 8-row groups of the A image that are materialised (2 Ci rows)
  const float2* xh;
  const float2* gyh;
  b2no_weights w;
  ModeMap mm;
};

__device__ __forceinline__ uint32_t dw_off(int r, int bb) {
  return (uint32_t)(r >> 3) * kDwSbo + (uint32_t)(bb >> 2) * kDwLbo + (uint32_t)(r & 7) * 16 + (uint32_t)(bb & 3) * 4;
}

__global__ void __launch_bounds__(kMixThreads, 2)
k_dw_tc(const DwTc p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Np = p.Np;
  // A image: 2 Ci rows (i, re|im) rounded up to 8; an M = 128 MMA reads 16 row groups from its base -- the groups beyond Ga
  // are whatever follows in shared memory (the lo image, the B images) and only feed accumulator rows nobody reads
  const uint32_t abytes = (uint32_t)p.Ga * kDwSbo;
  const uint32_t bbytes = (uint32_t)(Np / 8) * kDwSbo;     // B image: Np rows (o, re|im)
  uint8_t* s_ah = smem;
  uint8_t* s_al = s_ah + abytes;
  uint8_t* s_bh = s_al + abytes;
  uint8_t* s_bl = s_bh + bbytes;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)Np) ncols <<= 1;

  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  for (uint32_t i = tid; i < (2 * abytes + 2 * bbytes) / 16; i += kMixThreads) ((float4*)smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 0) tmem_alloc(&tslot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  const int quad = warp & 3, grp = warp >> 2;
  const int m = quad * 32 + lane;                          // accumulator row (i, re|im) read by this thread
  const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
  const uint32_t idesc = idesc_tf32(128, Np, 0, 0);
  const uint32_t sbase = smem_u32(smem);
  const int bb = tid & (kDwKB - 1), part = tid / kDwKB;    // fill: thread = sample of the chunk x channel residue (4 parts)
  const int nparts = kMixThreads / kDwKB;
  const int nchunks = (p.B + kDwKB - 1) / kDwKB;
  uint32_t phase = 0;
  bool pending = false;                                    // MMAs of the previous chunk still reading shared memory

  for (int k = blockIdx.x; k < p.Kt; k += gridDim.x) {
    uint32_t acc = 0;
    for (int ch = 0; ch < nchunks; ch++) {
      const int b = ch * kDwKB + bb;
      const bool live = b < p.B;
      const int bl = live ? b : 0;
      const float2* xs = p.layout ? p.xh + ((size_t)k * p.B + bl) * p.Ci : p.xh + ((size_t)bl * p.Ci) * p.Kt + k;
      const float2* gs = p.layout ? p.gyh + ((size_t)k * p.B + bl) * p.Co : p.gyh + ((size_t)bl * p.Co) * p.Kt + k;
      const size_t cs = p.layout ? 1 : (size_t)p.Kt;
      // the previous round of MMAs must have finished reading the operand images before they are overwritten
      if (pending) { mbar_wait(&bar, phase); phase ^= 1u; pending = false; }
      // gathers in batches of 8 (all in flight before the first shared-memory store)
      for (int i0 = part; i0 < p.Ci; i0 += 8 * nparts) {
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int i = i0 + u * nparts;
          v[u] = (live && i < p.Ci) ? __ldg(xs + (size_t)i * cs) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int i = i0 + u * nparts;
          if (i < p.Ci) {
            const uint32_t o0 = dw_off(2 * i, bb), o1 = dw_off(2 * i + 1, bb);
            const float hx = tf32_rna(v[u].x), hy = tf32_rna(v[u].y);
            *(float*)(s_ah + o0) = hx; *(float*)(s_ah + o1) = hy;
            *(float*)(s_al + o0) = tf32_rna(v[u].x - hx); *(float*)(s_al + o1) = tf32_rna(v[u].y - hy);
          }
        }
      }
      for (int o0c = part; o0c < p.Co; o0c += 8 * nparts) {
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int o = o0c + u * nparts;
          v[u] = (live && o < p.Co) ? __ldg(gs + (size_t)o * cs) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int o = o0c + u * nparts;
          if (o < p.Co) {
            const uint32_t o0 = dw_off(2 * o, bb), o1 = dw_off(2 * o + 1, bb);
            const float hx = tf32_rna(v[u].x), hy = tf32_rna(v[u].y);
            *(float*)(s_bh + o0) = hx; *(float*)(s_bh + o1) = hy;
            *(float*)(s_bl + o0) = tf32_rna(v[u].x - hx); *(float*)(s_bl + o1) = tf32_rna(v[u].y - hy);
          }
        }
      }
      fence_proxy_async();
      __syncthreads();
      if (warp == 0) {
        if (elect_one()) {
          _Pragma("unroll") for (int pass = 0; pass < 3; pass++) {
            if (pass >= p.npass) break;   // compile-time trip count (descriptors stay folded); 1-pass mode leaves early
            const uint32_t aoff = pass == 1 ? abytes : 0u;                       // A: hi, lo, hi
            const uint32_t boff = 2 * abytes + (pass == 2 ? bbytes : 0u);        // B: hi, hi, lo
            const uint64_t da = smem_desc(sbase + aoff, kDwLbo, kDwSbo, LAYOUT_NONE);
            const uint64_t db = smem_desc(sbase + boff, kDwLbo, kDwSbo, LAYOUT_NONE);
            for (int ks = 0; ks < kDwKB / 8; ks++) {
              mma_tf32_ss(tbase, da + (uint64_t)(ks * (2 * kDwLbo / 16)), db + (uint64_t)(ks * (2 * kDwLbo / 16)), idesc, acc);
              acc = 1;
            }
          }
          mma_commit(&bar);
        }
        __syncwarp();
      }
      pending = true;
    }
    // ---- read-out of the mode: combine the re / im rows of each input channel (adjacent lanes) ----
    mbar_wait(&bar, phase);
    phase ^= 1u;
    pending = false;
    tc_fence_after();
    {
      int corner;
      long woff;
      decode_mode(p.mm, p.w, k, &corner, &woff);
      float2* wb = (float2*)p.w.corner[corner] + woff;
      const int i = m >> 1;
      for (int c0 = grp * 8; c0 < Np; c0 += 16) {
        float v[8], pv[8];
        tmem_ld8(tbase + lane_base + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; j++) pv[j] = __shfl_xor_sync(0xffffffffu, v[j], 1);
        if (!(m & 1) && i < p.Ci) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int o = (c0 >> 1) + j;
            if (o < p.Co) {
              // v = sums with Re X (row 2i), pv = sums with Im X (row 2i+1):  dW = conj(X) G
              float2 r = make_float2(v[2 * j] + pv[2 * j + 1], v[2 * j + 1] - pv[2 * j]);
              float2* dst = wb + (long)i * p.w.stride_i + (long)o * p.w.stride_o;
              if (p.accumulate) { const float2 old = *dst; r.x += old.x; r.y += old.y; }
              *dst = r;
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();         // accumulator drained before the next mode's first MMA overwrites it
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, ncols);
}

int g_mix_min_batch = 96;    // below this a 128-sample tile is mostly padding: the CUDA-core kernels win

}  // namespace

// shape checks shared by the launchers and b2no_plan_layout_supported
int b2no_tc_mix_feasible(int batch, int ci, int co, int mode) {
  if (!b2no_tc_available() || batch < g_mix_min_batch) return 0;
  const int Cq = mode == 0 ? ci : co, Cp = mode == 0 ? co : ci;
  const int Kp = b2no_round_up(2 * Cq, 16), Np = b2no_round_up(2 * Cp, 16);
  if (Np > 256 || 2 * Kp + Np > 512) return 0;
  return 2 * (size_t)Np * Kp * 4 + 1024 <= 220 * 1024;
}
int b2no_tc_mix_dw_feasible(int batch, int ci, int co) {
  if (!b2no_tc_available() || batch < g_mix_min_batch) return 0;
  return 2 * ci <= 128 && b2no_round_up(2 * co, 16) <= 256;
}

int b2no_tc_mix(const b2no_plan* p, int mode, const float* in, const b2no_weights* w, float* out, int batch, int ci, int co,
                int accumulate, cudaStream_t st) {
  if (!b2no_tc_available() || batch < g_mix_min_batch) return 1;
  MixTc q;
  memset(&q, 0, sizeof(q));
  q.layout = p->g.spec_layout;
  q.B = batch; q.Cq = mode == 0 ? ci : co; q.Cp = mode == 0 ? co : ci; q.Kt = total_modes(p);
  q.Kp = b2no_round_up(2 * q.Cq, 16); q.Np = b2no_round_up(2 * q.Cp, 16);
  q.conjt = mode; q.accumulate = accumulate; q.npass = b2no_tc_passes();
  q.in = (const float2*)in; q.out = (float2*)out; q.w = *w; q.mm = make_mode_map(p);
  if (q.Np > 256 || 2 * q.Kp + q.Np > 512) return 1;
  const size_t smem = 2 * (size_t)q.Np * q.Kp * 4 + 1024;
  int dev = 0, max_smem = 0;
  B2NO_CHECK_CUDA(cudaGetDevice(&dev));
  B2NO_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if ((int)smem > max_smem) return 1;
  B2NO_CHECK_CUDA(cudaFuncSetAttribute(k_mix_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // two CTAs per SM when TMEM (<= 256 columns each) and shared memory allow: the kernel is a chain of gather latencies
  const int per_sm = (2 * q.Kp + q.Np <= 256 && smem <= 100 * 1024) ? 2 : 1;
  const long items = (long)q.Kt * ((batch + 127) / 128);
  long grid = (long)b2no_sm_count() * per_sm;
  if (grid > items) grid = items;
  // residency = TMEM budget: ask for enough shared memory that no more than per_sm CTAs fit an SM (a third CTA would sit
  // in tcgen05.alloc holding its slot)
  size_t smem_req = smem;
  const size_t floor_req = per_sm == 2 ? 100 * 1024 : 120 * 1024;
  if (smem_req < floor_req) smem_req = floor_req;
  B2NO_CHECK_CUDA(cudaFuncSetAttribute(k_mix_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_req));
  k_mix_tc<<<(unsigned)grid, kMixThreads, smem_req, st>>>(q);
  B2NO_LAUNCH_CHECK();
  b2no_tc_count_launch();
  return 0;
}

int b2no_tc_mix_dw(const b2no_plan* p, const float* xh, const float* gyh, const b2no_weights* dw, int batch, int ci, int co,
                   int accumulate, cudaStream_t st) {
  if (!b2no_tc_available() || batch < g_mix_min_batch) return 1;
  DwTc q;
  memset(&q, 0, sizeof(q));
  q.B = batch; q.Ci = ci; q.Co = co; q.Kt = total_modes(p); q.Np = b2no_round_up(2 * co, 16);
  q.accumulate = accumulate; q.npass = b2no_tc_passes(); q.layout = p->g.spec_layout;
  q.xh = (const float2*)xh; q.gyh = (const float2*)gyh; q.w = *dw; q.mm = make_mode_map(p);
  if (2 * ci > 128 || q.Np > 256) return 1;
  q.Ga = (2 * ci + 7) / 8;
  // the M = 128 MMA reads 16 row groups from the A base: keep them inside the allocation
  size_t smem = 2 * (size_t)q.Ga * kDwSbo + 2 * (size_t)(q.Np / 8) * kDwSbo;
  if (smem < 2 * (size_t)q.Ga * kDwSbo + 16 * (size_t)kDwSbo) smem = 2 * (size_t)q.Ga * kDwSbo + 16 * (size_t)kDwSbo;
  smem += 1024;
  int dev = 0, max_smem = 0;
  B2NO_CHECK_CUDA(cudaGetDevice(&dev));
  B2NO_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if ((int)smem > max_smem) return 1;
  B2NO_CHECK_CUDA(cudaFuncSetAttribute(k_dw_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_sm = (q.Np <= 256 && smem <= 100 * 1024) ? 2 : 1;
  int grid = b2no_sm_count() * per_sm;
  if (grid > q.Kt) grid = q.Kt;
  size_t smem_req = smem;
  const size_t floor_req = per_sm == 2 ? 100 * 1024 : 120 * 1024;
  if (smem_req < floor_req) smem_req = floor_req;
  B2NO_CHECK_CUDA(cudaFuncSetAttribute(k_dw_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_req));
  k_dw_tc<<<grid, kMixThreads, smem_req, st>>>(q);
  B2NO_LAUNCH_CHECK();
  b2no_tc_count_launch();
  return 0;
}
