// Fused Adam over ONE flat fp32 parameter / gradient buffer (the layout pde_policylearning_b200.parallel.GradBucket
// already gives the gradients; complex parameters are their (re, im) real view, exactly how torch.optim.Adam treats
// them).  Semantics of torch.optim.Adam as the reference configures it (run_pde_observers.py:134 Adam(lr, weight_decay),
// train_pino.py:205): L2 weight decay folded into the gradient, bias-corrected moments, eps outside the sqrt.
//   g   = grad * grad_scale + wd * p            (grad_scale = 1/world_size folds the all-reduce mean)
//   m   = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2
//   p  -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The step counter lives on the device so the whole training step is CUDA-graph capturable: the kernel reads t from
// `step` (already incremented by k_adam_tick on the same stream).  HBM-bound: 4 reads + 3 writes of 4 B per element.
#include <math.h>
#include "common.cuh"

__global__ void k_adam_tick(int* step) { *step += 1; }

__global__ void __launch_bounds__(256)
k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
       const int* __restrict__ step, float lr, float b1, float b2, float eps, float wd, float gscale) {
  const int t = *step;
  const float bc1 = 1.0f - powf(b1, (float)t);
  const float bc2s = sqrtf(1.0f - powf(b2, (float)t));
  const float step_size = lr / bc1;
  const long n4 = n >> 2;
  const long stride = (long)gridDim.x * blockDim.x;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg = fmaf(wd, pp, gg * gscale);
    mm = fmaf(b1, mm, (1.0f - b1) * gg);
    vv = fmaf(b2, vv, (1.0f - b2) * gg * gg);
    pp -= step_size * mm / (sqrtf(vv) / bc2s + eps);
  };
  for (long i = tid; i < n4; i += stride) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    upd(pp.x, gg.x, mm.x, vv.x);
    upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z);
    upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (long i = (n4 << 2) + tid; i < n; i += stride) upd(p[i], g[i], m[i], v[i]);
}

extern "C" int b2no_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                              int* step_counter, float lr, float beta1, float beta2, float eps, float weight_decay,
                              float grad_scale, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step_counter || n < 0) return B2NO_E_ARG;
  if (n == 0) return 0;
  if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) return B2NO_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  k_adam_tick<<<1, 1, 0, st>>>(step_counter);
  B2NO_LAUNCH_CHECK();
  long blocks = ((n + 3) / 4 + 255) / 256;
  const long cap = (long)b2no_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_adam<<<(unsigned)blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, step_counter, lr, beta1, beta2, eps,
                                           weight_decay, grad_scale);
  B2NO_LAUNCH_CHECK();
  return 0;
}

// Gathers per-parameter gradient tensors into the flat bucket: flat[offsets[s] .. offsets[s+1]) = src[s][0 ..) (a null
// src[s], or the part of a slot beyond counts[s], is zero-filled).  Replaces one torch accumulate kernel per parameter
// (autograd's `p.grad += g`) by ONE launch; the pointer table is constant inside a captured CUDA graph.
__global__ void __launch_bounds__(256)
k_gather_segments(float* __restrict__ flat, const float* const* __restrict__ src, const long long* __restrict__ offsets,
                  const long long* __restrict__ counts, int nseg) {
  for (int s = blockIdx.y; s < nseg; s += gridDim.y) {
    const long long o0 = offsets[s], n = offsets[s + 1] - o0, cnt = counts[s];
    const float* sp = src[s];
    float* dp = flat + o0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      dp[i] = (sp && i < cnt) ? sp[i] : 0.f;
  }
}

extern "C" int b2no_gather_segments(float* flat, const void* src_ptrs, const int64_t* offsets, const int64_t* counts,
                                    int nseg, void* stream) {
  if (!flat || !src_ptrs || !offsets || !counts || nseg < 0) return B2NO_E_ARG;
  if (nseg == 0) return 0;
  dim3 grid(32, (unsigned)(nseg < 1024 ? nseg : 1024));
  k_gather_segments<<<grid, 256, 0, (cudaStream_t)stream>>>(flat, (const float* const*)src_ptrs, (const long long*)offsets,
                                                           (const long long*)counts, nseg);
  B2NO_LAUNCH_CHECK();
  return 0;
}
