// fp32 CUDA-core kernels around the encoder:
//   * policy / baseline heads forward + backward (atari_model.py:104-107,126-127 and their autograd)
//   * bias-gradient column sums over bf16 dY
//   * clip_grad_norm_ (impala_atari.py:344-345) + RMSprop (impala_atari.py:99-105,346) / Adam update
#include "common.cuh"
#include "kernels.h"

namespace srl {

// ------------------------------------------------------------------------------------------------
// heads forward: one warp per frame.  core = [h(512), clamp(reward,-1,1), one_hot(action)(A)]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) head_fwd_kernel(const float* __restrict__ hpart, int nsplit, const float* __restrict__ bfc,
                                                       float* __restrict__ h, const float* __restrict__ reward,
                                                       const int64_t* __restrict__ action, const float* __restrict__ Wp,
                                                       const float* __restrict__ bp, const float* __restrict__ Wb,
                                                       const float* __restrict__ bb, int N, int A, float* __restrict__ logits,
                                                       float* __restrict__ baseline) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const int CORE = 513 + A;
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {     // fc epilogue: reduce the split-K partials in fixed order, + bias, ReLU (atari_model.py:100-101)
    const int j = i * 32 + lane;
    float s = __ldg(hpart + (size_t)n * 512 + j);
    for (int k = 1; k < nsplit; ++k) s += __ldg(hpart + ((size_t)k * N + n) * 512 + j);
    x[i] = fmaxf(s + __ldg(bfc + j), 0.f);
    h[(size_t)n * 512 + j] = x[i];
  }
  const float r = fminf(fmaxf(__ldg(reward + n), -1.f), 1.f);
  const int act = (int)__ldg(action + n);
  for (int a = 0; a <= A; ++a) {
    const float* w = a < A ? Wp + (size_t)a * CORE : Wb;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s = fmaf(x[i], __ldg(w + i * 32 + lane), s);
    s = warp_sum(s);
    if (lane == 0) {
      s += __ldg(w + 512) * r + __ldg(w + 513 + act) + (a < A ? __ldg(bp + a) : __ldg(bb));
      if (a < A) logits[(size_t)n * A + a] = s; else baseline[n] = s;
    }
  }
}

// dh[n][j] = (sum_a dlogits[n][a] Wp[a][j] + dV[n] Wb[j]) * (h[n][j] > 0)  -> bf16 (operand of the fc dgrad/wgrad GEMMs)
__global__ void __launch_bounds__(128) head_bwd_dh_kernel(const float* __restrict__ dlogits, const float* __restrict__ dbaseline,
                                                          const float* __restrict__ h, const float* __restrict__ Wp,
                                                          const float* __restrict__ Wb, int N, int A, __nv_bfloat16* __restrict__ dh) {
  const int n = blockIdx.x;
  const int j = blockIdx.y * 128 + threadIdx.x;
  const int CORE = 513 + A;
  float s = __ldg(dbaseline + n) * __ldg(Wb + j);
  for (int a = 0; a < A; ++a) s = fmaf(__ldg(dlogits + (size_t)n * A + a), __ldg(Wp + (size_t)a * CORE + j), s);
  if (!(__ldg(h + (size_t)n * 512 + j) > 0.f)) s = 0.f;
  dh[(size_t)n * 512 + j] = __float2bfloat16_rn(s);
}

// head weight/bias gradients: thread = one column j of `core` (j == CORE is the bias "ones" column),
// blockIdx.y = slab of frames; accumulates A+1 outputs and adds them atomically into the pre-zeroed gradient.
constexpr int HEAD_MAX_A = 32;
__global__ void __launch_bounds__(128) head_wgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ dbaseline,
                                                         const float* __restrict__ h, const float* __restrict__ reward,
                                                         const int64_t* __restrict__ action, int N, int A, int rows_per_slab,
                                                         float* __restrict__ gWp, float* __restrict__ gbp, float* __restrict__ gWb,
                                                         float* __restrict__ gbb) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int CORE = 513 + A;
  if (j > CORE) return;
  const int n0 = blockIdx.y * rows_per_slab, n1 = min(N, n0 + rows_per_slab);
  float acc[HEAD_MAX_A + 1];
#pragma unroll
  for (int a = 0; a <= HEAD_MAX_A; ++a) acc[a] = 0.f;
  for (int n = n0; n < n1; ++n) {
    float c;
    if (j < 512) c = __ldg(h + (size_t)n * 512 + j);
    else if (j == 512) c = fminf(fmaxf(__ldg(reward + n), -1.f), 1.f);
    else if (j < CORE) c = ((int)__ldg(action + n) == j - 513) ? 1.f : 0.f;
    else c = 1.f;
    if (c != 0.f) {
#pragma unroll
      for (int a = 0; a < HEAD_MAX_A; ++a)
        if (a < A) acc[a] = fmaf(__ldg(dlogits + (size_t)n * A + a), c, acc[a]);
      acc[HEAD_MAX_A] = fmaf(__ldg(dbaseline + n), c, acc[HEAD_MAX_A]);
    }
  }
#pragma unroll
  for (int a = 0; a < HEAD_MAX_A; ++a)
    if (a < A) {
      if (j < CORE) atomicAdd(gWp + (size_t)a * CORE + j, acc[a]); else atomicAdd(gbp + a, acc[a]);
    }
  if (j < CORE) atomicAdd(gWb + j, acc[HEAD_MAX_A]); else atomicAdd(gbb, acc[HEAD_MAX_A]);
}

// ------------------------------------------------------------------------------------------------
// optimizer
// ------------------------------------------------------------------------------------------------
// coef[0] = ||g||_2 ; coef[1] = min(1, max_norm / (||g|| + 1e-6))     (torch.nn.utils.clip_grad_norm_)
// scratch: [0] ticket (uint), [4 .. 4+grid) block partials
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, int64_t n, float max_norm, float* __restrict__ coef,
                                                         float* __restrict__ scratch) {
  float s = 0.f;
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[n4 * 4 + threadIdx.x]; s += v * v; }
  __shared__ float red[8];
  __shared__ bool is_last;
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    scratch[4 + blockIdx.x] = t;
    __threadfence();
    is_last = atomicAdd(reinterpret_cast<unsigned*>(scratch), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {   // fixed-order parallel sum of the block partials (deterministic)
    __threadfence();
    double t = 0.0;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += 32) t += (double)reinterpret_cast<volatile float*>(scratch)[4 + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) {
      const float norm = (float)sqrt(t);
      coef[0] = norm;
      coef[1] = max_norm >= 0.f ? fminf(max_norm / (norm + 1e-6f), 1.0f) : 1.0f;
      *reinterpret_cast<unsigned*>(scratch) = 0u;
    }
  }
}

// torch.optim.RMSprop(momentum=0, centered=False): v = alpha v + (1-alpha) g^2 ; p -= lr g / (sqrt(v) + eps); g pre-scaled by coef[1]
__global__ void __launch_bounds__(256) rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v, int64_t n,
                                                      const float* __restrict__ coef, float lr, float alpha, float eps) {
  const float c = coef ? __ldg(coef + 1) : 1.0f;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float* P = &pp.x; float* V = &vv.x; const float* G = &gg.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = G[k] * c;
      V[k] = alpha * V[k] + (1.f - alpha) * gk * gk;
      P[k] = P[k] - lr * (gk / (sqrtf(V[k]) + eps));
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = n4 * 4 + threadIdx.x;
    const float gk = g[i] * c;
    v[i] = alpha * v[i] + (1.f - alpha) * gk * gk;
    p[i] = p[i] - lr * (gk / (sqrtf(v[i]) + eps));
  }
}

// torch.optim.Adam: m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr/bc1) m / (sqrt(v)/sqrt(bc2) + eps)
// The 1-based step count is read from device memory (dstep, incremented by adam_step_inc_kernel) so that a captured
// CUDA graph replays with the right bias correction; dstep == nullptr uses the host-provided `step`.
__global__ void adam_step_inc_kernel(int* dstep) { *dstep += 1; }
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, const float* __restrict__ coef, float lr, float b1,
                                                   float b2, float eps, int step, const int* __restrict__ dstep) {
  const float c = coef ? __ldg(coef + 1) : 1.0f;
  const int t = dstep ? *dstep : step;
  const float inv_bc1 = 1.0f / (float)(1.0 - pow((double)b1, (double)t));
  const float inv_sqrt_bc2 = 1.0f / sqrtf((float)(1.0 - pow((double)b2, (double)t)));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gk = g[i] * c;
    const float mk = b1 * m[i] + (1.f - b1) * gk;
    const float vk = b2 * v[i] + (1.f - b2) * gk * gk;
    m[i] = mk; v[i] = vk;
    p[i] = p[i] - (lr * inv_bc1) * (mk / (sqrtf(vk) * inv_sqrt_bc2 + eps));
  }
}

// ------------------------------------------------------------------------------------------------
cudaError_t launch_head_fwd(const float* hpart, int nsplit, const float* bfc, float* h, const float* reward, const int64_t* action,
                            const float* Wp, const float* bp, const float* Wb, const float* bb, int N, int A, float* logits,
                            float* baseline, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  head_fwd_kernel<<<(N + 3) / 4, 128, 0, st>>>(hpart, nsplit, bfc, h, reward, action, Wp, bp, Wb, bb, N, A, logits, baseline);
  return cudaGetLastError();
}
cudaError_t launch_head_bwd(const float* dlogits, const float* dbaseline, const float* h, const float* reward, const int64_t* action,
                            const float* Wp, const float* Wb, int N, int A, __nv_bfloat16* dh, float* gWp, float* gbp, float* gWb,
                            float* gbb, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  head_bwd_dh_kernel<<<dim3(N, 4), 128, 0, st>>>(dlogits, dbaseline, h, Wp, Wb, N, A, dh);
  const int CORE = 513 + A;
  const int slabs = 32, rps = (N + slabs - 1) / slabs;
  head_wgrad_kernel<<<dim3((CORE + 1 + 127) / 128, slabs), 128, 0, st>>>(dlogits, dbaseline, h, reward, action, N, A, rps, gWp, gbp, gWb, gbb);
  return cudaGetLastError();
}
cudaError_t launch_grad_norm(const float* g, int64_t n, float max_norm, float* coef, float* scratch, cudaStream_t st) {
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 592) blocks = 592;
  if (blocks < 1) blocks = 1;
  grad_sumsq_kernel<<<blocks, 256, 0, st>>>(g, n, max_norm, coef, scratch);
  return cudaGetLastError();
}
static int ew_blocks(int64_t n) { int64_t b = (n / 4 + 255) / 256; return (int)(b < 1 ? 1 : (b > 1184 ? 1184 : b)); }
cudaError_t launch_rmsprop(float* p, const float* g, float* v, int64_t n, const float* coef, float lr, float alpha, float eps,
                           cudaStream_t st) {
  rmsprop_kernel<<<ew_blocks(n), 256, 0, st>>>(p, g, v, n, coef, lr, alpha, eps);
  return cudaGetLastError();
}
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* coef, float lr, float b1, float b2, float eps,
                        int step, int* dstep, cudaStream_t st) {
  if (dstep) adam_step_inc_kernel<<<1, 1, 0, st>>>(dstep);
  adam_kernel<<<ew_blocks(n * 4), 256, 0, st>>>(p, g, m, v, n, coef, lr, b1, b2, eps, step, dstep);
  return cudaGetLastError();
}

}  // namespace srl
