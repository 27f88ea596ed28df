// fp32 CUDA-core kernels around the encoder:
//   * policy / baseline heads forward + backward (atari_model.py:104-107,126-127 and their autograd)
//   * bias-gradient column sums over bf16 dY
//   * clip_grad_norm_ (impala_atari.py:344-345) + RMSprop (impala_atari.py:99-105,346) / Adam update
#include "common.cuh"
#include "kernels.h"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#ifndef SRL_TRY
#define SRL_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)
#endif

namespace srl {

constexpr int HEAD_MAX_A = 32;

// ------------------------------------------------------------------------------------------------
// heads forward: one warp per frame.  core = [h(512), clamp(reward,-1,1), one_hot(action)(A)]
// ------------------------------------------------------------------------------------------------
// Block = 256 threads = 2 frames x 4 warps; warp w of a frame owns features [128w, 128w+128), 4 per lane (float4 loads).
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ hpart, int nsplit, const float* __restrict__ bfc,
                                                       float* __restrict__ h, const float* __restrict__ reward,
                                                       const int64_t* __restrict__ action, const float* __restrict__ Wp,
                                                       const float* __restrict__ bp, const float* __restrict__ Wb,
                                                       const float* __restrict__ bb, int N, int A, float* __restrict__ logits,
                                                       float* __restrict__ baseline) {
  pdl_wait();      // launched with programmatic stream serialization: see common.cuh
  pdl_launch();
  __shared__ float part[2][4][HEAD_MAX_A + 1];
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 3, f = threadIdx.x >> 7;
  const int n = blockIdx.x * 2 + f;
  const int CORE = 513 + A;
  const int j = warp * 128 + lane * 4;
  if (n < N) {
    // fc epilogue: reduce the split-K partials in fixed order, + bias, ReLU (atari_model.py:100-101)
    float4 x = __ldg(reinterpret_cast<const float4*>(hpart + (size_t)n * 512 + j));
    for (int k = 1; k < nsplit; ++k) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(hpart + ((size_t)k * N + n) * 512 + j));
      x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w;
    }
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bfc + j));
    x.x = fmaxf(x.x + b4.x, 0.f); x.y = fmaxf(x.y + b4.y, 0.f); x.z = fmaxf(x.z + b4.z, 0.f); x.w = fmaxf(x.w + b4.w, 0.f);
    *reinterpret_cast<float4*>(h + (size_t)n * 512 + j) = x;
    for (int a = 0; a <= A; ++a) {
      const float* w = (a < A ? Wp + (size_t)a * CORE : Wb) + j;     // rows are not 16-byte aligned (CORE is odd): scalar loads
      float s = x.x * __ldg(w) + x.y * __ldg(w + 1) + x.z * __ldg(w + 2) + x.w * __ldg(w + 3);
      s = warp_sum(s);
      if (lane == 0) part[f][warp][a] = s;
    }
  }
  __syncthreads();
  if (n < N && warp == 0 && lane <= A) {
    const int a = lane;
    const float* w = a < A ? Wp + (size_t)a * CORE : Wb;
    const float r = fminf(fmaxf(__ldg(reward + n), -1.f), 1.f);
    const int act = (int)__ldg(action + n);
    float s = (part[f][0][a] + part[f][1][a]) + (part[f][2][a] + part[f][3][a]);
    s += __ldg(w + 512) * r + __ldg(w + 513 + act) + (a < A ? __ldg(bp + a) : __ldg(bb));
    if (a < A) logits[(size_t)n * A + a] = s; else baseline[n] = s;
  }
}

// dh[n][j] = (sum_a dlogits[n][a] Wp[a][j] + dV[n] Wb[j]) * (h[n][j] > 0)  -> bf16 (operand of the fc dgrad/wgrad GEMMs)
__global__ void __launch_bounds__(128) head_bwd_dh_kernel(const float* __restrict__ dlogits, const float* __restrict__ dbaseline,
                                                          const float* __restrict__ h, const float* __restrict__ Wp,
                                                          const float* __restrict__ Wb, int N, int A, __nv_bfloat16* __restrict__ dh) {
  pdl_wait();      // launched with programmatic stream serialization: see common.cuh
  pdl_launch();
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
constexpr int HEAD_SLAB = 16;   // frames per block
__global__ void __launch_bounds__(128) head_wgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ dbaseline,
                                                         const float* __restrict__ h, const float* __restrict__ reward,
                                                         const int64_t* __restrict__ action, int N, int A, int rows_per_slab,
                                                         float* __restrict__ gWp, float* __restrict__ gbp, float* __restrict__ gWb,
                                                         float* __restrict__ gbb) {
  __shared__ float sd[HEAD_SLAB][HEAD_MAX_A + 1];
  __shared__ float sr[HEAD_SLAB];
  __shared__ int sa[HEAD_SLAB];
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int CORE = 513 + A;
  const int n0 = blockIdx.y * HEAD_SLAB, cnt = min(HEAD_SLAB, N - n0);
  for (int i = threadIdx.x; i < HEAD_SLAB * (A + 1); i += 128) {   // rows past the ragged end are ZERO (0 * stale smem could be NaN)
    const int r = i / (A + 1), a = i - r * (A + 1);
    sd[r][a] = r < cnt ? (a < A ? __ldg(dlogits + (size_t)(n0 + r) * A + a) : __ldg(dbaseline + n0 + r)) : 0.f;
  }
  if (threadIdx.x < cnt) {
    sr[threadIdx.x] = fminf(fmaxf(__ldg(reward + n0 + threadIdx.x), -1.f), 1.f);
    sa[threadIdx.x] = (int)__ldg(action + n0 + threadIdx.x);
  }
  __syncthreads();
  if (j > CORE) return;
  float c[HEAD_SLAB];
#pragma unroll
  for (int r = 0; r < HEAD_SLAB; ++r) {       // all loads of the slab are independent: HEAD_SLAB requests in flight
    float v = 0.f;
    if (r < cnt) {
      if (j < 512) v = __ldg(h + (size_t)(n0 + r) * 512 + j);
      else if (j == 512) v = sr[r];
      else if (j < CORE) v = (sa[r] == j - 513) ? 1.f : 0.f;
      else v = 1.f;
    }
    c[r] = v;
  }
  for (int a = 0; a <= A; ++a) {
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < HEAD_SLAB; ++r) acc = fmaf(sd[r][a], c[r], acc);
    if (a < A) { if (j < CORE) atomicAdd(gWp + (size_t)a * CORE + j, acc); else atomicAdd(gbp + a, acc); }
    else       { if (j < CORE) atomicAdd(gWb + j, acc); else atomicAdd(gbb, acc); }
  }
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
// LSTM path (use_lstm): the heads read the LSTM output X [N][H] (H = 513 + A) instead of [h, reward, one-hot]
// ------------------------------------------------------------------------------------------------
// core[n] = [relu(sum_s hpart + bfc) (512), clamp(reward,-1,1), one_hot(action) (A)]  (atari_model.py:100-107); also stores h
__global__ void __launch_bounds__(128) core_build_kernel(const float* __restrict__ hpart, int nsplit, const float* __restrict__ bfc,
                                                         const float* __restrict__ reward, const int64_t* __restrict__ action, int N, int A,
                                                         float* __restrict__ h, float* __restrict__ core) {
  const int n = blockIdx.x, H = 513 + A;
  for (int j = threadIdx.x; j < H; j += 128) {
    float v;
    if (j < 512) {
      v = __ldg(hpart + (size_t)n * 512 + j);
      for (int k = 1; k < nsplit; ++k) v += __ldg(hpart + ((size_t)k * N + n) * 512 + j);
      v = fmaxf(v + __ldg(bfc + j), 0.f);
      h[(size_t)n * 512 + j] = v;
    } else if (j == 512) {
      v = fminf(fmaxf(__ldg(reward + n), -1.f), 1.f);
    } else {
      v = ((int)__ldg(action + n) == j - 513) ? 1.f : 0.f;
    }
    core[(size_t)n * H + j] = v;
  }
}
// logits[n][a] = X[n] . Wp[a] + bp[a];  baseline[n] = X[n] . Wb + bb      (one warp per frame)
__global__ void __launch_bounds__(256) head_dense_fwd_kernel(const float* __restrict__ X, const float* __restrict__ Wp, const float* __restrict__ bp,
                                                             const float* __restrict__ Wb, const float* __restrict__ bb, int N, int A,
                                                             float* __restrict__ logits, float* __restrict__ baseline) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  const int H = 513 + A;
  for (int a = 0; a <= A; ++a) {
    const float* w = a < A ? Wp + (size_t)a * H : Wb;
    float s = 0.f;
    for (int j = lane; j < H; j += 32) s = fmaf(__ldg(X + (size_t)n * H + j), __ldg(w + j), s);
    s = warp_sum(s);
    if (lane == 0) { if (a < A) logits[(size_t)n * A + a] = s + __ldg(bp + a); else baseline[n] = s + __ldg(bb); }
  }
}
// dX[n][j] = sum_a dlogits[n][a] Wp[a][j] + dV[n] Wb[j];  head weight/bias gradients accumulated atomically (slabs of 16 frames)
__global__ void __launch_bounds__(128) head_dense_bwd_kernel(const float* __restrict__ X, const float* __restrict__ dlogits,
                                                             const float* __restrict__ dbaseline, const float* __restrict__ Wp,
                                                             const float* __restrict__ Wb, int N, int A, float* __restrict__ dX,
                                                             float* __restrict__ gWp, float* __restrict__ gbp, float* __restrict__ gWb,
                                                             float* __restrict__ gbb) {
  __shared__ float sd[16][HEAD_MAX_A + 1];
  const int H = 513 + A, j = blockIdx.x * 128 + threadIdx.x;
  const int n0 = blockIdx.y * 16, cnt = min(16, N - n0);
  for (int i = threadIdx.x; i < HEAD_SLAB * (A + 1); i += 128) {   // rows past the ragged end are ZERO (0 * stale smem could be NaN)
    const int r = i / (A + 1), a = i - r * (A + 1);
    sd[r][a] = r < cnt ? (a < A ? __ldg(dlogits + (size_t)(n0 + r) * A + a) : __ldg(dbaseline + n0 + r)) : 0.f;
  }
  __syncthreads();
  if (j > H) return;
  float acc[HEAD_MAX_A + 1];
#pragma unroll
  for (int a = 0; a <= HEAD_MAX_A; ++a) acc[a] = 0.f;
  for (int r = 0; r < cnt; ++r) {
    const float x = j < H ? __ldg(X + (size_t)(n0 + r) * H + j) : 1.f;     // j == H: the bias "ones" column
    float dx = 0.f;
#pragma unroll
    for (int a = 0; a < HEAD_MAX_A; ++a)
      if (a < A) { acc[a] = fmaf(sd[r][a], x, acc[a]); if (j < H) dx = fmaf(sd[r][a], __ldg(Wp + (size_t)a * H + j), dx); }
    acc[HEAD_MAX_A] = fmaf(sd[r][A], x, acc[HEAD_MAX_A]);
    if (j < H) dX[(size_t)(n0 + r) * H + j] = dx + sd[r][A] * __ldg(Wb + j);
  }
#pragma unroll
  for (int a = 0; a < HEAD_MAX_A; ++a)
    if (a < A) { if (j < H) atomicAdd(gWp + (size_t)a * H + j, acc[a]); else atomicAdd(gbp + a, acc[a]); }
  if (j < H) atomicAdd(gWb + j, acc[HEAD_MAX_A]); else atomicAdd(gbb, acc[HEAD_MAX_A]);
}
// dh[n][j] = bf16(dcore[n][j] * (h[n][j] > 0)), j < 512 (the reward / one-hot columns of core have no parameters below them)
__global__ void __launch_bounds__(128) dcore_to_dh_kernel(const float* __restrict__ dcore, const float* __restrict__ h, int A,
                                                          __nv_bfloat16* __restrict__ dh) {
  const int n = blockIdx.x, H = 513 + A;
  for (int j = threadIdx.x; j < 512; j += 128) {
    const float v = __ldg(h + (size_t)n * 512 + j) > 0.f ? __ldg(dcore + (size_t)n * H + j) : 0.f;
    dh[(size_t)n * 512 + j] = __float2bfloat16_rn(v);
  }
}

cudaError_t launch_core_build(const float* hpart, int nsplit, const float* bfc, const float* reward, const int64_t* action, int N, int A, float* h,
                              float* core, cudaStream_t st) {
  core_build_kernel<<<N, 128, 0, st>>>(hpart, nsplit, bfc, reward, action, N, A, h, core);
  return cudaGetLastError();
}
cudaError_t launch_head_dense_fwd(const float* X, const float* Wp, const float* bp, const float* Wb, const float* bb, int N, int A, float* logits,
                                  float* baseline, cudaStream_t st) {
  head_dense_fwd_kernel<<<(N + 7) / 8, 256, 0, st>>>(X, Wp, bp, Wb, bb, N, A, logits, baseline);
  return cudaGetLastError();
}
cudaError_t launch_head_dense_bwd(const float* X, const float* dlogits, const float* dbaseline, const float* Wp, const float* Wb, int N, int A,
                                  float* dX, float* gWp, float* gbp, float* gWb, float* gbb, cudaStream_t st) {
  const int H = 513 + A;
  head_dense_bwd_kernel<<<dim3((H + 1 + 127) / 128, (N + 15) / 16), 128, 0, st>>>(X, dlogits, dbaseline, Wp, Wb, N, A, dX, gWp, gbp, gWb, gbb);
  return cudaGetLastError();
}
cudaError_t launch_dcore_to_dh(const float* dcore, const float* h, int N, int A, __nv_bfloat16* dh, cudaStream_t st) {
  dcore_to_dh_kernel<<<N, 128, 0, st>>>(dcore, h, A, dh);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// trajectory-slot unpack: B slots (one contiguous record per actor rollout, all keys of create_buffers,
// impala_atari.py:135-147) copied host->device as they lie, then scattered into the time-major [T+1, B, ...] batch.
// grid = (T+1, B): one block moves one 28,224-byte frame with 16-byte vectors; thread 0 moves the scalars.
// ------------------------------------------------------------------------------------------------
struct SlotOffsets { int64_t obs, reward, done, action, policy_logits, episode_return; };
__global__ void __launch_bounds__(256) unpack_slots_kernel(const uint8_t* __restrict__ staging, int64_t slot_bytes, SlotOffsets o, int B, int A,
                                                           uint8_t* __restrict__ obs, float* __restrict__ reward, uint8_t* __restrict__ done,
                                                           int64_t* __restrict__ action, float* __restrict__ logits,
                                                           float* __restrict__ episode_return) {
  const int t = blockIdx.x, b = blockIdx.y;
  const uint8_t* slot = staging + (size_t)b * slot_bytes;
  const uint4* src = reinterpret_cast<const uint4*>(slot + o.obs + (size_t)t * 28224);
  uint4* dst = reinterpret_cast<uint4*>(obs + ((size_t)t * B + b) * 28224);
  for (int i = threadIdx.x; i < 1764; i += 256) dst[i] = __ldg(src + i);
  const size_t n = (size_t)t * B + b;
  if (threadIdx.x == 0) {
    reward[n] = reinterpret_cast<const float*>(slot + o.reward)[t];
    done[n] = (slot + o.done)[t];
    action[n] = reinterpret_cast<const int64_t*>(slot + o.action)[t];
    if (episode_return) episode_return[n] = reinterpret_cast<const float*>(slot + o.episode_return)[t];
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + A) logits[n * A + threadIdx.x - 32] = reinterpret_cast<const float*>(slot + o.policy_logits)[t * A + threadIdx.x - 32];
}

cudaError_t launch_unpack_slots(const uint8_t* staging, int64_t slot_bytes, const int64_t* off6, int T, int B, int A, uint8_t* obs, float* reward,
                                uint8_t* done, int64_t* action, float* logits, float* episode_return, cudaStream_t st) {
  SlotOffsets o{off6[0], off6[1], off6[2], off6[3], off6[4], off6[5]};
  unpack_slots_kernel<<<dim3(T + 1, B), 256, 0, st>>>(staging, slot_bytes, o, B, A, obs, reward, done, action, logits, episode_return);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
cudaError_t launch_head_fwd(const float* hpart, int nsplit, const float* bfc, float* h, const float* reward, const int64_t* action,
                            const float* Wp, const float* bp, const float* Wb, const float* bb, int N, int A, float* logits,
                            float* baseline, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  return launch_chain<PDL_SIMT>(head_fwd_kernel, dim3((N + 1) / 2), dim3(256), 0, st, hpart, nsplit, bfc, h, reward, action, Wp, bp, Wb, bb, N, A, logits, baseline);
  return cudaGetLastError();
}
cudaError_t launch_head_bwd(const float* dlogits, const float* dbaseline, const float* h, const float* reward, const int64_t* action,
                            const float* Wp, const float* Wb, int N, int A, __nv_bfloat16* dh, float* gWp, float* gbp, float* gWb,
                            float* gbb, cudaStream_t st, cudaStream_t st_wgrad, bool do_dh) {
  if (N <= 0) return cudaSuccess;
  if (do_dh) SRL_TRY(launch_chain<PDL_SIMT>(head_bwd_dh_kernel, dim3(N, 4), dim3(128), 0, st, dlogits, dbaseline, h, Wp, Wb, N, A, dh));
  const int CORE = 513 + A;
  // the head weight gradients only feed the optimizer: they may run on a side stream (st_wgrad) beside the fc backward
  head_wgrad_kernel<<<dim3((CORE + 1 + 127) / 128, (N + HEAD_SLAB - 1) / HEAD_SLAB), 128, 0, st_wgrad>>>(dlogits, dbaseline, h, reward, action, N, A,
                                                                                                             HEAD_SLAB, gWp, gbp, gWb, gbb);
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

// ------------------------------------------------------------------------------------------------
// clip_grad_norm_ + optimizer step as ONE cooperative kernel (impala_atari.py:344-346): phase 1 sums g^2 (block partials
// in a fixed slot each), grid barrier, every block adds the partials in the same fixed order (deterministic, identical
// in all blocks), phase 2 applies the clipped update (g is re-read from L2).  OPT 0 = RMSprop, 1 = Adam.
// ------------------------------------------------------------------------------------------------
template <int OPT>
__global__ void __launch_bounds__(512) clip_optim_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ s0,
                                                         float* __restrict__ s1, int64_t n, float max_norm, float* __restrict__ coef,
                                                         float* __restrict__ scratch, float lr, float a, float b, float eps, int step,
                                                         int* __restrict__ dstep) {
  cg::grid_group grid = cg::this_grid();
  const int64_t n4 = n >> 2, stride = (int64_t)gridDim.x * blockDim.x, i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int t = OPT == 1 ? (dstep ? *dstep + 1 : step) : 0;
  float s = 0.f;
  for (int64_t i = i0; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[n4 * 4 + threadIdx.x]; s += v * v; }
  __shared__ float red[16];
  __shared__ float c_sh;
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tsum = 0.f;
    for (int w = 0; w < 16; ++w) tsum += red[w];
    scratch[4 + blockIdx.x] = tsum;
  }
  grid.sync();
  if (threadIdx.x < 32) {
    double tsum = 0.0;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += 32) tsum += (double)__ldcg(scratch + 4 + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
    if (threadIdx.x == 0) {
      const float norm = (float)sqrt(tsum);
      const float c = max_norm >= 0.f ? fminf(max_norm / (norm + 1e-6f), 1.0f) : 1.0f;
      c_sh = c;
      if (blockIdx.x == 0) {
        coef[0] = norm; coef[1] = c;
        if (OPT == 1 && dstep) *dstep = t;
      }
    }
  }
  __syncthreads();
  const float c = c_sh;
  if (OPT == 0) {
    for (int64_t i = i0; i < n4; i += stride) {
      float4 pp = reinterpret_cast<float4*>(p)[i], vv = reinterpret_cast<float4*>(s0)[i];
      const float4 gg = reinterpret_cast<const float4*>(g)[i];
      float* P = &pp.x; float* V = &vv.x; const float* G = &gg.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = G[k] * c;
        V[k] = a * V[k] + (1.f - a) * gk * gk;
        P[k] = P[k] - lr * (gk / (sqrtf(V[k]) + eps));
      }
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(s0)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
      const int64_t i = n4 * 4 + threadIdx.x;
      const float gk = g[i] * c;
      s0[i] = a * s0[i] + (1.f - a) * gk * gk;
      p[i] = p[i] - lr * (gk / (sqrtf(s0[i]) + eps));
    }
  } else {
    const float inv_bc1 = 1.0f / (float)(1.0 - pow((double)a, (double)t));
    const float inv_sqrt_bc2 = 1.0f / sqrtf((float)(1.0 - pow((double)b, (double)t)));
    for (int64_t i = i0; i < n; i += stride) {
      const float gk = g[i] * c;
      const float mk = a * s0[i] + (1.f - a) * gk;
      const float vk = b * s1[i] + (1.f - b) * gk * gk;
      s0[i] = mk; s1[i] = vk;
      p[i] = p[i] - (lr * inv_bc1) * (mk / (sqrtf(vk) * inv_sqrt_bc2 + eps));
    }
  }
}

template <int OPT>
static cudaError_t launch_clip_optim_t(float* p, const float* g, float* s0, float* s1, int64_t n, float max_norm, float* coef, float* scratch,
                                       float lr, float a, float b, float eps, int step, int* dstep, cudaStream_t st) {
  static int per_sm = 0, sms = 0;
  if (!per_sm) {
    int dev = 0;
    SRL_TRY(cudaGetDevice(&dev));
    SRL_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SRL_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, clip_optim_kernel<OPT>, 512, 0));
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  }
  int64_t need = (n / 4 + 511) / 512;
  int blocks = (int)(need < 1 ? 1 : need);
  int cap = per_sm * sms; if (cap > 592) cap = 592;          // scratch holds 592 partials
  if (blocks > cap) blocks = cap;
  void* args[] = {&p, &g, &s0, &s1, &n, &max_norm, &coef, &scratch, &lr, &a, &b, &eps, &step, &dstep};
  return cudaLaunchCooperativeKernel((const void*)clip_optim_kernel<OPT>, dim3(blocks), dim3(512), args, 0, st);
}
cudaError_t launch_clip_optim(int optimizer, float* p, const float* g, float* s0, float* s1, int64_t n, float max_norm, float* coef,
                              float* scratch, float lr, float a, float b, float eps, int step, int* dstep, cudaStream_t st) {
  return optimizer == 0 ? launch_clip_optim_t<0>(p, g, s0, s1, n, max_norm, coef, scratch, lr, a, b, eps, step, dstep, st)
                        : launch_clip_optim_t<1>(p, g, s0, s1, n, max_norm, coef, scratch, lr, a, b, eps, step, dstep, st);
}

}  // namespace srl
