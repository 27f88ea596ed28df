// Experiment / unit test: can a UMMA operand descriptor start at an arbitrary 128-byte ROW of a SWIZZLE_128B tile
// (i.e. not at a 1024-byte swizzle-atom boundary)?  If yes, every filter tap of a convolution can read a shifted
// window of ONE shared-memory copy of the input instead of a re-loaded im2col tile.
//   kmajor : D[128 x 64] = A[shift : shift+128, 0:64] * B[64 x 64]^T         (rows = M index)
//   mnmajor: D[128 x 64] = At[shift : shift+64, 0:128]^T * Bt[shift : shift+64, 0:64]   (rows = K index)
// base_offset_mode: 0 -> descriptor base_offset field 0; 1 -> base_offset = (start_address >> 7) & 7
#include "common.cuh"
#include "kernels.h"

namespace srl {

__global__ void __launch_bounds__(160) shift_test_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                         float* __restrict__ D, int shift, int mn_major, int bo_mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // K-major : sA = 160 rows x 128 B ; sB = 64 rows x 128 B
  // MN-major: sA = 2 blocks x 96 rows x 128 B ; sB = 96 rows x 128 B     (rows = contraction index, 64 + up to 32 shift)
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32768;
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < 128) {
    if (!mn_major) {
      for (int i = tid; i < 160 * 8; i += 128) {          // A is [160][64] in global
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sA + swz128(r, c)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
      }
      for (int i = tid; i < 64 * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sB + swz128(r, c)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
      }
    } else {
      for (int i = tid; i < 2 * 96 * 8; i += 128) {       // At is [96][128] in global; block b holds columns 64b..64b+63
        const int b = i / (96 * 8), r = (i >> 3) % 96, c = i & 7;
        *reinterpret_cast<uint4*>(sA + swz128(b * 96 + r, c)) = *reinterpret_cast<const uint4*>(A + r * 128 + b * 64 + c * 8);
      }
      for (int i = tid; i < 96 * 8; i += 128) {           // Bt is [96][64]
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sB + swz128(r, c)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
      }
    }
    fence_proxy_async_smem();
  }
  if (warp == 4) {
    if ((tid & 31) == 0) { mbar_init(done, 1); mbar_fence_init(); }
    __syncwarp();
    tmem_alloc(tmem_slot, 64);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 128) {
    const uint32_t idesc = make_idesc_bf16(128, 64, mn_major, mn_major);
    const uint32_t a0 = smem_u32(sA) + shift * 128, b0 = smem_u32(sB) + (mn_major ? shift * 128 : 0);
    for (int k = 0; k < 4; ++k) {
      const uint32_t aa = a0 + (mn_major ? k * 2048 : k * 32), bb = b0 + (mn_major ? k * 2048 : k * 32);
      uint64_t ad = mn_major ? make_smem_desc(aa, 96 * 128, 1024) : make_smem_desc(aa, 16, 1024);
      uint64_t bd = mn_major ? make_smem_desc(bb, 96 * 128, 1024) : make_smem_desc(bb, 16, 1024);
      if (bo_mode) {
        ad |= (uint64_t)((aa >> 7) & 7) << 49;
        bd |= (uint64_t)((bb >> 7) & 7) << 49;
      }
      umma_bf16(tmem_base, ad, bd, idesc, k != 0);
    }
    umma_commit(done);
  }
  if (tid < 128) {
    mbar_wait(done, 0);
    tc_fence_after();
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(lane_base + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[tid * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

cudaError_t test_shift(const void* A, const void* B, float* D, int shift, int mn_major, int bo_mode, cudaStream_t st) {
  const int smem = 49152 + 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(shift_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  shift_test_kernel<<<1, 160, smem, st>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, shift, mn_major, bo_mode);
  return cudaGetLastError();
}

// Test utility: leave NaN bit patterns in (almost) all shared memory of every SM, so that a kernel which multiplies
// stale shared memory by zero (instead of never reading it) shows up as NaN in the parity tests.
__global__ void poison_smem_kernel(uint32_t pattern) {
  extern __shared__ uint32_t poison[];
  for (int i = threadIdx.x; i < (200 * 1024) / 4; i += blockDim.x) poison[i] = pattern;
  __syncthreads();
  if (poison[(threadIdx.x * 37) % 1024] != pattern) __trap();    // keep the stores alive
}
cudaError_t test_poison_smem(cudaStream_t st) {
  const int smem = 200 * 1024;
  cudaError_t e = cudaFuncSetAttribute(poison_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  poison_smem_kernel<<<148 * 2, 256, smem, st>>>(0x7FC00000u);
  return cudaGetLastError();
}

// PDL self-test: kernel A spins ~delay_ns then sets flag = 1; kernel B (launched with programmatic stream serialization)
// executes griddepcontrol.wait and records the flag it sees.  out[0] must be 1 on every stream.
__global__ void pdl_test_a(volatile int* flag, unsigned delay_ns) {
  pdl_launch();
  if (threadIdx.x == 0) {
    const unsigned long long t0 = clock64();
    while (clock64() - t0 < (unsigned long long)delay_ns * 2) { }
    *flag = 1;
  }
}
__global__ void pdl_test_b(volatile int* flag, int* out) {
  pdl_wait();
  if (threadIdx.x == 0) out[blockIdx.x] = *flag;
}
cudaError_t test_pdl(int* flag, int* out, int nblk, unsigned delay_ns, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  pdl_test_a<<<1, 32, 0, st>>>(flag, delay_ns);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nblk); cfg.blockDim = dim3(32); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, pdl_test_b, (volatile int*)flag, out);
}

// MMA issue-rate microbenchmark: `issuers` warps (1 or 2) of ONE CTA each issue `reps` back-to-back tcgen05.mma (M = 128, N, K = 16,
// K-major SWIZZLE_128B operands already in shared memory, A descriptor starting `shift` rows into its tile) into their own TMEM
// accumulator and time issue + completion with clock64.  out[2*w] = cycles until the last MMA was ISSUED, out[2*w+1] = until the
// commit barrier fired.  Answers: what does one tcgen05.mma cost from a single thread, does a non-atom-aligned start row cost more,
// and do two issuing warps overlap?
__global__ void __launch_bounds__(96) mma_rate_kernel(int N, int shift, int reps, int issuers, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                   // 192 rows x 128 B
  uint8_t* sB = smem + 24576;           // 256 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 24576 + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (24576 + 32768) / 16; i += 96) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
  fence_proxy_async_smem();
  if (warp == 2) {
    if ((tid & 31) == 0) { for (int i = 0; i < 12; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // shift packs fields: bits 0-7 the A start row, bits 8-15 'commit every n MMAs' (0 = only at the end), bits 16-19 commits per point,
  // bits 20-23 the ISSUE-CODE variant: 0 = `if (lane == 0)` + make_smem_desc per MMA in a rolled loop (what the round-1 kernels do),
  // 1 = same branch, 16 MMAs unrolled, descriptor = base + constant, 2 = warp-converged code, elect.sync-predicated unrolled MMAs
  const int commit_every = (shift >> 8) & 255, ncommit = (shift >> 16) & 15, mode = (shift >> 20) & 15;
  shift &= 255;
  const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  const uint32_t a0 = smem_u32(sA) + shift * 128, b0 = smem_u32(sB);
  if (mode == 0 || mode == 1) {   // (modes 2, 3: the converged elect.sync form below)
  if (warp < issuers && (tid & 31) == 0) {
    const uint32_t acc = tmem_base + warp * 256;
    uint64_t* scratch = bars + 4 + warp * 4;      // barriers nobody waits on (phases just advance)
    const long long t0 = clock64();
    if (mode == 0) {
      for (int r = 0; r < reps; ++r) {
        const int k = r & 3;
        umma_bf16(acc, make_smem_desc(a0 + k * 32, 16, 1024), make_smem_desc(b0 + k * 32, 16, 1024), idesc, r != 0);
        if (commit_every && (r & (commit_every - 1)) == commit_every - 1)
          for (int c = 0; c < ncommit; ++c) umma_commit(&scratch[c]);
      }
    } else {
      const uint64_t ad = make_smem_desc(a0, 16, 1024), bd = make_smem_desc(b0, 16, 1024);
      for (int r = 0; r < reps; r += 16) {
#pragma unroll
        for (int i = 0; i < 16; ++i) umma_bf16(acc, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, (r | i) != 0);
        if (commit_every) for (int c = 0; c < ncommit; ++c) umma_commit(&scratch[c]);
      }
    }
    const long long t1 = clock64();
    umma_commit(&bars[warp]);
    mbar_wait(&bars[warp], 0);
    const long long t2 = clock64();
    out[2 * warp] = t1 - t0;
    out[2 * warp + 1] = t2 - t0;
  }
  } else if (warp < issuers) {
    // mode 3: both operands MN-major (the wgrad kernels' form: 16 contraction rows of 128 B per K step, M = two 64-wide blocks LBO apart)
    const bool mn = mode == 3;
    const uint32_t acc = tmem_base + warp * 256;
    uint64_t* scratch = bars + 4 + warp * 4;
    const uint64_t ad = mn ? make_smem_desc(a0, 128, 1024) : make_smem_desc(a0, 16, 1024), bd = mn ? make_smem_desc(b0, 8192, 1024) : make_smem_desc(b0, 16, 1024);
    const uint32_t idesc_v = mn ? make_idesc_bf16(128, N, 1, 1) : idesc;
    const uint32_t kstep = mn ? 128 : 2;
    const uint32_t leader = elect_one_sync();
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 16) {
      if (leader) {
#pragma unroll
        for (int i = 0; i < 16; ++i) umma_bf16(acc, ad + kstep * (i & 3), bd + kstep * (i & 3), idesc_v, (r | i) != 0);
        if (commit_every) for (int c = 0; c < ncommit; ++c) umma_commit(&scratch[c]);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (leader) umma_commit(&bars[warp]);
    __syncwarp();
    mbar_wait(&bars[warp], 0);
    const long long t2 = clock64();
    if (leader) { out[2 * warp] = t1 - t0; out[2 * warp + 1] = t2 - t0; }
  }
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}
cudaError_t test_mma_rate(int N, int shift, int reps, int issuers, long long* out, cudaStream_t st) {
  const int smem = 24576 + 32768 + 256 + 1024;
  cudaError_t e = cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  mma_rate_kernel<<<1, 96, smem, st>>>(N, shift, reps, issuers, out);
  return cudaGetLastError();
}

}  // namespace srl
