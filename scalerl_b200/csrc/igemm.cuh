// Implicit-GEMM mainloop on tcgen05 tensor cores (sm_100a), shared by every conv / fc kernel of the
// encoder (forward, dgrad, wgrad).
//
//   D[128 x BN] (fp32, TMEM)  =  sum over K-blocks  A_kb[128 x 64] * B_kb[BN x 64]^T      (bf16 operands)
//
// One CTA computes one 128 x BN output tile.  Warps 0-3 are producers (they gather the operand
// tiles -- im2col / transposed / u8->bf16 converted views of the NHWC activations -- into
// SWIZZLE_128B shared-memory tiles) and, after the mainloop, the epilogue (TMEM -> registers ->
// global).  Warp 4 owns TMEM allocation; its lane 0 issues tcgen05.mma and tcgen05.commit.
// Producer -> MMA hand-off: mbarrier full[s] (128 arrivals, after fence.proxy.async);
// MMA -> producer: tcgen05.commit on empty[s];  MMA -> epilogue: tcgen05.commit on done.
//
// A `Problem` type P supplies (all static):
//   BN                      UMMA N (multiple of 16, 16..256)
//   A_MN, B_MN              false: K-major tile (smem row = one M/N index, 64 contraction elems)
//                           true : MN-major tile (smem row = one contraction index, 64 M/N elems;
//                                  M=128 uses two 64-row blocks 8 KB apart)
//   STAGES                  smem pipeline depth
//   Params                  POD kernel argument
//   num_kblocks(p, tm, ty)  number of 64-deep K-blocks of this tile
//   RowA/RowB               per-(thread,row) context, built once per tile
//   make_rowA(p, tm, ty, srow) / make_rowB(...)       srow = smem row (0..127 / 0..B_ROWS-1)
//   load_A(p, rowctx, kb, chunk) -> uint4             8 bf16 for 16-byte chunk `chunk` of that row
//   load_B(p, rowctx, kb, chunk) -> uint4
//   epilogue16(p, tm, ty, row, col0, const float (&v)[16])   row = 0..127 of the tile
//   BIAS                    0: none.  1 / 2: the producers also accumulate fp32 column sums of the A / B operand
//                           rows they gather (MN-major wgrad problems: rows are dY pixels, so this is the bias
//                           gradient) and add them atomically to bias_dst(p, tm, ty) (nullptr: this CTA skips it)
//
// A second kernel, igemm_simt_kernel<P>, runs the SAME producers and epilogue around a plain
// CUDA-core inner product.  It exists only to triage (gather/epilogue bug vs descriptor/pipeline
// bug) from the tests; the product path always launches igemm_tc_kernel.
#pragma once
#include "common.cuh"

namespace srl {

constexpr int IG_PRODUCER_THREADS = 128;
constexpr int IG_THREADS = 160;
constexpr int IG_A_BYTES = 128 * 128;   // 128 smem rows x 128 B

template <class P>
struct IgemmCfg {
  static constexpr int B_ROWS = P::B_MN ? 64 * ((P::BN + 63) / 64) : P::BN;  // smem rows of the B tile
  static constexpr int B_BYTES = B_ROWS * 128;
  static constexpr int STAGE_BYTES = IG_A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = P::STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 512 /*bias sums*/;
  static constexpr int TMEM_COLS = P::BN <= 32 ? 32 : (P::BN <= 64 ? 64 : (P::BN <= 128 ? 128 : 256));
  static constexpr int A_PER_THREAD = 8;                 // 1024 chunks / 128 threads
  static constexpr int B_PER_THREAD = B_ROWS / 16;       // B_ROWS*8 chunks / 128 threads
  static_assert(P::BN % 16 == 0 && P::BN >= 16 && P::BN <= 256, "UMMA N");
  static_assert(B_ROWS % 16 == 0, "B rows");
};

SRL_DEVINL void acc_bf16x8(const uint4& v, float* s) {
  s[0] += bf16_lo(v.x); s[1] += bf16_hi(v.x); s[2] += bf16_lo(v.y); s[3] += bf16_hi(v.y);
  s[4] += bf16_lo(v.z); s[5] += bf16_hi(v.z); s[6] += bf16_lo(v.w); s[7] += bf16_hi(v.w);
}

template <class P>
SRL_DEVINL void ig_produce(const typename P::Params& p, const typename P::RowA (&ra)[8],
                           const typename P::RowB (&rb)[IgemmCfg<P>::B_PER_THREAD], int kb, uint8_t* sA, uint8_t* sB,
                           int tid, float* bsum = nullptr) {
  using C = IgemmCfg<P>;
  const int chunk = tid & 7;
  const int r0 = tid >> 3;
  uint4 va[8];
  uint4 vb[C::B_PER_THREAD];
#pragma unroll
  for (int i = 0; i < 8; ++i) va[i] = P::load_A(p, ra[i], kb, chunk);
#pragma unroll
  for (int i = 0; i < C::B_PER_THREAD; ++i) vb[i] = P::load_B(p, rb[i], kb, chunk);
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(sA + swz128(i * 16 + r0, chunk)) = va[i];
#pragma unroll
  for (int i = 0; i < C::B_PER_THREAD; ++i) *reinterpret_cast<uint4*>(sB + swz128(i * 16 + r0, chunk)) = vb[i];
  if (P::BIAS == 1 && bsum) {       // A rows: smem rows 0..63 are M-block 0, 64..127 M-block 1
#pragma unroll
    for (int i = 0; i < 8; ++i) acc_bf16x8(va[i], bsum + (i >= 4 ? 8 : 0));
  }
  if (P::BIAS == 2 && bsum) {
#pragma unroll
    for (int i = 0; i < C::B_PER_THREAD; ++i) acc_bf16x8(vb[i], bsum);
  }
}

template <class P>
__global__ void __launch_bounds__(IG_THREADS) igemm_tc_kernel(const typename P::Params p) {
  using C = IgemmCfg<P>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + P::STAGES;
  uint64_t* done = bars + 2 * P::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P::STAGES + 1);
  float* sbias = reinterpret_cast<float*>(bars + 2 * P::STAGES + 2);   // 128 floats (BIAS problems only)

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int tm = blockIdx.x, ty = blockIdx.y;
  const int nkb = P::num_kblocks(p, tm, ty);

  if (warp == 4) {
    if ((tid & 31) == 0) {
      for (int s = 0; s < P::STAGES; ++s) { mbar_init(&full[s], IG_PRODUCER_THREADS); mbar_init(&empty[s], 1); }
      mbar_init(done, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ---------------- producers ----------------
    typename P::RowA ra[8];
    typename P::RowB rb[C::B_PER_THREAD];
    const int r0 = tid >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) ra[i] = P::make_rowA(p, tm, ty, i * 16 + r0);
#pragma unroll
    for (int i = 0; i < C::B_PER_THREAD; ++i) rb[i] = P::make_rowB(p, tm, ty, i * 16 + r0);
    float bsum[16];
    float* bias_dst = nullptr;
    if constexpr (P::BIAS != 0) {
      bias_dst = P::bias_dst(p, tm, ty);
#pragma unroll
      for (int i = 0; i < 16; ++i) bsum[i] = 0.f;
      sbias[tid] = 0.f;
    }
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % P::STAGES;
      const uint32_t ph = (kb / P::STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      uint8_t* sA = smem + s * C::STAGE_BYTES;
      ig_produce<P>(p, ra, rb, kb, sA, sA + IG_A_BYTES, tid, (P::BIAS != 0 && bias_dst) ? bsum : nullptr);
      fence_proxy_async_smem();
      mbar_arrive(&full[s]);
    }
    if (P::BIAS != 0 && bias_dst) {   // block-uniform branch: 16 threads share each 8-channel chunk
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int chunk = tid & 7;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&sbias[chunk * 8 + e], bsum[e]);
        if (P::BIAS == 1) atomicAdd(&sbias[64 + chunk * 8 + e], bsum[8 + e]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid < P::BIAS_N) atomicAdd(bias_dst + tid, sbias[tid]);
    }
    // ---------------- epilogue ----------------
    mbar_wait(done, 0);
    tc_fence_after();
    const int row = tid;   // TMEM lane == tile row; warp w may only touch lanes 32w..32w+31
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < P::BN; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(lane_base + c0, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      P::epilogue16(p, tm, ty, row, c0, v);
    }
    tc_fence_before();
  } else if ((tid & 31) == 0) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = make_idesc_bf16(128, P::BN, P::A_MN ? 1 : 0, P::B_MN ? 1 : 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % P::STAGES;
      const uint32_t ph = (kb / P::STAGES) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint32_t a0 = smem_u32(smem + s * C::STAGE_BYTES);
      const uint32_t b0 = a0 + IG_A_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k) {   // 4 x (K=16) per 64-deep K-block
        // K-major: +32 B inside the 128 B swizzle row.  MN-major: 16 contraction rows = 2 groups of 8 = +2048 B.
        const uint64_t ad = P::A_MN ? make_smem_desc(a0 + k * 2048, 8192, 1024) : make_smem_desc(a0 + k * 32, 16, 1024);
        const uint64_t bd = P::B_MN ? make_smem_desc(b0 + k * 2048, 8192, 1024) : make_smem_desc(b0 + k * 32, 16, 1024);
        umma_bf16(tmem_base, ad, bd, idesc, (kb | k) != 0);
      }
      umma_commit(&empty[s]);
    }
    umma_commit(done);
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// --------------------------------------------------------------------------------------------
// triage kernel: same producers / smem image / epilogue, CUDA-core inner product
// --------------------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(IG_PRODUCER_THREADS) igemm_simt_kernel(const typename P::Params p) {
  using C = IgemmCfg<P>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + IG_A_BYTES;
  const int tid = threadIdx.x;
  const int tm = blockIdx.x, ty = blockIdx.y;
  const int nkb = P::num_kblocks(p, tm, ty);
  typename P::RowA ra[8];
  typename P::RowB rb[C::B_PER_THREAD];
  const int r0 = tid >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) ra[i] = P::make_rowA(p, tm, ty, i * 16 + r0);
#pragma unroll
  for (int i = 0; i < C::B_PER_THREAD; ++i) rb[i] = P::make_rowB(p, tm, ty, i * 16 + r0);
  float acc[P::BN];
#pragma unroll
  for (int n = 0; n < P::BN; ++n) acc[n] = 0.f;
  const int m = tid;
  __shared__ float sbias[128];
  float bsum[16];
  float* bias_dst = nullptr;
  if constexpr (P::BIAS != 0) {
    bias_dst = P::bias_dst(p, tm, ty);
#pragma unroll
    for (int i = 0; i < 16; ++i) bsum[i] = 0.f;
    sbias[tid] = 0.f;
  }
  for (int kb = 0; kb < nkb; ++kb) {
    __syncthreads();
    ig_produce<P>(p, ra, rb, kb, sA, sB, tid, (P::BIAS != 0 && bias_dst) ? bsum : nullptr);
    __syncthreads();
    for (int k = 0; k < 64; ++k) {
      uint32_t aoff = P::A_MN ? swz128((m >> 6) * 64 + k, (m & 63) >> 3) + (m & 7) * 2 : swz128(m, k >> 3) + (k & 7) * 2;
      const float a = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(sA + aoff));
#pragma unroll
      for (int n = 0; n < P::BN; ++n) {
        uint32_t boff = P::B_MN ? swz128((n >> 6) * 64 + k, (n & 63) >> 3) + (n & 7) * 2 : swz128(n, k >> 3) + (k & 7) * 2;
        acc[n] = fmaf(a, __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(sB + boff)), acc[n]);
      }
    }
  }
  if (P::BIAS != 0 && bias_dst) {
    __syncthreads();
    const int chunk = tid & 7;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&sbias[chunk * 8 + e], bsum[e]);
      if (P::BIAS == 1) atomicAdd(&sbias[64 + chunk * 8 + e], bsum[8 + e]);
    }
    __syncthreads();
    if (tid < P::BIAS_N) atomicAdd(bias_dst + tid, sbias[tid]);
  }
#pragma unroll
  for (int c0 = 0; c0 < P::BN; c0 += 16) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = acc[c0 + j];
    P::epilogue16(p, tm, ty, m, c0, v);
  }
}

template <class P>
cudaError_t igemm_launch(const typename P::Params& p, dim3 grid, cudaStream_t stream, bool simt = false) {
  using C = IgemmCfg<P>;
  if (grid.x == 0 || grid.y == 0) return cudaSuccess;
  if (simt) {
    const int smem = C::STAGE_BYTES + 1024;
    cudaError_t e = cudaFuncSetAttribute(igemm_simt_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    igemm_simt_kernel<P><<<grid, IG_PRODUCER_THREADS, smem, stream>>>(p);
  } else {
    cudaError_t e = cudaFuncSetAttribute(igemm_tc_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    igemm_tc_kernel<P><<<grid, IG_THREADS, C::SMEM_BYTES, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace srl
