// TMA-fed variant of the tcgen05 implicit-GEMM mainloop (sm_100a).
//
//   warp 4 lane 0 : producer -- one cp.async.bulk.tensor (TMA) box per operand block per stage; im2col is expressed
//                   as a multi-dimensional box over the NHWC activation (negative / out-of-range coordinates are
//                   zero-filled by the TMA unit, which gives convolution padding and batch tails for free)
//   warp 5 lane 0 : tcgen05.mma issuer (accumulators in TMEM), tcgen05.commit frees the stage / signals the epilogue
//   warps 0-3     : epilogue (tcgen05.ld -> registers -> global)
//   full[s]  : mbarrier, 1 arrival (producer's arrive.expect_tx) + TMA complete_tx bytes
//   empty[s] : mbarrier, 1 arrival (tcgen05.commit)
//
// Shared-memory tiles are SWIZZLE_128B (the tensor maps are encoded with CU_TENSOR_MAP_SWIZZLE_128B, so the bytes
// land exactly where the UMMA descriptors of igemm.cuh expect them).
//   K-major problems : A = 128 rows x 128 B (one row per output pixel, 64 contraction elements), 4 MMAs (K=16) per stage
//   MN-major problems: A = 2 blocks x KROWS rows x 128 B, B = KROWS rows x 128 B (row = one contraction index = one
//                      pixel/frame, 64 channels); KROWS/16 MMAs per stage.  Rows a box does not write stay zero
//                      (the stage buffers are zero-initialised once), so partial frames contribute nothing.
//
// A Problem P supplies: BN, A_MN, B_MN, STAGES, KROWS (MN-major only), ZERO_INIT, Params (holds the CUtensorMaps),
//   num_kblocks(p,tm,ty), issue(p,tm,ty,kb,sA,sB,bar), init_smem(p,tm,ty,stage_base,stage_bytes,tid) [ZERO_INIT only],
//   epilogue16(p,tm,ty,row,c0,v).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "kernels.h"

namespace srl {

SRL_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
SRL_DEVINL void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
SRL_DEVINL void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
SRL_DEVINL void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
SRL_DEVINL void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

constexpr int IGT_THREADS = 192;

// SPLIT = 1: fp32-accurate operand mode (see igemm_res.cuh): a stage holds [A hi][B hi][A lo][B lo]; the problem supplies
// issue_split(p, tm, ty, kb, stage_base, half_bytes, bar) (one expect_tx + the loads of all four tiles) and
// epilogue16<SPLIT>.  Problems that never run split (the LSTM GEMMs) only need the SPLIT = 0 interface.
template <class P, int SPLIT = 0>
struct TmaCfg {
  static constexpr int KROWS = P::A_MN ? P::KROWS : 64;
  static constexpr int A_BYTES = P::A_MN ? 2 * KROWS * 128 : 128 * 128;
  static constexpr int B_BLOCKS = (P::BN + 63) / 64;
  static constexpr int B_BYTES = P::B_MN ? B_BLOCKS * KROWS * 128 : P::BN * 128;
  static constexpr int HALF_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGE_BYTES = HALF_BYTES * (1 + SPLIT);
  static constexpr int SMEM_BYTES = P::STAGES * STAGE_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KB)");
  static constexpr int TMEM_COLS = P::BN <= 32 ? 32 : (P::BN <= 64 ? 64 : (P::BN <= 128 ? 128 : 256));
  static constexpr int MMAS = P::A_MN ? KROWS / 16 : 4;
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "operand tiles must keep 1024 B alignment (SWIZZLE_128B atoms)");
  static_assert(P::A_MN == P::B_MN, "mixed majors are not used");
  static_assert(KROWS % 16 == 0, "contraction rows per stage must be a multiple of UMMA K = 16");
};

template <class P, int SPLIT>
SRL_DEVINL void igt_epilogue(const typename P::Params& p, int tm, int ty, int row, int c0, float (&v)[16]) {
  if constexpr (SPLIT) P::template epilogue16<1>(p, tm, ty, row, c0, v); else P::epilogue16(p, tm, ty, row, c0, v);
}
template <class P, int SPLIT>
SRL_DEVINL void igt_epilogue(const typename P::Params& p, int tm, int ty, int row, int c0, float (&v)[16], const uint4 (&pre)[2]) {
  if constexpr (SPLIT) P::template epilogue16<1>(p, tm, ty, row, c0, v, pre); else P::epilogue16(p, tm, ty, row, c0, v, pre);
}

template <class P, int SPLIT = 0>
__global__ void __launch_bounds__(IGT_THREADS) igemm_tma_kernel(const __grid_constant__ typename P::Params p) {
  using C = TmaCfg<P, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + P::STAGES;
  uint64_t* done = bars + 2 * P::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P::STAGES + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int tm = blockIdx.x, ty = blockIdx.y;
  const int nkb = P::num_kblocks(p, tm, ty);

  if constexpr (P::ZERO_INIT) {
    if (P::zero_cta(ty)) {
      uint4* z = reinterpret_cast<uint4*>(smem);
      for (int i = tid; i < P::STAGES * C::STAGE_BYTES / 16; i += IGT_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    P::init_smem(p, tm, ty, smem, C::STAGE_BYTES, tid);
    fence_proxy_async_smem();
  }
  if (warp == 4) {
    if ((tid & 31) == 0) {
      for (int s = 0; s < P::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(done, 1);
      mbar_fence_init();
      P::prefetch(p);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait(P::KID);                   // prologue above overlaps the previous kernel's tail
  if (tid == 128) pdl_launch();

  if (warp == 4) {
    const uint32_t leader = elect_one_sync();
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % P::STAGES;
      mbar_wait(&empty[s], ((kb / P::STAGES) & 1) ^ 1);
      if (leader) {
        uint8_t* sA = smem + s * C::STAGE_BYTES;
        if constexpr (SPLIT) P::issue_split(p, tm, ty, kb, sA, C::A_BYTES, C::HALF_BYTES, &full[s]);
        else P::issue(p, tm, ty, kb, sA, sA + C::A_BYTES, &full[s]);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    const uint32_t leader = elect_one_sync();        // converged warp, one elected issuing lane (see igemm_res.cuh res_fwd_kernel)
    constexpr uint32_t idesc = make_idesc_bf16(128, P::BN, P::A_MN ? 1 : 0, P::B_MN ? 1 : 0);
    constexpr uint32_t ASTEP = P::A_MN ? 2048 / 16 : 32 / 16, BSTEP = P::B_MN ? 2048 / 16 : 32 / 16;     // descriptor address units per K = 16
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % P::STAGES;
      mbar_wait(&full[s], (kb / P::STAGES) & 1);
      tc_fence_after();
      const uint32_t a0 = smem_u32(smem + s * C::STAGE_BYTES);
      const uint64_t ad0 = P::A_MN ? make_smem_desc(a0, C::KROWS * 128, 1024) : make_smem_desc(a0, 16, 1024);
      const uint64_t bd0 = P::B_MN ? make_smem_desc(a0 + C::A_BYTES, C::KROWS * 128, 1024) : make_smem_desc(a0 + C::A_BYTES, 16, 1024);
      if (leader) {
#pragma unroll
        for (int k = 0; k < C::MMAS; ++k) {
          const uint64_t ad = ad0 + (uint64_t)(k * ASTEP), bd = bd0 + (uint64_t)(k * BSTEP);
          umma_bf16(tmem_base, ad, bd, idesc, (kb | k) != 0);
          if constexpr (SPLIT) {        // hi * lo, lo * hi (the low tiles sit HALF_BYTES further into the stage)
            umma_bf16(tmem_base, ad, bd + (uint64_t)(C::HALF_BYTES / 16), idesc, 1);
            umma_bf16(tmem_base, ad + (uint64_t)(C::HALF_BYTES / 16), bd, idesc, 1);
          }
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (leader) umma_commit(done);
    __syncwarp();
  } else {
    const int row = tid;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    if constexpr (P::PREFETCH) {     // epilogue with global operands (ReLU mask): requested before the accumulator is waited for
      uint4 pre[P::BN / 16][2];
#pragma unroll
      for (int c = 0; c < P::BN / 16; ++c) P::prefetch16(p, tm, ty, row, c * 16, pre[c]);
      mbar_wait(done, 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < P::BN / 16; ++c) {
        uint32_t r[16];
        tmem_ld16(lane_base + c * 16, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        igt_epilogue<P, SPLIT>(p, tm, ty, row, c * 16, v, pre[c]);
      }
    } else {
    if (nkb > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < P::BN; c0 += 16) {
      float v[16];
      if (nkb > 0) {
        uint32_t r[16];
        tmem_ld16(lane_base + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
      igt_epilogue<P, SPLIT>(p, tm, ty, row, c0, v);
    }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <class P, int SPLIT = 0>
cudaError_t igemm_tma_launch(const typename P::Params& p, dim3 grid, cudaStream_t stream) {
  using C = TmaCfg<P, SPLIT>;
  if (grid.x == 0 || grid.y == 0) return cudaSuccess;
  static PerDeviceOnce once;      // set once per device (outside any stream capture: the first step always runs eagerly)
  { cudaError_t e = ensure_max_dynamic_smem(once, igemm_tma_kernel<P, SPLIT>, C::SMEM_BYTES); if (e != cudaSuccess) return e; }
  return launch_chain<PDL_IGEMM>(igemm_tma_kernel<P, SPLIT>, grid, dim3(IGT_THREADS), C::SMEM_BYTES, stream, p);
}

}  // namespace srl
