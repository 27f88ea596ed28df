// "Resident window" convolution kernels on tcgen05 (sm_100a).
//
// Every activation tensor is stored as rows of 64 bf16 channels (128 B) on the spatial grid of the conv INPUT it
// belongs to (frames concatenated: row = n * GRID + y * PITCH + x).  A filter tap is then a CONSTANT ROW SHIFT of the
// whole batch, so one shared-memory window of the input (loaded once with one 2-D TMA box) serves every tap: the UMMA
// operand descriptor of tap j simply starts  shift_j * 128 B  further into the window.  (UMMA applies SWIZZLE_128B to
// absolute shared-memory address bits, so a descriptor may start at any 128-byte row -- verified by
// tests/exp_shifted_operand.py / srl_test_shifted_operand.)  Compared with one im2col box per tap this divides the
// L2->SM operand traffic by the number of taps (4 / 8 / 9) and keeps the weights stationary in shared memory.
// Output positions are enumerated on the same grid; positions outside the valid output range are computed and
// discarded (conv1 9 %, conv2 19 %, conv3 40 % of the MMA rows -- the MMA is not the limiter of these layers).
//
//   res_fwd_kernel<P>   K-major, persistent: forward convs and dgrads (dgrad = negative shifts over dY stored on the
//                       grid of the conv input with zeros outside the valid outputs, which doubles as padding).
//                       warp 4 = TMA producer (weights once, then one window per 128-position tile, S-deep ring),
//                       warp 5 = MMA issuer (NT taps x 4 MMAs into one of two TMEM accumulators),
//                       warps 0-3 = epilogue of the previous tile while the next one is being multiplied.
//   res_wgrad_kernel<P> MN-major: dW[tap] = sum over positions X[pos + shift_tap]^T dY[pos].  One CTA owns a contiguous
//                       range of positions, streams (window, dY) chunks of 128 positions through a ring and keeps ALL
//                       taps' accumulators in TMEM (tap pairs form M = 128; an all-ones block yields the bias gradient).
#pragma once
#ifdef SRL_WGRAD_STAMP
#include <cstdio>
#endif
#include "igemm_tma.cuh"

namespace srl {

constexpr int RES_THREADS = 192;
constexpr int RES_MAX_TAPS = 10;

// ------------------------------------------------------------------------------------------------------------------
// forward / dgrad
//   P: BN, NT (taps), NWIN (input windows per tile: 1, or 2 for conv2's two row-parity planes), WROWS (rows per window,
//      128 + max shift - min shift), SHIFT_MIN, STAGES, Params{ in[NWIN] maps, w map, ... },
//      tap_win(j), tap_shift(j) (relative to SHIFT_MIN, i.e. >= 0), num_tiles(p), epilogue16(p, tile, row, c0, v)
// ------------------------------------------------------------------------------------------------------------------
//
// SPLIT = 1 is the fp32-accurate operand mode (srl_config_t.precision = 1): every bf16 operand tensor has a second, "low"
// tensor holding bf16(v - bf16(v)), and each product is issued as hi*hi + hi*lo + lo*hi into the same fp32 TMEM accumulator
// (the dropped lo*lo term is ~2^-18 relative; operands carry 16 significant bits, tighter than kind::tf32's 11).  Layouts,
// descriptors and tensor maps are those of the bf16 mode -- the low tensors are simply a second copy of everything -- so
// the mode exercises exactly the same data paths the fast mode uses.  It is for whole-step parity, not speed: 3x (2x where an
// operand is exact: the u8 frames) the MMAs, twice the operand traffic, fewer pipeline stages.
template <class P, int SPLIT>
struct ResFwdCfg {
  static constexpr int ALO = (SPLIT && P::A_LO) ? 1 : 0;          // the A operand has a low tensor (everything but the u8 frames)
  static constexpr int STAGES = SPLIT ? P::SPLIT_STAGES : P::STAGES;
  static constexpr int WIN_BYTES = ((P::WROWS * 128 + 1023) / 1024) * 1024;
  static constexpr int IN_HI_BYTES = P::NWIN * WIN_BYTES;
  static constexpr int IN_BYTES = IN_HI_BYTES * (1 + ALO);
  static constexpr int W_HI_BYTES = P::NT * P::BN * 128;
  static constexpr int W_BYTES = W_HI_BYTES * (1 + SPLIT);
  static constexpr int SMEM_BYTES = W_BYTES + STAGES * IN_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KB)");
  static constexpr int TMEM_COLS = 2 * P::BN <= 32 ? 32 : (2 * P::BN <= 64 ? 64 : (2 * P::BN <= 128 ? 128 : 256));
  static_assert(W_BYTES % 1024 == 0, "weight block alignment");
  static_assert(2 * P::BN <= 256, "two accumulators must fit");
};

template <class P, int SPLIT>
__global__ void __launch_bounds__(RES_THREADS) res_fwd_kernel(const __grid_constant__ typename P::Params p) {
  using C = ResFwdCfg<P, SPLIT>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sIn = smem + C::W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sIn + STAGES * C::IN_BYTES);
  uint64_t* in_full = bars;
  uint64_t* in_empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ntiles = P::num_tiles(p);

  if (warp == 4) {
    if ((tid & 31) == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
      mbar_init(w_full, 1);
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
  if (warp != 4) pdl_wait(P::KID);

  if (warp == 4) {
    const uint32_t leader = elect_one_sync();      // converged warp, one elected issuing lane: no vote loop around every TMA instruction
    // the packed weights were complete before the first kernel of the chain started: their load overlaps the previous
    // kernel's tail; the activations are only touched after pdl_wait().  (W_AFTER_WAIT: the weights come from the stream predecessor.)
    auto load_weights = [&]() {
      mbar_arrive_expect_tx(w_full, C::W_BYTES);
      for (int j = 0; j < P::NT; ++j) tma_load_2d(sW + j * P::BN * 128, &p.w, w_full, j * 64, 0);
      if constexpr (SPLIT)
        for (int j = 0; j < P::NT; ++j) tma_load_2d(sW + C::W_HI_BYTES + j * P::BN * 128, &p.w_lo, w_full, j * 64, 0);
    };
    if (leader && !P::W_AFTER_WAIT) load_weights();
    __syncwarp();
    pdl_wait();
    if (leader) pdl_launch();
    if (leader && P::W_AFTER_WAIT) load_weights();
    __syncwarp();
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % STAGES;
      mbar_wait(&in_empty[s], ((it / STAGES) & 1) ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&in_full[s], P::NWIN * P::WROWS * 128 * (1 + C::ALO));
        P::load_windows(p, t, sIn + s * C::IN_BYTES, C::WIN_BYTES, &in_full[s], false);
        if constexpr (C::ALO) P::load_windows(p, t, sIn + s * C::IN_BYTES + C::IN_HI_BYTES, C::WIN_BYTES, &in_full[s], true);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // The warp stays converged and ONE elected lane issues: with elect.sync the compiler keeps the descriptors in uniform registers and
    // emits a tile's tcgen05.mma back to back (40 / 48 clk each for N = 32 / 64); under `if (lane == 0)` every MMA sits in a vote loop
    // (57+ clk) -- profiles/r02_mma_issue_rate.md.  Descriptors are base + constant: the 14-bit address field cannot carry.
    const uint32_t leader = elect_one_sync();
    constexpr uint32_t idesc = make_idesc_bf16(128, P::BN, 0, 0);
    mbar_wait(w_full, 0);
    const uint64_t wd = make_smem_desc(smem_u32(sW), 16, 1024);
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % STAGES, b = it & 1;
      mbar_wait(&acc_empty[b], ((it >> 1) & 1) ^ 1);       // epilogue drained this accumulator (two tiles ago)
      mbar_wait(&in_full[s], (it / STAGES) & 1);
      tc_fence_after();
      const uint64_t ind = make_smem_desc(smem_u32(sIn + s * C::IN_BYTES), 16, 1024);
      const uint32_t acc = tmem_base + b * P::BN;
      if (leader) {
#pragma unroll
        for (int j = 0; j < P::NT; ++j) {
          const uint64_t ad = ind + (uint64_t)((P::tap_win(j) * C::WIN_BYTES + P::tap_shift(j) * 128) / 16);
          const uint64_t bd = wd + (uint64_t)(j * P::BN * 128 / 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(acc, ad + 2 * k, bd + 2 * k, idesc, (j | k) != 0);
            if constexpr (SPLIT)          // hi * lo(weights)
              umma_bf16(acc, ad + 2 * k, bd + (uint64_t)(C::W_HI_BYTES / 16 + 2 * k), idesc, 1);
            if constexpr (C::ALO)         // lo(activations) * hi
              umma_bf16(acc, ad + (uint64_t)(C::IN_HI_BYTES / 16 + 2 * k), bd + 2 * k, idesc, 1);
          }
        }
        umma_commit(&in_empty[s]);
        umma_commit(&acc_full[b]);
      }
      __syncwarp();
    }
  } else {
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int b = it & 1;
      // global operands of the epilogue (ReLU masks of the dgrads) are requested BEFORE waiting for the accumulator, so
      // their L2 latency overlaps the MMAs instead of being paid once per 16-column chunk
      uint4 pre[P::BN / 16][2];
#pragma unroll
      for (int c = 0; c < P::BN / 16; ++c) P::prefetch16(p, t, tid, c * 16, pre[c]);
      mbar_wait(&acc_full[b], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16) + b * P::BN;
#pragma unroll
      for (int c = 0; c < P::BN / 16; ++c) {
        uint32_t r[16];
        tmem_ld16(lane_base + c * 16, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        P::template epilogue16<SPLIT>(p, t, tid, c * 16, v, pre[c]);
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[b]);
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <class P, int SPLIT>
cudaError_t res_fwd_launch_t(const typename P::Params& p, int ntiles, int max_ctas, cudaStream_t stream) {
  using C = ResFwdCfg<P, SPLIT>;
  if (ntiles <= 0) return cudaSuccess;
  static PerDeviceOnce once;
  { cudaError_t e = ensure_max_dynamic_smem(once, res_fwd_kernel<P, SPLIT>, C::SMEM_BYTES); if (e != cudaSuccess) return e; }
  const int grid = ntiles < max_ctas ? ntiles : max_ctas;
  return launch_chain<PDL_RESFWD>(res_fwd_kernel<P, SPLIT>, dim3(grid), dim3(RES_THREADS), C::SMEM_BYTES, stream, p);
}
template <class P>
cudaError_t res_fwd_launch(const typename P::Params& p, int ntiles, int max_ctas, cudaStream_t stream, int split = 0) {
  return split ? res_fwd_launch_t<P, 1>(p, ntiles, max_ctas, stream) : res_fwd_launch_t<P, 0>(p, ntiles, max_ctas, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// wgrad
//   P: NACC accumulators (each M = 128 = two 64-row blocks, N = 64), NWIN, WROWS (= 128 + max shift), STAGES,
//      acc_win(a), acc_shift0(a), acc_shift1(a) (block 1 = -1 -> the all-ones block: bias gradient),
//      Params{ in[NWIN] maps, dy map, P (positions), chunks_per_cta }, epilogue16(p, acc, row, c0, v)
// ------------------------------------------------------------------------------------------------------------------
template <class P, int SPLIT>
struct ResWgradCfg {
  static constexpr int ALO = (SPLIT && P::A_LO) ? 1 : 0;          // the input window has a low tensor (not conv1: exact u8 frames)
  static constexpr int STAGES = SPLIT ? P::SPLIT_STAGES : P::STAGES;
  static constexpr bool BIAS_SMEM = P::SMEM_BIAS || SPLIT;        // split mode: bias gradient always from the staged dY tiles
  static constexpr int WIN_BYTES = ((P::WROWS * 128 + 1023) / 1024) * 1024;
  static constexpr int DY_BYTES = 128 * P::DY_CH * 2;             // 128 positions x DY_CH channels (64: SWIZZLE_128B rows, 32: SWIZZLE_64B rows)
  static constexpr int X_HI_BYTES = P::NWIN * WIN_BYTES;
  static constexpr int X_BYTES = X_HI_BYTES * (1 + ALO);
  static constexpr int STAGE_BYTES = X_BYTES + DY_BYTES * (1 + SPLIT);     // [windows hi][windows lo][dY hi][dY lo]
  static constexpr int ONES_BYTES = 128 * 128;
  static constexpr int SMEM_BYTES = ONES_BYTES + STAGES * STAGE_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KB)");
  static constexpr int TMEM_COLS = P::NACC * 64 <= 64 ? 64 : (P::NACC * 64 <= 128 ? 128 : (P::NACC * 64 <= 256 ? 256 : 512));
  static_assert(P::NACC * 64 <= 512, "accumulators must fit TMEM");
};

// Diagnostics build (SRL_DEFINES=SRL_WGRAD_STAMP, tests/diag/diag_wgrad.py): CTA 0 of a two-accumulator wgrad (conv1) prints %globaltimer
// stamps of its phases.  Never defined in the product build.
#ifdef SRL_WGRAD_STAMP
#define WG_STAMP(var) do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(var)); } while (0)
#else
#define WG_STAMP(var) do { } while (0)
#endif

template <class P, int SPLIT>
__global__ void __launch_bounds__(RES_THREADS) res_wgrad_kernel(const __grid_constant__ typename P::Params p) {
  using C = ResWgradCfg<P, SPLIT>;
  constexpr int STAGES = C::STAGES;
  constexpr bool BIAS_SMEM = C::BIAS_SMEM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sSt = smem;
  uint8_t* sOnes = smem + STAGES * C::STAGE_BYTES;     // after the stages: block-1 - block-0 distances stay positive
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + C::ONES_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* done = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nchunks_total = (p.P + 127) >> 7;
  const int c_begin = blockIdx.x * p.chunks_per_cta;
  const int c_end = min(nchunks_total, c_begin + p.chunks_per_cta);
  const int nch = max(0, c_end - c_begin);

  if constexpr (!BIAS_SMEM) {  // all-ones block (bf16 1.0) for the bias-gradient accumulator
    uint4* q = reinterpret_cast<uint4*>(sOnes);
    for (int i = tid; i < C::ONES_BYTES / 16; i += RES_THREADS) q[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  if (warp == 4) {
    if ((tid & 31) == 0) {
      // SMEM_BIAS: the four epilogue warps also read every dy tile (bias-gradient column sums), so a stage is free
      // only after the MMA commit AND their four arrivals
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], BIAS_SMEM ? 5 : 1); }
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
  unsigned long long wg_t0 = 0; WG_STAMP(wg_t0);
  pdl_wait(P::KID);
  if (tid == 128) pdl_launch();
  unsigned long long wg_t1 = 0; WG_STAMP(wg_t1);

  if (warp == 4) {
    const uint32_t leader = elect_one_sync();
    for (int i = 0; i < nch; ++i) {
      const int s = i % STAGES;
      mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&full[s], P::NWIN * P::WROWS * 128 * (1 + C::ALO) + C::DY_BYTES * (1 + SPLIT));
        uint8_t* st = sSt + s * C::STAGE_BYTES;
        P::load_windows(p, c_begin + i, st, C::WIN_BYTES, &full[s], false);
        if constexpr (C::ALO) P::load_windows(p, c_begin + i, st + C::X_HI_BYTES, C::WIN_BYTES, &full[s], true);
        tma_load_2d(st + C::X_BYTES, &p.dy, &full[s], 0, (c_begin + i) * 128);
        if constexpr (SPLIT) tma_load_2d(st + C::X_BYTES + C::DY_BYTES, &p.dy_lo, &full[s], 0, (c_begin + i) * 128);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    const uint32_t leader = elect_one_sync();        // converged warp, one elected issuing lane (see res_fwd_kernel)
    constexpr uint32_t idesc = make_idesc_bf16(128, P::DY_CH, 1, 1);
    constexpr uint32_t DYK = P::DY_CH * 2;           // descriptor address units per K = 16 step of the dY tile (16 rows x row bytes / 16)
    const uint32_t ones = smem_u32(sOnes);
#ifdef SRL_WGRAD_STAMP
    unsigned long long wg_full[20] = {};
#endif
    for (int i = 0; i < nch; ++i) {
      const int s = i % STAGES;
      mbar_wait(&full[s], (i / STAGES) & 1);
#ifdef SRL_WGRAD_STAMP
      if (i < 20) WG_STAMP(wg_full[i]);
#endif
      tc_fence_after();
      const uint32_t st = smem_u32(sSt + s * C::STAGE_BYTES);
      const uint64_t dyd = P::DY_CH == 64 ? make_smem_desc(st + C::X_BYTES, 8192, 1024) : make_smem_desc_sw64(st + C::X_BYTES, 4096, 512);
      if (leader) {
#pragma unroll
        for (int a = 0; a < P::NACC; ++a) {
          const uint32_t blk0 = st + P::acc_win(a) * C::WIN_BYTES + P::acc_shift0(a) * 128;
          // a half-empty accumulator (conv3's tenth "tap"): the all-ones block (bias gradient) in the bf16 mode; in the split
          // mode the bias comes from the staged dY tiles and the spare half just re-reads the window one row further (ignored)
          const uint32_t blk1 = P::acc_shift1(a) >= 0 ? st + P::acc_win1(a) * C::WIN_BYTES + P::acc_shift1(a) * 128
                                                      : (BIAS_SMEM ? blk0 + 128 : ones);
          const uint32_t lbo = blk1 - blk0;      // byte distance between the two 64-row M blocks (any multiple of 16)
          const uint64_t xd = make_smem_desc(blk0, lbo, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k) {          // 128 positions = 8 x (K = 16): +2048 B per step
            umma_bf16(tmem_base + a * 64, xd + 128 * k, dyd + DYK * k, idesc, (i | k) != 0);
            if constexpr (SPLIT)         // hi(x) * lo(dy)
              umma_bf16(tmem_base + a * 64, xd + 128 * k, dyd + (uint64_t)(C::DY_BYTES / 16 + DYK * k), idesc, 1);
            if constexpr (C::ALO)        // lo(x) * hi(dy)
              umma_bf16(tmem_base + a * 64, xd + (uint64_t)(C::X_HI_BYTES / 16 + 128 * k), dyd + DYK * k, idesc, 1);
          }
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (leader) umma_commit(done);
    __syncwarp();
#ifdef SRL_WGRAD_STAMP
    if (P::NACC == 2 && blockIdx.x == 0 && leader) {
      mbar_wait(done, 0);
      unsigned long long td; WG_STAMP(td);
      printf("wgrad cta0: prologue->pdl %llu ns; chunks full at (ns after pdl):", wg_t1 - wg_t0);
      for (int i = 0; i < nch && i < 20; ++i) printf(" %llu", wg_full[i] - wg_t1);
      printf("; all MMAs done %llu\n", td - wg_t1);
    }
#endif
  } else {
    if constexpr (BIAS_SMEM) {
      // bias gradient = column sums of dy, taken from the staged dy tiles while the MMAs run (no all-ones accumulator).
      // dy tile: 128 position rows of 128 B (64 channels), SWIZZLE_128B.  thread -> 16-byte channel group g, rows q + 16k.
      // dy tile rows: DY_CH channels = NG 16-byte groups; thread -> group g, rows q + RP k (RP rows per pass, NG passes)
      constexpr int NG = P::DY_CH / 8, RP = 128 / NG;
      const int g = tid & (NG - 1), q = tid / NG;
      float bs[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) bs[j] = 0.f;
      // Releasing a stage lets the producer's TMA overwrite it.  mbarrier.arrive does NOT wait for this warp's loads that
      // are still in flight (seen on sm_100a: generic LD.E.128 of the tile overtaken by the arrive -> rows of the NEXT
      // chunk were summed).  So the tile is read with ld.shared (same pipe as the mbarrier op), and every lane stores a
      // value that depends on all of its loads before the warp arrives: the store cannot issue until the loads returned,
      // and the arrive (release) is ordered after the store.
      const uint32_t dep_slot = smem_u32(sOnes) + 4096 + tid * 4;
      for (int i = 0; i < nch; ++i) {
        const int s = i % STAGES;
        mbar_wait(&full[s], (i / STAGES) & 1);
        const uint32_t dyt = smem_u32(sSt + s * C::STAGE_BYTES + C::X_BYTES);
        float dep = 0.f;
#pragma unroll
        for (int part = 0; part <= SPLIT; ++part) {        // split mode: the low tile's column sums are added too
#pragma unroll
          for (int k = 0; k < NG; ++k) {
            const uint4 v = lds128(dyt + part * C::DY_BYTES + (P::DY_CH == 64 ? swz128(q + RP * k, g) : swz64(q + RP * k, g)));
            bs[0] += bf16_lo(v.x); bs[1] += bf16_hi(v.x); bs[2] += bf16_lo(v.y); bs[3] += bf16_hi(v.y);
            bs[4] += bf16_lo(v.z); bs[5] += bf16_hi(v.z); bs[6] += bf16_lo(v.w); bs[7] += bf16_hi(v.w);
            dep += __uint_as_float(v.x ^ v.y ^ v.z ^ v.w);
          }
        }
        sts_volatile_f32(dep_slot, dep);
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
      }
      float* red = reinterpret_cast<float*>(sOnes);          // the all-ones block is not used by these problems
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (NG == 4) bs[j] += __shfl_xor_sync(0xffffffffu, bs[j], 4);
        bs[j] += __shfl_xor_sync(0xffffffffu, bs[j], 8);
        bs[j] += __shfl_xor_sync(0xffffffffu, bs[j], 16);
      }
      if ((tid & 31) < NG) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[warp * 64 + g * 8 + j] = bs[j];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");        // epilogue warps only
      if (tid < P::BIAS_CH && nch > 0) atomicAdd(p.db + tid, red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid]);
    }
    if (nch > 0) {
      unsigned long long wg_e0 = 0, wg_e1 = 0, wg_e2 = 0; WG_STAMP(wg_e0);
      mbar_wait(done, 0);
      WG_STAMP(wg_e1);
      tc_fence_after();
      const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int a = 0; a < P::NACC; ++a) {
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_base + a * 64 + c0, r);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          P::template epilogue16<SPLIT>(p, a, tid, c0, v);
        }
      }
      tc_fence_before();
#ifdef SRL_WGRAD_STAMP
      __threadfence(); WG_STAMP(wg_e2);
      if (P::NACC == 2 && blockIdx.x == 0 && tid == 0) printf("wgrad cta0 epilogue: bias loop done %llu, accumulators ready %llu, reds issued+fenced %llu (ns after pdl)\n", wg_e0 - wg_t1, wg_e1 - wg_t1, wg_e2 - wg_t1);
#endif
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <class P, int SPLIT>
cudaError_t res_wgrad_launch_t(typename P::Params p, int target_ctas, cudaStream_t stream) {
  using C = ResWgradCfg<P, SPLIT>;
  const int nchunks = (p.P + 127) >> 7;
  if (nchunks <= 0) return cudaSuccess;
  static PerDeviceOnce once;
  { cudaError_t e = ensure_max_dynamic_smem(once, res_wgrad_kernel<P, SPLIT>, C::SMEM_BYTES); if (e != cudaSuccess) return e; }
  p.chunks_per_cta = (nchunks + target_ctas - 1) / target_ctas;
  const int grid = (nchunks + p.chunks_per_cta - 1) / p.chunks_per_cta;
  return launch_chain<PDL_RESWGRAD>(res_wgrad_kernel<P, SPLIT>, dim3(grid), dim3(RES_THREADS), C::SMEM_BYTES, stream, p);
}
template <class P>
cudaError_t res_wgrad_launch(const typename P::Params& p, int target_ctas, cudaStream_t stream, int split = 0) {
  return split ? res_wgrad_launch_t<P, 1>(p, target_ctas, stream) : res_wgrad_launch_t<P, 0>(p, target_ctas, stream);
}

}  // namespace srl
