// Resident-window problem definitions (see igemm_res.cuh) for conv1 / conv2 / conv3: forward, dgrad, wgrad.
//
// Row layouts (all bf16, 64 channels = 128 B per row, frames concatenated; "grid" = spatial grid of the conv input):
//   xs   [NF*441]   21x21 grid of conv1 (space-to-depth frame)            channel = (c,dy,dx)
//   a1   [2][NF*100] two row-parity planes of conv1's output, 10x10 grid of conv2: plane hp, row n*100 + (h>>1)*10 + (w>>1),
//                   channel = (w&1)*32 + c          (a stride-2 tap of conv2 = a unit row shift inside one plane)
//   a2   [NF*81]    9x9 grid of conv3
//   a3   [NF*49]    dense (the fc input)
//   da3g [NB*81]    d(conv3 out) on conv3's 9x9 grid, zeros outside the 7x7 valid outputs (those zeros ARE the padding of dgrad)
//   da2g [NB*100]   d(conv2 out) on conv2's 10x10 grid, zeros outside 9x9
//   da1g [NB*441]   d(conv1 out) on conv1's 21x21 grid, zeros outside 20x20; 32 channels = 64 B per row (SWIZZLE_64B tiles)
#pragma once
#include "igemm_res.cuh"
#include "encoder_problems.cuh"

namespace srl {

// bf16 store of 16 accumulator values; in the split (fp32-accurate) mode also the low tensor: lo = bf16(v - bf16(v))
template <int SPLIT>
SRL_DEVINL void store_act16(bf16* hi, bf16* lo, size_t elem_off, const float (&v)[16]) {
  store_bf16x16(hi + elem_off, v);
  if constexpr (SPLIT) {
    float r[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
    store_bf16x16(lo + elem_off, r);
  }
}

// ================================================================================================ forward
struct RConv1Fwd {   // 2x2 s1 over xs (== 8x8 s4 over the frame): taps (kh2,kw2) -> shifts {0, 1, 21, 22}
  static constexpr int KID = 11;        // diagnostics timeline id
  static constexpr int BN = 32, NT = 4, NWIN = 1, WROWS = 128 + 22, STAGES = 4, SPLIT_STAGES = 4;
  static constexpr bool A_LO = false;        // the frames are exact in bf16: only the weights have a low tensor
  struct Params { SRL_TMAP in0; SRL_TMAP w; SRL_TMAP w_lo; const float* bias; bf16* out; bf16* out_lo; int NF; int NFS; };   // NF frames now, NFS = frames the a1 planes are strided for
  // conv1's K-major weight copy (w1k) is written by the frame-conversion kernel's extra blocks (obs_s2d_kernel, encoder.cu), not by
  // pack_weights_kernel: conv1 then depends only on its stream predecessor and never waits for the re-pack (profiles/r02_timeline.md);
  // the weight tiles are therefore loaded AFTER griddepcontrol.wait
  static constexpr bool W_AFTER_WAIT = true;
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.in0); tma_prefetch_desc(&p.w); }
  SRL_DEVINL static int num_tiles(const Params& p) { return (p.NF * 441 + 127) >> 7; }
  SRL_DEVINL static constexpr int tap_win(int) { return 0; }
  SRL_DEVINL static constexpr int tap_shift(int j) { return (j >> 1) * 21 + (j & 1); }
  SRL_DEVINL static void load_windows(const Params& p, int t, uint8_t* dst, int, uint64_t* bar, bool) { tma_load_2d(dst, &p.in0, bar, 0, t * 128); }
  SRL_DEVINL static void prefetch16(const Params&, int, int, int, uint4 (&)[2]) {}
  template <int SPLIT>
  SRL_DEVINL static void epilogue16(const Params& p, int t, int row, int c0, float (&v)[16], const uint4 (&)[2]) {
    const int Q = t * 128 + row, n = Q / 441, r = Q - n * 441, oh = r / 21, ow = r - oh * 21;
    if (n >= p.NF || oh >= 20 || ow >= 20) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], 1.0f / 255.0f, __ldg(p.bias + c0 + j)), 0.f);
    const size_t prow = (size_t)(oh & 1) * p.NFS * 100 + (size_t)n * 100 + (oh >> 1) * 10 + (ow >> 1);    // plane stride: the buffer's frame capacity
    store_act16<SPLIT>(p.out, p.out_lo, prow * 64 + (ow & 1) * 32 + c0, v);
  }
};

struct RConv2Fwd {   // 4x4 s2 over a1: tap j = (kh, kww): plane kh&1, shift (kh>>1)*10 + kww, K-block = (kh, kw in {2kww, 2kww+1}, c)
  static constexpr int KID = 12;        // diagnostics timeline id
  static constexpr bool W_AFTER_WAIT = false;
  static constexpr int BN = 64, NT = 8, NWIN = 2, WROWS = 128 + 11, STAGES = 3, SPLIT_STAGES = 1;
  static constexpr bool A_LO = true;
  struct Params { SRL_TMAP in0; SRL_TMAP in1; SRL_TMAP w; SRL_TMAP in0_lo; SRL_TMAP in1_lo; SRL_TMAP w_lo; const float* bias; bf16* out; bf16* out_lo; int NF; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.in0); tma_prefetch_desc(&p.in1); tma_prefetch_desc(&p.w); }
  SRL_DEVINL static int num_tiles(const Params& p) { return (p.NF * 100 + 127) >> 7; }
  SRL_DEVINL static constexpr int tap_win(int j) { return (j >> 1) & 1; }
  SRL_DEVINL static constexpr int tap_shift(int j) { return (j >> 2) * 10 + (j & 1); }
  SRL_DEVINL static void load_windows(const Params& p, int t, uint8_t* dst, int win_bytes, uint64_t* bar, bool lo) {
    tma_load_2d(dst, lo ? &p.in0_lo : &p.in0, bar, 0, t * 128);
    tma_load_2d(dst + win_bytes, lo ? &p.in1_lo : &p.in1, bar, 0, t * 128);
  }
  SRL_DEVINL static void prefetch16(const Params&, int, int, int, uint4 (&)[2]) {}
  template <int SPLIT>
  SRL_DEVINL static void epilogue16(const Params& p, int t, int row, int c0, float (&v)[16], const uint4 (&)[2]) {
    const int Q = t * 128 + row, n = Q / 100, r = Q - n * 100, oh = r / 10, ow = r - oh * 10;
    if (n >= p.NF || oh >= 9 || ow >= 9) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + __ldg(p.bias + c0 + j), 0.f);
    store_act16<SPLIT>(p.out, p.out_lo, ((size_t)n * 81 + oh * 9 + ow) * 64 + c0, v);
  }
};

struct RConv3Fwd {   // 3x3 s1 over a2: tap (kh,kw) -> shift kh*9 + kw
  static constexpr int KID = 13;        // diagnostics timeline id
  static constexpr bool W_AFTER_WAIT = false;
  static constexpr int BN = 64, NT = 9, NWIN = 1, WROWS = 128 + 20, STAGES = 4, SPLIT_STAGES = 2;
  static constexpr bool A_LO = true;
  struct Params { SRL_TMAP in0; SRL_TMAP w; SRL_TMAP in0_lo; SRL_TMAP w_lo; const float* bias; bf16* out; bf16* out_lo; int NF; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.in0); tma_prefetch_desc(&p.w); }
  SRL_DEVINL static int num_tiles(const Params& p) { return (p.NF * 81 + 127) >> 7; }
  SRL_DEVINL static constexpr int tap_win(int) { return 0; }
  SRL_DEVINL static constexpr int tap_shift(int j) { return (j / 3) * 9 + j % 3; }
  SRL_DEVINL static void load_windows(const Params& p, int t, uint8_t* dst, int, uint64_t* bar, bool lo) { tma_load_2d(dst, lo ? &p.in0_lo : &p.in0, bar, 0, t * 128); }
  SRL_DEVINL static void prefetch16(const Params&, int, int, int, uint4 (&)[2]) {}
  template <int SPLIT>
  SRL_DEVINL static void epilogue16(const Params& p, int t, int row, int c0, float (&v)[16], const uint4 (&)[2]) {
    const int Q = t * 128 + row, n = Q / 81, r = Q - n * 81, oh = r / 9, ow = r - oh * 9;
    if (n >= p.NF || oh >= 7 || ow >= 7) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + __ldg(p.bias + c0 + j), 0.f);
    store_act16<SPLIT>(p.out, p.out_lo, ((size_t)n * 49 + oh * 7 + ow) * 64 + c0, v);
  }
};

// ================================================================================================ dgrad
struct RConv3Dgrad {   // da2[ih,iw] = sum_{kh,kw} da3g[(ih-kh),(iw-kw)] W3[:, :, kh, kw]: shifts -(kh*9+kw); window starts 20 rows early
  static constexpr int KID = 14;        // diagnostics timeline id
  static constexpr bool W_AFTER_WAIT = false;
  static constexpr int BN = 64, NT = 9, NWIN = 1, WROWS = 128 + 20, STAGES = 4, SPLIT_STAGES = 2;
  static constexpr bool A_LO = true;
  struct Params { SRL_TMAP in0; SRL_TMAP w; SRL_TMAP in0_lo; SRL_TMAP w_lo; const bf16* act; bf16* dx; bf16* dx_lo; int NB; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.in0); tma_prefetch_desc(&p.w); }
  SRL_DEVINL static int num_tiles(const Params& p) { return (p.NB * 81 + 127) >> 7; }
  SRL_DEVINL static constexpr int tap_win(int) { return 0; }
  SRL_DEVINL static constexpr int tap_shift(int j) { return 20 - ((j / 3) * 9 + j % 3); }
  SRL_DEVINL static void load_windows(const Params& p, int t, uint8_t* dst, int, uint64_t* bar, bool lo) { tma_load_2d(dst, lo ? &p.in0_lo : &p.in0, bar, 0, t * 128 - 20); }
  SRL_DEVINL static void prefetch16(const Params& p, int t, int row, int c0, uint4 (&m)[2]) {
    const int Q = t * 128 + row;
    if (Q < p.NB * 81) ld_mask16(p.act + (size_t)Q * 64 + c0, m);                  // a2 lives on the same 9x9 grid
  }
  template <int SPLIT>
  SRL_DEVINL static void epilogue16(const Params& p, int t, int row, int c0, float (&v)[16], const uint4 (&m)[2]) {
    const int Q = t * 128 + row, n = Q / 81, r = Q - n * 81, ih = r / 9, iw = r - ih * 9;
    if (n >= p.NB) return;
    relu_mask16_pre(m, v);
    store_act16<SPLIT>(p.dx, p.dx_lo, ((size_t)n * 100 + ih * 10 + iw) * 64 + c0, v);           // da2g: conv2's 10x10 grid
  }
};

struct RConv2Dgrad {   // the 4 stride-parity classes share A (da2g at (i'-kh', j'-kw')): one N = 4 x 32 GEMM; shifts -(kh'*10 + kw')
  static constexpr int KID = 15;        // diagnostics timeline id
  static constexpr bool W_AFTER_WAIT = false;
  static constexpr int BN = 128, NT = 4, NWIN = 1, WROWS = 128 + 11, STAGES = 3, SPLIT_STAGES = 2;
  static constexpr bool A_LO = true;
  struct Params { SRL_TMAP in0; SRL_TMAP w; SRL_TMAP in0_lo; SRL_TMAP w_lo; const bf16* act; bf16* dx; bf16* dx_lo; int NB; int NF; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.in0); tma_prefetch_desc(&p.w); }
  SRL_DEVINL static int num_tiles(const Params& p) { return (p.NB * 100 + 127) >> 7; }
  SRL_DEVINL static constexpr int tap_win(int) { return 0; }
  SRL_DEVINL static constexpr int tap_shift(int j) { return 11 - ((j >> 1) * 10 + (j & 1)); }
  SRL_DEVINL static void load_windows(const Params& p, int t, uint8_t* dst, int, uint64_t* bar, bool lo) { tma_load_2d(dst, lo ? &p.in0_lo : &p.in0, bar, 0, t * 128 - 11); }
  SRL_DEVINL static void prefetch16(const Params& p, int t, int row, int c0, uint4 (&m)[2]) {
    const int Q = t * 128 + row, cls = c0 >> 5, c = c0 & 31;
    if (Q < p.NB * 100) ld_mask16(p.act + ((size_t)(cls >> 1) * p.NF * 100 + Q) * 64 + (cls & 1) * 32 + c, m);   // a1 plane ph, same row Q
  }
  template <int SPLIT>
  SRL_DEVINL static void epilogue16(const Params& p, int t, int row, int c0, float (&v)[16], const uint4 (&m)[2]) {
    const int Q = t * 128 + row, n = Q / 100, r = Q - n * 100, i = r / 10, j = r - i * 10;
    if (n >= p.NB) return;
    const int cls = c0 >> 5, c = c0 & 31, ph = cls >> 1, pw = cls & 1;
    relu_mask16_pre(m, v);
    store_act16<SPLIT>(p.dx, p.dx_lo, ((size_t)n * 441 + (2 * i + ph) * 21 + 2 * j + pw) * 32 + c, v);   // da1g: conv1's 21x21 grid, 32-channel rows
  }
};

// ================================================================================================ wgrad
// The weight-gradient accumulators are added (16-byte vector reductions) into a zeroed fp32 workspace in the
// kernels' native [tap-block][row][co] order; conv_wgrad_finalize_kernel (encoder.cu) then writes the PyTorch-layout
// gradient tensors.  Workspace offsets (floats):
// (WS_W3 / WS_W2 / WS_W1 / WS_TOTAL are defined in kernels.h: the optimizer kernel reads the workspace too)
struct RConv3Wgrad {   // acc a = taps (2a, 2a+1); acc 4 = (tap 8, ones -> db3).  ws: [10 taps][64 c][64 co] fp32 (co contiguous)
  static constexpr int KID = 21;        // diagnostics timeline id
  static constexpr int NACC = 5, NWIN = 1, WROWS = 128 + 20, STAGES = 3, SPLIT_STAGES = 2;
  static constexpr bool A_LO = true;
  static constexpr bool SMEM_BIAS = false;     // the ninth tap leaves half an accumulator free: the all-ones block rides along
  static constexpr int BIAS_CH = 64, DY_CH = 64;
  struct Params { SRL_TMAP in0; SRL_TMAP dy; SRL_TMAP in0_lo; SRL_TMAP dy_lo; float* ws; float* db; int P; int chunks_per_cta; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.in0); tma_prefetch_desc(&p.dy); }
  SRL_DEVINL static constexpr int sh(int tap) { return (tap / 3) * 9 + tap % 3; }
  SRL_DEVINL static constexpr int acc_win(int) { return 0; }
  SRL_DEVINL static constexpr int acc_win1(int) { return 0; }
  SRL_DEVINL static constexpr int acc_shift0(int a) { return sh(2 * a); }
  SRL_DEVINL static constexpr int acc_shift1(int a) { return a < 4 ? sh(2 * a + 1) : -1; }
  SRL_DEVINL static void load_windows(const Params& p, int chunk, uint8_t* dst, int, uint64_t* bar, bool lo) { tma_load_2d(dst, lo ? &p.in0_lo : &p.in0, bar, 0, chunk * 128); }
  template <int SPLIT>
  SRL_DEVINL static void epilogue16(const Params& p, int a, int row, int c0, float (&v)[16]) {
    const int tap = 2 * a + (row >> 6), c = row & 63;
    if (tap < 9) {
      red_add_16(p.ws + ((size_t)(a * 128 + row)) * 64 + c0, v);
    } else if (!SPLIT && c == 0) {          // split mode: db3 comes from the staged dY tiles (igemm_res.cuh)
#pragma unroll
      for (int j = 0; j < 16; ++j) atomicAdd(p.db + c0 + j, v[j]);
    }
  }
};

struct RConv2Wgrad {   // acc a = kh (blocks kww = 0,1: rows = (kw = 2kww + wp, c)); db2 = column sums of the staged dy tiles
  static constexpr int KID = 22;        // diagnostics timeline id
  static constexpr int NACC = 4, NWIN = 2, WROWS = 128 + 11, STAGES = 3, SPLIT_STAGES = 1;
  static constexpr bool A_LO = true;
  static constexpr bool SMEM_BIAS = true;
  static constexpr int BIAS_CH = 64, DY_CH = 64;
  struct Params { SRL_TMAP in0; SRL_TMAP in1; SRL_TMAP dy; SRL_TMAP in0_lo; SRL_TMAP in1_lo; SRL_TMAP dy_lo; float* ws; float* db; int P; int chunks_per_cta; };   // ws: [4 kh][128 (kw,c)][64 co]
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.in0); tma_prefetch_desc(&p.in1); tma_prefetch_desc(&p.dy); }
  SRL_DEVINL static constexpr int acc_win(int a) { return a & 1; }
  SRL_DEVINL static constexpr int acc_win1(int a) { return a & 1; }
  SRL_DEVINL static constexpr int acc_shift0(int a) { return (a >> 1) * 10; }
  SRL_DEVINL static constexpr int acc_shift1(int a) { return (a >> 1) * 10 + 1; }
  SRL_DEVINL static void load_windows(const Params& p, int chunk, uint8_t* dst, int win_bytes, uint64_t* bar, bool lo) {
    tma_load_2d(dst, lo ? &p.in0_lo : &p.in0, bar, 0, chunk * 128);
    tma_load_2d(dst + win_bytes, lo ? &p.in1_lo : &p.in1, bar, 0, chunk * 128);
  }
  template <int SPLIT>
  SRL_DEVINL static void epilogue16(const Params& p, int a, int row, int c0, float (&v)[16]) {
    red_add_16(p.ws + ((size_t)(a * 128 + row)) * 64 + c0, v);
  }
};

struct RConv1Wgrad {   // acc a = kh2 (blocks kw2 = 0,1: rows = (kw2, c, dy, dx)); db1 = column sums of the staged dy tiles
  static constexpr int KID = 23;        // diagnostics timeline id
  static constexpr int NACC = 2, NWIN = 1, WROWS = 128 + 22, STAGES = 5, SPLIT_STAGES = 3;
  static constexpr bool A_LO = false;          // the frames are exact in bf16
  static constexpr bool SMEM_BIAS = true;
  static constexpr int BIAS_CH = 32;
  static constexpr int DY_CH = 32;             // da1g rows are 32 channels (64 B): SWIZZLE_64B dY tiles, N = 32 MMAs
  struct Params { SRL_TMAP in0; SRL_TMAP dy; SRL_TMAP dy_lo; float* ws; float* db; int P; int chunks_per_cta; };   // ws: [2 kh2][128 (kw2,c,dy,dx)][32 co]
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.in0); tma_prefetch_desc(&p.dy); }
  SRL_DEVINL static constexpr int acc_win(int) { return 0; }
  SRL_DEVINL static constexpr int acc_win1(int) { return 0; }
  SRL_DEVINL static constexpr int acc_shift0(int a) { return a * 21; }
  SRL_DEVINL static constexpr int acc_shift1(int a) { return a * 21 + 1; }
  SRL_DEVINL static void load_windows(const Params& p, int chunk, uint8_t* dst, int, uint64_t* bar, bool) { tma_load_2d(dst, &p.in0, bar, 0, chunk * 128); }
  template <int SPLIT>
  SRL_DEVINL static void epilogue16(const Params& p, int a, int row, int c0, float (&v)[16]) {
    if (c0 >= 32) return;                    // N = 32: accumulator columns 32..63 are not written by the MMAs
    red_add_16(p.ws + ((size_t)(a * 128 + row)) * 32 + c0, v);
  }
};

}  // namespace srl
