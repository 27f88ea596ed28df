// TMA box programs + epilogues of every encoder GEMM (see igemm_tma.cuh for the mainloop, encoder.cu for the
// tensor maps).  Reference arithmetic: scalerl/algorithms/utils/atari_model.py:30-46,93-101 and its backward.
//
// Tensor-map views (bf16, innermost dimension first, all SWIZZLE_128B, box inner extent = 64 elements = 128 B):
//   xs3  : xs  as [64][21][NF*21]            conv1 fwd rows   (frame rows are concatenated: row R = n*21 + y)
//   xs4  : xs  as [64][21][21][NF]           conv1 wgrad
//   a1v  : a1  as [64][10][2][10][NF]        (c+32*(w&1), w>>1, h&1, h>>1, n): the stride-2 taps of conv2 become unit-stride boxes
//   a2v  : a2  as [64][9][9][NF]
//   a3m  : a3  as [3136][NF]                 fc rows
//   dhm  : dh  as [512][NB];  da3v [64][7][7][NB], da3m [64][NB*49];  da2v [64][9][9][NB], da2m [64][NB*81];
//   da1m : da1 as [64][NB*400]               (da1 is stored with a 64-channel pitch, channels 32..63 are zero)
//   weights: w1k [256][32], w2k [512][64], w3k [576][64], wfk [3136][512], wfd [512][3136], w3d [576][64], w2d [256][128]
#pragma once
#include "igemm_tma.cuh"
#include "encoder_problems.cuh"   // store_bf16x16, relu_mask16, bf16 helpers

namespace srl {

#define SRL_TMAP alignas(64) CUtensorMap

// ============================================================================================ forward
struct TConv1Fwd {   // tile = 6 concatenated output rows x 20 columns (120 of 128 MMA rows); kb = (kh2, kw2)
  static constexpr int BN = 32, STAGES = 4;
  static constexpr bool A_MN = false, B_MN = false, ZERO_INIT = false;
  static constexpr int KROWS = 64;
  struct Params { SRL_TMAP xs3; SRL_TMAP w; const float* bias; bf16* out; int NF; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.xs3); tma_prefetch_desc(&p.w); }
  SRL_DEVINL static int num_kblocks(const Params&, int, int) { return 4; }
  SRL_DEVINL static void issue(const Params& p, int tm, int, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    mbar_arrive_expect_tx(bar, 120 * 128 + 32 * 128);
    tma_load_3d(sA, &p.xs3, bar, 0, kb & 1, tm * 6 + (kb >> 1));
    tma_load_2d(sB, &p.w, bar, kb * 64, 0);
  }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int, int row, int c0, float (&v)[16]) {
    if (row >= 120) return;
    const int rr = row / 20, ow = row - rr * 20, R = tm * 6 + rr, n = R / 21, oh = R - n * 21;
    if (oh >= 20 || n >= p.NF) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], 1.0f / 255.0f, __ldg(p.bias + c0 + j)), 0.f);
    store_bf16x16(p.out + ((size_t)(n * 20 + oh) * 20 + ow) * 32 + c0, v);
  }
};

struct TConv2Fwd {   // tile = one frame (81 rows); kb = (kh, kw pair)
  static constexpr int BN = 64, STAGES = 4;
  static constexpr bool A_MN = false, B_MN = false, ZERO_INIT = false;
  static constexpr int KROWS = 64;
  struct Params { SRL_TMAP a1v; SRL_TMAP w; const float* bias; bf16* out; int NF; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.a1v); tma_prefetch_desc(&p.w); }
  SRL_DEVINL static int num_kblocks(const Params&, int, int) { return 8; }
  SRL_DEVINL static void issue(const Params& p, int tm, int, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    const int kh = kb >> 1;
    mbar_arrive_expect_tx(bar, 81 * 128 + 64 * 128);
    tma_load_5d(sA, &p.a1v, bar, 0, kb & 1, kh & 1, kh >> 1, tm);
    tma_load_2d(sB, &p.w, bar, kb * 64, 0);
  }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int, int row, int c0, float (&v)[16]) {
    if (row >= 81) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + __ldg(p.bias + c0 + j), 0.f);
    store_bf16x16(p.out + ((size_t)tm * 81 + row) * 64 + c0, v);
  }
};

struct TConv3Fwd {   // tile = two frames (98 rows) in one 4-D box; kb = (kh, kw)
  static constexpr int BN = 64, STAGES = 4;
  static constexpr bool A_MN = false, B_MN = false, ZERO_INIT = false;
  static constexpr int KROWS = 64;
  struct Params { SRL_TMAP a2v; SRL_TMAP w; const float* bias; bf16* out; int NF; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.a2v); tma_prefetch_desc(&p.w); }
  SRL_DEVINL static int num_kblocks(const Params&, int, int) { return 9; }
  SRL_DEVINL static void issue(const Params& p, int tm, int, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    const int kh = kb / 3, kw = kb - kh * 3;
    mbar_arrive_expect_tx(bar, 98 * 128 + 64 * 128);
    tma_load_4d(sA, &p.a2v, bar, 0, kw, kh, tm * 2);
    tma_load_2d(sB, &p.w, bar, kb * 64, 0);
  }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int, int row, int c0, float (&v)[16]) {
    const int m = tm * 98 + row;
    if (row >= 98 || m >= p.NF * 49) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + __ldg(p.bias + c0 + j), 0.f);
    store_bf16x16(p.out + (size_t)m * 64 + c0, v);
  }
};

struct TFcFwd {   // split-K partials; grid.y = 8 N-tiles x FC_SPLITS, ty = nt*FC_SPLITS + split
  static constexpr int BN = 64, STAGES = 4, SPLITS = 4;
  static constexpr bool A_MN = false, B_MN = false, ZERO_INIT = false;
  static constexpr int KROWS = 64;
  struct Params { SRL_TMAP a3m; SRL_TMAP w; float* out; int M; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.a3m); tma_prefetch_desc(&p.w); }
  SRL_DEVINL static int kb_begin(int split) { return (49 * split) / SPLITS; }
  SRL_DEVINL static int num_kblocks(const Params&, int, int ty) { const int sp = ty % SPLITS; return kb_begin(sp + 1) - kb_begin(sp); }
  SRL_DEVINL static void issue(const Params& p, int tm, int ty, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    const int k = (kb_begin(ty % SPLITS) + kb) * 64;
    mbar_arrive_expect_tx(bar, 128 * 128 + 64 * 128);
    tma_load_2d(sA, &p.a3m, bar, k, tm * 128);
    tma_load_2d(sB, &p.w, bar, k, (ty / SPLITS) * 64);
  }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int ty, int row, int c0, float (&v)[16]) {
    const int m = tm * 128 + row;
    if (m >= p.M) return;
    float4* o = reinterpret_cast<float4*>(p.out + ((size_t)(ty % SPLITS) * p.M + m) * 512 + (ty / SPLITS) * 64 + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
};

// ============================================================================================ dgrad
struct TFcDgrad {   // da3[m][i] = (dh[m][:] . Wfc[:][i]) * (a3 > 0); grid = (ceil(M/128), 49)
  static constexpr int BN = 64, STAGES = 4;
  static constexpr bool A_MN = false, B_MN = false, ZERO_INIT = false;
  static constexpr int KROWS = 64;
  struct Params { SRL_TMAP dhm; SRL_TMAP w; const bf16* a3; bf16* da3; int M; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.dhm); tma_prefetch_desc(&p.w); }
  SRL_DEVINL static int num_kblocks(const Params&, int, int) { return 8; }
  SRL_DEVINL static void issue(const Params& p, int tm, int ty, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    mbar_arrive_expect_tx(bar, 128 * 128 + 64 * 128);
    tma_load_2d(sA, &p.dhm, bar, kb * 64, tm * 128);
    tma_load_2d(sB, &p.w, bar, kb * 64, ty * 64);
  }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int ty, int row, int c0, float (&v)[16]) {
    const int m = tm * 128 + row;
    if (m >= p.M) return;
    const size_t idx = (size_t)m * 3136 + ty * 64 + c0;
    relu_mask16(p.a3 + idx, v);
    store_bf16x16(p.da3 + idx, v);
  }
};

struct TConv3Dgrad {   // tile = one frame's 9x9 input grid (81 rows); tap (kh,kw) reads da3 at (ih-kh, iw-kw): box origin (-kw,-kh)
  static constexpr int BN = 64, STAGES = 4;
  static constexpr bool A_MN = false, B_MN = false, ZERO_INIT = false;
  static constexpr int KROWS = 64;
  struct Params { SRL_TMAP dyv; SRL_TMAP w; const bf16* act; bf16* dx; int NB; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.dyv); tma_prefetch_desc(&p.w); }
  SRL_DEVINL static int num_kblocks(const Params&, int, int) { return 9; }
  SRL_DEVINL static void issue(const Params& p, int tm, int, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    const int kh = kb / 3, kw = kb - kh * 3;
    mbar_arrive_expect_tx(bar, 81 * 128 + 64 * 128);
    tma_load_4d(sA, &p.dyv, bar, 0, -kw, -kh, tm);
    tma_load_2d(sB, &p.w, bar, kb * 64, 0);
  }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int, int row, int c0, float (&v)[16]) {
    if (row >= 81) return;
    const size_t idx = ((size_t)tm * 81 + row) * 64 + c0;
    relu_mask16(p.act + idx, v);
    store_bf16x16(p.dx + idx, v);
  }
};

struct TConv2Dgrad {   // grid = (frames, 1); rows = the 10x10 positions (i',j'); the four stride-parity classes share the same
                        // A operand (da2 at (i'-kh', j'-kw')) and differ only in the weights, so they are ONE GEMM with
                        // N = 4 classes x 32 channels = 128; column block cls goes to input pixel (2i'+ph, 2j'+pw).  kb = (kh', kw')
  static constexpr int BN = 128, STAGES = 3;
  static constexpr bool A_MN = false, B_MN = false, ZERO_INIT = false;
  static constexpr int KROWS = 64;
  struct Params { SRL_TMAP dyv; SRL_TMAP w; const bf16* act; bf16* dx; int NB; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.dyv); tma_prefetch_desc(&p.w); }
  SRL_DEVINL static int num_kblocks(const Params&, int, int) { return 4; }
  SRL_DEVINL static void issue(const Params& p, int tm, int, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    mbar_arrive_expect_tx(bar, 100 * 128 + 128 * 128);
    tma_load_4d(sA, &p.dyv, bar, 0, -(kb & 1), -(kb >> 1), tm);
    tma_load_2d(sB, &p.w, bar, kb * 64, 0);
  }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int, int row, int c0, float (&v)[16]) {
    if (row >= 100) return;
    const int i = row / 10, j = row - i * 10, cls = c0 >> 5, c = c0 & 31;
    const size_t pix = (size_t)(tm * 20 + 2 * i + (cls >> 1)) * 20 + 2 * j + (cls & 1);
    relu_mask16(p.act + pix * 32 + c, v);
    store_bf16x16(p.dx + pix * 64 + c, v);     // da1 has a 64-channel pitch (upper half stays zero)
  }
};

// ============================================================================================ wgrad (MN-major)
// One stage = one frame (or part of it).  blockIdx.x = split of the frame range, blockIdx.y = 128-row slice of dW,
// plus ONE extra slice whose A (or B) block is all ones: its accumulator rows are the column sums of dY = the bias gradient.
SRL_DEVINL void fill_ones(uint8_t* dst, int bytes, int tid) {   // bf16 1.0 = 0x3F80
  uint4* q = reinterpret_cast<uint4*>(dst);
  for (int i = tid; i < bytes / 16; i += IGT_THREADS) q[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
}

struct TConv3Wgrad {   // grid.y = 5 tap pairs (tap 9 of the last pair = ones -> db3); stage = 1 frame = 49 rows (+15 zero rows)
  static constexpr int BN = 64, STAGES = 4, KROWS = 64;
  static constexpr bool A_MN = true, B_MN = true, ZERO_INIT = true;
  struct Params { SRL_TMAP actv; SRL_TMAP dym; float* dw; float* db; int NB; int fps; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.actv); tma_prefetch_desc(&p.dym); }
  SRL_DEVINL static int num_kblocks(const Params& p, int tm, int) { return max(0, min(p.NB, (tm + 1) * p.fps) - tm * p.fps); }
  SRL_DEVINL static void init_smem(const Params&, int, int ty, uint8_t* smem, int stage_bytes, int tid) {
    if (ty == 4)
      for (int s = 0; s < STAGES; ++s) fill_ones(smem + s * stage_bytes + KROWS * 128, KROWS * 128, tid);
  }
  SRL_DEVINL static void issue(const Params& p, int tm, int ty, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    const int n = tm * p.fps + kb, t0 = 2 * ty, t1 = 2 * ty + 1;
    mbar_arrive_expect_tx(bar, (t1 < 9 ? 3 : 2) * 49 * 128);
    tma_load_4d(sA, &p.actv, bar, 0, t0 % 3, t0 / 3, n);
    if (t1 < 9) tma_load_4d(sA + KROWS * 128, &p.actv, bar, 0, t1 % 3, t1 / 3, n);
    tma_load_2d(sB, &p.dym, bar, 0, n * 49);
  }
  SRL_DEVINL static void epilogue16(const Params& p, int, int ty, int row, int c0, float (&v)[16]) {
    const int tap = 2 * ty + (row >> 6), c = row & 63;
    if (tap < 9) {
#pragma unroll
      for (int j = 0; j < 16; ++j) atomicAdd(p.dw + ((c0 + j) * 64 + c) * 9 + tap, v[j]);
    } else if (c == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) atomicAdd(p.db + c0 + j, v[j]);
    }
  }
};

struct TConv2Wgrad {   // grid.y = 4 kh slices (+1 ones slice -> db2); rows = (kw, c); stage = 1 frame = 81 rows (+15 zero rows)
  static constexpr int BN = 64, STAGES = 3, KROWS = 96;
  static constexpr bool A_MN = true, B_MN = true, ZERO_INIT = true;
  struct Params { SRL_TMAP actv; SRL_TMAP dym; float* dw; float* db; int NB; int fps; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.actv); tma_prefetch_desc(&p.dym); }
  SRL_DEVINL static int num_kblocks(const Params& p, int tm, int) { return max(0, min(p.NB, (tm + 1) * p.fps) - tm * p.fps); }
  SRL_DEVINL static void init_smem(const Params&, int, int ty, uint8_t* smem, int stage_bytes, int tid) {
    if (ty == 4)
      for (int s = 0; s < STAGES; ++s) fill_ones(smem + s * stage_bytes, 2 * KROWS * 128, tid);
  }
  SRL_DEVINL static void issue(const Params& p, int tm, int ty, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    const int n = tm * p.fps + kb;
    mbar_arrive_expect_tx(bar, (ty < 4 ? 3 : 1) * 81 * 128);
    if (ty < 4) {
      tma_load_5d(sA, &p.actv, bar, 0, 0, ty & 1, ty >> 1, n);                  // kw 0,1
      tma_load_5d(sA + KROWS * 128, &p.actv, bar, 0, 1, ty & 1, ty >> 1, n);    // kw 2,3
    }
    tma_load_2d(sB, &p.dym, bar, 0, n * 81);
  }
  SRL_DEVINL static void epilogue16(const Params& p, int, int ty, int row, int c0, float (&v)[16]) {
    if (ty < 4) {
      const int kw = row >> 5, c = row & 31;
#pragma unroll
      for (int j = 0; j < 16; ++j) atomicAdd(p.dw + (((c0 + j) * 32 + c) * 4 + ty) * 4 + kw, v[j]);
    } else if (row == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) atomicAdd(p.db + c0 + j, v[j]);
    }
  }
};

struct TConv1Wgrad {   // grid.y = 2 kh2 slices (+1 ones slice -> db1); rows = (kw2, c, dy, dx); stage = 4 output rows = 80 pixels
  static constexpr int BN = 64, STAGES = 3, KROWS = 80;
  static constexpr bool A_MN = true, B_MN = true, ZERO_INIT = true;
  struct Params { SRL_TMAP xs4; SRL_TMAP dym; float* dw; float* db; int NB; int fps; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.xs4); tma_prefetch_desc(&p.dym); }
  SRL_DEVINL static int num_kblocks(const Params& p, int tm, int) { return 5 * max(0, min(p.NB, (tm + 1) * p.fps) - tm * p.fps); }
  SRL_DEVINL static void init_smem(const Params&, int, int ty, uint8_t* smem, int stage_bytes, int tid) {
    if (ty == 2)
      for (int s = 0; s < STAGES; ++s) fill_ones(smem + s * stage_bytes, 2 * KROWS * 128, tid);
  }
  SRL_DEVINL static void issue(const Params& p, int tm, int ty, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    const int f = kb / 5, q = kb - f * 5, n = tm * p.fps + f;
    mbar_arrive_expect_tx(bar, (ty < 2 ? 3 : 1) * 80 * 128);
    if (ty < 2) {
      tma_load_4d(sA, &p.xs4, bar, 0, 0, 4 * q + ty, n);
      tma_load_4d(sA + KROWS * 128, &p.xs4, bar, 0, 1, 4 * q + ty, n);
    }
    tma_load_2d(sB, &p.dym, bar, 0, n * 400 + q * 80);
  }
  SRL_DEVINL static void epilogue16(const Params& p, int, int ty, int row, int c0, float (&v)[16]) {
    if (c0 >= 32) return;
    if (ty < 2) {
      const int kw2 = row >> 6, q = row & 63, c = q >> 4, dy = (q >> 2) & 3, dx = q & 3;
      const int k = c * 64 + (4 * ty + dy) * 8 + 4 * kw2 + dx;
#pragma unroll
      for (int j = 0; j < 16; ++j) atomicAdd(p.dw + (c0 + j) * 256 + k, v[j] * (1.0f / 255.0f));
    } else if (row == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) atomicAdd(p.db + c0 + j, v[j]);
    }
  }
};

struct TFcWgrad {   // grid = (1, 4*50): ty = hw*4 + jt, hw == 49 is the ones slice (B = ones -> dbfc); stage = 64 frames
  static constexpr int BN = 64, STAGES = 4, KROWS = 64;
  static constexpr bool A_MN = true, B_MN = true, ZERO_INIT = true;
  struct Params { SRL_TMAP dhm; SRL_TMAP a3m; float* dw; float* db; int M; };
  SRL_DEVINL static void prefetch(const Params& p) { tma_prefetch_desc(&p.dhm); tma_prefetch_desc(&p.a3m); }
  SRL_DEVINL static int num_kblocks(const Params& p, int, int) { return (p.M + 63) >> 6; }
  SRL_DEVINL static void init_smem(const Params&, int, int ty, uint8_t* smem, int stage_bytes, int tid) {
    if ((ty >> 2) == 49)
      for (int s = 0; s < STAGES; ++s) fill_ones(smem + s * stage_bytes + 2 * KROWS * 128, KROWS * 128, tid);
  }
  SRL_DEVINL static void issue(const Params& p, int, int ty, int kb, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    const int hw = ty >> 2, j0 = (ty & 3) * 128;
    mbar_arrive_expect_tx(bar, (hw < 49 ? 3 : 2) * 64 * 128);
    tma_load_2d(sA, &p.dhm, bar, j0, kb * 64);
    tma_load_2d(sA + KROWS * 128, &p.dhm, bar, j0 + 64, kb * 64);
    if (hw < 49) tma_load_2d(sB, &p.a3m, bar, hw * 64, kb * 64);
  }
  SRL_DEVINL static void epilogue16(const Params& p, int, int ty, int row, int c0, float (&v)[16]) {
    const int j = (ty & 3) * 128 + row, hw = ty >> 2;
    if (hw < 49) {
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) p.dw[(size_t)j * 3136 + (c0 + jj) * 49 + hw] = v[jj];
    } else if (c0 == 0) {
      p.db[j] = v[0];
    }
  }
};

}  // namespace srl
