// Problem definitions (operand gathers + epilogues) for every tensor-core GEMM of the AtariNet
// encoder (reference: scalerl/algorithms/utils/atari_model.py:30-46,93-101 and its autograd
// backward, impala_atari.py:343).  All activations are NHWC bf16; weights are pre-packed bf16
// copies of the fp32 master parameters (see pack_weights_kernel in encoder.cu for the layouts).
//
//   forward : Conv1Fwd (u8 NCHW frame -> a1), Conv2Fwd, Conv3Fwd, FcFwd            K-major operands
//   dgrad   : FcDgrad, Conv3Dgrad, Conv2Dgrad (4 stride-parity classes)            K-major operands
//   wgrad   : FcWgrad, Conv3Wgrad, Conv2Wgrad, Conv1Wgrad                          MN-major operands
//             (contraction over pixels/frames; NHWC rows are used as they lie in memory)
#pragma once
#include "igemm.cuh"

namespace srl {

typedef __nv_bfloat16 bf16;

SRL_DEVINL uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
SRL_DEVINL uint4 zero16() { return make_uint4(0, 0, 0, 0); }

// 8 consecutive u8 (two aligned u32 words) -> 8 bf16 (exact: 0..255 fit the 8-bit significand)
SRL_DEVINL uint4 u8x8_to_bf16x8(uint32_t w0, uint32_t w1) {
  float f[8];
  f[0] = __uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7540)) - 8388608.f;
  f[1] = __uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7541)) - 8388608.f;
  f[2] = __uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7542)) - 8388608.f;
  f[3] = __uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7543)) - 8388608.f;
  f[4] = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7540)) - 8388608.f;
  f[5] = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7541)) - 8388608.f;
  f[6] = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7542)) - 8388608.f;
  f[7] = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7543)) - 8388608.f;
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

SRL_DEVINL void store_bf16x16(bf16* dst, const float (&v)[16]) {
  uint4 a = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  uint4 b = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
  reinterpret_cast<uint4*>(dst)[0] = a;
  reinterpret_cast<uint4*>(dst)[1] = b;
}
// v[j] *= (mask[j] > 0) for 16 bf16 mask values
SRL_DEVINL void relu_mask16(const bf16* mask, float (&v)[16]) {
  uint4 a = ldg16(mask), b = ldg16(mask + 8);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (!(bf16_lo(w[i]) > 0.f)) v[2 * i] = 0.f;
    if (!(bf16_hi(w[i]) > 0.f)) v[2 * i + 1] = 0.f;
  }
}

// ============================================================================================
// forward
// ============================================================================================
// conv1 (8x8 s4 over the u8 frame == 2x2 s1 over the space-to-depth bf16 frame xs[N,21,21,64], 64 = (c,dy,dx)),
// conv2 (4x4 s2, 32->64) and conv3 (3x3 s1, 64->64): K ordered (kh, kw, c); one K-block = 64 contiguous bf16.
// SCALE255: conv1 multiplies the fp32 accumulator by 1/255 (the reference normalises its input, atari_model.py:94).
template <int IH, int OH, int CIN, int COUT, int KH, int KW, int STRIDE, bool SCALE255>
struct ConvFwd {   // atari_model.py:30-43,97-99
  static constexpr int BN = COUT, STAGES = COUT <= 32 ? 4 : 3;
  static constexpr bool A_MN = false, B_MN = false;
  static constexpr int BIAS = 0, BIAS_N = 0;
  static constexpr int KTOT = KH * KW * CIN;
  static constexpr int TAPS_PER_KB = 64 / CIN;       // conv2: 2 kw taps per K-block; conv1/conv3: 1
  static constexpr int KB_PER_KH = KW / TAPS_PER_KB;
  struct Params { const bf16* in; const bf16* w; const float* bias; bf16* out; int M; };
  typedef int RowA;
  typedef const bf16* RowB;
  SRL_DEVINL static int num_kblocks(const Params&, int, int) { return KTOT / 64; }
  SRL_DEVINL static RowA make_rowA(const Params& p, int tm, int, int srow) {
    const int m = tm * 128 + srow;
    if (m >= p.M) return -1;
    const int n = m / (OH * OH), r = m - n * (OH * OH), oh = r / OH, ow = r - oh * OH;
    return ((n * IH + oh * STRIDE) * IH + ow * STRIDE) * CIN;
  }
  SRL_DEVINL static RowB make_rowB(const Params& p, int, int, int srow) { return p.w + srow * KTOT; }
  SRL_DEVINL static uint4 load_A(const Params& p, RowA base, int kb, int chunk) {
    if (base < 0) return zero16();
    const int kh = kb / KB_PER_KH, kw0 = (kb - kh * KB_PER_KH) * TAPS_PER_KB;
    return ldg16(p.in + base + (kh * IH + kw0) * CIN + chunk * 8);
  }
  SRL_DEVINL static uint4 load_B(const Params&, RowB w, int kb, int chunk) { return ldg16(w + kb * 64 + chunk * 8); }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int, int row, int c0, float (&v)[16]) {
    const int m = tm * 128 + row;
    if (m >= p.M) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], SCALE255 ? 1.0f / 255.0f : 1.0f, __ldg(p.bias + c0 + j)), 0.f);
    store_bf16x16(p.out + (size_t)m * COUT + c0, v);
  }
};
typedef ConvFwd<21, 20, 64, 32, 2, 2, 1, true> Conv1Fwd;
typedef ConvFwd<20, 9, 32, 64, 4, 4, 2, false> Conv2Fwd;
typedef ConvFwd<9, 7, 64, 64, 3, 3, 1, false> Conv3Fwd;

struct FcFwd {   // atari_model.py:46,100 : split-K partials hpart[split][m][512] = a3_flat @ Wfc^T (bias + ReLU are applied by
                  // head_fwd_kernel, which reduces the FC_SPLITS partials);  grid.y = (512/64) * FC_SPLITS,  ty = nt*FC_SPLITS + split
  static constexpr int BN = 64, STAGES = 3, FC_SPLITS = 4;
  static constexpr bool A_MN = false, B_MN = false;
  static constexpr int BIAS = 0, BIAS_N = 0;
  struct Params { const bf16* in; const bf16* w; float* out; int M; };
  typedef int RowA;
  typedef const bf16* RowB;
  SRL_DEVINL static int kb_begin(int split) { return (49 * split) / FC_SPLITS; }
  SRL_DEVINL static int num_kblocks(const Params&, int, int ty) { const int sp = ty % FC_SPLITS; return kb_begin(sp + 1) - kb_begin(sp); }
  SRL_DEVINL static RowA make_rowA(const Params& p, int tm, int ty, int srow) {
    const int m = tm * 128 + srow;
    return m < p.M ? m * 3136 + kb_begin(ty % FC_SPLITS) * 64 : -1;
  }
  SRL_DEVINL static RowB make_rowB(const Params& p, int, int ty, int srow) {
    return p.w + (size_t)((ty / FC_SPLITS) * 64 + srow) * 3136 + kb_begin(ty % FC_SPLITS) * 64;
  }
  SRL_DEVINL static uint4 load_A(const Params& p, RowA base, int kb, int chunk) {
    return base < 0 ? zero16() : ldg16(p.in + base + kb * 64 + chunk * 8);
  }
  SRL_DEVINL static uint4 load_B(const Params&, RowB w, int kb, int chunk) { return ldg16(w + kb * 64 + chunk * 8); }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int ty, int row, int c0, float (&v)[16]) {
    const int m = tm * 128 + row;
    if (m >= p.M) return;
    float4* o = reinterpret_cast<float4*>(p.out + ((size_t)(ty % FC_SPLITS) * p.M + m) * 512 + (ty / FC_SPLITS) * 64 + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
};

// ============================================================================================
// dgrad (all K-major; the ReLU of the producing layer is applied as a mask in the epilogue)
// ============================================================================================
struct FcDgrad {   // da3[m][i] = (sum_j dh[m][j] * Wfc[j][i]) * (a3 > 0);  grid.y = 3136/64
  static constexpr int BN = 64, STAGES = 3;
  static constexpr bool A_MN = false, B_MN = false;
  static constexpr int BIAS = 0, BIAS_N = 0;
  struct Params { const bf16* dh; const bf16* wd; const bf16* a3; bf16* da3; int M; };
  typedef int RowA;
  typedef const bf16* RowB;
  SRL_DEVINL static int num_kblocks(const Params&, int, int) { return 8; }
  SRL_DEVINL static RowA make_rowA(const Params& p, int tm, int, int srow) {
    const int m = tm * 128 + srow;
    return m < p.M ? m * 512 : -1;
  }
  SRL_DEVINL static RowB make_rowB(const Params& p, int, int ty, int srow) { return p.wd + (size_t)(ty * 64 + srow) * 512; }
  SRL_DEVINL static uint4 load_A(const Params& p, RowA base, int kb, int chunk) {
    return base < 0 ? zero16() : ldg16(p.dh + base + kb * 64 + chunk * 8);
  }
  SRL_DEVINL static uint4 load_B(const Params&, RowB w, int kb, int chunk) { return ldg16(w + kb * 64 + chunk * 8); }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int ty, int row, int c0, float (&v)[16]) {
    const int m = tm * 128 + row;
    if (m >= p.M) return;
    const size_t idx = (size_t)m * 3136 + ty * 64 + c0;
    relu_mask16(p.a3 + idx, v);
    store_bf16x16(p.da3 + idx, v);
  }
};

struct Conv3Dgrad {   // da2[n,ih,iw,c] = sum_{kh,kw,co} da3[n,ih-kh,iw-kw,co] W3[co][c][kh][kw], masked by a2>0
  static constexpr int BN = 64, STAGES = 3;
  static constexpr bool A_MN = false, B_MN = false;
  static constexpr int BIAS = 0, BIAS_N = 0;
  struct Params { const bf16* dy; const bf16* wd; const bf16* act; bf16* dx; int M; };   // M = frames*81
  typedef int RowA;   // (n*7*7) << 8 | ih << 4 | iw, or -1
  typedef const bf16* RowB;
  SRL_DEVINL static int num_kblocks(const Params&, int, int) { return 9; }
  SRL_DEVINL static RowA make_rowA(const Params& p, int tm, int, int srow) {
    const int m = tm * 128 + srow;
    if (m >= p.M) return -1;
    const int n = m / 81, r = m - n * 81, ih = r / 9, iw = r - ih * 9;
    return ((n * 49) << 8) | (ih << 4) | iw;
  }
  SRL_DEVINL static RowB make_rowB(const Params& p, int, int, int srow) { return p.wd + srow * 576; }
  SRL_DEVINL static uint4 load_A(const Params& p, RowA rc, int kb, int chunk) {
    if (rc < 0) return zero16();
    const int kh = kb / 3, kw = kb - kh * 3;
    const int oh = ((rc >> 4) & 15) - kh, ow = (rc & 15) - kw;
    if ((unsigned)oh >= 7u || (unsigned)ow >= 7u) return zero16();
    return ldg16(p.dy + (size_t)((rc >> 8) + oh * 7 + ow) * 64 + chunk * 8);
  }
  SRL_DEVINL static uint4 load_B(const Params&, RowB w, int kb, int chunk) { return ldg16(w + kb * 64 + chunk * 8); }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int, int row, int c0, float (&v)[16]) {
    const int m = tm * 128 + row;
    if (m >= p.M) return;
    const size_t idx = (size_t)m * 64 + c0;
    relu_mask16(p.act + idx, v);
    store_bf16x16(p.dx + idx, v);
  }
};

struct Conv2Dgrad {   // stride-2 transposed conv split in 4 parity classes (grid.y = ph*2+pw); K = (kh',kw',co) = 256
  static constexpr int BN = 32, STAGES = 4;
  static constexpr bool A_MN = false, B_MN = false;
  static constexpr int BIAS = 0, BIAS_N = 0;
  struct Params { const bf16* dy; const bf16* wd; const bf16* act; bf16* dx; int M; };   // M = frames*100 per class
  typedef int RowA;   // (n*81) << 8 | i' << 4 | j'
  typedef const bf16* RowB;
  SRL_DEVINL static int num_kblocks(const Params&, int, int) { return 4; }
  SRL_DEVINL static RowA make_rowA(const Params& p, int tm, int, int srow) {
    const int m = tm * 128 + srow;
    if (m >= p.M) return -1;
    const int n = m / 100, r = m - n * 100, i = r / 10, j = r - i * 10;
    return ((n * 81) << 8) | (i << 4) | j;
  }
  SRL_DEVINL static RowB make_rowB(const Params& p, int, int ty, int srow) { return p.wd + (ty * 32 + srow) * 256; }
  SRL_DEVINL static uint4 load_A(const Params& p, RowA rc, int kb, int chunk) {
    if (rc < 0) return zero16();
    const int oh = ((rc >> 4) & 15) - (kb >> 1), ow = (rc & 15) - (kb & 1);
    if ((unsigned)oh >= 9u || (unsigned)ow >= 9u) return zero16();
    return ldg16(p.dy + (size_t)((rc >> 8) + oh * 9 + ow) * 64 + chunk * 8);
  }
  SRL_DEVINL static uint4 load_B(const Params&, RowB w, int kb, int chunk) { return ldg16(w + kb * 64 + chunk * 8); }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int ty, int row, int c0, float (&v)[16]) {
    const int m = tm * 128 + row;
    if (m >= p.M) return;
    const int n = m / 100, r = m - n * 100, i = r / 10, j = r - i * 10;
    const int ih = 2 * i + (ty >> 1), iw = 2 * j + (ty & 1);
    const size_t pix = (size_t)(n * 20 + ih) * 20 + iw;
    relu_mask16(p.act + pix * 32 + c0, v);
    store_bf16x16(p.dx + pix * 64 + c0, v);      // da1 has a 64-channel pitch (channels 32..63 stay zero)
  }
};

// ============================================================================================
// wgrad (MN-major operands: smem row = one pixel / frame of the contraction, 64 channels wide)
//   CTA (blockIdx.x = split s of the contraction range, blockIdx.y = which 128-row slice of dW)
//   conv wgrads accumulate with fp32 atomics into the (pre-zeroed) PyTorch-layout gradient.
// ============================================================================================
struct Conv3Wgrad {   // dW3[co][c][kh][kw] = sum_p da3[p][co] * a2[n,oh+kh,ow+kw,c];  grid.y = 5 tap pairs; db3 = colsum(da3)
  static constexpr int BN = 64, STAGES = 3;
  static constexpr bool A_MN = true, B_MN = true;
  static constexpr int BIAS = 2, BIAS_N = 64;
  struct Params { const bf16* act; const bf16* dy; float* dw; float* db; int P; int pps; };   // P = frames*49, pps % 64 == 0
  SRL_DEVINL static float* bias_dst(const Params& p, int, int ty) { return ty == 0 ? p.db : nullptr; }
  struct RowA { int krow; int tapoff; };   // tapoff = (kh*9+kw)*64 or -1 (tap 9 does not exist)
  typedef int RowB;
  SRL_DEVINL static int num_kblocks(const Params& p, int tm, int) {
    const int lo = tm * p.pps, hi = min(p.P, lo + p.pps);
    return hi > lo ? (hi - lo + 63) >> 6 : 0;
  }
  SRL_DEVINL static RowA make_rowA(const Params&, int, int ty, int srow) {
    const int tap = 2 * ty + (srow >> 6);
    RowA r; r.krow = srow & 63; r.tapoff = tap < 9 ? ((tap / 3) * 9 + tap % 3) * 64 : -1;
    return r;
  }
  SRL_DEVINL static RowB make_rowB(const Params&, int, int, int srow) { return srow; }
  SRL_DEVINL static uint4 load_A(const Params& p, const RowA& r, int kb, int chunk) {
    const int q = blockIdx.x * p.pps + kb * 64 + r.krow;
    if (r.tapoff < 0 || q >= p.P) return zero16();
    const int n = q / 49, s = q - n * 49, oh = s / 7, ow = s - oh * 7;
    return ldg16(p.act + (size_t)((n * 9 + oh) * 9 + ow) * 64 + r.tapoff + chunk * 8);
  }
  SRL_DEVINL static uint4 load_B(const Params& p, RowB krow, int kb, int chunk) {
    const int q = blockIdx.x * p.pps + kb * 64 + krow;
    return q < p.P ? ldg16(p.dy + (size_t)q * 64 + chunk * 8) : zero16();
  }
  SRL_DEVINL static void epilogue16(const Params& p, int, int ty, int row, int c0, float (&v)[16]) {
    const int tap = 2 * ty + (row >> 6), c = row & 63;
    if (tap >= 9) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) atomicAdd(p.dw + ((c0 + j) * 64 + c) * 9 + tap, v[j]);
  }
};

struct Conv2Wgrad {   // dW2[co][c][kh][kw] = sum_p da2[p][co] * a1[n,2oh+kh,2ow+kw,c];  grid.y = kh, rows = (kw, c); db2 = colsum(da2)
  static constexpr int BN = 64, STAGES = 3;
  static constexpr bool A_MN = true, B_MN = true;
  static constexpr int BIAS = 2, BIAS_N = 64;
  struct Params { const bf16* act; const bf16* dy; float* dw; float* db; int P; int pps; };   // P = frames*81
  SRL_DEVINL static float* bias_dst(const Params& p, int, int ty) { return ty == 0 ? p.db : nullptr; }
  struct RowA { int krow; int off; };    // off = kh*20*32 + block*64
  typedef int RowB;
  SRL_DEVINL static int num_kblocks(const Params& p, int tm, int) {
    const int lo = tm * p.pps, hi = min(p.P, lo + p.pps);
    return hi > lo ? (hi - lo + 63) >> 6 : 0;
  }
  SRL_DEVINL static RowA make_rowA(const Params&, int, int ty, int srow) {
    RowA r; r.krow = srow & 63; r.off = ty * 640 + (srow >> 6) * 64;
    return r;
  }
  SRL_DEVINL static RowB make_rowB(const Params&, int, int, int srow) { return srow; }
  SRL_DEVINL static uint4 load_A(const Params& p, const RowA& r, int kb, int chunk) {
    const int q = blockIdx.x * p.pps + kb * 64 + r.krow;
    if (q >= p.P) return zero16();
    const int n = q / 81, s = q - n * 81, oh = s / 9, ow = s - oh * 9;
    return ldg16(p.act + (size_t)((n * 20 + 2 * oh) * 20 + 2 * ow) * 32 + r.off + chunk * 8);
  }
  SRL_DEVINL static uint4 load_B(const Params& p, RowB krow, int kb, int chunk) {
    const int q = blockIdx.x * p.pps + kb * 64 + krow;
    return q < p.P ? ldg16(p.dy + (size_t)q * 64 + chunk * 8) : zero16();
  }
  SRL_DEVINL static void epilogue16(const Params& p, int, int ty, int row, int c0, float (&v)[16]) {
    const int kw = row >> 5, c = row & 31;
#pragma unroll
    for (int j = 0; j < 16; ++j) atomicAdd(p.dw + (((c0 + j) * 32 + c) * 4 + ty) * 4 + kw, v[j]);
  }
};

struct Conv1Wgrad {   // dW1[co][c][4kh2+dy][4kw2+dx] = (1/255) sum_p da1[p][co] * xs[n,oh+kh2,ow+kw2,(c,dy,dx)];  grid.y = kh2
  static constexpr int BN = 64, STAGES = 3;   // N padded 32 -> 64 (upper half zero) to keep the 128 B row form
  static constexpr bool A_MN = true, B_MN = true;
  static constexpr int BIAS = 2, BIAS_N = 32;   // db1 = colsum(da1)
  struct Params { const bf16* xs; const bf16* dy; float* dw; float* db; int P; int pps; };   // P = frames*400
  struct RowA { int krow; int tapoff; };   // tapoff = (kh2*21 + kw2)*64
  typedef int RowB;
  SRL_DEVINL static float* bias_dst(const Params& p, int, int ty) { return ty == 0 ? p.db : nullptr; }
  SRL_DEVINL static int num_kblocks(const Params& p, int tm, int) {
    const int lo = tm * p.pps, hi = min(p.P, lo + p.pps);
    return hi > lo ? (hi - lo + 63) >> 6 : 0;
  }
  SRL_DEVINL static RowA make_rowA(const Params&, int, int ty, int srow) {
    RowA r; r.krow = srow & 63; r.tapoff = (ty * 21 + (srow >> 6)) * 64;
    return r;
  }
  SRL_DEVINL static RowB make_rowB(const Params&, int, int, int srow) { return srow; }
  SRL_DEVINL static uint4 load_A(const Params& p, const RowA& r, int kb, int chunk) {
    const int q = blockIdx.x * p.pps + kb * 64 + r.krow;
    if (q >= p.P) return zero16();
    const int n = q / 400, s = q - n * 400, oh = s / 20, ow = s - oh * 20;
    return ldg16(p.xs + (size_t)((n * 21 + oh) * 21 + ow) * 64 + r.tapoff + chunk * 8);
  }
  SRL_DEVINL static uint4 load_B(const Params& p, RowB krow, int kb, int chunk) {
    const int q = blockIdx.x * p.pps + kb * 64 + krow;
    return (q < p.P && chunk < 4) ? ldg16(p.dy + (size_t)q * 64 + chunk * 8) : zero16();
  }
  SRL_DEVINL static void epilogue16(const Params& p, int, int ty, int row, int c0, float (&v)[16]) {
    if (c0 >= 32) return;
    const int kw2 = row >> 6, q = row & 63, c = q >> 4, dy = (q >> 2) & 3, dx = q & 3;
    const int k = c * 64 + (4 * ty + dy) * 8 + 4 * kw2 + dx;
#pragma unroll
    for (int j = 0; j < 16; ++j) atomicAdd(p.dw + (c0 + j) * 256 + k, v[j] * (1.0f / 255.0f));
  }
};

struct FcWgrad {   // dWfc[j][c*49+hw] = sum_m dh[m][j] * a3[m][hw*64+c];  grid = (1, 4*49): ty = hw*4 + jt; dbfc = colsum(dh)
  static constexpr int BN = 64, STAGES = 3;
  static constexpr bool A_MN = true, B_MN = true;
  static constexpr int BIAS = 1, BIAS_N = 128;
  struct Params { const bf16* dh; const bf16* a3; float* dw; float* db; int M; };
  SRL_DEVINL static float* bias_dst(const Params& p, int, int ty) { return (ty >> 2) == 0 ? p.db + (ty & 3) * 128 : nullptr; }
  struct RowA { int krow; int joff; };
  typedef int RowB;
  SRL_DEVINL static int num_kblocks(const Params& p, int, int) { return (p.M + 63) >> 6; }
  SRL_DEVINL static RowA make_rowA(const Params&, int, int ty, int srow) {
    RowA r; r.krow = srow & 63; r.joff = (ty & 3) * 128 + (srow >> 6) * 64;
    return r;
  }
  SRL_DEVINL static RowB make_rowB(const Params&, int, int, int srow) { return srow; }
  SRL_DEVINL static uint4 load_A(const Params& p, const RowA& r, int kb, int chunk) {
    const int m = kb * 64 + r.krow;
    return m < p.M ? ldg16(p.dh + (size_t)m * 512 + r.joff + chunk * 8) : zero16();
  }
  SRL_DEVINL static uint4 load_B(const Params& p, RowB krow, int kb, int chunk) {
    const int m = kb * 64 + krow;
    return m < p.M ? ldg16(p.a3 + (size_t)m * 3136 + (blockIdx.y >> 2) * 64 + chunk * 8) : zero16();
  }
  SRL_DEVINL static void epilogue16(const Params& p, int, int ty, int row, int c0, float (&v)[16]) {
    const int j = (ty & 3) * 128 + row, hw = ty >> 2;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) p.dw[(size_t)j * 3136 + (c0 + jj) * 49 + hw] = v[jj];
  }
};

// ============================================================================================
// plain GEMM problems used by the unit tests to validate descriptors / pipeline in isolation
// ============================================================================================
struct TestGemmK {    // D[M][N] = A[M][K] * B[N][K]^T, K % 64 == 0, N % 64 == 0; grid = (ceil(M/128), N/64)
  static constexpr int BN = 64, STAGES = 3;
  static constexpr bool A_MN = false, B_MN = false;
  static constexpr int BIAS = 0, BIAS_N = 0;
  struct Params { const bf16* A; const bf16* B; float* D; int M, N, K; };
  typedef int RowA;
  typedef int RowB;
  SRL_DEVINL static int num_kblocks(const Params& p, int, int) { return p.K / 64; }
  SRL_DEVINL static RowA make_rowA(const Params& p, int tm, int, int srow) { const int m = tm * 128 + srow; return m < p.M ? m : -1; }
  SRL_DEVINL static RowB make_rowB(const Params&, int, int ty, int srow) { return ty * 64 + srow; }
  SRL_DEVINL static uint4 load_A(const Params& p, RowA m, int kb, int chunk) {
    return m < 0 ? zero16() : ldg16(p.A + (size_t)m * p.K + kb * 64 + chunk * 8);
  }
  SRL_DEVINL static uint4 load_B(const Params& p, RowB n, int kb, int chunk) { return ldg16(p.B + (size_t)n * p.K + kb * 64 + chunk * 8); }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int ty, int row, int c0, float (&v)[16]) {
    const int m = tm * 128 + row;
    if (m >= p.M) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) p.D[(size_t)m * p.N + ty * 64 + c0 + j] = v[j];
  }
};
struct TestGemmMN {   // D[M][N] = At[K][M]^T * Bt[K][N], M % 128 == 0, N % 64 == 0, any K; grid = (M/128, N/64)
  static constexpr int BN = 64, STAGES = 3;
  static constexpr bool A_MN = true, B_MN = true;
  static constexpr int BIAS = 0, BIAS_N = 0;
  struct Params { const bf16* At; const bf16* Bt; float* D; int M, N, K; };
  typedef int RowA;
  typedef int RowB;
  SRL_DEVINL static int num_kblocks(const Params& p, int, int) { return (p.K + 63) / 64; }
  SRL_DEVINL static RowA make_rowA(const Params&, int, int, int srow) { return srow; }
  SRL_DEVINL static RowB make_rowB(const Params&, int, int, int srow) { return srow; }
  SRL_DEVINL static uint4 load_A(const Params& p, RowA srow, int kb, int chunk) {
    const int k = kb * 64 + (srow & 63);
    return k < p.K ? ldg16(p.At + (size_t)k * p.M + blockIdx.x * 128 + (srow >> 6) * 64 + chunk * 8) : zero16();
  }
  SRL_DEVINL static uint4 load_B(const Params& p, RowB srow, int kb, int chunk) {
    const int k = kb * 64 + srow;
    return k < p.K ? ldg16(p.Bt + (size_t)k * p.N + blockIdx.y * 64 + chunk * 8) : zero16();
  }
  SRL_DEVINL static void epilogue16(const Params& p, int tm, int ty, int row, int c0, float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) p.D[(size_t)(tm * 128 + row) * p.N + ty * 64 + c0 + j] = v[j];
  }
};

}  // namespace srl
