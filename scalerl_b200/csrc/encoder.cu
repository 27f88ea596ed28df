// Encoder forward / backward drivers: weight packing + the sequence of tcgen05 implicit-GEMM launches.
#include "encoder_problems.cuh"
#include "kernels.h"

namespace srl {

// fp32 master parameters (PyTorch layouts) -> bf16 operand copies in the layouts the GEMMs consume.
__global__ void __launch_bounds__(256) pack_weights_kernel(ParamPtrs p, bf16* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < WPack::TOTAL; i += (int64_t)gridDim.x * blockDim.x) {
    float v;
    if (i < WPack::W2K) {                       // w1k[co][(kh2*2+kw2)*64 + c*16 + dy*4 + dx] = W1[co][c][4kh2+dy][4kw2+dx]
      const int e = (int)(i - WPack::W1K), co = e >> 8, k = e & 255, tap = k >> 6, q = k & 63;
      const int c = q >> 4, dy = (q >> 2) & 3, dx = q & 3, kh = 4 * (tap >> 1) + dy, kw = 4 * (tap & 1) + dx;
      v = p.w1[co * 256 + c * 64 + kh * 8 + kw];
    } else if (i < WPack::W3K) {                // w2k[co][(kh*4+kw)*32 + c]
      const int e = (int)(i - WPack::W2K), co = e >> 9, k = e & 511, tap = k >> 5, c = k & 31;
      v = p.w2[((co * 32 + c) << 4) + tap];
    } else if (i < WPack::WFK) {                // w3k[co][(kh*3+kw)*64 + c]
      const int e = (int)(i - WPack::W3K), co = e / 576, k = e - co * 576, tap = k >> 6, c = k & 63;
      v = p.w3[(co * 64 + c) * 9 + tap];
    } else if (i < WPack::WFD) {                // wfk[j][hw*64 + c] = Wfc[j][c*49 + hw]
      const int e = (int)(i - WPack::WFK), j = e / 3136, k = e - j * 3136, hw = k >> 6, c = k & 63;
      v = p.wf[(size_t)j * 3136 + c * 49 + hw];
    } else if (i < WPack::W3D) {                // wfd[hw*64 + c][j]
      const int e = (int)(i - WPack::WFD), row = e >> 9, j = e & 511, hw = row >> 6, c = row & 63;
      v = p.wf[(size_t)j * 3136 + c * 49 + hw];
    } else if (i < WPack::W2D) {                // w3d[c][(kh*3+kw)*64 + co]
      const int e = (int)(i - WPack::W3D), c = e / 576, k = e - c * 576, tap = k >> 6, co = k & 63;
      v = p.w3[(co * 64 + c) * 9 + tap];
    } else {                                    // w2d[cls][c][(kh'*2+kw')*64 + co], kh = ph + 2kh', kw = pw + 2kw'
      const int e = (int)(i - WPack::W2D), cls = e >> 13, r = e & 8191, c = r >> 8, k = r & 255, t = k >> 6, co = k & 63;
      const int kh = (cls >> 1) + 2 * (t >> 1), kw = (cls & 1) + 2 * (t & 1);
      v = p.w2[((co * 32 + c) << 4) + kh * 4 + kw];
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

// u8 NCHW frames -> space-to-depth bf16 NHWC: xs[n][Y][X][c*16+dy*4+dx] = obs[n][c][4Y+dy][4X+dx]  (exact: u8 fits bf16).
// One thread moves one u32 (4 x dx) -> 4 bf16 (8 B); consecutive threads write consecutive 8 B.
__global__ void __launch_bounds__(256) obs_s2d_kernel(const uint8_t* __restrict__ obs, bf16* __restrict__ xs, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i & 15);               // (c, dy)
    const int64_t pix = i >> 4;                // (n, Y, X)
    const int X = (int)(pix % 21);
    const int64_t t = pix / 21;
    const int Y = (int)(t % 21);
    const int64_t n = t / 21;
    const int c = g >> 2, dy = g & 3;
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(obs + n * 28224 + c * 7056 + (4 * Y + dy) * 84 + 4 * X));
    const float f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540)) - 8388608.f;
    const float f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7541)) - 8388608.f;
    const float f2 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7542)) - 8388608.f;
    const float f3 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7543)) - 8388608.f;
    *reinterpret_cast<uint2*>(xs + i * 4) = make_uint2(pack_bf16x2(f0, f1), pack_bf16x2(f2, f3));
  }
}

cudaError_t launch_pack_weights(const ParamPtrs& p, bf16* wpack, cudaStream_t st) {
  pack_weights_kernel<<<1184, 256, 0, st>>>(p, wpack);
  return cudaGetLastError();
}

#define SRL_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

cudaError_t encoder_forward(const uint8_t* obs, int frames, const ParamPtrs& p, const EncoderBuffers& buf, bool simt, cudaStream_t st,
                            const Profiler& pf) {
  if (frames <= 0) return cudaSuccess;
  { const int64_t total = (int64_t)frames * 441 * 16;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    pf.b(PS_S2D); obs_s2d_kernel<<<blocks, 256, 0, st>>>(obs, buf.xs, total); SRL_TRY(cudaGetLastError()); pf.e(PS_S2D); }
  { Conv1Fwd::Params q{buf.xs, buf.wpack + WPack::W1K, p.b1, buf.a1, frames * 400};
    pf.b(PS_CONV1_FWD); SRL_TRY(igemm_launch<Conv1Fwd>(q, dim3(cdiv(q.M, 128), 1), st, simt)); pf.e(PS_CONV1_FWD); }
  { Conv2Fwd::Params q{buf.a1, buf.wpack + WPack::W2K, p.b2, buf.a2, frames * 81};
    pf.b(PS_CONV2_FWD); SRL_TRY(igemm_launch<Conv2Fwd>(q, dim3(cdiv(q.M, 128), 1), st, simt)); pf.e(PS_CONV2_FWD); }
  { Conv3Fwd::Params q{buf.a2, buf.wpack + WPack::W3K, p.b3, buf.a3, frames * 49};
    pf.b(PS_CONV3_FWD); SRL_TRY(igemm_launch<Conv3Fwd>(q, dim3(cdiv(q.M, 128), 1), st, simt)); pf.e(PS_CONV3_FWD); }
  { FcFwd::Params q{buf.a3, buf.wpack + WPack::WFK, buf.hpart, frames};
    static_assert(FcFwd::FC_SPLITS == FC_SPLITS, "split count");
    pf.b(PS_FC_FWD); SRL_TRY(igemm_launch<FcFwd>(q, dim3(cdiv(q.M, 128), 8 * FC_SPLITS), st, simt)); pf.e(PS_FC_FWD); }
  return cudaSuccess;
}

// split the contraction range P into ~target CTAs' worth of 64-aligned pieces
static inline void split_k(int P, int target, int* pps, int* nsplit) {
  int per = cdiv(cdiv(P, 64), target) * 64;
  if (per < 64) per = 64;
  *pps = per;
  *nsplit = cdiv(P, per);
}

cudaError_t encoder_backward(const uint8_t* obs, int frames, const EncoderBuffers& buf, const ParamPtrs& g, bool simt, cudaStream_t st,
                             const Profiler& pf) {
  if (frames <= 0) return cudaSuccess;
  int pps, ns;
  // ---- fc: bias, wgrad, dgrad
  { FcWgrad::Params q{buf.dh, buf.a3, g.wf, g.bf, frames};
    pf.b(PS_FC_WGRAD); SRL_TRY(igemm_launch<FcWgrad>(q, dim3(1, 4 * 49), st, simt)); pf.e(PS_FC_WGRAD); }
  { FcDgrad::Params q{buf.dh, buf.wpack + WPack::WFD, buf.a3, buf.da3, frames};
    pf.b(PS_FC_DGRAD); SRL_TRY(igemm_launch<FcDgrad>(q, dim3(cdiv(frames, 128), 49), st, simt)); pf.e(PS_FC_DGRAD); }
  // ---- conv3
  { split_k(frames * 49, 29, &pps, &ns);
    Conv3Wgrad::Params q{buf.a2, buf.da3, g.w3, g.b3, frames * 49, pps};
    pf.b(PS_CONV3_WGRAD); SRL_TRY(igemm_launch<Conv3Wgrad>(q, dim3(ns, 5), st, simt)); pf.e(PS_CONV3_WGRAD); }
  { Conv3Dgrad::Params q{buf.da3, buf.wpack + WPack::W3D, buf.a2, buf.da2, frames * 81};
    pf.b(PS_CONV3_DGRAD); SRL_TRY(igemm_launch<Conv3Dgrad>(q, dim3(cdiv(q.M, 128), 1), st, simt)); pf.e(PS_CONV3_DGRAD); }
  // ---- conv2
  { split_k(frames * 81, 37, &pps, &ns);
    Conv2Wgrad::Params q{buf.a1, buf.da2, g.w2, g.b2, frames * 81, pps};
    pf.b(PS_CONV2_WGRAD); SRL_TRY(igemm_launch<Conv2Wgrad>(q, dim3(ns, 4), st, simt)); pf.e(PS_CONV2_WGRAD); }
  { Conv2Dgrad::Params q{buf.da2, buf.wpack + WPack::W2D, buf.a1, buf.da1, frames * 100};
    pf.b(PS_CONV2_DGRAD); SRL_TRY(igemm_launch<Conv2Dgrad>(q, dim3(cdiv(q.M, 128), 4), st, simt)); pf.e(PS_CONV2_DGRAD); }
  // ---- conv1 (no dgrad: the frame is the network input)
  { split_k(frames * 400, 74, &pps, &ns);
    Conv1Wgrad::Params q{buf.xs, buf.da1, g.w1, g.b1, frames * 400, pps};
    pf.b(PS_CONV1_WGRAD); SRL_TRY(igemm_launch<Conv1Wgrad>(q, dim3(ns, 2), st, simt)); pf.e(PS_CONV1_WGRAD); }
  return cudaSuccess;
}

cudaError_t test_gemm(const void* A, const void* B, float* D, int M, int N, int K, bool mn_major, bool simt, cudaStream_t st) {
  if (mn_major) {
    TestGemmMN::Params q{(const bf16*)A, (const bf16*)B, D, M, N, K};
    return igemm_launch<TestGemmMN>(q, dim3(M / 128, N / 64), st, simt);
  }
  TestGemmK::Params q{(const bf16*)A, (const bf16*)B, D, M, N, K};
  return igemm_launch<TestGemmK>(q, dim3(cdiv(M, 128), N / 64), st, simt);
}

}  // namespace srl
