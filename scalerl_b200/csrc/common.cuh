// Common device helpers for the sm_100a kernels: mbarrier, proxy fences, tcgen05 (alloc / mma /
// commit / ld), UMMA shared-memory + instruction descriptors, bf16 packing.
// Hand-written inline PTX; descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" /
// "instruction descriptor" tables (cross-checked against cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define SRL_DEVINL __device__ __forceinline__

#ifndef SRL_SPIN_LIMIT
#define SRL_SPIN_LIMIT (1u << 24)   // bounded mbarrier spin: a broken pipeline traps instead of hanging the GPU
#endif

namespace srl {

SRL_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
SRL_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SRL_DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SRL_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
SRL_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
SRL_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
SRL_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SRL_SPIN_LIMIT) { __trap(); }
  }
}

// Programmatic dependent launch (PDL).  A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its stream predecessor is still running: everything before pdl_wait() (barrier init, TMEM allocation, descriptor
// prefetch, loads of data that was complete long before the predecessor started) overlaps the predecessor's tail;
// pdl_wait() returns once the predecessor grid has completed and its memory is visible.  pdl_launch() lets the NEXT
// kernel in the stream begin its own prologue.  Both are no-ops for a kernel launched without the attribute.
// Diagnostics build (SRL_DEFINES=SRL_KSTAMP, tests/diag/diag_timeline.py): thread 0 of block 0 of every kernel appends {kernel id, %globaltimer at
// entry, %globaltimer when its stream predecessor had completed} to a buffer -- the in-graph timeline of a step, which no profiler here can
// show (ncu serialises the launches, per-kernel CUDA events break the programmatic dependencies).  Never defined in the product build.
#ifdef SRL_KSTAMP
static __device__ unsigned long long* g_kstamp = nullptr;      // one copy per translation unit: kstamp_set_<unit>() (kernels.h)
SRL_DEVINL unsigned long long kstamp_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
SRL_DEVINL void kstamp_put(int kid, unsigned long long t0) {
  if (g_kstamp && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    const unsigned long long t1 = kstamp_now(), i = atomicAdd(g_kstamp, 1ull);
    if (i < 2000) { g_kstamp[1 + 3 * i] = (unsigned long long)kid; g_kstamp[2 + 3 * i] = t0; g_kstamp[3 + 3 * i] = t1; }
  }
}
#define SRL_KSTAMP_SETTER(fn) void fn(unsigned long long* q) { cudaMemcpyToSymbol(g_kstamp, &q, sizeof q); }
#else
#define SRL_KSTAMP_SETTER(fn) void fn(unsigned long long*) {}
#endif
// kid: the kernel's id in the diagnostics timeline (ignored by the product build)
SRL_DEVINL void pdl_wait(int kid = 0) {
#ifdef SRL_KSTAMP
  const unsigned long long t0 = kstamp_now();
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef SRL_KSTAMP
  if (kid) kstamp_put(kid, t0);
#else
  (void)kid;
#endif
}
SRL_DEVINL void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 16-byte shared-memory load through the shared pipe (LDS), never a generic LD: keeps it in order with mbarrier operations
SRL_DEVINL uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
SRL_DEVINL void sts_volatile_f32(uint32_t saddr, float v) { asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }

// generic-proxy smem writes -> visible to the async proxy (UMMA / TMA reads)
SRL_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// tcgen05
// ------------------------------------------------------------------------------------------
SRL_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SRL_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; ncols power of two in [32,512]; the allocated base address is written to *dst (smem)
SRL_DEVINL void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
SRL_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// one lane of the (converged) warp: the form the compiler recognises as 'exactly one thread' for the tcgen05 / TMA issue paths
SRL_DEVINL uint32_t elect_one_sync() {
  uint32_t pred;
  __syncwarp();      // elect.sync needs the full warp converged
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues.
SRL_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
SRL_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: 32 lanes x 32-bit, 16 consecutive columns (thread i of the warp gets lane base+i)
SRL_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
SRL_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// descriptors
// ------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B. byte offsets are encoded >>4.
//   K-major  tile: rows (M/N index) of 128 B (64 bf16 along K); 8-row groups 1024 B apart  -> SBO=1024, LBO=16 (ignored)
//   MN-major tile: rows (K index)   of 128 B (64 bf16 along M/N); 8-row groups 1024 B apart -> SBO=1024,
//                  LBO = byte distance between successive 64-element M/N blocks
SRL_DEVINL uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// the same descriptor for a SWIZZLE_64B tile (rows of 64 B, 8-row groups 512 B apart)
SRL_DEVINL uint64_t make_smem_desc_sw64(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (make_smem_desc(saddr, lbo_bytes, sbo_bytes) & ~((uint64_t)7 << 61)) | ((uint64_t)4 << 61);
}
// kind::f16 instruction descriptor: bf16 x bf16 -> f32, dense
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                     // D format f32
         | (1u << 7)                   // A format bf16
         | (1u << 10)                  // B format bf16
         | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// 128B-swizzle: 16-byte chunk c (0..7) of 128-byte row r lands at chunk position c ^ (r & 7)
SRL_DEVINL uint32_t swz128(uint32_t row, uint32_t chunk) { return row * 128u + ((chunk ^ (row & 7u)) << 4); }
// 64B-swizzle: 16-byte chunk c (0..3) of 64-byte row r lands at chunk position c ^ ((r >> 1) & 3)   (address bits [4:5] ^= bits [7:8])
SRL_DEVINL uint32_t swz64(uint32_t row, uint32_t chunk) { return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4); }

// ------------------------------------------------------------------------------------------
// small numeric helpers
// ------------------------------------------------------------------------------------------
SRL_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
SRL_DEVINL float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
SRL_DEVINL float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// 16-byte vector reduction: 4 fp32 adds in one L2 atomic transaction (sm_90+)
SRL_DEVINL void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
SRL_DEVINL void red_add_16(float* dst, const float (&v)[16], float scale = 1.0f) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) red_add_v4(dst + j, v[j] * scale, v[j + 1] * scale, v[j + 2] * scale, v[j + 3] * scale);
}

// Action index of a trajectory element, clamped to [0, A-1]: the reference's F.one_hot / gather raise on an out-of-range
// action (atari_model.py:104, vtrace.py:35-40); a kernel cannot raise, so it must at least never index out of bounds
// (the host side offers the raising check: B200ImpalaLearner(validate_inputs=True)).
SRL_DEVINL int ld_action(const int64_t* p, int A) {
  const long long a = __ldg(reinterpret_cast<const long long*>(p));
  return a < 0 ? 0 : (a >= A ? A - 1 : (int)a);
}

SRL_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace srl
