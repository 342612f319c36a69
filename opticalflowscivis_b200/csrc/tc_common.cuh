// tcgen05 / TMA / mbarrier PTX wrappers and the fused conv epilogue shared by conv_tc.cu and conv_halo.cu (sm_100a).
#pragma once
#include <cuda.h>

#include "ofsv_common.cuh"

namespace ofsv {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline bug must surface as a trap (cudaErrorLaunchFailure), never as a hung GPU.  The fast path (barrier
// already complete) is a single try_wait: the MMA-issuing thread calls this once per operand tile.
__device__ __forceinline__ uint32_t mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  return done;
}
// Waiting warps back off with nanosleep: at N = 64 the tensor core's operand fetch needs ~95 % of the shared-memory
// cycles, so nine warps hammering mbarrier words in shared memory measurably slow the MMAs down.
// (mbarrier.try_wait is itself a hardware-suspended wait with a time limit, not a poll: sleep_ns = 0 simply re-arms it.  A
// __nanosleep between tries costs far more than its argument — the stacked conv kernel's weight ring measured a ~2.8 us
// producer/consumer round trip with 32-64 ns back-offs, i.e. 0.35 us per stage whatever the stage size.)
static __device__ __noinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity, uint32_t sleep_ns) {
  const long long t0 = clock64();
  while (!mbar_try(addr, parity)) {
    if (sleep_ns) __nanosleep(sleep_ns);
    if (clock64() - t0 > 4000000000ll) asm volatile("trap;");
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t sleep_ns = 32) {
  const uint32_t addr = smem_u32(bar);
  if (!mbar_try(addr, parity)) mbar_wait_slow(addr, parity, sleep_ns);
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the two shared-memory descriptors passed as (lo, hi) 32-bit halves: advancing along K or to another tap is
// then ONE 32-bit add on `lo` (start address >> 4) — the single issuing thread is instruction-latency bound otherwise.
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred;
}
// upper 32 bits of a K-major descriptor: SBO>>4 @[0,14), version 1 @14, layout type @29
template <int KC>
__device__ __forceinline__ uint32_t kmajor_desc_hi(uint32_t sbo_bytes) {
  constexpr uint32_t layout = KC == 64 ? 2u : (KC == 32 ? 4u : 6u);
  return (sbo_bytes >> 4) | (1u << 14) | (layout << 29);
}
// lower 32 bits: start address >> 4 @[0,14), LBO (unused for swizzled K-major, 1 like CUTLASS) @16
__device__ __forceinline__ uint32_t kmajor_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 @16 | SBO>>4 @32 | version 1 @46
// | layout type @61 (SWIZZLE_128B = 2, 64B = 4, 32B = 6).  Rows are KC*2 bytes; 8-row groups are SBO = 8*row bytes apart.
template <int KC>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  constexpr uint64_t layout = KC == 64 ? 2 : (KC == 32 ? 4 : 6);
  constexpr uint64_t sbo = (8 * KC * 2) >> 4;
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}


// Epilogue of one 16-column chunk of one accumulator row: + bias, PReLU, (+ residual), convert, 16 B vector stores.
// `yo` = element offset of channel 0 of this row's output position; nstore = 16, or 8 for the 8-channel head tensor.
__device__ __forceinline__ void epilogue_store16(float* v, int c0, int64_t yo, int Cout_s, bool has_prelu, bool has_residual,
                                                 bool out_f32, const float* __restrict__ bias, const float* __restrict__ prelu,
                                                 const void* __restrict__ residual, void* __restrict__ y) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float a = v[j] + __ldg(bias + c0 + j);
    if (has_prelu) a = a > 0.0f ? a : a * __ldg(prelu + c0 + j);
    v[j] = a;
  }
  const int nstore = min(16, Cout_s - c0);
  if (nstore <= 0) return;
  if (out_f32) {
    float* o = reinterpret_cast<float*>(y) + yo + c0;
    if (has_residual) {
      const float* r = reinterpret_cast<const float*>(residual) + yo + c0;
      for (int j = 0; j < nstore; ++j) v[j] += __ldg(r + j);
    }
    for (int j = 0; j < nstore; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(y) + yo + c0;
    if (has_residual) {
      const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(residual) + yo + c0;
      for (int j = 0; j < nstore; j += 8) {
        const uint4 rr = __ldg(reinterpret_cast<const uint4*>(r + j));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[j + 2 * e] += __low2float(h[e]); v[j + 2 * e + 1] += __high2float(h[e]); }
      }
    }
    for (int j = 0; j < nstore; j += 8) {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[j + 2 * e], v[j + 2 * e + 1]);
        w[e] = *reinterpret_cast<const uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(o + j) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_tensor_map_encoder();   // conv_tc.cu
int validate_conv_desc(const ofsv_conv_desc* d, const char* who);  // conv_simt.cu
// conv_halo_ring.cu: same contract as ofsv_conv_halo (called by it for the layers the plane ring is faster on)
int conv_halo_ring(const ofsv_conv_desc* d, const void* x, const void* w, const float* bias, const float* prelu,
                   const void* residual, void* y, void* stream);

}  // namespace ofsv
