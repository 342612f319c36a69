// Shared host/device helpers of libofsv.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/ofsv.h"

namespace ofsv {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return OFSV_ECUDA;
  }
  return OFSV_OK;
}

#define OFSV_REQUIRE(cond, ...)   \
  do {                            \
    if (!(cond)) {                \
      ofsv::set_error(__VA_ARGS__); \
      return OFSV_EINVAL;         \
    }                             \
  } while (0)

// Shifted space-to-depth addressing (block_stage.cu, ifnet_glue.cu, conv_halo.cu): position (z,y,x) of a logical
// [Dn][Hn][Wn] grid lives in cell ((i+1)>>1) with sub-index ((i+1)&1) per axis, cells stored [N][Dn/2+1][Hn/2+1][Wn/2+1]
// [2^nd sub-cells] rows (2-D: the D axis has extent 1 and is not split).  Returns the ROW index; the sub-cells that
// correspond to i = -1 and i = n (the conv padding) are never written and must stay zero.
__host__ __device__ inline int64_t s2d_row(int nd, int n, int z, int y, int x, int Dn, int Hn, int Wn) {
  const int Hc = Hn / 2 + 1, Wc = Wn / 2 + 1;
  const int y1 = y + 1, x1 = x + 1;
  if (nd == 2) return (((int64_t)n * Hc + (y1 >> 1)) * Wc + (x1 >> 1)) * 4 + (((y1 & 1) << 1) | (x1 & 1));
  const int Dc = Dn / 2 + 1, z1 = z + 1;
  return ((((int64_t)n * Dc + (z1 >> 1)) * Hc + (y1 >> 1)) * Wc + (x1 >> 1)) * 8 + (((z1 & 1) << 2) | ((y1 & 1) << 1) | (x1 & 1));
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE property of a kernel: remember per kernel which device
// ordinals have it (one bit each; devices >= 64 re-set it on every launch) instead of a per-process flag, so that a second
// GPU used from the same process gets it too and concurrent host threads never race on a plain bool.
template <typename K>
inline int ensure_dyn_smem(std::atomic<uint64_t>& done, K kernel, int bytes, const char* who) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("%s: cudaGetDevice: %s", who, cudaGetErrorString(e)); return OFSV_ECUDA; }
  const uint64_t bit = dev < 64 ? (1ull << dev) : 0ull;
  if (bit && (done.load(std::memory_order_acquire) & bit)) return OFSV_OK;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute(%d B of dynamic shared memory): %s", who, bytes, cudaGetErrorString(e)); return OFSV_ECUDA; }
  done.fetch_or(bit, std::memory_order_release);
  return OFSV_OK;
}
int device_num_sms();   // api.cu: SM count of the CURRENT device (cached per device ordinal)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ---------------------------------------------------------------------------------
// Normalised-coordinate arithmetic of warplayer.py + ATen grid_sampler, every op individually rounded
// (__f*_rn intrinsics are never contracted into FMAs by nvcc).  SURVEY.md Appendix A.
// f / half_extent as the CPU reference computes it (true IEEE division).  For the constants c = (S-1)/2 the Markstein
// sequence q = RN(f*rc), r = f - q*c (exact, FMA), RN(q + r*rc) with rc = RN(1/c) IS the correctly rounded quotient for every
// float 2^-60 <= |f| <= 2^20 (oracle/check_div_const.c checks all of them for S = 2..1024); it costs 3 instructions and no
// branch where __fdiv_rn costs ~20 with a slow-path call on the critical path of every warp tap.  Outside that range the
// result can differ from IEEE division in the last bit (or be NaN instead of +-inf beyond 1e38), which cannot change the
// warp: quotients below 2^-60/c vanish in `lin + q`, and |f| > 2^20 px lands far outside the volume and is border-clipped.
__device__ __forceinline__ float norm_flow(float f, float half_extent, float rcp_half_extent, int ref_mode) {
  const float q = __fmul_rn(f, rcp_half_extent);
  if (ref_mode == OFSV_REF_CUDA) return q;
  const float r = __fmaf_rn(-q, half_extent, f);
  return __fmaf_rn(r, rcp_half_extent, q);
}
// grid_sampler_unnormalize(align_corners=True) + clip_coordinates (border)
__device__ __forceinline__ float unnorm_clip_ac(float g, float size_m1) {
  float p = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), size_m1);  // ((g+1)/2)*(S-1); /2 == *0.5 exactly
  p = fmaxf(p, 0.0f);                                                 // NaN -> 0, like ATen's clamp order
  return fminf(size_m1, p);
}

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace ofsv
