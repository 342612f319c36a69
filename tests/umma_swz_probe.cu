// Hardware probe (design tool, run on the GPU box): cycles per tcgen05.mma M128 x N x K16 (bf16, SS mode) for the two operand
// layouts the stacked conv kernel can use — K-major SWIZZLE_128B (128 B rows, 64 channels) and SWIZZLE_64B (64 B rows, 32
// channels) — with the halo A pitch (8-row groups 10 rows apart, tap-shifted start) and stacked N = 64 .. 256.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/umma_swz_probe tests/umma_swz_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000ll) asm volatile("trap;");
  }
}
// rowb = 128 (layout 2) or 64 (layout 4); KSTEPS = rowb / 32 K16 steps per row
template <int ROWB>
__global__ void __launch_bounds__(128, 1) swz_rate_kernel(int N, int halo, int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                  // 64 KB
  uint8_t* sB = smem + 65536;          // 128 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 131072);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  uint32_t s = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
  for (int i = threadIdx.x; i < (65536 + 131072) / 4; i += 128) {
    s = s * 1664525u + 1013904223u;
    const uint32_t lo = 0x3F80u | ((s >> 9) & 0x807Fu), hi = 0x3F80u | ((s >> 17) & 0x807Fu);
    reinterpret_cast<uint32_t*>(smem)[i] = lo | (hi << 16);
  }
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (warp == 0) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(leader));
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint64_t layout = ROWB == 128 ? 2 : 4;
    constexpr int KS = ROWB / 32;
    const uint32_t sbo_a = halo ? 10 * ROWB : 8 * ROWB, sbo_b = 8 * ROWB;
    uint64_t da[4][KS], db[4][KS];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const uint32_t a = smem_u32(sA) + (halo ? (uint32_t)((t + 1) * 11 * ROWB) : (uint32_t)(t * 128 * ROWB / 4));   // tap-shifted starts
        const uint32_t b = smem_u32(sB) + (uint32_t)(t * 256 * ROWB / 4);
        da[t][k] = ((uint64_t)((a & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_a >> 4) << 32) | (1ull << 46) | (layout << 61)) + 2 * k;
        db[t][k] = ((uint64_t)((b & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_b >> 4) << 32) | (1ull << 46) | (layout << 61)) + 2 * k;
      }
    if (leader) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da[0][0]), "l"(db[0][0]), "r"(idesc) : "memory");
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < reps; i += 4 * KS) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          if (!leader) continue;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da[t][k]), "l"(db[t][k]), "r"(idesc) : "memory");
        }
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    __syncwarp();
    mbar_wait(bar, 0);
    if (blockIdx.x == 0 && leader) cycles[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
int main() {
  const size_t smem = 65536 + 131072 + 1024 + 64;
  CHECK(cudaFuncSetAttribute(swz_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CHECK(cudaFuncSetAttribute(swz_rate_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long* dC;
  CHECK(cudaMalloc(&dC, 8));
  const int reps = 4096;
  for (int rowb : {128, 64})
    for (int halo : {0, 1})
      for (int N : {64, 128, 192, 256}) {
        for (int r : {64, reps}) {
          if (rowb == 128) swz_rate_kernel<128><<<148, 128, smem>>>(N, halo, r, dC); else swz_rate_kernel<64><<<148, 128, smem>>>(N, halo, r, dC);
        }
        CHECK(cudaDeviceSynchronize());
        long long hC;
        CHECK(cudaMemcpy(&hC, dC, 8, cudaMemcpyDeviceToHost));
        const double cyc = (double)hC / reps;
        printf("rows %3d B (%s) %s N=%3d : %6.1f cycles per MMA (%5.1f %% of 4096 MAC/clk)\n", rowb, rowb == 128 ? "SWIZZLE_128B" : "SWIZZLE_64B", halo ? "halo pitch" : "dense     ", N, cyc, 100.0 * 128.0 * N * 16 / cyc / 4096.0);
      }
  printf("exit 0\n");
  return 0;
}
