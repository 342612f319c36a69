// Hardware probe (bring-up / design tool, run on the GPU box): issue rate of tcgen05.mma M128 x N x K16 (bf16 -> fp32) as a
// function of N, operand source (SS: A and B from shared memory; TS: A from tensor memory; A copied into tensor memory by
// tcgen05.cp in front of every MMA) and, in the halo layout the conv engine uses, of the tap shift of the A tile, its 8-row
// group pitch and the number of issuing warps.  Result on B200 (profiles/r01o_umma_rate_probe.txt): an SS-mode MMA costs
// max(N/2, (4 KB + N*32 B) / 128 B/clk) cycles — bound by the shared-memory read port below N = 128 — whatever the shift,
// pitch or number of issuers.
// CAVEAT: the per-MMA numbers of the generic kernel's TS / mixed modes include one or two R2UR (vector -> uniform register)
// moves per MMA on the issuing thread, each worth 10-15 cycles; only the SS rows and the halo kernel are straight-line
// UTCHMMA sequences.  The first version of this probe had such moves in its SS path too and over-stated the SS cost by 26
// cycles.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate_probe tests/umma_rate_probe.cu ; run: ./umma_rate_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x)                                                                     \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000ll) asm volatile("trap;");
  }
}
// K-major SWIZZLE_128B descriptor, rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t kdesc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// mode 0: SS, mode 1: TS (A in TMEM columns 256..), mode 2: every MMA is preceded by a tcgen05.cp (UTCCP) of its 128 x 32 B
// A slice from shared memory into a 4-slot TMEM ring and reads A from there (does staging A through TMEM hide the fetch?),
// mode 3: like 2 but one 128x256b copy feeds two MMAs
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int mode, int dense, int reps, int nacc, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                  // 4 tiles of 128 rows x 128 B (K = 64), walked round robin
  uint8_t* sB = smem + 4 * 16384;      // 4 tiles of 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 4 * 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  // operand bytes: zeros or pseudo-random bf16 in [-2, 2)
  uint32_t s = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += 128) {
    s = s * 1664525u + 1013904223u;
    uint32_t v = 0;
    if (dense) {
      const uint32_t lo = 0x3F80u | ((s >> 9) & 0x807Fu) | (((s >> 3) & 1u) << 7), hi = 0x3F80u | ((s >> 17) & 0x807Fu);
      v = lo | (hi << 16);
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (mode >= 1) {
    // A operand (128 lanes x 4 k-steps x 8 columns) into TMEM columns 256..287: each warp writes its 32 lanes
    uint32_t r[8];
    for (int k = 0; k < 4; ++k) {
      for (int e = 0; e < 8; ++e) { s = s * 1664525u + 1013904223u; r[e] = dense ? ((0x3F80u | ((s >> 9) & 0x807Fu)) | ((0x3F80u | ((s >> 17) & 0x807Fu)) << 16)) : 0u; }
      const uint32_t taddr = tmem + 256 + k * 8 + ((uint32_t)(warp * 32) << 16);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                   "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) {
    // the whole warp walks the (warp-uniform) loop and one elected lane issues: operands then live in uniform registers;
    // a divergent single-thread loop makes the compiler wrap every UTCHMMA in an R2UR.BROADCAST waterfall (~110 cycles each)
    uint32_t leader;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(leader));
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // all operand descriptors live in registers and the loop body is 16 straight-line MMAs: the single issuing thread must
    // not be the bottleneck of what is measured
    uint64_t da[16], db[16];
    uint32_t ta[16], td[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int t = j & 3, k = j >> 2;
      da[j] = kdesc(smem_u32(sA + t * 16384)) + 2 * k;
      db[j] = kdesc(smem_u32(sB + t * 32768)) + 2 * k;
      ta[j] = tmem + 256 + k * 8;
      td[j] = tmem + (j % nacc) * N;          // nacc independent accumulators: consecutive MMAs do not depend on each other
    }
    // first MMA overwrites the accumulator
    if (leader) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                 "l"(da[0]), "l"(db[0]), "r"(idesc) : "memory");
    const long long t0 = clock64();
    for (int i = 0; i < reps; i += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (!leader) continue;
        if (mode == 0) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(td[j]),
                       "l"(da[j]), "l"(db[j]), "r"(idesc) : "memory");
        } else {
          if (mode == 2 || (mode == 3 && (j & 1) == 0))
            asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(ta[j]), "l"(da[j]) : "memory");
          asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(td[j]),
                       "r"(ta[j]), "l"(db[j]), "r"(idesc) : "memory");
        }
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

// Halo-layout variant (what conv_halo.cu issues): N = 64, K = 64 per tap = 4 dependent MMAs on one accumulator, the A tile
// starts `ashift` bytes into the buffer (a tap shift of ashift/128 halo rows; 0 = 1024 B-aligned 8-row groups) and its 8-row
// groups are `sbo` bytes apart (1024 = dense, 1280 = the 10-row halo pitch); nw issuing warps with one accumulator each.
__global__ void __launch_bounds__(128, 1) halo_rate_kernel(int N, int ashift, int sbo, int nw, int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                  // 64 KB
  uint8_t* sB = smem + 4 * 16384;      // 4 tiles of 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 4 * 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  const int warp = threadIdx.x >> 5;
  uint32_t s = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += 128) {
    s = s * 1664525u + 1013904223u;
    const uint32_t lo = 0x3F80u | ((s >> 9) & 0x807Fu), hi = 0x3F80u | ((s >> 17) & 0x807Fu);
    reinterpret_cast<uint32_t*>(smem)[i] = lo | (hi << 16);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (warp < nw) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(leader));
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint64_t da[4][4], db[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // tap t: A shifted by a further t * 128 B (dx walk) from the base shift; weights tile t
        const uint32_t a = smem_u32(sA) + (uint32_t)ashift * (uint32_t)(t + 1) + (uint32_t)warp * 0;
        da[t][k] = ((uint64_t)((a & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61)) + 2 * k;
        db[t][k] = kdesc(smem_u32(sB + t * 32768)) + 2 * k;
      }
    const uint32_t td = tmem + warp * N;
    if (leader) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(td),
                 "l"(da[0][0]), "l"(db[0][0]), "r"(idesc) : "memory");
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < reps; i += 16) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (!leader) continue;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(td),
                       "l"(da[t][k]), "l"(db[t][k]), "r"(idesc) : "memory");
        }
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + warp)) : "memory");
    __syncwarp();
    mbar_wait(bar + warp, 0);
    if (blockIdx.x == 0 && leader) cycles[warp] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  const size_t smem = 4 * 16384 + 4 * 32768 + 1024 + 64;
  CHECK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long* dC;
  CHECK(cudaMalloc(&dC, 8));
  const int reps = 4096;
  const int Ns[] = {16, 32, 64, 128, 256};
  for (int grid : {1, 148})
    for (int mode : {0, 1, 2, 3})
      for (int dense : {0, 1})
        for (int N : Ns)
          for (int nacc : {1, 2, 4}) {
            if (nacc * N > 256 || (grid == 1 && dense == 0)) continue;
            rate_kernel<<<grid, 128, smem>>>(N, mode, dense, 64, nacc, dC);   // warm-up
            rate_kernel<<<grid, 128, smem>>>(N, mode, dense, reps, nacc, dC);
            CHECK(cudaDeviceSynchronize());
            long long hC;
            CHECK(cudaMemcpy(&hC, dC, 8, cudaMemcpyDeviceToHost));
            const double cyc = (double)hC / reps;
            printf("grid %3d %s %s N=%3d acc=%d : %6.1f cycles per MMA -> %5.0f MAC/cycle/SM (%4.1f %% of 4096)\n", grid, mode == 0 ? "SS" : mode == 1 ? "TS" : mode == 2 ? "CP+TS" : "CP+2TS",
                   dense ? "dense" : "zeros", N, nacc, cyc, 128.0 * N * 16 / cyc, 100.0 * 128.0 * N * 16 / cyc / 4096.0);
          }
  CHECK(cudaFuncSetAttribute(halo_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long* dC4;
  CHECK(cudaMalloc(&dC4, 32));
  for (int N : {64, 128})
    for (int nw : {1, 4})
      for (int sbo : {1024, 1280})
        for (int ashift : {0, 128, 384, 512, 1024, 1280}) {
          if (nw * N > 512) continue;
          halo_rate_kernel<<<148, 128, smem>>>(N, ashift, sbo, nw, 64, dC4);
          halo_rate_kernel<<<148, 128, smem>>>(N, ashift, sbo, nw, reps, dC4);
          CHECK(cudaDeviceSynchronize());
          long long h4[4];
          CHECK(cudaMemcpy(h4, dC4, 32, cudaMemcpyDeviceToHost));
          const double cyc = (double)h4[0] / (reps * nw);
          printf("halo N=%3d warps=%d sbo=%4d ashift=%4d : %6.1f cycles per MMA (%4.1f %% of peak)\n", N, nw, sbo, ashift, cyc,
                 100.0 * 128.0 * N * 16 / cyc / 4096.0);
        }
  printf("exit 0\n");
  return 0;
}
