// Hardware probe (bring-up tool, run on the GPU box): does a K-major SWIZZLE_128B UMMA descriptor address shared
// memory by ABSOLUTE address bits?  i.e. may the A tile start at any 128 B row of a TMA-written (1024 B-aligned) box,
// with an 8-row-group stride (SBO) that is not a multiple of 1024 B?
//
// If yes, an implicit-GEMM convolution can keep ONE halo tile of the input in shared memory and express every filter
// tap as a shifted descriptor (start += row shift, SBO = halo row pitch) instead of re-loading the A tile per tap.
//
//   X  [ROWS][64] bf16 small integers, TMA-loaded (SWIZZLE_128B) to smem;  B = 64x64 identity  =>  D[m][n] = A[m][n]
//   for each (shift, sbo_rows): A descriptor start = smem + shift*128, SBO = sbo_rows*128
//   expectation if address-based: D[m][n] == X[(m/8)*sbo_rows + shift + m%8][n]
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tests/umma_probe.cu ; run: ./umma_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ROWS 256
#define CHECK(x)                                                                  \
  do {                                                                            \
    cudaError_t e = (x);                                                          \
    if (e != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
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
    if (clock64() - t0 > 2000000000ll) asm volatile("trap;");
  }
}
__device__ __forceinline__ uint64_t kdesc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         ((uint64_t)(base_offset & 7) << 49) | (2ull << 61);
}

__global__ void __launch_bounds__(128, 1)
    probe_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB, float* out, int shift,
                 int sbo_rows, int use_base_offset, int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                 // ROWS x 128 B
  uint8_t* sB = smem + ROWS * 128;    // 64 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint64_t* bar2 = bar + 1;
  uint64_t* bar3 = bar + 2;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar2)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar3)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(ROWS * 128 + 64 * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(sX)), "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(sB)), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a0 = smem_u32(sX) + shift * 128, b0 = smem_u32(sB);
    const uint32_t bo = use_base_offset == 1 ? ((a0 >> 7) & 7) : 0;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
      for (int k = 0; k < 4; ++k) {
        // reps < 0 never; "vary" mode (use_base_offset == 2): a different A tile (tap-like row shift) on every repetition
        const uint32_t av = a0 + (use_base_offset == 2 ? ((rep & 15) * 5) * 128 : 0);
        const uint64_t da = kdesc(av + k * 32, sbo_rows * 128, bo == 2 ? 0 : bo), db = kdesc(b0 + k * 32, 1024, 0);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                     "l"(da), "l"(db), "r"(idesc), "r"((k > 0 || rep > 0) ? 1u : 0u) : "memory");
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar2)) : "memory");
    mbar_wait(bar2, 0);
    if (cycles) *cycles = clock64() - t0;
    if (reps > 1) {   // timing mode: re-arm nothing, the epilogue below waits on a second commit
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar3)) : "memory");
    }
    if (reps <= 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar3)) : "memory");
  }
  __syncwarp();
  mbar_wait(bar3, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[row * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}


// ---- issue-pattern timing: how many cycles does one M128 x N x K16 MMA cost under the halo kernel's issue patterns? ----
__device__ __forceinline__ void umma_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1)
    pattern_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB, int variant, int N, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sB = smem + ROWS * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint64_t* bar2 = bar + 1;
  uint64_t* bar3 = bar + 2;   // commit sink
  uint64_t* bar4 = bar + 3;   // never armed: parity-1 waits return immediately
  uint64_t* tbar = bar + 4;   // [4] concurrent-TMA barriers (variant 5/6)
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  volatile uint32_t* stop = slot + 1;
  uint8_t* scratch = smem + ROWS * 128 + 64 * 128 + 1024;   // 4 x 8 KB
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    *stop = 0;
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (warp == 0) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(leader));
    if (leader) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(ROWS * 128 + 64 * 128) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                       smem_u32(sX)), "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                       smem_u32(sB)), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    }
    __syncwarp();
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hi = (1280u >> 4) | (1u << 14) | (2u << 29), b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = ((smem_u32(sX) & 0x3FFFF) >> 4) | (1u << 16), b_lo0 = ((smem_u32(sB) & 0x3FFFF) >> 4) | (1u << 16);
    const int blocks = 64;                       // 64 blocks x 16 MMAs
    const long long t0 = clock64();
    for (int blk = 0; blk < blocks; ++blk) {
      if (variant >= 3) { mbar_wait(bar4, 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      if (leader) {
        const uint32_t a_lo = a_lo0 + (uint32_t)(blk & 7) * (variant >= 4 ? 11u : 8u) * 8u;      // tap-like shift: rows*8 (16 B units); v4+: 11-row steps (not atom aligned)
        // v6: the halo kernel's footprint — 4 slices in planes 23552 B apart (from the big scratch area), B tiles from an 8-slot ring
        const uint32_t b_lo = b_lo0 + (variant >= 6 ? 2560u + (uint32_t)(blk & 7) * 512u : 0u);
        for (int j = 0; j < 4; ++j) {
          const uint32_t dj = tmem + (variant >= 1 ? j * N : 0);
          const uint32_t aj = a_lo + (variant >= 6 ? 6656u + (uint32_t)j * 1472u : (variant >= 1 ? (uint32_t)j * 80 : 0));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_lohi(dj, aj + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, (blk | k) ? 1u : 0u);
        }
        if (variant >= 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar3)) : "memory");
      }
      __syncwarp();
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar2)) : "memory");
    __syncwarp();
    mbar_wait(bar2, 0);
    if (leader && blockIdx.x == 0) *cycles = clock64() - t0;
  }
  if (warp == 0 && (threadIdx.x & 31) == 0) *stop = 1;
  if (warp == 1 && threadIdx.x == 32 && variant >= 5) {
    // concurrent TMA traffic into shared memory: 8 KB tiles, up to 4 in flight, as fast as they complete
    uint32_t n = 0;
    while (!*stop) {
      const uint32_t sidx = n & 3;
      if (n >= 4) mbar_wait(&tbar[sidx], ((n >> 2) - 1) & 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tbar[sidx])), "r"(64 * 128) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                       smem_u32(scratch + sidx * 8192)), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&tbar[sidx])), "r"(0), "r"(0) : "memory");
      ++n;
    }
    for (uint32_t i = (n > 4 ? n - 4 : 0); i < n; ++i) mbar_wait(&tbar[i & 3], (i >> 2) & 1);
    if (blockIdx.x == 0) cycles[1] = n;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
  PFN_encodeTiled encode = (PFN_encodeTiled)fp;
  static __nv_bfloat16 hX[ROWS * 64], hB[64 * 64];
  static float X[ROWS * 64];
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < 64; ++c) {
      X[r * 64 + c] = (float)(((r * 7 + c * 3) % 251) - 125);
      hX[r * 64 + c] = __float2bfloat16(X[r * 64 + c]);
    }
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.0f : 0.0f);
  __nv_bfloat16 *dX, *dB;
  float* dO;
  CHECK(cudaMalloc(&dX, sizeof(hX))); CHECK(cudaMalloc(&dB, sizeof(hB))); CHECK(cudaMalloc(&dO, 128 * 64 * 4));
  CHECK(cudaMemcpy(dX, hX, sizeof(hX), cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
  CUtensorMap tmX, tmB;
  {
    cuuint64_t gd[2] = {64, ROWS}; cuuint64_t gs[1] = {128}; cuuint32_t box[2] = {64, ROWS}; cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dX, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode X failed %d\n", (int)r); return 2; }
    cuuint64_t gd2[2] = {64, 64}; cuuint32_t box2[2] = {64, 64};
    r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, gd2, gs, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); return 2; }
  }
  const size_t smem = 220 * 1024;
  CHECK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static float hO[128 * 64];
  int all_ok = 1;
  const int sbos[] = {8, 10, 9, 12, 16, 18};
  for (int ubo = 0; ubo < 2; ++ubo)
    for (int si = 0; si < 6; ++si)
      for (int shift = 0; shift < 11; ++shift) {
        const int sbo = sbos[si];
        if (15 * sbo + shift + 8 > ROWS) continue;
        probe_kernel<<<1, 128, smem>>>(tmX, tmB, dO, shift, sbo, ubo, 1, nullptr);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("shift %d sbo %d ubo %d: CUDA error %s\n", shift, sbo, ubo, cudaGetErrorString(e)); return 3; }
        CHECK(cudaMemcpy(hO, dO, sizeof(hO), cudaMemcpyDeviceToHost));
        int bad = 0, first_m = -1, first_n = -1;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            const float exp = X[((m / 8) * sbo + shift + m % 8) * 64 + n];
            if (hO[m * 64 + n] != exp) { if (!bad) { first_m = m; first_n = n; } ++bad; }
          }
        printf("base_offset_field=%d sbo_rows=%2d shift=%2d : %s (%d mismatches, first at m=%d n=%d)\n", ubo, sbo, shift,
               bad ? "MISMATCH" : "ok", bad, first_m, first_n);
        if (bad && !ubo) all_ok = 0;
      }
  printf("RESULT address-based swizzle with base_offset=0: %s\n", all_ok ? "CONFIRMED" : "NOT confirmed");
  // ---- timing: cycles per M128 x N64 x K16 MMA for aligned vs shifted / non-1024 B-stride A descriptors
  long long* dC; long long hC; long long hC2[2];
  CHECK(cudaMalloc(&dC, 16));
  CHECK(cudaMemset(dC, 0, 16));
  CHECK(cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const char* vn[] = {"v0 back-to-back, 1 accumulator", "v1 4 accumulators + per-slice A", "v2 v1 + commit per 16 MMAs", "v3 v2 + mbarrier try_wait per 16 MMAs", "v4 v3 with A starts not 1024B-aligned", "v5 v4 + concurrent 8 KB TMA loads into smem", "v6 v4 with the halo kernel's smem footprint (4 planes, B ring)"};
  for (int pass2 = 0; pass2 < 1; ++pass2) {
  if (pass2 == 1) {   // dense random operands (the integer / identity operands above are mostly zero bits)
    srand(1);
    for (int i = 0; i < ROWS * 64; ++i) hX[i] = __float2bfloat16((float)rand() / RAND_MAX - 0.5f);
    for (int i = 0; i < 64 * 64; ++i) hB[i] = __float2bfloat16((float)rand() / RAND_MAX - 0.5f);
    CHECK(cudaMemcpy(dX, hX, sizeof(hX), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
  }
  for (int grid : {1})
  for (int N : {64, 16})
    for (int v = 4; v < 7; v += 2) {
      printf("[%s data, grid %3d] ", pass2 ? "random" : "sparse", grid);
      pattern_kernel<<<grid, 128, smem>>>(tmX, tmB, v, N, dC);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pattern v%d N=%d: CUDA error %s\n", v, N, cudaGetErrorString(e)); return 3; }
      CHECK(cudaMemcpy(&hC, dC, 8, cudaMemcpyDeviceToHost));
      CHECK(cudaMemcpy(hC2, dC, 16, cudaMemcpyDeviceToHost));
      printf("PATTERN N=%3d %-42s : %.1f cycles per MMA  (%lld concurrent TMA tiles = %.0f B per MMA)\n", N, vn[v], (double)hC / 1024.0, v >= 5 ? hC2[1] : 0ll, v >= 5 ? hC2[1] * 8192.0 / 1024.0 : 0.0);
    }
  }
  const int cfg[][2] = {{0, 8}, {1, 8}, {4, 8}, {0, 10}, {1, 10}, {11, 10}, {0, 16}, {3, 12}};
  for (int rep_i = 0; rep_i < 2; ++rep_i)
    for (auto& c : cfg) {
      const int reps = 256;
      probe_kernel<<<1, 128, smem>>>(tmX, tmB, dO, c[0], c[1], rep_i == 1 ? 2 : 0, reps, dC);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("timing shift %d sbo %d: CUDA error %s\n", c[0], c[1], cudaGetErrorString(e)); return 3; }
      CHECK(cudaMemcpy(&hC, dC, 8, cudaMemcpyDeviceToHost));
      printf("TIMING %s shift=%2d sbo_rows=%2d : %.1f cycles per MMA (M128 N64 K16, %d MMAs back to back)\n", rep_i ? "varyA" : "sameA", c[0], c[1],
             (double)hC / (reps * 4), reps * 4);
    }
  return 0;
}
