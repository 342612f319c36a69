// tcgen05 / TMEM implicit-GEMM convolution engine for sm_100a (ofsv.h: ofsv_conv_tc).
//
// GEMM view of one layer in tap form:  D[m, co] = sum_{tap} sum_{ci} X[pos(m) * in_stride + off(tap), ci] * W[tap][ci][co]
//   M tile  = 128 virtual output positions = a (TW x TH x TD) box (8x4x4 in 3-D, 16x8x1 in 2-D)
//   N       = Cout_w (16..128, one UMMA N), K loop = taps x (Cin_s / KC) chunks of KC in {64,32,16} channels
//   A tile  = the input box shifted by the tap offset, fetched by ONE 5-D TMA box load per (tap, chunk) from the
//             channels-last activation tensor [N][D][H][W][C]: out-of-range coordinates are zero-filled by TMA, which
//             is exactly the convolution's zero padding; stride-2 layers use the tensor map's element strides.
//             The box lands in shared memory as 128 rows x KC channels = the K-major SWIZZLE_{128,64,32}B UMMA layout.
//   B tile  = W[tap][chunk] stored [Cout_w][KC] (K-major), one 2-D TMA load.
//   D       = fp32 accumulator in tensor memory (128 lanes x N columns), read back with tcgen05.ld by 4 epilogue warps
//             that fuse bias + PReLU + residual + dtype conversion and write channels-last output rows.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include <algorithm>
#include <mutex>

#include "tc_common.cuh"

namespace ofsv {

std::atomic<int> g_tc_pair{-1};      // -1 auto, 0 never, 1 whenever there are two tiles (ofsv_set_tuning("tc_pair", v))
std::atomic<int> g_tc_stages{0};     // 0 auto, 2..4 forces the pipeline depth (ofsv_set_tuning("tc_stages", v))
constexpr int TC_MAX_STAGES = 16;   // pipeline depth is chosen per layer: small K chunks need more loads in flight
constexpr int TC_M = 128;

struct TcParams {
  int N, Do, Ho, Wo, Dy, Hy, Wy, Cout_s, Cout_w;
  int in_stride, out_stride, ntaps, nkc;
  int tw, th, td, tiles_w, tiles_h, tiles_d;
  int has_prelu, has_residual, out_f32, stages;
  int8_t tap_off[OFSV_MAX_TAPS][4];
};

// MT = M tiles per CTA (1 or 2).  With 2, the CTA owns two consecutive tiles that SHARE every weight tile: the B operand — as many
// bytes per K step as one A tile when Cout_w = 128 — is fetched once for two MMAs, into two accumulators (2 x Cout_w TMEM columns).
// Layers with many tiles and large K (UPFlow's dense estimator: 1.3 MB of weights per tile) are bound by that L2 -> SM stream.
template <int KC, int MT>
__global__ void __launch_bounds__(192, 1)
    conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p,
                   const float* __restrict__ bias, const float* __restrict__ prelu, const void* __restrict__ residual,
                   void* __restrict__ y) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int A_BYTES = TC_M * KC * 2;
  const int b_bytes = (p.Cout_w * KC * 2 + 1023) & ~1023;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  const int S = p.stages;
  uint8_t* sB = smem + S * MT * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + S * b_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + TC_MAX_STAGES;
  uint64_t* accum_full = bars + 2 * TC_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_MAX_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ph = blockIdx.z;
  // tile -> (n, tz, ty, tx); the CTA's tiles are MT * blockIdx.x + m
  int tn[MT], ox0[MT], oy0[MT], oz0[MT];
  const int ntiles = p.tiles_w * p.tiles_h * p.tiles_d * p.N;
  const int nmt = min(MT, ntiles - (int)blockIdx.x * MT);      // live tiles of this CTA (the last CTA may own one)
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    int tile = min((int)blockIdx.x * MT + m, ntiles - 1);
    const int tx = tile % p.tiles_w; tile /= p.tiles_w;
    const int ty = tile % p.tiles_h; tile /= p.tiles_h;
    const int tz = tile % p.tiles_d;
    tn[m] = tile / p.tiles_d;
    ox0[m] = tx * p.tw; oy0[m] = ty * p.th; oz0[m] = tz * p.td;
  }
  const int kiters = p.ntaps * p.nkc;
  uint32_t ncol1 = 32;
  while ((int)ncol1 < p.Cout_w) ncol1 <<= 1;
  const uint32_t ncols = ncol1 * MT;                           // power of two: MT is 1 or 2

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (one lane) =================
    if (lane == 0) {
      const uint32_t tx_bytes = nmt * A_BYTES + p.Cout_w * KC * 2;
      for (int it = 0; it < kiters; ++it) {
        const int s = it % S;
        if (it >= S) mbar_wait(&empty[s], ((it / S) - 1) & 1);
        const int t = it / p.nkc, kc = it - t * p.nkc;
        const int8_t* off = p.tap_off[ph * p.ntaps + t];
        mbar_expect_tx(&full[s], tx_bytes);
#pragma unroll
        for (int m = 0; m < MT; ++m)
          if (m < nmt)
            tma_load_5d(&tmA, &full[s], sA + (s * MT + m) * A_BYTES, kc * KC, ox0[m] * p.in_stride + off[2], oy0[m] * p.in_stride + off[1],
                        oz0[m] * p.in_stride + off[0], tn[m]);
        tma_load_2d(&tmB, &full[s], sB + s * b_bytes, 0, ((ph * p.ntaps + t) * p.nkc + kc) * p.Cout_w);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: warp-uniform loop, one elected lane issues =================
    const uint32_t leader = elect_one_sync();
    // cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b_format BF16 (1) @7/@10, K-major A and B, N>>3 @17, M>>4 @24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Cout_w >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    const uint32_t d_hi = kmajor_desc_hi<KC>(8 * KC * 2);
    const uint32_t a_lo0 = kmajor_desc_lo(smem_u32(sA)), b_lo0 = kmajor_desc_lo(smem_u32(sB));
    for (int it = 0; it < kiters; ++it) {
      const int s = it % S;
      mbar_wait(&full[s], (it / S) & 1);
      tcgen05_fence_after();
      if (leader) {
        const uint32_t b_lo = b_lo0 + s * (b_bytes >> 4);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          if (m >= nmt) break;
          const uint32_t a_lo = a_lo0 + (s * MT + m) * (A_BYTES >> 4), acc = tmem_base + m * ncol1;
          umma_bf16_lohi(acc, a_lo, d_hi, b_lo, d_hi, idesc, it > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 1; k < KC / 16; ++k) umma_bf16_lohi(acc, a_lo + 2 * k, d_hi, b_lo + 2 * k, d_hi, idesc, 1u);
        }
        tcgen05_commit(&empty[s]);          // frees the smem slot when these MMAs have read it
      }
      __syncwarp();
    }
    if (leader) tcgen05_commit(accum_full);   // accumulator complete
    __syncwarp();
  } else {
    // ================= epilogue: TMEM -> registers -> global =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int rx = row % p.tw, ry = (row / p.tw) % p.th, rz = row / (p.tw * p.th);
    mbar_wait(accum_full, 0);
    tcgen05_fence_after();
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m >= nmt) break;
      const int ox = ox0[m] + rx, oy = oy0[m] + ry, oz = oz0[m] + rz;
      const bool valid = ox < p.Wo && oy < p.Ho && oz < p.Do;
      const int yz = oz * p.out_stride + ((ph >> 2) & 1), yy = oy * p.out_stride + ((ph >> 1) & 1), yx = ox * p.out_stride + (ph & 1);
      const int64_t yo = ((((int64_t)tn[m] * p.Dy + yz) * p.Hy + yy) * p.Wy + yx) * p.Cout_s;
      for (int c0 = 0; c0 < p.Cout_w; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * ncol1 + c0), v);
        if (!valid) continue;
        epilogue_store16(v, c0, yo, p.Cout_s, p.has_prelu, p.has_residual, p.out_f32, bias, prelu, residual, y);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
PFN_encodeTiled get_tensor_map_encoder() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

template <int KC, int MT>
static int launch_tc(const TcParams& P, const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias, const float* prelu,
                     const void* residual, void* y, dim3 grid, cudaStream_t st) {
  const int b_bytes = (P.Cout_w * KC * 2 + 1023) & ~1023;
  const size_t smem = 1024 + (size_t)P.stages * (MT * TC_M * KC * 2 + b_bytes) + (2 * TC_MAX_STAGES + 1) * 8 + 16;
  static std::atomic<uint64_t> attr_done{0};   // per template instance, one bit per device
  if (int e = ensure_dyn_smem(attr_done, conv_tc_kernel<KC, MT>, 200 * 1024, "ofsv_conv_tc")) return e;
  conv_tc_kernel<KC, MT><<<grid, 192, smem, st>>>(tmA, tmB, P, bias, prelu, residual, y);
  return check_launch("conv_tc_kernel");
}

}  // namespace ofsv

using namespace ofsv;

extern "C" int ofsv_conv_tc(const ofsv_conv_desc* d, const void* x, const void* w, const float* bias, const float* prelu,
                            const void* residual, void* y, void* stream) {
  if (int e = validate_conv_desc(d, "ofsv_conv_tc")) return e;
  if (d->out_shuffle) { set_error("ofsv_conv_tc: depth-to-space heads are only implemented by ofsv_conv_halo"); return OFSV_ENOSUP; }
  if (d->out_s2d) { set_error("ofsv_conv_tc: space-to-depth outputs are only implemented by ofsv_conv_halo"); return OFSV_ENOSUP; }
  if (d->N == 0) return OFSV_OK;
  OFSV_REQUIRE(x && w && bias && y, "ofsv_conv_tc: null pointer");
  OFSV_REQUIRE(!d->has_prelu || prelu, "ofsv_conv_tc: has_prelu without prelu slopes");
  OFSV_REQUIRE(!d->has_residual || residual, "ofsv_conv_tc: has_residual without residual");
  OFSV_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && (!residual || aligned16(residual)),
               "ofsv_conv_tc: pointers must be 16-byte aligned");
  if (d->in_dtype != OFSV_BF16) { set_error("ofsv_conv_tc: activations must be bf16"); return OFSV_ENOSUP; }
  if (d->Cout_w > 128) { set_error("ofsv_conv_tc: Cout_w=%d > 128 not supported", d->Cout_w); return OFSV_ENOSUP; }
  PFN_encodeTiled encode = get_tensor_map_encoder();
  if (!encode) { set_error("ofsv_conv_tc: cuTensorMapEncodeTiled unavailable (driver too old?)"); return OFSV_ECUDA; }

  const int KC = d->Cin_s % 64 == 0 ? 64 : (d->Cin_s % 32 == 0 ? 32 : 16);
  TcParams P;
  P.N = d->N; P.Do = d->Do; P.Ho = d->Ho; P.Wo = d->Wo; P.Dy = d->Dy; P.Hy = d->Hy; P.Wy = d->Wy;
  P.Cout_s = d->Cout_s; P.Cout_w = d->Cout_w; P.in_stride = d->in_stride; P.out_stride = d->out_stride;
  P.ntaps = d->ntaps; P.nkc = d->Cin_s / KC;
  if (d->nd == 3) { P.tw = 8; P.th = 4; P.td = 4; } else { P.tw = 16; P.th = 8; P.td = 1; }
  P.tiles_w = (int)cdiv(d->Wo, P.tw); P.tiles_h = (int)cdiv(d->Ho, P.th); P.tiles_d = (int)cdiv(d->Do, P.td);
  P.has_prelu = d->has_prelu; P.has_residual = d->has_residual; P.out_f32 = d->out_dtype == OFSV_F32;
  memcpy(P.tap_off, d->tap_off, sizeof(P.tap_off));
  const int64_t ntiles = (int64_t)P.tiles_w * P.tiles_h * P.tiles_d * d->N;
  OFSV_REQUIRE(ntiles < (1ll << 31), "ofsv_conv_tc: too many tiles");
  // two tiles per CTA when there are enough tiles to keep every SM busy with pairs and the K loop is long enough for the shared
  // weight stream to matter (ofsv_set_tuning("tc_pair", 0 | 1) forces it off / on where it fits)
  const int pair_mode = g_tc_pair.load(std::memory_order_relaxed);
  const bool pair = pair_mode != 0 && ntiles >= 2 && (pair_mode > 0 || (ntiles >= 4 * (int64_t)device_num_sms() && d->ntaps * P.nkc >= 8));
  {  // Pipeline depth.  The kernel is one tile (pair) per CTA, not persistent: while a CTA runs its prologue or its epilogue the tensor
     // pipe of its SM idles unless ANOTHER CTA is resident.  Small stages get that from occupancy (4 stages, up to ~10 CTAs per SM,
     // measured faster than deep rings); 32-48 KB stages (KC = 64 with 64-128 output channels) are cut to the depth at which two
     // CTAs fit the 227 KB of an SM.  ofsv_set_tuning("tc_stages", 2..4) forces a depth.
    const int kiters = d->ntaps * P.nkc;
    const int stage_bytes = (pair ? 2 : 1) * TC_M * KC * 2 + ((d->Cout_w * KC * 2 + 1023) & ~1023);
    int depth = g_tc_stages.load(std::memory_order_relaxed);
    if (depth <= 0) depth = std::max(2, std::min(4, (112 * 1024 - 2048) / stage_bytes));
    depth = std::max(1, std::min(depth, std::min(4, (200 * 1024 - 2048) / stage_bytes)));
    P.stages = kiters < depth ? kiters : depth;
  }

  const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmA, tmB;
  {
    const cuuint64_t gdim[5] = {(cuuint64_t)d->Cin_s, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->Di, (cuuint64_t)d->N};
    const cuuint64_t es = 2;
    const cuuint64_t gstr[4] = {d->Cin_s * es, (cuuint64_t)d->Wi * d->Cin_s * es, (cuuint64_t)d->Hi * d->Wi * d->Cin_s * es,
                                (cuuint64_t)d->Di * d->Hi * d->Wi * d->Cin_s * es};
    const cuuint32_t s = (cuuint32_t)d->in_stride, sd = d->nd == 3 ? s : 1;   // 2-D: the D axis has extent 1
    const cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)P.tw * s, (cuuint32_t)P.th * s, (cuuint32_t)P.td * sd, 1};
    const cuuint32_t estr[5] = {1, s, s, sd, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("ofsv_conv_tc: cuTensorMapEncodeTiled(A) failed with %d", (int)r); return OFSV_ECUDA; }
  }
  {
    const cuuint64_t rows = (cuuint64_t)d->nphase * d->ntaps * P.nkc * d->Cout_w;
    const cuuint64_t gdim[2] = {(cuuint64_t)KC, rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)KC * 2};
    const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)d->Cout_w};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("ofsv_conv_tc: cuTensorMapEncodeTiled(B) failed with %d", (int)r); return OFSV_ECUDA; }
  }
  const dim3 grid((unsigned)(pair ? cdiv(ntiles, 2) : ntiles), 1, (unsigned)d->nphase);
  cudaStream_t st = (cudaStream_t)stream;
  if (pair) {
    if (KC == 64) return launch_tc<64, 2>(P, tmA, tmB, bias, prelu, residual, y, grid, st);
    if (KC == 32) return launch_tc<32, 2>(P, tmA, tmB, bias, prelu, residual, y, grid, st);
    return launch_tc<16, 2>(P, tmA, tmB, bias, prelu, residual, y, grid, st);
  }
  if (KC == 64) return launch_tc<64, 1>(P, tmA, tmB, bias, prelu, residual, y, grid, st);
  if (KC == 32) return launch_tc<32, 1>(P, tmA, tmB, bias, prelu, residual, y, grid, st);
  return launch_tc<16, 1>(P, tmA, tmB, bias, prelu, residual, y, grid, st);
}
