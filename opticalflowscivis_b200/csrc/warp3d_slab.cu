// a2 on cubic volumes: 3-D warp with the SOURCE staged in shared memory by TMA (the north-star design for the gather
// kernels).  The generic kernel in warp.cu gathers straight from global memory: 8 LDG per voxel, each with a 64-bit address
// (a quarter of its instructions) and 10+ sectors per request on realistic flows (L1 data pipe at 67 %).  Here
//
//   * the reference warp rotates axes: output (d,h,w) samples source (z,y,x) ~ (w,d,h) + flow.  A CTA owns a
//     32(h) x 32(w) output column and WALKS ALONG d; what it needs of the source for plane d is the y-slab
//     src[z in w0-5 .. w0+38][y = d-5 .. d+6][x in h0-8 .. h0+39] — as d advances, one new 44 x 48 slab per plane enters a
//     16-slab ring (slot = y & 15).  Slabs arrive by ONE cp.async.bulk.tensor each (4-D box {48,1,44,1}, out-of-volume
//     elements zero-filled = the weight-0 neighbour of a border-clipped sample), two planes ahead of their use;
//   * the flow tile of a plane (3 x 32 x 32 floats) arrives by TMA too, SWIZZLE_128B so that lane = h reads its four
//     consecutive w values with one conflict-free LDS.128 per channel;
//   * a tap is an LDS with an immediate offset from one of two 32-bit bases (slab y0, slab y0+1): 8 LDS + ~10 address
//     instructions per voxel instead of 8 LDG + ~40;
//   * a voxel whose cell leaves the staged window (|flow| beyond ~5 voxels) takes the global gather of warp.cu, per lane —
//     correctness never depends on the window;
//   * results cross a swizzled shared tile and leave as fully coalesced 16 B stores along w.
//
// Coordinates, weights and the summation order are the shared helpers of warp_device.cuh: results are bit-identical to the
// generic kernel (tests/test_gpu_parity.py::test_warp3d_slab_equals_generic) and to the C oracle.
#include <cstdlib>

#include "tc_common.cuh"
#include "warp_device.cuh"

namespace ofsv {

constexpr int SL_TH = 32;                      // output tile height (lanes)
constexpr int SL_NS = 16;                      // ring slots
constexpr int SL_XLO = 8, SL_NX = 48;          // x window [h0-8, h0+40)
#ifndef OFSV_SLAB_G
#define OFSV_SLAB_G 2
#endif
constexpr int SL_G = OFSV_SLAB_G;               // voxels of a thread whose coordinate math / taps / sums are interleaved
constexpr int SL_ZLO = 5;                      // z window starts at w0-5
// Tile width TW (voxels along w): 32 -> 512 threads, 197 KB, one CTA per SM; 16 -> 256 threads, 111 KB, TWO CTAs per SM whose
// barrier / TMA waits overlap each other's arithmetic (the source window per output voxel grows from 2.06x to 2.44x).
template <int TW, int VPT = 4>
struct SlabCfg {
  // TMA pipeline depth PF (plane pairs in flight ahead of the one being computed) and y margin MY: a pair at planes (d, d+1)
  // reads slabs d-MY .. d+MY+2, the loads of pair +PF overwrite slots of slabs <= d+2PF+MY+2-16, so 2PF + 2MY < 14.  With one
  // pair in flight (41 KB per SM) the kernel was latency-bound at ~4 TB/s of L2->SM traffic; two pairs need MY = 4.
  static constexpr int PF = TW == 32 ? 2 : 1;
  static constexpr int MY = TW == 32 ? 4 : 5;
  static constexpr int NST = PF + 1;                         // flow stages / full barriers
  static_assert(2 * PF + 2 * MY < 14, "ring slots would be overwritten while live");
  static constexpr int NZ = TW == 32 ? 44 : 26;              // z window [w0-5, w0+TW+NZ-TW-5): margins 5 below, 6 / 4 above (+1 tap)
  static constexpr int SLAB = NZ * SL_NX * 4;                // bytes; multiple of 128
  static constexpr int FTILE = SL_TH * TW * 4;               // bytes per (plane, channel)
  static constexpr int FSTAGE = 2 * 3 * FTILE;               // two planes x three channels
  static constexpr int OUT = 2 * FTILE;                      // two planes of results
  static constexpr int QG = TW / VPT;                        // column groups (warps) per plane, VPT voxels per thread
  static constexpr int CPR = TW / 4;                         // 16 B chunks per tile row
  static constexpr int THREADS = 2 * QG * 32;
  static constexpr int ROWB = TW * 4;                        // bytes per tile row: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
  static constexpr int SMEM = SL_NS * SLAB + NST * FSTAGE + 2 * OUT + 64 + 1024;   // + barriers + alignment slack
  static_assert(SLAB % 128 == 0, "slab size");
  // 16 B chunk c of tile row `row` lives at chunk c ^ swz(row): TMA's 128 B / 64 B swizzle patterns
  __device__ static __forceinline__ uint32_t swz(int row) { return TW == 32 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3); }
};

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32x2(uint32_t addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts_f32x4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct SlabParams {
  int N, C, S;        // cubic: D = H = W = S
  int ref_mode;
  int nchunk;         // d chunks per tile column; chunk k covers planes [2*floor(k*(S/2)/nchunk), 2*floor((k+1)*(S/2)/nchunk))
  uint32_t ntasks;
  int flow_nc, flow_c0;   // the flow tensor has flow_nc channels per sample, this warp uses channels flow_c0 .. flow_c0+2
  int dbg_skip;       // probe builds only (-DOFSV_SLAB_PROBE + OFSV_SLAB_DBG_SKIP=1): no arithmetic, out = flow channel 0 — the
                      // kernel's pure data-movement time (212-222 us for 4 x 256^3); always 0 in the shipped library
  float hs[6];
};

template <int TW, int VPT, bool FMA>
__global__ void __launch_bounds__(SlabCfg<TW, VPT>::THREADS, (TW == 32 || VPT == 2) ? 1 : 2)
    warp3d_slab_kernel(const __grid_constant__ CUtensorMap tm_src, const __grid_constant__ CUtensorMap tm_flow,
                       const float* __restrict__ src, const float* __restrict__ lin_h, const float* __restrict__ lin_d,
                       const float* __restrict__ lin_w, float* __restrict__ out, const SlabParams P) {
  using K = SlabCfg<TW, VPT>;
  constexpr int SL_NZ = K::NZ, SL_SLAB = K::SLAB, SL_FTILE = K::FTILE, SL_FSTAGE = K::FSTAGE, SL_OUT = K::OUT;
  constexpr int SL_MY = K::MY, PF = K::PF, NST = K::NST;
  extern __shared__ uint8_t sl_raw[];
  const uint32_t s_base = (smem_u32(sl_raw) + 1023u) & ~1023u;            // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t s_flow = s_base;                                         // [2 stages][2 planes][3 ch][32 h][32 w] swizzled
  const uint32_t s_out = s_flow + NST * SL_FSTAGE;                          // [2 buffers][2 planes][32 h][32 w] swizzled
  const uint32_t s_slab = s_out + 2 * SL_OUT;                             // [16 slots][44 z][48 x]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sl_raw + (s_slab + SL_NS * SL_SLAB - smem_u32(sl_raw)));   // full[2]

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int pl = wid / K::QG, q = wid % K::QG;                                   // plane of the pair, 4-voxel column group
  const int S = P.S;
  const int64_t V = (int64_t)S * S * S;
  const int nth = S / 32, ntw = S / TW, nchunk = P.nchunk;
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t g = 0;                                                         // pair-iteration counter across tasks: stage = g & 1
  for (uint32_t task = blockIdx.x; task < P.ntasks; task += gridDim.x) {
    // task -> (volume nc, d chunk, h tile, w tile); w tile fastest: concurrent CTAs read neighbouring flow rows / source slabs
    uint32_t r = task;
    const int tw = r % ntw; r /= ntw;
    const int th = r % nth; r /= nth;
    const int ck = r % nchunk;
    const int nc = r / nchunk;
    const int n = nc / P.C;
    const int h0 = th * 32, w0 = tw * TW;
    const int d0 = 2 * ((ck * (S / 2)) / nchunk), npair = ((ck + 1) * (S / 2)) / nchunk - d0 / 2;
    const int xorg = h0 - SL_XLO, zorg = w0 - SL_ZLO;
    const float* sp = src + (int64_t)nc * V;
    float* op = out + (int64_t)nc * V;

    auto load_pair = [&](int d, uint32_t gi, int y_first, int y_last) {   // one thread: flow of planes d, d+1 + slabs y_first..y_last
      uint64_t* bar = &bars[gi % NST];
      mbar_expect_tx(bar, (uint32_t)(SL_FSTAGE + (y_last - y_first + 1) * SL_SLAB));
      const uint32_t fs = s_flow + (gi % NST) * SL_FSTAGE;
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int c = 0; c < 3; ++c) tma_load_4d(&tm_flow, bar, fs + (p * 3 + c) * SL_FTILE, w0, h0, d + p, n * P.flow_nc + P.flow_c0 + c);
      for (int y = y_first; y <= y_last; ++y) tma_load_4d(&tm_src, bar, s_slab + (uint32_t)(y & (SL_NS - 1)) * SL_SLAB, xorg, y, zorg, nc);
    };

    __syncthreads();                       // every thread is done with the previous task's slabs / flow stages
    if (tid == 0) {
      load_pair(d0, g, d0 - SL_MY, d0 + SL_MY + 2);
      for (int a = 1; a < PF && a < npair; ++a) load_pair(d0 + 2 * a, g + a, d0 + 2 * a + SL_MY + 1, d0 + 2 * a + SL_MY + 2);
    }

    const int h = h0 + lane;
    const float lh = __ldg(lin_h + h);
    float lw[VPT];
#pragma unroll
    for (int k = 0; k < VPT; ++k) lw[k] = __ldg(lin_w + w0 + q * VPT + k);
    for (int it = 0; it < npair; ++it, ++g) {
      const int dc = d0 + 2 * it;          // planes dc, dc + 1
      const int d = dc + pl;
      const float ld = __ldg(lin_d + d);   // issued before the barrier / stores / TMA wait so that its latency hides there
      __syncthreads();                     // pair it-1 fully computed: its results are in s_out, its oldest two slabs are dead
      if (tid == 0 && it + PF < npair) load_pair(dc + 2 * PF, g + PF, dc + 2 * PF + SL_MY + 1, dc + 2 * PF + SL_MY + 2);
      if (it > 0) {                        // coalesced stores of the previous pair
        const uint32_t ob = s_out + ((g - 1) & 1) * SL_OUT;
        const int p = tid / (K::THREADS / 2), tp = tid % (K::THREADS / 2), row = tp / K::CPR, c = tp % K::CPR;
        if (tp < 32 * K::CPR) {
          const float4 v = lds_f32x4(ob + p * SL_FTILE + row * K::ROWB + ((c ^ K::swz(row)) << 4));
          stg_stream4(op + ((int64_t)(dc - 2 + p) * S + (h0 + row)) * S + w0 + c * 4, v);
        }
      }
      mbar_wait(&bars[g % NST], (g / NST) & 1);

      // this thread's VPT consecutive w values of the three flow channels: 16 B chunk (q*VPT/4) ^ swz(row), VPT*4 B inside it
      const uint32_t toff = lane * K::ROWB + ((((q * VPT) >> 2) ^ K::swz(lane)) << 4) + ((q * VPT) & 3) * 4;
      const uint32_t fa = s_flow + (g % NST) * SL_FSTAGE + pl * 3 * SL_FTILE + toff;
      float f0[VPT], f1[VPT], f2[VPT];
      if (VPT == 4) {
        const float4 F0 = lds_f32x4(fa), F1 = lds_f32x4(fa + SL_FTILE), F2 = lds_f32x4(fa + 2 * SL_FTILE);
        f0[0] = F0.x; f0[1] = F0.y; f0[VPT - 2] = F0.z; f0[VPT - 1] = F0.w;
        f1[0] = F1.x; f1[1] = F1.y; f1[VPT - 2] = F1.z; f1[VPT - 1] = F1.w;
        f2[0] = F2.x; f2[1] = F2.y; f2[VPT - 2] = F2.z; f2[VPT - 1] = F2.w;
      } else {
        const float2 F0 = lds_f32x2(fa), F1 = lds_f32x2(fa + SL_FTILE), F2 = lds_f32x2(fa + 2 * SL_FTILE);
        f0[0] = F0.x; f0[1] = F0.y; f1[0] = F1.x; f1[1] = F1.y; f2[0] = F2.x; f2[1] = F2.y;
      }
      float res[VPT];
#pragma unroll
      for (int k = 0; k < VPT; ++k) res[k] = f0[k];
#ifdef OFSV_SLAB_PROBE
      if (!P.dbg_skip)
#endif
#pragma unroll
      for (int half = 0; half < VPT / SL_G; ++half) {
        TrilinCell cell[SL_G];
        Taps8 tp[SL_G];
#pragma unroll
        for (int j = 0; j < SL_G; ++j) {
          const int k = half * SL_G + j;
          cell[j] = trilin_cell<true>(f0[k], f1[k], f2[k], lh, ld, lw[k], S, S, S, P.hs, P.ref_mode);
        }
#pragma unroll
        for (int j = 0; j < SL_G; ++j) {
          const TrilinCell& c = cell[j];
          const uint32_t xr = (uint32_t)(c.x0 - xorg), zr = (uint32_t)(c.z0 - zorg), yr = (uint32_t)(c.y0 - (dc - SL_MY));
          if (xr <= (uint32_t)(SL_NX - 2) && zr <= (uint32_t)(SL_NZ - 2) && yr <= (uint32_t)(2 * SL_MY + 1)) {
            const uint32_t in = (zr * SL_NX + xr) * 4;
            const uint32_t a0 = s_slab + (uint32_t)(c.y0 & (SL_NS - 1)) * SL_SLAB + in;
            const uint32_t a1 = s_slab + (uint32_t)((c.y0 + 1) & (SL_NS - 1)) * SL_SLAB + in;
            tp[j].v[0] = lds_f32(a0); tp[j].v[1] = lds_f32(a0 + 4);
            tp[j].v[2] = lds_f32(a1); tp[j].v[3] = lds_f32(a1 + 4);
            tp[j].v[4] = lds_f32(a0 + SL_NX * 4); tp[j].v[5] = lds_f32(a0 + SL_NX * 4 + 4);
            tp[j].v[6] = lds_f32(a1 + SL_NX * 4); tp[j].v[7] = lds_f32(a1 + SL_NX * 4 + 4);
          } else {
            tp[j] = trilin_gather(sp, trilin_from_cell(c, S, S, S));      // cell outside the staged window
          }
        }
#pragma unroll
        for (int j = 0; j < SL_G; ++j) {
          Trilin t;
          t.ex = cell[j].ex; t.wx = cell[j].wx; t.ey = cell[j].ey; t.wy = cell[j].wy; t.ez = cell[j].ez; t.wz = cell[j].wz;
          t.base = t.dx = t.dy = t.dz = 0;
          res[half * SL_G + j] = trilin_reduce<FMA>(tp[j], t);
        }
      }
      if (VPT == 4) sts_f32x4(s_out + (g & 1) * SL_OUT + pl * SL_FTILE + toff, make_float4(res[0], res[1], res[VPT - 2], res[VPT - 1]));
      else sts_f32x2(s_out + (g & 1) * SL_OUT + pl * SL_FTILE + toff, make_float2(res[0], res[1]));
    }
    __syncthreads();
    {                                      // last pair of the task
      const uint32_t ob = s_out + ((g - 1) & 1) * SL_OUT;
      const int p = tid / (K::THREADS / 2), tp = tid % (K::THREADS / 2), row = tp / K::CPR, c = tp % K::CPR;
      if (tp < 32 * K::CPR) {
        const float4 v = lds_f32x4(ob + p * SL_FTILE + row * K::ROWB + ((c ^ K::swz(row)) << 4));
        stg_stream4(op + ((int64_t)(d0 + 2 * npair - 2 + p) * S + (h0 + row)) * S + w0 + c * 4, v);
      }
    }
  }
}

// returns 1 when the slab kernel was launched, 0 when the shape is not eligible (caller falls back), < 0 on error
template <int TW, int VPT>
static int warp3d_slab_launch(const float* src, const float* flow, const float* lin_h, const float* lin_d, const float* lin_w,
                              float* out, int N, int C, int S, int ref_mode, int flow_nc, int flow_c0, cudaStream_t st) {
  using K = SlabCfg<TW, VPT>;
  PFN_encodeTiled encode = get_tensor_map_encoder();
  if (!encode) { set_error("ofsv_warp3d_f32: cuTensorMapEncodeTiled unavailable"); return OFSV_ECUDA; }
  CUtensorMap tm_src, tm_flow;
  {
    const cuuint64_t gdim[4] = {(cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)N * C};
    const cuuint64_t gstr[3] = {(cuuint64_t)S * 4, (cuuint64_t)S * S * 4, (cuuint64_t)S * S * S * 4};
    const cuuint32_t box[4] = {SL_NX, 1, (cuuint32_t)K::NZ, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm_src, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(src), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("ofsv_warp3d_f32: cuTensorMapEncodeTiled(src) failed (%d)", (int)r); return OFSV_ECUDA; }
  }
  {
    const cuuint64_t gdim[4] = {(cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)N * flow_nc};
    const cuuint64_t gstr[3] = {(cuuint64_t)S * 4, (cuuint64_t)S * S * 4, (cuuint64_t)S * S * S * 4};
    const cuuint32_t box[4] = {(cuuint32_t)TW, SL_TH, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm_flow, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(flow), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, TW == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("ofsv_warp3d_f32: cuTensorMapEncodeTiled(flow) failed (%d)", (int)r); return OFSV_ECUDA; }
  }
  SlabParams P;
  P.N = N; P.C = C; P.S = S; P.ref_mode = ref_mode; P.flow_nc = flow_nc; P.flow_c0 = flow_c0;
  P.dbg_skip = 0;
#ifdef OFSV_SLAB_PROBE
  { const char* e = getenv("OFSV_SLAB_DBG_SKIP"); P.dbg_skip = e ? atoi(e) : 0; }   // probe builds only, never in the shipped library
#endif
  const Warp3dParams wp = make_warp3d_params(N, C, S, S, S, ref_mode);
  for (int i = 0; i < 6; ++i) P.hs[i] = wp.hs[i];
  const int64_t tiles = (int64_t)N * C * (S / 32) * (S / TW);
  const int slots = device_num_sms() * ((TW == 32 || VPT == 2) ? 1 : 2);      // resident CTAs of the persistent grid
  // d chunks per tile column: every task pays a prologue (13 slabs before its first plane pair, ~6 plane times) and the last
  // round of the persistent grid may be partly empty — minimise rounds x (planes per task + prologue)
  int nchunk = 1;
  double best = 1e30;
  for (int k = 1; k <= S / 4; ++k) {
    const int64_t rounds = cdiv(tiles * k, slots);
    const double cost = (double)rounds * (2.0 * (double)cdiv(S / 2, k) + 6.0);
    if (cost < best) { best = cost; nchunk = k; }
  }
  P.nchunk = nchunk;
  const int64_t ntasks = tiles * nchunk;
  if (ntasks >= (1ll << 31)) return 0;
  P.ntasks = (uint32_t)ntasks;
  const int grid = (int)(ntasks < slots ? ntasks : slots);
  static std::atomic<uint64_t> attr_a{0}, attr_b{0};
  if (int e = ensure_dyn_smem(attr_a, warp3d_slab_kernel<TW, VPT, true>, K::SMEM, "ofsv_warp3d_f32")) return e;
  if (int e = ensure_dyn_smem(attr_b, warp3d_slab_kernel<TW, VPT, false>, K::SMEM, "ofsv_warp3d_f32")) return e;
  if (ref_mode == OFSV_REF_CUDA)
    warp3d_slab_kernel<TW, VPT, true><<<grid, K::THREADS, K::SMEM, st>>>(tm_src, tm_flow, src, lin_h, lin_d, lin_w, out, P);
  else
    warp3d_slab_kernel<TW, VPT, false><<<grid, K::THREADS, K::SMEM, st>>>(tm_src, tm_flow, src, lin_h, lin_d, lin_w, out, P);
  const int rc = check_launch("warp3d_slab_kernel");
  return rc == OFSV_OK ? 1 : rc;
}

#ifndef OFSV_SLAB_TW
#define OFSV_SLAB_TW 16
#endif
std::atomic<int> g_warp_slab{1};
int warp3d_slab_try(const float* src, const float* flow, const float* lin_h, const float* lin_d, const float* lin_w, float* out,
                    int N, int C, int D, int H, int W, int ref_mode, cudaStream_t st, int flow_nc, int flow_c0) {
  if (!(D == H && H == W && W % 32 == 0 && W >= 32 && W <= 1024)) return 0;
  if (!aligned16(src) || !aligned16(flow) || !aligned16(out) || !aligned16(lin_w)) return 0;
  if ((int64_t)N * C > (1 << 20) || (int64_t)N * flow_nc > (1 << 20)) return 0;
  if (!g_warp_slab.load(std::memory_order_relaxed)) return 0;      // ofsv_set_tuning("warp_slab", 0): gather kernel only
  if (OFSV_SLAB_TW == 32) return warp3d_slab_launch<32, 4>(src, flow, lin_h, lin_d, lin_w, out, N, C, W, ref_mode, flow_nc, flow_c0, st);
  return warp3d_slab_launch<16, 4>(src, flow, lin_h, lin_d, lin_w, out, N, C, W, ref_mode, flow_nc, flow_c0, st);
}

}  // namespace ofsv
