// Backward of the conv()/deconv() + PReLU layers of IFBlock: what autograd runs under
// `loss_G.backward()` in Model.update — Flow-3D/model/RIFE.py:255-259, Flow-2D/model/RIFE.py:315-317 through
// Flow-{2D,3D}/model/IFNet.py:16-27 (conv = nn.Conv + nn.PReLU, deconv = nn.ConvTranspose + nn.PReLU).
//
//   ofsv_prelu_bias_bwd_bf16 : gradient through bias + per-channel PReLU of one layer: g_pre, d bias, d slope in one pass.
//   ofsv_conv_wgrad_bf16     : weight gradient of a layer in the tap form of ofsv_conv_desc,
//                                dW[ph*ntaps+t][ci][co] = sum_{n,o} x[n, o*in_stride + tap_off][ci] * g[n, o*out_stride + parity(ph)][co]
//                              as a bf16 tensor-core GEMM whose K axis is the output positions (fp32 accumulate), split over CTAs
//                              along K with a fixed-order second pass (deterministic, no atomics).
// The INPUT gradient of every layer type is itself a tap-form convolution (conv <-> transposed conv with the same weights) and
// runs on the forward engines (ofsv_conv_halo / ofsv_conv_tc) — opticalflowscivis_b200/train.py builds those descriptors.
#include <algorithm>
#include <cstring>

#include <cuda.h>

#include "ofsv_common.cuh"
#include "tc_common.cuh"

namespace ofsv {
namespace {

// ------------------------------------------------------------------------------------------------ bias + PReLU backward
constexpr int PB_THREADS = 256;

// y is the layer's output AFTER PReLU (what the forward pass kept): for slope > 0 the sign of the pre-activation is the sign of
// y, and the pre-activation of a negative output is y / slope.  (A non-positive slope would need the pre-activation itself;
// the reference initialises 0.25 and trains with lr ~1e-4: the host side checks slope > 0 when it re-packs the weights.)
__global__ void __launch_bounds__(PB_THREADS) prelu_bias_bwd_kernel(const __nv_bfloat16* __restrict__ gy, const __nv_bfloat16* __restrict__ y,
                                                                     const float* __restrict__ slope, __nv_bfloat16* __restrict__ gpre,
                                                                     float* __restrict__ partial, int64_t P, int Cs) {
  // thread = 8 consecutive channels of a row; rows strided over (thread rows, blocks)
  const int tpr = Cs / 8;                        // threads per row
  const int rows_per_pass = PB_THREADS / tpr;
  const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
  float db[8], ds[8], sl[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    db[i] = ds[i] = 0.f;
    sl[i] = slope ? slope[tc * 8 + i] : 1.f;
  }
  if (tr < rows_per_pass) {
    for (int64_t r = (int64_t)blockIdx.x * rows_per_pass + tr; r < P; r += (int64_t)gridDim.x * rows_per_pass) {
      const int64_t off = r * Cs + tc * 8;
      uint4 gv = *reinterpret_cast<const uint4*>(gy + off);
      const __nv_bfloat16* g8 = reinterpret_cast<const __nv_bfloat16*>(&gv);
      uint4 ov;
      __nv_bfloat16* o8 = reinterpret_cast<__nv_bfloat16*>(&ov);
      if (slope) {
        uint4 yv = *reinterpret_cast<const uint4*>(y + off);
        const __nv_bfloat16* y8 = reinterpret_cast<const __nv_bfloat16*>(&yv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float g = __bfloat162float(g8[i]), yy = __bfloat162float(y8[i]);
          const bool pos = yy > 0.f;
          const float gp = pos ? g : g * sl[i];
          db[i] += gp;
          ds[i] += pos ? 0.f : g * (yy / sl[i]);
          o8[i] = __float2bfloat16(gp);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          db[i] += __bfloat162float(g8[i]);
          o8[i] = g8[i];
        }
      }
      if (gpre != gy || slope) *reinterpret_cast<uint4*>(gpre + off) = ov;
    }
  }
  // block reduction over the thread rows, in a fixed order
  __shared__ float red[PB_THREADS * 16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x * 16 + i] = db[i];
    red[threadIdx.x * 16 + 8 + i] = ds[i];
  }
  __syncthreads();
  if (threadIdx.x < Cs * 2) {
    const int which = threadIdx.x / Cs, c = threadIdx.x % Cs;       // 0 = bias, 1 = slope
    float s = 0.f;
    for (int r = 0; r < rows_per_pass; ++r) s += red[(r * tpr + c / 8) * 16 + which * 8 + (c % 8)];
    partial[((int64_t)blockIdx.x * 2 + which) * Cs + c] = s;
  }
}

// one warp per output (bias | slope, channel): lanes stride over the per-CTA partials, then a fixed shuffle tree (deterministic)
__global__ void prelu_bias_bwd_finalize(const float* __restrict__ partial, float* __restrict__ dbias, float* __restrict__ dslope, int nblk,
                                        int Cs) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= 2 * Cs) return;
  const int which = i / Cs, c = i % Cs;
  float s = 0.f;
  for (int b = lane; b < nblk; b += 32) s += partial[((int64_t)b * 2 + which) * Cs + c];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (which == 0) dbias[c] = s;
    else if (dslope) dslope[c] = s;
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
constexpr int WG_THREADS = 128;
constexpr int WG_KC = 32;          // output positions per pipeline stage (two k16 MMA steps)

struct WgradParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* g;
  float* out;                      // [splits][T][Cin_s][Cout_w]
  int N, Di, Hi, Wi, Cin_s;
  int Do, Ho, Wo;
  int Dy, Hy, Wy, g_cs, Cout_w;
  int in_stride, out_stride;
  int T, ntaps, mt, nt;            // tiles along Cin_s / Cout_w
  int64_t K, Kper;                 // virtual output positions, positions per K split (multiple of WG_KC)
  int8_t tap[OFSV_MAX_TAPS][4];
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Warp grid over the MT x NT tile (see the table in ofsv_conv_wgrad_bf16).
template <int MT, int NT>
struct WgTile {
  static constexpr int WM = (NT == 16) ? (MT >= 64 ? 4 : (MT >= 32 ? 2 : 1)) : (MT >= 32 ? 2 : 1);
  static constexpr int WN = (NT == 16 && MT == 16) ? 2 : 4 / WM;
  static constexpr int WMT = MT / WM, WNT = NT / WN;     // warp tile
  static_assert(WMT % 16 == 0 && WNT % 8 == 0, "warp tile");
  static constexpr int XP = MT * 2 + 16, GP = NT * 2 + 16;   // row pitches in bytes (+16: ldmatrix rows fall in distinct banks)
  static constexpr int KC = 64;                               // positions per pipeline stage (four k16 MMA steps)
  static constexpr int STAGE = KC * (XP + GP);
};

template <int MT, int NT>
__global__ void __launch_bounds__(WG_THREADS) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  using TL = WgTile<MT, NT>;
  __shared__ __align__(16) unsigned char smem[2 * TL::STAGE];
  constexpr int KC = TL::KC;
  __shared__ int xrow_tab[2][KC], grow_tab[2][KC];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int tile = blockIdx.x;
  const int ntile = tile % p.nt; tile /= p.nt;
  const int mtile = tile % p.mt; tile /= p.mt;
  const int tap = tile;                                    // ph * ntaps + t
  const int ph = tap / p.ntaps;
  const int pz = (ph >> 2) & 1, py = (ph >> 1) & 1, px = ph & 1;
  const int oz = p.tap[tap][0], oy = p.tap[tap][1], ox = p.tap[tap][2];
  const int64_t k_begin = (int64_t)blockIdx.y * p.Kper;
  const int64_t k_end = min(p.K, k_begin + p.Kper);
  const int nchunk = (int)((k_end - k_begin + KC - 1) / KC);

  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);

  auto fill_table = [&](int chunk) {                      // threads 0..31: source rows of the positions of `chunk`
    if (tid < KC) {
      const int64_t q = k_begin + (int64_t)chunk * KC + tid;
      int xr = -1, gr = -1;
      if (q < k_end) {
        int r = (int)q;
        const int vx = r % p.Wo; r /= p.Wo;
        const int vy = r % p.Ho; r /= p.Ho;
        const int vz = r % p.Do; const int n = r / p.Do;
        gr = ((n * p.Dy + vz * p.out_stride + pz) * p.Hy + vy * p.out_stride + py) * p.Wy + vx * p.out_stride + px;
        const int iz = vz * p.in_stride + oz, iy = vy * p.in_stride + oy, ix = vx * p.in_stride + ox;
        if (iz >= 0 && iz < p.Di && iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) xr = ((n * p.Di + iz) * p.Hi + iy) * p.Wi + ix;
        else gr = -1;                                      // the product is zero either way: skip both loads
      }
      xrow_tab[chunk & 1][tid] = xr;
      grow_tab[chunk & 1][tid] = gr;
    }
  };
  auto issue_loads = [&](int chunk) {
    const uint32_t st = sbase + (chunk & 1) * TL::STAGE;
    constexpr int XPIECES = MT / 8, GPIECES = NT / 8;
    for (int i = tid; i < KC * XPIECES; i += WG_THREADS) {
      const int row = i / XPIECES, pc = i % XPIECES;
      const int xr = xrow_tab[chunk & 1][row];
      const __nv_bfloat16* src = xr >= 0 ? p.x + (int64_t)xr * p.Cin_s + mtile * MT + pc * 8 : p.x;
      cp_async16(st + row * TL::XP + pc * 16, src, xr >= 0 ? 16 : 0);
    }
    for (int i = tid; i < KC * GPIECES; i += WG_THREADS) {
      const int row = i / GPIECES, pc = i % GPIECES;
      const int gr = grow_tab[chunk & 1][row];
      const __nv_bfloat16* src = gr >= 0 ? p.g + (int64_t)gr * p.g_cs + ntile * NT + pc * 8 : p.g;
      cp_async16(st + KC * TL::XP + row * TL::GP + pc * 16, src, gr >= 0 ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  constexpr int MI = TL::WMT / 16, NI = TL::WNT / 8;
  float acc[MI][NI][4];
#pragma unroll
  for (int a = 0; a < MI; ++a)
#pragma unroll
    for (int b = 0; b < NI; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
  const bool active = warp < TL::WM * TL::WN;
  const int wm = warp / TL::WN, wn = warp % TL::WN;
  const int m0 = wm * TL::WMT, n0 = wn * TL::WNT;

  if (nchunk > 0) {
    fill_table(0);
    __syncthreads();
    issue_loads(0);
    if (nchunk > 1) fill_table(1);
    for (int c = 0; c < nchunk; ++c) {
      __syncthreads();                                     // table(c+1) visible; stage (c+1)&1 no longer read by anyone
      if (c + 1 < nchunk) {
        issue_loads(c + 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();                                     // chunk c landed for every thread; table slot c&1 is free again
      if (c + 2 < nchunk) fill_table(c + 2);
      if (active) {
        const uint32_t xs = sbase + (c & 1) * TL::STAGE, gs = xs + KC * TL::XP;
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks) {
          uint32_t afr[MI][4];
          const int mat = lane >> 3, r = lane & 7;
#pragma unroll
          for (int a = 0; a < MI; ++a) {
            // A = x^T: matrices (m-block, k-block) = (mat&1, mat>>1); stored [k][m] -> .trans
            const uint32_t addr = xs + (ks * 16 + (mat >> 1) * 8 + r) * TL::XP + (m0 + a * 16 + (mat & 1) * 8) * 2;
            ldmatrix_x4_t(addr, afr[a][0], afr[a][1], afr[a][2], afr[a][3]);
          }
          if constexpr (NI % 2 == 0) {
#pragma unroll
            for (int b = 0; b < NI; b += 2) {
              // B = g: matrices (k-block, n-tile) = (mat&1, mat>>1); stored [k][n] -> .trans
              uint32_t b0, b1, b2, b3;
              const uint32_t addr = gs + (ks * 16 + (mat & 1) * 8 + r) * TL::GP + (n0 + (b + (mat >> 1)) * 8) * 2;
              ldmatrix_x4_t(addr, b0, b1, b2, b3);
#pragma unroll
              for (int a = 0; a < MI; ++a) {
                mma_bf16_16816(acc[a][b], afr[a], b0, b1);
                mma_bf16_16816(acc[a][b + 1], afr[a], b2, b3);
              }
            }
          } else {
            static_assert(NI == 1 || NI % 2 == 0, "NI");
            uint32_t b0, b1;
            const uint32_t addr = gs + (ks * 16 + ((lane >> 3) & 1) * 8 + r) * TL::GP + n0 * 2;
            ldmatrix_x2_t(addr, b0, b1);
#pragma unroll
            for (int a = 0; a < MI; ++a) mma_bf16_16816(acc[a][0], afr[a], b0, b1);
          }
        }
      }
    }
  }
  if (active) {
    float* out = p.out + (((int64_t)blockIdx.y * p.T + tap) * p.Cin_s + mtile * MT) * p.Cout_w + ntile * NT;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int a = 0; a < MI; ++a)
#pragma unroll
      for (int b = 0; b < NI; ++b) {
        const int row = m0 + a * 16 + g, col = n0 + b * 8 + 2 * t;
        *reinterpret_cast<float2*>(out + (int64_t)row * p.Cout_w + col) = make_float2(acc[a][b][0], acc[a][b][1]);
        *reinterpret_cast<float2*>(out + (int64_t)(row + 8) * p.Cout_w + col) = make_float2(acc[a][b][2], acc[a][b][3]);
      }
  }
}

// ---- tap-group variant -------------------------------------------------------------------------------------------------------
// For layers with many taps and small channel tiles (conv0.0: 64 taps of 16 x 32, the heads: 8 x 8 taps of 64 x 16) the kernel above
// spends its time re-loading the gradient rows and synchronising once per tap.  Here a CTA owns TG = 4 * TPW taps of one phase:
// the 32 gradient rows of a position chunk are loaded ONCE and their B fragments reused for every tap, each warp accumulates the
// full MT x NT tile of its TPW taps (A fragments from that tap's shifted input rows), and the position -> (n, z, y, x) divisions
// are done once per row and shared by the taps.
template <int MT, int NT, int TPW>
struct WgTg {
  static constexpr int TG = 4 * TPW;
  static constexpr int XP = MT * 2 + 16, GP = NT * 2 + 16;
  static constexpr int STAGE = WG_KC * GP + TG * WG_KC * XP;
  static constexpr int MI = MT / 16, NI = NT / 8;
  static_assert(TPW * MI * NI * 4 <= 64, "accumulator registers");
};

template <int MT, int NT, int TPW>
__global__ void __launch_bounds__(WG_THREADS) conv_wgrad_tg_kernel(const __grid_constant__ WgradParams p, int ngroups) {
  using TL = WgTg<MT, NT, TPW>;
  constexpr int TG = TL::TG, MI = TL::MI, NI = TL::NI;
  extern __shared__ __align__(16) unsigned char dsmem[];
  __shared__ int xrow_tab[2][TG][WG_KC], grow_tab[2][WG_KC];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int tile = blockIdx.x;
  const int ntile = tile % p.nt; tile /= p.nt;
  const int mtile = tile % p.mt; tile /= p.mt;
  const int grp = tile % ngroups, ph = tile / ngroups;
  const int tap0 = grp * TG;                               // first tap (within the phase) of this CTA
  const int ntap = min(TG, p.ntaps - tap0);
  const int pz = (ph >> 2) & 1, py = (ph >> 1) & 1, px = ph & 1;
  const int64_t k_begin = (int64_t)blockIdx.y * p.Kper;
  const int64_t k_end = min(p.K, k_begin + p.Kper);
  const int nchunk = (int)((k_end - k_begin + WG_KC - 1) / WG_KC);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dsmem);

  auto fill_table = [&](int chunk) {                      // thread = (row, tap quarter): one position decomposition, TPW taps
    const int row = tid & 31, tq = tid >> 5;
    const int64_t q = k_begin + (int64_t)chunk * WG_KC + row;
    int n = 0, vz = 0, vy = 0, vx = 0;
    const bool live = q < k_end;
    if (live) {
      int r = (int)q;
      vx = r % p.Wo; r /= p.Wo;
      vy = r % p.Ho; r /= p.Ho;
      vz = r % p.Do; n = r / p.Do;
    }
    if (tq == 0)
      grow_tab[chunk & 1][row] = live ? ((n * p.Dy + vz * p.out_stride + pz) * p.Hy + vy * p.out_stride + py) * p.Wy + vx * p.out_stride + px : -1;
#pragma unroll
    for (int j = 0; j < TPW; ++j) {
      const int tl = tq + 4 * j;
      int xr = -1;
      if (live && tl < ntap) {
        const int8_t* o = p.tap[ph * p.ntaps + tap0 + tl];
        const int iz = vz * p.in_stride + o[0], iy = vy * p.in_stride + o[1], ix = vx * p.in_stride + o[2];
        if (iz >= 0 && iz < p.Di && iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) xr = ((n * p.Di + iz) * p.Hi + iy) * p.Wi + ix;
      }
      xrow_tab[chunk & 1][tl][row] = xr;
    }
  };
  auto issue_loads = [&](int chunk) {
    const uint32_t st = sbase + (chunk & 1) * TL::STAGE;
    constexpr int XPIECES = MT / 8, GPIECES = NT / 8;
    for (int i = tid; i < WG_KC * GPIECES; i += WG_THREADS) {
      const int row = i / GPIECES, pc = i % GPIECES;
      const int gr = grow_tab[chunk & 1][row];
      const __nv_bfloat16* src = gr >= 0 ? p.g + (int64_t)gr * p.g_cs + ntile * NT + pc * 8 : p.g;
      cp_async16(st + row * TL::GP + pc * 16, src, gr >= 0 ? 16 : 0);
    }
    const uint32_t xs = st + WG_KC * TL::GP;
    for (int i = tid; i < ntap * WG_KC * XPIECES; i += WG_THREADS) {
      const int tl = i / (WG_KC * XPIECES), rem = i % (WG_KC * XPIECES);
      const int row = rem / XPIECES, pc = rem % XPIECES;
      const int xr = xrow_tab[chunk & 1][tl][row];
      const __nv_bfloat16* src = xr >= 0 ? p.x + (int64_t)xr * p.Cin_s + mtile * MT + pc * 8 : p.x;
      cp_async16(xs + (tl * WG_KC + row) * TL::XP + pc * 16, src, xr >= 0 ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[TPW][MI][NI][4];
#pragma unroll
  for (int j = 0; j < TPW; ++j)
#pragma unroll
    for (int a = 0; a < MI; ++a)
#pragma unroll
      for (int b = 0; b < NI; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[j][a][b][c] = 0.f;

  if (nchunk > 0) {
    fill_table(0);
    __syncthreads();
    issue_loads(0);
    if (nchunk > 1) fill_table(1);
    for (int c = 0; c < nchunk; ++c) {
      __syncthreads();
      if (c + 1 < nchunk) {
        issue_loads(c + 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      if (c + 2 < nchunk) fill_table(c + 2);
      const uint32_t gs = sbase + (c & 1) * TL::STAGE, xs0 = gs + WG_KC * TL::GP;
      const int mat = lane >> 3, r = lane & 7;
#pragma unroll
      for (int ks = 0; ks < WG_KC / 16; ++ks) {
        uint32_t bfr[NI][2];
        if constexpr (NI % 2 == 0) {
#pragma unroll
          for (int b = 0; b < NI; b += 2)
            ldmatrix_x4_t(gs + (ks * 16 + (mat & 1) * 8 + r) * TL::GP + ((b + (mat >> 1)) * 8) * 2, bfr[b][0], bfr[b][1], bfr[b + 1][0], bfr[b + 1][1]);
        } else {
          ldmatrix_x2_t(gs + (ks * 16 + (mat & 1) * 8 + r) * TL::GP, bfr[0][0], bfr[0][1]);
        }
#pragma unroll
        for (int j = 0; j < TPW; ++j) {
          const int tl = warp + 4 * j;
          if (tl < ntap) {                                 // warp-uniform
            const uint32_t xs = xs0 + tl * WG_KC * TL::XP;
#pragma unroll
            for (int a = 0; a < MI; ++a) {
              uint32_t afr[4];
              ldmatrix_x4_t(xs + (ks * 16 + (mat >> 1) * 8 + r) * TL::XP + (a * 16 + (mat & 1) * 8) * 2, afr[0], afr[1], afr[2], afr[3]);
#pragma unroll
              for (int b = 0; b < NI; ++b) mma_bf16_16816(acc[j][a][b], afr, bfr[b][0], bfr[b][1]);
            }
          }
        }
      }
    }
  }
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int j = 0; j < TPW; ++j) {
    const int tl = warp + 4 * j;
    if (tl >= ntap) continue;
    const int tap = ph * p.ntaps + tap0 + tl;
    float* out = p.out + (((int64_t)blockIdx.y * p.T + tap) * p.Cin_s + mtile * MT) * p.Cout_w + ntile * NT;
#pragma unroll
    for (int a = 0; a < MI; ++a)
#pragma unroll
      for (int b = 0; b < NI; ++b) {
        const int row = a * 16 + g, col = b * 8 + 2 * t;
        *reinterpret_cast<float2*>(out + (int64_t)row * p.Cout_w + col) = make_float2(acc[j][a][b][0], acc[j][a][b][1]);
        *reinterpret_cast<float2*>(out + (int64_t)(row + 8) * p.Cout_w + col) = make_float2(acc[j][a][b][2], acc[j][a][b][3]);
      }
  }
}

// ---- brick-window variant ----------------------------------------------------------------------------------------------------
// The two kernels above load the input row of every (position, tap) separately: 64 taps read the input 64 times (2.1 GB of L2 -> SM
// traffic for the heads of a 64^3 block, 0.53 ms).  Here a CTA owns a 16 x 16 channel tile for ALL taps and walks BRICKS of 4 x 4 x 4
// (2-D: 8 x 8) output positions: the input WINDOW a brick can touch (6^3 rows for a 3^3 / transposed layer, 10^3 for a 4^3 stride-2
// conv) and the brick's gradient rows are staged in shared memory once, and every tap's A fragment is an ldmatrix whose eight row
// addresses are window rows — window row = (position part) + (tap part), both linear, so a tap costs one add per fragment.  The
// eight warps split the taps (<= 8 each: 64 accumulator registers), double-buffered windows, fixed-order K split as above.
// Measured on the layers of one 8 x 64^3 step (tests/bench_wgrad.py, profiles/r02x_bench_wgrad.txt): the windows cut the L2 -> SM
// bytes 6-17x, but with 16-channel tiles every window row is a 32-byte request and the kernel is bound by the REQUEST rate of the
// TMA unit (and of cp.async before it: the same times): heads 458 us vs 535 us for the tap-group kernel, conv0.0 246 vs 232, the
// 64-channel layers 1.7-2.6x slower.  Policy: -1 (default) = only where it wins (Cout_w == 16 and >= 32 taps: the heads);
// ofsv_set_tuning("wgrad_brick", 1 | 0) forces it on (wherever the windows fit) / off.
std::atomic<int> g_wgrad_brick{-1};
constexpr int WB_THREADS = 256, WB_PITCH = 32;    // window rows are the 16 channels of the tile: dense, as TMA writes them

struct WbParams {
  float* out;
  int N, Cin_s, Cout_w;
  int s, os, T, ntaps, mt, nt;
  int lby, lbx, bz, by, bx;         // brick dims (bz * by * bx == 64) and log2 of by, bx
  int nbz, nby, nbx, nbricks;       // bricks per axis, N * nbz * nby * nbx
  int ominz, ominy, ominx;          // smallest tap offset per axis
  int wz, wy, wx, gz, gy, gx;       // window dims (rows) of the input and of the gradient
  int8_t tap[OFSV_MAX_TAPS][4];
};

__global__ void __launch_bounds__(WB_THREADS) conv_wgrad_brick_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                                                                       const __grid_constant__ WbParams p) {
  extern __shared__ __align__(128) unsigned char wsmem[];
  __shared__ __align__(8) uint64_t full[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntile = blockIdx.x % p.nt, mtile = blockIdx.x / p.nt;
  const int xrows = p.wz * p.wy * p.wx, grows = p.gz * p.gy * p.gx;
  const int xbytes = (xrows * WB_PITCH + 127) & ~127;                 // TMA destinations are 128-byte aligned
  const int stage_bytes = xbytes + ((grows * WB_PITCH + 127) & ~127);
  const uint32_t sbase = (smem_u32(wsmem) + 127u) & ~127u;

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmG)) : "memory");
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // one thread issues the two window loads of a brick: a 5-D box each, out-of-volume rows zero-filled by the TMA unit
  auto issue = [&](int brick, int st) {
    int r = brick;
    const int bxi = r % p.nbx; r /= p.nbx;
    const int byi = r % p.nby; r /= p.nby;
    const int bzi = r % p.nbz; const int n = r / p.nbz;
    const uint32_t xs = sbase + st * stage_bytes, gs = xs + xbytes;
    mbar_expect_tx(&full[st], (uint32_t)((xrows + grows) * WB_PITCH));
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(xs),
                 "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(smem_u32(&full[st])), "r"(mtile * 16), "r"(p.s * bxi * p.bx + p.ominx),
                 "r"(p.s * byi * p.by + p.ominy), "r"(p.s * bzi * p.bz + p.ominz), "r"(n)
                 : "memory");
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(gs),
                 "l"(reinterpret_cast<uint64_t>(&tmG)), "r"(smem_u32(&full[st])), "r"(ntile * 16), "r"(p.os * bxi * p.bx), "r"(p.os * byi * p.by),
                 "r"(p.os * bzi * p.bz), "r"(n)
                 : "memory");
  };

  // per-lane position parts of the fragment row addresses, for the four k16 steps of a brick (64 positions)
  const int mat = lane >> 3, r8 = lane & 7;
  uint32_t a_pos[4], b_pos[4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int pa = ks * 16 + (mat >> 1) * 8 + r8;          // A: matrices (m-block, k-block) = (mat & 1, mat >> 1)
    const int az = pa >> (p.lby + p.lbx), ay = (pa >> p.lbx) & (p.by - 1), ax = pa & (p.bx - 1);
    a_pos[ks] = (uint32_t)(((p.s * az) * p.wy + p.s * ay) * p.wx + p.s * ax) * WB_PITCH + (mat & 1) * 16;
    const int pb = ks * 16 + (mat & 1) * 8 + r8;           // B: matrices (k-block, n8-tile) = (mat & 1, mat >> 1)
    const int bz_ = pb >> (p.lby + p.lbx), by_ = (pb >> p.lbx) & (p.by - 1), bx_ = pb & (p.bx - 1);
    b_pos[ks] = (uint32_t)(((p.os * bz_) * p.gy + p.os * by_) * p.gx + p.os * bx_) * WB_PITCH + (mat >> 1) * 16;
  }
  // this warp's taps: tap = warp + 8 j; tap parts of the addresses
  uint32_t a_tap[8], b_tap[8];
  int ntw = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int tap = warp + 8 * j;
    a_tap[j] = b_tap[j] = 0;
    if (tap < p.T) {
      ntw = j + 1;
      const int ph = tap / p.ntaps;
      a_tap[j] = (uint32_t)(((p.tap[tap][0] - p.ominz) * p.wy + (p.tap[tap][1] - p.ominy)) * p.wx + (p.tap[tap][2] - p.ominx)) * WB_PITCH;
      b_tap[j] = (uint32_t)((((ph >> 2) & 1) * p.gy + ((ph >> 1) & 1)) * p.gx + (ph & 1)) * WB_PITCH;
    }
  }
  float acc[8][2][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][b][c] = 0.f;

  int brick = blockIdx.y, it = 0;
  if (tid == 0 && brick < p.nbricks) issue(brick, 0);
  for (; brick < p.nbricks; brick += gridDim.y, ++it) {
    const int st = it & 1;
    if (tid == 0 && brick + (int)gridDim.y < p.nbricks) issue(brick + gridDim.y, st ^ 1);   // stage st^1 was released by the barrier below
    mbar_wait(&full[st], (uint32_t)(it >> 1) & 1u);
    const uint32_t xs = sbase + st * stage_bytes, gs = xs + xbytes;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < ntw) {                                        // warp-uniform
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t afr[4], b0, b1, b2, b3;
          ldmatrix_x4_t(xs + a_pos[ks] + a_tap[j], afr[0], afr[1], afr[2], afr[3]);
          ldmatrix_x4_t(gs + b_pos[ks] + b_tap[j], b0, b1, b2, b3);
          mma_bf16_16816(acc[j][0], afr, b0, b1);
          mma_bf16_16816(acc[j][1], afr, b2, b3);
        }
      }
    }
    __syncthreads();                                        // every warp is done with stage st before it is loaded again
  }
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int tap = warp + 8 * j;
    if (tap >= p.T) continue;
    float* out = p.out + (((int64_t)blockIdx.y * p.T + tap) * p.Cin_s + mtile * 16) * p.Cout_w + ntile * 16;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int col = b * 8 + 2 * t;
      *reinterpret_cast<float2*>(out + (int64_t)g * p.Cout_w + col) = make_float2(acc[j][b][0], acc[j][b][1]);
      *reinterpret_cast<float2*>(out + (int64_t)(g + 8) * p.Cout_w + col) = make_float2(acc[j][b][2], acc[j][b][3]);
    }
  }
}

__global__ void conv_wgrad_finalize(const float4* __restrict__ work, float4* __restrict__ dw, int64_t n4, int splits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 s = work[i];
  for (int k = 1; k < splits; ++k) {
    const float4 v = work[(int64_t)k * n4 + i];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  dw[i] = s;
}

int tile_of(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : 16); }

int wgrad_validate(const ofsv_conv_desc* d) {
  OFSV_REQUIRE(d != nullptr, "conv_wgrad: null descriptor");
  OFSV_REQUIRE(d->nd == 2 || d->nd == 3, "conv_wgrad: nd must be 2 or 3");
  OFSV_REQUIRE(d->Cin_s > 0 && d->Cin_s % 16 == 0 && d->Cout_w > 0 && d->Cout_w % 16 == 0, "conv_wgrad: Cin_s / Cout_w must be multiples of 16");
  OFSV_REQUIRE(d->nphase >= 1 && d->ntaps >= 1 && d->nphase * d->ntaps <= OFSV_MAX_TAPS, "conv_wgrad: too many taps");
  OFSV_REQUIRE(d->nphase == 1 || d->nphase == (1 << d->nd), "conv_wgrad: nphase must be 1 or 2^nd");
  OFSV_REQUIRE(!d->out_shuffle && !d->out_s2d, "conv_wgrad: depth-to-space / space-to-depth outputs have no weight-gradient form (use the phase form)");
  OFSV_REQUIRE(d->N > 0 && d->Do > 0 && d->Ho > 0 && d->Wo > 0, "conv_wgrad: empty output grid");
  const int64_t K = (int64_t)d->N * d->Do * d->Ho * d->Wo;
  const int64_t xin = (int64_t)d->N * d->Di * d->Hi * d->Wi, yo = (int64_t)d->N * d->Dy * d->Hy * d->Wy;
  OFSV_REQUIRE(K < (1ll << 31) && xin < (1ll << 31) && yo < (1ll << 31), "conv_wgrad: more than 2^31 positions");
  const int par = d->nphase > 1 ? 1 : 0;
  OFSV_REQUIRE((d->nd == 2 || (d->Do - 1) * d->out_stride + par < d->Dy) && (d->Ho - 1) * d->out_stride + par < d->Hy &&
                   (d->Wo - 1) * d->out_stride + par < d->Wy,
               "conv_wgrad: virtual output grid exceeds the output tensor");
  return OFSV_OK;
}

// Taps per warp of the tap-group kernel for an MT x NT tile (0 = use the per-tap kernel): 64 accumulator registers per thread.
int tg_tpw(int MT, int NT) {
  const int per_tap = (MT / 16) * (NT / 8) * 4;
  return per_tap > 64 ? 0 : (64 / per_tap > 4 ? 4 : 64 / per_tap);
}
struct WgPlan { int MT, NT, tpw, ngroups, splits; int64_t tiles; bool brick; WbParams wb; size_t wb_smem; };
// Brick-window kernel: any layer with >= 4 taps whose windows fit shared memory (all layers of the IFBlocks do).
bool wgrad_brick_plan(const ofsv_conv_desc* d, WbParams* w, size_t* smem) {
  const int T = d->nphase * d->ntaps;
  if (d->ntaps < 4 || T > OFSV_MAX_TAPS) return false;
  int omin[3] = {127, 127, 127}, omax[3] = {-127, -127, -127};
  for (int i = 0; i < T; ++i)
    for (int a = 0; a < 3; ++a) {
      const int o = (d->nd == 2 && a == 0) ? 0 : d->tap_off[i][a];
      omin[a] = std::min(omin[a], o); omax[a] = std::max(omax[a], o);
    }
  memset(w, 0, sizeof(*w));
  w->s = d->in_stride; w->os = d->out_stride; w->T = T; w->ntaps = d->ntaps;
  if (d->nd == 3) { w->bz = 4; w->by = 4; w->bx = 4; w->lby = 2; w->lbx = 2; }
  else { w->bz = 1; w->by = 8; w->bx = 8; w->lby = 3; w->lbx = 3; }
  const int Do = d->nd == 2 ? 1 : d->Do;
  w->nbz = (int)cdiv(Do, w->bz); w->nby = (int)cdiv(d->Ho, w->by); w->nbx = (int)cdiv(d->Wo, w->bx);
  const int64_t nb = (int64_t)d->N * w->nbz * w->nby * w->nbx;
  if (nb >= (1ll << 31)) return false;
  w->nbricks = (int)nb;
  w->ominz = omin[0]; w->ominy = omin[1]; w->ominx = omin[2];
  w->wz = w->s * (w->bz - 1) + omax[0] - omin[0] + 1;
  w->wy = w->s * (w->by - 1) + omax[1] - omin[1] + 1;
  w->wx = w->s * (w->bx - 1) + omax[2] - omin[2] + 1;
  const int span = d->nphase > 1 ? w->os : 1;
  w->gz = d->nd == 2 ? 1 : w->os * (w->bz - 1) + span;
  w->gy = w->os * (w->by - 1) + span;
  w->gx = w->os * (w->bx - 1) + span;
  const size_t xb = ((size_t)w->wz * w->wy * w->wx * WB_PITCH + 127) & ~(size_t)127, gb = ((size_t)w->gz * w->gy * w->gx * WB_PITCH + 127) & ~(size_t)127;
  *smem = 2 * (xb + gb) + 128;                             // + alignment slack
  return *smem <= 200 * 1024 && w->wx <= 256 && w->wy <= 256 && w->wz <= 256 && w->gx <= 256 && w->gy <= 256 && w->gz <= 256;
}
WgPlan wgrad_plan(const ofsv_conv_desc* d) {
  WgPlan pl;
  const int64_t K = (int64_t)d->N * d->Do * d->Ho * d->Wo;
  const int mode = g_wgrad_brick.load(std::memory_order_relaxed);
  pl.brick = (mode > 0 || (mode < 0 && d->Cout_w == 16 && d->nphase * d->ntaps >= 32)) && wgrad_brick_plan(d, &pl.wb, &pl.wb_smem);
  if (pl.brick) {
    pl.MT = pl.NT = 16; pl.tpw = 0; pl.ngroups = 0;
    pl.tiles = (int64_t)(d->Cin_s / 16) * (d->Cout_w / 16);
    const int64_t want = cdiv((int64_t)device_num_sms() * 2, pl.tiles);
    pl.splits = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, pl.wb.nbricks), 64));
    return pl;
  }
  pl.MT = tile_of(d->Cin_s); pl.NT = tile_of(d->Cout_w);
  pl.tpw = d->ntaps >= 4 ? tg_tpw(pl.MT, pl.NT) : 0;
  const int64_t mn = (int64_t)(d->Cin_s / pl.MT) * (d->Cout_w / pl.NT);
  if (pl.tpw) {
    pl.ngroups = (int)cdiv(d->ntaps, 4 * pl.tpw);
    pl.tiles = (int64_t)d->nphase * pl.ngroups * mn;
  } else {
    pl.ngroups = 0;
    pl.tiles = (int64_t)d->nphase * d->ntaps * mn;
  }
  const int64_t want = cdiv((int64_t)device_num_sms() * (pl.tpw ? 4 : 8), pl.tiles);
  int64_t s = std::min<int64_t>(want, cdiv(K, 8 * WG_KC));
  pl.splits = (int)std::max<int64_t>(1, std::min<int64_t>(s, 64));
  return pl;
}

template <int MT, int NT>
void launch_wgrad(const WgradParams& p, dim3 grid, cudaStream_t st) {
  conv_wgrad_kernel<MT, NT><<<grid, WG_THREADS, 0, st>>>(p);
}
template <int MT, int NT, int TPW>
int launch_wgrad_tg(const WgradParams& p, dim3 grid, int ngroups, cudaStream_t st) {
  static std::atomic<uint64_t> attr_done{0};
  constexpr int smem = 2 * WgTg<MT, NT, TPW>::STAGE;
  if (int e = ensure_dyn_smem(attr_done, conv_wgrad_tg_kernel<MT, NT, TPW>, smem, "ofsv_conv_wgrad_bf16")) return e;
  conv_wgrad_tg_kernel<MT, NT, TPW><<<grid, WG_THREADS, smem, st>>>(p, ngroups);
  return OFSV_OK;
}

}  // namespace
void ofsv_set_wgrad_brick(int v) { g_wgrad_brick.store(v, std::memory_order_relaxed); }
}  // namespace ofsv

namespace ofsv {
namespace {
// ------------------------------------------------------------------------------------------------ per-step weight refresh
__global__ void __launch_bounds__(256) conv_refresh_kernel(const ofsv_refresh_rec* __restrict__ recs) {
  const ofsv_refresh_rec& r = recs[blockIdx.y];
  if (r.kind == 1) {                                      // vector copy (bias, PReLU slopes)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < r.n; i += gridDim.x * blockDim.x) r.dst[i] = r.src[i];
    return;
  }
  const int cin = r.swap ? r.B : r.A, cout = r.swap ? r.A : r.B;
  const int64_t total = (int64_t)r.T * cin * cout;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(e % cout);
    const int64_t q = e / cout;
    const int i = (int)(q % cin), t = (int)(q / cin);
    const int k = r.kidx[t];
    const int a = r.swap ? o : i, b = r.swap ? i : o;
    r.dst[((int64_t)t * r.Cin_s + r.ci0 + i) * r.Cout_w + r.co0 + o] = k >= 0 ? __ldg(r.src + ((int64_t)a * r.B + b) * r.K + k) : 0.f;
  }
}
}  // namespace
}  // namespace ofsv

using namespace ofsv;

extern "C" int ofsv_conv_refresh_tapform(const ofsv_refresh_rec* recs_dev, int nrec, void* stream) {
  OFSV_REQUIRE(nrec >= 0 && nrec <= 65535, "ofsv_conv_refresh_tapform: bad record count");
  if (nrec == 0) return OFSV_OK;
  OFSV_REQUIRE(recs_dev != nullptr && (reinterpret_cast<uintptr_t>(recs_dev) & 7u) == 0, "ofsv_conv_refresh_tapform: null / misaligned record table");
  conv_refresh_kernel<<<dim3(32, (unsigned)nrec), 256, 0, static_cast<cudaStream_t>(stream)>>>(recs_dev);
  return check_launch("conv_refresh_kernel");
}

extern "C" int ofsv_prelu_bias_bwd_blocks(void) { return 2 * device_num_sms(); }

extern "C" int ofsv_prelu_bias_bwd_bf16(const void* gy, const void* y, const float* slope, void* gpre, float* dbias, float* dslope,
                                        float* work, int64_t P, int Cs, void* stream) {
  OFSV_REQUIRE(gy && gpre && dbias && work, "prelu_bias_bwd: null pointer");
  OFSV_REQUIRE(!slope || (y && dslope), "prelu_bias_bwd: a slope needs the layer output y and a dslope buffer");
  OFSV_REQUIRE(P > 0 && Cs >= 8 && Cs % 8 == 0 && Cs <= 128, "prelu_bias_bwd: need P > 0 and 8 <= Cs <= 128, Cs %% 8 == 0 (got P=%lld Cs=%d)", (long long)P, Cs);
  OFSV_REQUIRE(aligned16(gy) && aligned16(gpre) && (!y || aligned16(y)), "prelu_bias_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows_per_pass = PB_THREADS / (Cs / 8);
  int nblk = (int)std::min<int64_t>(ofsv_prelu_bias_bwd_blocks(), cdiv(P, rows_per_pass));
  prelu_bias_bwd_kernel<<<nblk, PB_THREADS, 0, st>>>(static_cast<const __nv_bfloat16*>(gy), static_cast<const __nv_bfloat16*>(y), slope,
                                                        static_cast<__nv_bfloat16*>(gpre), work, P, Cs);
  int rc = check_launch("prelu_bias_bwd_kernel");
  if (rc) return rc;
  prelu_bias_bwd_finalize<<<(int)cdiv(2 * Cs * 32, 256), 256, 0, st>>>(work, dbias, slope ? dslope : nullptr, nblk, Cs);
  return check_launch("prelu_bias_bwd_finalize");
}

extern "C" int ofsv_conv_wgrad_splits(const ofsv_conv_desc* d) {
  int rc = wgrad_validate(d);
  if (rc) return rc;
  return wgrad_plan(d).splits;
}

extern "C" int ofsv_conv_wgrad_bf16(const ofsv_conv_desc* d, const void* x, const void* gy, int gy_cs, float* dw, float* work, void* stream) {
  int rc = wgrad_validate(d);
  if (rc) return rc;
  OFSV_REQUIRE(x && gy && dw, "conv_wgrad: null pointer");
  OFSV_REQUIRE(gy_cs >= d->Cout_w && gy_cs % 8 == 0, "conv_wgrad: gy_cs (%d) must be a multiple of 8 and >= Cout_w (%d)", gy_cs, d->Cout_w);
  OFSV_REQUIRE(aligned16(x) && aligned16(gy) && aligned16(dw) && (!work || aligned16(work)), "conv_wgrad: pointers must be 16-byte aligned");
  WgPlan pl = wgrad_plan(d);
  const int splits = pl.splits;
  OFSV_REQUIRE(splits == 1 || work, "conv_wgrad: %d K splits need a work buffer of ofsv_conv_wgrad_splits(d) * T * Cin_s * Cout_w floats", splits);
  if (pl.brick) {
    WbParams& w = pl.wb;
    w.out = splits == 1 ? dw : work;
    w.N = d->N; w.Cin_s = d->Cin_s; w.Cout_w = d->Cout_w;
    w.mt = d->Cin_s / 16; w.nt = d->Cout_w / 16;
    for (int i = 0; i < w.T; ++i) {
      for (int j = 0; j < 4; ++j) w.tap[i][j] = d->tap_off[i][j];
      if (d->nd == 2) w.tap[i][0] = 0;
    }
    PFN_encodeTiled encode = get_tensor_map_encoder();
    if (!encode) { set_error("ofsv_conv_wgrad_bf16: cuTensorMapEncodeTiled unavailable (driver too old?)"); return OFSV_ECUDA; }
    CUtensorMap tmX, tmG;
    {
      const int Di = d->nd == 2 ? 1 : d->Di, Dy = d->nd == 2 ? 1 : d->Dy;
      const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      const cuuint64_t xdim[5] = {(cuuint64_t)d->Cin_s, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)Di, (cuuint64_t)d->N};
      const cuuint64_t xstr[4] = {(cuuint64_t)d->Cin_s * 2, (cuuint64_t)d->Wi * d->Cin_s * 2, (cuuint64_t)d->Hi * d->Wi * d->Cin_s * 2,
                                  (cuuint64_t)Di * d->Hi * d->Wi * d->Cin_s * 2};
      const cuuint32_t xbox[5] = {16, (cuuint32_t)w.wx, (cuuint32_t)w.wy, (cuuint32_t)w.wz, 1};
      CUresult r = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), xdim, xstr, xbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("ofsv_conv_wgrad_bf16: cuTensorMapEncodeTiled(x) failed with %d", (int)r); return OFSV_ECUDA; }
      const cuuint64_t gdim[5] = {(cuuint64_t)gy_cs, (cuuint64_t)d->Wy, (cuuint64_t)d->Hy, (cuuint64_t)Dy, (cuuint64_t)d->N};
      const cuuint64_t gstr[4] = {(cuuint64_t)gy_cs * 2, (cuuint64_t)d->Wy * gy_cs * 2, (cuuint64_t)d->Hy * d->Wy * gy_cs * 2,
                                  (cuuint64_t)Dy * d->Hy * d->Wy * gy_cs * 2};
      const cuuint32_t gbox[5] = {16, (cuuint32_t)w.gx, (cuuint32_t)w.gy, (cuuint32_t)w.gz, 1};
      r = encode(&tmG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(gy), gdim, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("ofsv_conv_wgrad_bf16: cuTensorMapEncodeTiled(gy) failed with %d", (int)r); return OFSV_ECUDA; }
    }
    static std::atomic<uint64_t> attr_done{0};
    if (int e = ensure_dyn_smem(attr_done, conv_wgrad_brick_kernel, 200 * 1024, "ofsv_conv_wgrad_bf16")) return e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    conv_wgrad_brick_kernel<<<dim3((unsigned)pl.tiles, (unsigned)splits), WB_THREADS, pl.wb_smem, st>>>(tmX, tmG, w);
    rc = check_launch("conv_wgrad_brick_kernel");
    if (rc) return rc;
    if (splits > 1) {
      const int64_t n4 = (int64_t)w.T * w.Cin_s * w.Cout_w / 4;
      conv_wgrad_finalize<<<(unsigned)cdiv(n4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(work), reinterpret_cast<float4*>(dw), n4, splits);
      rc = check_launch("conv_wgrad_finalize");
    }
    return rc;
  }
  WgradParams p;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.g = static_cast<const __nv_bfloat16*>(gy);
  p.out = splits == 1 ? dw : work;
  p.N = d->N; p.Di = d->nd == 2 ? 1 : d->Di; p.Hi = d->Hi; p.Wi = d->Wi; p.Cin_s = d->Cin_s;
  p.Do = d->nd == 2 ? 1 : d->Do; p.Ho = d->Ho; p.Wo = d->Wo;
  p.Dy = d->nd == 2 ? 1 : d->Dy; p.Hy = d->Hy; p.Wy = d->Wy; p.g_cs = gy_cs; p.Cout_w = d->Cout_w;
  p.in_stride = d->in_stride; p.out_stride = d->out_stride;
  p.T = d->nphase * d->ntaps; p.ntaps = d->ntaps;
  const int MT = pl.MT, NT = pl.NT;
  p.mt = d->Cin_s / MT; p.nt = d->Cout_w / NT;
  p.K = (int64_t)p.N * p.Do * p.Ho * p.Wo;
  p.Kper = cdiv(cdiv(p.K, splits), 64) * 64;              // a multiple of both kernels' chunk sizes
  for (int i = 0; i < p.T; ++i)
    for (int j = 0; j < 4; ++j) p.tap[i][j] = d->tap_off[i][j];
  if (d->nd == 2)
    for (int i = 0; i < p.T; ++i) p.tap[i][0] = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)pl.tiles, (unsigned)splits);
  if (pl.tpw) {
    int e = OFSV_ENOSUP;
#define OFSV_TG_CASE(M, Nn, TP) if (MT == M && NT == Nn && pl.tpw == TP) e = launch_wgrad_tg<M, Nn, TP>(p, grid, pl.ngroups, st); else
    OFSV_TG_CASE(64, 32, 1) OFSV_TG_CASE(64, 16, 2) OFSV_TG_CASE(32, 64, 1) OFSV_TG_CASE(32, 32, 2) OFSV_TG_CASE(32, 16, 4)
    OFSV_TG_CASE(16, 64, 2) OFSV_TG_CASE(16, 32, 4) OFSV_TG_CASE(16, 16, 4) { set_error("conv_wgrad: no tap-group kernel for %d x %d x %d", MT, NT, pl.tpw); }
#undef OFSV_TG_CASE
    if (e) return e;
  } else {
#define OFSV_WG_CASE(M, Nn) if (MT == M && NT == Nn) launch_wgrad<M, Nn>(p, grid, st); else
    OFSV_WG_CASE(64, 64) OFSV_WG_CASE(64, 32) OFSV_WG_CASE(64, 16) OFSV_WG_CASE(32, 64) OFSV_WG_CASE(32, 32) OFSV_WG_CASE(32, 16)
    OFSV_WG_CASE(16, 64) OFSV_WG_CASE(16, 32) OFSV_WG_CASE(16, 16) { set_error("conv_wgrad: no tile for %d x %d", MT, NT); return OFSV_ENOSUP; }
#undef OFSV_WG_CASE
  }
  rc = check_launch("conv_wgrad_kernel");
  if (rc) return rc;
  if (splits > 1) {
    const int64_t n4 = (int64_t)p.T * p.Cin_s * p.Cout_w / 4;
    conv_wgrad_finalize<<<(unsigned)cdiv(n4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(work), reinterpret_cast<float4*>(dw), n4, splits);
    rc = check_launch("conv_wgrad_finalize");
  }
  return rc;
}
