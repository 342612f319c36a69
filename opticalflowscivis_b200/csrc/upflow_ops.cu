// UPFlow L1 operators for sm_100a: 9x9 local correlation cost volume (a8) fwd/bwd, flow up-sampling (a10),
// feature warping with validity mask (a11).
#include "ofsv_common.cuh"

namespace ofsv {

// ----------------------------------------------------------------------------------------------------
// correlation-81 forward.
//   out[b,(dy+4)*9+(dx+4),y,x] = (1/C) sum_c f1[b,c,y,x] * f2[b,c,y+dy,x+dx]   (zero outside), optional LeakyReLU.
// CTA = 32(x) x 8(y) output pixels of one sample, 288 threads = 9 warps: warp = displacement row dy, lane = (pair of pixel
// rows, quad of 4 consecutive x).  Channel chunks of CC are staged in shared memory with their 4-pixel halo (16 B global
// loads, one fixed tile position per thread, channels walked by pointer increment); per channel a thread reads its 8 f1
// values and the 2 x 12 f2 values its 8 pixels x 9 dx need with eight 16 B shared loads and does 72 FMAs into registers
// (9 FMAs per shared load instead of 1), so the kernel is FMA-bound rather than shared-memory-bound.
// ----------------------------------------------------------------------------------------------------
constexpr int CTW = 32, CTH = 8, CMD = 4, CC = 8;
constexpr int CHW = CTW + 2 * CMD, CHH = CTH + 2 * CMD;

__device__ __forceinline__ uint32_t up_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// cp.async with zero fill: `bytes` valid source bytes (0 = write zeros; the source pointer must still be a valid address)
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Channel chunks are DOUBLE-BUFFERED with cp.async (the chunk after the one being multiplied is in flight; out-of-image elements
// are zero-filled by the copy itself), and the channels can be SPLIT over blockIdx.z: split s of nsplit accumulates channels
// [s*cps, (s+1)*cps) and — when nsplit > 1 — writes raw partial sums to work[s][b][81][H][W]; corr81_finalize_kernel adds the
// splits in a fixed order, divides by C and applies the LeakyReLU.  The coarse pyramid levels (4 x 13 ... 16 x 52 at B = 16) are
// 16-64 tiles: one CTA per tile walking 96-196 channels with a load -> barrier -> multiply -> barrier loop took 96-160 us per
// launch for a few hundred KB; with the split every SM gets a few chunks.
constexpr int CORR_FWD_SMEM = 2 * (CC * CHH * CHW + CC * CTH * CTW) * 4;      // 57344 B
template <bool VEC>
__global__ void __launch_bounds__(288, 2)
    corr81_fwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out, float* __restrict__ work,
                      int B, int C, int H, int W, int nsplit, int cps, float leaky_slope, int apply_leaky, int64_t out_batch_stride) {
  extern __shared__ __align__(16) float cf_smem[];
  float (*s2)[CC][CHH][CHW] = reinterpret_cast<float (*)[CC][CHH][CHW]>(cf_smem);                         // [2][CC][16][40]
  float (*s1)[CC][CTH][CTW] = reinterpret_cast<float (*)[CC][CTH][CTW]>(cf_smem + 2 * CC * CHH * CHW);    // [2][CC][8][32]
  const int b = blockIdx.z / nsplit, split = blockIdx.z - b * nsplit;
  const int x0 = blockIdx.x * CTW, y0 = blockIdx.y * CTH;
  const int c_beg = split * cps, c_end = min(C, c_beg + cps);
  const int tid = threadIdx.x;
  const int q = tid & 7, rp = (tid >> 3) & 3, dy = tid >> 5;       // quad, pixel-row pair, displacement row (warp-uniform)
  const int64_t HW = (int64_t)H * W;
  const float* p1 = f1 + (int64_t)b * C * HW;
  const float* p2 = f2 + (int64_t)b * C * HW;
  // loader role: threads 0..159 own one 16 B column of the f2 halo tile (16 rows x 10 quads), threads 160..223 one of the f1
  // tile (8 rows x 8 quads); the position (and its bounds check) is fixed, only the channel pointer moves
  const bool ld2 = tid < CHH * (CHW / 4), ld1 = tid >= 160 && tid < 160 + CTH * (CTW / 4);
  int lyy = 0, lxx = 0, lx_glob = 0;
  bool lrow = false;
  const float* lsrc = p1;
  if (ld2) { lyy = tid / (CHW / 4); lxx = (tid % (CHW / 4)) * 4; const int y = y0 + lyy - CMD; lx_glob = x0 + lxx - CMD;
             lrow = y >= 0 && y < H; lsrc = p2 + (int64_t)y * W + lx_glob; }
  if (ld1) { const int t = tid - 160; lyy = t / (CTW / 4); lxx = (t % (CTW / 4)) * 4; const int y = y0 + lyy; lx_glob = x0 + lxx;
             lrow = y < H; lsrc = p1 + (int64_t)y * W + lx_glob; }
  bool lok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) lok[i] = lrow && lx_glob + i >= 0 && lx_glob + i < W;
  const bool lok16 = lok[0] && lok[3];                              // VEC: W % 4 == 0 and x0 - 4 + 4k: a quad is all in or all out

  auto issue = [&](int c0, int buf) {                               // chunk [c0, c0 + CC) -> buffer buf
    if (ld1 || ld2) {
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        const bool cin = c0 + c < c_end;
        const float* g = lsrc + (int64_t)(c0 + c) * HW;
        const uint32_t dst = ld2 ? up_smem_u32(&s2[buf][c][lyy][lxx]) : up_smem_u32(&s1[buf][c][lyy][lxx]);
        if (VEC) {
          cp_async_16(dst, (cin && lok16) ? (const void*)g : (const void*)p1, (cin && lok16) ? 16 : 0);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) cp_async_4(dst + 4 * i, (cin && lok[i]) ? (const void*)(g + i) : (const void*)p1, (cin && lok[i]) ? 4 : 0);
        }
      }
    }
    cp_async_commit();
  };

  float acc[2][4][9];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[j][i][k] = 0.0f;

  issue(c_beg, 0);
  int buf = 0;
  for (int c0 = c_beg; c0 < c_end; c0 += CC, buf ^= 1) {
    if (c0 + CC < c_end) { issue(c0 + CC, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
#pragma unroll 2
    for (int c = 0; c < CC; ++c) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = 2 * rp + j;
        const float4 a4 = *reinterpret_cast<const float4*>(&s1[buf][c][r][4 * q]);
        const float4 r0 = *reinterpret_cast<const float4*>(&s2[buf][c][r + dy][4 * q]);
        const float4 r1 = *reinterpret_cast<const float4*>(&s2[buf][c][r + dy][4 * q + 4]);
        const float4 r2 = *reinterpret_cast<const float4*>(&s2[buf][c][r + dy][4 * q + 8]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float row[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int k = 0; k < 9; ++k) acc[j][i][k] = fmaf(a[i], row[i + k], acc[j][i][k]);
      }
    }
    __syncthreads();                                                // buffer `buf` is refilled by the next iteration's issue
  }
  const float cf = (float)C;
  const bool partial = nsplit > 1;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int x = x0 + 4 * q, y = y0 + 2 * rp + j;
    if (y < H && x < W) {
      float* o = partial ? work + ((int64_t)(split * B + b) * 81 + dy * 9) * HW + (int64_t)y * W + x
                         : out + (int64_t)b * out_batch_stride + (int64_t)(dy * 9) * HW + (int64_t)y * W + x;
      const bool vec = VEC && (x + 3 < W) && ((reinterpret_cast<uintptr_t>(o) & 15u) == 0) && ((HW & 3) == 0);
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = acc[j][i][k];
          if (!partial) {
            v[i] = v[i] / cf;  // torch.mean = sum / C
            if (apply_leaky && v[i] < 0.0f) v[i] *= leaky_slope;
          }
        }
        float* ok = o + (int64_t)k * HW;
        if (vec) {
          *reinterpret_cast<float4*>(ok) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (x + i < W) ok[i] = v[i];
        }
      }
    }
  }
}

// out[b][k][p] = leaky((sum_s work[s][b][k][p]) / C), splits added in ascending order
__global__ void corr81_finalize_kernel(const float* __restrict__ work, float* __restrict__ out, int B, int C, int64_t HW, int nsplit,
                                       float leaky_slope, int apply_leaky, int64_t out_batch_stride) {
  const int64_t per = 81 * HW, total = (int64_t)B * per;
  const float cf = (float)C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per, r = i - b * per;
    float v = __ldg(work + i);
    for (int s = 1; s < nsplit; ++s) v += __ldg(work + (int64_t)s * total + i);
    v = v / cf;
    if (apply_leaky && v < 0.0f) v *= leaky_slope;
    out[b * out_batch_stride + r] = v;
  }
}

// channel splits of the forward kernel for this shape: enough CTAs for two per SM, at least two chunks of channels per split
static int corr81_splits(int B, int C, int H, int W) {
  const int64_t ctas = cdiv(W, CTW) * cdiv(H, CTH) * (int64_t)B;
  const int64_t want = cdiv(2 * (int64_t)device_num_sms(), ctas);
  const int64_t cap = C / (2 * CC) > 1 ? C / (2 * CC) : 1;
  int64_t n = want < cap ? want : cap;
  if (n < 1) n = 1;
  if (ctas * n > 65535) n = 65535 / ctas > 1 ? 65535 / ctas : 1;
  return (int)n;
}

// backward (what autograd derives through Corr_pyTorch):
//   g1[b,c,y,x] = (1/C) sum_k gout[b,k,y,x]       * f2[b,c,y+dy,x+dx]
//   g2[b,c,y,x] = (1/C) sum_k gout[b,k,y-dy,x-dx] * f1[b,c,y-dy,x-dx]
// Both are the same stencil: with Gs[k][y][x] = gout[k][y-dy_k][x-dx_k] (the k-th plane pre-shifted while it is staged) and
// k' = 80-k, g2[c] = sum_k' Gs[80-k'][y][x] * f1[c][y+dy_k'][x+dx_k'].  CTA = 32x8 pixels of one sample: the 81 gradient planes
// of the tile stay in shared memory (83 KB) for all channels; channel chunks of the feature map are staged with their halo;
// thread = (quad of 4 x, row, channel slot) accumulates 4 pixels over the 81 displacements with one 16 B load per gradient
// plane and three per displacement row of the feature map (3 FMAs per shared load; 40x the naive gather kernel).
constexpr int BTW = 32, BTH = 8, BCC = 8;               // tile, channels per chunk
constexpr int BHW = BTW + 2 * CMD, BHH = BTH + 2 * CMD;
constexpr int CORR_BWD_SMEM = (81 * BTH * BTW + BCC * BHH * BHW) * 4;

template <bool G2>
__global__ void __launch_bounds__(256)
    corr81_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ gout, float* __restrict__ gin, int C, int H,
                      int W, int tiles_y, int c_per_cta) {
  extern __shared__ __align__(16) float bw_smem[];
  float (*sg)[BTH][BTW] = reinterpret_cast<float (*)[BTH][BTW]>(bw_smem);                      // [81][8][32]
  float (*sf)[BHH][BHW] = reinterpret_cast<float (*)[BHH][BHW]>(bw_smem + 81 * BTH * BTW);     // [BCC][16][40]
  // blockIdx.y = channel split * tiles_y + tile row: the channels are independent outputs, so small images are spread over the
  // machine by giving every CTA only a slice of them
  const int b = blockIdx.z, x0 = blockIdx.x * BTW, y0 = (blockIdx.y % tiles_y) * BTH;
  const int c_beg = (blockIdx.y / tiles_y) * c_per_cta, c_end = min(C, c_beg + c_per_cta);
  const int tid = threadIdx.x;
  const int q = tid & 7, cs = (tid >> 3) & 3, r = tid >> 5;          // quad, channel slot (0..3), row: a warp = one row
  const int64_t HW = (int64_t)H * W;
  const float* pf = feat + (int64_t)b * C * HW;
  const float* pg = gout + (int64_t)b * 81 * HW;
  float* po = gin + (int64_t)b * C * HW;
  // gradient planes of the tile, plane k shifted by (-dy_k, -dx_k) for g2
  for (int i = tid; i < 81 * BTH * BTW; i += 256) {
    const int k = i / (BTH * BTW), rr = i - k * (BTH * BTW), yy = rr / BTW, xx = rr - yy * BTW;
    const int dyk = k / 9 - CMD, dxk = k % 9 - CMD;
    const int y = y0 + yy - (G2 ? dyk : 0), x = x0 + xx - (G2 ? dxk : 0);
    float v = 0.0f;
    if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(pg + (int64_t)k * HW + (int64_t)y * W + x);
    sg[k][yy][xx] = v;
  }
  const float cf = (float)C;
  for (int c0 = c_beg; c0 < c_end; c0 += BCC) {
    const int cc = min(BCC, c_end - c0);
    __syncthreads();
    for (int i = tid; i < BCC * BHH * BHW; i += 256) {
      const int c = i / (BHH * BHW), rr = i - c * (BHH * BHW), yy = rr / BHW, xx = rr - yy * BHW;
      const int y = y0 + yy - CMD, x = x0 + xx - CMD;
      float v = 0.0f;
      if (c < cc && y >= 0 && y < H && x >= 0 && x < W) v = __ldg(pf + (int64_t)(c0 + c) * HW + (int64_t)y * W + x);
      sf[c][yy][xx] = v;
    }
    __syncthreads();
    // per displacement row: the 9 gradient quads are loaded ONCE (a warp = 8 quads x 4 channel slots of one row: the four slots
    // read the same 128 B, one shared-memory wavefront) and feed both channels of the slot
    float acc[BCC / 4][4];
#pragma unroll
    for (int ci = 0; ci < BCC / 4; ++ci) { acc[ci][0] = 0.f; acc[ci][1] = 0.f; acc[ci][2] = 0.f; acc[ci][3] = 0.f; }
#pragma unroll 3
    for (int dy = 0; dy < 9; ++dy) {
      float4 g4[9];
#pragma unroll
      for (int dx = 0; dx < 9; ++dx) {
        const int k = G2 ? 80 - (dy * 9 + dx) : dy * 9 + dx;
        g4[dx] = *reinterpret_cast<const float4*>(&sg[k][r][4 * q]);
      }
#pragma unroll
      for (int ci = 0; ci < BCC / 4; ++ci) {
        const int c = cs + 4 * ci;
        const float4 r0 = *reinterpret_cast<const float4*>(&sf[c][r + dy][4 * q]);
        const float4 r1 = *reinterpret_cast<const float4*>(&sf[c][r + dy][4 * q + 4]);
        const float4 r2 = *reinterpret_cast<const float4*>(&sf[c][r + dy][4 * q + 8]);
        const float row[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
#pragma unroll
        for (int dx = 0; dx < 9; ++dx) {
          acc[ci][0] = fmaf(g4[dx].x, row[dx], acc[ci][0]); acc[ci][1] = fmaf(g4[dx].y, row[dx + 1], acc[ci][1]);
          acc[ci][2] = fmaf(g4[dx].z, row[dx + 2], acc[ci][2]); acc[ci][3] = fmaf(g4[dx].w, row[dx + 3], acc[ci][3]);
        }
      }
    }
#pragma unroll
    for (int ci = 0; ci < BCC / 4; ++ci) {
      const int c = cs + 4 * ci;
      const int x = x0 + 4 * q, y = y0 + r;
      if (c < cc && y < H) {
        float* o = po + (int64_t)(c0 + c) * HW + (int64_t)y * W + x;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (x + i < W) o[i] = acc[ci][i] / cf;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------------
// upsample2d_flow_as: bilinear, align_corners=True; u *= w/w_, v *= h/h_ (pwc_modules.py:77-90).
// ATen: ratio = (in-1)/(out-1) in fp32; src = ratio*dst; i0 = (int)src; lambda1 = src - i0.
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    upsample_flow_ac_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int h_, int w_, int h, int w,
                            float ry, float rx, float us, float vs, int if_rate) {
  const int64_t total = (int64_t)B * 2 * h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const int64_t bc = i / ((int64_t)h * w);
    const int c = (int)(bc & 1);
    const float* p = in + bc * h_ * w_;
    const float sy = __fmul_rn(ry, (float)y), sx = __fmul_rn(rx, (float)x);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < h_ - 1 ? 1 : 0), x1 = x0 + (x0 < w_ - 1 ? 1 : 0);
    const float ly1 = __fsub_rn(sy, (float)y0), ly0 = __fsub_rn(1.0f, ly1);
    const float lx1 = __fsub_rn(sx, (float)x0), lx0 = __fsub_rn(1.0f, lx1);
    const float top = __fadd_rn(__fmul_rn(lx0, __ldg(p + y0 * w_ + x0)), __fmul_rn(lx1, __ldg(p + y0 * w_ + x1)));
    const float bot = __fadd_rn(__fmul_rn(lx0, __ldg(p + y1 * w_ + x0)), __fmul_rn(lx1, __ldg(p + y1 * w_ + x1)));
    float v = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
    if (if_rate) v = __fmul_rn(v, c == 0 ? us : vs);
    out[i] = v;
  }
}

// ----------------------------------------------------------------------------------------------------
// WarpingLayer_no_div (pwc_modules.py:184-207): zeros padding, align_corners=False, validity mask (sum of in-bounds
// corner weights >= 1).  The >= 1 test is decided by the last bit of fp32 arithmetic, so the op order (including the
// one contracted FMA in ATen's unnormalize) is replicated exactly — see oracle/ofsv_oracle.c.
// ----------------------------------------------------------------------------------------------------
// Sampling position, corner weights, in-bounds flags and the >= 1 validity of one output pixel (shared by the kernels below).
struct NoDivTaps {
  float nw, ne, sw, se, valid;
  int x0, y0;
  bool in00, in01, in10, in11;
};
__device__ __forceinline__ NoDivTaps nodiv_taps(int x, int y, float fx, float fy, int H, int W, float dw, float dh, float rdw, float rdh,
                                                int ref_mode) {
  NoDivTaps t;
  const float vx = __fadd_rn((float)x, fx), vy = __fadd_rn((float)y, fy);
  float gx, gy;
  if (ref_mode == OFSV_REF_CUDA) {
    gx = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, vx), rdw), 1.0f);
    gy = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, vy), rdh), 1.0f);
  } else {
    gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, vx), dw), 1.0f);
    gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, vy), dh), 1.0f);
  }
  const float ix = __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.0f), (float)W, -1.0f), 0.5f);
  const float iy = __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.0f), (float)H, -1.0f), 0.5f);
  const float xw = floorf(ix), yn = floorf(iy);
  const float w = __fsub_rn(ix, xw), e = __fsub_rn(1.0f, w), n = __fsub_rn(iy, yn), s = __fsub_rn(1.0f, n);
  t.nw = __fmul_rn(s, e); t.ne = __fmul_rn(s, w); t.sw = __fmul_rn(n, e); t.se = __fmul_rn(n, w);
  const float xc = fminf(fmaxf(xw, -2.0f), (float)W + 1.0f), yc = fminf(fmaxf(yn, -2.0f), (float)H + 1.0f);
  t.x0 = (int)xc; t.y0 = (int)yc;
  const int x1 = t.x0 + 1, y1 = t.y0 + 1;
  const bool ix0 = t.x0 >= 0 && t.x0 < W, ix1 = x1 >= 0 && x1 < W, iy0 = t.y0 >= 0 && t.y0 < H, iy1 = y1 >= 0 && y1 < H;
  t.in00 = ix0 && iy0; t.in01 = ix1 && iy0; t.in10 = ix0 && iy1; t.in11 = ix1 && iy1;
  const float msum = __fadd_rn(__fadd_rn(__fadd_rn(t.in00 ? t.nw : 0.0f, t.in01 ? t.ne : 0.0f), t.in10 ? t.sw : 0.0f), t.in11 ? t.se : 0.0f);
  t.valid = msum >= 1.0f ? 1.0f : 0.0f;
  return t;
}
__device__ __forceinline__ float nodiv_sample(const NoDivTaps& t, const float* __restrict__ p, int W) {
  const float p00 = t.in00 ? __ldg(p + (int64_t)t.y0 * W + t.x0) : 0.0f, p01 = t.in01 ? __ldg(p + (int64_t)t.y0 * W + t.x0 + 1) : 0.0f;
  const float p10 = t.in10 ? __ldg(p + (int64_t)(t.y0 + 1) * W + t.x0) : 0.0f, p11 = t.in11 ? __ldg(p + (int64_t)(t.y0 + 1) * W + t.x0 + 1) : 0.0f;
  return __fmaf_rn(p11, t.se, __fmaf_rn(p10, t.sw, __fmaf_rn(p01, t.ne, __fmul_rn(p00, t.nw))));
}

// apply_mask = 1: WarpingLayer_no_div; 0: tools.torch_warp (UPFlow/utils/tools.py:1317-1361), the same sampling without the mask
__global__ void __launch_bounds__(256)
    warping_no_div_kernel(const float* __restrict__ src, const float* __restrict__ flow, float* __restrict__ out, int B,
                          int C, int H, int W, float dw, float dh, float rdw, float rdh, int ref_mode, int c_per, int apply_mask) {
  // blockIdx.y = channel chunk: the coarse pyramid levels have few pixels and many channels (4 x 13 x 196), one thread per
  // pixel looping over all channels left the GPU empty (43 us for 160 KB)
  const int c_begin = blockIdx.y * c_per, c_end = min(C, c_begin + c_per);
  const int64_t HW = (int64_t)H * W, total = (int64_t)B * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int r = (int)(i - (int64_t)b * HW);
    const int y = r / W, x = r - y * W;
    const NoDivTaps t = nodiv_taps(x, y, ldg_stream(flow + ((int64_t)b * 2 + 0) * HW + r), ldg_stream(flow + ((int64_t)b * 2 + 1) * HW + r), H, W,
                                   dw, dh, rdw, rdh, ref_mode);
    const float valid = apply_mask ? t.valid : 1.0f;
    for (int c = c_begin; c < c_end; ++c) {
      const float v = nodiv_sample(t, src + ((int64_t)b * C + c) * HW, W);
      out[((int64_t)b * C + c) * HW + r] = apply_mask ? __fmul_rn(v, valid) : v;
    }
  }
}

// ----------------------------------------------------------------------------------------------------
// Producer of the cost-volume inputs of one pyramid level (UPFlow/model/upflow.py:621-640): feature_2_warp =
// WarpingLayer_no_div(feature_2, flow) followed by network_tools.normalize_features((feature_1, feature_2_warp), normalize = center
// = True, moments_across_channels = moments_across_images = False) (upflow.py:95-138): per (sample, channel) plane
//   out = (f - mean(f)) / sqrt(var_unbiased(f) + 1e-16).
// One CTA per (tensor, sample, channel): the plane (warped on the fly for the second tensor) is staged in shared memory, so the
// source is read once and the three passes (mean, centred sum of squares, normalise) never touch HBM again.
// ----------------------------------------------------------------------------------------------------
constexpr int FN_THREADS = 256;
__device__ __forceinline__ float fn_block_sum(float v, float* red) {
  // fixed-order reduction: lanes by shuffle, warps by one thread
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();                                   // `red` may still be read from the previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < FN_THREADS / 32; ++w) s += red[w];
  return s;
}
__global__ void __launch_bounds__(FN_THREADS)
    feature_norm_pair_kernel(const float* __restrict__ f_plain, const float* __restrict__ f_src, const float* __restrict__ flow,
                             float* __restrict__ out_plain, float* __restrict__ out_warp, int C, int H, int W, float dw, float dh,
                             float rdw, float rdh, int ref_mode) {
  extern __shared__ float plane[];
  __shared__ float red[FN_THREADS / 32];
  const int c = blockIdx.x, b = blockIdx.y, which = blockIdx.z;
  const int HW = H * W;
  const float* src = (which == 0 ? f_plain : f_src) + ((int64_t)b * C + c) * HW;
  float* out = (which == 0 ? out_plain : out_warp) + ((int64_t)b * C + c) * HW;
  const bool warp = which == 1 && flow != nullptr;
  float sum = 0.f;
  for (int r = threadIdx.x; r < HW; r += FN_THREADS) {
    float v;
    if (warp) {
      const int y = r / W, x = r - y * W;
      const NoDivTaps t = nodiv_taps(x, y, __ldg(flow + ((int64_t)b * 2 + 0) * HW + r), __ldg(flow + ((int64_t)b * 2 + 1) * HW + r), H, W, dw, dh,
                                     rdw, rdh, ref_mode);
      v = __fmul_rn(nodiv_sample(t, src, W), t.valid);
    } else {
      v = __ldg(src + r);
    }
    plane[r] = v;
    sum += v;
  }
  const float mean = fn_block_sum(sum, red) / (float)HW;
  float ss = 0.f;
  for (int r = threadIdx.x; r < HW; r += FN_THREADS) {
    const float d = plane[r] - mean;
    ss += d * d;
  }
  const float var = fn_block_sum(ss, red) / (float)(HW - 1);     // torch.var: unbiased
  const float sd = sqrtf(var + 1e-16f);
  for (int r = threadIdx.x; r < HW; r += FN_THREADS) out[r] = __fdiv_rn(__fsub_rn(plane[r], mean), sd);
}

static inline int grid_1d(int64_t total) {
  int64_t b = cdiv(total, 256);
  const int64_t cap = device_num_sms() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace ofsv

using namespace ofsv;

extern "C" int ofsv_corr81_fwd_splits(int B, int C, int H, int W) {
  if (B < 1 || C < 1 || H < 1 || W < 1) return 1;
  return corr81_splits(B, C, H, W);
}

extern "C" int ofsv_corr81_fwd_f32(const float* f1, const float* f2, float* out, int B, int C, int H, int W,
                                   float leaky_slope, int apply_leaky, int64_t out_batch_stride, float* work, void* stream) {
  OFSV_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1, "ofsv_corr81_fwd_f32: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  if (B == 0) return OFSV_OK;
  OFSV_REQUIRE(f1 && f2 && out, "ofsv_corr81_fwd_f32: null pointer");
  OFSV_REQUIRE(out_batch_stride >= (int64_t)81 * H * W, "ofsv_corr81_fwd_f32: out_batch_stride %lld < 81*H*W",
               (long long)out_batch_stride);
  OFSV_REQUIRE(cdiv(H, CTH) <= 65535, "ofsv_corr81_fwd_f32: H too large");
  // without a workspace the channels are not split (one CTA per tile walks all of them)
  const int nsplit = work ? corr81_splits(B, C, H, W) : 1;
  OFSV_REQUIRE((int64_t)B * nsplit <= 65535, "ofsv_corr81_fwd_f32: batch %d exceeds grid.z", B);
  const int cps = (int)(cdiv(cdiv(C, nsplit), CC) * CC);             // channels per split, whole chunks
  const dim3 grid((unsigned)cdiv(W, CTW), (unsigned)cdiv(H, CTH), (unsigned)(B * nsplit));
  const bool vec = (W % 4 == 0) && aligned16(f1) && aligned16(f2);
  static std::atomic<uint64_t> attr_a{0}, attr_b{0};
  if (int e = ensure_dyn_smem(attr_a, corr81_fwd_kernel<true>, CORR_FWD_SMEM, "ofsv_corr81_fwd_f32")) return e;
  if (int e = ensure_dyn_smem(attr_b, corr81_fwd_kernel<false>, CORR_FWD_SMEM, "ofsv_corr81_fwd_f32")) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) corr81_fwd_kernel<true><<<grid, 288, CORR_FWD_SMEM, st>>>(f1, f2, out, work, B, C, H, W, nsplit, cps, leaky_slope, apply_leaky, out_batch_stride);
  else corr81_fwd_kernel<false><<<grid, 288, CORR_FWD_SMEM, st>>>(f1, f2, out, work, B, C, H, W, nsplit, cps, leaky_slope, apply_leaky, out_batch_stride);
  if (int e = check_launch("corr81_fwd_kernel")) return e;
  if (nsplit > 1) {
    const int64_t total = (int64_t)B * 81 * H * W;
    const int fgrid = (int)(cdiv(total, 256) < 1184 ? cdiv(total, 256) : 1184);
    corr81_finalize_kernel<<<fgrid, 256, 0, st>>>(work, out, B, C, (int64_t)H * W, nsplit, leaky_slope, apply_leaky, out_batch_stride);
    return check_launch("corr81_finalize_kernel");
  }
  return OFSV_OK;
}

extern "C" int ofsv_corr81_bwd_f32(const float* f1, const float* f2, const float* gout, float* g1, float* g2, int B,
                                   int C, int H, int W, void* stream) {
  OFSV_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1, "ofsv_corr81_bwd_f32: bad shape");
  if (B == 0) return OFSV_OK;
  OFSV_REQUIRE(f1 && f2 && gout && g1 && g2, "ofsv_corr81_bwd_f32: null pointer");
  OFSV_REQUIRE(B <= 65535 && cdiv(H, BTH) <= 65535, "ofsv_corr81_bwd_f32: batch / height exceed the grid");
  static std::atomic<uint64_t> attr_a{0}, attr_b{0};
  if (int e = ensure_dyn_smem(attr_a, corr81_bwd_kernel<false>, CORR_BWD_SMEM, "ofsv_corr81_bwd_f32")) return e;
  if (int e = ensure_dyn_smem(attr_b, corr81_bwd_kernel<true>, CORR_BWD_SMEM, "ofsv_corr81_bwd_f32")) return e;
  const int tiles_x = (int)cdiv(W, BTW), tiles_y = (int)cdiv(H, BTH);
  int nsplit = (int)cdiv(2 * device_num_sms(), (int64_t)tiles_x * tiles_y * B);          // aim at two CTAs per SM
  const int max_split = (int)cdiv(C, BCC);
  nsplit = nsplit < 1 ? 1 : (nsplit > max_split ? max_split : nsplit);
  const int c_per_cta = (int)cdiv(cdiv(C, nsplit), BCC) * BCC;
  nsplit = (int)cdiv(C, c_per_cta);
  const dim3 grid((unsigned)tiles_x, (unsigned)(tiles_y * nsplit), (unsigned)B);
  corr81_bwd_kernel<false><<<grid, 256, CORR_BWD_SMEM, (cudaStream_t)stream>>>(f2, gout, g1, C, H, W, tiles_y, c_per_cta);
  if (int e = check_launch("corr81_bwd_kernel<g1>")) return e;
  corr81_bwd_kernel<true><<<grid, 256, CORR_BWD_SMEM, (cudaStream_t)stream>>>(f1, gout, g2, C, H, W, tiles_y, c_per_cta);
  return check_launch("corr81_bwd_kernel<g2>");
}

extern "C" int ofsv_upsample_flow_ac_f32(const float* in, float* out, int B, int h_in, int w_in, int h_out, int w_out,
                                         int if_rate, void* stream) {
  OFSV_REQUIRE(B >= 0 && h_in >= 1 && w_in >= 1 && h_out >= 1 && w_out >= 1, "ofsv_upsample_flow_ac_f32: bad shape");
  if (B == 0) return OFSV_OK;
  OFSV_REQUIRE(in && out, "ofsv_upsample_flow_ac_f32: null pointer");
  const float ry = h_out > 1 ? (float)(h_in - 1) / (float)(h_out - 1) : 0.0f;
  const float rx = w_out > 1 ? (float)(w_in - 1) / (float)(w_out - 1) : 0.0f;
  const float us = (float)((double)w_out / (double)w_in), vs = (float)((double)h_out / (double)h_in);
  upsample_flow_ac_kernel<<<grid_1d((int64_t)B * 2 * h_out * w_out), 256, 0, (cudaStream_t)stream>>>(
      in, out, B, h_in, w_in, h_out, w_out, ry, rx, us, vs, if_rate);
  return check_launch("upsample_flow_ac_kernel");
}

static int warping_no_div_launch(const char* who, const float* src, const float* flow, float* out, int B, int C, int H, int W, int ref_mode,
                                 int apply_mask, void* stream) {
  OFSV_REQUIRE(B >= 0 && C >= 0 && H >= 1 && W >= 1 && (int64_t)H * W < (1ll << 31), "%s: bad shape", who);
  OFSV_REQUIRE(ref_mode == OFSV_REF_CPU || ref_mode == OFSV_REF_CUDA, "%s: bad ref_mode", who);
  if ((int64_t)B * C == 0) return OFSV_OK;
  OFSV_REQUIRE(src && flow && out, "%s: null pointer", who);
  const int dw = W - 1 > 1 ? W - 1 : 1, dh = H - 1 > 1 ? H - 1 : 1;
  const int gx = grid_1d((int64_t)B * H * W);
  int nsplit = (int)(cdiv(device_num_sms() * 8, gx));                 // aim at >= 8 CTAs per SM; at least 4 channels per thread
  nsplit = nsplit < 1 ? 1 : (nsplit > cdiv(C, 4) ? (int)cdiv(C, 4) : nsplit);
  const int c_per = (int)cdiv(C, nsplit);
  warping_no_div_kernel<<<dim3((unsigned)gx, (unsigned)cdiv(C, c_per)), 256, 0, (cudaStream_t)stream>>>(
      src, flow, out, B, C, H, W, (float)dw, (float)dh, (float)(1.0 / (double)dw), (float)(1.0 / (double)dh), ref_mode, c_per, apply_mask);
  return check_launch("warping_no_div_kernel");
}

extern "C" int ofsv_warping_no_div_f32(const float* src, const float* flow, float* out, int B, int C, int H, int W,
                                       int ref_mode, void* stream) {
  return warping_no_div_launch("ofsv_warping_no_div_f32", src, flow, out, B, C, H, W, ref_mode, 1, stream);
}

extern "C" int ofsv_torch_warp_f32(const float* src, const float* flow, float* out, int B, int C, int H, int W, int ref_mode,
                                   void* stream) {
  return warping_no_div_launch("ofsv_torch_warp_f32", src, flow, out, B, C, H, W, ref_mode, 0, stream);
}

extern "C" int ofsv_feature_norm_pair_f32(const float* f_plain, const float* f_src, const float* flow, float* out_plain, float* out_warp,
                                          int B, int C, int H, int W, int ref_mode, void* stream) {
  OFSV_REQUIRE(B >= 0 && C >= 0 && H >= 1 && W >= 1 && (int64_t)H * W >= 2, "ofsv_feature_norm_pair_f32: bad shape (a plane needs >= 2 pixels)");
  OFSV_REQUIRE(ref_mode == OFSV_REF_CPU || ref_mode == OFSV_REF_CUDA, "ofsv_feature_norm_pair_f32: bad ref_mode");
  if ((int64_t)B * C == 0) return OFSV_OK;
  OFSV_REQUIRE(f_plain && f_src && out_plain && out_warp, "ofsv_feature_norm_pair_f32: null pointer");
  OFSV_REQUIRE(B <= 65535, "ofsv_feature_norm_pair_f32: B > 65535");
  const int64_t bytes = (int64_t)H * W * 4;
  if (bytes > 200 * 1024) {
    set_error("ofsv_feature_norm_pair_f32: a %d x %d plane does not fit the 200 KB shared-memory staging (pyramid levels of UPFlow are <= 64 x 208)", H, W);
    return OFSV_ENOSUP;
  }
  static std::atomic<uint64_t> attr_done{0};
  if (int e = ensure_dyn_smem(attr_done, feature_norm_pair_kernel, 200 * 1024, "ofsv_feature_norm_pair_f32")) return e;
  const int dw = W - 1 > 1 ? W - 1 : 1, dh = H - 1 > 1 ? H - 1 : 1;
  feature_norm_pair_kernel<<<dim3((unsigned)C, (unsigned)B, 2), FN_THREADS, (size_t)bytes, (cudaStream_t)stream>>>(
      f_plain, f_src, flow, out_plain, out_warp, C, H, W, (float)dw, (float)dh, (float)(1.0 / (double)dw), (float)(1.0 / (double)dh), ref_mode);
  return check_launch("feature_norm_pair_kernel");
}
