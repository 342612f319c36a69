// Backward of the warp layer (a1, a2): what autograd runs under `warp(tenInput, tenFlow)` in the reference's training
// step (Flow-2D/model/RIFE.py:80-336, Flow-3D/model/RIFE.py:81-275 through Flow-*/model/warplayer.py), i.e. ATen's
// grid_sampler_{2,3}d_backward (bilinear, border, align_corners=True) followed by the `flow / ((S-1)/2)` division.
//
//   gsrc[n,c,tap]  += weight(tap) * gout[n,c,voxel]                        (scatter; red.global.add.f32)
//   gflow[n,a,vox]  = (sum_c gout * d value / d i_a) * clipgrad_a * ((S_a-1)/2) / half_extent_a
//
// clipgrad_a is ATen's clip_coordinates_set_grad: 0 where the un-normalised coordinate is <= 0 or >= S-1, else 1.  The
// coordinates themselves are computed by the SAME helpers as the forward kernels (bit-identical cell and weights); the
// gradient sums are not order-identical to ATen (and the scatter is atomic), parity is to 1e-5 relative.
//
// 3-D keeps the forward kernel's mapping, lanes along h = the contiguous SOURCE axis of the rotated warp, so the 8 gathers
// and the 8 reductions of a warp each fall into one or two 128 B lines.
#include "ofsv_common.cuh"
#include "warp_device.cuh"

namespace ofsv {

__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float div_back(float g, float half_extent, float rcp_half_extent, int ref_mode) {
  return ref_mode == OFSV_REF_CUDA ? __fmul_rn(g, rcp_half_extent) : __fdiv_rn(g, half_extent);
}

// ----------------------------------------------------------------------------------------------------
// 2-D: one thread per (n, y, x), loop over channels
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    warp2d_bwd_kernel(const float* __restrict__ src, const float* __restrict__ flow, const float* __restrict__ gout,
                      const float* __restrict__ lin_x, const float* __restrict__ lin_y, float* __restrict__ gsrc,
                      float* __restrict__ gflow, int N, int C, int H, int W, int ref_mode) {
  const int64_t HW = (int64_t)H * W;
  const int64_t total = (int64_t)N * HW;
  const float hx = (float)((W - 1.0) / 2.0), hy = (float)((H - 1.0) / 2.0);
  const float rhx = (float)(1.0 / ((W - 1.0) / 2.0)), rhy = (float)(1.0 / ((H - 1.0) / 2.0));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW);
    const int r = (int)(i - (int64_t)n * HW);
    const int y = r / W, x = r - y * W;
    const float fx = ldg_stream(flow + ((int64_t)n * 2 + 0) * HW + r);
    const float fy = ldg_stream(flow + ((int64_t)n * 2 + 1) * HW + r);
    const float gx = __fadd_rn(__ldg(lin_x + x), norm_flow(fx, hx, rhx, ref_mode));
    const float gy = __fadd_rn(__ldg(lin_y + y), norm_flow(fy, hy, rhy, ref_mode));
    const float ux = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), (float)(W - 1));
    const float uy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), (float)(H - 1));
    const float cgx = (ux > 0.0f && ux < (float)(W - 1)) ? 1.0f : 0.0f;     // clip_coordinates_set_grad
    const float cgy = (uy > 0.0f && uy < (float)(H - 1)) ? 1.0f : 0.0f;
    const float ix = fminf((float)(W - 1), fmaxf(ux, 0.0f)), iy = fminf((float)(H - 1), fmaxf(uy, 0.0f));
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float wx = ix - x0f, ex = 1.0f - wx, wy = iy - y0f, ey = 1.0f - wy;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int dx = x0 + 1 <= W - 1 ? 1 : 0, dy = y0 + 1 <= H - 1 ? W : 0;   // out-of-range +1 taps carry weight 0
    const int b = y0 * W + x0;
    float gix = 0.0f, giy = 0.0f;
    for (int c = 0; c < C; ++c) {
      const int64_t pl = ((int64_t)n * C + c) * HW;
      const float go = ldg_stream(gout + pl + r);
      if (gsrc) {
        float* g = gsrc + pl + b;
        red_add(g, ex * ey * go);
        if (dx) red_add(g + 1, wx * ey * go);
        if (dy) red_add(g + dy, ex * wy * go);
        if (dx && dy) red_add(g + dy + 1, wx * wy * go);
      }
      if (gflow) {
        const float* p = src + pl + b;
        const float v00 = __ldg(p), v01 = __ldg(p + dx), v10 = __ldg(p + dy), v11 = __ldg(p + dy + dx);
        gix += ((v01 - v00) * ey + (v11 - v10) * wy) * go;
        giy += ((v10 - v00) * ex + (v11 - v01) * wx) * go;
      }
    }
    if (gflow) {
      gflow[((int64_t)n * 2 + 0) * HW + r] = div_back(gix * (cgx * hx), hx, rhx, ref_mode);
      gflow[((int64_t)n * 2 + 1) * HW + r] = div_back(giy * (cgy * hy), hy, rhy, ref_mode);
    }
  }
}

// ----------------------------------------------------------------------------------------------------
// 3-D: CTA = 32(h) x 8(w) voxels at fixed (n, d); flow / gout / gflow tiles cross shared memory so that global traffic is
// coalesced along w while the gathers and reductions run with lanes along h (see warp.cu)
// ----------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
    warp3d_bwd_kernel(const float* __restrict__ src, const float* __restrict__ flow, const float* __restrict__ gout,
                      const float* __restrict__ lin_h, const float* __restrict__ lin_d, const float* __restrict__ lin_w,
                      float* __restrict__ gsrc, float* __restrict__ gflow, const Warp3dParams P) {
  __shared__ float si[3][T3H][T3P];   // flow 0..2
  __shared__ float sg[1][T3H][T3P];   // gout of the current channel
  __shared__ float so[3][T3H][T3P];   // gflow 0..2
  const int H = P.H, W = P.W, D = P.D, HW = H * W;
  const int64_t V = (int64_t)D * HW;
  const int n = blockIdx.z / D, d = blockIdx.z - n * D;
  const int h0 = blockIdx.y * T3H, w0 = blockIdx.x * T3W;
  const float* fl = flow + (int64_t)n * 3 * V + (int64_t)d * HW;
  load_planes<3, VEC>(si, [&](int k) -> const float* { return fl + (int64_t)k * V; }, h0, w0, H, W);
  __syncthreads();
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int h = h0 + lane, w = w0 + wl;
  const bool ok = h < H && w < W;
  Trilin t{};
  float cg0 = 0.f, cg1 = 0.f, cg2 = 0.f;
  if (ok) {
    const float lh = __ldg(lin_h + h), ld = __ldg(lin_d + d), lw = __ldg(lin_w + w);
    const float f0 = si[0][lane][wl], f1 = si[1][lane][wl], f2 = si[2][lane][wl];
    t = trilin_setup(f0, f1, f2, lh, ld, lw, D, H, W, P.hs, P.ref_mode);
    // un-clipped coordinates again for clip_coordinates_set_grad (same expressions as trilin_setup / unnorm_clip_ac)
    const float u0 = __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(lh, norm_flow(f0, P.hs[0], P.hs[3], P.ref_mode)), 1.0f), 0.5f), (float)(W - 1));
    const float u1 = __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(ld, norm_flow(f1, P.hs[1], P.hs[4], P.ref_mode)), 1.0f), 0.5f), (float)(H - 1));
    const float u2 = __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(lw, norm_flow(f2, P.hs[2], P.hs[5], P.ref_mode)), 1.0f), 0.5f), (float)(D - 1));
    cg0 = (u0 > 0.0f && u0 < (float)(W - 1)) ? 1.0f : 0.0f;
    cg1 = (u1 > 0.0f && u1 < (float)(H - 1)) ? 1.0f : 0.0f;
    cg2 = (u2 > 0.0f && u2 < (float)(D - 1)) ? 1.0f : 0.0f;
  }
  float gix = 0.f, giy = 0.f, giz = 0.f;
  for (int c = 0; c < P.C; ++c) {
    const int64_t vol = ((int64_t)n * P.C + c) * V;
    __syncthreads();                       // previous channel's sg fully consumed
    load_planes<1, VEC>(sg, [&](int) -> const float* { return gout + vol + (int64_t)d * HW; }, h0, w0, H, W);
    __syncthreads();
    if (ok) {
      const float go = sg[0][lane][wl];
      const float xy00 = t.ex * t.ey, xy10 = t.wx * t.ey, xy01 = t.ex * t.wy, xy11 = t.wx * t.wy;
      if (gflow) {
        const Taps8 r = trilin_gather(src + vol, t);
        // d/dx: (+1 tap) - (0 tap) weighted by the other two axes, etc.
        const float dxv = ((r.v[1] - r.v[0]) * t.ey + (r.v[3] - r.v[2]) * t.wy) * t.ez +
                          ((r.v[5] - r.v[4]) * t.ey + (r.v[7] - r.v[6]) * t.wy) * t.wz;
        const float dyv = ((r.v[2] - r.v[0]) * t.ex + (r.v[3] - r.v[1]) * t.wx) * t.ez +
                          ((r.v[6] - r.v[4]) * t.ex + (r.v[7] - r.v[5]) * t.wx) * t.wz;
        const float dzv = (r.v[4] - r.v[0]) * xy00 + (r.v[5] - r.v[1]) * xy10 + (r.v[6] - r.v[2]) * xy01 + (r.v[7] - r.v[3]) * xy11;
        gix += dxv * go; giy += dyv * go; giz += dzv * go;
      }
      if (gsrc) {
        float* g = gsrc + vol + t.base;
        // an out-of-range +1 neighbour has offset 0 and weight exactly 0: adding 0 to the in-range sample is harmless, but
        // skip the traffic
        const bool px = t.dx != 0, py = t.dy != 0, pz = t.dz != 0;
        red_add(g, xy00 * t.ez * go);
        if (px) red_add(g + 1, xy10 * t.ez * go);
        if (py) red_add(g + t.dy, xy01 * t.ez * go);
        if (px && py) red_add(g + t.dy + 1, xy11 * t.ez * go);
        if (pz) {
          g += t.dz;
          red_add(g, xy00 * t.wz * go);
          if (px) red_add(g + 1, xy10 * t.wz * go);
          if (py) red_add(g + t.dy, xy01 * t.wz * go);
          if (px && py) red_add(g + t.dy + 1, xy11 * t.wz * go);
        }
      }
    }
  }
  if (gflow) {
    if (ok) {
      // grad_grid = g_i * clipgrad * (S_sampled - 1)/2 ; grad_flow = grad_grid / half_extent (Flow-3D/model/warplayer.py:24-26:
      // channel 0 is normalised with (H-1)/2 but sampled along W, 1 with (D-1)/2 along H, 2 with (W-1)/2 along D)
      so[0][lane][wl] = div_back(gix * (cg0 * (float)((W - 1.0) / 2.0)), P.hs[0], P.hs[3], P.ref_mode);
      so[1][lane][wl] = div_back(giy * (cg1 * (float)((H - 1.0) / 2.0)), P.hs[1], P.hs[4], P.ref_mode);
      so[2][lane][wl] = div_back(giz * (cg2 * (float)((D - 1.0) / 2.0)), P.hs[2], P.hs[5], P.ref_mode);
    }
    __syncthreads();
    float* gf = gflow + (int64_t)n * 3 * V + (int64_t)d * HW;
    store_planes<3, VEC>(so, [&](int k) -> float* { return gf + (int64_t)k * V; }, h0, w0, H, W);
  }
}

}  // namespace ofsv

using namespace ofsv;

extern "C" int ofsv_warp2d_bwd_f32(const float* src, const float* flow, const float* gout, const float* lin_x,
                                   const float* lin_y, float* gsrc, float* gflow, int N, int C, int H, int W, int ref_mode,
                                   void* stream) {
  OFSV_REQUIRE(N >= 0 && C >= 0 && H >= 1 && W >= 1, "ofsv_warp2d_bwd_f32: bad shape N=%d C=%d H=%d W=%d", N, C, H, W);
  OFSV_REQUIRE((int64_t)H * W < (1ll << 31), "ofsv_warp2d_bwd_f32: plane too large");
  OFSV_REQUIRE(ref_mode == OFSV_REF_CPU || ref_mode == OFSV_REF_CUDA, "ofsv_warp2d_bwd_f32: bad ref_mode %d", ref_mode);
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) return OFSV_OK;
  if (gflow && C == 0) {
    cudaError_t e = cudaMemsetAsync(gflow, 0, sizeof(float) * (size_t)N * 2 * H * W, st);
    if (e != cudaSuccess) { set_error("ofsv_warp2d_bwd_f32: memset: %s", cudaGetErrorString(e)); return OFSV_ECUDA; }
  }
  if (C == 0 || (!gsrc && !gflow)) return OFSV_OK;
  OFSV_REQUIRE(flow && gout && lin_x && lin_y, "ofsv_warp2d_bwd_f32: null pointer");
  OFSV_REQUIRE(src || !gflow, "ofsv_warp2d_bwd_f32: gflow needs src");
  if (gsrc) {
    cudaError_t e = cudaMemsetAsync(gsrc, 0, sizeof(float) * (size_t)N * C * H * W, st);
    if (e != cudaSuccess) { set_error("ofsv_warp2d_bwd_f32: memset: %s", cudaGetErrorString(e)); return OFSV_ECUDA; }
  }
  int64_t b = cdiv((int64_t)N * H * W, 256);
  const int64_t cap = device_num_sms() * 16;
  const int grid = (int)(b < cap ? b : cap);
  warp2d_bwd_kernel<<<grid, 256, 0, st>>>(src, flow, gout, lin_x, lin_y, gsrc, gflow, N, C, H, W, ref_mode);
  return check_launch("warp2d_bwd_kernel");
}

extern "C" int ofsv_warp3d_bwd_f32(const float* src, const float* flow, const float* gout, const float* lin_h,
                                   const float* lin_d, const float* lin_w, float* gsrc, float* gflow, int N, int C, int D,
                                   int H, int W, int ref_mode, void* stream) {
  OFSV_REQUIRE(N >= 0 && C >= 0 && D >= 1 && H >= 1 && W >= 1, "ofsv_warp3d_bwd_f32: bad shape");
  OFSV_REQUIRE((int64_t)D * H * W < (1ll << 31), "ofsv_warp3d_bwd_f32: volume too large for 32-bit voxel offsets");
  OFSV_REQUIRE(ref_mode == OFSV_REF_CPU || ref_mode == OFSV_REF_CUDA, "ofsv_warp3d_bwd_f32: bad ref_mode %d", ref_mode);
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) return OFSV_OK;
  const size_t V = (size_t)D * H * W;
  if (gflow && C == 0) {
    cudaError_t e = cudaMemsetAsync(gflow, 0, sizeof(float) * (size_t)N * 3 * V, st);
    if (e != cudaSuccess) { set_error("ofsv_warp3d_bwd_f32: memset: %s", cudaGetErrorString(e)); return OFSV_ECUDA; }
  }
  if (C == 0 || (!gsrc && !gflow)) return OFSV_OK;
  OFSV_REQUIRE(flow && gout && lin_h && lin_d && lin_w, "ofsv_warp3d_bwd_f32: null pointer");
  OFSV_REQUIRE(src || !gflow, "ofsv_warp3d_bwd_f32: gflow needs src");
  if (gsrc) {
    cudaError_t e = cudaMemsetAsync(gsrc, 0, sizeof(float) * (size_t)N * C * V, st);
    if (e != cudaSuccess) { set_error("ofsv_warp3d_bwd_f32: memset: %s", cudaGetErrorString(e)); return OFSV_ECUDA; }
  }
  const Warp3dParams P = make_warp3d_params(N, C, D, H, W, ref_mode);
  const dim3 grid((unsigned)cdiv(W, T3W), (unsigned)cdiv(H, T3H), (unsigned)(N * D));
  if (grid.z > 65535u) { set_error("ofsv_warp3d_bwd_f32: N*D=%u exceeds grid.z", grid.z); return OFSV_ENOSUP; }
  const bool vec = (W % 4 == 0) && aligned16(flow) && aligned16(gout) && (!gflow || aligned16(gflow));
  if (vec) warp3d_bwd_kernel<true><<<grid, 256, 0, st>>>(src, flow, gout, lin_h, lin_d, lin_w, gsrc, gflow, P);
  else warp3d_bwd_kernel<false><<<grid, 256, 0, st>>>(src, flow, gout, lin_h, lin_d, lin_w, gsrc, gflow, P);
  return check_launch("warp3d_bwd_kernel");
}
