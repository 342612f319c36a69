// Backward of the UPFlow flow-path operators a10 / a11 (the training step of BASELINE cfg 5 differentiates through both):
//   upsample2d_flow_as  — UPFlow/model/pwc_modules.py:77-90 (F.interpolate bilinear align_corners=True, then * (w/w_, h/h_))
//   WarpingLayer_no_div — UPFlow/model/pwc_modules.py:184-207 (grid_sample zeros / align_corners=False, times the constant
//                         validity mask of grid_sample(ones) >= 1)
// i.e. ATen's upsample_bilinear2d_backward and grid_sampler_2d_backward chained with the element-wise ops around them.
// Coordinates / cells / weights are computed exactly as in the forward kernels (upflow_ops.cu); the scatter sums use
// red.global.add.f32 (not order-deterministic), parity is to 1e-5 relative.
#include "ofsv_common.cuh"

namespace ofsv {

__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
static inline int grid_1d_bwd(int64_t total) {
  int64_t b = cdiv(total, 256);
  const int64_t cap = device_num_sms() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// gin (B,2,h_,w_) += weights * gout (B,2,h,w) * (u_scale | v_scale)
__global__ void __launch_bounds__(256)
    upsample_flow_ac_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int B, int h_, int w_, int h, int w,
                                float ry, float rx, float us, float vs, int if_rate) {
  const int64_t total = (int64_t)B * 2 * h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const int64_t bc = i / ((int64_t)h * w);
    const int c = (int)(bc & 1);
    float* p = gin + bc * h_ * w_;
    const float sy = __fmul_rn(ry, (float)y), sx = __fmul_rn(rx, (float)x);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < h_ - 1 ? 1 : 0), x1 = x0 + (x0 < w_ - 1 ? 1 : 0);
    const float ly1 = __fsub_rn(sy, (float)y0), ly0 = __fsub_rn(1.0f, ly1);
    const float lx1 = __fsub_rn(sx, (float)x0), lx0 = __fsub_rn(1.0f, lx1);
    float g = ldg_stream(gout + i);
    if (if_rate) g *= (c == 0 ? us : vs);
    red_add_f32(p + y0 * w_ + x0, ly0 * lx0 * g);
    red_add_f32(p + y0 * w_ + x1, ly0 * lx1 * g);
    red_add_f32(p + y1 * w_ + x0, ly1 * lx0 * g);
    red_add_f32(p + y1 * w_ + x1, ly1 * lx1 * g);
  }
}

__global__ void __launch_bounds__(256)
    warping_no_div_bwd_kernel(const float* __restrict__ src, const float* __restrict__ flow, const float* __restrict__ gout,
                              float* __restrict__ gsrc, float* __restrict__ gflow, int B, int C, int H, int W, float dw, float dh,
                              float rdw, float rdh, int ref_mode, int c_per) {
  // blockIdx.y = channel chunk (see warping_no_div_kernel); with more than one chunk the flow gradient is accumulated into the
  // zero-filled gflow with red.add
  const int c_begin = blockIdx.y * c_per, c_end = min(C, c_begin + c_per);
  const bool split = gridDim.y > 1;
  const int64_t HW = (int64_t)H * W, total = (int64_t)B * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int r = (int)(i - (int64_t)b * HW);
    const int y = r / W, x = r - y * W;
    const float vx = __fadd_rn((float)x, ldg_stream(flow + ((int64_t)b * 2 + 0) * HW + r));
    const float vy = __fadd_rn((float)y, ldg_stream(flow + ((int64_t)b * 2 + 1) * HW + r));
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
    const float nw = __fmul_rn(s, e), ne = __fmul_rn(s, w), sw = __fmul_rn(n, e), se = __fmul_rn(n, w);
    const float xc = fminf(fmaxf(xw, -2.0f), (float)W + 1.0f), yc = fminf(fmaxf(yn, -2.0f), (float)H + 1.0f);
    const int x0 = (int)xc, y0 = (int)yc, x1 = x0 + 1, y1 = y0 + 1;
    const bool ix0 = x0 >= 0 && x0 < W, ix1 = x1 >= 0 && x1 < W, iy0 = y0 >= 0 && y0 < H, iy1 = y1 >= 0 && y1 < H;
    const bool in00 = ix0 && iy0, in01 = ix1 && iy0, in10 = ix0 && iy1, in11 = ix1 && iy1;
    const float msum = __fadd_rn(__fadd_rn(__fadd_rn(in00 ? nw : 0.0f, in01 ? ne : 0.0f), in10 ? sw : 0.0f), in11 ? se : 0.0f);
    const bool valid = msum >= 1.0f;          // the mask is a constant of the graph ((mask >= 1).float())
    float gix = 0.0f, giy = 0.0f;
    if (valid) {
      for (int c = c_begin; c < c_end; ++c) {
        const int64_t pl = ((int64_t)b * C + c) * HW;
        const float g = ldg_stream(gout + pl + r);
        if (gsrc) {
          float* q = gsrc + pl;
          if (in00) red_add_f32(q + (int64_t)y0 * W + x0, nw * g);
          if (in01) red_add_f32(q + (int64_t)y0 * W + x1, ne * g);
          if (in10) red_add_f32(q + (int64_t)y1 * W + x0, sw * g);
          if (in11) red_add_f32(q + (int64_t)y1 * W + x1, se * g);
        }
        if (gflow) {
          const float* p = src + pl;
          const float p00 = in00 ? __ldg(p + (int64_t)y0 * W + x0) : 0.0f, p01 = in01 ? __ldg(p + (int64_t)y0 * W + x1) : 0.0f;
          const float p10 = in10 ? __ldg(p + (int64_t)y1 * W + x0) : 0.0f, p11 = in11 ? __ldg(p + (int64_t)y1 * W + x1) : 0.0f;
          gix += ((p01 - p00) * s + (p11 - p10) * n) * g;
          giy += ((p10 - p00) * e + (p11 - p01) * w) * g;
        }
      }
    }
    if (gflow) {
      // grid_sampler backward: * W/2 (align_corners=False un-normalisation, no clipping in zeros mode); then the backward of
      // 2 * v / max(W-1, 1) - 1 (pwc_modules.py:198-199): / dw, * 2
      float fx = gix * (0.5f * (float)W), fy = giy * (0.5f * (float)H);
      fx = ref_mode == OFSV_REF_CUDA ? fx * rdw : __fdiv_rn(fx, dw);
      fy = ref_mode == OFSV_REF_CUDA ? fy * rdh : __fdiv_rn(fy, dh);
      if (split) {
        if (valid) { red_add_f32(gflow + ((int64_t)b * 2 + 0) * HW + r, 2.0f * fx); red_add_f32(gflow + ((int64_t)b * 2 + 1) * HW + r, 2.0f * fy); }
      } else {
        gflow[((int64_t)b * 2 + 0) * HW + r] = 2.0f * fx;
        gflow[((int64_t)b * 2 + 1) * HW + r] = 2.0f * fy;
      }
    }
  }
}

}  // namespace ofsv

using namespace ofsv;

extern "C" int ofsv_upsample_flow_ac_bwd_f32(const float* gout, float* gin, int B, int h_in, int w_in, int h_out, int w_out,
                                             int if_rate, void* stream) {
  OFSV_REQUIRE(B >= 0 && h_in >= 1 && w_in >= 1 && h_out >= 1 && w_out >= 1, "ofsv_upsample_flow_ac_bwd_f32: bad shape");
  if (B == 0) return OFSV_OK;
  OFSV_REQUIRE(gout && gin, "ofsv_upsample_flow_ac_bwd_f32: null pointer");
  OFSV_REQUIRE((int64_t)h_in * w_in < (1ll << 31) && (int64_t)h_out * w_out < (1ll << 31), "ofsv_upsample_flow_ac_bwd_f32: plane too large");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(gin, 0, sizeof(float) * (size_t)B * 2 * h_in * w_in, st);
  if (e != cudaSuccess) { set_error("ofsv_upsample_flow_ac_bwd_f32: memset: %s", cudaGetErrorString(e)); return OFSV_ECUDA; }
  // same source-index ratios as the forward (ATen area_pixel_compute_scale, align_corners=True)
  const float ry = h_out > 1 ? (float)(h_in - 1) / (float)(h_out - 1) : 0.0f;
  const float rx = w_out > 1 ? (float)(w_in - 1) / (float)(w_out - 1) : 0.0f;
  const float us = (float)((double)w_out / (double)w_in), vs = (float)((double)h_out / (double)h_in);
  upsample_flow_ac_bwd_kernel<<<grid_1d_bwd((int64_t)B * 2 * h_out * w_out), 256, 0, st>>>(gout, gin, B, h_in, w_in, h_out, w_out,
                                                                                          ry, rx, us, vs, if_rate);
  return check_launch("upsample_flow_ac_bwd_kernel");
}

extern "C" int ofsv_warping_no_div_bwd_f32(const float* src, const float* flow, const float* gout, float* gsrc, float* gflow,
                                           int B, int C, int H, int W, int ref_mode, void* stream) {
  OFSV_REQUIRE(B >= 0 && C >= 0 && H >= 1 && W >= 1 && (int64_t)H * W < (1ll << 31), "ofsv_warping_no_div_bwd_f32: bad shape");
  OFSV_REQUIRE(ref_mode == OFSV_REF_CPU || ref_mode == OFSV_REF_CUDA, "ofsv_warping_no_div_bwd_f32: bad ref_mode");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return OFSV_OK;
  if (gflow && C == 0) {
    cudaError_t e = cudaMemsetAsync(gflow, 0, sizeof(float) * (size_t)B * 2 * H * W, st);
    if (e != cudaSuccess) { set_error("ofsv_warping_no_div_bwd_f32: memset: %s", cudaGetErrorString(e)); return OFSV_ECUDA; }
  }
  if (C == 0 || (!gsrc && !gflow)) return OFSV_OK;
  OFSV_REQUIRE(flow && gout, "ofsv_warping_no_div_bwd_f32: null pointer");
  OFSV_REQUIRE(src || !gflow, "ofsv_warping_no_div_bwd_f32: gflow needs src");
  if (gsrc) {
    cudaError_t e = cudaMemsetAsync(gsrc, 0, sizeof(float) * (size_t)B * C * H * W, st);
    if (e != cudaSuccess) { set_error("ofsv_warping_no_div_bwd_f32: memset: %s", cudaGetErrorString(e)); return OFSV_ECUDA; }
  }
  const int dw = W - 1 > 1 ? W - 1 : 1, dh = H - 1 > 1 ? H - 1 : 1;
  const int gx = grid_1d_bwd((int64_t)B * H * W);
  int nsplit = (int)(cdiv(device_num_sms() * 8, gx));
  nsplit = nsplit < 1 ? 1 : (nsplit > cdiv(C, 4) ? (int)cdiv(C, 4) : nsplit);
  const int c_per = (int)cdiv(C, nsplit);
  const unsigned gy = (unsigned)cdiv(C, c_per);
  if (gflow && gy > 1) {
    cudaError_t e = cudaMemsetAsync(gflow, 0, sizeof(float) * (size_t)B * 2 * H * W, st);
    if (e != cudaSuccess) { set_error("ofsv_warping_no_div_bwd_f32: memset: %s", cudaGetErrorString(e)); return OFSV_ECUDA; }
  }
  warping_no_div_bwd_kernel<<<dim3((unsigned)gx, gy), 256, 0, st>>>(
      src, flow, gout, gsrc, gflow, B, C, H, W, (float)dw, (float)dh, (float)(1.0 / (double)dw), (float)(1.0 / (double)dh), ref_mode, c_per);
  return check_launch("warping_no_div_bwd_kernel");
}
