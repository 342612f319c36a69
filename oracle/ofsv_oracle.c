/* oracle/ofsv_oracle.c — plain-C restatement of the reference's L1 operator arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): built by oracle/Makefile into
 * oracle/libofsv_oracle.so, loaded by oracle/c_oracle.py, used by tests/, smoke() and the
 * cpu_baseline leg of bench.py.  Never linked into or called from the product library.
 *
 * Compiled with -ffp-contract=off: every fp32 operation below is rounded on its own, in the
 * order the reference (PyTorch ATen on CPU) performs it.  Pinned against the imported reference
 * by tests/golden/make_golden.py.
 *
 * Reference lines restated:
 *   warp 2D        Flow-2D/model/warplayer.py:7-26   + ATen grid_sampler (bilinear, border, align_corners=True)
 *   warp 3D        Flow-3D/model/warplayer.py:9-41   + ATen grid_sampler_3d (axis rotation: SURVEY.md fact 2)
 *   blend          Flow-2D/model/IFNet.py:189,240 ; Flow-3D/model/IFNet.py:186,242
 *   correlation    UPFlow/utils/pytorch_correlation.py:27-50 ; call sites UPFlow/model/upflow.py:649,652,655-656
 *   flow upsample  UPFlow/model/pwc_modules.py:77-90
 *   feature warp   UPFlow/model/pwc_modules.py:184-207
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OFSV_DIV_TRUE 0 /* CPU reference: flow / float((S-1)/2) */
#define OFSV_DIV_RCP 1  /* CUDA-eager reference: flow * float(1.0 / double((S-1)/2)) (ATen div_true_kernel_cuda) */

static inline float norm_flow(float f, double half_extent, int div_mode) {
  if (div_mode == OFSV_DIV_RCP) return f * (float)(1.0 / half_extent);
  return f / (float)half_extent;
}

/* grid_sampler_unnormalize(align_corners=True) then clip_coordinates (border padding). */
static inline float unnorm_clip_ac(float g, int size) {
  float p = ((g + 1.0f) / 2.0f) * (float)(size - 1);
  p = fmaxf(p, 0.0f); /* NaN -> 0 like ATen's clamp order */
  p = fminf((float)(size - 1), p);
  return p;
}

int ofsv_oracle_warp2d(const float* src, const float* flow, const float* lin_x, const float* lin_y, float* out,
                       int N, int C, int H, int W, int div_mode) {
  const double sx = (W - 1.0) / 2.0, sy = (H - 1.0) / 2.0; /* warplayer.py:19-20 */
  const int64_t HW = (int64_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const float fx = flow[((int64_t)n * 2 + 0) * HW + (int64_t)y * W + x];
        const float fy = flow[((int64_t)n * 2 + 1) * HW + (int64_t)y * W + x];
        const float gx = lin_x[x] + norm_flow(fx, sx, div_mode);
        const float gy = lin_y[y] + norm_flow(fy, sy, div_mode);
        const float ix = unnorm_clip_ac(gx, W), iy = unnorm_clip_ac(gy, H);
        const float xw = floorf(ix), yn = floorf(iy);
        const float w = ix - xw, e = 1.0f - w, nn = iy - yn, s = 1.0f - nn;
        const float nw = s * e, ne = s * w, sw = nn * e, se = nn * w;
        const int x0 = (int)xw, y0 = (int)yn, x1 = x0 + 1, y1 = y0 + 1;
        const int okx = x1 <= W - 1, oky = y1 <= H - 1;
        for (int c = 0; c < C; ++c) {
          const float* p = src + ((int64_t)n * C + c) * HW;
          const float v00 = p[(int64_t)y0 * W + x0];
          const float v01 = okx ? p[(int64_t)y0 * W + x1] : 0.0f;
          const float v10 = oky ? p[(int64_t)y1 * W + x0] : 0.0f;
          const float v11 = (okx && oky) ? p[(int64_t)y1 * W + x1] : 0.0f;
          /* ATen's vectorised 2-D kernel: (nw_val*nw)+(ne_val*ne)+(sw_val*sw)+(se_val*se), contracted to an FMA chain */
          out[((int64_t)n * C + c) * HW + (int64_t)y * W + x] = fmaf(v11, se, fmaf(v10, sw, fmaf(v01, ne, v00 * nw)));
        }
      }
  return 0;
}

/* Tensor dims (N,C,D,H,W).  lin_h has H entries, lin_d has D, lin_w has W (torch.linspace(-1,1,·)).
 * Grid channel 0 = lin_h[h] + f0/((H-1)/2)  -> sampled along the W axis (size W)
 * Grid channel 1 = lin_d[d] + f1/((D-1)/2)  -> sampled along the H axis (size H)
 * Grid channel 2 = lin_w[w] + f2/((W-1)/2)  -> sampled along the D axis (size D)          */
int ofsv_oracle_warp3d(const float* src, const float* flow, const float* lin_h, const float* lin_d,
                       const float* lin_w, float* out, int N, int C, int D, int H, int W, int div_mode) {
  const double s0 = (H - 1.0) / 2.0, s1 = (D - 1.0) / 2.0, s2 = (W - 1.0) / 2.0; /* warplayer.py:24-26 */
  const int64_t HW = (int64_t)H * W, V = (int64_t)D * HW;
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int d = 0; d < D; ++d)
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          const int64_t o = (int64_t)d * HW + (int64_t)h * W + w;
          const float g0 = lin_h[h] + norm_flow(flow[((int64_t)n * 3 + 0) * V + o], s0, div_mode);
          const float g1 = lin_d[d] + norm_flow(flow[((int64_t)n * 3 + 1) * V + o], s1, div_mode);
          const float g2 = lin_w[w] + norm_flow(flow[((int64_t)n * 3 + 2) * V + o], s2, div_mode);
          const float ix = unnorm_clip_ac(g0, W), iy = unnorm_clip_ac(g1, H), iz = unnorm_clip_ac(g2, D);
          const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
          const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz, x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
          /* ATen grid_sampler_3d_cpu corner weights: products of distances to the opposite corner */
          const float ex = (fx + 1.0f) - ix, wx = ix - fx;
          const float ey = (fy + 1.0f) - iy, wy = iy - fy;
          const float ez = (fz + 1.0f) - iz, wz = iz - fz;
          const float tnw = ex * ey * ez, tne = wx * ey * ez, tsw = ex * wy * ez, tse = wx * wy * ez;
          const float bnw = ex * ey * wz, bne = wx * ey * wz, bsw = ex * wy * wz, bse = wx * wy * wz;
          const int okx = x1 <= W - 1, oky = y1 <= H - 1, okz = z1 <= D - 1;
          for (int c = 0; c < C; ++c) {
            const float* p = src + ((int64_t)n * C + c) * V;
#define AT(z, y, x) p[(int64_t)(z)*HW + (int64_t)(y)*W + (x)]
            float acc = 0.0f;
            acc += AT(z0, y0, x0) * tnw;
            if (okx) acc += AT(z0, y0, x1) * tne;
            if (oky) acc += AT(z0, y1, x0) * tsw;
            if (okx && oky) acc += AT(z0, y1, x1) * tse;
            if (okz) acc += AT(z1, y0, x0) * bnw;
            if (okz && okx) acc += AT(z1, y0, x1) * bne;
            if (okz && oky) acc += AT(z1, y1, x0) * bsw;
            if (okz && okx && oky) acc += AT(z1, y1, x1) * bse;
#undef AT
            out[((int64_t)n * C + c) * V + o] = acc;
          }
        }
  return 0;
}

/* merged = w0*sigmoid(m) + w1*(1-sigmoid(m)) */
int ofsv_oracle_blend(const float* w0, const float* w1, const float* mask_logit, float* merged, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const float m = 1.0f / (1.0f + expf(-mask_logit[i]));
    merged[i] = w0[i] * m + w1[i] * (1.0f - m);
  }
  return 0;
}

/* out[b,(dy+4)*9+(dx+4),y,x] = (1/C) sum_c f1[b,c,y,x]*f2[b,c,y+dy,x+dx], zero outside; optional LeakyReLU. */
int ofsv_oracle_corr81(const float* f1, const float* f2, float* out, int B, int C, int H, int W, float leaky_slope,
                       int apply_leaky) {
  const int64_t HW = (int64_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < 81; ++k) {
      const int dy = k / 9 - 4, dx = k % 9 - 4;
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          const int y2 = y + dy, x2 = x + dx;
          float acc = 0.0f;
          if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W)
            for (int c = 0; c < C; ++c)
              acc += f1[((int64_t)b * C + c) * HW + (int64_t)y * W + x] * f2[((int64_t)b * C + c) * HW + (int64_t)y2 * W + x2];
          float v = acc / (float)C;
          if (apply_leaky && v < 0.0f) v *= leaky_slope;
          out[((int64_t)b * 81 + k) * HW + (int64_t)y * W + x] = v;
        }
    }
  return 0;
}

/* Gradients of the (non-activated) cost volume wrt both inputs — what autograd derives through Corr_pyTorch. */
int ofsv_oracle_corr81_bwd(const float* f1, const float* f2, const float* gout, float* g1, float* g2, int B, int C,
                           int H, int W) {
  const int64_t HW = (int64_t)H * W;
  memset(g1, 0, sizeof(float) * (size_t)B * C * HW);
  memset(g2, 0, sizeof(float) * (size_t)B * C * HW);
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float* a1 = f1 + ((int64_t)b * C + c) * HW;
      const float* a2 = f2 + ((int64_t)b * C + c) * HW;
      float* o1 = g1 + ((int64_t)b * C + c) * HW;
      float* o2 = g2 + ((int64_t)b * C + c) * HW;
      for (int k = 0; k < 81; ++k) {
        const int dy = k / 9 - 4, dx = k % 9 - 4;
        const float* g = gout + ((int64_t)b * 81 + k) * HW;
        for (int y = 0; y < H; ++y) {
          const int y2 = y + dy;
          if (y2 < 0 || y2 >= H) continue;
          for (int x = 0; x < W; ++x) {
            const int x2 = x + dx;
            if (x2 < 0 || x2 >= W) continue;
            const float gv = g[(int64_t)y * W + x] / (float)C;
            o1[(int64_t)y * W + x] += gv * a2[(int64_t)y2 * W + x2];
            o2[(int64_t)y2 * W + x2] += gv * a1[(int64_t)y * W + x];
          }
        }
      }
    }
  return 0;
}

/* bilinear align_corners=True resize of a 2-channel flow (B,2,h_,w_) -> (B,2,h,w); u *= w/w_, v *= h/h_. */
int ofsv_oracle_upsample_flow_ac(const float* in, float* out, int B, int h_, int w_, int h, int w, int if_rate) {
  /* ATen area_pixel_compute_scale(align_corners=True): (in-1)/(out-1) in fp32, 0 if out==1 */
  const float ry = h > 1 ? (float)(h_ - 1) / (float)(h - 1) : 0.0f;
  const float rx = w > 1 ? (float)(w_ - 1) / (float)(w - 1) : 0.0f;
  const float us = (float)((double)w / (double)w_), vs = (float)((double)h / (double)h_);
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < 2; ++c) {
      const float* p = in + ((int64_t)b * 2 + c) * h_ * w_;
      float* q = out + ((int64_t)b * 2 + c) * h * w;
      const float rate = if_rate ? (c == 0 ? us : vs) : 1.0f;
      for (int y = 0; y < h; ++y) {
        const float sy = ry * (float)y;
        const int y0 = (int)sy, y1 = y0 + (y0 < h_ - 1 ? 1 : 0);
        const float ly1 = sy - (float)y0, ly0 = 1.0f - ly1;
        for (int x = 0; x < w; ++x) {
          const float sx = rx * (float)x;
          const int x0 = (int)sx, x1 = x0 + (x0 < w_ - 1 ? 1 : 0);
          const float lx1 = sx - (float)x0, lx0 = 1.0f - lx1;
          const float v = ly0 * (lx0 * p[y0 * w_ + x0] + lx1 * p[y0 * w_ + x1]) +
                          ly1 * (lx0 * p[y1 * w_ + x0] + lx1 * p[y1 * w_ + x1]);
          q[y * w + x] = if_rate ? v * rate : v;
        }
      }
    }
  return 0;
}

/* WarpingLayer_no_div: vgrid = (x,y)+flow ; g = 2*v/max(S-1,1) - 1 ; grid_sample(bilinear, zeros, align_corners=False);
 * output multiplied by [sum of in-bounds corner weights >= 1]. */
int ofsv_oracle_warping_no_div(const float* src, const float* flow, float* out, int B, int C, int H, int W,
                               int div_mode) {
  const int64_t HW = (int64_t)H * W;
  const int dw = W - 1 > 1 ? W - 1 : 1, dh = H - 1 > 1 ? H - 1 : 1;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const float vx = (float)x + flow[((int64_t)b * 2 + 0) * HW + (int64_t)y * W + x];
        const float vy = (float)y + flow[((int64_t)b * 2 + 1) * HW + (int64_t)y * W + x];
        float gx, gy;
        if (div_mode == OFSV_DIV_RCP) {
          gx = (2.0f * vx) * (float)(1.0 / (double)dw) - 1.0f;
          gy = (2.0f * vy) * (float)(1.0 / (double)dh) - 1.0f;
        } else {
          gx = (2.0f * vx) / (float)dw - 1.0f;
          gy = (2.0f * vy) / (float)dh - 1.0f;
        }
        /* unnormalize, align_corners=False.  Both ATen builds fuse this: the AVX2 CPU kernel computes
         * (g+1)*(S/2)-0.5 with a contracted FMA, the CUDA kernel ((g+1)*S-1)/2 with FMAD — the same single rounding. */
        const float ix = fmaf(gx + 1.0f, (float)W, -1.0f) / 2.0f;
        const float iy = fmaf(gy + 1.0f, (float)H, -1.0f) / 2.0f;
        const float xw = floorf(ix), yn = floorf(iy);
        const float w = ix - xw, e = 1.0f - w, nn = iy - yn, s = 1.0f - nn;
        const float nw = s * e, ne = s * w, sw = nn * e, se = nn * w;
        /* floorf of a huge |coordinate| may not fit an int: clamp before the cast (all taps are then out of range) */
        const float xc = fminf(fmaxf(xw, -2.0f), (float)W + 1.0f), yc = fminf(fmaxf(yn, -2.0f), (float)H + 1.0f);
        const int x0 = (int)xc, y0 = (int)yc, x1 = x0 + 1, y1 = y0 + 1;
        const int in00 = x0 >= 0 && x0 < W && y0 >= 0 && y0 < H, in01 = x1 >= 0 && x1 < W && y0 >= 0 && y0 < H;
        const int in10 = x0 >= 0 && x0 < W && y1 >= 0 && y1 < H, in11 = x1 >= 0 && x1 < W && y1 >= 0 && y1 < H;
        const float msum = (((in00 ? nw : 0.0f) + (in01 ? ne : 0.0f)) + (in10 ? sw : 0.0f)) + (in11 ? se : 0.0f);
        const float valid = msum >= 1.0f ? 1.0f : 0.0f;
        for (int c = 0; c < C; ++c) {
          const float* p = src + ((int64_t)b * C + c) * HW;
          const float p00 = in00 ? p[(int64_t)y0 * W + x0] : 0.0f, p01 = in01 ? p[(int64_t)y0 * W + x1] : 0.0f;
          const float p10 = in10 ? p[(int64_t)y1 * W + x0] : 0.0f, p11 = in11 ? p[(int64_t)y1 * W + x1] : 0.0f;
          const float v = fmaf(p11, se, fmaf(p10, sw, fmaf(p01, ne, p00 * nw))); /* same FMA chain as warp2d */
          out[((int64_t)b * C + c) * HW + (int64_t)y * W + x] = v * valid;
        }
      }
  return 0;
}

const char* ofsv_oracle_version(void) { return "ofsv-oracle 1"; }
