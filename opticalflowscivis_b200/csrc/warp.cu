// Backward-warp gather kernels (a1, a2) and the fused sigmoid/warp/blend stage (a6) for sm_100a.
//
// Layout notes
//  * 2-D: output x runs along source x, so a thread owns 4 consecutive x (float4 flow loads / stores) and the four
//    bilinear taps of neighbouring threads fall in the same 128 B lines.
//  * 3-D: the reference warp ROTATES axes: output (d,h,w) samples source (z,y,x) ~ (w,d,h).
//    Flow/out are contiguous along w but the source is contiguous along x ~ h.  A CTA therefore owns a 32(h) x 32(w)
//    tile at fixed (n,d): flow is read coalesced along w (float4) into shared memory, the gather runs with lanes along
//    h (coalesced 128 B source reads, the z+1 neighbour of column w is the z tap of column w+1 -> L1 hits), and the
//    results go back through shared memory so the stores are again float4 along w.
//  * All coordinate arithmetic replicates the reference op-for-op (ofsv_common.cuh); 1e-5 parity needs it.
#include "ofsv_common.cuh"
#include "warp_device.cuh"

namespace ofsv {

// ----------------------------------------------------------------------------------------------------
// 2-D
// ----------------------------------------------------------------------------------------------------
struct Bilin {
  int i00, i01, i10, i11;  // element offsets inside one (H,W) plane; -1 = tap outside (contributes 0)
  float nw, ne, sw, se;
};

__device__ __forceinline__ Bilin bilin_setup(float fx, float fy, float lx, float ly, int H, int W, float hx, float hy,
                                             float rhx, float rhy, int ref_mode) {
  const float gx = __fadd_rn(lx, norm_flow(fx, hx, rhx, ref_mode));
  const float gy = __fadd_rn(ly, norm_flow(fy, hy, rhy, ref_mode));
  const float ix = unnorm_clip_ac(gx, (float)(W - 1)), iy = unnorm_clip_ac(gy, (float)(H - 1));
  const float xw = floorf(ix), yn = floorf(iy);
  const float w = __fsub_rn(ix, xw), e = __fsub_rn(1.0f, w), n = __fsub_rn(iy, yn), s = __fsub_rn(1.0f, n);
  Bilin b;
  b.nw = __fmul_rn(s, e); b.ne = __fmul_rn(s, w); b.sw = __fmul_rn(n, e); b.se = __fmul_rn(n, w);
  const int x0 = (int)xw, y0 = (int)yn;
  const bool okx = x0 + 1 <= W - 1, oky = y0 + 1 <= H - 1;
  b.i00 = y0 * W + x0;
  b.i01 = okx ? b.i00 + 1 : -1;
  b.i10 = oky ? b.i00 + W : -1;
  b.i11 = (okx && oky) ? b.i00 + W + 1 : -1;
  return b;
}
__device__ __forceinline__ float bilin_sample(const float* __restrict__ p, const Bilin& b) {
  const float v00 = __ldg(p + b.i00);
  const float v01 = b.i01 >= 0 ? __ldg(p + b.i01) : 0.0f;
  const float v10 = b.i10 >= 0 ? __ldg(p + b.i10) : 0.0f;
  const float v11 = b.i11 >= 0 ? __ldg(p + b.i11) : 0.0f;
  // both ATen builds evaluate nw*v00 + ne*v01 + sw*v10 + se*v11 as an FMA chain (AVX2 contraction / FMAD)
  return __fmaf_rn(v11, b.se, __fmaf_rn(v10, b.sw, __fmaf_rn(v01, b.ne, __fmul_rn(v00, b.nw))));
}

// one thread = one (n, y, x) pixel, loops over channels
__global__ void __launch_bounds__(256) warp2d_kernel(const float* __restrict__ src, const float* __restrict__ flow,
                                                     const float* __restrict__ lin_x, const float* __restrict__ lin_y,
                                                     float* __restrict__ out, int N, int C, int H, int W, int ref_mode) {
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
    const Bilin b = bilin_setup(fx, fy, __ldg(lin_x + x), __ldg(lin_y + y), H, W, hx, hy, rhx, rhy, ref_mode);
    for (int c = 0; c < C; ++c) out[((int64_t)n * C + c) * HW + r] = bilin_sample(src + ((int64_t)n * C + c) * HW, b);
  }
}

__global__ void __launch_bounds__(256)
    warp_blend_2d_kernel(const float* __restrict__ img0, const float* __restrict__ img1, const float* __restrict__ flow,
                         const float* __restrict__ mask_logit, const float* __restrict__ lin_x,
                         const float* __restrict__ lin_y, float* __restrict__ warped0, float* __restrict__ warped1,
                         float* __restrict__ merged, float* __restrict__ mask_sig, int N, int H, int W, int ref_mode) {
  const int64_t HW = (int64_t)H * W;
  const int64_t total = (int64_t)N * HW;
  const float hx = (float)((W - 1.0) / 2.0), hy = (float)((H - 1.0) / 2.0);
  const float rhx = (float)(1.0 / ((W - 1.0) / 2.0)), rhy = (float)(1.0 / ((H - 1.0) / 2.0));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW);
    const int r = (int)(i - (int64_t)n * HW);
    const int y = r / W, x = r - y * W;
    const float* f = flow + (int64_t)n * 4 * HW + r;
    const float lx = __ldg(lin_x + x), ly = __ldg(lin_y + y);
    const Bilin b0 = bilin_setup(ldg_stream(f), ldg_stream(f + HW), lx, ly, H, W, hx, hy, rhx, rhy, ref_mode);
    const Bilin b1 = bilin_setup(ldg_stream(f + 2 * HW), ldg_stream(f + 3 * HW), lx, ly, H, W, hx, hy, rhx, rhy, ref_mode);
    const float w0 = bilin_sample(img0 + (int64_t)n * HW, b0);
    const float w1 = bilin_sample(img1 + (int64_t)n * HW, b1);
    if (warped0) warped0[i] = w0;
    if (warped1) warped1[i] = w1;
    if (merged || mask_sig) {
      const float m = sigmoidf_ref(ldg_stream(mask_logit + i));
      if (mask_sig) mask_sig[i] = m;
      if (merged) merged[i] = __fadd_rn(__fmul_rn(w0, m), __fmul_rn(w1, __fsub_rn(1.0f, m)));
    }
  }
}

// ----------------------------------------------------------------------------------------------------
// 3-D
// ----------------------------------------------------------------------------------------------------
template <bool VEC, bool FMA>
__global__ void __launch_bounds__(256)
    warp3d_kernel(const float* __restrict__ src, const float* __restrict__ flow, const float* __restrict__ lin_h,
                  const float* __restrict__ lin_d, const float* __restrict__ lin_w, float* __restrict__ out,
                  const Warp3dParams P) {
  __shared__ float sf[3][T3H][T3P];
  __shared__ float so[1][T3H][T3P];
  const int H = P.H, W = P.W, D = P.D, HW = H * W;
  const int64_t V = (int64_t)D * HW;
  const int n = blockIdx.z / D, d = blockIdx.z - n * D;
  const int h0 = blockIdx.y * T3H, w0 = blockIdx.x * T3W;
  const float* fl = flow + (int64_t)n * 3 * V + (int64_t)d * HW;
  load_planes<3, VEC>(sf, [&](int k) { return fl + (int64_t)k * V; }, h0, w0, H, W);
  __syncthreads();
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int h = h0 + lane, w = w0 + wl;
  const bool ok = h < H && w < W;
  Trilin t;
  if (ok) t = trilin_setup(sf[0][lane][wl], sf[1][lane][wl], sf[2][lane][wl], __ldg(lin_h + h), __ldg(lin_d + d),
                           __ldg(lin_w + w), D, H, W, P.hs, P.ref_mode);
  for (int c = 0; c < P.C; ++c) {
    if (ok) so[0][lane][wl] = trilin_sample<FMA>(src + ((int64_t)n * P.C + c) * V, t, W, HW);
    __syncthreads();
    float* o = out + ((int64_t)n * P.C + c) * V + (int64_t)d * HW;
    store_planes<1, VEC>(so, [&](int) { return o; }, h0, w0, H, W);
    __syncthreads();
  }
}

// fused: sigmoid(mask), warp(img0, flow[:, :3]), warp(img1, flow[:, 3:6]), merged.  7 input planes -> up to 4 output planes.
template <bool VEC, bool FMA>
__global__ void __launch_bounds__(256)
    warp_blend_3d_kernel(const float* __restrict__ img0, const float* __restrict__ img1, const float* __restrict__ flow,
                         const float* __restrict__ mask_logit, const float* __restrict__ lin_h,
                         const float* __restrict__ lin_d, const float* __restrict__ lin_w, float* __restrict__ warped0,
                         float* __restrict__ warped1, float* __restrict__ merged, float* __restrict__ mask_sig,
                         const Warp3dParams P) {
  __shared__ float si[7][T3H][T3P];  // flow0..5, mask logit
  __shared__ float so[4][T3H][T3P];  // warped0, warped1, merged, sigmoid(mask)
  const int H = P.H, W = P.W, D = P.D, HW = H * W;
  const int64_t V = (int64_t)D * HW;
  const int n = blockIdx.z / D, d = blockIdx.z - n * D;
  const int h0 = blockIdx.y * T3H, w0 = blockIdx.x * T3W;
  const int64_t plane = (int64_t)n * V + (int64_t)d * HW;
  const float* fl = flow + (int64_t)n * 6 * V + (int64_t)d * HW;
  const bool need_m = (merged != nullptr) || (mask_sig != nullptr);
  load_planes<7, VEC>(si, [&](int k) -> const float* {
    return k < 6 ? fl + (int64_t)k * V : (need_m ? mask_logit + plane : nullptr); }, h0, w0, H, W);
  __syncthreads();
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int h = h0 + lane, w = w0 + wl;
  if (h < H && w < W) {
    const float lh = __ldg(lin_h + h), ld = __ldg(lin_d + d), lw = __ldg(lin_w + w);
    const Trilin t0 = trilin_setup(si[0][lane][wl], si[1][lane][wl], si[2][lane][wl], lh, ld, lw, D, H, W, P.hs, P.ref_mode);
    const Trilin t1 = trilin_setup(si[3][lane][wl], si[4][lane][wl], si[5][lane][wl], lh, ld, lw, D, H, W, P.hs, P.ref_mode);
    const Taps8 g0 = trilin_gather(img0 + (int64_t)n * V, t0), g1 = trilin_gather(img1 + (int64_t)n * V, t1);
    const float a = trilin_reduce<FMA>(g0, t0), b = trilin_reduce<FMA>(g1, t1);
    float m = 0.0f, mg = 0.0f;
    if (need_m) {
      m = sigmoidf_ref(si[6][lane][wl]);
      mg = __fadd_rn(__fmul_rn(a, m), __fmul_rn(b, __fsub_rn(1.0f, m)));
    }
    so[0][lane][wl] = a; so[1][lane][wl] = b; so[2][lane][wl] = mg; so[3][lane][wl] = m;
  }
  __syncthreads();
  store_planes<4, VEC>(so, [&](int k) -> float* {
    float* b = k == 0 ? warped0 : (k == 1 ? warped1 : (k == 2 ? merged : mask_sig));
    return b ? b + plane : nullptr; }, h0, w0, H, W);
}

__global__ void __launch_bounds__(256) blend_kernel(const float* __restrict__ w0, const float* __restrict__ w1,
                                                    const float* __restrict__ mask_logit, float* __restrict__ merged,
                                                    int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = sigmoidf_ref(ldg_stream(mask_logit + i));
    merged[i] = __fadd_rn(__fmul_rn(ldg_stream(w0 + i), m), __fmul_rn(ldg_stream(w1 + i), __fsub_rn(1.0f, m)));
  }
}

static inline int grid_1d(int64_t total) {
  int64_t b = cdiv(total, 256);
  const int64_t cap = 148 * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace ofsv

using namespace ofsv;

extern "C" int ofsv_warp2d_f32(const float* src, const float* flow, const float* lin_x, const float* lin_y, float* out,
                               int N, int C, int H, int W, int ref_mode, void* stream) {
  OFSV_REQUIRE(N >= 0 && C >= 0 && H >= 1 && W >= 1, "ofsv_warp2d_f32: bad shape N=%d C=%d H=%d W=%d", N, C, H, W);
  if ((int64_t)N * C == 0) return OFSV_OK;   // empty batch: nothing to do (pointers may be null)
  OFSV_REQUIRE(src && flow && lin_x && lin_y && out, "ofsv_warp2d_f32: null pointer");
  OFSV_REQUIRE((int64_t)H * W < (1ll << 31), "ofsv_warp2d_f32: plane too large");
  OFSV_REQUIRE(ref_mode == OFSV_REF_CPU || ref_mode == OFSV_REF_CUDA, "ofsv_warp2d_f32: bad ref_mode %d", ref_mode);
  warp2d_kernel<<<grid_1d((int64_t)N * H * W), 256, 0, (cudaStream_t)stream>>>(src, flow, lin_x, lin_y, out, N, C, H, W,
                                                                                ref_mode);
  return check_launch("warp2d_kernel");
}

extern "C" int ofsv_warp_blend_2d_f32(const float* img0, const float* img1, const float* flow, const float* mask_logit,
                                      const float* lin_x, const float* lin_y, float* warped0, float* warped1,
                                      float* merged, float* mask_sig, int N, int H, int W, int ref_mode, void* stream) {
  OFSV_REQUIRE(N >= 0 && H >= 1 && W >= 1 && (int64_t)H * W < (1ll << 31), "ofsv_warp_blend_2d_f32: bad shape");
  OFSV_REQUIRE(ref_mode == OFSV_REF_CPU || ref_mode == OFSV_REF_CUDA, "ofsv_warp_blend_2d_f32: bad ref_mode");
  if (N == 0) return OFSV_OK;
  OFSV_REQUIRE(img0 && img1 && flow && lin_x && lin_y, "ofsv_warp_blend_2d_f32: null pointer");
  OFSV_REQUIRE(mask_logit || (!merged && !mask_sig), "ofsv_warp_blend_2d_f32: merged/mask_sig need mask_logit");
  warp_blend_2d_kernel<<<grid_1d((int64_t)N * H * W), 256, 0, (cudaStream_t)stream>>>(
      img0, img1, flow, mask_logit, lin_x, lin_y, warped0, warped1, merged, mask_sig, N, H, W, ref_mode);
  return check_launch("warp_blend_2d_kernel");
}

extern "C" int ofsv_warp3d_f32(const float* src, const float* flow, const float* lin_h, const float* lin_d,
                               const float* lin_w, float* out, int N, int C, int D, int H, int W, int ref_mode,
                               void* stream) {
  OFSV_REQUIRE(N >= 0 && C >= 0 && D >= 1 && H >= 1 && W >= 1, "ofsv_warp3d_f32: bad shape");
  if ((int64_t)N * C == 0) return OFSV_OK;   // empty batch
  OFSV_REQUIRE(src && flow && lin_h && lin_d && lin_w && out, "ofsv_warp3d_f32: null pointer");
  OFSV_REQUIRE((int64_t)D * H * W < (1ll << 31), "ofsv_warp3d_f32: volume too large for 32-bit voxel offsets");
  OFSV_REQUIRE((int64_t)N * D <= 65535 * 1ll * 65535, "ofsv_warp3d_f32: N*D too large");
  OFSV_REQUIRE(ref_mode == OFSV_REF_CPU || ref_mode == OFSV_REF_CUDA, "ofsv_warp3d_f32: bad ref_mode %d", ref_mode);
  const Warp3dParams P = make_warp3d_params(N, C, D, H, W, ref_mode);
  const dim3 grid((unsigned)cdiv(W, T3W), (unsigned)cdiv(H, T3H), (unsigned)(N * D));
  const bool vec = (W % 4 == 0) && aligned16(flow) && aligned16(out);
  cudaStream_t st = (cudaStream_t)stream;
  if (grid.z > 65535u) { set_error("ofsv_warp3d_f32: N*D=%u exceeds grid.z", grid.z); return OFSV_ENOSUP; }
#define LAUNCH(V, F) warp3d_kernel<V, F><<<grid, 256, 0, st>>>(src, flow, lin_h, lin_d, lin_w, out, P)
  if (vec) { if (ref_mode == OFSV_REF_CUDA) LAUNCH(true, true); else LAUNCH(true, false); }
  else     { if (ref_mode == OFSV_REF_CUDA) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
  return check_launch("warp3d_kernel");
}

extern "C" int ofsv_warp_blend_3d_f32(const float* img0, const float* img1, const float* flow, const float* mask_logit,
                                      const float* lin_h, const float* lin_d, const float* lin_w, float* warped0,
                                      float* warped1, float* merged, float* mask_sig, int N, int D, int H, int W,
                                      int ref_mode, void* stream) {
  OFSV_REQUIRE(N >= 0 && D >= 1 && H >= 1 && W >= 1 && (int64_t)D * H * W < (1ll << 31), "ofsv_warp_blend_3d_f32: bad shape");
  OFSV_REQUIRE(ref_mode == OFSV_REF_CPU || ref_mode == OFSV_REF_CUDA, "ofsv_warp_blend_3d_f32: bad ref_mode");
  if (N == 0) return OFSV_OK;
  OFSV_REQUIRE(img0 && img1 && flow && lin_h && lin_d && lin_w, "ofsv_warp_blend_3d_f32: null pointer");
  OFSV_REQUIRE(mask_logit || (!merged && !mask_sig), "ofsv_warp_blend_3d_f32: merged/mask_sig need mask_logit");
  const Warp3dParams P = make_warp3d_params(N, 1, D, H, W, ref_mode);
  const dim3 grid((unsigned)cdiv(W, T3W), (unsigned)cdiv(H, T3H), (unsigned)(N * D));
  if (grid.z > 65535u) { set_error("ofsv_warp_blend_3d_f32: N*D=%u exceeds grid.z", grid.z); return OFSV_ENOSUP; }
  bool vec = (W % 4 == 0) && aligned16(flow);
  const void* ptrs[] = {mask_logit, warped0, warped1, merged, mask_sig};
  for (const void* p : ptrs) vec = vec && (p == nullptr || aligned16(p));
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(V, F) \
  warp_blend_3d_kernel<V, F><<<grid, 256, 0, st>>>(img0, img1, flow, mask_logit, lin_h, lin_d, lin_w, warped0, warped1, merged, mask_sig, P)
  if (vec) { if (ref_mode == OFSV_REF_CUDA) LAUNCH(true, true); else LAUNCH(true, false); }
  else     { if (ref_mode == OFSV_REF_CUDA) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
  return check_launch("warp_blend_3d_kernel");
}

extern "C" int ofsv_blend_f32(const float* w0, const float* w1, const float* mask_logit, float* merged, int64_t n,
                              void* stream) {
  OFSV_REQUIRE(n >= 0, "ofsv_blend_f32: negative size");
  if (n == 0) return OFSV_OK;
  OFSV_REQUIRE(w0 && w1 && mask_logit && merged, "ofsv_blend_f32: null pointer");
  blend_kernel<<<grid_1d(n), 256, 0, (cudaStream_t)stream>>>(w0, w1, mask_logit, merged, n);
  return check_launch("blend_kernel");
}
