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
// Warp-autonomous 3-D warp: no block-level barrier anywhere.  A task = (n, d, 32 consecutive h, W3_WT consecutive w), owned
// by one WARP; tasks are numbered with the w tile fastest so that warps running at the same time read neighbouring pieces of
// the same flow rows (DRAM page hits) and gather from the same few source planes (L2 hits).  Per task
//   * the three flow planes (32 rows x 64 B) arrive by cp.async into the warp's private, double-buffered shared tile — the
//     tile of the warp's NEXT task is in flight while the current one is processed, so streaming reads never stall the warp;
//   * lane = row h: it reads its flow vectors with 16 B shared loads and processes 4 voxels at a time (32 independent
//     gathers in flight per lane, 12 warps per SM); lanes along h make every tap a 128 B-coalesced read of the rotated
//     source, and walking along w steps through consecutive source z planes (the z+1 taps of one voxel are the z taps of
//     the next: L1 hits);
//   * each lane stores its 16 results as four 16 B vectors (64 contiguous bytes of its output row).
constexpr int W3_WT = 16;                 // voxels along w per tile
constexpr int W3_ROW = W3_WT + 4;         // floats per tile row in shared memory (80 B: 16 B-aligned, conflict-free for LDS.128)
constexpr int W3_WARPS = 4;
constexpr int W3_PLANE = 32 * W3_ROW;     // floats per (plane, tile)
constexpr int W3_SMEM_PER_WARP = 2 * 3 * W3_PLANE * 4;

__device__ __forceinline__ void w3_cp16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void w3_cp4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

#ifndef OFSV_W3_MINB
#define OFSV_W3_MINB 3
#endif
template <bool VEC, bool FMA>
__global__ void __launch_bounds__(32 * W3_WARPS, OFSV_W3_MINB)
    warp3d_kernel(const float* __restrict__ src, const float* __restrict__ flow, const float* __restrict__ lin_h,
                  const float* __restrict__ lin_d, const float* __restrict__ lin_w, float* __restrict__ out,
                  const Warp3dParams P, const uint32_t ntasks) {
  extern __shared__ __align__(16) float w3_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* sf = w3_smem + wid * (W3_SMEM_PER_WARP / 4);     // [2][3][32][W3_ROW]
  const int H = P.H, W = P.W, D = P.D, HW = H * W;
  const int64_t V = (int64_t)D * HW;
  const int hblocks = (H + 31) >> 5;
  const int ntw = (W + W3_WT - 1) / W3_WT;
  const uint32_t stride = gridDim.x * W3_WARPS;      // task ids are 32-bit (host check): the decode is plain 32-bit division
  struct Task { int n, d, h0, w0; };
  auto decode = [&](uint32_t id) {
    Task k;
    k.w0 = (int)(id % ntw) * W3_WT; id /= ntw;
    k.h0 = (int)(id % hblocks) << 5; id /= hblocks;
    k.d = (int)(id % D);
    k.n = (int)(id / D);
    return k;
  };
  // async copy of the flow tile of task `id` (3 planes x 32 rows x 16 floats) into stage st: row = i*8 + lane/4, chunk = lane%4
  auto issue = [&](uint32_t id, int st) {
    if (id < ntasks) {
      const Task k = decode(id);
      const float* fl = flow + (int64_t)k.n * 3 * V + (int64_t)k.d * HW;
      float* dst = sf + st * 3 * W3_PLANE;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = i * 8 + (lane >> 2), c4 = (lane & 3) * 4;
        if (k.h0 + row < H) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* g = fl + (int64_t)c * V + (int64_t)(k.h0 + row) * W + k.w0 + c4;
            float* s = dst + c * W3_PLANE + row * W3_ROW + c4;
            if (VEC && k.w0 + c4 + 3 < W) {
              w3_cp16(s, g);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (k.w0 + c4 + e < W) w3_cp4(s + e, g + e);
            }
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  uint32_t task = blockIdx.x * W3_WARPS + wid;
  issue(task, 0);
  for (int it = 0; task < ntasks; task += stride, ++it) {
    const int st = it & 1;
    issue(task + stride, st ^ 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    const Task k = decode(task);
    const int h = k.h0 + lane, w0 = k.w0;
    const bool okh = h < H;
    const float lh = okh ? __ldg(lin_h + h) : 0.0f, ld = __ldg(lin_d + k.d);
    const float* f0 = sf + st * 3 * W3_PLANE + lane * W3_ROW;
    for (int c = 0; c < P.C; ++c) {
      const float* sp = src + ((int64_t)k.n * P.C + c) * V;
      if (okh) {
        float* o = out + ((int64_t)k.n * P.C + c) * V + (int64_t)k.d * HW + (int64_t)h * W + w0;
#pragma unroll
        for (int q4 = 0; q4 < W3_WT / 4; ++q4) {
          const float4 a = *reinterpret_cast<const float4*>(f0 + q4 * 4);
          const float4 b = *reinterpret_cast<const float4*>(f0 + W3_PLANE + q4 * 4);
          const float4 e = *reinterpret_cast<const float4*>(f0 + 2 * W3_PLANE + q4 * 4);
          const float fa[4] = {a.x, a.y, a.z, a.w}, fb[4] = {b.x, b.y, b.z, b.w}, fe[4] = {e.x, e.y, e.z, e.w};
          float r[4];
          {
            Trilin tr[4];
            Taps8 tp[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int w = min(w0 + q4 * 4 + j, W - 1);        // columns past W reuse the last one (never stored)
              tr[j] = trilin_setup(fa[j], fb[j], fe[j], lh, ld, __ldg(lin_w + w), D, H, W, P.hs, P.ref_mode);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) tp[j] = trilin_gather(sp, tr[j]);     // 32 independent gathers in flight
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = trilin_reduce<FMA>(tp[j], tr[j]);
          }
          if (VEC && w0 + q4 * 4 + 3 < W) {
            stg_stream4(o + q4 * 4, make_float4(r[0], r[1], r[2], r[3]));
          } else {
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2)
              if (w0 + q4 * 4 + e2 < W) o[q4 * 4 + e2] = r[e2];
          }
        }
      }
    }
    __syncwarp();            // every lane has read stage st before the next iteration's copies overwrite it
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
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
                                                    float* __restrict__ mask_sig, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = sigmoidf_ref(ldg_stream(mask_logit + i));
    if (mask_sig) mask_sig[i] = m;
    if (merged) merged[i] = __fadd_rn(__fmul_rn(ldg_stream(w0 + i), m), __fmul_rn(ldg_stream(w1 + i), __fsub_rn(1.0f, m)));
  }
}

static inline int grid_1d(int64_t total) {
  int64_t b = cdiv(total, 256);
  const int64_t cap = device_num_sms() * 16;
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

namespace ofsv {
int warp3d_slab_try(const float* src, const float* flow, const float* lin_h, const float* lin_d, const float* lin_w, float* out,
                    int N, int C, int D, int H, int W, int ref_mode, cudaStream_t st, int flow_nc = 3, int flow_c0 = 0);   // warp3d_slab.cu
}
// engine: 0 = pick (the TMA slab kernel on cubic volumes it supports, else the global-gather kernel), 1 = global gather only
static int warp3d_dispatch(const float* src, const float* flow, const float* lin_h, const float* lin_d, const float* lin_w,
                           float* out, int N, int C, int D, int H, int W, int ref_mode, void* stream, int engine) {
  OFSV_REQUIRE(N >= 0 && C >= 0 && D >= 1 && H >= 1 && W >= 1, "ofsv_warp3d_f32: bad shape");
  if ((int64_t)N * C == 0) return OFSV_OK;   // empty batch
  OFSV_REQUIRE(src && flow && lin_h && lin_d && lin_w && out, "ofsv_warp3d_f32: null pointer");
  OFSV_REQUIRE((int64_t)D * H * W < (1ll << 31), "ofsv_warp3d_f32: volume too large for 32-bit voxel offsets");
  OFSV_REQUIRE(ref_mode == OFSV_REF_CPU || ref_mode == OFSV_REF_CUDA, "ofsv_warp3d_f32: bad ref_mode %d", ref_mode);
  if (engine == 0) {
    const int rc = warp3d_slab_try(src, flow, lin_h, lin_d, lin_w, out, N, C, D, H, W, ref_mode, (cudaStream_t)stream);
    if (rc != 0) return rc < 0 ? rc : OFSV_OK;
  }
  const Warp3dParams P = make_warp3d_params(N, C, D, H, W, ref_mode);
  const bool vec = (W % 4 == 0) && aligned16(flow) && aligned16(out);
  cudaStream_t st = (cudaStream_t)stream;
  const long long ntasks64 = (long long)N * D * cdiv(H, 32) * cdiv(W, W3_WT);
  OFSV_REQUIRE(ntasks64 < (1ll << 31) - 148 * 64, "ofsv_warp3d_f32: too many tiles for 32-bit task ids");
  const uint32_t ntasks = (uint32_t)ntasks64;
  const int64_t ctas = cdiv(ntasks64, W3_WARPS);
  const int grid = (int)(ctas < device_num_sms() * OFSV_W3_MINB ? ctas : device_num_sms() * OFSV_W3_MINB);      // OFSV_W3_MINB CTAs of 4 warps resident per SM, persistent over the task list
  const int smem = W3_WARPS * W3_SMEM_PER_WARP;
  static std::atomic<uint64_t> attr_done[4];
  if (int e = ensure_dyn_smem(attr_done[0], warp3d_kernel<true, true>, smem, "ofsv_warp3d_f32")) return e;
  if (int e = ensure_dyn_smem(attr_done[1], warp3d_kernel<true, false>, smem, "ofsv_warp3d_f32")) return e;
  if (int e = ensure_dyn_smem(attr_done[2], warp3d_kernel<false, true>, smem, "ofsv_warp3d_f32")) return e;
  if (int e = ensure_dyn_smem(attr_done[3], warp3d_kernel<false, false>, smem, "ofsv_warp3d_f32")) return e;
#define LAUNCH(V, F) warp3d_kernel<V, F><<<grid, 32 * W3_WARPS, smem, st>>>(src, flow, lin_h, lin_d, lin_w, out, P, ntasks)
  if (vec) { if (ref_mode == OFSV_REF_CUDA) LAUNCH(true, true); else LAUNCH(true, false); }
  else     { if (ref_mode == OFSV_REF_CUDA) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
  return check_launch("warp3d_kernel");
}

extern "C" int ofsv_warp3d_f32(const float* src, const float* flow, const float* lin_h, const float* lin_d,
                               const float* lin_w, float* out, int N, int C, int D, int H, int W, int ref_mode,
                               void* stream) {
  return warp3d_dispatch(src, flow, lin_h, lin_d, lin_w, out, N, C, D, H, W, ref_mode, stream, 0);
}
extern "C" int ofsv_warp3d_gather_f32(const float* src, const float* flow, const float* lin_h, const float* lin_d,
                                      const float* lin_w, float* out, int N, int C, int D, int H, int W, int ref_mode,
                                      void* stream) {
  return warp3d_dispatch(src, flow, lin_h, lin_d, lin_w, out, N, C, D, H, W, ref_mode, stream, 1);
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
  cudaStream_t st0 = (cudaStream_t)stream;
  if (warped0 && warped1) {
    // cubic volumes with both warped outputs wanted: two TMA slab warps (flow channels 0-2 / 3-5 of the 6-channel tensor) and
    // one streaming sigmoid/blend pass — 245 us per 256^3 volume where the single gather kernel below takes 379 us; results
    // are bit-identical (shared coordinate / weight / sum helpers, same blend expression)
    const int r0 = warp3d_slab_try(img0, flow, lin_h, lin_d, lin_w, warped0, N, 1, D, H, W, ref_mode, st0, 6, 0);
    if (r0 < 0) return r0;
    if (r0 == 1) {
      const int r1 = warp3d_slab_try(img1, flow, lin_h, lin_d, lin_w, warped1, N, 1, D, H, W, ref_mode, st0, 6, 3);
      if (r1 < 0) return r1;
      if (r1 == 1) {
        if (merged || mask_sig) {
          const int64_t n = (int64_t)N * D * H * W;
          blend_kernel<<<grid_1d(n), 256, 0, st0>>>(warped0, warped1, mask_logit, merged, mask_sig, n);
          return check_launch("blend_kernel");
        }
        return OFSV_OK;
      }
    }
  }
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
  blend_kernel<<<grid_1d(n), 256, 0, (cudaStream_t)stream>>>(w0, w1, mask_logit, merged, nullptr, n);
  return check_launch("blend_kernel");
}
