// Memory-bound stages around the IFBlock convolutions (a4, a5): the block-input builder (down-resize + flow rescale +
// concat -> channels-last) and the block-output stage (up-resize + flow*scale + residual add -> fp32 NC(D)HW).
// Both follow ATen's upsample_{bi,tri}linear (align_corners=False) index/weight arithmetic:
//   src = rscale*(dst+0.5)-0.5 (clamped at 0), i0 = min(floor(src), n-1), i1 = i0 + (i0 < n-1), l1 = src-i0, l0 = 1-l1,
//   value = nested  t0*l0 + t1*l1  with W innermost (SURVEY.md Appendix A "Resize identities").
#include <algorithm>

#include "ofsv_common.cuh"

namespace ofsv {

template <typename T>
__device__ __forceinline__ T cvt(float v);
template <>
__device__ __forceinline__ float cvt<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ void store16(float* o, const float* v) {
#pragma unroll
  for (int k = 0; k < 4; ++k) reinterpret_cast<float4*>(o)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}
__device__ __forceinline__ void store16(__nv_bfloat16* o, const float* v) {
  uint32_t w[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    w[k] = *reinterpret_cast<const uint32_t*>(&h);
  }
  reinterpret_cast<uint4*>(o)[0] = make_uint4(w[0], w[1], w[2], w[3]);
  reinterpret_cast<uint4*>(o)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// mean of the 2^nd samples {s*o + s/2 - 1, s*o + s/2} per axis == F.interpolate(1/s) for s in {2,4}; s = 1 reads directly.
template <int ND>
__device__ __forceinline__ float down_sample(const float* __restrict__ vol, int H, int W, int oz, int oy, int ox, int s) {
  if (s == 1) return __ldg(vol + ((int64_t)oz * H + oy) * W + ox);
  const int o = s / 2 - 1;
  const int x = ox * s + o, y = oy * s + o, z = ND == 3 ? oz * s + o : 0;
  float r[2];
#pragma unroll
  for (int dz = 0; dz < (ND == 3 ? 2 : 1); ++dz) {
    float rows[2];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const float* p = vol + ((int64_t)(z + dz) * H + (y + dy)) * W + x;
      rows[dy] = __fadd_rn(__fmul_rn(__ldg(p), 0.5f), __fmul_rn(__ldg(p + 1), 0.5f));
    }
    r[dz] = __fadd_rn(__fmul_rn(rows[0], 0.5f), __fmul_rn(rows[1], 0.5f));
  }
  if (ND == 3) return __fadd_rn(__fmul_rn(r[0], 0.5f), __fmul_rn(r[1], 0.5f));
  return r[0];
}

template <int ND, typename T>
__global__ void __launch_bounds__(256)
    pack_block_input_kernel(const float* __restrict__ img0, const float* __restrict__ img1,
                            const float* __restrict__ warped0, const float* __restrict__ warped1,
                            const float* __restrict__ mask, const float* __restrict__ flow, T* __restrict__ dst, int N,
                            int D, int H, int W, int s, int Cs, int s2d) {
  const int Do = ND == 3 ? D / s : 1, Ho = H / s, Wo = W / s;
  const int64_t V = (int64_t)D * H * W, Vo = (int64_t)Do * Ho * Wo, total = (int64_t)N * Vo;
  const float inv_s = 1.0f / (float)s;
  constexpr int NF = 2 * ND;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / Vo);
    int r = (int)(i - (int64_t)n * Vo);
    const int ox = r % Wo; r /= Wo;
    const int oy = r % Ho;
    const int oz = r / Ho;
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = 0.0f;
    v[0] = down_sample<ND>(img0 + (int64_t)n * V, H, W, oz, oy, ox, s);
    v[1] = down_sample<ND>(img1 + (int64_t)n * V, H, W, oz, oy, ox, s);
    if (flow != nullptr) {
      v[2] = down_sample<ND>(warped0 + (int64_t)n * V, H, W, oz, oy, ox, s);
      v[3] = down_sample<ND>(warped1 + (int64_t)n * V, H, W, oz, oy, ox, s);
      v[4] = down_sample<ND>(mask + (int64_t)n * V, H, W, oz, oy, ox, s);
#pragma unroll
      for (int c = 0; c < NF; ++c)  // F.interpolate(flow, 1/scale) * 1. / scale   (IFNet.py:92 / :88)
        v[5 + c] = __fmul_rn(down_sample<ND>(flow + ((int64_t)n * NF + c) * V, H, W, oz, oy, ox, s), inv_s);
    }
    T* o = dst + (s2d ? s2d_row(ND, n, oz, oy, ox, Do, Ho, Wo) : i) * Cs;
    store16(o, v);                                  // one 32 B (bf16) / 64 B (fp32) vector row per position
    for (int c = 16; c < Cs; ++c) o[c] = cvt<T>(0.0f);
  }
}

__device__ __forceinline__ void ld8(const float* p, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp up_index(int dst, int n_in, float rscale) {
  float src = __fsub_rn(__fmul_rn(rscale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  src = src < 0.0f ? 0.0f : src;
  Lerp L;
  L.i0 = min((int)src, n_in - 1);
  L.i1 = L.i0 + (L.i0 < n_in - 1 ? 1 : 0);
  L.l1 = fminf(fmaxf(__fsub_rn(src, (float)L.i0), 0.0f), 1.0f);
  L.l0 = __fsub_rn(1.0f, L.l1);
  return L;
}

// head: [N][Dh][Hh][Wh][Cs] fp32 channels-last, channels 0..NF-1 flow delta, NF mask delta.
template <int ND>
__global__ void __launch_bounds__(256)
    head_upsample_add_kernel(const float* __restrict__ head, int Cs, const float* __restrict__ flow_prev,
                             const float* __restrict__ mask_prev, float* __restrict__ flow_out,
                             float* __restrict__ mask_out, int N, int D, int H, int W, int s) {
  constexpr int NF = 2 * ND, NC = NF + 1;
  const int Dh = ND == 3 ? D / s : 1, Hh = H / s, Wh = W / s;
  const int64_t V = (int64_t)D * H * W, total = (int64_t)N * V;
  const float rscale = 1.0f / (float)s, fs = (float)s;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / V);
    const int64_t r = i - (int64_t)n * V;
    const int x = (int)(r % W), y = (int)((r / W) % H), z = (int)(r / ((int64_t)W * H));
    float v[NC];
    const float* hb = head + (int64_t)n * Dh * Hh * Wh * Cs;
    if (s == 1) {
      float t8[8];
      ld8(hb + r * Cs, t8);
#pragma unroll
      for (int c = 0; c < NC; ++c) v[c] = t8[c];
    } else {
      const Lerp lx = up_index(x, Wh, rscale), ly = up_index(y, Hh, rscale);
      Lerp lz; lz.i0 = lz.i1 = 0; lz.l0 = 1.0f; lz.l1 = 0.0f;
      if (ND == 3) lz = up_index(z, Dh, rscale);
      float acc_z[2][NC];
#pragma unroll
      for (int dz = 0; dz < (ND == 3 ? 2 : 1); ++dz) {
        const int zz = dz ? lz.i1 : lz.i0;
        float acc_y[2][NC];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const int yy = dy ? ly.i1 : ly.i0;
          float c0[8], c1[8];
          ld8(hb + (((int64_t)zz * Hh + yy) * Wh + lx.i0) * Cs, c0);
          ld8(hb + (((int64_t)zz * Hh + yy) * Wh + lx.i1) * Cs, c1);
#pragma unroll
          for (int c = 0; c < NC; ++c) acc_y[dy][c] = __fmaf_rn(c1[c], lx.l1, __fmul_rn(c0[c], lx.l0));
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) acc_z[dz][c] = __fmaf_rn(acc_y[1][c], ly.l1, __fmul_rn(acc_y[0][c], ly.l0));
      }
#pragma unroll
      for (int c = 0; c < NC; ++c)
        v[c] = ND == 3 ? __fmaf_rn(acc_z[1][c], lz.l1, __fmul_rn(acc_z[0][c], lz.l0)) : acc_z[0][c];
    }
#pragma unroll
    for (int c = 0; c < NF; ++c) {
      float f = __fmul_rn(v[c], fs);  // F.interpolate(flow, scale) * scale
      if (flow_prev) f = __fadd_rn(ldg_stream(flow_prev + ((int64_t)n * NF + c) * V + r), f);  // flow = flow + flow_d
      flow_out[((int64_t)n * NF + c) * V + r] = f;
    }
    float m = v[NF];
    if (mask_prev) m = __fadd_rn(ldg_stream(mask_prev + i), m);
    mask_out[i] = m;
  }
}

// ---- backward of the two stages above (training tier) ----------------------------------------------------
// head_upsample_add: flow_out = flow_prev + s * up_s(head[0..NF)), mask_out = mask_prev + up_s(head[NF]).  The gradient w.r.t. the
// previous flow / mask is the incoming gradient itself; w.r.t. the head it is the ADJOINT of the trilinear up-sampling, evaluated
// as a gather (deterministic): coarse cell c collects every fine position whose two taps (up_index — the forward's own index /
// weight function) include c.  One thread per coarse cell, all NF + 1 channels.
// (A warp-per-cell variant with a shuffle reduction was measured slower at every scale: 0.44 - 0.67 ms against 0.39 ms per step.)
template <int ND>
__global__ void __launch_bounds__(128)
    head_upsample_add_bwd_kernel(const float* __restrict__ gflow, const float* __restrict__ gmask, float* __restrict__ ghead, int N, int D,
                                 int H, int W, int s) {
  constexpr int NF = 2 * ND, NC = NF + 1;
  const int Dh = ND == 3 ? D / s : 1, Hh = H / s, Wh = W / s;
  const int64_t V = (int64_t)D * H * W, Vh = (int64_t)Dh * Hh * Wh, total = (int64_t)N * Vh;
  const float rscale = 1.0f / (float)s, fs = (float)s;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / Vh);
    int r = (int)(i - (int64_t)n * Vh);
    const int cx = r % Wh; r /= Wh;
    const int cy = r % Hh;
    const int cz = r / Hh;
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.f;
    auto weight = [&](int j, int c, int n_in) -> float {
      if (s == 1) return j == c ? 1.f : 0.f;
      const Lerp L = up_index(j, n_in, rscale);
      return (L.i0 == c ? L.l0 : 0.f) + (L.i1 == c ? L.l1 : 0.f);       // both taps on c at the clamped far edge: l0 + l1 = 1
    };
    const int m = s == 1 ? 0 : s;                                        // candidate margin around the cell's own s fine positions
    const int z_lo = ND == 3 ? max(0, s * cz - m) : 0, z_hi = ND == 3 ? min(D, s * cz + s + m) : 1;
    const int y_lo = max(0, s * cy - m), ny = min(H, s * cy + s + m) - y_lo;
    const int x_lo = max(0, s * cx - m), nx = min(W, s * cx + s + m) - x_lo;
    for (int k = 0; k < ny * nx; ++k) {
      const int y = y_lo + k / nx, x = x_lo + k % nx;
      const float wyx = weight(y, cy, Hh) * weight(x, cx, Wh);
      if (wyx == 0.f) continue;
      for (int z = z_lo; z < z_hi; ++z) {
        const float wz = ND == 3 ? weight(z, cz, Dh) : 1.f;
        if (wz == 0.f) continue;
        const float w = wz * wyx;
        const int64_t o = ((int64_t)z * H + y) * W + x;
#pragma unroll
        for (int c = 0; c < NF; ++c) acc[c] = fmaf(w, __ldg(gflow + ((int64_t)n * NF + c) * V + o), acc[c]);
        acc[NF] = fmaf(w, __ldg(gmask + (int64_t)n * V + o), acc[NF]);
      }
    }
    float out[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) out[c] = c < NF ? acc[c] * fs : (c == NF ? acc[NF] : 0.f);
    float4* po = reinterpret_cast<float4*>(ghead + i * 8);
    po[0] = make_float4(out[0], out[1], out[2], out[3]);
    po[1] = make_float4(out[4], out[5], out[6], out[7]);
  }
}

// pack_block_input: channel c of the packed row = resize_{1/s}(source c) (flow channels additionally / s); the resize is the mean of
// the 2^nd samples at offsets {s/2 - 1, s/2} of each s-cell, so a fine voxel receives gx[cell] / 2^nd if it is
// one of them and nothing otherwise.  One thread per fine voxel; img0 / img1 (and a teacher's gt) take no gradient.
template <int ND>
__global__ void __launch_bounds__(256)
    pack_block_input_bwd_kernel(const __nv_bfloat16* __restrict__ gx, float* __restrict__ gw0, float* __restrict__ gw1, float* __restrict__ gmask,
                                float* __restrict__ gflow, int N, int D, int H, int W, int s) {
  constexpr int NF = 2 * ND;
  const int Do = ND == 3 ? D / s : 1, Ho = H / s, Wo = W / s;
  const int64_t V = (int64_t)D * H * W, total = (int64_t)N * V;
  const float wgt = s == 1 ? 1.0f : (ND == 3 ? 0.125f : 0.25f), inv_s = 1.0f / (float)s;
  const int lo = s / 2 - 1, hi = s / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / V);
    const int64_t r = i - (int64_t)n * V;
    const int x = (int)(r % W), y = (int)((r / W) % H), z = (int)(r / ((int64_t)W * H));
    auto inside = [&](int j) { const int o = j % s; return s == 1 || o == lo || o == hi; };
    float v[3 + NF];
#pragma unroll
    for (int c = 0; c < 3 + NF; ++c) v[c] = 0.f;
    if (inside(x) && inside(y) && (ND == 2 || inside(z))) {
      const int64_t cell = (((int64_t)n * Do + (ND == 3 ? z / s : 0)) * Ho + y / s) * Wo + x / s;
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(gx + cell * 16)), b = __ldg(reinterpret_cast<const uint4*>(gx + cell * 16) + 1);
      const __nv_bfloat16* ha = reinterpret_cast<const __nv_bfloat16*>(&a);
      const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&b);
      float row[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) { row[c] = __bfloat162float(ha[c]); row[8 + c] = __bfloat162float(hb[c]); }
      v[0] = row[2] * wgt; v[1] = row[3] * wgt; v[2] = row[4] * wgt;
#pragma unroll
      for (int c = 0; c < NF; ++c) v[3 + c] = row[5 + c] * wgt * inv_s;
    }
    gw0[i] = v[0]; gw1[i] = v[1]; gmask[i] = v[2];
#pragma unroll
    for (int c = 0; c < NF; ++c) gflow[((int64_t)n * NF + c) * V + r] = v[3 + c];
  }
}

// uint8 volume -> fp32 in [0,1]: float(x) / 255.0f (true division, what the reference loaders compute on the host:
// Datasets/read_data.py, Flow-3D/load_datasets.py).  16 voxels per thread: one 16 B load, four 16 B stores.
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t n, float div) {
  const int64_t nv = n >> 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float4* o = reinterpret_cast<float4*>(dst) + i * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = make_float4(__fdiv_rn((float)(w[k] & 0xFFu), div), __fdiv_rn((float)((w[k] >> 8) & 0xFFu), div),
                         __fdiv_rn((float)((w[k] >> 16) & 0xFFu), div), __fdiv_rn((float)(w[k] >> 24), div));
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 15)) {
    const int64_t i = (nv << 4) + threadIdx.x;
    dst[i] = __fdiv_rn((float)src[i], div);
  }
}

static inline int grid_1d(int64_t total) {
  int64_t b = cdiv(total, 256);
  const int64_t cap = device_num_sms() * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// fp32 volume -> uint8: (uint8)(x * mul) with the product clamped to [0, 255] and truncated toward zero — the reference's
// export `(img * 255).byte()` (Flow-3D/inference_img.py:105, Flow-2D/inference_img.py) done BEFORE the download, so that an
// interpolated byte volume crosses PCIe as bytes.  16 voxels per thread.
__global__ void __launch_bounds__(256) f32_to_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int64_t n, float mul) {
  const int64_t nv = n >> 4;
  auto q = [mul](float x) -> uint32_t { return (uint32_t)fminf(fmaxf(__fmul_rn(x, mul), 0.0f), 255.0f); };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    const float4* in = reinterpret_cast<const float4*>(src) + i * 4;
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 v = __ldg(in + k);
      w[k] = q(v.x) | (q(v.y) << 8) | (q(v.z) << 16) | (q(v.w) << 24);
    }
    reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 15)) {
    const int64_t i = (nv << 4) + threadIdx.x;
    dst[i] = (uint8_t)q(src[i]);
  }
}

}  // namespace ofsv

using namespace ofsv;

extern "C" int ofsv_pack_block_input(const float* img0, const float* img1, const float* warped0, const float* warped1,
                                     const float* mask, const float* flow, void* dst, int act_dtype, int nd, int N,
                                     int D, int H, int W, int scale, int Cs, int s2d, void* stream) {
  OFSV_REQUIRE(nd == 2 || nd == 3, "ofsv_pack_block_input: nd must be 2 or 3");
  OFSV_REQUIRE(!s2d || ((nd == 2 || D % (2 * scale) == 0) && H % (2 * scale) == 0 && W % (2 * scale) == 0),
               "ofsv_pack_block_input: space-to-depth packing needs dims that are multiples of 2*scale");
  OFSV_REQUIRE(scale == 1 || scale == 2 || scale == 4, "ofsv_pack_block_input: scale %d not in {1,2,4}", scale);
  OFSV_REQUIRE(N >= 0 && D >= 1 && H >= 1 && W >= 1, "ofsv_pack_block_input: bad shape");
  OFSV_REQUIRE((nd == 2 ? D == 1 : D % scale == 0) && H % scale == 0 && W % scale == 0,
               "ofsv_pack_block_input: spatial dims must be multiples of scale");
  OFSV_REQUIRE(Cs >= 16 && Cs % 16 == 0, "ofsv_pack_block_input: Cs must be a multiple of 16");
  if (N == 0) return OFSV_OK;
  OFSV_REQUIRE(img0 && img1 && dst, "ofsv_pack_block_input: null pointer");
  OFSV_REQUIRE(aligned16(dst), "ofsv_pack_block_input: dst must be 16-byte aligned");
  OFSV_REQUIRE(flow == nullptr || (warped0 && warped1 && mask), "ofsv_pack_block_input: flow needs warped0/1 and mask");
  const int64_t total = (int64_t)N * (nd == 3 ? D / scale : 1) * (H / scale) * (W / scale);
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_1d(total);
#define GO(ND, T) pack_block_input_kernel<ND, T><<<g, 256, 0, st>>>(img0, img1, warped0, warped1, mask, flow, (T*)dst, N, D, H, W, scale, Cs, s2d)
  if (act_dtype == OFSV_F32) { if (nd == 2) GO(2, float); else GO(3, float); }
  else if (act_dtype == OFSV_BF16) { if (nd == 2) GO(2, __nv_bfloat16); else GO(3, __nv_bfloat16); }
  else { set_error("ofsv_pack_block_input: bad act_dtype %d", act_dtype); return OFSV_EINVAL; }
#undef GO
  return check_launch("pack_block_input_kernel");
}

extern "C" int ofsv_head_upsample_add(const float* head, int Cs, const float* flow_prev, const float* mask_prev,
                                      float* flow_out, float* mask_out, int nd, int N, int D, int H, int W, int scale,
                                      void* stream) {
  OFSV_REQUIRE(nd == 2 || nd == 3, "ofsv_head_upsample_add: nd must be 2 or 3");
  OFSV_REQUIRE(scale == 1 || scale == 2 || scale == 4, "ofsv_head_upsample_add: scale %d not in {1,2,4}", scale);
  OFSV_REQUIRE(N >= 0 && D >= 1 && H >= 1 && W >= 1 && Cs >= 8 && Cs % 4 == 0, "ofsv_head_upsample_add: bad shape (Cs must be >= 8, multiple of 4)");
  OFSV_REQUIRE((nd == 2 ? D == 1 : D % scale == 0) && H % scale == 0 && W % scale == 0,
               "ofsv_head_upsample_add: spatial dims must be multiples of scale");
  OFSV_REQUIRE((flow_prev == nullptr) == (mask_prev == nullptr), "ofsv_head_upsample_add: flow_prev and mask_prev go together");
  if (N == 0) return OFSV_OK;
  OFSV_REQUIRE(head && flow_out && mask_out, "ofsv_head_upsample_add: null pointer");
  OFSV_REQUIRE(aligned16(head), "ofsv_head_upsample_add: head must be 16-byte aligned");
  const int g = grid_1d((int64_t)N * D * H * W);
  cudaStream_t st = (cudaStream_t)stream;
  if (nd == 2) head_upsample_add_kernel<2><<<g, 256, 0, st>>>(head, Cs, flow_prev, mask_prev, flow_out, mask_out, N, D, H, W, scale);
  else head_upsample_add_kernel<3><<<g, 256, 0, st>>>(head, Cs, flow_prev, mask_prev, flow_out, mask_out, N, D, H, W, scale);
  return check_launch("head_upsample_add_kernel");
}

// ------------------------------------------------------------------------------------------------ layout changes
// NC(P) fp32 (P = product of the spatial dims) <-> channels-last bf16 [N][P][Cs], through a 32-pixel x 64-channel shared tile so that
// both sides move 128-byte segments.  pack concatenates up to eight sources along the channel axis (torch.cat + zero padding + cast
// of the estimator input torch.cat([corr, x_1x1, flow], 1), UPFlow/model/upflow.py:657, in one pass) and zero-fills the padding.
namespace ofsv {
struct NhwcSrc { const float* p[8]; int c[8]; int nsrc; };
constexpr int NHWC_TILES = 8;
__global__ void __launch_bounds__(256) pack_nhwc_kernel(const NhwcSrc S, __nv_bfloat16* __restrict__ dst, int64_t P, int Cs, int tiles) {
  __shared__ float tile[32][65];
  const int n = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // ty: 8 channel rows per pass
  for (int t = 0; t < tiles; ++t) {                             // a CTA walks `tiles` consecutive 32-pixel tiles (8 on large tensors)
  const int64_t p0 = ((int64_t)blockIdx.x * tiles + t) * 32;
  if (p0 >= P) break;
  for (int c0 = 0; c0 < Cs; c0 += 64) {
    const int ncc = min(64, Cs - c0);
    for (int cc = ty; cc < ncc; cc += 8) {
      const int c = c0 + cc;
      float v = 0.f;
      if (p0 + tx < P) {
        int base = 0;
        for (int s = 0; s < S.nsrc; ++s) {
          if (c >= base && c < base + S.c[s]) v = __ldg(S.p[s] + ((int64_t)n * S.c[s] + (c - base)) * P + p0 + tx);
          base += S.c[s];
        }
      }
      tile[tx][cc] = v;
    }
    __syncthreads();
    // 32 pixels x 64 channels -> rows of 128 bytes: thread = (pixel, 8-channel group)
    const int px = threadIdx.x >> 3, g8 = threadIdx.x & 7;
    if (p0 + px < P && c0 + g8 * 8 < Cs) {
      uint4 o;
      __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(tile[px][g8 * 8 + 2 * i], tile[px][g8 * 8 + 2 * i + 1]);
      *reinterpret_cast<uint4*>(dst + ((int64_t)n * P + p0 + px) * Cs + c0 + g8 * 8) = o;
    }
    __syncthreads();
  }
  }
}
__global__ void __launch_bounds__(256) unpack_nhwc_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t P, int Cs, int C, int tiles) {
  __shared__ float tile[32][65];
  const int n = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int t = 0; t < tiles; ++t) {
  const int64_t p0 = ((int64_t)blockIdx.x * tiles + t) * 32;
  if (p0 >= P) break;
  for (int c0 = 0; c0 < C; c0 += 64) {
    const int px = threadIdx.x >> 3, g8 = threadIdx.x & 7;
    if (p0 + px < P && c0 + g8 * 8 < Cs) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + ((int64_t)n * P + p0 + px) * Cs + c0 + g8 * 8);
      const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(v2[i]);
        tile[px][g8 * 8 + 2 * i] = f.x;
        tile[px][g8 * 8 + 2 * i + 1] = f.y;
      }
    }
    __syncthreads();
    for (int cc = ty; cc < 64; cc += 8) {
      const int c = c0 + cc;
      if (c < C && p0 + tx < P) dst[((int64_t)n * C + c) * P + p0 + tx] = tile[tx][cc];
    }
    __syncthreads();
  }
  }
}
}  // namespace ofsv

extern "C" int ofsv_pack_nhwc_bf16(const float* const* srcs, const int* channels, int nsrc, void* dst, int N, int64_t P, int Cs, void* stream) {
  OFSV_REQUIRE(srcs && channels && nsrc >= 1 && nsrc <= 8, "ofsv_pack_nhwc_bf16: 1..8 sources");
  OFSV_REQUIRE(N >= 0 && N <= 65535 && P >= 0 && Cs >= 8 && Cs % 8 == 0, "ofsv_pack_nhwc_bf16: bad shape");
  ofsv::NhwcSrc S;
  int total = 0;
  for (int i = 0; i < 8; ++i) { S.p[i] = i < nsrc ? srcs[i] : nullptr; S.c[i] = i < nsrc ? channels[i] : 0; total += S.c[i]; }
  S.nsrc = nsrc;
  OFSV_REQUIRE(total <= Cs, "ofsv_pack_nhwc_bf16: %d source channels do not fit Cs = %d", total, Cs);
  if ((int64_t)N * P == 0) return OFSV_OK;
  for (int i = 0; i < nsrc; ++i) OFSV_REQUIRE(srcs[i] != nullptr && channels[i] > 0, "ofsv_pack_nhwc_bf16: null / empty source %d", i);
  OFSV_REQUIRE(dst && aligned16(dst), "ofsv_pack_nhwc_bf16: dst must be 16-byte aligned");
  const int tiles = (int64_t)N * P >= (1 << 20) ? ofsv::NHWC_TILES : 1;
  ofsv::pack_nhwc_kernel<<<dim3((unsigned)cdiv(P, 32 * tiles), (unsigned)N), 256, 0, (cudaStream_t)stream>>>(S, static_cast<__nv_bfloat16*>(dst), P, Cs, tiles);
  return check_launch("pack_nhwc_kernel");
}

extern "C" int ofsv_unpack_nhwc_f32(const void* src, float* dst, int N, int64_t P, int Cs, int C, void* stream) {
  OFSV_REQUIRE(N >= 0 && N <= 65535 && P >= 0 && Cs >= 8 && Cs % 8 == 0 && C >= 1 && C <= Cs, "ofsv_unpack_nhwc_f32: bad shape");
  if ((int64_t)N * P == 0) return OFSV_OK;
  OFSV_REQUIRE(src && dst && aligned16(src), "ofsv_unpack_nhwc_f32: null / misaligned pointer");
  const int tiles = (int64_t)N * P >= (1 << 20) ? ofsv::NHWC_TILES : 1;
  ofsv::unpack_nhwc_kernel<<<dim3((unsigned)cdiv(P, 32 * tiles), (unsigned)N), 256, 0, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16*>(src), dst, P, Cs, C, tiles);
  return check_launch("unpack_nhwc_kernel");
}

extern "C" int ofsv_head_upsample_add_bwd(const float* gflow, const float* gmask, float* ghead, int nd, int N, int D, int H, int W, int scale,
                                          void* stream) {
  OFSV_REQUIRE(nd == 2 || nd == 3, "ofsv_head_upsample_add_bwd: nd must be 2 or 3");
  OFSV_REQUIRE(scale == 1 || scale == 2 || scale == 4, "ofsv_head_upsample_add_bwd: scale must be 1, 2 or 4");
  OFSV_REQUIRE(N >= 0 && D >= 1 && H >= 1 && W >= 1 && (nd == 2 ? D == 1 : D % scale == 0) && H % scale == 0 && W % scale == 0,
               "ofsv_head_upsample_add_bwd: spatial dims must be multiples of scale");
  if (N == 0) return OFSV_OK;
  OFSV_REQUIRE(gflow && gmask && ghead && aligned16(ghead), "ofsv_head_upsample_add_bwd: null / misaligned pointer");
  const int64_t cells = (int64_t)N * (nd == 3 ? D / scale : 1) * (H / scale) * (W / scale);
  const int g = (int)std::min<int64_t>(cdiv(cells, 128), (int64_t)device_num_sms() * 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (nd == 2) head_upsample_add_bwd_kernel<2><<<g, 128, 0, st>>>(gflow, gmask, ghead, N, D, H, W, scale);
  else head_upsample_add_bwd_kernel<3><<<g, 128, 0, st>>>(gflow, gmask, ghead, N, D, H, W, scale);
  return check_launch("head_upsample_add_bwd_kernel");
}

extern "C" int ofsv_pack_block_input_bwd(const void* gx, float* gw0, float* gw1, float* gmask, float* gflow, int nd, int N, int D, int H, int W,
                                         int scale, void* stream) {
  OFSV_REQUIRE(nd == 2 || nd == 3, "ofsv_pack_block_input_bwd: nd must be 2 or 3");
  OFSV_REQUIRE(scale == 1 || scale == 2 || scale == 4, "ofsv_pack_block_input_bwd: scale must be 1, 2 or 4");
  OFSV_REQUIRE(N >= 0 && D >= 1 && H >= 1 && W >= 1 && (nd == 2 ? D == 1 : D % scale == 0) && H % scale == 0 && W % scale == 0,
               "ofsv_pack_block_input_bwd: spatial dims must be multiples of scale");
  if (N == 0) return OFSV_OK;
  OFSV_REQUIRE(gx && gw0 && gw1 && gmask && gflow && aligned16(gx), "ofsv_pack_block_input_bwd: null / misaligned pointer");
  const int g = grid_1d((int64_t)N * D * H * W);
  cudaStream_t st = (cudaStream_t)stream;
  if (nd == 2) pack_block_input_bwd_kernel<2><<<g, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(gx), gw0, gw1, gmask, gflow, N, D, H, W, scale);
  else pack_block_input_bwd_kernel<3><<<g, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(gx), gw0, gw1, gmask, gflow, N, D, H, W, scale);
  return check_launch("pack_block_input_bwd_kernel");
}

extern "C" int ofsv_u8_to_f32(const uint8_t* src, float* dst, int64_t n, float div, void* stream) {
  OFSV_REQUIRE(n >= 0 && div != 0.0f, "ofsv_u8_to_f32: bad arguments");
  if (n == 0) return OFSV_OK;
  OFSV_REQUIRE(src && dst, "ofsv_u8_to_f32: null pointer");
  OFSV_REQUIRE(aligned16(src) && aligned16(dst), "ofsv_u8_to_f32: pointers must be 16-byte aligned");
  u8_to_f32_kernel<<<grid_1d(cdiv(n, 16)), 256, 0, (cudaStream_t)stream>>>(src, dst, n, div);
  return check_launch("u8_to_f32_kernel");
}

extern "C" int ofsv_f32_to_u8(const float* src, uint8_t* dst, int64_t n, float mul, void* stream) {
  OFSV_REQUIRE(n >= 0, "ofsv_f32_to_u8: negative size");
  if (n == 0) return OFSV_OK;
  OFSV_REQUIRE(src && dst, "ofsv_f32_to_u8: null pointer");
  OFSV_REQUIRE(aligned16(src) && aligned16(dst), "ofsv_f32_to_u8: pointers must be 16-byte aligned");
  f32_to_u8_kernel<<<grid_1d(cdiv(n, 16)), 256, 0, (cudaStream_t)stream>>>(src, dst, n, mul);
  return check_launch("f32_to_u8_kernel");
}
