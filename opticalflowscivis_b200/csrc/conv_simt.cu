// CUDA-core convolution engine in tap form (ofsv.h: ofsv_conv_desc).  fp32 accumulation, fp32 or bf16 activations,
// fp32 weights [ntaps][Cin_s][Cout_s].  This is the exact-order validation path for the tensor-core engine
// (conv_tc.cu) and the engine used for fp32 (`act_dtype = OFSV_F32`) inference.
//
// Thread = one virtual output position x 16 output channels.  Weight rows are warp-uniform float4 loads (broadcast
// from L1), activations are 16 B vector loads walking the channel axis of one input voxel.
#include "ofsv_common.cuh"

namespace ofsv {

constexpr int COT = 16;  // output channels per thread

template <typename T>
struct Act;
template <>
struct Act<float> {
  static __device__ __forceinline__ void load4(const float* p, float* v) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
  static __device__ __forceinline__ void store16(float* p, const float* v) {
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  static __device__ __forceinline__ void store8(float* p, const float* v) {
#pragma unroll
    for (int i = 0; i < 2; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  static __device__ __forceinline__ void load16(const float* p, float* v) {
#pragma unroll
    for (int i = 0; i < 4; ++i) load4(p + 4 * i, v + 4 * i);
  }
};
template <>
struct Act<__nv_bfloat16> {
  static __device__ __forceinline__ void load4(const __nv_bfloat16* p, float* v) {
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&a.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&a.y);
    v[0] = __low2float(lo); v[1] = __high2float(lo); v[2] = __low2float(hi); v[3] = __high2float(hi);
  }
  static __device__ __forceinline__ void store16(__nv_bfloat16* p, const float* v) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    reinterpret_cast<uint4*>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4*>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
  static __device__ __forceinline__ void store8(__nv_bfloat16* p, const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    reinterpret_cast<uint4*>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  static __device__ __forceinline__ void load16(const __nv_bfloat16* p, float* v) {
#pragma unroll
    for (int i = 0; i < 4; ++i) load4(p + 4 * i, v + 4 * i);
  }
};

template <typename TIN, typename TOUT>
__global__ void __launch_bounds__(128)
    conv_simt_kernel(const ofsv_conv_desc d, const TIN* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ bias, const float* __restrict__ prelu, const TOUT* __restrict__ residual,
                     TOUT* __restrict__ y) {
  const int64_t Vo = (int64_t)d.Do * d.Ho * d.Wo, total = (int64_t)d.N * Vo;
  const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= total) return;
  const int co0 = blockIdx.y * COT;
  const int ph = blockIdx.z;
  const int n = (int)(pos / Vo);
  int r = (int)(pos - (int64_t)n * Vo);
  const int ox = r % d.Wo; r /= d.Wo;
  const int oy = r % d.Ho;
  const int oz = r / d.Ho;

  float acc[COT];
#pragma unroll
  for (int j = 0; j < COT; ++j) acc[j] = __ldg(bias + co0 + j);

  for (int t = 0; t < d.ntaps; ++t) {
    const int8_t* off = d.tap_off[ph * d.ntaps + t];
    const int iz = oz * d.in_stride + off[0], iy = oy * d.in_stride + off[1], ix = ox * d.in_stride + off[2];
    if (iz < 0 || iz >= d.Di || iy < 0 || iy >= d.Hi || ix < 0 || ix >= d.Wi) continue;
    const TIN* xp = x + ((((int64_t)n * d.Di + iz) * d.Hi + iy) * d.Wi + ix) * d.Cin_s;
    const float* wt = w + (int64_t)(ph * d.ntaps + t) * d.Cin_s * d.Cout_w + co0;
    for (int ci = 0; ci < d.Cin_s; ci += 4) {
      float xv[4];
      Act<TIN>::load4(xp + ci, xv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4* wr = reinterpret_cast<const float4*>(wt + (int64_t)(ci + k) * d.Cout_w);
#pragma unroll
        for (int q = 0; q < COT / 4; ++q) {
          const float4 ww = __ldg(wr + q);
          acc[4 * q + 0] = fmaf(xv[k], ww.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(xv[k], ww.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(xv[k], ww.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(xv[k], ww.w, acc[4 * q + 3]);
        }
      }
    }
  }
  if (d.has_prelu) {
#pragma unroll
    for (int j = 0; j < COT; ++j) acc[j] = acc[j] > 0.0f ? acc[j] : acc[j] * __ldg(prelu + co0 + j);
  }
  const int yz = oz * d.out_stride + ((ph >> 2) & 1), yy = oy * d.out_stride + ((ph >> 1) & 1),
            yx = ox * d.out_stride + (ph & 1);
  const int64_t yo = ((((int64_t)n * d.Dy + yz) * d.Hy + yy) * d.Wy + yx) * d.Cout_s + co0;
  if (co0 + COT <= d.Cout_s) {
    if (d.has_residual) {
      float rv[COT];
      Act<TOUT>::load16(residual + yo, rv);
#pragma unroll
      for (int j = 0; j < COT; ++j) acc[j] += rv[j];
    }
    Act<TOUT>::store16(y + yo, acc);
  } else {  // Cout_s % 16 == 8: the last channel group stores 8 channels
    if (d.has_residual) {
      float rv[8];
      Act<TOUT>::load4(residual + yo, rv); Act<TOUT>::load4(residual + yo + 4, rv + 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += rv[j];
    }
    Act<TOUT>::store8(y + yo, acc);
  }
}

int validate_conv_desc(const ofsv_conv_desc* d, const char* who) {
  OFSV_REQUIRE(d != nullptr, "%s: null descriptor", who);
  OFSV_REQUIRE(d->nd == 2 || d->nd == 3, "%s: nd must be 2 or 3", who);
  OFSV_REQUIRE(d->N >= 0 && d->Di >= 1 && d->Hi >= 1 && d->Wi >= 1 && d->Do >= 1 && d->Ho >= 1 && d->Wo >= 1,
               "%s: bad spatial shape", who);
  OFSV_REQUIRE(d->Cin_s >= 16 && d->Cin_s % 16 == 0, "%s: Cin_s=%d must be a multiple of 16", who, d->Cin_s);
  OFSV_REQUIRE(d->Cout_s >= 8 && d->Cout_s % 8 == 0 && d->Cout_w >= 16 && d->Cout_w % 16 == 0 && d->Cout_s <= d->Cout_w &&
                   (d->Cout_w - d->Cout_s < 16 || d->out_shuffle == d->Cout_s),
               "%s: bad output channels (Cout_s=%d Cout_w=%d)", who, d->Cout_s, d->Cout_w);
  OFSV_REQUIRE(d->nphase >= 1 && d->ntaps >= 1 && d->nphase * d->ntaps <= OFSV_MAX_TAPS, "%s: nphase*ntaps out of range", who);
  OFSV_REQUIRE(d->nphase == 1 || d->nphase == (1 << d->nd), "%s: nphase must be 1 or 2^nd", who);
  OFSV_REQUIRE(d->in_stride >= 1 && d->out_stride >= 1 && (d->nphase == 1 || d->out_stride == 2), "%s: bad strides", who);
  {
    const int pz = d->nphase > 1 && d->nd == 3 ? 1 : 0, pyx = d->nphase > 1 ? 1 : 0;
    OFSV_REQUIRE((int64_t)(d->Do - 1) * d->out_stride + pz < d->Dy && (int64_t)(d->Ho - 1) * d->out_stride + pyx < d->Hy &&
                     (int64_t)(d->Wo - 1) * d->out_stride + pyx < d->Wy,
                 "%s: virtual output grid does not fit the output tensor", who);
  }
  OFSV_REQUIRE((d->in_dtype == OFSV_F32 || d->in_dtype == OFSV_BF16) && (d->out_dtype == OFSV_F32 || d->out_dtype == OFSV_BF16),
               "%s: bad dtype", who);
  return OFSV_OK;
}

}  // namespace ofsv

using namespace ofsv;

extern "C" int ofsv_conv_simt(const ofsv_conv_desc* d, const void* x, const float* w, const float* bias,
                              const float* prelu, const void* residual, void* y, void* stream) {
  if (int e = validate_conv_desc(d, "ofsv_conv_simt")) return e;
  if (d->out_shuffle) { set_error("ofsv_conv_simt: depth-to-space heads are only implemented by ofsv_conv_halo"); return OFSV_ENOSUP; }
  if (d->out_s2d) { set_error("ofsv_conv_simt: space-to-depth outputs are only implemented by ofsv_conv_halo"); return OFSV_ENOSUP; }
  if (d->N == 0) return OFSV_OK;
  OFSV_REQUIRE(x && w && bias && y, "ofsv_conv_simt: null pointer");
  OFSV_REQUIRE(!d->has_prelu || prelu, "ofsv_conv_simt: has_prelu without prelu slopes");
  OFSV_REQUIRE(!d->has_residual || residual, "ofsv_conv_simt: has_residual without residual");
  OFSV_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && (!residual || aligned16(residual)),
               "ofsv_conv_simt: pointers must be 16-byte aligned");
  const int64_t total = (int64_t)d->N * d->Do * d->Ho * d->Wo;
  const dim3 grid((unsigned)cdiv(total, 128), (unsigned)(d->Cout_w / COT), (unsigned)d->nphase);
  cudaStream_t st = (cudaStream_t)stream;
#define GO(TI, TO) conv_simt_kernel<TI, TO><<<grid, 128, 0, st>>>(*d, (const TI*)x, w, bias, prelu, (const TO*)residual, (TO*)y)
  if (d->in_dtype == OFSV_F32 && d->out_dtype == OFSV_F32) GO(float, float);
  else if (d->in_dtype == OFSV_BF16 && d->out_dtype == OFSV_BF16) GO(__nv_bfloat16, __nv_bfloat16);
  else if (d->in_dtype == OFSV_BF16 && d->out_dtype == OFSV_F32) GO(__nv_bfloat16, float);
  else GO(float, __nv_bfloat16);
#undef GO
  return check_launch("conv_simt_kernel");
}
