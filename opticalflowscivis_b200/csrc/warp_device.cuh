// Device helpers shared by the 3-D warp kernels (warp.cu) and the fused IFBlock output stage (block_finish.cu).
#pragma once
#include "ofsv_common.cuh"

namespace ofsv {

struct Trilin {
  // unsigned 32-bit ELEMENT indices: a tap address is then one IMAD.WIDE.U32 (index * 4 + base pointer) instead of a
  // sign-extended 64-bit add + LEA pair per tap (a quarter of the gather kernels' instructions before)
  uint32_t base;       // z0*HW + y0*W + x0
  uint32_t dx, dy, dz; // element offset of the +1 neighbour along each source axis, 0 when it lies outside the volume
  float ex, wx, ey, wy, ez, wz;
};

// (f0,f1,f2) = flow channels; lh/ld/lw = linspace entries of THIS output voxel's (h,d,w).
// Integer cell (x0,y0,z0) of the sample in the source volume + the six 1-D weights, op for op like the reference.
struct TrilinCell {
  int x0, y0, z0;
  float ex, wx, ey, wy, ez, wz;
};
// MAGIC: floor through the 2^23 trick instead of FRND + F2I (identical results).  One instruction more per axis but none on
// the quarter-rate conversion unit: 2 % faster in the latency-bound slab kernel, 4-5 % slower in the issue-bound kernels.
template <bool MAGIC = false>
__device__ __forceinline__ TrilinCell trilin_cell(float f0, float f1, float f2, float lh, float ld, float lw, int D, int H,
                                                  int W, const float* hs, int ref_mode) {
  const float g0 = __fadd_rn(lh, norm_flow(f0, hs[0], hs[3], ref_mode));  // sampled along the W axis
  const float g1 = __fadd_rn(ld, norm_flow(f1, hs[1], hs[4], ref_mode));  // along H
  const float g2 = __fadd_rn(lw, norm_flow(f2, hs[2], hs[5], ref_mode));  // along D
  const float ix = unnorm_clip_ac(g0, (float)(W - 1)), iy = unnorm_clip_ac(g1, (float)(H - 1)),
              iz = unnorm_clip_ac(g2, (float)(D - 1));
  float fx, fy, fz;
  int xi, yi, zi;
  if (MAGIC) {
  // floor of a clipped coordinate 0 <= v <= S-1 < 2^22 without the quarter-rate conversion unit (FRND + F2I per axis): adding
  // 2^23 with round-down leaves 2^23 + floor(v) exactly (the ulp there is 1), whose low mantissa bits are the integer
    const float tx = __fadd_rd(ix, 8388608.0f), ty = __fadd_rd(iy, 8388608.0f), tz = __fadd_rd(iz, 8388608.0f);
    fx = __fsub_rn(tx, 8388608.0f); fy = __fsub_rn(ty, 8388608.0f); fz = __fsub_rn(tz, 8388608.0f);
    xi = __float_as_int(tx) - 0x4B000000; yi = __float_as_int(ty) - 0x4B000000; zi = __float_as_int(tz) - 0x4B000000;
  } else {
    fx = floorf(ix); fy = floorf(iy); fz = floorf(iz);
    xi = (int)fx; yi = (int)fy; zi = (int)fz;
  }
  TrilinCell t;
  t.ex = __fsub_rn(__fadd_rn(fx, 1.0f), ix); t.wx = __fsub_rn(ix, fx);
  t.ey = __fsub_rn(__fadd_rn(fy, 1.0f), iy); t.wy = __fsub_rn(iy, fy);
  t.ez = __fsub_rn(__fadd_rn(fz, 1.0f), iz); t.wz = __fsub_rn(iz, fz);
  t.x0 = xi; t.y0 = yi; t.z0 = zi;
  return t;
}
__device__ __forceinline__ Trilin trilin_from_cell(const TrilinCell& c, int D, int H, int W) {
  Trilin t;
  t.ex = c.ex; t.wx = c.wx; t.ey = c.ey; t.wy = c.wy; t.ez = c.ez; t.wz = c.wz;
  // A +1 neighbour outside the volume only occurs when the clipped coordinate sits exactly on the last sample, where its
  // weight (wx / wy / wz) is exactly 0: ATen skips the tap, here it re-reads the in-range sample with weight 0.  That keeps
  // the 8 taps branch-free so that all gathers of a voxel are in flight together (one memory round trip, not four).
  t.dx = c.x0 + 1 <= W - 1 ? 1u : 0u;
  t.dy = c.y0 + 1 <= H - 1 ? (uint32_t)W : 0u;
  t.dz = c.z0 + 1 <= D - 1 ? (uint32_t)(H * W) : 0u;
  t.base = (uint32_t)((c.z0 * H + c.y0) * W + c.x0);
  return t;
}
__device__ __forceinline__ Trilin trilin_setup(float f0, float f1, float f2, float lh, float ld, float lw, int D, int H,
                                               int W, const float* hs, int ref_mode) {
  return trilin_from_cell(trilin_cell(f0, f1, f2, lh, ld, lw, D, H, W, hs, ref_mode), D, H, W);
}

template <bool FMA>
__device__ __forceinline__ float acc_tap(float acc, float v, float w) {
  return FMA ? __fmaf_rn(v, w, acc) : __fadd_rn(acc, __fmul_rn(v, w));
}

struct Taps8 {
  float v[8];
};
__device__ __forceinline__ Taps8 trilin_gather(const float* __restrict__ p, const Trilin& t) {
  const uint32_t i0 = t.base, i1 = i0 + t.dx, i2 = i0 + t.dy, i3 = i2 + t.dx;
  Taps8 r;
  r.v[0] = __ldg(p + i0); r.v[1] = __ldg(p + i1); r.v[2] = __ldg(p + i2); r.v[3] = __ldg(p + i3);
  r.v[4] = __ldg(p + (i0 + t.dz)); r.v[5] = __ldg(p + (i1 + t.dz)); r.v[6] = __ldg(p + (i2 + t.dz)); r.v[7] = __ldg(p + (i3 + t.dz));
  return r;
}
// ATen grid_sampler_3d corner order tnw,tne,tsw,tse,bnw,bne,bsw,bse; weights = product of 3 distances, left to right.
template <bool FMA>
__device__ __forceinline__ float trilin_reduce(const Taps8& r, const Trilin& t) {
  const float xy00 = __fmul_rn(t.ex, t.ey), xy10 = __fmul_rn(t.wx, t.ey), xy01 = __fmul_rn(t.ex, t.wy),
              xy11 = __fmul_rn(t.wx, t.wy);
  float acc = 0.0f;
  acc = acc_tap<FMA>(acc, r.v[0], __fmul_rn(xy00, t.ez));
  acc = acc_tap<FMA>(acc, r.v[1], __fmul_rn(xy10, t.ez));
  acc = acc_tap<FMA>(acc, r.v[2], __fmul_rn(xy01, t.ez));
  acc = acc_tap<FMA>(acc, r.v[3], __fmul_rn(xy11, t.ez));
  acc = acc_tap<FMA>(acc, r.v[4], __fmul_rn(xy00, t.wz));
  acc = acc_tap<FMA>(acc, r.v[5], __fmul_rn(xy10, t.wz));
  acc = acc_tap<FMA>(acc, r.v[6], __fmul_rn(xy01, t.wz));
  acc = acc_tap<FMA>(acc, r.v[7], __fmul_rn(xy11, t.wz));
  return acc;
}
template <bool FMA>
__device__ __forceinline__ float trilin_sample(const float* __restrict__ p, const Trilin& t, int W, int HW) {
  (void)W; (void)HW;
  const Taps8 r = trilin_gather(p, t);
  return trilin_reduce<FMA>(r, t);
}

// CTA tile of the 3-D kernels: 32 (h) x 8 (w) voxels at fixed (n, d); 256 threads = one voxel each, warp q owns the
// w column q with its lanes along h (coalesced 128 B source gathers).  Global I/O of the tile is 32 rows x 32 B (one
// full sector per row and plane) moved as float4 by the first P*64 threads; shared planes are padded to 9 floats.
constexpr int T3H = 32, T3W = 8, T3P = T3W + 1;

// planes[k] (k < NP) -> s[k][32][9]; a null plane pointer zero-fills.  `poff` = offset of the tile's (n,d) plane.
// `plane_ptr(k)` returns the global pointer of plane k's (n,d) slice, or nullptr to zero-fill / skip.
template <int NP, bool VEC, typename F>
__device__ __forceinline__ void load_planes(float (*s)[T3H][T3P], F plane_ptr, int h0, int w0, int H, int W) {
#pragma unroll
  for (int k = 0; k < (NP * 64 + 255) / 256; ++k) {
    const int idx = threadIdx.x + k * 256;
    if (idx < NP * 64) {
      const int pl = idx >> 6, row = (idx & 63) >> 1, c4 = (idx & 1) * 4;
      const int h = h0 + row, w = w0 + c4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* base = plane_ptr(pl);
      if (base != nullptr && h < H) {
        const float* g = base + (int64_t)h * W + w;
        if (VEC && w + 3 < W) {
          v = ldg_stream4(g);
        } else {
          if (w < W) v.x = ldg_stream(g);
          if (w + 1 < W) v.y = ldg_stream(g + 1);
          if (w + 2 < W) v.z = ldg_stream(g + 2);
          if (w + 3 < W) v.w = ldg_stream(g + 3);
        }
      }
      s[pl][row][c4] = v.x; s[pl][row][c4 + 1] = v.y; s[pl][row][c4 + 2] = v.z; s[pl][row][c4 + 3] = v.w;
    }
  }
}
template <int NP, bool VEC, typename F>
__device__ __forceinline__ void store_planes(const float (*s)[T3H][T3P], F plane_ptr, int h0, int w0, int H, int W) {
#pragma unroll
  for (int k = 0; k < (NP * 64 + 255) / 256; ++k) {
    const int idx = threadIdx.x + k * 256;
    if (idx < NP * 64) {
      const int pl = idx >> 6, row = (idx & 63) >> 1, c4 = (idx & 1) * 4;
      const int h = h0 + row, w = w0 + c4;
      float* base = plane_ptr(pl);
      if (base != nullptr && h < H) {
        float* g = base + (int64_t)h * W + w;
        if (VEC && w + 3 < W) {
          stg_stream4(g, make_float4(s[pl][row][c4], s[pl][row][c4 + 1], s[pl][row][c4 + 2], s[pl][row][c4 + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (w + i < W) g[i] = s[pl][row][c4 + i];
        }
      }
    }
  }
}

struct Warp3dParams {
  int N, C, D, H, W, ref_mode;
  float hs[6];  // (H-1)/2, (D-1)/2, (W-1)/2 and their fp32 reciprocals (computed in double like ATen)
};


static inline Warp3dParams make_warp3d_params(int N, int C, int D, int H, int W, int ref_mode) {
  Warp3dParams P;
  P.N = N; P.C = C; P.D = D; P.H = H; P.W = W; P.ref_mode = ref_mode;
  const double h0 = (H - 1.0) / 2.0, h1 = (D - 1.0) / 2.0, h2 = (W - 1.0) / 2.0;  // Flow-3D/model/warplayer.py:24-26
  P.hs[0] = (float)h0; P.hs[1] = (float)h1; P.hs[2] = (float)h2;
  P.hs[3] = (float)(1.0 / h0); P.hs[4] = (float)(1.0 / h1); P.hs[5] = (float)(1.0 / h2);
  return P;
}

}  // namespace ofsv
