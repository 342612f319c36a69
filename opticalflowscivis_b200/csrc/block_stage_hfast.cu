// Fused IFBlock output stage for 3-D volumes on the H-FASTEST channels-last flow/mask state fm[N][D][W][H][8] fp32
// (state_layout = OFSV_STATE_DWH8).  Same arithmetic, op for op, as block_stage.cu (Flow-3D/model/IFNet.py:118-119 resize and
// *scale, :169-170 accumulate, :186-191 sigmoid and warp x2, :242 blend, :82-90,166 the next block's resized concat).
//
// Why another layout: the reference warp rotates axes — output (d,h,w) samples source (z,y,x) ~ (w,d,h) — so
// the gathers are only coalesced with LANES ALONG h.  With the state stored [D][H][W][8] a lane-along-h thread reads its voxel
// 8 KB away from its neighbour's, and block_stage.cu therefore moves every state tile through shared memory twice (cp.async in,
// padded tile, a second thread mapping for the update): 450 of its ~840 instructions per voxel are integer address work.  With
// H as the fastest spatial axis of the STATE (the user-facing flow tensors are permuted views of it either way, and the head
// conv's depth-to-space epilogue writes whichever layout it is told) one thread owns one voxel for the whole stage:
//   prev state / full-resolution head : ONE 256-bit load each, a warp reads 1 KB contiguous
//   up-sampled head (scale 2, 4)      : the x (W) lerp of ATen's x-y-z order has warp-uniform weights and depends only on the
//                                       coarse h row, so lane L computes it once for coarse row L and the fine lanes fetch their
//                                       two rows by shuffle (4 loads + 7 lerps instead of 8 loads + 14 lerps, same rounding)
//   fm_out                            : ONE 256-bit store
//   warps                             : lanes along h, 16 gathers in flight (warp_device.cuh)
//   next block's input row (bf16)     : ONE 256-bit store (space-to-depth rows are 32 B sectors)
// Only the planar outputs (merged, sigmoid(mask)) and the 2x2x2 pooling still cross warps through a small shared tile.
#include "warp_device.cuh"

namespace ofsv {

constexpr int HF_H = 32, HF_W = 8;

struct V8 { float v[8]; };
__device__ __forceinline__ V8 ldg256(const float* p) {
  V8 r;
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ V8 ldg256_stream(const float* p) {
  V8 r;
  asm("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ldg_stream2(const float* p) {
  float2 v;
#if OFSV_HF_STREAM
  asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
#else
  v = __ldg(reinterpret_cast<const float2*>(p));
#endif
  return v;
}
__device__ __forceinline__ void stg256(float* p, const V8& r) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
               "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]));
}
__device__ __forceinline__ void stg256_b32(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g,
                                           uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h));
}
__device__ __forceinline__ uint32_t hf_pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Pure (non-volatile, no memory clobber) read-only loads: the compiler may hoist them over the kernel's stores — none of which is
// ever read back by the kernel — so that the gathers of the next voxel overlap the arithmetic and the stores of the current one.
__device__ __forceinline__ float hf_ldg(const float* p) {
  float v;
  asm("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ Taps8 hf_gather(const float* p, const Trilin& t) {
  const uint32_t i0 = t.base, i1 = i0 + t.dx, i2 = i0 + t.dy, i3 = i2 + t.dx;
  Taps8 r;
  r.v[0] = hf_ldg(p + i0); r.v[1] = hf_ldg(p + i1); r.v[2] = hf_ldg(p + i2); r.v[3] = hf_ldg(p + i3);
  r.v[4] = hf_ldg(p + (i0 + t.dz)); r.v[5] = hf_ldg(p + (i1 + t.dz)); r.v[6] = hf_ldg(p + (i2 + t.dz)); r.v[7] = hf_ldg(p + (i3 + t.dz));
  return r;
}
__device__ __forceinline__ void hf_st2(float* p, float2 v) {
  asm volatile("st.global.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y));
}

// U consecutive floats (U = 2, 4, 8) as one 8 / 16 / 32-byte access
template <int U>
struct VecF { float v[U]; };
template <int U>
__device__ __forceinline__ VecF<U> ldg_vec(const float* p) {
  VecF<U> r;
  if (U == 8) {
    const V8 t = ldg256_stream(p);
#pragma unroll
    for (int i = 0; i < U; ++i) r.v[i] = t.v[i];
  } else if (U == 4) {
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2 % U]), "=f"(r.v[3 % U]) : "l"(p));
  } else {
    const float2 t = ldg_stream2(p);
    r.v[0] = t.x; r.v[1] = t.y;
  }
  return r;
}
template <int U>
__device__ __forceinline__ void stg_vec(float* p, const VecF<U>& r) {
  if (U == 8) {
    V8 t;
#pragma unroll
    for (int i = 0; i < U; ++i) t.v[i] = r.v[i];
    stg256(p, t);
  } else if (U == 4) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2 % U]), "f"(r.v[3 % U]));
  } else {
    hf_st2(p, make_float2(r.v[0], r.v[1]));
  }
}

struct HfLerp {
  int i0, i1;
  float l0, l1;
};
// ATen area_pixel_compute_source_index (align_corners=False) + guard_index_and_lambda — identical to block_stage.cu / ifnet_glue.cu
__device__ __forceinline__ HfLerp hf_up_index(int dst, int n_in, float rscale) {
  float src = __fsub_rn(__fmul_rn(rscale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  src = src < 0.0f ? 0.0f : src;
  HfLerp L;
  L.i0 = min((int)src, n_in - 1);
  L.i1 = L.i0 + (L.i0 < n_in - 1 ? 1 : 0);
  L.l1 = fminf(fmaxf(__fsub_rn(src, (float)L.i0), 0.0f), 1.0f);
  L.l0 = __fsub_rn(1.0f, L.l1);
  return L;
}
// c0*l0 + c1*l1: the product c0*l0 rounded, then one FMA (the expression of head_upsample_add_kernel and block_stage.cu)
__device__ __forceinline__ V8 hf_lerp8(const V8& a, float la, const V8& b, float lb) {
  V8 r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = __fmaf_rn(b.v[i], lb, __fmul_rn(a.v[i], la));
  return r;
}
__device__ __forceinline__ V8 hf_shfl8(const V8& a, int src) {
  V8 r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
  return r;
}

struct HfPtrs {
  const float* head; const float* fm_prev; const float* img0; const float* img1;
  const float* lin_h; const float* lin_d; const float* lin_w;
  float* fm_out; float* merged; float* mask_sig; __nv_bfloat16* pack_out;
};

// Keeps a per-sample base pointer in ONE 64-bit register pair: without it nvcc re-adds the (uniform) sample offset to every tap
// address — IADD3 + IADD3.X + LEA + LEA.HI.X per gather instead of one IMAD.WIDE.U32 (48 of the final stage's 344 instructions).
template <typename T>
__device__ __forceinline__ T* hf_pin(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}

// SH: scale of the head (1, 2, 4), or 0 = fm_prev already holds the accumulated state (nothing is added, fm_out not written).
// SN: 0 = no packed output, 1 = next block at full resolution, 2 = at half resolution (2x2x2 mean).
//
// Work distribution.  A CTA owns a 32 (h) x 8 (w) column of HF_DZ = 16 planes; warp q owns the plane pair (2q, 2q+1) and a THREAD
// owns the 8 consecutive w of its (d, h): everything that is laid out [D][H][W] — the voxel's own image values going into the next
// block's input, merged, sigmoid(mask) — is then ONE 256-bit access per lane per plane (a full 32 B sector) instead of 8 scalar
// accesses spread over 8 warps, each of which costs an L1 wavefront per lane (the first version of this kernel spent 64 of its
// ~160 L1 wavefronts per 32 voxels on the own-image loads and needed a shared tile + block barrier per plane for the planar
// outputs: ncu l1tex throughput 83-86 % on the two up-sampling stages).  Walking along w also walks through consecutive source
// planes of the axis-rotating warp (z ~ w), whose z+1 taps are the next voxel's z taps.  The 2x2x2 mean of SN == 2 never leaves the
// warp: the w pair is two consecutive iterations, the h pair the neighbouring lane (one shuffle per channel), the d pair the warp's
// second plane.  No shared memory, no barrier.
#ifndef OFSV_HF_U_UP
#define OFSV_HF_U_UP 8        // columns of a thread processed as one unrolled group (stages that write a state / a packed input)
#endif
#ifndef OFSV_HF_U_FIN
#define OFSV_HF_U_FIN 8       // same, final stage (state in, merged + mask out)
#endif
#ifndef OFSV_HF_PP
#define OFSV_HF_PP 2          // planes per warp
#endif
#ifndef OFSV_HF_MINB_UP
#define OFSV_HF_MINB_UP 2     // resident CTAs per SM the register allocation aims at: stages that write a state / a packed input ...
#endif
#ifndef OFSV_HF_MINB_POOL
#define OFSV_HF_MINB_POOL 2   // ... the stage with the 2x2x2 mean ...
#endif
#ifndef OFSV_HF_MINB_FIN
#define OFSV_HF_MINB_FIN 3    // ... and the final stage (state in, merged + mask out)
#endif
#ifndef OFSV_HF_STREAM
#define OFSV_HF_STREAM 1      // 1: streaming loads (previous state, own image values) do not allocate in L1, which is left to the gathers
#endif
constexpr int HF_PP = OFSV_HF_PP;
constexpr int HF_DZ = 8 * HF_PP;   // planes per CTA: 8 warps x HF_PP
struct HfVox {
  V8 st;          // flow 0..5, mask logit, 0
  float a, b;     // warped img0 / img1
};

template <int SH, int SN, bool S2D, bool FMA>
__global__ void __launch_bounds__(256, (SH == 0 && SN == 0) ? OFSV_HF_MINB_FIN : (SN == 2 ? OFSV_HF_MINB_POOL : OFSV_HF_MINB_UP)) stage3d_hfast_kernel(const HfPtrs q, const Warp3dParams P) {
  const int H = P.H, W = P.W, D = P.D, HW = H * W;
  const int V = D * HW;                                             // < 2^28 (host check)
  const int nzb = (D + HF_DZ - 1) / HF_DZ;
  const int n = blockIdx.z / nzb, dbeg = (blockIdx.z - n * nzb) * HF_DZ;
  const int h0 = blockIdx.y * HF_H, w0 = blockIdx.x * HF_W;         // W % 8 == 0 (host check): the 8 columns of a thread all exist
  const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
  const int h = h0 + lane;
  const bool okh = h < H;
  const int hc = min(h, H - 1);                                     // clamped: out-of-tile lanes compute on a valid voxel and store nothing
  constexpr int SHD = SH > 1 ? SH : 1;
  constexpr int NROWS = SH > 1 ? HF_H / SHD + 2 : 0;                // coarse head rows a 32-row tile can touch
  const int Dh = D / SHD, Hh = H / SHD, Wh = W / SHD;
  const float* hb = SH ? hf_pin(q.head + (int64_t)n * Dh * Hh * Wh * 8) : nullptr;
  const float* fprev = q.fm_prev ? hf_pin(q.fm_prev + (int64_t)n * V * 8) : nullptr;
  float* fout = SH ? hf_pin(q.fm_out + (int64_t)n * V * 8) : nullptr;
  const float* i0p = hf_pin(q.img0 + (int64_t)n * V);
  const float* i1p = hf_pin(q.img1 + (int64_t)n * V);
  const bool has_prev = fprev != nullptr;
  const bool need_m = q.merged != nullptr || q.mask_sig != nullptr;
  const float lh = __ldg(q.lin_h + hc);
  const int d0 = dbeg + HF_PP * wq;                                 // this warp's plane group
  if (d0 >= D) return;                                              // warp-uniform; nothing below synchronises across warps

  // up-sampled head: per-axis taps.  w (x) and d (z) taps are warp-uniform; the h (y) taps are per lane.  Lane L < NROWS also OWNS
  // coarse row hb0 + L of the tile: it evaluates the x lerp of that row for both z taps, the fine lanes pick rows (y0 - hb0, y1 - hb0).
  HfLerp Ly{0, 0, 0.f, 0.f};
  int hb0 = 0, own = 0;
  if (SH > 1) {
    const float rs = 1.0f / (float)SHD;
    Ly = hf_up_index(hc, Hh, rs);
    hb0 = hf_up_index(h0, Hh, rs).i0;
    own = min(hb0 + lane, Hh - 1);
  }
  const int s0 = Ly.i0 - hb0, s1 = Ly.i1 - hb0;                     // owner lanes of this lane's two coarse rows (< NROWS)

  // one voxel (d, hc, w): state update + the two warps.  `cur` = its previous state (prefetched by the caller).
  auto voxel = [&](int d, int w, float ld, const HfLerp& Lz, const V8& cur) -> HfVox {
    HfVox o;
    const int so = w * H + hc;
    const float lw = hf_ldg(q.lin_w + w);
    if (SH != 0) {
      V8 hv;
      if (SH == 1) {
        hv = ldg256(hb + ((int64_t)d * HW + so) * 8);
      } else {
        const HfLerp Lx = hf_up_index(w, Wh, 1.0f / (float)SHD);
        V8 a0, a1;
#pragma unroll
        for (int i = 0; i < 8; ++i) { a0.v[i] = 0.f; a1.v[i] = 0.f; }
        if (lane < NROWS) {
          // x lerp of coarse row `own` for the two z taps (ATen order: x, then y, then z)
          const int c0 = Lx.i0 * Hh + own, c1 = Lx.i1 * Hh + own;
          const int r0 = Lz.i0 * Wh * Hh, r1 = Lz.i1 * Wh * Hh;
          a0 = hf_lerp8(ldg256(hb + (int64_t)(r0 + c0) * 8), Lx.l0, ldg256(hb + (int64_t)(r0 + c1) * 8), Lx.l1);
          a1 = hf_lerp8(ldg256(hb + (int64_t)(r1 + c0) * 8), Lx.l0, ldg256(hb + (int64_t)(r1 + c1) * 8), Lx.l1);
        }
        const V8 b0 = hf_lerp8(hf_shfl8(a0, s0), Ly.l0, hf_shfl8(a0, s1), Ly.l1);
        const V8 b1 = hf_lerp8(hf_shfl8(a1, s0), Ly.l0, hf_shfl8(a1, s1), Ly.l1);
        hv = hf_lerp8(b0, Lz.l0, b1, Lz.l1);
      }
      const float sh = (float)SHD;
      if (has_prev) {
#pragma unroll
        for (int i = 0; i < 6; ++i) o.st.v[i] = __fadd_rn(cur.v[i], __fmul_rn(hv.v[i], sh));
        o.st.v[6] = __fadd_rn(cur.v[6], hv.v[6]);
      } else {                                                      // block 0: flow = flow_d exactly
#pragma unroll
        for (int i = 0; i < 6; ++i) o.st.v[i] = __fmul_rn(hv.v[i], sh);
        o.st.v[6] = hv.v[6];
      }
      o.st.v[7] = 0.0f;
      if (okh) stg256(fout + ((int64_t)d * HW + so) * 8, o.st);
    } else {
      o.st = cur;
    }
    const Trilin t0 = trilin_setup(o.st.v[0], o.st.v[1], o.st.v[2], lh, ld, lw, D, H, W, P.hs, P.ref_mode);
    const Trilin t1 = trilin_setup(o.st.v[3], o.st.v[4], o.st.v[5], lh, ld, lw, D, H, W, P.hs, P.ref_mode);
    const Taps8 g0 = hf_gather(i0p, t0), g1 = hf_gather(i1p, t1);       // 16 independent loads in flight
    o.a = trilin_reduce<FMA>(g0, t0);
    o.b = trilin_reduce<FMA>(g1, t1);
    return o;
  };
  auto state_at = [&](int d, int w) -> V8 {
    const float* p = fprev + ((int64_t)d * HW + w * H + hc) * 8;
    return OFSV_HF_STREAM ? ldg256_stream(p) : ldg256(p);
  };
  auto blend = [&](const HfVox& o, float& mg, float& ms) {
    ms = sigmoidf_ref(o.st.v[6]);
    mg = __fadd_rn(__fmul_rn(o.a, ms), __fmul_rn(o.b, __fsub_rn(1.0f, ms)));
  };

  if (SN != 2) {
    // ------------------------------------------------------------------ plane by plane, U columns of the thread as one unrolled group
    constexpr int U = (SH == 0 && SN == 0) ? OFSV_HF_U_FIN : OFSV_HF_U_UP;
    V8 pv;
    if (has_prev) pv = state_at(d0, w0);
#pragma unroll 1
    for (int p = 0; p < HF_PP; ++p) {
      const int d = d0 + p;                                         // D % HF_PP == 0 (host check)
      const float ld = __ldg(q.lin_d + d);
      const HfLerp Lz = SH > 1 ? hf_up_index(d, Dh, 1.0f / (float)SHD) : HfLerp{0, 0, 0.f, 0.f};
#pragma unroll 1
      for (int kc = 0; kc < HF_W; kc += U) {
        const int64_t prow = (int64_t)d * HW + hc * W + w0 + kc;    // U of this thread's voxels in the [D][H][W] planes (4U-byte aligned)
        VecF<U> o0, o1, mg, ms;
        if (SN == 1) { o0 = ldg_vec<U>(i0p + prow); o1 = ldg_vec<U>(i1p + prow); }
#pragma unroll
        for (int ku = 0; ku < U; ++ku) {
          const int k = kc + ku, w = w0 + k;
          const V8 cur = pv;
          const bool more = k + 1 < HF_W || p + 1 < HF_PP;
          if (has_prev && more) pv = (k + 1 < HF_W) ? state_at(d, w + 1) : state_at(d + 1, w0);
          const HfVox o = voxel(d, w, ld, Lz, cur);
          if (need_m) blend(o, mg.v[ku], ms.v[ku]);
          if (SN == 1 && okh) {
            int64_t ro;
            if (S2D) ro = s2d_row(3, n, d, h, w, D, H, W) * 16;
            else ro = ((((int64_t)n * D + d) * H + h) * W + w) * 16;
            stg256_b32(q.pack_out + ro, hf_pack2(o0.v[ku], o1.v[ku]), hf_pack2(o.a, o.b), hf_pack2(o.st.v[6], o.st.v[0]),
                       hf_pack2(o.st.v[1], o.st.v[2]), hf_pack2(o.st.v[3], o.st.v[4]), hf_pack2(o.st.v[5], 0.0f), 0u, 0u);
          }
        }
        if (need_m && okh) {
          if (q.merged) stg_vec<U>(q.merged + (int64_t)n * V + prow, mg);
          if (q.mask_sig) stg_vec<U>(q.mask_sig + (int64_t)n * V + prow, ms);
        }
      }
    }
    return;
  }
  // Order: w pair -> the PP planes of the warp -> the two columns of the pair.  Neighbouring planes and columns share source rows
  // of the axis-rotating warp (voxel (d, w) taps rows (z, y) = (w..w+1, d..d+1)): walking w with the planes innermost touches
  // PP + 1 new 128-byte row segments per image for PP voxels and needs only the last few iterations to still be in L1.
  // The 2x2x2 mean of SN == 2 (== F.interpolate(., 0.5)) nests W, H, D like ATen, each level 0.5 * a + 0.5 * b, lower index first.
  constexpr int PP = HF_PP;
  float ldv[PP];
  HfLerp Lzv[PP];
#pragma unroll
  for (int p = 0; p < PP; ++p) {
    ldv[p] = __ldg(q.lin_d + d0 + p);
    Lzv[p] = SH > 1 ? hf_up_index(d0 + p, Dh, 1.0f / (float)SHD) : HfLerp{0, 0, 0.f, 0.f};
  }
  V8 pv;
  if (has_prev) pv = state_at(d0, w0);
#pragma unroll 1
  for (int kp = 0; kp < HF_W / 2; ++kp) {
    float pd[11];
#pragma unroll
    for (int p = 0; p < PP; ++p) {
      const int d = d0 + p;                                         // D % PP == 0 (host check): every plane of the group exists
      const int64_t prow = (int64_t)d * HW + hc * W + w0 + 2 * kp;  // this thread's column pair in the [D][H][W] planes (8 B aligned)
      float2 o0 = make_float2(0.f, 0.f), o1 = o0;
      if (SN != 0) { o0 = ldg_stream2(i0p + prow); o1 = ldg_stream2(i1p + prow); }
      float pw[11];
      float2 mg, ms;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int w = w0 + 2 * kp + kk;
        const V8 cur = pv;
        const bool more = !(kk == 1 && p == PP - 1 && kp == HF_W / 2 - 1);
        if (has_prev && more) pv = kk == 0 ? state_at(d, w + 1) : (p < PP - 1 ? state_at(d + 1, w - 1) : state_at(d0, w + 1));
        const HfVox o = voxel(d, w, ldv[p], Lzv[p], cur);
        if (need_m) blend(o, kk ? mg.y : mg.x, kk ? ms.y : ms.x);
        if (SN == 1 && okh) {
          int64_t ro;
          if (S2D) ro = s2d_row(3, n, d, h, w, D, H, W) * 16;
          else ro = ((((int64_t)n * D + d) * H + h) * W + w) * 16;
          stg256_b32(q.pack_out + ro, hf_pack2(kk ? o0.y : o0.x, kk ? o1.y : o1.x), hf_pack2(o.a, o.b), hf_pack2(o.st.v[6], o.st.v[0]),
                     hf_pack2(o.st.v[1], o.st.v[2]), hf_pack2(o.st.v[3], o.st.v[4]), hf_pack2(o.st.v[5], 0.0f), 0u, 0u);
        }
        if (SN == 2) {
          const float c11[11] = {kk ? o0.y : o0.x, kk ? o1.y : o1.x, o.a, o.b, o.st.v[6], o.st.v[0], o.st.v[1], o.st.v[2], o.st.v[3],
                                 o.st.v[4], o.st.v[5]};
#pragma unroll
          for (int c = 0; c < 11; ++c) pw[c] = kk == 0 ? __fmul_rn(c11[c], 0.5f) : __fadd_rn(pw[c], __fmul_rn(c11[c], 0.5f));
        }
      }
      if (need_m && okh) {
        if (q.merged) hf_st2(q.merged + (int64_t)n * V + prow, mg);
        if (q.mask_sig) hf_st2(q.mask_sig + (int64_t)n * V + prow, ms);
      }
      if (SN == 2) {
#pragma unroll
        for (int c = 0; c < 11; ++c) {
          const float up = __shfl_down_sync(0xffffffffu, pw[c], 1); // row h + 1 (used by the even lanes only)
          const float ph = __fadd_rn(__fmul_rn(pw[c], 0.5f), __fmul_rn(up, 0.5f));
          if ((p & 1) == 0) pd[c] = ph;
          else pd[c] = __fadd_rn(__fmul_rn(pd[c], 0.5f), __fmul_rn(ph, 0.5f));
        }
        if ((p & 1) && !(lane & 1) && okh) {
          const int oh = h >> 1, ow = (w0 >> 1) + kp, od = d >> 1;
#pragma unroll
          for (int c = 5; c < 11; ++c) pd[c] = __fmul_rn(pd[c], 0.5f);             // flow channels additionally * 0.5
          int64_t ro;
          if (S2D) ro = s2d_row(3, n, od, oh, ow, D / 2, H / 2, W / 2) * 16;
          else ro = ((((int64_t)n * (D / 2) + od) * (H / 2) + oh) * (W / 2) + ow) * 16;
          stg256_b32(q.pack_out + ro, hf_pack2(pd[0], pd[1]), hf_pack2(pd[2], pd[3]), hf_pack2(pd[4], pd[5]), hf_pack2(pd[6], pd[7]),
                     hf_pack2(pd[8], pd[9]), hf_pack2(pd[10], 0.0f), 0u, 0u);
        }
      }
    }
  }
}

// Column variant for the stages that write the next block's input at FULL resolution (SN == 1: one 32-byte packed row per voxel,
// nothing to vectorise along w): warp q owns ONE w column of a 32 (h) x 8 (w) tile and walks HF_CDZ planes, so everything that depends
// on w only (x taps of the head interpolation, linspace entry, plane offsets) is hoisted out of the loop.  Measured against the
// row variant above on the block1 -> block2 stage of a 4 x 256^3 batch: 1773 us vs 1840-2150 us (any unroll / occupancy choice).
// merged / sigmoid(mask), when requested (IFNet.forward with all three blends), go through a shared tile + one barrier per plane.
constexpr int HF_CDZ = 8;
#ifndef OFSV_HF_MINB_COLS
#define OFSV_HF_MINB_COLS 3
#endif
template <int SH, bool S2D, bool FMA>
__global__ void __launch_bounds__(256, OFSV_HF_MINB_COLS) stage3d_hfast_cols_kernel(const HfPtrs q, const Warp3dParams P) {
  __shared__ float s_out[2][2][HF_H][HF_W + 1];                     // [buffer][merged | mask][h][w]
  const int H = P.H, W = P.W, D = P.D, HW = H * W;
  const int V = D * HW;
  const int nzb = (D + HF_CDZ - 1) / HF_CDZ;
  const int n = blockIdx.z / nzb, dbeg = (blockIdx.z - n * nzb) * HF_CDZ;
  const int nplanes = min(HF_CDZ, D - dbeg);
  const int h0 = blockIdx.y * HF_H, w0 = blockIdx.x * HF_W;
  const int tid = threadIdx.x, lane = tid & 31, wl = tid >> 5;
  const int h = h0 + lane, w = w0 + wl;                             // W % 8 == 0 (host check): w < W
  const bool okh = h < H;
  const int hc = min(h, H - 1);
  const int rP = tid >> 3, cP = tid & 7;                            // planar-output mapping: 8 consecutive threads = one tile row
  const bool okP = (h0 + rP) < H;
  constexpr int SHD = SH > 1 ? SH : 1;
  constexpr int NROWS = SH > 1 ? HF_H / SHD + 2 : 0;
  const int Dh = D / SHD, Hh = H / SHD, Wh = W / SHD;
  const float* hb = hf_pin(q.head + (int64_t)n * Dh * Hh * Wh * 8);
  const float* fprev = q.fm_prev ? hf_pin(q.fm_prev + (int64_t)n * V * 8) : nullptr;
  float* fout = hf_pin(q.fm_out + (int64_t)n * V * 8);
  const float* i0p = hf_pin(q.img0 + (int64_t)n * V);
  const float* i1p = hf_pin(q.img1 + (int64_t)n * V);
  const bool has_prev = fprev != nullptr;
  const bool need_m = q.merged != nullptr || q.mask_sig != nullptr;
  const float lh = __ldg(q.lin_h + hc), lw = __ldg(q.lin_w + w);
  const int so = w * H + hc, io = hc * W + w;                       // in-plane offsets: [W][H] state planes, [H][W] image planes
  HfLerp Ly{0, 0, 0.f, 0.f}, Lx{0, 0, 0.f, 0.f};
  int hb0 = 0, own = 0;
  if (SH > 1) {
    const float rs = 1.0f / (float)SHD;
    Ly = hf_up_index(hc, Hh, rs);
    Lx = hf_up_index(w, Wh, rs);
    hb0 = hf_up_index(h0, Hh, rs).i0;
    own = min(hb0 + lane, Hh - 1);
  }
  const int s0 = Ly.i0 - hb0, s1 = Ly.i1 - hb0;
  const int c0 = Lx.i0 * Hh + own, c1 = Lx.i1 * Hh + own;
  V8 pv;
  if (has_prev) pv = ldg256(fprev + ((int64_t)dbeg * HW + so) * 8);
  float pi0 = hf_ldg(i0p + dbeg * HW + io), pi1 = hf_ldg(i1p + dbeg * HW + io);     // own image values, one plane ahead
  for (int it = 0; it < nplanes; ++it) {
    const int d = dbeg + it;
    const float ld = __ldg(q.lin_d + d);
    const V8 cur = pv;
    if (has_prev && it + 1 < nplanes) pv = ldg256(fprev + ((int64_t)(d + 1) * HW + so) * 8);
    V8 hv, st;
    if (SH == 1) {
      hv = ldg256(hb + ((int64_t)d * HW + so) * 8);
    } else {
      const HfLerp Lz = hf_up_index(d, Dh, 1.0f / (float)SHD);
      V8 a0, a1;
#pragma unroll
      for (int i = 0; i < 8; ++i) { a0.v[i] = 0.f; a1.v[i] = 0.f; }
      if (lane < NROWS) {                                           // x lerp of coarse row `own` for the two z taps (ATen order: x, y, z)
        const int r0 = Lz.i0 * Wh * Hh, r1 = Lz.i1 * Wh * Hh;
        a0 = hf_lerp8(ldg256(hb + (int64_t)(r0 + c0) * 8), Lx.l0, ldg256(hb + (int64_t)(r0 + c1) * 8), Lx.l1);
        a1 = hf_lerp8(ldg256(hb + (int64_t)(r1 + c0) * 8), Lx.l0, ldg256(hb + (int64_t)(r1 + c1) * 8), Lx.l1);
      }
      const V8 b0 = hf_lerp8(hf_shfl8(a0, s0), Ly.l0, hf_shfl8(a0, s1), Ly.l1);
      const V8 b1 = hf_lerp8(hf_shfl8(a1, s0), Ly.l0, hf_shfl8(a1, s1), Ly.l1);
      hv = hf_lerp8(b0, Lz.l0, b1, Lz.l1);
    }
    const float sh = (float)SHD;
    if (has_prev) {
#pragma unroll
      for (int i = 0; i < 6; ++i) st.v[i] = __fadd_rn(cur.v[i], __fmul_rn(hv.v[i], sh));
      st.v[6] = __fadd_rn(cur.v[6], hv.v[6]);
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) st.v[i] = __fmul_rn(hv.v[i], sh);
      st.v[6] = hv.v[6];
    }
    st.v[7] = 0.0f;
    if (okh) stg256(fout + ((int64_t)d * HW + so) * 8, st);
    const float m = st.v[6];
    const Trilin t0 = trilin_setup(st.v[0], st.v[1], st.v[2], lh, ld, lw, D, H, W, P.hs, P.ref_mode);
    const Trilin t1 = trilin_setup(st.v[3], st.v[4], st.v[5], lh, ld, lw, D, H, W, P.hs, P.ref_mode);
    const Taps8 g0 = hf_gather(i0p, t0), g1 = hf_gather(i1p, t1);
    const float i0v = pi0, i1v = pi1;
    if (it + 1 < nplanes) { pi0 = hf_ldg(i0p + (d + 1) * HW + io); pi1 = hf_ldg(i1p + (d + 1) * HW + io); }
    const float a = trilin_reduce<FMA>(g0, t0), b = trilin_reduce<FMA>(g1, t1);
    const int ob = it & 1;
    if (need_m) {
      const float ms = sigmoidf_ref(m);
      s_out[ob][0][lane][wl] = __fadd_rn(__fmul_rn(a, ms), __fmul_rn(b, __fsub_rn(1.0f, ms)));
      s_out[ob][1][lane][wl] = ms;
    }
    if (okh) {
      int64_t ro;
      if (S2D) ro = s2d_row(3, n, d, h, w, D, H, W) * 16;
      else ro = ((((int64_t)n * D + d) * H + h) * W + w) * 16;
      stg256_b32(q.pack_out + ro, hf_pack2(i0v, i1v), hf_pack2(a, b), hf_pack2(m, st.v[0]), hf_pack2(st.v[1], st.v[2]),
                 hf_pack2(st.v[3], st.v[4]), hf_pack2(st.v[5], 0.0f), 0u, 0u);
    }
    if (need_m) {
      __syncthreads();                                              // s_out is double-buffered by plane parity: one barrier per plane
      if (okP) {
        const int64_t g = (int64_t)n * V + (int64_t)d * HW + (h0 + rP) * W + w0 + cP;
        if (q.merged) q.merged[g] = s_out[ob][0][rP][cP];
        if (q.mask_sig) q.mask_sig[g] = s_out[ob][1][rP][cP];
      }
    }
  }
}

// Column variant with CPT (2 or 4) ADJACENT columns per thread: the warp still walks the planes of its columns (the L1-friendly order:
// a plane step of the CTA touches the source rows (z, y) = (w0 .. w0 + 8 CPT, d .. d + 1), half of which the previous step loaded),
// but the voxel's own image values are ONE 8 / 16-byte load per thread and plane instead of CPT strided 4-byte loads (a lane-along-h
// warp touches 32 lines per request either way), the CPT voxels of a plane are independent chains for the scheduler to interleave,
// and merged / sigmoid(mask) leave as 8 / 16-byte stores without the shared tile and its barrier.
#ifndef OFSV_HF_CPT
#define OFSV_HF_CPT 2
#endif
#ifndef OFSV_HF_MINB_CPT
#define OFSV_HF_MINB_CPT 3
#endif
template <int SH, int SN, bool S2D, bool FMA, int CPT>
__global__ void __launch_bounds__(256, OFSV_HF_MINB_CPT) stage3d_hfast_colsn_kernel(const HfPtrs q, const Warp3dParams P) {
  constexpr int TW = HF_W * CPT;
  const int H = P.H, W = P.W, D = P.D, HW = H * W;
  const int V = D * HW;
  const int nzb = (D + HF_CDZ - 1) / HF_CDZ;
  const int n = blockIdx.z / nzb, dbeg = (blockIdx.z - n * nzb) * HF_CDZ;
  const int nplanes = min(HF_CDZ, D - dbeg);
  const int h0 = blockIdx.y * HF_H, w0 = blockIdx.x * TW;
  const int tid = threadIdx.x, lane = tid & 31, wl = tid >> 5;
  const int h = h0 + lane, wb = w0 + wl * CPT;                      // this thread's columns wb .. wb + CPT - 1 (W % 8 == 0: all in or all out)
  if (wb >= W) return;                                              // warp-uniform; no block-wide synchronisation below
  const bool okh = h < H;
  const int hc = min(h, H - 1);
  constexpr int SHD = SH > 1 ? SH : 1;
  constexpr int NROWS = SH > 1 ? HF_H / SHD + 2 : 0;
  const int Dh = D / SHD, Hh = H / SHD, Wh = W / SHD;
  const float* hb = SH ? hf_pin(q.head + (int64_t)n * Dh * Hh * Wh * 8) : nullptr;
  const float* fprev = q.fm_prev ? hf_pin(q.fm_prev + (int64_t)n * V * 8) : nullptr;
  float* fout = SH ? hf_pin(q.fm_out + (int64_t)n * V * 8) : nullptr;
  const float* i0p = hf_pin(q.img0 + (int64_t)n * V);
  const float* i1p = hf_pin(q.img1 + (int64_t)n * V);
  const bool has_prev = fprev != nullptr;
  const bool need_m = q.merged != nullptr || q.mask_sig != nullptr;
  const float lh = __ldg(q.lin_h + hc);
  const int io = hc * W + wb;                                       // [H][W] image planes: the CPT columns are consecutive floats
  HfLerp Ly{0, 0, 0.f, 0.f};
  int hb0 = 0, own = 0;
  if (SH > 1) {
    const float rs = 1.0f / (float)SHD;
    Ly = hf_up_index(hc, Hh, rs);
    hb0 = hf_up_index(h0, Hh, rs).i0;
    own = min(hb0 + lane, Hh - 1);
  }
  const int s0 = Ly.i0 - hb0, s1 = Ly.i1 - hb0;
  float lw[CPT], xl0[CPT], xl1[CPT];
  int so[CPT], c0[CPT], c1[CPT];
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    lw[k] = __ldg(q.lin_w + wb + k);
    so[k] = (wb + k) * H + hc;                                      // [W][H] state planes
    const HfLerp Lx = SH > 1 ? hf_up_index(wb + k, Wh, 1.0f / (float)SHD) : HfLerp{0, 0, 0.f, 0.f};
    c0[k] = Lx.i0 * Hh + own; c1[k] = Lx.i1 * Hh + own; xl0[k] = Lx.l0; xl1[k] = Lx.l1;
  }
  float pd[11];                                                     // SN == 2: (w, h)-pair sums of the even plane
#pragma unroll
  for (int c = 0; c < 11; ++c) pd[c] = 0.f;
  V8 pv;
  if (has_prev) pv = ldg256(fprev + ((int64_t)dbeg * HW + so[0]) * 8);
  VecF<CPT> pi0, pi1;                                                // own image values, one plane ahead
  if (SN != 0) { pi0 = ldg_vec<CPT>(i0p + dbeg * HW + io); pi1 = ldg_vec<CPT>(i1p + dbeg * HW + io); }
  for (int it = 0; it < nplanes; ++it) {
    const int d = dbeg + it;
    const float ld = __ldg(q.lin_d + d);
    const HfLerp Lz = SH > 1 ? hf_up_index(d, Dh, 1.0f / (float)SHD) : HfLerp{0, 0, 0.f, 0.f};
    const int r0 = Lz.i0 * Wh * Hh, r1 = Lz.i1 * Wh * Hh;
    const VecF<CPT> o0 = pi0, o1 = pi1;
    if (SN != 0 && it + 1 < nplanes) { pi0 = ldg_vec<CPT>(i0p + (d + 1) * HW + io); pi1 = ldg_vec<CPT>(i1p + (d + 1) * HW + io); }
    VecF<CPT> mg, ms;
    float pw[11];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const V8 cur = pv;
      const bool more = k + 1 < CPT || it + 1 < nplanes;
      if (has_prev && more) pv = ldg256(fprev + ((int64_t)(k + 1 < CPT ? d : d + 1) * HW + so[k + 1 < CPT ? k + 1 : 0]) * 8);
      V8 hv, st;
      if (SH == 0) {
        st = cur;
      } else if (SH == 1) {
        hv = ldg256(hb + ((int64_t)d * HW + so[k]) * 8);
      } else {
        V8 a0, a1;
#pragma unroll
        for (int i = 0; i < 8; ++i) { a0.v[i] = 0.f; a1.v[i] = 0.f; }
        if (lane < NROWS) {                                         // x lerp of coarse row `own` for the two z taps (ATen order: x, y, z)
          a0 = hf_lerp8(ldg256(hb + (int64_t)(r0 + c0[k]) * 8), xl0[k], ldg256(hb + (int64_t)(r0 + c1[k]) * 8), xl1[k]);
          a1 = hf_lerp8(ldg256(hb + (int64_t)(r1 + c0[k]) * 8), xl0[k], ldg256(hb + (int64_t)(r1 + c1[k]) * 8), xl1[k]);
        }
        const V8 b0 = hf_lerp8(hf_shfl8(a0, s0), Ly.l0, hf_shfl8(a0, s1), Ly.l1);
        const V8 b1 = hf_lerp8(hf_shfl8(a1, s0), Ly.l0, hf_shfl8(a1, s1), Ly.l1);
        hv = hf_lerp8(b0, Lz.l0, b1, Lz.l1);
      }
      if (SH != 0) {
        const float sh = (float)SHD;
        if (has_prev) {
#pragma unroll
          for (int i = 0; i < 6; ++i) st.v[i] = __fadd_rn(cur.v[i], __fmul_rn(hv.v[i], sh));
          st.v[6] = __fadd_rn(cur.v[6], hv.v[6]);
        } else {
#pragma unroll
          for (int i = 0; i < 6; ++i) st.v[i] = __fmul_rn(hv.v[i], sh);
          st.v[6] = hv.v[6];
        }
        st.v[7] = 0.0f;
        if (okh) stg256(fout + ((int64_t)d * HW + so[k]) * 8, st);
      }
      const float m = st.v[6];
      const Trilin t0 = trilin_setup(st.v[0], st.v[1], st.v[2], lh, ld, lw[k], D, H, W, P.hs, P.ref_mode);
      const Trilin t1 = trilin_setup(st.v[3], st.v[4], st.v[5], lh, ld, lw[k], D, H, W, P.hs, P.ref_mode);
      const Taps8 g0 = hf_gather(i0p, t0), g1 = hf_gather(i1p, t1);
      const float a = trilin_reduce<FMA>(g0, t0), b = trilin_reduce<FMA>(g1, t1);
      if (need_m) {
        ms.v[k] = sigmoidf_ref(m);
        mg.v[k] = __fadd_rn(__fmul_rn(a, ms.v[k]), __fmul_rn(b, __fsub_rn(1.0f, ms.v[k])));
      }
      if (SN == 2) {                                                // 2x2x2 mean, W level: CPT == 2, the pair is this thread's two columns
        const float c11[11] = {o0.v[k], o1.v[k], a, b, m, st.v[0], st.v[1], st.v[2], st.v[3], st.v[4], st.v[5]};
#pragma unroll
        for (int c = 0; c < 11; ++c) pw[c] = k == 0 ? __fmul_rn(c11[c], 0.5f) : __fadd_rn(pw[c], __fmul_rn(c11[c], 0.5f));
      }
      if (SN == 1 && okh) {
        int64_t ro;
        if (S2D) ro = s2d_row(3, n, d, h, wb + k, D, H, W) * 16;
        else ro = ((((int64_t)n * D + d) * H + h) * W + wb + k) * 16;
        stg256_b32(q.pack_out + ro, hf_pack2(o0.v[k], o1.v[k]), hf_pack2(a, b), hf_pack2(m, st.v[0]), hf_pack2(st.v[1], st.v[2]),
                   hf_pack2(st.v[3], st.v[4]), hf_pack2(st.v[5], 0.0f), 0u, 0u);
      }
    }
    if (need_m && okh) {
      const int64_t g = (int64_t)n * V + (int64_t)d * HW + io;
      if (q.merged) stg_vec<CPT>(q.merged + g, mg);
      if (q.mask_sig) stg_vec<CPT>(q.mask_sig + g, ms);
    }
    if (SN == 2) {                                                  // H level: the neighbouring lane; D level: the previous (even) plane
      static_assert(SN != 2 || CPT == 2, "the pooled output needs the w pair in one thread");
#pragma unroll
      for (int c = 0; c < 11; ++c) {
        const float up = __shfl_down_sync(0xffffffffu, pw[c], 1);
        const float ph = __fadd_rn(__fmul_rn(pw[c], 0.5f), __fmul_rn(up, 0.5f));
        pd[c] = (it & 1) ? __fadd_rn(__fmul_rn(pd[c], 0.5f), __fmul_rn(ph, 0.5f)) : ph;
      }
      if ((it & 1) && !(lane & 1) && okh) {
        float r[11];
#pragma unroll
        for (int c = 0; c < 11; ++c) r[c] = c >= 5 ? __fmul_rn(pd[c], 0.5f) : pd[c];     // flow channels additionally * 0.5
        const int oh = h >> 1, ow = wb >> 1, od = d >> 1;
        int64_t ro;
        if (S2D) ro = s2d_row(3, n, od, oh, ow, D / 2, H / 2, W / 2) * 16;
        else ro = ((((int64_t)n * (D / 2) + od) * (H / 2) + oh) * (W / 2) + ow) * 16;
        stg256_b32(q.pack_out + ro, hf_pack2(r[0], r[1]), hf_pack2(r[2], r[3]), hf_pack2(r[4], r[5]), hf_pack2(r[6], r[7]),
                   hf_pack2(r[8], r[9]), hf_pack2(r[10], 0.0f), 0u, 0u);
      }
    }
  }
}

#ifndef OFSV_HF_POOL_COLS
#define OFSV_HF_POOL_COLS 0     // 1: the stage with the 2x2x2 mean runs the column kernel too
#endif
#ifndef OFSV_HF_FIN_COLS
#define OFSV_HF_FIN_COLS 0      // 1: the stages without a packed output also run the column kernel
#endif
template <int SH, int SN, bool S2D, bool FMA>
static int launch_hfast(const HfPtrs& q, const Warp3dParams& P, cudaStream_t st) {
  if constexpr ((SN == 1 && SH != 0) || (SN == 0 && OFSV_HF_FIN_COLS) || (SN == 2 && OFSV_HF_POOL_COLS && OFSV_HF_CPT == 2)) {
    constexpr int CPT = OFSV_HF_CPT > 1 ? OFSV_HF_CPT : 2;
    const dim3 grid((unsigned)cdiv(P.W, HF_W * (OFSV_HF_CPT > 1 ? CPT : 1)), (unsigned)cdiv(P.H, HF_H), (unsigned)(P.N * cdiv(P.D, HF_CDZ)));
    if (grid.z > 65535u) { set_error("ofsv_block_stage_3d: N*D=%u exceeds grid.z", grid.z); return OFSV_ENOSUP; }
    if (OFSV_HF_CPT > 1 || SN != 1) {
      stage3d_hfast_colsn_kernel<SH, SN, S2D, FMA, CPT><<<grid, 256, 0, st>>>(q, P);
      return check_launch("stage3d_hfast_colsn_kernel");
    }
    if constexpr (SN == 1) {
      stage3d_hfast_cols_kernel<SH, S2D, FMA><<<grid, 256, 0, st>>>(q, P);
      return check_launch("stage3d_hfast_cols_kernel");
    }
    return OFSV_EINVAL;
  } else {
    const dim3 grid((unsigned)cdiv(P.W, HF_W), (unsigned)cdiv(P.H, HF_H), (unsigned)(P.N * cdiv(P.D, HF_DZ)));
    if (grid.z > 65535u) { set_error("ofsv_block_stage_3d: N*D=%u exceeds grid.z", grid.z); return OFSV_ENOSUP; }
    stage3d_hfast_kernel<SH, SN, S2D, FMA><<<grid, 256, 0, st>>>(q, P);
    return check_launch("stage3d_hfast_kernel");
  }
}

int block_stage_hfast(const float* head, const float* fm_prev, const float* img0, const float* img1, const float* lin_h,
                      const float* lin_d, const float* lin_w, float* fm_out, float* merged, float* mask_sig, void* pack_out, int N,
                      int D, int H, int W, int scale_head, int scale_next, int pack_s2d, int ref_mode, cudaStream_t st) {
  OFSV_REQUIRE((!head || (reinterpret_cast<uintptr_t>(head) & 31) == 0) && (!fm_out || (reinterpret_cast<uintptr_t>(fm_out) & 31) == 0) &&
                   (!fm_prev || (reinterpret_cast<uintptr_t>(fm_prev) & 31) == 0) && (!pack_out || (reinterpret_cast<uintptr_t>(pack_out) & 31) == 0),
               "ofsv_block_stage_3d: head / fm / pack_out must be 32-byte aligned in the H-fastest state layout");
  OFSV_REQUIRE(D % HF_PP == 0, "ofsv_block_stage_3d: the H-fastest state layout needs D %% %d == 0", HF_PP);
  OFSV_REQUIRE(W % 8 == 0 && (reinterpret_cast<uintptr_t>(img0) & 31) == 0 && (reinterpret_cast<uintptr_t>(img1) & 31) == 0 &&
                   (!merged || (reinterpret_cast<uintptr_t>(merged) & 31) == 0) && (!mask_sig || (reinterpret_cast<uintptr_t>(mask_sig) & 31) == 0),
               "ofsv_block_stage_3d: the H-fastest state layout needs W %% 8 == 0 and 32-byte aligned img0 / img1 / merged / mask_sig");
  const Warp3dParams P = make_warp3d_params(N, 1, D, H, W, ref_mode);
  HfPtrs q{head, fm_prev, img0, img1, lin_h, lin_d, lin_w, fm_out, merged, mask_sig, reinterpret_cast<__nv_bfloat16*>(pack_out)};
  const bool fma = ref_mode == OFSV_REF_CUDA;
  const bool s2d = pack_s2d != 0;
#define GO3(SH, SN, S2) return fma ? launch_hfast<SH, SN, S2, true>(q, P, st) : launch_hfast<SH, SN, S2, false>(q, P, st)
#define GO(SH)                                                                                         \
  do {                                                                                                 \
    if (scale_next == 0) GO3(SH, 0, false);                                                            \
    else if (scale_next == 1) { if (s2d) GO3(SH, 1, true); else GO3(SH, 1, false); }                   \
    else { if (s2d) GO3(SH, 2, true); else GO3(SH, 2, false); }                                        \
  } while (0)
  if (scale_head == 0) GO(0); else if (scale_head == 1) GO(1); else if (scale_head == 2) GO(2); else GO(4);
#undef GO
#undef GO3
  return OFSV_OK;
}

}  // namespace ofsv
