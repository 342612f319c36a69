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

constexpr int HF_H = 32, HF_W = 8, HF_DZ = 8;
#ifndef OFSV_HF_HEAD_MODE
#define OFSV_HF_HEAD_MODE 0      // 0: both x-lerped coarse rows re-loaded every plane; 1: kept across planes that share a coarse z tap
#endif
#ifndef OFSV_HF_OWN_PREFETCH
#define OFSV_HF_OWN_PREFETCH 1   // 1: the voxel's own image values are loaded one plane ahead
#endif

struct V8 { float v[8]; };
__device__ __forceinline__ V8 ldg256(const float* p) {
  V8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg256(float* p, const V8& r) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
               "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
               : "memory");
}
__device__ __forceinline__ void stg256_b32(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g,
                                           uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h)
               : "memory");
}
__device__ __forceinline__ uint32_t hf_pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
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

// SH: scale of the head (1, 2, 4), or 0 = fm_prev already holds the accumulated state (nothing is added, fm_out not written).
// SN: 0 = no packed output, 1 = next block at full resolution, 2 = at half resolution (2x2x2 mean).
template <int SH, int SN, bool S2D, bool FMA>
__global__ void __launch_bounds__(256, (SH == 0 && SN == 0) ? 4 : 3) stage3d_hfast_kernel(const HfPtrs q, const Warp3dParams P) {
  __shared__ float s_out[2][2][HF_H][HF_W + 1];                     // [buffer][merged | mask][h][w]
  __shared__ float s_pool[SN == 2 ? 22 : 1][HF_H][HF_W + 1];        // [plane parity * 11 + channel][h][w]

  const int H = P.H, W = P.W, D = P.D, HW = H * W;
  const int V = D * HW;                                             // < 2^28 (host check)
  const int nzb = (D + HF_DZ - 1) / HF_DZ;
  const int n = blockIdx.z / nzb, dbeg = (blockIdx.z - n * nzb) * HF_DZ;
  const int nplanes = min(HF_DZ, D - dbeg);
  const int h0 = blockIdx.y * HF_H, w0 = blockIdx.x * HF_W;
  const int tid = threadIdx.x, lane = tid & 31, wl = tid >> 5;
  const int h = h0 + lane, w = w0 + wl;
  const bool ok = h < H && w < W;
  const int hc = min(h, H - 1), wc = min(w, W - 1);                 // clamped: out-of-tile lanes compute on a valid voxel and store nothing
  const int rP = tid >> 3, cP = tid & 7;                            // planar-output mapping: 8 consecutive threads = one tile row
  const bool okP = (h0 + rP) < H && (w0 + cP) < W;
  constexpr int SHD = SH > 1 ? SH : 1;
  const int Dh = D / SHD, Hh = H / SHD, Wh = W / SHD;
  const float* hb = SH ? q.head + (int64_t)n * Dh * Hh * Wh * 8 : nullptr;
  const float* fprev = q.fm_prev ? q.fm_prev + (int64_t)n * V * 8 : nullptr;
  float* fout = SH ? q.fm_out + (int64_t)n * V * 8 : nullptr;
  const float* i0p = q.img0 + (int64_t)n * V;
  const float* i1p = q.img1 + (int64_t)n * V;
  const bool has_prev = fprev != nullptr;
  const bool need_m = q.merged != nullptr || q.mask_sig != nullptr;
  const float lh = __ldg(q.lin_h + hc), lw = __ldg(q.lin_w + wc);
  // in-plane element offset of this thread's voxel in the [W][H] state planes and the [H][W] image planes
  const int so = wc * H + hc;

  // up-sampled head: per-axis taps.  w (x) and d (z) taps are warp-uniform; the h (y) taps are per lane.  Lane L also OWNS coarse
  // row hb0 + L of the tile: it evaluates the x lerp of that row for both z taps, the fine lanes pick rows (y0 - hb0, y1 - hb0).
  HfLerp Ly{0, 0, 0.f, 0.f}, Lx{0, 0, 0.f, 0.f};
  int hb0 = 0, own = 0;
  if (SH > 1) {
    const float rs = 1.0f / (float)SHD;
    Ly = hf_up_index(hc, Hh, rs);
    Lx = hf_up_index(wc, Wh, rs);
    hb0 = hf_up_index(h0, Hh, rs).i0;
    own = min(hb0 + lane, Hh - 1);
  }

  V8 pv;                                                            // previous state of the plane being processed (prefetched)
  if (has_prev) pv = ldg256(fprev + ((int64_t)dbeg * HW + so) * 8);
  // own image values (channels 0, 1 of the next block's input): lanes along h read 32 different lines per request; moving them
  // through a coalescing shared tile needs a block barrier per plane, which measured slower (568 vs 485 us per 256^3 pair for the
  // middle stage) than the direct loads issued one plane ahead
  const int io = hc * W + wc;
#if OFSV_HF_OWN_PREFETCH
  float pi0 = 0.f, pi1 = 0.f;
  if (SN != 0) { pi0 = __ldg(i0p + dbeg * HW + io); pi1 = __ldg(i1p + dbeg * HW + io); }
#endif
  // x-lerped coarse head rows of the two z taps, kept across planes (consecutive planes share one or both coarse z taps)
  V8 a0, a1;
  int cz0 = -1, cz1 = -1;
#pragma unroll
  for (int i = 0; i < 8; ++i) { a0.v[i] = 0.f; a1.v[i] = 0.f; }

  for (int it = 0; it < nplanes; ++it) {
    const int d = dbeg + it;
    const float ld = __ldg(q.lin_d + d);
    V8 st;                                                          // flow 0..5, mask logit, 0
    V8 cur = pv;
    if (has_prev && it + 1 < nplanes) pv = ldg256(fprev + ((int64_t)(d + 1) * HW + so) * 8);
    if (SH != 0) {
      V8 hv;
      if (SH == 1) {
        hv = ldg256(hb + ((int64_t)d * HW + so) * 8);
      } else {
        const HfLerp Lz = hf_up_index(d, Dh, 1.0f / (float)SHD);
        const int c0 = Lx.i0 * Hh + own, c1 = Lx.i1 * Hh + own;
        // x lerp of coarse row `own` for the two z taps (ATen order: x, then y, then z); warp-uniform reuse across planes
        auto xrow = [&](int z) -> V8 {
          const int64_t r = ((int64_t)z * Wh) * Hh;
          return hf_lerp8(ldg256(hb + (r + c0) * 8), Lx.l0, ldg256(hb + (r + c1) * 8), Lx.l1);
        };
#if OFSV_HF_HEAD_MODE == 0
        a0 = xrow(Lz.i0);                                          // every lane loads (rows beyond the tile's 32/SH + 2 are never read)
        a1 = xrow(Lz.i1);
#else
        if (Lz.i0 != cz0 || Lz.i1 != cz1) {                         // warp-uniform: consecutive planes share one or both coarse z taps
          if (Lz.i0 == cz1) a0 = a1; else a0 = xrow(Lz.i0);
          if (Lz.i1 == Lz.i0) a1 = a0; else a1 = xrow(Lz.i1);
          cz0 = Lz.i0; cz1 = Lz.i1;
        }
#endif
        const int s0 = Ly.i0 - hb0, s1 = Ly.i1 - hb0;              // owner lanes of this lane's two coarse rows (< 32/SH + 2)
        const V8 b0 = hf_lerp8(hf_shfl8(a0, s0), Ly.l0, hf_shfl8(a0, s1), Ly.l1);
        const V8 b1 = hf_lerp8(hf_shfl8(a1, s0), Ly.l0, hf_shfl8(a1, s1), Ly.l1);
        hv = hf_lerp8(b0, Lz.l0, b1, Lz.l1);
      }
      const float sh = (float)SHD;
      if (has_prev) {
#pragma unroll
        for (int i = 0; i < 6; ++i) st.v[i] = __fadd_rn(cur.v[i], __fmul_rn(hv.v[i], sh));
        st.v[6] = __fadd_rn(cur.v[6], hv.v[6]);
      } else {                                                      // block 0: flow = flow_d exactly
#pragma unroll
        for (int i = 0; i < 6; ++i) st.v[i] = __fmul_rn(hv.v[i], sh);
        st.v[6] = hv.v[6];
      }
      st.v[7] = 0.0f;
      if (ok) stg256(fout + ((int64_t)d * HW + so) * 8, st);
    } else {
      st = cur;
    }
    // ---------------- warps / blend
    const float m = st.v[6];
    const Trilin t0 = trilin_setup(st.v[0], st.v[1], st.v[2], lh, ld, lw, D, H, W, P.hs, P.ref_mode);
    const Trilin t1 = trilin_setup(st.v[3], st.v[4], st.v[5], lh, ld, lw, D, H, W, P.hs, P.ref_mode);
    const Taps8 g0 = trilin_gather(i0p, t0), g1 = trilin_gather(i1p, t1);     // 16 independent loads in flight
#if OFSV_HF_OWN_PREFETCH
    const float i0v = pi0, i1v = pi1;
    if (SN != 0 && it + 1 < nplanes) { pi0 = __ldg(i0p + (d + 1) * HW + io); pi1 = __ldg(i1p + (d + 1) * HW + io); }
#else
    float i0v = 0.f, i1v = 0.f;
    if (SN != 0) { i0v = __ldg(i0p + d * HW + io); i1v = __ldg(i1p + d * HW + io); }
#endif
    const float a = trilin_reduce<FMA>(g0, t0), b = trilin_reduce<FMA>(g1, t1);
    const int ob = it & 1;
    if (need_m) {
      const float ms = sigmoidf_ref(m);
      s_out[ob][0][lane][wl] = __fadd_rn(__fmul_rn(a, ms), __fmul_rn(b, __fsub_rn(1.0f, ms)));
      s_out[ob][1][lane][wl] = ms;
    }
    if (SN == 1) {
      if (ok) {
        int64_t ro;
        if (S2D) ro = s2d_row(3, n, d, h, w, D, H, W) * 16;
        else ro = ((((int64_t)n * D + d) * H + h) * W + w) * 16;
        stg256_b32(q.pack_out + ro, hf_pack2(i0v, i1v), hf_pack2(a, b), hf_pack2(m, st.v[0]), hf_pack2(st.v[1], st.v[2]),
                   hf_pack2(st.v[3], st.v[4]), hf_pack2(st.v[5], 0.0f), 0u, 0u);
      }
    } else if (SN == 2) {
      const float c11[11] = {i0v, i1v, a, b, m, st.v[0], st.v[1], st.v[2], st.v[3], st.v[4], st.v[5]};
#pragma unroll
      for (int c = 0; c < 11; ++c) s_pool[(it & 1) * 11 + c][lane][wl] = c11[c];
    }
    if (need_m || SN == 2) __syncthreads();
    // ---------------- planar outputs: coalesced 32 B rows
    if (need_m && okP) {
      const int64_t g = (int64_t)n * V + (int64_t)d * HW + (h0 + rP) * W + w0 + cP;
      if (q.merged) q.merged[g] = s_out[ob][0][rP][cP];
      if (q.mask_sig) q.mask_sig[g] = s_out[ob][1][rP][cP];
    }
    if (SN == 2 && (it & 1) && tid < 64) {
      // 2x2x2 mean == F.interpolate(., 0.5): nested W, H, D like ATen; flow channels additionally * 0.5
      const int ph = tid >> 2, pw = tid & 3;
      const int oh = h0 / 2 + ph, ow = w0 / 2 + pw;
      if (oh < H / 2 && ow < W / 2) {
        float c11[11];
#pragma unroll
        for (int c = 0; c < 11; ++c) {
          float rz[2];
#pragma unroll
          for (int dz = 0; dz < 2; ++dz) {
            const float (*pl)[HF_W + 1] = s_pool[dz * 11 + c];
            const float r0 = __fadd_rn(__fmul_rn(pl[2 * ph][2 * pw], 0.5f), __fmul_rn(pl[2 * ph][2 * pw + 1], 0.5f));
            const float r1 = __fadd_rn(__fmul_rn(pl[2 * ph + 1][2 * pw], 0.5f), __fmul_rn(pl[2 * ph + 1][2 * pw + 1], 0.5f));
            rz[dz] = __fadd_rn(__fmul_rn(r0, 0.5f), __fmul_rn(r1, 0.5f));
          }
          float r = __fadd_rn(__fmul_rn(rz[0], 0.5f), __fmul_rn(rz[1], 0.5f));
          if (c >= 5) r = __fmul_rn(r, 0.5f);
          c11[c] = r;
        }
        int64_t ro;
        if (S2D) ro = s2d_row(3, n, d / 2, oh, ow, D / 2, H / 2, W / 2) * 16;
        else ro = ((((int64_t)n * (D / 2) + d / 2) * (H / 2) + oh) * (W / 2) + ow) * 16;
        stg256_b32(q.pack_out + ro, hf_pack2(c11[0], c11[1]), hf_pack2(c11[2], c11[3]), hf_pack2(c11[4], c11[5]), hf_pack2(c11[6], c11[7]),
                   hf_pack2(c11[8], c11[9]), hf_pack2(c11[10], 0.0f), 0u, 0u);
      }
    }
    // s_out is double-buffered (plane parity); s_pool's two plane halves are written on alternating planes and read right after
    // the odd plane's barrier: the next write to either half happens after the NEXT plane's barrier only for s_out, so guard s_pool
    if (SN == 2 && (it & 1)) __syncthreads();
  }
}

template <int SH, int SN, bool S2D, bool FMA>
static int launch_hfast(const HfPtrs& q, const Warp3dParams& P, dim3 grid, cudaStream_t st) {
  stage3d_hfast_kernel<SH, SN, S2D, FMA><<<grid, 256, 0, st>>>(q, P);
  return check_launch("stage3d_hfast_kernel");
}

int block_stage_hfast(const float* head, const float* fm_prev, const float* img0, const float* img1, const float* lin_h,
                      const float* lin_d, const float* lin_w, float* fm_out, float* merged, float* mask_sig, void* pack_out, int N,
                      int D, int H, int W, int scale_head, int scale_next, int pack_s2d, int ref_mode, cudaStream_t st) {
  OFSV_REQUIRE((!head || (reinterpret_cast<uintptr_t>(head) & 31) == 0) && (!fm_out || (reinterpret_cast<uintptr_t>(fm_out) & 31) == 0) &&
                   (!fm_prev || (reinterpret_cast<uintptr_t>(fm_prev) & 31) == 0) && (!pack_out || (reinterpret_cast<uintptr_t>(pack_out) & 31) == 0),
               "ofsv_block_stage_3d: head / fm / pack_out must be 32-byte aligned in the H-fastest state layout");
  const Warp3dParams P = make_warp3d_params(N, 1, D, H, W, ref_mode);
  const dim3 grid((unsigned)cdiv(W, HF_W), (unsigned)cdiv(H, HF_H), (unsigned)(N * cdiv(D, HF_DZ)));
  if (grid.z > 65535u) { set_error("ofsv_block_stage_3d: N*D=%u exceeds grid.z", grid.z); return OFSV_ENOSUP; }
  HfPtrs q{head, fm_prev, img0, img1, lin_h, lin_d, lin_w, fm_out, merged, mask_sig, reinterpret_cast<__nv_bfloat16*>(pack_out)};
  const bool fma = ref_mode == OFSV_REF_CUDA;
  const bool s2d = pack_s2d != 0;
#define GO3(SH, SN, S2) return fma ? launch_hfast<SH, SN, S2, true>(q, P, grid, st) : launch_hfast<SH, SN, S2, false>(q, P, grid, st)
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
