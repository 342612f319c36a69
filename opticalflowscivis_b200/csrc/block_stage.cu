// Fused IFBlock output stage for 3-D volumes on a CHANNELS-LAST flow/mask state (a4/a5/a6 glue, one pass per scale):
//
//   flow_d, mask_d = F.interpolate(head, scale) (flow_d *= scale)                 Flow-3D/model/IFNet.py:118-119
//   flow = flow + flow_d ; mask = mask + mask_d                                    :169-170
//   warped0 = warp(img0, flow[:, :3]) ; warped1 = warp(img1, flow[:, 3:6])         :190-191
//   mask_sig = sigmoid(mask) ; merged = warped0*mask_sig + warped1*(1-mask_sig)    :186,242      (optional)
//   next block's input = cat(img0,img1,warped0,warped1,mask,flow) resized by 1/s_next, flow/s_next   :82-90,166
//
// State layout: fm[N][D][H][W][8] fp32 = (flow 0..5, mask logit, 0).  One voxel = one 32 B sector, so the state is read
// and written with fully coalesced 16 B accesses (two threads per voxel) instead of seven strided 4 B planes, and the
// depth-to-space epilogue of the head conv produces exactly this layout.  The user-facing (N,6,D,H,W) flow tensors are
// permuted views of it (ops.py).
//
// CTA = 32(h) x 8(w) voxels at fixed (n, d) (two d planes when the next block's input is pooled 2x):
//   phase A  one voxel per thread: prev (+) up-sampled head -> fm_out (global, streaming) and, in place, the shared state tile
//   phase B  one voxel per thread, LANES ALONG h: the reference warp rotates axes, output h is the
//            contiguous source axis, so the 16 trilinear taps of a warp are 128 B-coalesced; sigmoid / blend; the 11
//            block-input channels go to shared memory (bf16 rows, or fp32 planes for the 2x2x2 mean)
//   phase C  coalesced stores of merged / sigmoid(mask) / the packed bf16 rows.
// The packed output is either plain channels-last [N][D/s][H/s][W/s][16] or the SHIFTED SPACE-TO-DEPTH form
// [N][Dn/2+1][Hn/2+1][Wn/2+1][2x2x2][16] (cell = (i+1)>>1, sub = (i+1)&1 per axis; border sub-cells stay zero) that turns
// the next conv0 (k=4, s=2, p=1) into a stride-1 2^3-tap conv over 128 channels for the halo tcgen05 kernel.
#include "warp_device.cuh"

namespace ofsv {

constexpr int BS_H = 32, BS_W = 8;
constexpr int BS_PKROW = BS_W * 2 + 1; // uint4 per tile row of the packed tile (32 B per voxel + 16 B pad)

struct Lerp1s {
  int i0, i1;
  float l0, l1;
};
// ATen area_pixel_compute_source_index (align_corners=False) + guard_index_and_lambda
__device__ __forceinline__ Lerp1s up_index1s(int dst, int n_in, float rscale) {
  float src = __fsub_rn(__fmul_rn(rscale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  src = src < 0.0f ? 0.0f : src;
  Lerp1s L;
  L.i0 = min((int)src, n_in - 1);
  L.i1 = L.i0 + (L.i0 < n_in - 1 ? 1 : 0);
  L.l1 = fminf(fmaxf(__fsub_rn(src, (float)L.i0), 0.0f), 1.0f);
  L.l0 = __fsub_rn(1.0f, L.l1);
  return L;
}

__device__ __forceinline__ uint32_t bs_pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// element (bf16) offset of the 16-channel row of position (z,y,x) of sample n in the packed tensor of logical size (Dn,Hn,Wn)
template <bool S2D>
__device__ __forceinline__ int64_t pack_row_off(int n, int z, int y, int x, int Dn, int Hn, int Wn) {
  if (!S2D) return ((((int64_t)n * Dn + z) * Hn + y) * Wn + x) * 16;
  return s2d_row(3, n, z, y, x, Dn, Hn, Wn) * 16;
}

struct StagePtrs {
  const float* head; const float* fm_prev; const float* img0; const float* img1;
  const float* lin_h; const float* lin_d; const float* lin_w;
  float* fm_out; float* merged; float* mask_sig; __nv_bfloat16* pack_out;
};

// c0*l0 + c1*l1 per component (the product c0*l0 rounded, then one FMA — the same expression in head_upsample_add_kernel)
__device__ __forceinline__ float4 lerp4(float4 a, float la, float4 b, float lb) {
  return make_float4(__fmaf_rn(b.x, lb, __fmul_rn(a.x, la)), __fmaf_rn(b.y, lb, __fmul_rn(a.y, la)),
                     __fmaf_rn(b.z, lb, __fmul_rn(a.z, la)), __fmaf_rn(b.w, lb, __fmul_rn(a.w, la)));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifndef OFSV_BS_MINB
#define OFSV_BS_MINB 4
#endif
#ifndef OFSV_BS_NST
#define OFSV_BS_NST 2
#endif
constexpr int BS_NST = OFSV_BS_NST;          // cp.async stages (planes in flight per CTA)
#ifndef OFSV_BS_DZ
#define OFSV_BS_DZ 8
#endif
constexpr int BS_DZ = OFSV_BS_DZ;            // d planes walked by one CTA
constexpr int BS_FROW = BS_W * 4 + 4;        // floats per tile row of a half-state tile (+16 B pad: lanes along h hit distinct banks)
constexpr int BS_HALF = BS_H * BS_FROW;      // floats per half-state tile (full-resolution head, SH == 1)
constexpr int BS_SROW = BS_W * 8 + 4;        // floats per row of the state tile: 8 voxels x 32 B as they lie in memory + 16 B pad
constexpr int BS_STATE = BS_H * BS_SROW;     // floats per state tile

// head tile of one (tile, plane) for SH > 1: the 2 x (32/SH + 2) x (8/SH + 2) coarse voxels all trilinear taps fall into
__host__ __device__ constexpr int bs_head_rows(int SH) { return SH > 1 ? BS_H / SH + 2 : 0; }
__host__ __device__ constexpr int bs_head_cols(int SH) { return SH > 1 ? BS_W / SH + 2 : 0; }
__host__ __device__ constexpr int bs_head_rowf(int SH) { return SH > 1 ? bs_head_cols(SH) * 8 + 4 : 0; }   // floats per row (+16 B pad)
__host__ __device__ constexpr int bs_head_tile(int SH) { return 2 * bs_head_rows(SH) * bs_head_rowf(SH); }

// shared-memory carve-up (bytes) — must match the kernel
__host__ __device__ constexpr int bs_smem_bytes(int SH, int SN) {
  return BS_NST * BS_STATE * 4                          // s_f          (prev state, updated in place)
         + (SH == 1 ? BS_NST * 2 * BS_HALF * 4 : 0)     // s_ha, s_hb   (full-resolution head tiles)
         + BS_NST * 2 * BS_H * (BS_W + 1) * 4           // s_img
         + 2 * BS_H * (BS_W + 1) * 4                    // s_out
         + (SN == 1 ? BS_H * BS_PKROW * 16 : 0)         // s_pk
         + (SN == 2 ? 22 * BS_H * (BS_W + 1) * 4 : 0)   // s_pool
         + 4 * (BS_H + BS_W + BS_DZ) * 4                // tap tables
         + BS_NST * bs_head_tile(SH) * 4;               // coarse head voxels of the plane's taps (SH > 1)
}

// SH: scale of the head (1, 2, 4), or 0 = the state is already accumulated (fm_prev holds flow/mask, nothing is added and
// fm_out is not written: the warp/blend-only pass after a head conv whose epilogue did `fm = fm_prev + head`).
//
// One CTA walks BS_DZ consecutive d planes of its 32(h) x 8(w) tile.  The streaming inputs of plane p+2 (previous state,
// full-resolution head, the voxel's own img0/img1 values) are in flight as cp.async copies while plane p is processed, so
// the only exposed latencies are the data-dependent gathers of phase B (hidden by the other resident CTAs).
template <int SH, int SN, bool S2D, bool FMA>
__global__ void __launch_bounds__(256, (SH == 0 && SN == 0) ? 4 : OFSV_BS_MINB) block_stage_3d_kernel(const StagePtrs q, const Warp3dParams P) {
  extern __shared__ __align__(16) uint8_t bs_smem[];
  float* s_f = reinterpret_cast<float*>(bs_smem);                  // [NST][BS_H][BS_SROW]: rows of 8 voxels x (flow 0..5, mask, 0)
  float* s_ha = s_f + BS_NST * BS_STATE;                           // [NST][BS_HALF] (SH == 1)
  float* s_hb = s_ha + (SH == 1 ? BS_NST * BS_HALF : 0);
  float (*s_img)[2][BS_H][BS_W + 1] = reinterpret_cast<float (*)[2][BS_H][BS_W + 1]>(s_hb + (SH == 1 ? BS_NST * BS_HALF : 0));
  float (*s_out)[BS_H][BS_W + 1] = reinterpret_cast<float (*)[BS_H][BS_W + 1]>(&s_img[BS_NST][0][0][0]);
  uint4* s_pk = reinterpret_cast<uint4*>(&s_out[2][0][0]);
  float (*s_pool)[BS_H][BS_W + 1] = reinterpret_cast<float (*)[BS_H][BS_W + 1]>(s_pk + (SN == 1 ? BS_H * BS_PKROW : 0));
  int* s_li = reinterpret_cast<int*>(&s_pool[SN == 2 ? 22 : 0][0][0]);   // i0, i1 element offsets of the head taps per tile row / col / plane
  float* s_ll = reinterpret_cast<float*>(s_li + 2 * (BS_H + BS_W + BS_DZ));
  float* s_head = s_ll + 2 * (BS_H + BS_W + BS_DZ);                // [NST][2 z taps][HNR][2 halves][HNC][4] (+16 B row pad)
  constexpr int HNR = bs_head_rows(SH), HNC = bs_head_cols(SH), HRF = bs_head_rowf(SH), HTILE = bs_head_tile(SH);

  const int H = P.H, W = P.W, D = P.D, HW = H * W;
  const int V = D * HW;                                     // < 2^28 (host check): 32-bit offsets inside one sample
  const int nzb = (D + BS_DZ - 1) / BS_DZ;
  const int n = blockIdx.z / nzb, dbeg = (blockIdx.z - n * nzb) * BS_DZ;
  const int nplanes = min(BS_DZ, D - dbeg);
  const int h0 = blockIdx.y * BS_H, w0 = blockIdx.x * BS_W;
  const int tid = threadIdx.x;
  // phase B mapping: lanes along h
  const int lane = tid & 31, wl = tid >> 5;
  const int hB = h0 + lane, wB = w0 + wl;
  const bool okB = hB < H && wB < W;
  // phase A/C mapping: 8 consecutive threads = one tile row (8 voxels = 256 B of state, 32 B of a planar volume)
  const int rP = tid >> 3, cP = tid & 7;
  const bool okP = (h0 + rP) < H && (w0 + cP) < W;
  const int sP = rP * BS_FROW + cP * 4;                     // this thread's slot in a half-state tile
  constexpr int SHD = SH > 1 ? SH : 1;
  const int Dh = D / SHD, Hh = H / SHD, Wh = W / SHD;
  const float* hb = SH ? q.head + (int64_t)n * Dh * Hh * Wh * 8 : nullptr;
  const float* fprev = q.fm_prev ? q.fm_prev + (int64_t)n * V * 8 : nullptr;
  float* fout = SH ? q.fm_out + (int64_t)n * V * 8 : nullptr;
  const float* i0p = q.img0 + (int64_t)n * V;
  const float* i1p = q.img1 + (int64_t)n * V;
  const bool has_prev = fprev != nullptr;
  const bool need_m = q.merged != nullptr || q.mask_sig != nullptr;
  const int gP0 = (h0 + rP) * W + w0 + cP;                  // in-plane voxel offset of this thread's phase A/C voxel

  // head tile copies (SH > 1): every thread owns at most two fixed 16 B pieces of the tile; their position inside the tile
  // and inside a coarse plane never changes, only the two coarse z planes do — decode once, not per plane
  constexpr int HPIECES = 2 * bs_head_rows(SH) * bs_head_cols(SH) * 2;      // (z tap, row, col, half)
  int hsm[2] = {-1, -1}, hgl[2] = {0, 0}, hz[2] = {0, 0};
  if (SH > 1) {
    const float rs = 1.0f / (float)SHD;
    const int yb = up_index1s(h0, Hh, rs).i0, xb = up_index1s(w0, Wh, rs).i0;
    constexpr int NCd = HNC > 0 ? HNC : 1, NRd = HNR > 0 ? HNR : 1;          // (dead code when SH <= 1)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int idx = tid + k * 256;
      if (idx < HPIECES) {
        const int half = idx & 1, c = (idx >> 1) % NCd, r = ((idx >> 1) / NCd) % NRd, z = (idx >> 1) / (NCd * NRd);
        const int yy = min(yb + r, Hh - 1), xx = min(xb + c, Wh - 1);
        hsm[k] = (z * HNR + r) * HRF + half * (HNC * 4) + c * 4;
        hgl[k] = (yy * Wh + xx) * 8 + half * 4;
        hz[k] = z;
      }
    }
  }
  static_assert(HPIECES <= 512, "head tile pieces per thread");
  // previous state: a tile row is 8 voxels x 32 B = 256 contiguous bytes, copied as it lies in memory.  One cp.async
  // instruction covers two whole rows (512 contiguous bytes = 16 sectors): LDGSTS costs one shared-memory wavefront per
  // global sector it touches, so the voxel-per-thread mapping (every thread one half sector, 32 sectors per instruction)
  // spent a third of the kernel's L1 data-pipe cycles on these copies.  (One bulk copy per row measured slower.)
  const int rows_valid = min(BS_H, H - h0), row_bytes = min(BS_W, W - w0) * 32;
  auto issue = [&](int it) {            // async copies of plane dbeg + it into stage it % NST (always commits a group)
    if (has_prev && it < nplanes) {
      const int st = it % BS_NST;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int pc = k * 256 + tid, row = pc >> 4, c = pc & 15;
        if (row < rows_valid && c * 16 < row_bytes)
          cp_async16(s_f + (st * BS_H + row) * BS_SROW + c * 4, fprev + ((int64_t)(dbeg + it) * HW + (int64_t)(h0 + row) * W + w0) * 8 + c * 4);
      }
    }
    if (it < nplanes && okP) {
      const int st = it % BS_NST, g = (dbeg + it) * HW + gP0;
      if (SH == 1) { cp_async16(s_ha + st * BS_HALF + sP, hb + g * 8); cp_async16(s_hb + st * BS_HALF + sP, hb + g * 8 + 4); }
      if (SN != 0) { cp_async4(&s_img[st][0][rP][cP], i0p + g); cp_async4(&s_img[st][1][rP][cP], i1p + g); }
    }
    if (SH > 1 && it < nplanes) {
      // the coarse head voxels this plane's taps need, so that phase A interpolates from shared memory
      const int st = it % BS_NST;
      const Lerp1s lz = up_index1s(dbeg + it, Dh, 1.0f / (float)SHD);
      const int64_t zo0 = (int64_t)lz.i0 * Hh * Wh * 8, zo1 = (int64_t)lz.i1 * Hh * Wh * 8;
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (hsm[k] >= 0) cp_async16(s_head + st * HTILE + hsm[k], hb + (hz[k] ? zo1 : zo0) + hgl[k]);
    }
    cp_async_commit();
  };
  for (int i = 0; i < BS_NST - 1; ++i) issue(i);

  if (SH > 1) {
    // per-axis tap tables of F.interpolate(scale_factor = SH, align_corners = False) for this tile column
    constexpr int NT = BS_H + BS_W + BS_DZ;
    if (tid < NT) {
      // offsets inside the plane's head tile [z tap][row - yb][col - xb][8]
      int dst, n_in, stride, base;
      const float rs = 1.0f / (float)SHD;
      if (tid < BS_H) { dst = h0 + tid; n_in = Hh; stride = HRF; base = up_index1s(h0, Hh, rs).i0; }
      else if (tid < BS_H + BS_W) { dst = w0 + tid - BS_H; n_in = Wh; stride = 4; base = up_index1s(w0, Wh, rs).i0; }
      else { dst = dbeg + tid - BS_H - BS_W; n_in = Dh; stride = 0; base = 0; }
      const Lerp1s L = up_index1s(dst, n_in, rs);
      s_li[2 * tid] = (L.i0 - base) * stride; s_li[2 * tid + 1] = (L.i1 - base) * stride;
      s_ll[2 * tid] = L.l0; s_ll[2 * tid + 1] = L.l1;
    }
  }

  for (int it = 0; it < nplanes; ++it) {
    const int d = dbeg + it, st = it % BS_NST;
    const int gP = d * HW + gP0;
    float* sf = s_f + st * BS_STATE;
    // a voxel's two 16 B halves are read / written in opposite order by lanes 0-3 and 4-7 of a quarter warp: with the rows
    // stored as they lie in memory (32 B per voxel) this is what keeps the 16 B accesses of phase A / C bank-conflict-free
    const int sw = (cP >> 2) & 1;
    float* myv = sf + rP * BS_SROW + cP * 8;
    issue(it + BS_NST - 1);
    cp_async_wait<BS_NST - 1>();
    __syncthreads();
    // ---------------- phase A: state update in place, one voxel (32 B) per thread
    if (SH != 0) {
      float4 oa = make_float4(0.f, 0.f, 0.f, 0.f), ob = oa;
      if (okP) {
        float4 va, vb;
        if (SH == 1) {
          va = *reinterpret_cast<const float4*>(s_ha + st * BS_HALF + sP);
          vb = *reinterpret_cast<const float4*>(s_hb + st * BS_HALF + sP);
        } else {
          const int ty = rP, tx = BS_H + cP, tz = BS_H + BS_W + it;
          const int y0 = s_li[2 * ty], y1 = s_li[2 * ty + 1], x0 = s_li[2 * tx], x1 = s_li[2 * tx + 1];
          const int z0 = 0, z1 = HNR * HRF;
          const float ly0 = s_ll[2 * ty], ly1 = s_ll[2 * ty + 1], lx0 = s_ll[2 * tx], lx1 = s_ll[2 * tx + 1];
          const float lz0 = s_ll[2 * tz], lz1 = s_ll[2 * tz + 1];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const float* r = s_head + st * HTILE + half * (HNC * 4);     // [z tap][row][half][col][4]: the 16 B taps of a tile row are contiguous
            auto L4 = [&](int o) { return *reinterpret_cast<const float4*>(r + o); };
            const float4 a00 = lerp4(L4(z0 + y0 + x0), lx0, L4(z0 + y0 + x1), lx1);
            const float4 a01 = lerp4(L4(z0 + y1 + x0), lx0, L4(z0 + y1 + x1), lx1);
            const float4 a10 = lerp4(L4(z1 + y0 + x0), lx0, L4(z1 + y0 + x1), lx1);
            const float4 a11 = lerp4(L4(z1 + y1 + x0), lx0, L4(z1 + y1 + x1), lx1);
            const float4 v = lerp4(lerp4(a00, ly0, a01, ly1), lz0, lerp4(a10, ly0, a11, ly1), lz1);
            if (half == 0) va = v; else vb = v;
          }
        }
        const float sh = (float)SHD;
        if (has_prev) {
          const float4 p0 = *reinterpret_cast<const float4*>(myv + sw * 4), p1 = *reinterpret_cast<const float4*>(myv + (sw ^ 1) * 4);
          const float4 pa = sw ? p1 : p0, pb = sw ? p0 : p1;
          oa = make_float4(__fadd_rn(pa.x, __fmul_rn(va.x, sh)), __fadd_rn(pa.y, __fmul_rn(va.y, sh)),
                           __fadd_rn(pa.z, __fmul_rn(va.z, sh)), __fadd_rn(pa.w, __fmul_rn(va.w, sh)));
          ob = make_float4(__fadd_rn(pb.x, __fmul_rn(vb.x, sh)), __fadd_rn(pb.y, __fmul_rn(vb.y, sh)), __fadd_rn(pb.z, vb.z), 0.0f);
        } else {   // block 0: flow = flow_d exactly
          oa = make_float4(__fmul_rn(va.x, sh), __fmul_rn(va.y, sh), __fmul_rn(va.z, sh), __fmul_rn(va.w, sh));
          ob = make_float4(__fmul_rn(vb.x, sh), __fmul_rn(vb.y, sh), vb.z, 0.0f);
        }
        stg_stream4(fout + gP * 8, oa);
        stg_stream4(fout + gP * 8 + 4, ob);
      }
      *reinterpret_cast<float4*>(myv + sw * 4) = sw ? ob : oa;
      *reinterpret_cast<float4*>(myv + (sw ^ 1) * 4) = sw ? oa : ob;
      __syncthreads();
    }
    // ---------------- phase B: warps / blend, one voxel per thread, lanes along h
    if (okB) {
      const float4 va = *reinterpret_cast<const float4*>(sf + lane * BS_SROW + wl * 8);
      const float4 vb = *reinterpret_cast<const float4*>(sf + lane * BS_SROW + wl * 8 + 4);
      const float m = vb.z;
      const float lh = __ldg(q.lin_h + hB), ld = __ldg(q.lin_d + d), lw = __ldg(q.lin_w + wB);
      const Trilin t0 = trilin_setup(va.x, va.y, va.z, lh, ld, lw, D, H, W, P.hs, P.ref_mode);
      const Trilin t1 = trilin_setup(va.w, vb.x, vb.y, lh, ld, lw, D, H, W, P.hs, P.ref_mode);
      const Taps8 g0 = trilin_gather(i0p, t0), g1 = trilin_gather(i1p, t1);     // 16 independent loads in flight
      const float a = trilin_reduce<FMA>(g0, t0), b = trilin_reduce<FMA>(g1, t1);
      if (need_m) {
        const float ms = sigmoidf_ref(m);
        s_out[0][lane][wl] = __fadd_rn(__fmul_rn(a, ms), __fmul_rn(b, __fsub_rn(1.0f, ms)));
        s_out[1][lane][wl] = ms;
      }
      if (SN == 1) {
        const float i0v = s_img[st][0][lane][wl], i1v = s_img[st][1][lane][wl];
        uint4 lo, hi;
        lo.x = bs_pack2(i0v, i1v); lo.y = bs_pack2(a, b); lo.z = bs_pack2(m, va.x); lo.w = bs_pack2(va.y, va.z);
        hi.x = bs_pack2(va.w, vb.x); hi.y = bs_pack2(vb.y, 0.0f); hi.z = 0u; hi.w = 0u;
        s_pk[lane * BS_PKROW + wl * 2] = lo;
        s_pk[lane * BS_PKROW + wl * 2 + 1] = hi;
      } else if (SN == 2) {
        const float c11[11] = {s_img[st][0][lane][wl], s_img[st][1][lane][wl], a, b, m, va.x, va.y, va.z, va.w, vb.x, vb.y};
#pragma unroll
        for (int c = 0; c < 11; ++c) s_pool[(it & 1) * 11 + c][lane][wl] = c11[c];
      }
    }
    __syncthreads();
    // ---------------- phase C: coalesced stores
    if (okP) {
      const int64_t g = (int64_t)n * V + gP;
      if (q.merged) q.merged[g] = s_out[0][rP][cP];
      if (q.mask_sig) q.mask_sig[g] = s_out[1][rP][cP];
      if (SN == 1) {
        uint4* o = reinterpret_cast<uint4*>(q.pack_out + pack_row_off<S2D>(n, d, h0 + rP, w0 + cP, D, H, W));
        o[sw] = s_pk[rP * BS_PKROW + cP * 2 + sw];
        o[sw ^ 1] = s_pk[rP * BS_PKROW + cP * 2 + (sw ^ 1)];
      }
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
            const float (*pl)[BS_W + 1] = s_pool[dz * 11 + c];
            const float r0 = __fadd_rn(__fmul_rn(pl[2 * ph][2 * pw], 0.5f), __fmul_rn(pl[2 * ph][2 * pw + 1], 0.5f));
            const float r1 = __fadd_rn(__fmul_rn(pl[2 * ph + 1][2 * pw], 0.5f), __fmul_rn(pl[2 * ph + 1][2 * pw + 1], 0.5f));
            rz[dz] = __fadd_rn(__fmul_rn(r0, 0.5f), __fmul_rn(r1, 0.5f));
          }
          float r = __fadd_rn(__fmul_rn(rz[0], 0.5f), __fmul_rn(rz[1], 0.5f));
          if (c >= 5) r = __fmul_rn(r, 0.5f);
          c11[c] = r;
        }
        uint4 lo, hi;
        lo.x = bs_pack2(c11[0], c11[1]); lo.y = bs_pack2(c11[2], c11[3]); lo.z = bs_pack2(c11[4], c11[5]); lo.w = bs_pack2(c11[6], c11[7]);
        hi.x = bs_pack2(c11[8], c11[9]); hi.y = bs_pack2(c11[10], 0.0f); hi.z = 0u; hi.w = 0u;
        uint4* o = reinterpret_cast<uint4*>(q.pack_out + pack_row_off<S2D>(n, d / 2, oh, ow, D / 2, H / 2, W / 2));
        o[0] = lo; o[1] = hi;
      }
    }
  }
  cp_async_wait<0>();
}

template <int SH, int SN, bool S2D, bool FMA>
static int launch_stage(const StagePtrs& q, const Warp3dParams& P, dim3 grid, cudaStream_t st) {
  constexpr int smem = bs_smem_bytes(SH, SN);
  static std::atomic<uint64_t> attr_done{0};
  if (int e = ensure_dyn_smem(attr_done, block_stage_3d_kernel<SH, SN, S2D, FMA>, smem, "ofsv_block_stage_3d")) return e;
  block_stage_3d_kernel<SH, SN, S2D, FMA><<<grid, 256, smem, st>>>(q, P);
  return OFSV_OK;
}

// block_stage_hfast.cu: the same stage on the H-fastest state layout
int block_stage_hfast(const float* head, const float* fm_prev, const float* img0, const float* img1, const float* lin_h,
                      const float* lin_d, const float* lin_w, float* fm_out, float* merged, float* mask_sig, void* pack_out, int N,
                      int D, int H, int W, int scale_head, int scale_next, int pack_s2d, int ref_mode, cudaStream_t st);

}  // namespace ofsv

using namespace ofsv;

extern "C" int ofsv_block_stage_3d(const float* head, const float* fm_prev, const float* img0, const float* img1,
                                   const float* lin_h, const float* lin_d, const float* lin_w, float* fm_out, float* merged,
                                   float* mask_sig, void* pack_out, int N, int D, int H, int W, int scale_head,
                                   int scale_next, int pack_s2d, int ref_mode, int state_layout, void* stream) {
  OFSV_REQUIRE(state_layout == OFSV_STATE_DHW8 || state_layout == OFSV_STATE_DWH8, "ofsv_block_stage_3d: bad state_layout");
  OFSV_REQUIRE(N >= 0 && D >= 1 && H >= 1 && W >= 1 && (int64_t)D * H * W < (1ll << 28), "ofsv_block_stage_3d: bad shape (D*H*W must be < 2^28)");
  OFSV_REQUIRE(scale_head == 0 || scale_head == 1 || scale_head == 2 || scale_head == 4, "ofsv_block_stage_3d: scale_head %d not in {0,1,2,4}", scale_head);
  OFSV_REQUIRE(scale_next == 0 || scale_next == 1 || scale_next == 2, "ofsv_block_stage_3d: scale_next %d not in {0,1,2}", scale_next);
  OFSV_REQUIRE(scale_head == 0 || (D % scale_head == 0 && H % scale_head == 0 && W % scale_head == 0), "ofsv_block_stage_3d: dims must be multiples of scale_head");
  OFSV_REQUIRE(scale_next != 2 || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "ofsv_block_stage_3d: dims must be even for scale_next = 2");
  OFSV_REQUIRE(!pack_s2d || (scale_next != 0 && D % (2 * scale_next) == 0 && H % (2 * scale_next) == 0 && W % (2 * scale_next) == 0),
               "ofsv_block_stage_3d: space-to-depth packing needs dims that are multiples of 2*scale_next");
  OFSV_REQUIRE(ref_mode == OFSV_REF_CPU || ref_mode == OFSV_REF_CUDA, "ofsv_block_stage_3d: bad ref_mode");
  if (N == 0) return OFSV_OK;
  OFSV_REQUIRE(img0 && img1 && lin_h && lin_d && lin_w, "ofsv_block_stage_3d: null pointer");
  if (scale_head == 0) OFSV_REQUIRE(fm_prev && !head && !fm_out, "ofsv_block_stage_3d: scale_head = 0 takes the accumulated state in fm_prev (head and fm_out must be NULL)");
  else OFSV_REQUIRE(head && fm_out, "ofsv_block_stage_3d: null pointer");
  OFSV_REQUIRE((scale_next == 0) == (pack_out == nullptr), "ofsv_block_stage_3d: pack_out must be given iff scale_next != 0");
  OFSV_REQUIRE((!head || aligned16(head)) && (!fm_out || aligned16(fm_out)) && (!fm_prev || aligned16(fm_prev)) && (!pack_out || aligned16(pack_out)),
               "ofsv_block_stage_3d: head / fm / pack_out must be 16-byte aligned");
  if (state_layout == OFSV_STATE_DWH8)
    return block_stage_hfast(head, fm_prev, img0, img1, lin_h, lin_d, lin_w, fm_out, merged, mask_sig, pack_out, N, D, H, W, scale_head,
                             scale_next, pack_s2d, ref_mode, (cudaStream_t)stream);
  const Warp3dParams P = make_warp3d_params(N, 1, D, H, W, ref_mode);
  const dim3 grid((unsigned)cdiv(W, BS_W), (unsigned)cdiv(H, BS_H), (unsigned)(N * cdiv(D, BS_DZ)));
  if (grid.z > 65535u) { set_error("ofsv_block_stage_3d: N*D=%u exceeds grid.z", grid.z); return OFSV_ENOSUP; }
  StagePtrs q{head, fm_prev, img0, img1, lin_h, lin_d, lin_w, fm_out, merged, mask_sig, reinterpret_cast<__nv_bfloat16*>(pack_out)};
  cudaStream_t st = (cudaStream_t)stream;
  const bool fma = ref_mode == OFSV_REF_CUDA;
  const bool s2d = pack_s2d != 0;
  int rc = OFSV_OK;
#define GO3(SH, SN, S2)                                                                                \
  do {                                                                                                 \
    rc = fma ? launch_stage<SH, SN, S2, true>(q, P, grid, st) : launch_stage<SH, SN, S2, false>(q, P, grid, st);  \
  } while (0)
#define GO(SH)                                                                                         \
  do {                                                                                                 \
    if (scale_next == 0) GO3(SH, 0, false);                                                            \
    else if (scale_next == 1) { if (s2d) GO3(SH, 1, true); else GO3(SH, 1, false); }                   \
    else { if (s2d) GO3(SH, 2, true); else GO3(SH, 2, false); }                                        \
  } while (0)
  if (scale_head == 0) GO(0); else if (scale_head == 1) GO(1); else if (scale_head == 2) GO(2); else GO(4);
#undef GO
#undef GO3
  if (rc != OFSV_OK) return rc;
  return check_launch("block_stage_3d_kernel");
}
