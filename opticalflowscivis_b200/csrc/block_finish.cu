// Fused IFBlock output stage for 3-D volumes (a4/a5/a6 glue in ONE pass over the full-resolution voxels):
//
//   flow_d, mask_d = F.interpolate(head, scale) (flow_d *= scale)           Flow-3D/model/IFNet.py:118-119
//   flow = flow + flow_d ; mask = mask + mask_d                              :169-170
//   warped0 = warp(img0, flow[:, :3]) ; warped1 = warp(img1, flow[:, 3:6])   :190-191
//   mask_sig = sigmoid(mask) ; merged = warped0*mask_sig + warped1*(1-mask_sig)   :186,242      (optional)
//   next block's input = cat(img0,img1,warped0,warped1,mask, flow) resized by 1/s_next, flow/s_next   :82-90,166
//       written straight into the channels-last bf16 tensor the next conv0 reads                   (optional)
//
// so flow/mask are read once and written once per scale, the warped volumes never touch HBM, and the 11-channel concat
// of the reference (738 MB at 256^3) is never materialised in fp32.  Tile/thread mapping as in warp.cu: 32(h) x 8(w)
// voxels per CTA, lanes along h so the rotated source gathers are 128 B-coalesced.
#include "warp_device.cuh"

namespace ofsv {

struct Lerp1 {
  int i0, i1;
  float l0, l1;
};
// ATen area_pixel_compute_source_index (align_corners=False) + guard_index_and_lambda
__device__ __forceinline__ Lerp1 up_index1(int dst, int n_in, float rscale) {
  float src = __fsub_rn(__fmul_rn(rscale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  src = src < 0.0f ? 0.0f : src;
  Lerp1 L;
  L.i0 = min((int)src, n_in - 1);
  L.i1 = L.i0 + (L.i0 < n_in - 1 ? 1 : 0);
  L.l1 = fminf(fmaxf(__fsub_rn(src, (float)L.i0), 0.0f), 1.0f);
  L.l0 = __fsub_rn(1.0f, L.l1);
  return L;
}

__device__ __forceinline__ void ld8(const float* p, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// 16 channels (11 used) of one block-input position -> 32 B
__device__ __forceinline__ void store_pack_row(__nv_bfloat16* dst, const float* c11) {
  uint4 lo, hi;
  lo.x = pack_bf16x2(c11[0], c11[1]); lo.y = pack_bf16x2(c11[2], c11[3]);
  lo.z = pack_bf16x2(c11[4], c11[5]); lo.w = pack_bf16x2(c11[6], c11[7]);
  hi.x = pack_bf16x2(c11[8], c11[9]); hi.y = pack_bf16x2(c11[10], 0.0f); hi.z = 0u; hi.w = 0u;
  reinterpret_cast<uint4*>(dst)[0] = lo;
  reinterpret_cast<uint4*>(dst)[1] = hi;
}


// Lean tile I/O for the vectorised path: thread t owns tile row (t & 63) >> 1, columns ((t & 1) * 4 .. +3) of planes
// g, g+4, g+8 with g = t >> 6 — the row/column arithmetic is done once per CTA, each plane costs one 16 B access.
struct TileIO {
  int row, c4, goff;   // tile-local row / first column, element offset of (h0+row, w0+c4) inside a (H, W) plane
  bool in;             // inside the volume (W % 4 == 0 on this path)
};
__device__ __forceinline__ TileIO make_tile_io(int h0, int w0, int H, int W) {
  TileIO t;
  const int r = threadIdx.x & 63;
  t.row = r >> 1; t.c4 = (r & 1) * 4;
  t.in = (h0 + t.row) < H && (w0 + t.c4) < W;
  t.goff = (h0 + t.row) * W + w0 + t.c4;
  return t;
}
__device__ __forceinline__ void tile_load(float (*s)[T3P], const float* __restrict__ plane, const TileIO& t) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (plane != nullptr && t.in) v = ldg_stream4(plane + t.goff);
  float* d = &s[t.row][t.c4];
  d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
}
__device__ __forceinline__ void tile_store(const float (*s)[T3P], float* __restrict__ plane, const TileIO& t) {
  if (plane != nullptr && t.in) {
    const float* d = &s[t.row][t.c4];
    stg_stream4(plane + t.goff, make_float4(d[0], d[1], d[2], d[3]));
  }
}

struct FinishPtrs {
  const float* head; const float* flow_prev; const float* mask_prev; const float* img0; const float* img1;
  const float* lin_h; const float* lin_d; const float* lin_w;
  float* flow_out; float* mask_out; float* merged; float* mask_sig; __nv_bfloat16* pack_out;
};

template <int SH, int SN, bool VEC, bool FMA>
__global__ void __launch_bounds__(256, 4)
    block_finish_3d_kernel(const FinishPtrs q, const Warp3dParams P, const int Cs) {
  constexpr int TDZ = SN == 2 ? 2 : 1;
  __shared__ float s[9][T3H][T3P];                          // in: flow_prev 0..5, mask_prev, img0, img1 ; out: flow 0..5, mask, merged, sigmoid
  __shared__ float sp[SN == 2 ? 2 * 11 : 1][T3H][T3P];      // SN == 2: the 11 block-input channels of both d planes, for the 2x2x2 mean
  const int H = P.H, W = P.W, D = P.D, HW = H * W;
  const int64_t V = (int64_t)D * HW;
  const int nzb = D / TDZ;
  const int n = blockIdx.z / nzb, d0 = (blockIdx.z - n * nzb) * TDZ;
  const int h0 = blockIdx.y * T3H, w0 = blockIdx.x * T3W;
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int h = h0 + lane, w = w0 + wl;
  const bool ok = h < H && w < W;
  const int Dh = D / SH, Hh = H / SH, Wh = W / SH;
  const float* hb = q.head + (int64_t)n * Dh * Hh * Wh * Cs;
  const bool has_prev = q.flow_prev != nullptr;
  const bool need_m = q.merged != nullptr || q.mask_sig != nullptr;
  const TileIO tio = make_tile_io(h0, w0, H, W);

#pragma unroll
  for (int dz = 0; dz < TDZ; ++dz) {
    const int d = d0 + dz;
    const int64_t plane = (int64_t)n * V + (int64_t)d * HW;
    // ---- up-sampled head (flow delta x6, mask delta): independent of the shared tile, issued first so its latency
    //      overlaps the plane loads below
    float v[8];
    if (ok) {
      if (SH == 1) {
        ld8(hb + (((int64_t)d * H + h) * W + w) * Cs, v);
      } else {
        const float rs = 1.0f / (float)SH;
        const Lerp1 lx = up_index1(w, Wh, rs), ly = up_index1(h, Hh, rs), lz = up_index1(d, Dh, rs);
        float az[2][7];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int zz = a ? lz.i1 : lz.i0;
          float ay[2][7];
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int yy = b ? ly.i1 : ly.i0;
            float c0[8], c1[8];
            ld8(hb + (((int64_t)zz * Hh + yy) * Wh + lx.i0) * Cs, c0);
            ld8(hb + (((int64_t)zz * Hh + yy) * Wh + lx.i1) * Cs, c1);
#pragma unroll
            for (int c = 0; c < 7; ++c) ay[b][c] = __fadd_rn(__fmul_rn(c0[c], lx.l0), __fmul_rn(c1[c], lx.l1));
          }
#pragma unroll
          for (int c = 0; c < 7; ++c) az[a][c] = __fadd_rn(__fmul_rn(ay[0][c], ly.l0), __fmul_rn(ay[1][c], ly.l1));
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) v[c] = __fadd_rn(__fmul_rn(az[0][c], lz.l0), __fmul_rn(az[1][c], lz.l1));
      }
    }
    if (dz > 0) __syncthreads();
    if (VEC) {
      const int g = threadIdx.x >> 6;
      const float* fp = has_prev ? q.flow_prev + (int64_t)n * 6 * V + (int64_t)d * HW : nullptr;
      // planes g, g+4, g+8 of {flow_prev 0..5, mask_prev, img0, img1}
      tile_load(s[g], fp ? fp + (int64_t)g * V : nullptr, tio);
      const int k1 = g + 4;
      tile_load(s[k1], k1 < 6 ? (fp ? fp + (int64_t)k1 * V : nullptr)
                              : (k1 == 6 ? (has_prev ? q.mask_prev + plane : nullptr) : (SN == 0 ? nullptr : q.img0 + plane)), tio);
      if (g == 0) tile_load(s[8], SN == 0 ? nullptr : q.img1 + plane, tio);
    } else {
      load_planes<9, false>(s, [&](int k) -> const float* {
        if (k < 6) return has_prev ? q.flow_prev + (int64_t)n * 6 * V + (int64_t)k * V + (int64_t)d * HW : nullptr;
        if (k == 6) return has_prev ? q.mask_prev + plane : nullptr;
        if (SN == 0) return nullptr;
        return (k == 7 ? q.img0 : q.img1) + plane; }, h0, w0, H, W);
    }
    __syncthreads();
    if (ok) {
      // ---- flow / mask accumulation (fp32)
      float f[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const float fd = __fmul_rn(v[c], (float)SH);
        f[c] = has_prev ? __fadd_rn(s[c][lane][wl], fd) : fd;
      }
      const float m = has_prev ? __fadd_rn(s[6][lane][wl], v[6]) : v[6];
      const float i0v = s[7][lane][wl], i1v = s[8][lane][wl];
      // ---- warps
      const float lh = __ldg(q.lin_h + h), ld = __ldg(q.lin_d + d), lw = __ldg(q.lin_w + w);
      const Trilin t0 = trilin_setup(f[0], f[1], f[2], lh, ld, lw, D, H, W, P.hs, P.ref_mode);
      const Trilin t1 = trilin_setup(f[3], f[4], f[5], lh, ld, lw, D, H, W, P.hs, P.ref_mode);
      const float a = trilin_sample<FMA>(q.img0 + (int64_t)n * V, t0, W, HW);
      const float b = trilin_sample<FMA>(q.img1 + (int64_t)n * V, t1, W, HW);
      float ms = 0.0f, mg = 0.0f;
      if (need_m) {
        ms = sigmoidf_ref(m);
        mg = __fadd_rn(__fmul_rn(a, ms), __fmul_rn(b, __fsub_rn(1.0f, ms)));
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) s[c][lane][wl] = f[c];
      s[6][lane][wl] = m; s[7][lane][wl] = mg; s[8][lane][wl] = ms;
      if (SN == 1) {
        const float c11[11] = {i0v, i1v, a, b, m, f[0], f[1], f[2], f[3], f[4], f[5]};
        store_pack_row(q.pack_out + ((((int64_t)n * D + d) * H + h) * W + w) * 16, c11);
      } else if (SN == 2) {
        const float c11[11] = {i0v, i1v, a, b, m, f[0], f[1], f[2], f[3], f[4], f[5]};
#pragma unroll
        for (int c = 0; c < 11; ++c) sp[dz * 11 + c][lane][wl] = c11[c];
      }
    }
    __syncthreads();
    if (VEC) {
      const int g = threadIdx.x >> 6;
      float* fo = q.flow_out + (int64_t)n * 6 * V + (int64_t)d * HW;
      tile_store(s[g], fo + (int64_t)g * V, tio);
      const int k1 = g + 4;
      tile_store(s[k1], k1 < 6 ? fo + (int64_t)k1 * V : (k1 == 6 ? q.mask_out + plane : (q.merged ? q.merged + plane : nullptr)), tio);
      if (g == 0) tile_store(s[8], q.mask_sig ? q.mask_sig + plane : nullptr, tio);
    } else {
      store_planes<9, false>(s, [&](int k) -> float* {
        if (k < 6) return q.flow_out + (int64_t)n * 6 * V + (int64_t)k * V + (int64_t)d * HW;
        float* b = k == 6 ? q.mask_out : (k == 7 ? q.merged : q.mask_sig);
        return b ? b + plane : nullptr; }, h0, w0, H, W);
    }
  }
  if (SN == 2) {
    // 2x2x2 mean == F.interpolate(., 0.5): nested W, H, D like ATen; flow channels additionally * 0.5
    __syncthreads();
    if (threadIdx.x < 64) {
      const int ph = threadIdx.x >> 2, pw = threadIdx.x & 3;
      const int oh = h0 / 2 + ph, ow = w0 / 2 + pw;
      if (oh < H / 2 && ow < W / 2) {
        float c11[11];
#pragma unroll
        for (int c = 0; c < 11; ++c) {
          float rz[2];
#pragma unroll
          for (int dz = 0; dz < 2; ++dz) {
            const float (*pl)[T3P] = sp[dz * 11 + c];
            const float r0 = __fadd_rn(__fmul_rn(pl[2 * ph][2 * pw], 0.5f), __fmul_rn(pl[2 * ph][2 * pw + 1], 0.5f));
            const float r1 = __fadd_rn(__fmul_rn(pl[2 * ph + 1][2 * pw], 0.5f), __fmul_rn(pl[2 * ph + 1][2 * pw + 1], 0.5f));
            rz[dz] = __fadd_rn(__fmul_rn(r0, 0.5f), __fmul_rn(r1, 0.5f));
          }
          float r = __fadd_rn(__fmul_rn(rz[0], 0.5f), __fmul_rn(rz[1], 0.5f));
          if (c >= 5) r = __fmul_rn(r, 0.5f);
          c11[c] = r;
        }
        store_pack_row(q.pack_out + ((((int64_t)n * (D / 2) + d0 / 2) * (H / 2) + oh) * (W / 2) + ow) * 16, c11);
      }
    }
  }
}

}  // namespace ofsv

using namespace ofsv;

extern "C" int ofsv_block_finish_3d(const float* head, int Cs, const float* flow_prev, const float* mask_prev,
                                    const float* img0, const float* img1, const float* lin_h, const float* lin_d,
                                    const float* lin_w, float* flow_out, float* mask_out, float* merged, float* mask_sig,
                                    void* pack_out, int N, int D, int H, int W, int scale_head, int scale_next,
                                    int ref_mode, void* stream) {
  OFSV_REQUIRE(N >= 0 && D >= 1 && H >= 1 && W >= 1 && (int64_t)D * H * W < (1ll << 31), "ofsv_block_finish_3d: bad shape");
  OFSV_REQUIRE(scale_head == 1 || scale_head == 2 || scale_head == 4, "ofsv_block_finish_3d: scale_head %d not in {1,2,4}", scale_head);
  OFSV_REQUIRE(scale_next == 0 || scale_next == 1 || scale_next == 2, "ofsv_block_finish_3d: scale_next %d not in {0,1,2}", scale_next);
  OFSV_REQUIRE(D % scale_head == 0 && H % scale_head == 0 && W % scale_head == 0, "ofsv_block_finish_3d: dims must be multiples of scale_head");
  OFSV_REQUIRE(scale_next != 2 || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "ofsv_block_finish_3d: dims must be even for scale_next = 2");
  OFSV_REQUIRE(Cs >= 8 && Cs % 4 == 0, "ofsv_block_finish_3d: head channel stride must be >= 8 and a multiple of 4");
  OFSV_REQUIRE(ref_mode == OFSV_REF_CPU || ref_mode == OFSV_REF_CUDA, "ofsv_block_finish_3d: bad ref_mode");
  OFSV_REQUIRE((flow_prev == nullptr) == (mask_prev == nullptr), "ofsv_block_finish_3d: flow_prev and mask_prev go together");
  if (N == 0) return OFSV_OK;
  OFSV_REQUIRE(head && img0 && img1 && lin_h && lin_d && lin_w && flow_out && mask_out, "ofsv_block_finish_3d: null pointer");
  OFSV_REQUIRE((scale_next == 0) == (pack_out == nullptr), "ofsv_block_finish_3d: pack_out must be given iff scale_next != 0");
  OFSV_REQUIRE(aligned16(head) && (!pack_out || aligned16(pack_out)), "ofsv_block_finish_3d: head / pack_out must be 16-byte aligned");
  const Warp3dParams P = make_warp3d_params(N, 1, D, H, W, ref_mode);
  const int tdz = scale_next == 2 ? 2 : 1;
  const dim3 grid((unsigned)cdiv(W, T3W), (unsigned)cdiv(H, T3H), (unsigned)(N * (D / tdz)));
  if (grid.z > 65535u) { set_error("ofsv_block_finish_3d: N*D=%u exceeds grid.z", grid.z); return OFSV_ENOSUP; }
  bool vec = (W % 4 == 0);
  const void* ptrs[] = {flow_prev, mask_prev, img0, img1, flow_out, mask_out, merged, mask_sig};
  for (const void* p : ptrs) vec = vec && (p == nullptr || aligned16(p));
  FinishPtrs q{head, flow_prev, mask_prev, img0, img1, lin_h, lin_d, lin_w, flow_out, mask_out, merged, mask_sig,
               reinterpret_cast<__nv_bfloat16*>(pack_out)};
  cudaStream_t st = (cudaStream_t)stream;
  const bool fma = ref_mode == OFSV_REF_CUDA;
#define GO4(SH, SN)                                                                         \
  do {                                                                                      \
    if (vec) { if (fma) block_finish_3d_kernel<SH, SN, true, true><<<grid, 256, 0, st>>>(q, P, Cs);    \
               else block_finish_3d_kernel<SH, SN, true, false><<<grid, 256, 0, st>>>(q, P, Cs); }     \
    else     { if (fma) block_finish_3d_kernel<SH, SN, false, true><<<grid, 256, 0, st>>>(q, P, Cs);   \
               else block_finish_3d_kernel<SH, SN, false, false><<<grid, 256, 0, st>>>(q, P, Cs); }    \
  } while (0)
#define GO(SH)                                                  \
  do {                                                          \
    if (scale_next == 0) GO4(SH, 0); else if (scale_next == 1) GO4(SH, 1); else GO4(SH, 2); \
  } while (0)
  if (scale_head == 1) GO(1); else if (scale_head == 2) GO(2); else GO(4);
#undef GO
#undef GO4
  return check_launch("block_finish_3d_kernel");
}
