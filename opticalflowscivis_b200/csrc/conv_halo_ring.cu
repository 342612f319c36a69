// PLANE-RING variant of the halo-reuse tcgen05 convolution (conv_halo.cu), used for the layers with few taps per super-tile
// (the space-to-depth conv0 layers): z-fastest tile runs that share boundary planes, per-plane release, resident weights.
// tcgen05 implicit-GEMM convolution with shared-memory HALO REUSE for the stride-1 layers of an IFBlock
// (3^d convs of convblock0..3, and the 2^d-phase form of ConvTranspose(4,2,1): every tap offset is in {-1,0,+1}^d).
//
// conv_tc.cu re-loads the 128 x KC activation tile from L2 for every filter tap (27x re-read, ~650 KB of L2->SM traffic
// per 128 output positions): it is L2/latency bound at ~25 % tensor-pipe utilisation (profiles/r01_*).  Here each input
// plane of a super-tile is loaded ONCE and every tap is just a shifted UMMA shared-memory descriptor:
//
//   super-tile  = 16(h) x 8(w) x TD(d) output positions of one sample (TD in {1,2,4} accumulators of 128 rows)
//   halo plane  = the (16+2) x (8+2) x KC input box of one input depth slice, ONE 5-D TMA load (OOB zero-fill = conv
//                 padding), SWIZZLE_128B/64B/32B, laid out [18][10] rows of KC*2 bytes
//   tap (dz,dy,dx) for output slice j  ->  A descriptor start = plane[j+dz-dzmin] + ((dy+1)*10 + (dx+1)) * rowbytes,
//                 8-row-group stride SBO = 10 rows.  tests/umma_probe.cu verified on B200 that UMMA K-major swizzled
//                 descriptors address shared memory by absolute address bits (base_offset = 0), so neither the start nor
//                 SBO has to be a multiple of the 1024 B swizzle atom.
//   B tile      = W[pass][tap][chunk] ([Cout_w][KC], K-major) streamed once per super-tile through a small TMA ring and
//                 shared by the TD accumulators.
//   passes      = 1 (conv) or 2^d (ConvTranspose output parities): the planes stay resident across all passes.
//   TMEM        = TD x Cout_w fp32 columns per pass, double-buffered when it fits so the epilogue of pass i overlaps
//                 the MMAs of pass i+1; CTAs are persistent over super-tiles (grid = #SMs).
//   plane ring  = every CTA walks a contiguous run of super-tiles in z-FASTEST order, so consecutive super-tiles of a column
//                 share their (dzmax - dzmin) boundary planes: the np = TD + dz-span plane slots form a ring indexed by the
//                 input depth, only TD new planes are loaded per super-tile, and (single-pass layers, taps sorted by dz) a
//                 plane's slot is handed back to the TMA producer right after the last tap that reads it — the next
//                 super-tile's planes stream in underneath the current super-tile's MMAs instead of after them.
// L2->SM traffic per 128 outputs drops from ~650 KB to ~(23 KB x (TD+2)/TD + 216 KB/TD).
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

namespace ofsv {

__device__ __forceinline__ uint32_t ring_pack2_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

constexpr int HT_H = 16, HT_W = 8, HP_H = HT_H + 2, HP_W = HT_W + 2, HP_ROWS = HP_H * HP_W;  // 180 halo rows
constexpr int H_MAX_PLANES = 8;
constexpr int H_MAX_BSTAGES = 32;
#ifndef OFSV_RING_MMA_WARPS
#define OFSV_RING_MMA_WARPS 8
#endif
constexpr int H_MMA_WARPS = OFSV_RING_MMA_WARPS, H_EPI_WARPS = 8, H_EPI_W0 = 1 + H_MMA_WARPS;
constexpr int H_THREADS = 32 * (H_EPI_W0 + H_EPI_WARPS);   // warp 0 TMA, then the MMA issuer warps (slices x tap groups), then 8 epilogue warps

struct RingParams {
  int N, Do, Ho, Wo, Dy, Hy, Wy, Cout_s, Cout_w;
  int out_stride, nphase, ntaps, nkc;
  int td, np, dzmin;                 // output slices per super-tile, input planes per super-tile, min dz over taps
  int tiles_w, tiles_h, tiles_d;     // super-tile grid per sample
  int nb, plane_stride, b_stride;    // B ring depth, smem strides (bytes, multiples of 1024)
  int nbuf, acc_stride;              // TMEM accumulator double buffering
  int has_prelu, has_residual, out_f32, shuffle;
  int nd, out_s2d;
  int dzspan;                        // dzmax - dzmin
  int sliding;                       // 1: z-fastest contiguous tile runs with shared boundary planes and early plane release
  int b_resident;                    // all weight tiles fit the ring: loaded once per CTA, never recycled
  int ng;                            // tap groups (resident weights only): the taps of a super-tile are split over ng issuer
                                     // warps per slice, each with its own partial accumulator, summed by the epilogue
  uint8_t tap_order[OFSV_MAX_TAPS];   // per pass: taps in ascending dz (position -> original tap index)
  uint8_t rel_after[OFSV_MAX_TAPS];   // last pass only: bitmask of non-shared plane indices q < td whose last reader is this position
  int8_t tap_off[OFSV_MAX_TAPS][4];
};

template <int KC>
__global__ void __launch_bounds__(H_THREADS, 1)
    conv_halo_ring_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const RingParams p,
                     const float* __restrict__ bias, const float* __restrict__ prelu, const void* __restrict__ residual,
                     void* __restrict__ y, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int ROWB = KC * 2;
#ifdef OFSV_RING_TRACE   // compile-time: clock reads inside the MMA issue loop cost ~10 % on the 27-tap layers
  const bool trace = dbg != nullptr && blockIdx.x == 0;   // OFSV_HALO_TRACE: per-super-tile clock64 timeline of CTA 0
#else
  constexpr bool trace = false;
  (void)dbg;
#endif
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nplanes = p.np * p.nkc;
  uint8_t* sP = smem;
  uint8_t* sB = smem + nplanes * p.plane_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + p.nb * p.b_stride);
  uint64_t* plane_full = bars;                            // [H_MAX_PLANES]  (indexed by ring position)
  uint64_t* plane_empty = bars + H_MAX_PLANES;            // [H_MAX_PLANES]
  uint64_t* b_full = plane_empty + H_MAX_PLANES;          // [H_MAX_BSTAGES]
  uint64_t* b_empty = b_full + H_MAX_BSTAGES;             // [H_MAX_BSTAGES]
  uint64_t* acc_full = b_empty + H_MAX_BSTAGES;           // [2]
  uint64_t* acc_empty = acc_full + 2;                     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sBias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));   // [128], 16 B aligned
  float* sPrelu = sBias + 128;                              // [128]
  uint32_t* sTap = reinterpret_cast<uint32_t*>(sPrelu + 128); // [OFSV_MAX_TAPS]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = p.tiles_w * p.tiles_h * p.tiles_d * p.N;
  // sliding: contiguous run of super-tiles of this CTA in z-fastest order, L = ((n * tiles_h + ty) * tiles_w + tx) * tiles_d + tz.
  // otherwise: tiles blockIdx.x + k * gridDim.x of the x-fastest order (neighbouring CTAs work on neighbouring tiles), every
  // tile loads all of its planes.  Both are expressed as a run [L_begin, L_end) of per-CTA sequence numbers.
  const int L_begin = p.sliding ? (int)(((long long)blockIdx.x * total) / gridDim.x) : 0;
  const int L_end = p.sliding ? (int)(((long long)(blockIdx.x + 1) * total) / gridDim.x)
                              : (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  struct TileC { int tx, ty, tz, n; };
  auto decode = [&](int L) {
    TileC c;
    if (p.sliding) {
      c.tz = L % p.tiles_d; L /= p.tiles_d;
      c.tx = L % p.tiles_w; L /= p.tiles_w;
      c.ty = L % p.tiles_h;
      c.n = L / p.tiles_h;
    } else {
      int r = (int)blockIdx.x + L * (int)gridDim.x;
      c.tx = r % p.tiles_w; r /= p.tiles_w;
      c.ty = r % p.tiles_h; r /= p.tiles_h;
      c.tz = r % p.tiles_d;
      c.n = r / p.tiles_d;
    }
    return c;
  };

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < H_MAX_PLANES; ++i) { mbar_init(&plane_full[i], 1); mbar_init(&plane_empty[i], H_MMA_WARPS); }
    for (int i = 0; i < H_MAX_BSTAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], H_MMA_WARPS); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], H_MMA_WARPS); mbar_init(&acc_empty[i], H_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t bcount = 0, bs = 0, bphase = 0;
      // plane cursor: planes are issued strictly in (tile, plane index) order; it may run at most one tile ahead of the B cursor
      int cL = L_begin, cq = 0, c_ring0 = 0;               // tile / plane index of the next plane to issue, ring position of its plane 0
      uint32_t used = 0, eph = 0;                           // per ring slot: loaded before? / parity of the next empty-wait
      auto tile_is_start = [&](int L) { return !p.sliding || L == L_begin || (L % p.tiles_d) == 0; };
      auto cursor_norm = [&]() {                            // skip the planes the tile shares with its predecessor
        if (cL < L_end && cq == 0 && !tile_is_start(cL)) cq = p.dzspan;
      };
      TileC cc = decode(cL < L_end ? cL : L_begin);         // coordinates of the cursor's tile (decoded once per tile)
      auto cursor_try = [&](bool block) -> bool {           // issue the plane under the cursor if its slot is free
        int r = c_ring0 + cq; if (r >= p.np) r -= p.np;
        if ((used >> r) & 1u) {
          if (block) mbar_wait(&plane_empty[r], (eph >> r) & 1u, 64);
          else if (!mbar_try(smem_u32(&plane_empty[r]), (eph >> r) & 1u)) return false;
          eph ^= 1u << r;
        }
        used |= 1u << r;
        const TileC c = cc;
        mbar_expect_tx(&plane_full[r], p.nkc * HP_ROWS * ROWB);
        for (int kc = 0; kc < p.nkc; ++kc)
          tma_load_5d(&tmA, &plane_full[r], sP + (r * p.nkc + kc) * p.plane_stride, kc * KC, c.tx * HT_W - 1, c.ty * HT_H - 1,
                      c.tz * p.td + p.dzmin + cq, c.n);
        if (++cq == p.np) {                                  // next tile: its plane 0 sits td ring positions further
          cq = 0; ++cL;
          if (cL < L_end) {
            cc = decode(cL);
            if (tile_is_start(cL)) c_ring0 = 0; else { c_ring0 += p.td; if (c_ring0 >= p.np) c_ring0 -= p.np; }
          }
          cursor_norm();
        }
        return true;
      };
      for (int L = L_begin; L < L_end; ++L) {
        for (int pass = 0; pass < p.nphase; ++pass)
          for (int tp = 0; tp < p.ntaps; ++tp) {
            const int t = p.tap_order[pass * p.ntaps + tp];
            // planes this tap needs must have been issued before its weights (else the MMA warps could wait on a plane that
            // is never requested while this thread blocks on the weight ring)
            const int q_need = ((p.nphase == 1 && p.sliding) ? (p.tap_off[t][0] - p.dzmin) : p.dzspan) + p.td - 1;
            while (cL == L && cq <= q_need) cursor_try(true);
            for (int kc = 0; kc < p.nkc; ++kc, ++bcount) {
              if (p.b_resident && L > L_begin) {               // weights already in shared memory: only keep planes streaming
                while (cL < L_end && cL <= L + 1 && cursor_try(false)) {}
                continue;
              }
              if (bcount >= (uint32_t)p.nb) {
                if (!p.sliding) {
                  mbar_wait(&b_empty[bs], bphase ^ 1u, 100);
                } else {
                  const uint32_t addr = smem_u32(&b_empty[bs]);
                  while (!mbar_try(addr, bphase ^ 1u)) {     // while the ring is full: stream planes of this / the next tile
                    if (!(cL < L_end && cL <= L + 1 && cursor_try(false))) __nanosleep(64);
                  }
                }
              }
              mbar_expect_tx(&b_full[bs], p.Cout_w * ROWB);
              tma_load_2d(&tmB, &b_full[bs], sB + bs * p.b_stride, 0, ((pass * p.ntaps + t) * p.nkc + kc) * p.Cout_w);
              if (++bs == (uint32_t)p.nb) { bs = 0; bphase ^= 1u; }
              if (p.sliding) while (cL < L_end && cL <= L + 1 && cursor_try(false)) {}
            }
          }
      }
    }
  } else if (warp < H_EPI_W0) {
    const int iss = warp - 1;   // issuer index: owns output slices j with j % H_MMA_WARPS == iss
    // ================= MMA issuer: the whole warp walks the (warp-uniform) loops, one elected lane issues =================
    // The issuing thread is latency-bound on its own scalar code, so everything per operand tile is table-driven:
    // sTap[pass*ntaps + t] = (descriptor offset of the tap inside a plane) | (dz - dzmin) << 16, ring indices are counters.
    for (int i = lane; iss == 0 && i < p.nphase * p.ntaps; i += 32) {
      const int8_t* off = p.tap_off[(i / p.ntaps) * p.ntaps + p.tap_order[i]];
      sTap[i] = ((uint32_t)(((off[1] + 1) * HP_W + (off[2] + 1)) * ROWB) >> 4) | ((uint32_t)(off[0] - p.dzmin) << 16);
    }
    asm volatile("bar.sync 2, %0;" ::"n"(32 * H_MMA_WARPS) : "memory");
    const uint32_t leader = elect_one_sync();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Cout_w >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_hi = kmajor_desc_hi<KC>(HP_W * ROWB);      // SBO = 10 halo rows
    const uint32_t b_hi = kmajor_desc_hi<KC>(8 * ROWB);
    const uint32_t plane_lo0 = kmajor_desc_lo(smem_u32(sP)), plane_step = (uint32_t)p.plane_stride >> 4;
    const uint32_t b_lo0 = kmajor_desc_lo(smem_u32(sB)), b_step = (uint32_t)p.b_stride >> 4;
    const int ntap_total = p.ntaps;
    const int nj = p.td;                     // Do % td == 0 (host)
    // tap groups: the single issuing thread needs ~600 cycles of scalar work per (tap, chunk) whatever the MMA sizes, so a
    // short super-tile (few slices, small N) is issue-bound; with ng > 1 (H_MMA_WARPS = ng * td) warp iss issues slice
    // iss % td for the taps of group iss / td only, into its own accumulator
    const int grp = p.ng > 1 ? iss / nj : 0, tg = ntap_total / p.ng;
    const int tbeg = (p.ng > 1 && grp >= p.ng) ? ntap_total : grp * tg;          // warps beyond ng * td issue nothing
    const int jbeg = p.ng > 1 ? iss % nj : iss, jstep = p.ng > 1 ? nj : H_MMA_WARPS;
    const uint32_t jmask = (1u << nj) - 1u;
    uint32_t bs = 0, bphase = 0;             // B ring slot / parity
    uint32_t buf = 0, acc_use = 0;           // accumulator buffer, number of uses so far
    uint32_t fph = 0;                        // per ring slot: parity of the next full-wait
    int ring0 = 0;
    int iter = 0;
    for (int L = L_begin; L < L_end; ++L, ++iter) {
      const int tz = L % p.tiles_d;      // (only meaningful when sliding)
      const bool col_start = !p.sliding || L == L_begin || tz == 0, col_end = !p.sliding || tz == p.tiles_d - 1;
      if (col_start) ring0 = 0; else { ring0 += p.td; if (ring0 >= p.np) ring0 -= p.np; }
      uint32_t waited = col_start ? 0u : ((1u << p.dzspan) - 1u);   // plane indices q already resident (shared with the previous tile)
      const long long t_in = trace ? clock64() : 0;
      long long tw_plane = 0, tw_b = 0, tw_acc = 0;
      for (int pass = 0; pass < p.nphase; ++pass) {
        { const long long t0 = trace ? clock64() : 0;
          if (acc_use >= (uint32_t)p.nbuf) mbar_wait(&acc_empty[buf], ((acc_use / p.nbuf) - 1) & 1);
          if (trace) tw_acc += clock64() - t0; }
        const uint32_t acc0 = tmem_base + buf * p.acc_stride;
        const uint32_t* taps = sTap + pass * ntap_total;
        const bool last_pass = pass == p.nphase - 1;
        for (int t = 0; t < ntap_total; ++t) {
          const uint32_t te = taps[t];
          const uint32_t tap16 = te & 0xFFFFu, dzr = te >> 16;
          // planes this tap touches for the first time in this super-tile: q = dzr .. dzr + nj - 1
          const long long tp0 = trace ? clock64() : 0;
          uint32_t need = (jmask << dzr) & ~waited;
          while (need) {
            const int q = __ffs(need) - 1;
            int r = ring0 + q; if (r >= p.np) r -= p.np;
            mbar_wait(&plane_full[r], (fph >> r) & 1u);
            fph ^= 1u << r;
            need &= need - 1;
          }
          waited |= jmask << dzr;
          if (trace) tw_plane += clock64() - tp0;
          if (p.ng > 1 && (t < tbeg || t >= tbeg + tg)) {          // another group's tap (resident weights: slot = (t, kc))
            bs += (uint32_t)p.nkc;
            if (bs >= (uint32_t)p.nb) bs -= (uint32_t)p.nb;
          } else
          for (int kc = 0; kc < p.nkc; ++kc) {
            const long long tb0 = trace ? clock64() : 0;
            mbar_wait(&b_full[bs], p.b_resident ? 0u : bphase);
            if (trace) tw_b += clock64() - tb0;
            tcgen05_fence_after();
            if (leader) {
              const uint32_t b_lo = b_lo0 + bs * b_step;
              const uint32_t first = ((t - tbeg) | kc) ? 1u : 0u;
              for (int j = jbeg; j < nj; j += jstep) {
                int r = ring0 + j + (int)dzr; if (r >= p.np) r -= p.np;
                const uint32_t aj = plane_lo0 + (uint32_t)(r * p.nkc + kc) * plane_step + tap16, dj = acc0 + (grp * nj + j) * p.Cout_w;
                umma_bf16_lohi(dj, aj, a_hi, b_lo, b_hi, idesc, first);
#pragma unroll
                for (int k = 1; k < KC / 16; ++k) umma_bf16_lohi(dj, aj + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, 1u);
              }
              if (!p.b_resident) tcgen05_commit(&b_empty[bs]);
            }
            __syncwarp();
            if (++bs == (uint32_t)p.nb) { bs = 0; bphase ^= 1u; }
          }
          if (last_pass) {
            // hand plane slots back to the producer: planes whose last reader was this tap; at the end of a column also the
            // planes that would otherwise be kept for the next super-tile
            uint32_t rel = p.rel_after[t];
            if (t == ntap_total - 1 && col_end) rel |= ((1u << p.np) - 1u) & ~((1u << p.td) - 1u);
            while (rel) {
              const int q = __ffs(rel) - 1;
              int r = ring0 + q; if (r >= p.np) r -= p.np;
              if (leader) tcgen05_commit(&plane_empty[r]);
              rel &= rel - 1;
            }
            __syncwarp();
          }
        }
        if (leader) tcgen05_commit(&acc_full[buf]);
        __syncwarp();
        ++acc_use;
        if (++buf == (uint32_t)p.nbuf) buf = 0;
      }
      if (trace && iter < 16 && lane == 0 && iss == 0) {
        dbg[iter * 16 + 2] = t_in; dbg[iter * 16 + 3] = clock64();
        dbg[iter * 16 + 4] = tw_plane; dbg[iter * 16 + 5] = tw_b; dbg[iter * 16 + 6] = tw_acc;
      }
    }
  } else {
    // ================= epilogue: 8 warps, two per TMEM lane quarter, splitting the (slice, 16-column chunk) list ==========
    const int q = warp & 3, half = (warp - H_EPI_W0) >> 2;
    const int row = q * 32 + lane;
    const int rx = row & 7, ry = row >> 3;
    for (int i = threadIdx.x - 32 * H_EPI_W0; i < p.Cout_w; i += 32 * H_EPI_WARPS) {
      sBias[i] = __ldg(bias + i);
      sPrelu[i] = p.has_prelu ? __ldg(prelu + i) : 1.0f;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * H_EPI_WARPS) : "memory");
    const int nch = p.Cout_w >> 4;
    const __nv_bfloat16* resb = reinterpret_cast<const __nv_bfloat16*>(residual);
    uint32_t acc_it = 0;
    for (int L = L_begin; L < L_end; ++L) {
      const TileC tc = decode(L);
      const int tx = tc.tx, ty = tc.ty, tz = tc.tz, n = tc.n;
      const int ox = tx * HT_W + rx, oy = ty * HT_H + ry;
      const bool valid_xy = ox < p.Wo && oy < p.Ho;
      const int nitems = min(p.td, p.Do - tz * p.td) * nch;
      for (int pass = 0; pass < p.nphase; ++pass, ++acc_it) {
        const int buf = acc_it % p.nbuf;
        const int pz = (pass >> 2) & 1, py = (pass >> 1) & 1, px = pass & 1;
        // row offset of output slice j (non-shuffled layers)
        auto row_off = [&](int j) -> int64_t {
          const int oz = tz * p.td + j;
          if (p.out_s2d) return s2d_row(p.nd, n, oz, oy, ox, p.Dy, p.Hy, p.Wy) * p.Cout_s;
          return ((((int64_t)n * p.Dy + oz * p.out_stride + pz) * p.Hy + oy * p.out_stride + py) * p.Wy + ox * p.out_stride + px) * p.Cout_s;
        };
        uint4 rnext[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        const bool res_bf16 = p.has_residual && !p.shuffle, res_f32 = p.has_residual && p.shuffle;
        // depth-to-space heads: element offset of output parity ph (z,y,x bits) of slice j, row (oy, ox)
        auto shuf_off = [&](int j, int ph) -> int64_t {
          const int oz = tz * p.td + j;
          return ((((int64_t)n * p.Dy + oz * 2 + ((ph >> 2) & 1)) * p.Hy + oy * 2 + ((ph >> 1) & 1)) * p.Wy + ox * 2 + (ph & 1)) * p.Cout_s;
        };
        const float* resf = reinterpret_cast<const float*>(residual);
        float4 fnext[4];
        auto load_state = [&](int it) {       // fp32 state rows (fm_prev) of the two parities of chunk `it`
          const int j = it / nch, c0 = (it - j * nch) << 4;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float4* r = reinterpret_cast<const float4*>(resf + shuf_off(j, (c0 >> 3) + e));
            fnext[2 * e] = __ldg(r); fnext[2 * e + 1] = __ldg(r + 1);
          }
        };
        if (res_f32 && valid_xy && half < nitems) load_state(half);
        if (res_bf16 && valid_xy && half < nitems) {
          const __nv_bfloat16* rp = resb + row_off(half / nch) + (half % nch) * 16;
          rnext[0] = __ldg(reinterpret_cast<const uint4*>(rp)); rnext[1] = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
        }
        const long long te0 = clock64();
        mbar_wait(&acc_full[buf], (acc_it / p.nbuf) & 1, 128);   // long waits: sleep between polls
        tcgen05_fence_after();
        const long long te1 = clock64();
        for (int it = half; it < nitems; it += 2) {
          const int j = it / nch, c0 = (it - j * nch) << 4;
          const uint4 rcur0 = rnext[0], rcur1 = rnext[1];
          float4 fcur[4];
          if (res_f32) {
#pragma unroll
            for (int e = 0; e < 4; ++e) fcur[e] = fnext[e];
            if (valid_xy && it + 2 < nitems) load_state(it + 2);
          }
          if (res_bf16 && valid_xy && it + 2 < nitems) {       // prefetch the next chunk's residual row
            const int jn = (it + 2) / nch;
            const __nv_bfloat16* rp = resb + row_off(jn) + ((it + 2) - jn * nch) * 16;
            rnext[0] = __ldg(reinterpret_cast<const uint4*>(rp)); rnext[1] = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
          }
          float v[16];
          tmem_ld16(tmem_base + buf * p.acc_stride + j * p.Cout_w + c0 + ((uint32_t)(q * 32) << 16), v);
          for (int g = 1; g < p.ng; ++g) {      // partial sums of the other tap groups, fixed order
            float u[16];
            tmem_ld16(tmem_base + buf * p.acc_stride + (g * p.td + j) * p.Cout_w + c0 + ((uint32_t)(q * 32) << 16), u);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] += u[e];
          }
          if (!valid_xy) continue;
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {          // 16 B shared loads: the epilogue competes with the UMMA operand fetch
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + e4 * 4);
            const float4 p4 = *reinterpret_cast<const float4*>(sPrelu + c0 + e4 * 4);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = v[e4 * 4 + e] + bb[e];
              v[e4 * 4 + e] = fmaxf(a, 0.0f) + pp[e] * fminf(a, 0.0f);
            }
          }
          if (res_f32) {            // flow/mask state accumulation: fm = fm_prev + head (Flow-3D/model/IFNet.py:169-170)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[4 * e] = __fadd_rn(fcur[e].x, v[4 * e]); v[4 * e + 1] = __fadd_rn(fcur[e].y, v[4 * e + 1]);
              v[4 * e + 2] = __fadd_rn(fcur[e].z, v[4 * e + 2]); v[4 * e + 3] = __fadd_rn(fcur[e].w, v[4 * e + 3]);
            }
          }
          if (res_bf16) {
            const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&rcur0);
            const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&rcur1);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[2 * e] += __low2float(h0[e]); v[2 * e + 1] += __high2float(h0[e]);
              v[8 + 2 * e] += __low2float(h1[e]); v[8 + 2 * e + 1] += __high2float(h1[e]);
            }
          }
          if (p.shuffle) {
            // depth-to-space heads: columns = [output parity (z,y,x)][8 channels]; this chunk holds parities c0/8 and c0/8+1
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int64_t yo = shuf_off(j, (c0 >> 3) + e);
              if (p.out_f32) {
                float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + yo);
                o[0] = make_float4(v[8 * e], v[8 * e + 1], v[8 * e + 2], v[8 * e + 3]);
                o[1] = make_float4(v[8 * e + 4], v[8 * e + 5], v[8 * e + 6], v[8 * e + 7]);
              } else {
                uint4 w;
                w.x = ring_pack2_bf16(v[8 * e], v[8 * e + 1]); w.y = ring_pack2_bf16(v[8 * e + 2], v[8 * e + 3]);
                w.z = ring_pack2_bf16(v[8 * e + 4], v[8 * e + 5]); w.w = ring_pack2_bf16(v[8 * e + 6], v[8 * e + 7]);
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + yo) = w;
              }
            }
          } else {
            const int64_t yo = row_off(j) + c0;
            const int nstore = min(16, p.Cout_s - c0);
            if (p.out_f32) {
              float* o = reinterpret_cast<float*>(y) + yo;
              for (int e = 0; e < nstore; e += 4) *reinterpret_cast<float4*>(o + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
            } else {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(y) + yo;
              for (int e = 0; e < nstore; e += 8) {
                uint4 w;
                w.x = ring_pack2_bf16(v[e], v[e + 1]); w.y = ring_pack2_bf16(v[e + 2], v[e + 3]);
                w.z = ring_pack2_bf16(v[e + 4], v[e + 5]); w.w = ring_pack2_bf16(v[e + 6], v[e + 7]);
                *reinterpret_cast<uint4*>(o + e) = w;
              }
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[buf])) : "memory");
        if (trace && warp == H_EPI_W0 && lane == 0 && acc_it < 16) { dbg[acc_it * 16 + 8] = te0; dbg[acc_it * 16 + 9] = te1; dbg[acc_it * 16 + 10] = clock64(); }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <int KC>
static int launch_halo_ring(const RingParams& P, const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias,
                       const float* prelu, const void* residual, void* y, int grid, size_t smem, cudaStream_t st) {
  static std::atomic<uint64_t> attr_done{0};
  if (int e = ensure_dyn_smem(attr_done, conv_halo_ring_kernel<KC>, 227 * 1024, "ofsv_conv_halo(ring)")) return e;
  conv_halo_ring_kernel<KC><<<grid, H_THREADS, smem, st>>>(tmA, tmB, P, bias, prelu, residual, y, nullptr);
  return check_launch("conv_halo_ring_kernel");
}

}  // namespace ofsv

using namespace ofsv;

// Same contract as ofsv_conv_tc; returns OFSV_ENOSUP (nothing launched) for layers outside this kernel's domain
// (strided input, tap offsets outside {-1,0,1}, Cout_w > 128, planes that do not fit shared memory).
int ofsv::conv_halo_ring(const ofsv_conv_desc* d, const void* x, const void* w, const float* bias, const float* prelu,
                         const void* residual, void* y, void* stream) {
  if (int e = validate_conv_desc(d, "ofsv_conv_halo")) return e;
  if (d->N == 0) return OFSV_OK;
  OFSV_REQUIRE(x && w && bias && y, "ofsv_conv_halo(ring): null pointer");
  OFSV_REQUIRE(!d->has_prelu || prelu, "ofsv_conv_halo(ring): has_prelu without prelu slopes");
  OFSV_REQUIRE(!d->has_residual || residual, "ofsv_conv_halo(ring): has_residual without residual");
  OFSV_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && (!residual || aligned16(residual)),
               "ofsv_conv_halo(ring): pointers must be 16-byte aligned");
  if (d->in_dtype != OFSV_BF16) { set_error("ofsv_conv_halo(ring): activations must be bf16"); return OFSV_ENOSUP; }
  if (d->in_stride != 1) { set_error("ofsv_conv_halo(ring): in_stride must be 1"); return OFSV_ENOSUP; }
  if (d->out_shuffle_hfast) { set_error("ofsv_conv_halo(ring): H-fastest depth-to-space output is implemented by the stacked kernel only"); return OFSV_ENOSUP; }
  if (d->Cout_w > 128) { set_error("ofsv_conv_halo(ring): Cout_w=%d > 128", d->Cout_w); return OFSV_ENOSUP; }
  if (d->has_residual && d->out_dtype != OFSV_BF16 && !d->out_shuffle) { set_error("ofsv_conv_halo(ring): residual needs a bf16 output"); return OFSV_ENOSUP; }
  if (d->has_residual && d->out_shuffle && d->out_dtype != OFSV_F32) { set_error("ofsv_conv_halo(ring): depth-to-space residual (flow/mask state) is fp32"); return OFSV_ENOSUP; }
  if (d->out_shuffle && (d->out_shuffle != 8 || d->nphase != 1 || d->Cout_w != 8 * (1 << d->nd) || d->Cout_s != 8)) {
    set_error("ofsv_conv_halo(ring): bad depth-to-space configuration");
    return OFSV_EINVAL;
  }
  if (d->out_s2d && (d->nphase != 1 || d->out_stride != 1 || d->has_residual || d->out_shuffle || d->Hy % 2 || d->Wy % 2 ||
                     (d->nd == 3 && d->Dy % 2))) {
    set_error("ofsv_conv_halo(ring): bad space-to-depth output configuration");
    return OFSV_EINVAL;
  }
  int dzmin = 1, dzmax = -1;
  for (int i = 0; i < d->nphase * d->ntaps; ++i) {
    const int8_t* o = d->tap_off[i];
    if (o[0] < -1 || o[0] > 1 || o[1] < -1 || o[1] > 1 || o[2] < -1 || o[2] > 1) {
      set_error("ofsv_conv_halo(ring): tap offset outside {-1,0,1}");
      return OFSV_ENOSUP;
    }
    dzmin = o[0] < dzmin ? o[0] : dzmin;
    dzmax = o[0] > dzmax ? o[0] : dzmax;
  }
  PFN_encodeTiled encode = get_tensor_map_encoder();
  if (!encode) { set_error("ofsv_conv_halo(ring): cuTensorMapEncodeTiled unavailable"); return OFSV_ECUDA; }

  const int KC = d->Cin_s % 64 == 0 ? 64 : (d->Cin_s % 32 == 0 ? 32 : 16);
  RingParams P;
  memset(&P, 0, sizeof(P));
  P.N = d->N; P.Do = d->Do; P.Ho = d->Ho; P.Wo = d->Wo; P.Dy = d->Dy; P.Hy = d->Hy; P.Wy = d->Wy;
  P.Cout_s = d->Cout_s; P.Cout_w = d->Cout_w; P.out_stride = d->out_stride; P.nphase = d->nphase; P.ntaps = d->ntaps;
  P.nkc = d->Cin_s / KC; P.dzmin = dzmin;
  P.plane_stride = (HP_ROWS * KC * 2 + 1023) & ~1023;
  P.b_stride = (d->Cout_w * KC * 2 + 1023) & ~1023;
  P.tiles_w = (int)cdiv(d->Wo, HT_W); P.tiles_h = (int)cdiv(d->Ho, HT_H);
  P.has_prelu = d->has_prelu; P.has_residual = d->has_residual; P.out_f32 = d->out_dtype == OFSV_F32;
  P.shuffle = d->out_shuffle; P.nd = d->nd; P.out_s2d = d->out_s2d;
  memcpy(P.tap_off, d->tap_off, sizeof(P.tap_off));
  const int sms = device_num_sms();
  const size_t smem_cap = 227 * 1024 - 2048;
  const size_t bar_bytes = (2 * H_MAX_PLANES + 2 * H_MAX_BSTAGES + 4) * 8 + 32 + 2 * 128 * 4 + OFSV_MAX_TAPS * 4;
  // TD: as many output slices per B-tile load as fit (TMEM columns, shared memory) while keeping >= 2 waves of super-tiles
  int td = 0;
  const int forced = 0;
  for (int cand = 4; cand >= 1; cand >>= 1) {
    if (forced && cand != forced && cand > 1) continue;
    if (d->Do % cand) continue;                       // the plane ring assumes full super-tiles along z
    const int np = cand + (dzmax - dzmin);
    if (np * P.nkc > H_MAX_PLANES) continue;
    if (cand * d->Cout_w > 512) continue;
    if ((size_t)np * P.nkc * P.plane_stride + 3 * (size_t)P.b_stride + bar_bytes + 1024 > smem_cap) continue;
    const int64_t nst = (int64_t)P.tiles_w * P.tiles_h * cdiv(d->Do, cand) * d->N;
    if (cand > 1 && nst < 2 * sms && !forced) continue;
    td = cand;
    break;
  }
  if (td == 0) { set_error("ofsv_conv_halo(ring): layer does not fit (Cin_s=%d Cout_w=%d)", d->Cin_s, d->Cout_w); return OFSV_ENOSUP; }
  P.td = td; P.np = td + (dzmax - dzmin); P.dzspan = dzmax - dzmin;
  P.tiles_d = d->Do / td;
  // tap order per pass: ascending dz (stable), so that the lowest planes of a super-tile are finished first; single-pass
  // layers release plane q < td right after the last tap that reads it (dz = dzmin + min(q, dzspan))
  for (int ph = 0; ph < d->nphase; ++ph) {
    int pos = 0;
    for (int dz = dzmin; dz <= dzmax; ++dz)
      for (int t = 0; t < d->ntaps; ++t)
        if (d->tap_off[ph * d->ntaps + t][0] == dz) P.tap_order[ph * d->ntaps + pos++] = (uint8_t)t;
  }
  {  // the plane ring pays off where a super-tile has little tensor work per loaded plane (few taps, several channel chunks);
     // layers with a long tap loop hide their plane loads anyway and measured ~10 % slower with it on B200
    const int mmas = d->nphase * d->ntaps * P.nkc * (KC / 16) * td;
    P.sliding = mmas <= 256 ? 1 : 0;
  }
  if (d->nphase == 1 && P.sliding) {
    for (int q = 0; q < td; ++q) {
      const int last_dz = dzmin + (q < P.dzspan ? q : P.dzspan);
      int pos = -1;
      for (int t = 0; t < d->ntaps; ++t)
        if (d->tap_off[P.tap_order[t]][0] == last_dz) pos = t;
      if (pos < 0) pos = d->ntaps - 1;
      P.rel_after[pos] |= (uint8_t)(1u << q);
    }
  } else {
    P.rel_after[d->ntaps - 1] = (uint8_t)((1u << td) - 1u);
  }
  const size_t fixed = (size_t)P.np * P.nkc * P.plane_stride + bar_bytes + 1024;
  int nb = (int)((smem_cap - fixed) / P.b_stride);
  nb = nb > H_MAX_BSTAGES ? H_MAX_BSTAGES : nb;
  const int nbt = d->nphase * d->ntaps * P.nkc;       // weight tiles of the layer
  if (nb >= nbt) { nb = nbt; P.b_resident = 1; }
  else if (nb > 8) nb = 8;
  P.nb = nb;
  P.ng = 1;
  // The partition of the taps decides the order of the fp32 partial sums, so it must not depend on TD (which depends on the
  // batch size through the wave count): exactly two groups, and only for layers whose accumulators fit at every TD <= 4.
  // (A pair then gives bit-identical results alone and inside any batch — tests/test_gpu_parity.py, batch invariance.)
  if (P.b_resident && d->nphase == 1 && d->ntaps % 2 == 0 && 2 * 4 * d->Cout_w <= 256 && 2 * td <= H_MMA_WARPS) {
    P.ng = 2;
  }
  P.nbuf = (2 * P.ng * td * d->Cout_w <= 512) ? 2 : 1;
  P.acc_stride = 256;
  const int64_t total = (int64_t)P.tiles_w * P.tiles_h * P.tiles_d * d->N;
  OFSV_REQUIRE(total < (1ll << 31), "ofsv_conv_halo(ring): too many super-tiles");
  const size_t smem = fixed + (size_t)nb * P.b_stride;

  const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmA, tmB;
  {
    const cuuint64_t gdim[5] = {(cuuint64_t)d->Cin_s, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->Di, (cuuint64_t)d->N};
    const cuuint64_t es = 2;
    const cuuint64_t gstr[4] = {d->Cin_s * es, (cuuint64_t)d->Wi * d->Cin_s * es, (cuuint64_t)d->Hi * d->Wi * d->Cin_s * es,
                                (cuuint64_t)d->Di * d->Hi * d->Wi * d->Cin_s * es};
    const cuuint32_t box[5] = {(cuuint32_t)KC, HP_W, HP_H, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("ofsv_conv_halo(ring): cuTensorMapEncodeTiled(A) failed with %d", (int)r); return OFSV_ECUDA; }
  }
  {
    const cuuint64_t rows = (cuuint64_t)d->nphase * d->ntaps * P.nkc * d->Cout_w;
    const cuuint64_t gdim[2] = {(cuuint64_t)KC, rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)KC * 2};
    const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)d->Cout_w};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("ofsv_conv_halo(ring): cuTensorMapEncodeTiled(B) failed with %d", (int)r); return OFSV_ECUDA; }
  }
  const int grid = (int)(total < sms ? total : sms);
  cudaStream_t st = (cudaStream_t)stream;
  if (KC == 64) return launch_halo_ring<64>(P, tmA, tmB, bias, prelu, residual, y, grid, smem, st);
  if (KC == 32) return launch_halo_ring<32>(P, tmA, tmB, bias, prelu, residual, y, grid, smem, st);
  return launch_halo_ring<16>(P, tmA, tmB, bias, prelu, residual, y, grid, smem, st);
}
