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
// L2->SM traffic per 128 outputs drops from ~650 KB to ~(23 KB x (TD+2)/TD + 216 KB/TD).
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

namespace ofsv {

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

constexpr int HT_H = 16, HT_W = 8, HP_H = HT_H + 2, HP_W = HT_W + 2, HP_ROWS = HP_H * HP_W;  // 180 halo rows
constexpr int H_MAX_PLANES = 8;
constexpr int H_MAX_BSTAGES = 8;
constexpr int H_MMA_WARPS = 4, H_EPI_WARPS = 8, H_EPI_W0 = 1 + H_MMA_WARPS;
constexpr int H_THREADS = 32 * (H_EPI_W0 + H_EPI_WARPS);   // warp 0 TMA, warps 1..2 MMA issuers (output slices split by parity), warps 3..10 epilogue

struct HaloParams {
  int N, Do, Ho, Wo, Dy, Hy, Wy, Cout_s, Cout_w;
  int out_stride, nphase, ntaps, nkc;
  int td, np, dzmin;                 // output slices per super-tile, input planes per super-tile, min dz over taps
  int tiles_w, tiles_h, tiles_d;     // super-tile grid per sample
  int nb, plane_stride, b_stride;    // B ring depth, smem strides (bytes, multiples of 1024)
  int nbuf, acc_stride;              // TMEM accumulator double buffering
  int has_prelu, has_residual, out_f32, shuffle, dbg_flags;
  int nd, out_s2d;
  int8_t tap_off[OFSV_MAX_TAPS][4];
};

template <int KC>
__global__ void __launch_bounds__(H_THREADS, 1)
    conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p,
                     const float* __restrict__ bias, const float* __restrict__ prelu, const void* __restrict__ residual,
                     void* __restrict__ y, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int ROWB = KC * 2;
  const bool trace = dbg != nullptr && blockIdx.x == 0;   // OFSV_HALO_TRACE: per-super-tile clock64 timeline of CTA 0
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nplanes = p.np * p.nkc;
  uint8_t* sP = smem;
  uint8_t* sB = smem + nplanes * p.plane_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + p.nb * p.b_stride);
  uint64_t* plane_full = bars;                            // [H_MAX_PLANES]
  uint64_t* planes_empty = bars + H_MAX_PLANES;           // [1]
  uint64_t* b_full = planes_empty + 1;                    // [H_MAX_BSTAGES]
  uint64_t* b_empty = b_full + H_MAX_BSTAGES;             // [H_MAX_BSTAGES]
  uint64_t* acc_full = b_empty + H_MAX_BSTAGES;           // [2]
  uint64_t* acc_empty = acc_full + 2;                     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sBias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));   // [128], 16 B aligned
  float* sPrelu = sBias + 128;                              // [128]
  uint32_t* sTap = reinterpret_cast<uint32_t*>(sPrelu + 128); // [OFSV_MAX_TAPS]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_sample = p.tiles_w * p.tiles_h * p.tiles_d;
  const int total = per_sample * p.N;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < H_MAX_PLANES; ++i) mbar_init(&plane_full[i], 1);
    mbar_init(planes_empty, H_MMA_WARPS);
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
      int iter = 0;
      for (int st = blockIdx.x; st < total; st += gridDim.x, ++iter) {
        int r = st;
        const int tx = r % p.tiles_w; r /= p.tiles_w;
        const int ty = r % p.tiles_h; r /= p.tiles_h;
        const int tz = r % p.tiles_d;
        const int n = r / p.tiles_d;
        if (trace && iter < 16) dbg[iter * 16 + 0] = clock64();
        if (iter > 0) mbar_wait(planes_empty, (iter - 1) & 1, 200);     // previous super-tile's MMAs have read the planes
        if (trace && iter < 16) dbg[iter * 16 + 1] = clock64();
        for (int pl = 0; pl < p.np; ++pl)
          for (int kc = 0; kc < p.nkc; ++kc) {
            uint64_t* bar = &plane_full[pl * p.nkc + kc];
            mbar_expect_tx(bar, HP_ROWS * ROWB);
            tma_load_5d(&tmA, bar, sP + (pl * p.nkc + kc) * p.plane_stride, kc * KC, tx * HT_W - 1, ty * HT_H - 1,
                        tz * p.td + p.dzmin + pl, n);
          }
        for (int pass = 0; pass < p.nphase; ++pass)
          for (int t = 0; t < p.ntaps; ++t)
            for (int kc = 0; kc < p.nkc; ++kc, ++bcount) {
              if (bcount >= (uint32_t)p.nb) mbar_wait(&b_empty[bs], bphase ^ 1u, 100);
              mbar_expect_tx(&b_full[bs], p.Cout_w * ROWB);
              tma_load_2d(&tmB, &b_full[bs], sB + bs * p.b_stride, 0, ((pass * p.ntaps + t) * p.nkc + kc) * p.Cout_w);
              if (++bs == (uint32_t)p.nb) { bs = 0; bphase ^= 1u; }
            }
      }
    }
  } else if (warp < H_EPI_W0) {
    const int iss = warp - 1;   // issuer index: owns output slices j with j % H_MMA_WARPS == iss
    // ================= MMA issuer: the whole warp walks the (warp-uniform) loops, one elected lane issues =================
    // The issuing thread is latency-bound on its own scalar code, so everything per operand tile is table-driven:
    // sTap[pass*ntaps + t] = (descriptor offset of the tap inside a plane) | (dz - dzmin) << 16, ring indices are counters.
    for (int i = lane; iss == 0 && i < p.nphase * p.ntaps; i += 32) {
      const int8_t* off = p.tap_off[i];
      sTap[i] = ((uint32_t)(((off[1] + 1) * HP_W + (off[2] + 1)) * ROWB) >> 4) | ((uint32_t)(off[0] - p.dzmin) << 16);
    }
    asm volatile("bar.sync 2, %0;" ::"n"(32 * H_MMA_WARPS) : "memory");
    const uint32_t leader = elect_one_sync();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Cout_w >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_hi = kmajor_desc_hi<KC>(HP_W * ROWB);      // SBO = 10 halo rows
    const uint32_t b_hi = kmajor_desc_hi<KC>(8 * ROWB);
    const uint32_t plane_lo0 = kmajor_desc_lo(smem_u32(sP)), plane_step = (uint32_t)p.plane_stride >> 4;
    const uint32_t b_lo0 = kmajor_desc_lo(smem_u32(sB)), b_step = (uint32_t)p.b_stride >> 4;
    const uint32_t jstep = (uint32_t)p.nkc * plane_step;
    const int ntap_total = p.ntaps;
    uint32_t bs = 0, bphase = 0;             // B ring slot / parity
    uint32_t buf = 0, acc_use = 0;           // accumulator buffer, number of uses so far
    int iter = 0;
    for (int st = blockIdx.x; st < total; st += gridDim.x, ++iter) {
      const int tz = (st / (p.tiles_w * p.tiles_h)) % p.tiles_d;
      const int nj = min(p.td, p.Do - tz * p.td);
      const uint32_t jmask = (1u << nj) - 1u;
      uint32_t waited = 0;                   // bit (slice plane * nkc + kc)
      const long long t_in = trace ? clock64() : 0;
      long long w_plane = 0, w_b = 0, w_acc = 0, tw = 0;
      for (int pass = 0; pass < p.nphase; ++pass) {
        if (trace) tw = clock64();
        if (acc_use >= (uint32_t)p.nbuf) mbar_wait(&acc_empty[buf], ((acc_use / p.nbuf) - 1) & 1);
        if (trace) w_acc += clock64() - tw;
        const uint32_t acc0 = tmem_base + buf * p.acc_stride;
        const uint32_t* taps = sTap + pass * ntap_total;
        for (int t = 0; t < ntap_total; ++t) {
          const uint32_t te = taps[t];
          const uint32_t tap16 = te & 0xFFFFu, dzr = te >> 16;
          for (int kc = 0; kc < p.nkc; ++kc) {
            // planes this (tap, chunk) touches for the first time in this super-tile (nkc == 1: planes dzr .. dzr+nj-1)
            if (trace) tw = clock64();
            if (p.nkc == 1) {
              uint32_t need = (jmask << dzr) & ~waited;
              while (need) { const int pl = __ffs(need) - 1; mbar_wait(&plane_full[pl], iter & 1); need &= need - 1; }
              waited |= jmask << dzr;
            } else {
              for (int j = 0; j < nj; ++j) {
                const int pl = (j + dzr) * p.nkc + kc;
                if (!((waited >> pl) & 1u)) { mbar_wait(&plane_full[pl], iter & 1); waited |= 1u << pl; }
              }
            }
            if (trace) { const long long t1 = clock64(); w_plane += t1 - tw; tw = t1; }
            mbar_wait(&b_full[bs], bphase);
            if (trace) w_b += clock64() - tw;
            tcgen05_fence_after();
            if (leader) {
              const uint32_t b_lo = b_lo0 + bs * b_step;
              const uint32_t a_lo = plane_lo0 + (dzr * p.nkc + kc) * plane_step + tap16;
              const uint32_t first = (t | kc) ? 1u : 0u;
              for (int j = iss; j < nj; j += H_MMA_WARPS) {
                const uint32_t aj = a_lo + j * jstep, dj = acc0 + j * p.Cout_w;
                umma_bf16_lohi(dj, aj, a_hi, b_lo, b_hi, idesc, first);
#pragma unroll
                for (int k = 1; k < KC / 16; ++k) umma_bf16_lohi(dj, aj + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, 1u);
              }
              tcgen05_commit(&b_empty[bs]);
            }
            __syncwarp();
            if (++bs == (uint32_t)p.nb) { bs = 0; bphase ^= 1u; }
          }
        }
        if (leader) tcgen05_commit(&acc_full[buf]);
        __syncwarp();
        ++acc_use;
        if (++buf == (uint32_t)p.nbuf) buf = 0;
      }
      if (leader) tcgen05_commit(planes_empty);
      __syncwarp();
      if (trace && iter < 16 && lane == 0 && iss == 0) {
        dbg[iter * 16 + 2] = t_in; dbg[iter * 16 + 3] = clock64();
        dbg[iter * 16 + 4] = w_plane; dbg[iter * 16 + 5] = w_b; dbg[iter * 16 + 6] = w_acc;
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
    for (int st = blockIdx.x; st < total; st += gridDim.x) {
      int r = st;
      const int tx = r % p.tiles_w; r /= p.tiles_w;
      const int ty = r % p.tiles_h; r /= p.tiles_h;
      const int tz = r % p.tiles_d;
      const int n = r / p.tiles_d;
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
        for (int it = half; it < ((p.dbg_flags & 1) ? 0 : nitems); it += 2) {
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
          if (!valid_xy || (p.dbg_flags & 2)) continue;
          if (!(p.dbg_flags & 4)) {
            // bias / PReLU slopes as 16 B shared loads (8 instead of 32 LDS per chunk: the epilogue's shared-memory traffic
            // competes with the tensor core's operand fetch)
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + e4 * 4);
              const float4 p4 = *reinterpret_cast<const float4*>(sPrelu + c0 + e4 * 4);
              const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = v[e4 * 4 + e] + bb[e];
                v[e4 * 4 + e] = fmaxf(a, 0.0f) + pp[e] * fminf(a, 0.0f);
              }
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
                w.x = pack2_bf16(v[8 * e], v[8 * e + 1]); w.y = pack2_bf16(v[8 * e + 2], v[8 * e + 3]);
                w.z = pack2_bf16(v[8 * e + 4], v[8 * e + 5]); w.w = pack2_bf16(v[8 * e + 6], v[8 * e + 7]);
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
                w.x = pack2_bf16(v[e], v[e + 1]); w.y = pack2_bf16(v[e + 2], v[e + 3]);
                w.z = pack2_bf16(v[e + 4], v[e + 5]); w.w = pack2_bf16(v[e + 6], v[e + 7]);
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
static int launch_halo(const HaloParams& P, const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias,
                       const float* prelu, const void* residual, void* y, int grid, size_t smem, cudaStream_t st) {
  long long* dbg = nullptr;
  const char* trace_path = getenv("OFSV_HALO_TRACE");
  if (trace_path) { cudaMalloc(&dbg, 16 * 16 * 8); cudaMemset(dbg, 0, 16 * 16 * 8); }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("ofsv_conv_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return OFSV_ECUDA; }
    attr_done = true;
  }
  conv_halo_kernel<KC><<<grid, H_THREADS, smem, st>>>(tmA, tmB, P, bias, prelu, residual, y, dbg);
  const int rc = check_launch("conv_halo_kernel");
  if (trace_path) {   // debug only: synchronous dump of CTA 0's timeline
    static long long h[256];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(dbg);
    if (FILE* f = fopen(trace_path, "a")) {
      fprintf(f, "launch N=%d Cout_w=%d td=%d nphase=%d ntaps=%d tiles=%dx%dx%d grid=%d nb=%d nbuf=%d\n", P.N, P.Cout_w, P.td, P.nphase, P.ntaps, P.tiles_w, P.tiles_h, P.tiles_d, grid, P.nb, P.nbuf);
      const long long t0 = h[0];
      for (int i = 0; i < 16 && h[i * 16 + 3]; ++i)
        fprintf(f, "  st %2d: prod wait_planes_empty [%lld..%lld]  mma [%lld..%lld] wait plane %lld b %lld acc %lld | epi(pass %d) wait_full [%lld..%lld] done %lld\n", i, h[i * 16] - t0, h[i * 16 + 1] - t0, h[i * 16 + 2] - t0, h[i * 16 + 3] - t0, h[i * 16 + 4], h[i * 16 + 5], h[i * 16 + 6], i, h[i * 16 + 8] - t0, h[i * 16 + 9] - t0, h[i * 16 + 10] - t0);
      fclose(f);
    }
  }
  return rc;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace ofsv

using namespace ofsv;

// Same contract as ofsv_conv_tc; returns OFSV_ENOSUP (nothing launched) for layers outside this kernel's domain
// (strided input, tap offsets outside {-1,0,1}, Cout_w > 128, planes that do not fit shared memory).
extern "C" int ofsv_conv_halo(const ofsv_conv_desc* d, const void* x, const void* w, const float* bias, const float* prelu,
                              const void* residual, void* y, void* stream) {
  if (int e = validate_conv_desc(d, "ofsv_conv_halo")) return e;
  if (d->N == 0) return OFSV_OK;
  OFSV_REQUIRE(x && w && bias && y, "ofsv_conv_halo: null pointer");
  OFSV_REQUIRE(!d->has_prelu || prelu, "ofsv_conv_halo: has_prelu without prelu slopes");
  OFSV_REQUIRE(!d->has_residual || residual, "ofsv_conv_halo: has_residual without residual");
  OFSV_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && (!residual || aligned16(residual)),
               "ofsv_conv_halo: pointers must be 16-byte aligned");
  if (d->in_dtype != OFSV_BF16) { set_error("ofsv_conv_halo: activations must be bf16"); return OFSV_ENOSUP; }
  if (d->in_stride != 1) { set_error("ofsv_conv_halo: in_stride must be 1"); return OFSV_ENOSUP; }
  if (d->Cout_w > 128) { set_error("ofsv_conv_halo: Cout_w=%d > 128", d->Cout_w); return OFSV_ENOSUP; }
  if (d->has_residual && d->out_dtype != OFSV_BF16 && !d->out_shuffle) { set_error("ofsv_conv_halo: residual needs a bf16 output"); return OFSV_ENOSUP; }
  if (d->has_residual && d->out_shuffle && d->out_dtype != OFSV_F32) { set_error("ofsv_conv_halo: depth-to-space residual (flow/mask state) is fp32"); return OFSV_ENOSUP; }
  if (d->out_shuffle && (d->out_shuffle != 8 || d->nphase != 1 || d->Cout_w != 8 * (1 << d->nd) || d->Cout_s != 8)) {
    set_error("ofsv_conv_halo: bad depth-to-space configuration");
    return OFSV_EINVAL;
  }
  if (d->out_s2d && (d->nphase != 1 || d->out_stride != 1 || d->has_residual || d->out_shuffle || d->Hy % 2 || d->Wy % 2 ||
                     (d->nd == 3 && d->Dy % 2))) {
    set_error("ofsv_conv_halo: bad space-to-depth output configuration");
    return OFSV_EINVAL;
  }
  int dzmin = 1, dzmax = -1;
  for (int i = 0; i < d->nphase * d->ntaps; ++i) {
    const int8_t* o = d->tap_off[i];
    if (o[0] < -1 || o[0] > 1 || o[1] < -1 || o[1] > 1 || o[2] < -1 || o[2] > 1) {
      set_error("ofsv_conv_halo: tap offset outside {-1,0,1}");
      return OFSV_ENOSUP;
    }
    dzmin = o[0] < dzmin ? o[0] : dzmin;
    dzmax = o[0] > dzmax ? o[0] : dzmax;
  }
  {  // layers with little tensor work per loaded plane (the 2^d-tap space-to-depth conv0 layers): plane-ring kernel
    const char* f = getenv("OFSV_HALO_RING");
    const bool ring = f ? atoi(f) != 0 : (d->nphase == 1 && d->ntaps <= 8);
    if (ring) return conv_halo_ring(d, x, w, bias, prelu, residual, y, stream);
  }
  PFN_encodeTiled encode = get_tensor_map_encoder();
  if (!encode) { set_error("ofsv_conv_halo: cuTensorMapEncodeTiled unavailable"); return OFSV_ECUDA; }

  const int KC = d->Cin_s % 64 == 0 ? 64 : (d->Cin_s % 32 == 0 ? 32 : 16);
  HaloParams P;
  memset(&P, 0, sizeof(P));
  P.N = d->N; P.Do = d->Do; P.Ho = d->Ho; P.Wo = d->Wo; P.Dy = d->Dy; P.Hy = d->Hy; P.Wy = d->Wy;
  P.Cout_s = d->Cout_s; P.Cout_w = d->Cout_w; P.out_stride = d->out_stride; P.nphase = d->nphase; P.ntaps = d->ntaps;
  P.nkc = d->Cin_s / KC; P.dzmin = dzmin;
  P.plane_stride = (HP_ROWS * KC * 2 + 1023) & ~1023;
  P.b_stride = (d->Cout_w * KC * 2 + 1023) & ~1023;
  P.tiles_w = (int)cdiv(d->Wo, HT_W); P.tiles_h = (int)cdiv(d->Ho, HT_H);
  P.has_prelu = d->has_prelu; P.has_residual = d->has_residual; P.out_f32 = d->out_dtype == OFSV_F32;
  P.shuffle = d->out_shuffle; P.nd = d->nd; P.out_s2d = d->out_s2d;
  { const char* f = getenv("OFSV_HALO_DBGFLAGS"); P.dbg_flags = f ? atoi(f) : 0; }   // bring-up switches (1 = skip epilogue work)
  memcpy(P.tap_off, d->tap_off, sizeof(P.tap_off));
  const int sms = num_sms();
  const size_t smem_cap = 227 * 1024 - 2048;
  const size_t bar_bytes = (H_MAX_PLANES + 1 + 2 * H_MAX_BSTAGES + 4) * 8 + 32 + 2 * 128 * 4 + OFSV_MAX_TAPS * 4;
  // TD (output slices per super-tile, all sharing one pass over the weight tiles): pick the feasible candidate with the
  // smallest modelled time = rounds of super-tiles per SM x (max(tensor time, weight-stream time) + exposed plane loads).
  // Constants measured on B200: an SS-mode M128.N.K16 MMA takes 44 + 0.49 N cycles (tests/umma_rate_probe.cu); weight tiles
  // arrive from L2 at ~10 B/clk per SM when every CTA streams the same tensor, halo planes at ~30 B/clk (tests/halo_trace.py).
  int td = 0;
  const char* force = getenv("OFSV_HALO_TD");   // test hook: force the super-tile depth (1, 2 or 4) when it fits
  const int forced = force ? atoi(force) : 0;
  double best = 0.0;
  for (int cand = 4; cand >= 1; cand >>= 1) {
    if (forced && cand != forced && cand > 1) continue;
    if (cand > d->Do && cand > 1) continue;
    const int np = cand + (dzmax - dzmin);
    if (np * P.nkc > H_MAX_PLANES) continue;
    if (cand * d->Cout_w > 512) continue;
    if ((size_t)np * P.nkc * P.plane_stride + 3 * (size_t)P.b_stride + bar_bytes + 1024 > smem_cap) continue;
    const int64_t nst = (int64_t)P.tiles_w * P.tiles_h * cdiv(d->Do, cand) * d->N;
    const double ntiles_b = (double)d->nphase * d->ntaps * P.nkc;
    const double mma = ntiles_b * (KC / 16) * (44.0 + 0.49 * d->Cout_w) * cand;
    const double wstream = ntiles_b * P.b_stride / 10.0;
    const double planes = (double)np * P.nkc * P.plane_stride / 30.0;
    const double cost = (double)cdiv(nst, sms) * ((mma > wstream ? mma : wstream) + planes);
    if (td == 0 || cost < best) { td = cand; best = cost; }
    if (forced && cand == forced) break;
  }
  if (td == 0) { set_error("ofsv_conv_halo: layer does not fit (Cin_s=%d Cout_w=%d)", d->Cin_s, d->Cout_w); return OFSV_ENOSUP; }
  P.td = td; P.np = td + (dzmax - dzmin);
  P.tiles_d = (int)cdiv(d->Do, td);
  const size_t fixed = (size_t)P.np * P.nkc * P.plane_stride + bar_bytes + 1024;
  int nb = (int)((smem_cap - fixed) / P.b_stride);
  nb = nb > H_MAX_BSTAGES ? H_MAX_BSTAGES : nb;
  P.nb = nb;
  P.nbuf = (2 * td * d->Cout_w <= 512) ? 2 : 1;
  P.acc_stride = 256;
  const int64_t total = (int64_t)P.tiles_w * P.tiles_h * P.tiles_d * d->N;
  OFSV_REQUIRE(total < (1ll << 31), "ofsv_conv_halo: too many super-tiles");
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
    if (r != CUDA_SUCCESS) { set_error("ofsv_conv_halo: cuTensorMapEncodeTiled(A) failed with %d", (int)r); return OFSV_ECUDA; }
  }
  {
    const cuuint64_t rows = (cuuint64_t)d->nphase * d->ntaps * P.nkc * d->Cout_w;
    const cuuint64_t gdim[2] = {(cuuint64_t)KC, rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)KC * 2};
    const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)d->Cout_w};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("ofsv_conv_halo: cuTensorMapEncodeTiled(B) failed with %d", (int)r); return OFSV_ECUDA; }
  }
  const int grid = (int)(total < sms ? total : sms);
  cudaStream_t st = (cudaStream_t)stream;
  if (KC == 64) return launch_halo<64>(P, tmA, tmB, bias, prelu, residual, y, grid, smem, st);
  if (KC == 32) return launch_halo<32>(P, tmA, tmB, bias, prelu, residual, y, grid, smem, st);
  return launch_halo<16>(P, tmA, tmB, bias, prelu, residual, y, grid, smem, st);
}
