// STACKED-N halo-reuse tcgen05 convolution: the stride-1 layers of an IFBlock (3^d convs, the 2^d-phase form of
// ConvTranspose(4,2,1), the depth-to-space heads) — Flow-2D/model/IFNet.py:16-27,95-116, Flow-3D/model/IFNet.py:15-27,92-119.
//
// What bounds an SS-mode tcgen05.mma M128.N.K16 on B200 is max(N/2, (4 KB + 32 N B) / 128 B/clk) cycles
// (tests/umma_rate_probe.cu): below N = 128 the MMA waits for the shared-memory port, and 95 % of this network's FLOPs sit
// in layers with Cout = 64 (67 % of the tensor peak at best).  This kernel makes N large instead of shrinking the operand
// traffic: a shifted window of ONE input plane is the A operand of every output depth slice that reads it, so
//
//   super-tile  = 16(h) x 8(w) x TD(d) output positions; the TD + dz-span input halo planes (18 x 10 rows of KC channels,
//                 one 5-D TMA box each, OOB zero fill = conv padding) are loaded once per super-tile
//   group       = one in-plane tap offset (dy,dx) [of one pass]; its B stage holds the weight "slots" of every (phase, dz)
//                 that uses that offset, stacked along N in (phase asc, dz desc) order
//   op          = ONE MMA per (plane, run of slots): A = the plane's window shifted by (dy,dx), B = the run of slots,
//                 D = the TMEM columns of the output slices those slots feed — consecutive slices sit in consecutive
//                 column blocks, so a plane in the middle of a 3^3 conv feeds three slices with one N = 192 MMA
//                 (96 cycles of tensor work, 80 of operand fetch) instead of three N = 64 MMAs (3 x 48, port-bound).
//                 The op list is built on the host (build_ops) and checked exhaustively by ofsv_conv_stack_selfcheck.
//   pass        = the phases evaluated together (ConvTranspose: the x-parity pair of one (z,y) parity)
//   kc-outer    = channel chunks are the OUTER loop of a pass, so the planes of chunk kc are dead after their half of the
//                 super-tile and the next super-tile's planes of that chunk stream in under the other half's MMAs
//                 (a dedicated plane-producer warp; the legacy kernel exposed ~6 k cycles of plane loads per super-tile)
//   epilogue    = tcgen05.ld -> bias / PReLU / residual -> bf16 rows (or fp32 depth-to-space heads with the flow/mask state
//                 accumulation) staged per warp in swizzled shared memory and written by TMA tensor stores (one 2-8 KB box
//                 per warp and block) instead of 32-line scattered 16 B stores.
// The per-slice accumulation order (group, chunk, plane = dz ascending, k) does not depend on TD, so results are
// bit-identical across batch sizes (tests/test_gpu_parity.py batch invariance).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "tc_common.cuh"

namespace ofsv {

constexpr int SK_HT_H = 16, SK_HT_W = 8, SK_HP_H = SK_HT_H + 2, SK_HP_W = SK_HT_W + 2, SK_HP_ROWS = SK_HP_H * SK_HP_W;
constexpr int SK_MAX_GROUPS = 32, SK_MAX_OPS = 224, SK_MAX_SLOTS = 6, SK_MAX_PASS = 8;
constexpr int SK_MAX_PBARS = 24, SK_MAX_KC = 8, SK_MAX_RING = 8, SK_MAX_BST = 8;
constexpr int SK_EPI_W0 = 4, SK_EPI_WARPS = 8;
constexpr int SK_THREADS = 32 * (SK_EPI_W0 + SK_EPI_WARPS);   // warp 0 weights, 1 MMA issuer, 2 planes, 3 idle, 4..11 epilogue

enum { SK_EPI_DIRECT = 0, SK_EPI_TMA_ROWS = 1, SK_EPI_TMA_SHUF = 2 };

// ---------------------------------------------------------------------------------------------------- host-side plan
struct SkSlot {
  int8_t p, oz; uint8_t tap;                           // phase index inside the pass, dz of the tap, ph * ntaps + t
  int16_t c_lo, c_hi;                                  // columns [c_lo, c_hi) of the slot's Cout_w can be non-zero (structural zeros outside)
};
struct SkGroup {
  int pass, oy, ox, nslots, row0;                      // row0: first row of the group's kc = 0 block in the packed weights
  SkSlot slots[SK_MAX_SLOTS];
};
struct SkPlan {
  int KC, nkc, P, npass, ngroups, dzmin, dzmax;
  int group_first[SK_MAX_PASS + 1];
  SkGroup g[SK_MAX_GROUPS];
};

// 32-channel chunks (64 B rows, SWIZZLE_64B): Cin = 64 gives two chunks whose planes double-buffer under the kc-outer loop,
// and a weight stage of (up to 4 slots x 128 rows) stays <= 32 KB so that the ring keeps >= 2 stages next to the planes.
static int sk_kc(int Cin_s) { return Cin_s % 32 == 0 ? 32 : 16; }

// Pure function of (nd, nphase, ntaps, tap_off, Cin_s, Cout_w): groups, slot order and the packed weight layout.
// Returns false when the layer is outside the kernel's domain.
static bool sk_make_plan(const ofsv_conv_desc* d, SkPlan* pl) {
  memset(pl, 0, sizeof(*pl));
  pl->KC = sk_kc(d->Cin_s);
  pl->nkc = d->Cin_s / pl->KC;
  if (pl->nkc > SK_MAX_KC) return false;
  pl->P = (d->nphase > 1 && 2 * d->Cout_w <= 256) ? 2 : 1;
  pl->npass = d->nphase / pl->P;
  if (pl->npass > SK_MAX_PASS) return false;
  pl->dzmin = 1; pl->dzmax = -1;
  for (int i = 0; i < d->nphase * d->ntaps; ++i) {
    const int8_t* o = d->tap_off[i];
    if (o[0] < -1 || o[0] > 1 || o[1] < -1 || o[1] > 1 || o[2] < -1 || o[2] > 1) return false;
    pl->dzmin = o[0] < pl->dzmin ? o[0] : pl->dzmin;
    pl->dzmax = o[0] > pl->dzmax ? o[0] : pl->dzmax;
  }
  int row = 0;
  for (int pass = 0; pass < pl->npass; ++pass) {
    pl->group_first[pass] = pl->ngroups;
    for (int oy = -1; oy <= 1; ++oy)
      for (int ox = -1; ox <= 1; ++ox) {
        SkGroup G;
        memset(&G, 0, sizeof(G));
        G.pass = pass; G.oy = oy; G.ox = ox;
        for (int p = 0; p < pl->P; ++p)                 // slot order: phase ascending, dz DESCENDING (= output slice ascending)
          for (int oz = 1; oz >= -1; --oz) {
            const int ph = pass * pl->P + p;
            for (int t = 0; t < d->ntaps; ++t) {
              const int8_t* o = d->tap_off[ph * d->ntaps + t];
              if (o[0] == oz && o[1] == oy && o[2] == ox) {
                if (G.nslots == SK_MAX_SLOTS) return false;
                G.slots[G.nslots].p = (int8_t)p; G.slots[G.nslots].oz = (int8_t)oz; G.slots[G.nslots].tap = (uint8_t)(ph * d->ntaps + t);
                // depth-to-space heads (out_shuffle: columns [parity (z,y,x)][8 ch], include/ofsv.h): a tap with dz = +1 only feeds the
                // z-parity-1 half of the columns, dz = -1 only the z-parity-0 half — the rest of its weights are zero BY CONTRACT
                G.slots[G.nslots].c_lo = (int16_t)((d->out_shuffle && d->nd == 3 && oz == 1) ? d->Cout_w / 2 : 0);
                G.slots[G.nslots].c_hi = (int16_t)((d->out_shuffle && d->nd == 3 && oz == -1) ? d->Cout_w / 2 : d->Cout_w);
                ++G.nslots;
              }
            }
          }
        if (G.nslots == 0) continue;
        if (pl->ngroups == SK_MAX_GROUPS) return false;
        G.row0 = row;
        row += pl->nkc * G.nslots * d->Cout_w;
        pl->g[pl->ngroups++] = G;
      }
  }
  pl->group_first[pl->npass] = pl->ngroups;
  return true;
}

struct SkOp { uint8_t group, q, slot0, nsl, col0, fresh, issuer; int8_t oy, ox; uint8_t lo, hi; };   // lo / hi: zero columns skipped at the ends of the run
static int sk_op_cols(const SkOp& o, int Cout_w) { return o.nsl * Cout_w - o.lo - o.hi; }

constexpr int SK_NI = 2;      // MMA-issuing warps

// The single thread that issues tcgen05.mma spends ~50 cycles of uniform-datapath work per op and the tensor queue it feeds is
// shallow: with one issuer the convblocks lose a quarter of the tensor time to it and the small-N layers (N = 32 / 64 conv0)
// run at 83 cycles per MMA against 42-48 of operand fetch.  So the column blocks of a pass are split in two halves, each owned
// by ONE issuing warp (every accumulator still receives its MMAs from one thread, in the same order: results stay deterministic
// and independent of the super-tile depth); runs never cross the boundary between the halves.
static int sk_issuers(const SkPlan& pl, int td) { return pl.P * td >= 2 ? SK_NI : 1; }
static int sk_block_issuer(const SkPlan& pl, int td, int col) { const int nblk = pl.P * td; return sk_issuers(pl, td) * col / nblk; }

// MMA list for super-tile depth td: per group and issuer, per plane, greedy runs of (consecutive slots, consecutive column
// blocks of one issuer, same initialisation state).  Column block of (phase p, slice j) = p * td + j; slice j of plane q for a
// slot with dz = oz is j = q + dzmin - oz.  op_first[g * SK_NI + w] = first op of (group g, issuer w), ops of a group are
// stored issuer-major.  Returns the number of ops or -1 when the table would overflow.
static int sk_build_ops(const ofsv_conv_desc* d, const SkPlan& pl, int td, SkOp* ops, int* op_first /* [ngroups * SK_NI + 1] */) {
  const int np = td + pl.dzmax - pl.dzmin;
  const int ni = sk_issuers(pl, td);
  int n = 0;
  for (int pass = 0; pass < pl.npass; ++pass) {
    bool init[64] = {false};
    for (int gi = pl.group_first[pass]; gi < pl.group_first[pass + 1]; ++gi) {
      const SkGroup& G = pl.g[gi];
      for (int w = 0; w < SK_NI; ++w) {
        op_first[gi * SK_NI + w] = n;
        if (w >= ni) continue;
        for (int q = 0; q < np; ++q) {
          int s = 0;
          while (s < G.nslots) {
            const int j = q + pl.dzmin - G.slots[s].oz;
            const int col = G.slots[s].p * td + j;
            if (j < 0 || j >= td || sk_block_issuer(pl, td, col) != w) { ++s; continue; }
            const bool fresh = !init[col];
            int len = 1;
            while (s + len < G.nslots) {
              const int j2 = q + pl.dzmin - G.slots[s + len].oz;
              if (j2 < 0 || j2 >= td) break;
              const int col2 = G.slots[s + len].p * td + j2;
              if (col2 != col + len || sk_block_issuer(pl, td, col2) != w || (!init[col2]) != fresh || (len + 1) * d->Cout_w > 256) break;
              ++len;
            }
            if (n == SK_MAX_OPS) return -1;
            SkOp o;
            o.group = (uint8_t)gi; o.q = (uint8_t)q; o.slot0 = (uint8_t)s; o.nsl = (uint8_t)len; o.col0 = (uint8_t)col; o.fresh = fresh;
            o.issuer = (uint8_t)w; o.oy = (int8_t)G.oy; o.ox = (int8_t)G.ox;
            // structural zeros at the two ends of the run are not multiplied — unless this MMA initialises its accumulator columns
            o.lo = fresh ? 0 : (uint8_t)G.slots[s].c_lo;
            o.hi = fresh ? 0 : (uint8_t)(d->Cout_w - G.slots[s + len - 1].c_hi);
            ops[n++] = o;
            for (int i = 0; i < len; ++i) init[col + i] = true;
            s += len;
          }
        }
      }
    }
  }
  op_first[pl.ngroups * SK_NI] = n;
  return n;
}

// cycles of one K16 step of an SS-mode M128 MMA with N columns (tests/umma_rate_probe.cu)
static double sk_mma_cycles(int N) {
  double c = (4096.0 + 32.0 * N) / 128.0;            // operand fetch through the 128 B/clk shared-memory port
  if (c < 41.0) c = 41.0;                            // measured issue floor (N = 16: 41.0, N = 32: 41.6, N = 64: 48.1)
  return c > N / 2.0 ? c : N / 2.0;                  // tensor pipe from N = 128 (64.1) on
}

// ---------------------------------------------------------------------------------------------------- device side
struct SkGroupRec { uint32_t row0; uint8_t op_begin[SK_NI], nops[SK_NI]; uint8_t nslots, pad[3]; };
struct SkParams {
  int N, Do, Ho, Wo, Dy, Hy, Wy, Cout_s, Cout_w;
  int out_stride, nd, nkc, td, np, dzmin, P, npass, nring;
  int tiles_w, tiles_h, tiles_d;
  int nb, plane_stride, b_stride, stg_stride, stg_rowb;
  int nbuf, acc_stride;
  int has_prelu, has_residual, out_f32, shuffle, out_s2d, epi_mode, hfast, probe, b_resident, ni;
  uint16_t group_first[SK_MAX_PASS + 1];
  SkGroupRec groups[SK_MAX_GROUPS];
  // MMA list, READY TO USE: the single issuing thread runs a dependent chain of uniform-datapath instructions per MMA (one warp
  // cannot hide their latency: ~25 field extractions / multiplies per op cost ~200 cycles against ~130 of tensor work), so every
  // field is a whole 32-bit word — per op the issuer does three independent adds.
  //   [0] A descriptor offset ((plane q, chunk 0) + in-plane tap, >> 4)   [1] B descriptor offset inside the stage (>> 4)
  //   [2] TMEM column offset | fresh << 16                               [3] instruction descriptor (N of this run)
  alignas(16) uint32_t ops[SK_MAX_OPS][4];
};
struct SkOutMaps { CUtensorMap m[8]; };

__device__ __forceinline__ uint32_t sk_pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void sk_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int KC>
__global__ void __launch_bounds__(SK_THREADS, 1)
    conv_stack_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ SkOutMaps tmO, const __grid_constant__ SkParams p, const float* __restrict__ bias,
                      const float* __restrict__ prelu, const void* __restrict__ residual, void* __restrict__ y) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int ROWB = KC * 2;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nplanes = p.nring * p.np;
  uint8_t* sP = smem;                                                  // [ring slot][q] halo plane chunks
  uint8_t* sB = sP + (size_t)nplanes * p.plane_stride;                 // [nb] weight stages
  uint8_t* sS = sB + (size_t)p.nb * p.b_stride;                        // [8 warps] epilogue staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(sS + (size_t)SK_EPI_WARPS * p.stg_stride);
  uint64_t* plane_full = bars;                                         // [SK_MAX_PBARS]  index ring slot * np + q
  uint64_t* plane_empty = plane_full + SK_MAX_PBARS;                   // [SK_MAX_RING]   index ring slot
  uint64_t* b_full = plane_empty + SK_MAX_RING;                        // [SK_MAX_BST]
  uint64_t* b_empty = b_full + SK_MAX_BST;                             // [SK_MAX_BST]
  uint64_t* acc_full = b_empty + SK_MAX_BST;                           // [2]
  uint64_t* acc_empty = acc_full + 2;                                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sBias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));   // [128]
  float* sPrelu = sBias + 128;                                         // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_sample = p.tiles_w * p.tiles_h * p.tiles_d;
  const int total = per_sample * p.N;
  const int ngroups_all = p.group_first[p.npass];

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < SK_MAX_PBARS; ++i) mbar_init(&plane_full[i], 1);
    for (int i = 0; i < SK_MAX_RING; ++i) mbar_init(&plane_empty[i], p.ni);
    for (int i = 0; i < SK_MAX_BST; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], p.ni); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], p.ni); mbar_init(&acc_empty[i], SK_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.Cout_w; i += SK_THREADS) {
    sBias[i] = __ldg(bias + i);
    sPrelu[i] = p.has_prelu ? __ldg(prelu + i) : 1.0f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= weight producer: one B stage per (pass, chunk, group), slot by slot =================
    if (lane == 0) {
      uint32_t bcount = 0, bs = 0, bphase = 0;
      const uint32_t slot_bytes = (uint32_t)p.Cout_w * ROWB;
      if (p.b_resident) {
        // the whole packed weight tensor fits next to the planes: loaded ONCE per CTA in its global order, no ring, no hand-shakes
        // (a stage boundary costs the issuing thread ~370 cycles of barrier / commit latency — as much as the tensor work of a
        //  2-slot stage of the 32-channel conv0 layers)
        uint32_t nblk = 0;
        for (int g = 0; g < ngroups_all; ++g) nblk += (uint32_t)p.groups[g].nslots * (uint32_t)p.nkc;
        mbar_expect_tx(&b_full[0], nblk * slot_bytes);
        for (uint32_t b = 0; b < nblk; ++b) tma_load_2d(&tmB, &b_full[0], sB + (size_t)b * slot_bytes, 0, (int)(b * (uint32_t)p.Cout_w));
      } else
      for (int st = blockIdx.x; st < total; st += gridDim.x)
        for (int pass = 0; pass < p.npass; ++pass)
          for (int kc = 0; kc < p.nkc; ++kc)
            for (int g = p.group_first[pass]; g < p.group_first[pass + 1]; ++g, ++bcount) {
              const SkGroupRec gr = p.groups[g];
              if (bcount >= (uint32_t)p.nb) mbar_wait(&b_empty[bs], bphase ^ 1u, 0);
#ifdef OFSV_STACK_PROBE
              if ((p.probe & 1) && st != (int)blockIdx.x) { mbar_expect_tx(&b_full[bs], 0); if (++bs == (uint32_t)p.nb) { bs = 0; bphase ^= 1u; } continue; }
#endif
              mbar_expect_tx(&b_full[bs], gr.nslots * slot_bytes);
              const int row = (int)gr.row0 + kc * gr.nslots * p.Cout_w;
              for (int s = 0; s < gr.nslots; ++s)
                tma_load_2d(&tmB, &b_full[bs], sB + (size_t)bs * p.b_stride + (size_t)s * slot_bytes, 0, row + s * p.Cout_w);
              if (++bs == (uint32_t)p.nb) { bs = 0; bphase ^= 1u; }
            }
    }
  } else if (warp == 2) {
    // ================= plane producer: the halo planes of chunk c = (super-tile, kc) go to ring slot c % nring =================
    // (a slot is reused when the MMAs of the chunk nring places earlier have read it; layers with several passes keep every chunk
    //  of a super-tile resident: nring is then a multiple of nkc and the slot is released after the last pass)
    if (lane == 0) {
      int it = 0;
      for (int st = blockIdx.x; st < total; st += gridDim.x, ++it) {
        int r = st;
        const int tx = r % p.tiles_w; r /= p.tiles_w;
        const int ty = r % p.tiles_h; r /= p.tiles_h;
        const int tz = r % p.tiles_d;
        const int n = r / p.tiles_d;
        for (int kc = 0; kc < p.nkc; ++kc) {
          const uint32_t c = (uint32_t)(it * p.nkc + kc), slot = c % (uint32_t)p.nring, use = c / (uint32_t)p.nring;
          if (use > 0) mbar_wait(&plane_empty[slot], (use - 1) & 1u, 0);
          for (int q = 0; q < p.np; ++q) {
            const int idx = (int)slot * p.np + q;
#ifdef OFSV_STACK_PROBE
            if ((p.probe & 2) && it > 0) { mbar_expect_tx(&plane_full[idx], 0); continue; }
#endif
            mbar_expect_tx(&plane_full[idx], SK_HP_ROWS * ROWB);
            tma_load_5d(&tmA, &plane_full[idx], sP + (size_t)idx * p.plane_stride, kc * KC, tx * SK_HT_W - 1, ty * SK_HT_H - 1,
                        tz * p.td + p.dzmin + q, n);
          }
        }
      }
    }
  } else if (warp == 1 || (warp == 3 && p.ni > 1)) {
    // ================= MMA issuers (warp 1: issuer 0, warp 3: issuer 1): warp-uniform control flow, one elected lane issues =========
    const int iss = warp == 1 ? 0 : 1;
    const uint32_t leader = elect_one_sync();
    const uint32_t a_hi = kmajor_desc_hi<KC>(SK_HP_W * ROWB);      // 8-row groups of A are 10 halo rows apart
    const uint32_t b_hi = kmajor_desc_hi<KC>(8 * ROWB);
    const uint32_t plane_lo0 = kmajor_desc_lo(smem_u32(sP)), plane_step = (uint32_t)p.plane_stride >> 4;
    const uint32_t b_lo0 = kmajor_desc_lo(smem_u32(sB)), b_step = (uint32_t)p.b_stride >> 4;
    uint32_t bs = 0, bphase = 0, buf = 0, acc_use = 0;
    int it = 0;
    for (int st = blockIdx.x; st < total; st += gridDim.x, ++it) {
      for (int pass = 0; pass < p.npass; ++pass) {
        if (acc_use >= (uint32_t)p.nbuf) mbar_wait(&acc_empty[buf], ((acc_use / p.nbuf) - 1) & 1, 0);
        const uint32_t acc0 = tmem_base + buf * p.acc_stride;
        for (int kc = 0; kc < p.nkc; ++kc) {
          const uint32_t c = (uint32_t)(it * p.nkc + kc), slot = c % (uint32_t)p.nring, use = c / (uint32_t)p.nring;
          if (pass == 0)                                             // first use of this chunk's planes in this super-tile
            for (int q = 0; q < p.np; ++q) mbar_wait(&plane_full[(int)slot * p.np + q], use & 1u, 0);
          const uint32_t a_kc = plane_lo0 + slot * (uint32_t)p.np * plane_step;
          for (int g = p.group_first[pass]; g < p.group_first[pass + 1]; ++g) {
            const SkGroupRec gr = p.groups[g];
            if (!p.b_resident) mbar_wait(&b_full[bs], bphase, 0);
            else if (it == 0 && pass == 0 && kc == 0 && g == 0) mbar_wait(&b_full[0], 0u, 0);
            tcgen05_fence_after();
            if (leader) {
              const uint32_t b_lo = p.b_resident ? b_lo0 + (((uint32_t)gr.row0 + (uint32_t)(kc * gr.nslots * p.Cout_w)) * (uint32_t)ROWB >> 4)
                                                 : b_lo0 + bs * b_step;
              const uint32_t fresh_ok = kc == 0 ? 1u : 0u;
#ifdef OFSV_STACK_PROBE
              if (!(p.probe & 8))
#endif
              {
                // the op words come from the kernel-parameter bank (LDCU, ~100 cycles): they are fetched TWO ops ahead of their
                // MMAs, otherwise the single issuing thread waits for a constant load per op (84 cycles per MMA measured on
                // the N = 32 / 64 conv0 layers against 42-48 of operand fetch)
                const int ob = gr.op_begin[iss], no = gr.nops[iss];
                uint4 w_a = *reinterpret_cast<const uint4*>(p.ops[no > 0 ? ob : 0]);
                uint4 w_b = *reinterpret_cast<const uint4*>(p.ops[no > 1 ? ob + 1 : 0]);
                for (int i = 0; i < no; ++i) {
                  const uint4 w = w_a;
                  w_a = w_b;
                  if (i + 2 < no) w_b = *reinterpret_cast<const uint4*>(p.ops[ob + i + 2]);
                  const uint32_t a = a_kc + w.x, b = b_lo + w.y;
                  const uint32_t dcol = acc0 + (w.z & 0xFFFFu);
                  const uint32_t acc = (fresh_ok & (w.z >> 16)) ^ 1u;
                  umma_bf16_lohi(dcol, a, a_hi, b, b_hi, w.w, acc);
#pragma unroll
                  for (int k = 1; k < KC / 16; ++k) umma_bf16_lohi(dcol, a + 2 * k, a_hi, b + 2 * k, b_hi, w.w, 1u);
                }
              }
              if (!p.b_resident) tcgen05_commit(&b_empty[bs]);
            }
            __syncwarp();
            if (++bs == (uint32_t)p.nb) { bs = 0; bphase ^= 1u; }
          }
          if (pass == p.npass - 1) {                                 // this chunk's planes are dead: hand them to the producer
            if (leader) tcgen05_commit(&plane_empty[slot]);
            __syncwarp();
          }
        }
        if (leader) tcgen05_commit(&acc_full[buf]);
        __syncwarp();
        ++acc_use;
        if (++buf == (uint32_t)p.nbuf) buf = 0;
      }
    }
    (void)ngroups_all;
  } else if (warp >= SK_EPI_W0) {
    // ================= epilogue: 8 warps, two per TMEM lane quarter, splitting the column blocks of a pass =================
    const int qq = warp & 3, ew = warp - SK_EPI_W0, half = ew >> 2;
    const int row = qq * 32 + lane;
    const int rx = row & 7, ry = row >> 3;
    const int nch = p.Cout_w >> 4;
    const int nblk = p.P * p.td;
    const uint32_t stg = smem_u32(sS) + (uint32_t)ew * (uint32_t)p.stg_stride;
    const __nv_bfloat16* resb = reinterpret_cast<const __nv_bfloat16*>(residual);
    const float* resf = reinterpret_cast<const float*>(residual);
    const bool res_bf16 = p.has_residual && !p.shuffle, res_f32 = p.has_residual && p.shuffle;
    const uint32_t lane_tm = (uint32_t)(qq * 32) << 16;
    uint32_t acc_it = 0;
    bool store_pending = false;
    for (int st = blockIdx.x; st < total; st += gridDim.x) {
      int r = st;
      const int tx = r % p.tiles_w; r /= p.tiles_w;
      const int ty = r % p.tiles_h; r /= p.tiles_h;
      const int tz = r % p.tiles_d;
      const int n = r / p.tiles_d;
      const int ox = tx * SK_HT_W + rx, oy = ty * SK_HT_H + ry;
      const bool valid_xy = ox < p.Wo && oy < p.Ho;
      const int nj = min(p.td, p.Do - tz * p.td);
      for (int pass = 0; pass < p.npass; ++pass, ++acc_it) {
        const int buf = acc_it % p.nbuf;
        const uint32_t acc0 = tmem_base + buf * p.acc_stride + lane_tm;
        mbar_wait(&acc_full[buf], (acc_it / p.nbuf) & 1, 0);
        tcgen05_fence_after();
#ifdef OFSV_STACK_PROBE
        if (!(p.probe & 4))
#endif
        for (int blk = half; blk < nblk; blk += 2) {
          const int pp = blk / p.td, j = blk - pp * p.td;
          if (j >= nj) continue;
          const int ph = pass * p.P + pp;
          const int pz = (ph >> 2) & 1, py = (ph >> 1) & 1, px = ph & 1;
          const int oz = tz * p.td + j;
          const uint32_t tcol = acc0 + (uint32_t)(blk * p.Cout_w);
          if (p.epi_mode == SK_EPI_TMA_ROWS) {
            // ---- bf16 channels-last rows: this warp's 32 rows (4 y x 8 x) x stg_rowb bytes per TMA box
            const int cps = p.stg_rowb >> 5;                       // 16-column chunks per staged row piece
            const int64_t ro = ((((int64_t)n * p.Dy + oz) * p.Hy + oy) * p.Wy + ox) * p.Cout_s;   // residual layers: out_stride 1, nphase 1
            uint4 rn0 = make_uint4(0, 0, 0, 0), rn1 = rn0;
            if (res_bf16 && valid_xy) { rn0 = __ldg(reinterpret_cast<const uint4*>(resb + ro)); rn1 = __ldg(reinterpret_cast<const uint4*>(resb + ro) + 1); }
            for (int c = 0; c < nch; ++c) {
              const int c0 = c << 4;
              float v[16];
              tmem_ld16(tcol + (uint32_t)c0, v);
              const uint4 r0 = rn0, r1 = rn1;
              if (res_bf16 && valid_xy && c + 1 < nch) {
                rn0 = __ldg(reinterpret_cast<const uint4*>(resb + ro + c0 + 16)); rn1 = __ldg(reinterpret_cast<const uint4*>(resb + ro + c0 + 16) + 1);
              }
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + e4 * 4);
                const float4 p4 = *reinterpret_cast<const float4*>(sPrelu + c0 + e4 * 4);
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, sl[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float a = v[e4 * 4 + e] + bb[e];
                  v[e4 * 4 + e] = fmaxf(a, 0.0f) + sl[e] * fminf(a, 0.0f);
                }
              }
              if (res_bf16) {
                const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
                const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  v[2 * e] += __low2float(h0[e]); v[2 * e + 1] += __high2float(h0[e]);
                  v[8 + 2 * e] += __low2float(h1[e]); v[8 + 2 * e + 1] += __high2float(h1[e]);
                }
              }
              const int cc = c % cps;                              // chunk inside the staged piece
              if (cc == 0 && store_pending) {                      // the previous box must have left the staging tile
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
                store_pending = false;
              }
              // 16 B unit u of staged row `lane` lives at unit u ^ swz(lane) (TMA SWIZZLE_128B / 64B / 32B patterns)
              const uint32_t sw = p.stg_rowb == 128 ? (uint32_t)(lane & 7) : (p.stg_rowb == 64 ? (uint32_t)((lane >> 1) & 3) : (uint32_t)((lane >> 2) & 1));
              const uint32_t rowa = stg + (uint32_t)lane * (uint32_t)p.stg_rowb;
              sk_sts128(rowa + ((((uint32_t)(2 * cc)) ^ sw) << 4), sk_pack2(v[0], v[1]), sk_pack2(v[2], v[3]), sk_pack2(v[4], v[5]), sk_pack2(v[6], v[7]));
              sk_sts128(rowa + ((((uint32_t)(2 * cc + 1)) ^ sw) << 4), sk_pack2(v[8], v[9]), sk_pack2(v[10], v[11]), sk_pack2(v[12], v[13]), sk_pack2(v[14], v[15]));
              if (cc == cps - 1 || c == nch - 1) {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                  const int cbase = (c / cps) * (p.stg_rowb >> 1);
                  if (cbase < p.Cout_s) {
                    tma_store_5d(&tmO.m[p.out_stride == 1 ? 0 : ph], stg, cbase, tx * SK_HT_W, ty * SK_HT_H + 4 * qq, oz, n);
                    tma_store_commit();
                  }
                }
                store_pending = true;
              }
            }
          } else if (p.epi_mode == SK_EPI_TMA_SHUF) {
            // ---- fp32 depth-to-space heads: columns [parity (z,y,x)][8 ch]; this warp's 4 x 8 rows cover 8 (y) x 16 (x) x nz output
            //      voxels = one {32 floats, 4, 8, 1} box per z parity of the [N][Dy][Hy][Wy/4][32] view of the state tensor
            auto shuf_off = [&](int par) -> int64_t {
              const int zz = oz * (p.nd == 3 ? 2 : 1) + ((par >> 2) & 1), yy = oy * 2 + ((par >> 1) & 1), xx = ox * 2 + (par & 1);
              if (p.hfast) return ((((int64_t)n * p.Dy + zz) * p.Wy + xx) * p.Hy + yy) * 8;       // [N][D][W][H][8]
              return ((((int64_t)n * p.Dy + zz) * p.Hy + yy) * p.Wy + xx) * 8;
            };
            float4 fn[4];
            auto load_state = [&](int c) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float4* rp = reinterpret_cast<const float4*>(resf + shuf_off(2 * c + e));
                fn[2 * e] = __ldg(rp); fn[2 * e + 1] = __ldg(rp + 1);
              }
            };
            if (res_f32 && valid_xy) load_state(0);
            for (int c = 0; c < nch; ++c) {
              const int c0 = c << 4;
              float v[16];
              tmem_ld16(tcol + (uint32_t)c0, v);
              float4 fc[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) fc[e] = fn[e];
              if (res_f32 && valid_xy && c + 1 < nch) load_state(c + 1);
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + e4 * 4);
                v[e4 * 4] += b4.x; v[e4 * 4 + 1] += b4.y; v[e4 * 4 + 2] += b4.z; v[e4 * 4 + 3] += b4.w;
              }
              if (res_f32 && valid_xy) {          // fm = fm_prev + head (Flow-3D/model/IFNet.py:169-170)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  v[4 * e] = __fadd_rn(fc[e].x, v[4 * e]); v[4 * e + 1] = __fadd_rn(fc[e].y, v[4 * e + 1]);
                  v[4 * e + 2] = __fadd_rn(fc[e].z, v[4 * e + 2]); v[4 * e + 3] = __fadd_rn(fc[e].w, v[4 * e + 3]);
                }
              }
              // chunk c = parities (2c, 2c+1) = (cz, cy, x parity 0 / 1).  The staging tile holds ONE z parity: rows R = (y' in the
              // warp's 8 output rows) * 4 + (16-voxel x quarter), 128 B each; this chunk's 64 B = units (rx & 1) * 4 + 0..3 of row R
              const int cz = p.nd == 3 ? (c >> 1) : 0, cy = c & 1;
              if (cy == 0 && store_pending) {                      // the previous box must have left the staging tile
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
                store_pending = false;
              }
              if (!p.hfast) {
                const uint32_t R = (uint32_t)(((2 * (lane >> 3) + cy) << 2) + (rx >> 1));
                const uint32_t rowa = stg + R * 128u, sw = R & 7u, u0 = (uint32_t)(rx & 1) * 4u;
                sk_sts128(rowa + (((u0 + 0) ^ sw) << 4), __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
                sk_sts128(rowa + (((u0 + 1) ^ sw) << 4), __float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
                sk_sts128(rowa + (((u0 + 2) ^ sw) << 4), __float_as_uint(v[8]), __float_as_uint(v[9]), __float_as_uint(v[10]), __float_as_uint(v[11]));
                sk_sts128(rowa + (((u0 + 3) ^ sw) << 4), __float_as_uint(v[12]), __float_as_uint(v[13]), __float_as_uint(v[14]), __float_as_uint(v[15]));
              } else {
                // H-fastest state [N][D][W][H/4][32 floats]: box {32, 2 (quarters of the warp's 8 y'), 16 (x'), 1}; staged row
                // R = x' * 2 + (y' >> 2), this voxel's 8 floats = units (y' & 3) * 2 + {0, 1}; x parity 0 / 1 are different rows
                const uint32_t yl = (uint32_t)(2 * (lane >> 3) + cy), u0 = (yl & 3u) * 2u;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const uint32_t R = (uint32_t)(2 * rx + e) * 2u + (yl >> 2);
                  const uint32_t rowa = stg + R * 128u, sw = R & 7u;
                  sk_sts128(rowa + (((u0 + 0) ^ sw) << 4), __float_as_uint(v[8 * e]), __float_as_uint(v[8 * e + 1]), __float_as_uint(v[8 * e + 2]), __float_as_uint(v[8 * e + 3]));
                  sk_sts128(rowa + (((u0 + 1) ^ sw) << 4), __float_as_uint(v[8 * e + 4]), __float_as_uint(v[8 * e + 5]), __float_as_uint(v[8 * e + 6]), __float_as_uint(v[8 * e + 7]));
                }
              }
              if (cy == 1) {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                  const int zc = p.nd == 3 ? 2 * oz + cz : 0;
                  if (!p.hfast) tma_store_5d(&tmO.m[0], stg, 0, tx * (SK_HT_W / 2), 2 * (ty * SK_HT_H + 4 * qq), zc, n);
                  else tma_store_5d(&tmO.m[0], stg, 0, ty * (SK_HT_H / 2) + 2 * qq, tx * (2 * SK_HT_W), zc, n);
                  tma_store_commit();
                }
                store_pending = true;
              }
            }
          } else {
            // ---- direct per-thread stores (space-to-depth outputs, fp32 planar heads, anything without a tensor-store form)
            auto row_off = [&]() -> int64_t {
              if (p.out_s2d) return s2d_row(p.nd, n, oz, oy, ox, p.Dy, p.Hy, p.Wy) * p.Cout_s;
              return ((((int64_t)n * p.Dy + oz * p.out_stride + pz) * p.Hy + oy * p.out_stride + py) * p.Wy + ox * p.out_stride + px) * p.Cout_s;
            };
            auto shuf_off = [&](int par) -> int64_t {
              const int zz = oz * (p.nd == 3 ? 2 : 1) + ((par >> 2) & 1), yy = oy * 2 + ((par >> 1) & 1), xx = ox * 2 + (par & 1);
              if (p.hfast) return ((((int64_t)n * p.Dy + zz) * p.Wy + xx) * p.Hy + yy) * p.Cout_s;
              return ((((int64_t)n * p.Dy + zz) * p.Hy + yy) * p.Wy + xx) * p.Cout_s;
            };
            for (int c = 0; c < nch; ++c) {
              const int c0 = c << 4;
              float v[16];
              tmem_ld16(tcol + (uint32_t)c0, v);
              if (!valid_xy) continue;
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float a = v[e] + sBias[c0 + e];
                v[e] = fmaxf(a, 0.0f) + sPrelu[c0 + e] * fminf(a, 0.0f);
              }
              if (p.shuffle) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int64_t yo = shuf_off(2 * c + e);
                  if (res_f32) {
                    const float4 a0 = __ldg(reinterpret_cast<const float4*>(resf + yo)), a1 = __ldg(reinterpret_cast<const float4*>(resf + yo) + 1);
                    v[8 * e] = __fadd_rn(a0.x, v[8 * e]); v[8 * e + 1] = __fadd_rn(a0.y, v[8 * e + 1]); v[8 * e + 2] = __fadd_rn(a0.z, v[8 * e + 2]); v[8 * e + 3] = __fadd_rn(a0.w, v[8 * e + 3]);
                    v[8 * e + 4] = __fadd_rn(a1.x, v[8 * e + 4]); v[8 * e + 5] = __fadd_rn(a1.y, v[8 * e + 5]); v[8 * e + 6] = __fadd_rn(a1.z, v[8 * e + 6]); v[8 * e + 7] = __fadd_rn(a1.w, v[8 * e + 7]);
                  }
                  if (p.out_f32) {
                    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + yo);
                    o[0] = make_float4(v[8 * e], v[8 * e + 1], v[8 * e + 2], v[8 * e + 3]);
                    o[1] = make_float4(v[8 * e + 4], v[8 * e + 5], v[8 * e + 6], v[8 * e + 7]);
                  } else {
                    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + yo) =
                        make_uint4(sk_pack2(v[8 * e], v[8 * e + 1]), sk_pack2(v[8 * e + 2], v[8 * e + 3]), sk_pack2(v[8 * e + 4], v[8 * e + 5]), sk_pack2(v[8 * e + 6], v[8 * e + 7]));
                  }
                }
              } else {
                const int64_t yo = row_off() + c0;
                const int nstore = min(16, p.Cout_s - c0);
                if (nstore <= 0) continue;
                if (res_bf16) {
                  for (int e = 0; e < nstore; e += 8) {
                    const uint4 rr = __ldg(reinterpret_cast<const uint4*>(resb + yo + e));
                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { v[e + 2 * k] += __low2float(h[k]); v[e + 2 * k + 1] += __high2float(h[k]); }
                  }
                }
                if (p.out_f32) {
                  float* o = reinterpret_cast<float*>(y) + yo;
                  for (int e = 0; e < nstore; e += 4) *reinterpret_cast<float4*>(o + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                } else {
                  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(y) + yo;
                  for (int e = 0; e < nstore; e += 8)
                    *reinterpret_cast<uint4*>(o + e) = make_uint4(sk_pack2(v[e], v[e + 1]), sk_pack2(v[e + 2], v[e + 3]), sk_pack2(v[e + 4], v[e + 5]), sk_pack2(v[e + 6], v[e + 7]));
                }
              }
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[buf])) : "memory");
      }
    }
    if (lane == 0) tma_store_wait_all();       // the staging tile must outlive the last store's reads
    __syncwarp();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// weights fp32 tap form [T][Cin_s][Cout_w] -> bf16 blocks of [Cout_w][KC]: out block b = (tap, kc) = table[b]
struct SkPackTable { uint16_t blk[OFSV_MAX_TAPS * SK_MAX_KC]; };   // (tap << 6) | kc
__global__ void conv_pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, const SkPackTable T, int nblocks,
                                         int Cin_s, int Cout_w, int KC) {
  const int per = Cout_w * KC;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)nblocks * per; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per), e = (int)(i - (int64_t)b * per);
    const int r = e / KC, k = e - r * KC;
    const int tap = T.blk[b] >> 6, kc = T.blk[b] & 63;
    out[i] = __float2bfloat16_rn(__ldg(w + ((int64_t)tap * Cin_s + kc * KC + k) * Cout_w + r));
  }
}

static std::atomic<int> g_stack_epi{-1};     // -1 auto; 0 forces the direct epilogue (ofsv_set_tuning("stack_epilogue", v))
static std::atomic<int> g_stack_td{0};
void ofsv_set_wgrad_brick(int v);            // conv_bwd.cu
extern std::atomic<int> g_tc_pair, g_tc_stages;   // conv_tc.cu
extern std::atomic<int> g_warp_slab;         // warp3d_slab.cu       // 0 auto; 1 / 2 / 4 forces the super-tile depth when it is feasible

struct SkConfig {
  int td, nring, nb, nbuf, epi_mode, stg_stride, stg_rowb, plane_stride, b_stride, b_resident;
  size_t smem;
  double cost;
};

// Chooses the super-tile depth and the shared-memory carve-up.  `sms` and d->N only enter the wave count of the cost model.
static bool sk_configure(const ofsv_conv_desc* d, const SkPlan& pl, int sms, SkConfig* out) {
  const int KC = pl.KC, ROWB = KC * 2;
  const size_t smem_cap = 227 * 1024 - 1024;           // - alignment slack
  const size_t misc = (SK_MAX_PBARS + SK_MAX_RING + 2 * SK_MAX_BST + 4) * 8 + 64 + 2 * 128 * 4;
  const int plane_stride = (SK_HP_ROWS * ROWB + 1023) & ~1023;
  int max_slots = 0;
  for (int g = 0; g < pl.ngroups; ++g) max_slots = pl.g[g].nslots > max_slots ? pl.g[g].nslots : max_slots;
  const int b_stride = (max_slots * d->Cout_w * ROWB + 1023) & ~1023;
  // epilogue candidates: the TMA-store form of this layer type (if it has one), then per-thread stores
  int epis[2], nepi = 0;
  if (g_stack_epi.load(std::memory_order_relaxed) != 0) {
    if (d->out_shuffle && d->out_dtype == OFSV_F32 && (d->out_shuffle_hfast ? d->Hy : d->Wy) % 4 == 0 && !d->has_prelu) epis[nepi++] = SK_EPI_TMA_SHUF;
    else if (!d->out_shuffle && !d->out_s2d && d->out_dtype == OFSV_BF16 && (d->Cout_s * 2) % 32 == 0 && d->Cout_w == d->Cout_s) epis[nepi++] = SK_EPI_TMA_ROWS;
  }
  epis[nepi++] = SK_EPI_DIRECT;
  bool found = false;
  SkConfig best;
  memset(&best, 0, sizeof(best));
  const int forced_td = g_stack_td.load(std::memory_order_relaxed);
  for (int td = (d->nd == 3 ? 4 : 1); td >= 1; td >>= 1) {
    if (forced_td && td != forced_td && td > 1) continue;
    if (td > 1 && td > d->Do) continue;
    const int cols = pl.P * td * d->Cout_w;
    if (cols > 256 && td > 1) continue;                // keep the TMEM accumulators double-buffered (epilogue under the next pass)
    if (cols > 512) continue;
    const int np = td + pl.dzmax - pl.dzmin;
    SkOp ops[SK_MAX_OPS];
    int op_first[SK_MAX_GROUPS * SK_NI + 1];
    const int nops = sk_build_ops(d, pl, td, ops, op_first);
    if (nops < 0) continue;
    // tensor / operand-port time of all MMAs vs the issue time of the busier issuer (~50 cycles of scalar work per op)
    double mma = 0.0, iss_t[SK_NI] = {0.0, 0.0};
    for (int i = 0; i < nops; ++i) {
      const double c = sk_mma_cycles(sk_op_cols(ops[i], d->Cout_w)) * (KC / 16);
      mma += c;
      iss_t[ops[i].issuer] += c + 50.0;
    }
    { const double im = iss_t[0] > iss_t[1] ? iss_t[0] : iss_t[1]; mma = (mma > im ? mma : im) * pl.nkc; }
    const int64_t nst = cdiv(d->Wo, SK_HT_W) * cdiv(d->Ho, SK_HT_H) * cdiv(d->Do, td) * d->N;
    // plane-chunk ring depth: single-pass layers stream chunks (2 slots overlap the loads of chunk c+1 with the MMAs of chunk c,
    // more than 2 * nkc never helps); multi-pass layers need every chunk of a super-tile resident (multiples of nkc)
    int rings[4], nrings = 0;
    if (pl.npass == 1) { rings[nrings++] = pl.nkc >= 2 ? 2 : 2; rings[nrings++] = 1; }
    else { rings[nrings++] = 2 * pl.nkc; rings[nrings++] = pl.nkc; }
    for (int ei = 0; ei < nepi; ++ei)
      for (int ri = 0; ri < nrings; ++ri) {
        const int epi = epis[ei], nring = rings[ri];
        if (nring > SK_MAX_RING || nring * np > SK_MAX_PBARS) continue;
        const size_t planes = (size_t)nring * np * plane_stride;
        // staging: full rows when they fit next to >= 4 weight stages, else half rows
        int stg_rowb = 0, stg_stride = 0;
        if (epi == SK_EPI_TMA_SHUF) { stg_rowb = 128; stg_stride = 4096; }
        else if (epi == SK_EPI_TMA_ROWS) {
          const int rb = d->Cout_s * 2, full = rb % 128 == 0 ? 128 : (rb % 64 == 0 ? 64 : 32);
          stg_rowb = full;
          if (planes + 4 * (size_t)b_stride + 8 * 32 * (size_t)full + misc > smem_cap && full > 64) stg_rowb = 64;
          stg_stride = (32 * stg_rowb + 1023) & ~1023;
        }
        const size_t fixed = planes + (size_t)SK_EPI_WARPS * stg_stride + misc;
        const size_t w_all = (((size_t)d->nphase * d->ntaps * d->Cin_s * d->Cout_w * 2) + 1023) & ~(size_t)1023;
        const bool resident = fixed + w_all <= smem_cap && w_all < (1u << 20);
        if (!resident && fixed + 2 * (size_t)b_stride > smem_cap) continue;
        int nb = (int)((smem_cap - fixed) / b_stride);
        nb = nb > SK_MAX_BST ? SK_MAX_BST : nb;
        // cost model: rounds x (tensor time + exposed plane loads + a thin weight ring + scattered stores)
        const bool overlapped = pl.npass == 1 ? nring >= 2 : (nring >= 2 * pl.nkc || pl.nkc >= 2);
        const double exposed = overlapped ? 0.0 : (double)np * pl.nkc * plane_stride / 30.0;
        // a streamed stage costs the issuing thread ~370 cycles of barrier / commit latency that resident weights do not pay
        const double stall = resident ? 0.0 : 370.0 * pl.ngroups * pl.nkc + (nb < 3 ? 0.15 * mma : 0.0);
        const double scatter = epi == SK_EPI_DIRECT ? 0.25 * mma : 0.0;
        const double cost = (double)cdiv(nst, sms) * (mma + exposed + stall + scatter + 1500.0);
        if (!found || cost < best.cost) {
          found = true;
          best.td = td; best.nring = nring; best.nb = nb; best.nbuf = cols <= 256 ? 2 : 1; best.epi_mode = epi;
          best.stg_stride = stg_stride; best.stg_rowb = stg_rowb; best.plane_stride = plane_stride; best.b_stride = b_stride;
          best.b_resident = resident ? 1 : 0;
          if (resident) { best.nb = 1; best.b_stride = (int)w_all; }      // the B area is the whole packed weight tensor
          best.smem = fixed + (resident ? w_all : (size_t)nb * b_stride) + 1024; best.cost = cost;
        }
      }
    if (forced_td && td == forced_td && found) break;
  }
  if (found) *out = best;
  return found;
}

static std::atomic<int> g_conv0_ring{0};     // 1: the 2^d-tap space-to-depth conv0 layers on the legacy plane-ring kernel (A/B)
static bool sk_wants_ring(const ofsv_conv_desc* d) { return g_conv0_ring.load(std::memory_order_relaxed) != 0 && d->nphase == 1 && d->ntaps <= 8; }

template <int KC>
static int sk_launch(const SkParams& P, const CUtensorMap& tmA, const CUtensorMap& tmB, const SkOutMaps& tmO, const float* bias,
                     const float* prelu, const void* residual, void* y, int grid, size_t smem, cudaStream_t st) {
  static std::atomic<uint64_t> attr_done{0};
  if (int e = ensure_dyn_smem(attr_done, conv_stack_kernel<KC>, 227 * 1024, "ofsv_conv_halo")) return e;
  conv_stack_kernel<KC><<<grid, SK_THREADS, smem, st>>>(tmA, tmB, tmO, P, bias, prelu, residual, y);
  return check_launch("conv_stack_kernel");
}

// launch-plan cache (see ofsv_conv_halo): a few dozen distinct (layer, shape) descriptors per model
struct SkCacheEntry { bool used; int sms, tuning; ofsv_conv_desc d; SkPlan pl; SkConfig cfg; SkParams P; };
constexpr int SK_CACHE_N = 128;
static SkCacheEntry* g_sk_cache = nullptr;
static int g_sk_cache_next = 0;
static std::mutex g_sk_cache_mu;
static int sk_tuning_stamp() { return (g_stack_epi.load(std::memory_order_relaxed) + 2) * 16 + g_stack_td.load(std::memory_order_relaxed); }
static bool sk_cache_get(const ofsv_conv_desc* d, int sms, SkPlan* pl, SkConfig* cfg, SkParams* P) {
  std::lock_guard<std::mutex> lk(g_sk_cache_mu);
  if (!g_sk_cache) return false;
  const int stamp = sk_tuning_stamp();
  for (int i = 0; i < SK_CACHE_N; ++i) {
    const SkCacheEntry& e = g_sk_cache[i];
    if (e.used && e.sms == sms && e.tuning == stamp && memcmp(&e.d, d, sizeof(*d)) == 0) { *pl = e.pl; *cfg = e.cfg; *P = e.P; return true; }
  }
  return false;
}
static void sk_cache_put(const ofsv_conv_desc* d, int sms, const SkPlan& pl, const SkConfig& cfg, const SkParams& P) {
  std::lock_guard<std::mutex> lk(g_sk_cache_mu);
  if (!g_sk_cache) g_sk_cache = new SkCacheEntry[SK_CACHE_N]();
  SkCacheEntry& e = g_sk_cache[g_sk_cache_next];
  g_sk_cache_next = (g_sk_cache_next + 1) % SK_CACHE_N;
  e.used = true; e.sms = sms; e.tuning = sk_tuning_stamp(); e.d = *d; e.pl = pl; e.cfg = cfg; e.P = P;
}

}  // namespace ofsv

using namespace ofsv;

extern "C" int ofsv_set_tuning(const char* key, int value) {
  if (!key) return OFSV_EINVAL;
  if (!strcmp(key, "stack_epilogue")) { g_stack_epi.store(value); return OFSV_OK; }
  if (!strcmp(key, "stack_td")) { g_stack_td.store(value); return OFSV_OK; }
  if (!strcmp(key, "warp_slab")) { g_warp_slab.store(value); return OFSV_OK; }
  if (!strcmp(key, "conv0_ring")) { g_conv0_ring.store(value); return OFSV_OK; }
  if (!strcmp(key, "wgrad_brick")) { ofsv_set_wgrad_brick(value); return OFSV_OK; }
  if (!strcmp(key, "tc_pair")) { g_tc_pair.store(value); return OFSV_OK; }
  if (!strcmp(key, "tc_stages")) { g_tc_stages.store(value); return OFSV_OK; }
  set_error("ofsv_set_tuning: unknown key '%s'", key);
  return OFSV_EINVAL;
}

extern "C" int ofsv_conv_halo_weight_layout(const ofsv_conv_desc* d) {
  if (!d) return OFSV_EINVAL;
  if (sk_wants_ring(d)) return OFSV_WL_TAP;
  SkPlan pl;
  return sk_make_plan(d, &pl) ? OFSV_WL_STACK : OFSV_WL_TAP;
}

// (tap, kc) order of the bf16 blocks of `layout` for the layer structure of `d`; returns the block count or a negative error.
static int sk_pack_table(const ofsv_conv_desc* d, int layout, SkPackTable* T, int* KC_out) {
  OFSV_REQUIRE(d != nullptr, "ofsv_conv_pack_weights: null descriptor");
  OFSV_REQUIRE(d->Cin_s >= 16 && d->Cin_s % 16 == 0 && d->Cout_w >= 16 && d->Cout_w % 16 == 0, "ofsv_conv_pack_weights: bad channels");
  OFSV_REQUIRE(d->nphase >= 1 && d->ntaps >= 1 && d->nphase * d->ntaps <= OFSV_MAX_TAPS, "ofsv_conv_pack_weights: nphase*ntaps out of range");
  OFSV_REQUIRE(layout == OFSV_WL_TAP || layout == OFSV_WL_STACK, "ofsv_conv_pack_weights: bad layout");
  memset(T, 0, sizeof(*T));
  int KC, nkc, nblocks = 0;
  const int ntap = d->nphase * d->ntaps;
  if (layout == OFSV_WL_TAP) {
    KC = d->Cin_s % 64 == 0 ? 64 : (d->Cin_s % 32 == 0 ? 32 : 16);
    nkc = d->Cin_s / KC;
    // the table packs (tap << 6) | kc into 16 bits and holds OFSV_MAX_TAPS * SK_MAX_KC entries: the per-tap kernel itself has no limit
    // on Cin_s (UPFlow's dense estimator reaches 576 channels = 9 chunks of 64 with 9 taps)
    OFSV_REQUIRE(nkc <= 64 && ntap * nkc <= OFSV_MAX_TAPS * SK_MAX_KC, "ofsv_conv_pack_weights: too many channel chunks (%d taps x %d chunks)", ntap, nkc);
    for (int t = 0; t < ntap; ++t)
      for (int kc = 0; kc < nkc; ++kc) T->blk[nblocks++] = (uint16_t)((t << 6) | kc);
  } else {
    SkPlan pl;
    if (!sk_make_plan(d, &pl)) { set_error("ofsv_conv_pack_weights: layer has no stacked form"); return OFSV_ENOSUP; }
    KC = pl.KC; nkc = pl.nkc;
    for (int g = 0; g < pl.ngroups; ++g)
      for (int kc = 0; kc < nkc; ++kc)
        for (int s = 0; s < pl.g[g].nslots; ++s) T->blk[nblocks++] = (uint16_t)((pl.g[g].slots[s].tap << 6) | kc);
    OFSV_REQUIRE(nblocks == ntap * nkc, "ofsv_conv_pack_weights: internal error (slot count %d != %d)", nblocks, ntap * nkc);
  }
  *KC_out = KC;
  return nblocks;
}

extern "C" int ofsv_conv_pack_weights(const ofsv_conv_desc* d, const float* w_tap, void* w_out, int layout, void* stream) {
  OFSV_REQUIRE(d && w_tap && w_out, "ofsv_conv_pack_weights: null pointer");
  SkPackTable T;
  int KC = 0;
  const int nblocks = sk_pack_table(d, layout, &T, &KC);
  if (nblocks < 0) return nblocks;
  const int64_t total = (int64_t)nblocks * d->Cout_w * KC;
  const int grid = (int)(cdiv(total, 256) < 1184 ? cdiv(total, 256) : 1184);
  conv_pack_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w_tap, reinterpret_cast<__nv_bfloat16*>(w_out), T, nblocks, d->Cin_s, d->Cout_w, KC);
  return check_launch("conv_pack_weights_kernel");
}

// Batched form for the training step (every layer of a block is re-packed every step): ofsv_conv_pack_record fills one HOST record
// per (layer, layout), the caller uploads the array once, ofsv_conv_pack_weights_batched re-packs all of them in ONE launch.
static_assert(sizeof(ofsv_pack_rec) == 16 + 16 + sizeof(SkPackTable), "ofsv_pack_rec layout");
__global__ void __launch_bounds__(256) conv_pack_weights_batched_kernel(const ofsv_pack_rec* __restrict__ recs) {
  const ofsv_pack_rec& r = recs[blockIdx.y];
  const int per = r.Cout_w * r.KC;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(r.w_out);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)r.nblocks * per; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per), e = (int)(i - (int64_t)b * per);
    const int row = e / r.KC, k = e - row * r.KC;
    const int tap = r.blk[b] >> 6, kc = r.blk[b] & 63;
    out[i] = __float2bfloat16_rn(__ldg(r.w_tap + ((int64_t)tap * r.Cin_s + kc * r.KC + k) * r.Cout_w + row));
  }
}

extern "C" int ofsv_conv_pack_record(const ofsv_conv_desc* d, int layout, const float* w_tap, void* w_out, ofsv_pack_rec* rec) {
  OFSV_REQUIRE(w_tap && w_out && rec, "ofsv_conv_pack_record: null pointer");
  SkPackTable T;
  int KC = 0;
  const int nblocks = sk_pack_table(d, layout, &T, &KC);
  if (nblocks < 0) return nblocks;
  rec->w_tap = w_tap; rec->w_out = w_out;
  rec->nblocks = nblocks; rec->Cin_s = d->Cin_s; rec->Cout_w = d->Cout_w; rec->KC = KC;
  memcpy(rec->blk, T.blk, sizeof(T.blk));
  return OFSV_OK;
}

extern "C" int ofsv_conv_pack_weights_batched(const ofsv_pack_rec* recs_dev, int nrec, void* stream) {
  OFSV_REQUIRE(nrec >= 0 && nrec <= 65535, "ofsv_conv_pack_weights_batched: bad record count");
  if (nrec == 0) return OFSV_OK;
  OFSV_REQUIRE(recs_dev != nullptr && (reinterpret_cast<uintptr_t>(recs_dev) & 7u) == 0, "ofsv_conv_pack_weights_batched: null / misaligned record table");
  conv_pack_weights_batched_kernel<<<dim3(64, (unsigned)nrec), 256, 0, (cudaStream_t)stream>>>(recs_dev);
  return check_launch("conv_pack_weights_batched_kernel");
}

// Host-only check of the op list (no GPU): every (phase, tap, output slice) term must be covered exactly once, the first MMA
// into every column block must be the (only) fresh one, runs must be contiguous in weight rows and TMEM columns.
extern "C" int ofsv_conv_stack_selfcheck(const ofsv_conv_desc* d, int td, int* nops_out, double* mma_cycles_out) {
  OFSV_REQUIRE(d != nullptr, "ofsv_conv_stack_selfcheck: null descriptor");
  SkPlan pl;
  if (!sk_make_plan(d, &pl)) { set_error("ofsv_conv_stack_selfcheck: no stacked form"); return OFSV_ENOSUP; }
  SkOp ops[SK_MAX_OPS];
  int op_first[SK_MAX_GROUPS * SK_NI + 1];
  const int nops = sk_build_ops(d, pl, td, ops, op_first);
  if (nops < 0) { set_error("ofsv_conv_stack_selfcheck: op table overflow"); return OFSV_ENOSUP; }
  // covered[ph * ntaps + t][j]
  int covered[OFSV_MAX_TAPS][8];
  memset(covered, 0, sizeof(covered));
  double cyc = 0.0;
  for (int pass = 0; pass < pl.npass; ++pass) {
    bool init[64] = {false};
    for (int g = pl.group_first[pass]; g < pl.group_first[pass + 1]; ++g)
      for (int i = op_first[g * SK_NI]; i < op_first[(g + 1) * SK_NI]; ++i) {
        const SkOp& o = ops[i];
        const SkGroup& G = pl.g[g];
        OFSV_REQUIRE(o.group == g && o.nsl >= 1 && o.slot0 + o.nsl <= G.nslots && o.nsl * d->Cout_w <= 256, "selfcheck: bad op %d", i);
        OFSV_REQUIRE(sk_op_cols(o, d->Cout_w) >= 16 && sk_op_cols(o, d->Cout_w) % 16 == 0 && o.lo % 16 == 0 && (!o.fresh || (o.lo == 0 && o.hi == 0)),
                     "selfcheck: op %d trimmed to a bad column range", i);
        for (int k = 0; k < o.nsl; ++k) {
          const SkSlot& s = G.slots[o.slot0 + k];
          const int j = o.q + pl.dzmin - s.oz, col = s.p * td + j;
          OFSV_REQUIRE(j >= 0 && j < td, "selfcheck: op %d slot %d leaves the super-tile", i, k);
          OFSV_REQUIRE(col == o.col0 + k, "selfcheck: op %d columns not contiguous", i);
          OFSV_REQUIRE((!init[col]) == (o.fresh != 0), "selfcheck: op %d fresh flag inconsistent", i);
          // the columns this op really multiplies inside slot k must include every column that can be non-zero
          const int lo_k = k == 0 ? o.lo : 0, hi_k = k == o.nsl - 1 ? d->Cout_w - o.hi : d->Cout_w;
          OFSV_REQUIRE(lo_k <= s.c_lo && hi_k >= s.c_hi, "selfcheck: op %d skips non-zero columns of slot %d", i, k);
          covered[s.tap][j] += 1;
        }
        for (int k = 0; k < o.nsl; ++k) init[o.col0 + k] = true;
        cyc += sk_mma_cycles(sk_op_cols(o, d->Cout_w));
      }
    for (int c = 0; c < pl.P * td; ++c) OFSV_REQUIRE(init[c], "selfcheck: column block %d of pass %d never written", c, pass);
  }
  for (int t = 0; t < d->nphase * d->ntaps; ++t)
    for (int j = 0; j < td; ++j) OFSV_REQUIRE(covered[t][j] == 1, "selfcheck: tap %d slice %d covered %d times", t, j, covered[t][j]);
  if (nops_out) *nops_out = nops;
  if (mma_cycles_out) *mma_cycles_out = cyc * pl.nkc * (pl.KC / 16);
  return OFSV_OK;
}

// Human-readable launch configuration of ofsv_conv_halo for this descriptor (tests); no GPU work.
extern "C" int ofsv_conv_halo_describe(const ofsv_conv_desc* d, char* buf, int buflen) {
  OFSV_REQUIRE(d && buf && buflen > 0, "ofsv_conv_halo_describe: null pointer");
  if (sk_wants_ring(d)) { snprintf(buf, buflen, "ring"); return OFSV_OK; }
  SkPlan pl;
  if (!sk_make_plan(d, &pl)) { snprintf(buf, buflen, "unsupported"); return OFSV_OK; }
  SkConfig cfg;
  if (!sk_configure(d, pl, device_num_sms(), &cfg)) { snprintf(buf, buflen, "does-not-fit"); return OFSV_OK; }
  SkOp ops[SK_MAX_OPS];
  int op_first[SK_MAX_GROUPS * SK_NI + 1];
  const int nops = sk_build_ops(d, pl, cfg.td, ops, op_first);
  double mma = 0.0, ideal = 0.0;
  for (int i = 0; i < nops; ++i) { mma += sk_mma_cycles(sk_op_cols(ops[i], d->Cout_w)); ideal += sk_op_cols(ops[i], d->Cout_w) / 2.0; }
  snprintf(buf, buflen, "stack KC=%d nkc=%d P=%d npass=%d groups=%d td=%d nring=%d nb=%d%s nbuf=%d epi=%d rowb=%d smem=%zu ops=%d tensor_frac=%.2f",
           pl.KC, pl.nkc, pl.P, pl.npass, pl.ngroups, cfg.td, cfg.nring, cfg.nb, cfg.b_resident ? "(resident)" : "", cfg.nbuf, cfg.epi_mode, cfg.stg_rowb, cfg.smem, nops, ideal / mma);
  return OFSV_OK;
}

// Same contract as ofsv_conv_tc for the stride-1 layers; `w` in the layout ofsv_conv_halo_weight_layout(d) names
// (ofsv_conv_pack_weights produces it).  Returns OFSV_ENOSUP (nothing launched) for layers outside the kernels' domain.
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
  if (d->has_residual && !d->out_shuffle && (d->nphase != 1 || d->out_stride != 1)) { set_error("ofsv_conv_halo: residual needs a plain stride-1 layer"); return OFSV_ENOSUP; }
  if (d->out_shuffle && (d->out_shuffle != 8 || d->nphase != 1 || d->Cout_w != 8 * (1 << d->nd) || d->Cout_s != 8)) {
    set_error("ofsv_conv_halo: bad depth-to-space configuration");
    return OFSV_EINVAL;
  }
  if (d->out_s2d && (d->nphase != 1 || d->out_stride != 1 || d->has_residual || d->out_shuffle || d->Hy % 2 || d->Wy % 2 ||
                     (d->nd == 3 && d->Dy % 2))) {
    set_error("ofsv_conv_halo: bad space-to-depth output configuration");
    return OFSV_EINVAL;
  }
  if (d->out_shuffle_hfast && (!d->out_shuffle || d->out_dtype != OFSV_F32)) {
    set_error("ofsv_conv_halo: out_shuffle_hfast needs the fp32 depth-to-space output");
    return OFSV_EINVAL;
  }
  for (int i = 0; i < d->nphase * d->ntaps; ++i) {
    const int8_t* o = d->tap_off[i];
    if (o[0] < -1 || o[0] > 1 || o[1] < -1 || o[1] > 1 || o[2] < -1 || o[2] > 1) {
      set_error("ofsv_conv_halo: tap offset outside {-1,0,1}");
      return OFSV_ENOSUP;
    }
  }
  if (sk_wants_ring(d)) return conv_halo_ring(d, x, w, bias, prelu, residual, y, stream);   // 2^d-tap space-to-depth conv0 layers

  PFN_encodeTiled encode = get_tensor_map_encoder();
  if (!encode) { set_error("ofsv_conv_halo: cuTensorMapEncodeTiled unavailable"); return OFSV_ECUDA; }
  const int sms = device_num_sms();
  // plan / configuration / MMA list are pure functions of (descriptor, SM count): ~15 us of host work per launch, which is what
  // bounds the small configurations (a 160x224 frame pair is 36 conv launches of ~10 us of GPU time each) — remembered per
  // descriptor; only the tensor maps (which hold the pointers) are built per call
  SkPlan pl;
  SkConfig cfg;
  SkParams P;
  if (!sk_cache_get(d, sms, &pl, &cfg, &P)) {
  if (!sk_make_plan(d, &pl)) { set_error("ofsv_conv_halo: layer has no stacked form"); return OFSV_ENOSUP; }
  if (!sk_configure(d, pl, sms, &cfg)) { set_error("ofsv_conv_halo: layer does not fit (Cin_s=%d Cout_w=%d)", d->Cin_s, d->Cout_w); return OFSV_ENOSUP; }

  const int KC = pl.KC, ROWB = KC * 2;
  memset(&P, 0, sizeof(P));
  P.N = d->N; P.Do = d->Do; P.Ho = d->Ho; P.Wo = d->Wo; P.Dy = d->Dy; P.Hy = d->Hy; P.Wy = d->Wy;
  P.Cout_s = d->Cout_s; P.Cout_w = d->Cout_w; P.out_stride = d->out_stride; P.nd = d->nd;
  P.nkc = pl.nkc; P.td = cfg.td; P.np = cfg.td + pl.dzmax - pl.dzmin; P.dzmin = pl.dzmin; P.P = pl.P; P.npass = pl.npass; P.nring = cfg.nring;
  P.tiles_w = (int)cdiv(d->Wo, SK_HT_W); P.tiles_h = (int)cdiv(d->Ho, SK_HT_H); P.tiles_d = (int)cdiv(d->Do, cfg.td);
  P.nb = cfg.nb; P.plane_stride = cfg.plane_stride; P.b_stride = cfg.b_stride; P.stg_stride = cfg.stg_stride; P.stg_rowb = cfg.stg_rowb;
  P.nbuf = cfg.nbuf; P.acc_stride = 256; P.b_resident = cfg.b_resident; P.ni = sk_issuers(pl, cfg.td);
  P.has_prelu = d->has_prelu; P.has_residual = d->has_residual; P.out_f32 = d->out_dtype == OFSV_F32;
  P.shuffle = d->out_shuffle; P.out_s2d = d->out_s2d; P.epi_mode = cfg.epi_mode; P.hfast = d->out_shuffle_hfast;
#ifdef OFSV_STACK_PROBE   // probe builds only (tests/ab_build.sh): timing experiments that produce WRONG results; never in the shipped library
  { const char* e = getenv("OFSV_STACK_PROBE_BITS"); P.probe = e ? atoi(e) : 0; }
#endif
  {
    SkOp ops[SK_MAX_OPS];
    int op_first[SK_MAX_GROUPS * SK_NI + 1];
    const int nops = sk_build_ops(d, pl, cfg.td, ops, op_first);
    OFSV_REQUIRE(nops > 0, "ofsv_conv_halo: internal error (op list)");
    for (int i = 0; i <= pl.npass; ++i) P.group_first[i] = (uint16_t)pl.group_first[i];
    for (int g = 0; g < pl.ngroups; ++g) {
      P.groups[g].row0 = (uint32_t)pl.g[g].row0;
      for (int w = 0; w < SK_NI; ++w) {
        P.groups[g].op_begin[w] = (uint8_t)op_first[g * SK_NI + w];
        P.groups[g].nops[w] = (uint8_t)(op_first[g * SK_NI + w + 1] - op_first[g * SK_NI + w]);
      }
      P.groups[g].nslots = (uint8_t)pl.g[g].nslots;
    }
    const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);   // fp32 accumulate, bf16 x bf16, K-major, M = 128
    for (int i = 0; i < nops; ++i) {
      const SkOp& o = ops[i];
      const uint32_t a_off = (uint32_t)(((o.oy + 1) * SK_HP_W + (o.ox + 1)) * ROWB) >> 4;
      const uint32_t ncols = (uint32_t)sk_op_cols(o, d->Cout_w);
      P.ops[i][0] = a_off + (uint32_t)o.q * ((uint32_t)cfg.plane_stride >> 4);
      P.ops[i][1] = (uint32_t)((o.slot0 * d->Cout_w + o.lo) * ROWB) >> 4;      // o.lo % 8 == 0: whole 8-row swizzle atoms
      P.ops[i][2] = (uint32_t)(o.col0 * d->Cout_w + o.lo) | ((uint32_t)(o.fresh ? 1 : 0) << 16);
      P.ops[i][3] = idesc0 | ((ncols >> 3) << 17);
      OFSV_REQUIRE(ncols >= 16 && ncols % 16 == 0 && ncols <= 256 && o.col0 * d->Cout_w + o.lo + ncols <= 512, "ofsv_conv_halo: internal error (op encoding)");
    }
  }
#ifndef OFSV_STACK_PROBE
  sk_cache_put(d, sms, pl, cfg, P);
#endif
  }
  const int KC = pl.KC;
  const int64_t total = (int64_t)P.tiles_w * P.tiles_h * P.tiles_d * d->N;
  OFSV_REQUIRE(total < (1ll << 31), "ofsv_conv_halo: too many super-tiles");

  const CUtensorMapSwizzle swz = KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap tmA, tmB;
  SkOutMaps tmO;
  memset(&tmO, 0, sizeof(tmO));
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  {
    const cuuint64_t gdim[5] = {(cuuint64_t)d->Cin_s, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->Di, (cuuint64_t)d->N};
    const cuuint64_t es = 2;
    const cuuint64_t gstr[4] = {d->Cin_s * es, (cuuint64_t)d->Wi * d->Cin_s * es, (cuuint64_t)d->Hi * d->Wi * d->Cin_s * es,
                                (cuuint64_t)d->Di * d->Hi * d->Wi * d->Cin_s * es};
    const cuuint32_t box[5] = {(cuuint32_t)KC, SK_HP_W, SK_HP_H, 1, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("ofsv_conv_halo: cuTensorMapEncodeTiled(A) failed with %d", (int)r); return OFSV_ECUDA; }
  }
  {
    const cuuint64_t rows = (cuuint64_t)d->nphase * d->ntaps * pl.nkc * d->Cout_w;
    const cuuint64_t gdim[2] = {(cuuint64_t)KC, rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)KC * 2};
    const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)d->Cout_w};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("ofsv_conv_halo: cuTensorMapEncodeTiled(B) failed with %d", (int)r); return OFSV_ECUDA; }
  }
  if (cfg.epi_mode == SK_EPI_TMA_ROWS) {
    // one map per output parity: [N][Dv][Hv][Wv][Cout_s] view of y with the parity folded into the base pointer
    const int os = d->out_stride;
    const CUtensorMapSwizzle oswz = cfg.stg_rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (cfg.stg_rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const int nmaps = os == 1 ? 1 : d->nphase;
    for (int ph = 0; ph < nmaps; ++ph) {
      const int pz = (ph >> 2) & 1, py = (ph >> 1) & 1, px = ph & 1;
      char* base = reinterpret_cast<char*>(y) + ((((int64_t)pz * d->Hy + py) * d->Wy + px) * d->Cout_s) * 2;
      const cuuint64_t cs = (cuuint64_t)d->Cout_s * 2;
      const cuuint64_t gdim[5] = {(cuuint64_t)d->Cout_s, (cuuint64_t)(os == 1 ? d->Wy : d->Wo), (cuuint64_t)(os == 1 ? d->Hy : d->Ho),
                                  (cuuint64_t)(os == 1 ? d->Dy : d->Do), (cuuint64_t)d->N};
      const cuuint64_t gstr[4] = {cs * os, cs * d->Wy * os, cs * d->Wy * d->Hy * (d->nd == 3 ? os : 1), cs * d->Wy * d->Hy * d->Dy};
      const cuuint32_t box[5] = {(cuuint32_t)(cfg.stg_rowb / 2), SK_HT_W, 4, 1, 1};
      CUresult r = encode(&tmO.m[ph], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, oswz,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("ofsv_conv_halo: cuTensorMapEncodeTiled(out %d) failed with %d", ph, (int)r); return OFSV_ECUDA; }
    }
  } else if (cfg.epi_mode == SK_EPI_TMA_SHUF) {
    // state tensor [N][Dy][Hy][Wy][8] fp32 viewed as [N][Dy][Hy][Wy/4][32 floats]
    // ... or, H-fastest, [N][Dy][Wy][Hy][8] viewed as [N][Dy][Wy][Hy/4][32 floats]
    const bool hf = d->out_shuffle_hfast != 0;
    const cuuint64_t row = (cuuint64_t)(hf ? d->Hy : d->Wy) * 32, plane = row * (hf ? d->Wy : d->Hy);
    const cuuint64_t gdim[5] = {32, (cuuint64_t)((hf ? d->Hy : d->Wy) / 4), (cuuint64_t)(hf ? d->Wy : d->Hy), (cuuint64_t)d->Dy, (cuuint64_t)d->N};
    const cuuint64_t gstr[4] = {128, row, plane, plane * d->Dy};
    const cuuint32_t box[5] = {32, (cuuint32_t)(hf ? 2 : 4), (cuuint32_t)(hf ? 16 : 8), 1, 1};
    CUresult r = encode(&tmO.m[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, y, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("ofsv_conv_halo: cuTensorMapEncodeTiled(state) failed with %d", (int)r); return OFSV_ECUDA; }
  }
  const int grid = (int)(total < sms ? total : sms);
  cudaStream_t st = (cudaStream_t)stream;
  if (KC == 32) return sk_launch<32>(P, tmA, tmB, tmO, bias, prelu, residual, y, grid, cfg.smem, st);
  return sk_launch<16>(P, tmA, tmB, tmO, bias, prelu, residual, y, grid, cfg.smem, st);
}
