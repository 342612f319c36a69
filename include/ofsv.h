/* ofsv.h — C ABI of libofsv.so, the B200 (sm_100a) implementation of the OpticalFlowSciVis hot path.
 *
 * This is the drop-in boundary.  The reference is pure Python; its only native FFI on this path
 * is the un-vendored pybind module `correlation_cuda` (UPFlow/model/correlation_package/correlation.py:26-27,42-43),
 * everything else bottoms out in ATen calls (F.grid_sample, F.interpolate, nn.Conv*, nn.ConvTranspose*, nn.PReLU).
 * Each entry point below names the reference call site it replaces.
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is DEVICE memory owned by the caller (PyTorch allocates it); the
 *     library never allocates or frees caller-visible memory;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream); work is enqueued, not
 *     synchronised; the library is stateless apart from a thread-local error string;
 *   - every function returns 0 on success, a negative OFSV_E* code otherwise (nothing is launched on a
 *     validation error); `ofsv_last_error()` returns the message for the calling thread;
 *   - fp32 tensors are contiguous NCHW / NCDHW exactly as the reference passes them; the conv engine uses
 *     channels-last activations (NHWC / NDHWC) that only ever live inside the library's callers.
 */
#ifndef OFSV_H_
#define OFSV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFSV_OK 0
#define OFSV_EINVAL (-1)  /* bad shape / null pointer / misalignment */
#define OFSV_ECUDA (-2)   /* CUDA runtime or driver error (message has the cudaError string) */
#define OFSV_ENOSUP (-3)  /* configuration outside what the kernels implement */

/* Which ATen build the fp32 warp arithmetic reproduces bit-for-bit:
 *   OFSV_REF_CPU  : flow / float((S-1)/2) true division, trilinear sum without FMA (ATen CPU kernels)
 *   OFSV_REF_CUDA : flow * float(1.0/((S-1)/2)) (ATen div_true_kernel_cuda fast path), FMA-contracted sums */
#define OFSV_REF_CPU 0
#define OFSV_REF_CUDA 1

const char* ofsv_version(void);
const char* ofsv_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t ofsv_launch_count(void);

/* ---- a1: warp(tenInput, tenFlow), 2-D — Flow-2D/model/warplayer.py:7-26 (F.grid_sample bilinear/border/align_corners).
 * src,out (N,C,H,W); flow (N,2,H,W); lin_x = torch.linspace(-1,1,W), lin_y = torch.linspace(-1,1,H) (device). */
int ofsv_warp2d_f32(const float* src, const float* flow, const float* lin_x, const float* lin_y, float* out,
                    int N, int C, int H, int W, int ref_mode, void* stream);

/* ---- a2: warp, 3-D — Flow-3D/model/warplayer.py:9-41.  Axis-rotating:
 * src,out (N,C,D,H,W); flow (N,3,D,H,W); lin_h/lin_d/lin_w = torch.linspace(-1,1,H|D|W).
 * out[n,c,d,h,w] = trilinear(src[n,c]; x = u(lin_h[h]+f0/((H-1)/2), W), y = u(lin_d[d]+f1/((D-1)/2), H),
 *                                      z = u(lin_w[w]+f2/((W-1)/2), D)),  u(g,S) = clip(((g+1)/2)*(S-1), 0, S-1). */
int ofsv_warp3d_f32(const float* src, const float* flow, const float* lin_h, const float* lin_d, const float* lin_w,
                    float* out, int N, int C, int D, int H, int W, int ref_mode, void* stream);

/* Same contract as ofsv_warp3d_f32, always on the global-gather kernel (ofsv_warp3d_f32 picks the TMA slab kernel of
 * csrc/warp3d_slab.cu on cubic volumes with S % 32 == 0; the two are bit-identical — this entry exists so that tests and
 * benchmarks can run both). */
int ofsv_warp3d_gather_f32(const float* src, const float* flow, const float* lin_h, const float* lin_d, const float* lin_w,
                           float* out, int N, int C, int D, int H, int W, int ref_mode, void* stream);

/* ---- backward of a1 / a2: what autograd runs under warp() in the reference's training step (Flow-2D/model/RIFE.py:80-336,
 * Flow-3D/model/RIFE.py:81-275 through Flow-{2D,3D}/model/warplayer.py:26 / :37) = ATen grid_sampler_{2,3}d_backward (bilinear,
 * border, align_corners=True, incl. the zero gradient of clipped coordinates) followed by the backward of
 * `flow / ((S-1)/2)`.  gout has the shape of out.  gsrc (shape of src; zero-filled by the call, then accumulated with
 * red.global.add) and gflow (shape of flow) may each be NULL = not wanted; src may be NULL when gflow is NULL. */
int ofsv_warp2d_bwd_f32(const float* src, const float* flow, const float* gout, const float* lin_x, const float* lin_y,
                        float* gsrc, float* gflow, int N, int C, int H, int W, int ref_mode, void* stream);
int ofsv_warp3d_bwd_f32(const float* src, const float* flow, const float* gout, const float* lin_h, const float* lin_d,
                        const float* lin_w, float* gsrc, float* gflow, int N, int C, int D, int H, int W, int ref_mode,
                        void* stream);

/* ---- a6 (+a1/a2 fused): sigmoid(mask) ; warp(img0, flow[:, :nd]) ; warp(img1, flow[:, nd:2nd]) ; blend.
 * Flow-2D/model/IFNet.py:189-192,240 ; Flow-3D/model/IFNet.py:186-191,242.
 * img0,img1,mask_logit (N,1,·); flow (N,2*nd,·).  Any of warped0/warped1/merged/mask_sig may be NULL (not written). */
int ofsv_warp_blend_2d_f32(const float* img0, const float* img1, const float* flow, const float* mask_logit,
                           const float* lin_x, const float* lin_y, float* warped0, float* warped1, float* merged,
                           float* mask_sig, int N, int H, int W, int ref_mode, void* stream);
int ofsv_warp_blend_3d_f32(const float* img0, const float* img1, const float* flow, const float* mask_logit,
                           const float* lin_h, const float* lin_d, const float* lin_w, float* warped0, float* warped1,
                           float* merged, float* mask_sig, int N, int D, int H, int W, int ref_mode, void* stream);
/* plain blend of already-warped frames: merged = w0*sigmoid(m) + w1*(1-sigmoid(m)) */
int ofsv_blend_f32(const float* w0, const float* w1, const float* mask_logit, float* merged, int64_t n, void* stream);

/* ---- a8: correlation_cuda.forward / .backward — UPFlow/model/correlation_package/correlation.py:26-27,42-43,
 * call sites UPFlow/model/upflow.py:649,652 with (pad 4, kernel 1, max_disp 4, stride1 1, stride2 1, mult 1) only.
 * f1,f2 (B,C,H,W); out (B,81,H,W) written at out + b*out_batch_stride (elements) so the caller can point it into the
 * channel-concatenated estimator input (upflow.py:657).  apply_leaky fuses LeakyReLU(leaky_slope) (upflow.py:655-656).
 * work: optional caller-owned scratch of ofsv_corr81_fwd_splits(B,C,H,W) * B*81*H*W floats (NULL = none): with it the channels
 * of the coarse pyramid levels are split over CTAs (partial sums, added in a fixed order by a second launch — deterministic). */
int ofsv_corr81_fwd_splits(int B, int C, int H, int W);
int ofsv_corr81_fwd_f32(const float* f1, const float* f2, float* out, int B, int C, int H, int W, float leaky_slope,
                        int apply_leaky, int64_t out_batch_stride, float* work, void* stream);
int ofsv_corr81_bwd_f32(const float* f1, const float* f2, const float* gout, float* g1, float* g2, int B, int C, int H,
                        int W, void* stream);

/* ---- a10: upsample2d_flow_as(inputs, target_as, "bilinear", if_rate) — UPFlow/model/pwc_modules.py:77-90. */
int ofsv_upsample_flow_ac_f32(const float* in, float* out, int B, int h_in, int w_in, int h_out, int w_out, int if_rate,
                              void* stream);

/* ---- a11: WarpingLayer_no_div.forward — UPFlow/model/pwc_modules.py:184-207 (zeros padding, default
 * align_corners=False, times the grid_sample(ones) >= 1 validity mask). */
int ofsv_warping_no_div_f32(const float* src, const float* flow, float* out, int B, int C, int H, int W, int ref_mode,
                            void* stream);

/* tools.torch_warp(x, flo) — UPFlow/utils/tools.py:1317-1361: the sampling of WarpingLayer_no_div WITHOUT the validity mask (used by
 * the occlusion check, tools.py:592-630, and by sgu_model, upflow.py:86). */
int ofsv_torch_warp_f32(const float* src, const float* flow, float* out, int B, int C, int H, int W, int ref_mode, void* stream);

/* ---- f.2: producer of the cost-volume inputs of one pyramid level — UPFlow/model/upflow.py:621-640:
 *   out_warp  = normalize(WarpingLayer_no_div(f_src, flow))      (flow NULL = level 0: no warp)
 *   out_plain = normalize(f_plain)
 * with network_tools.normalize_features(normalize = center = True, moments_across_channels = moments_across_images = False)
 * (upflow.py:95-138; the configuration of scripts/simple_train.py:321-329 and test.py:116-118): per (sample, channel) plane
 * (f - mean) / sqrt(var_unbiased + 1e-16).  All tensors (B,C,H,W) fp32, flow (B,2,H,W).  One launch; planes are staged in shared
 * memory (H*W*4 <= 200 KB, else OFSV_ENOSUP). */
int ofsv_feature_norm_pair_f32(const float* f_plain, const float* f_src, const float* flow, float* out_plain, float* out_warp, int B,
                               int C, int H, int W, int ref_mode, void* stream);

/* ---- backward of a10 / a11 (autograd of UPFlow's training step, BASELINE cfg 5, runs through both).
 * ofsv_upsample_flow_ac_bwd_f32: gin (B,2,h_in,w_in) = upsample_bilinear2d_backward(gout (B,2,h_out,w_out)) with the
 *   (w/w_, h/h_) factors of pwc_modules.py:83-88; gin is zero-filled by the call.
 * ofsv_warping_no_div_bwd_f32: grid_sampler_2d_backward (zeros padding, align_corners=False) times the constant validity
 *   mask, chained with the backward of `2*v/max(S-1,1) - 1` (pwc_modules.py:198-199).  gsrc (zero-filled by the call) and
 *   gflow may each be NULL; src may be NULL when gflow is NULL. */
int ofsv_upsample_flow_ac_bwd_f32(const float* gout, float* gin, int B, int h_in, int w_in, int h_out, int w_out, int if_rate,
                                  void* stream);
int ofsv_warping_no_div_bwd_f32(const float* src, const float* flow, const float* gout, float* gsrc, float* gflow, int B,
                                int C, int H, int W, int ref_mode, void* stream);

/* ---- training tier: the optimizer step of Model.update — torch.optim.AdamW(lr=1e-6, weight_decay=1e-3),
 * Flow-2D/model/RIFE.py:26,317 ; Flow-3D/model/RIFE.py:29,259 — for EVERY parameter tensor in one launch.
 * `tensors`: device array of ntensors records {float* p; const float* g; float* m; float* v; int64_t n} (40 bytes each);
 * `chunks`: device array of nchunks (tensor index, chunk index) int32 pairs covering every tensor in 4096-element chunks.
 * step >= 1 is the step count AFTER this update (bias corrections 1 - beta^step); grad_scale multiplies the gradients
 * (1/world after a SUM allreduce). */
#define OFSV_ADAMW_CHUNK 4096
int ofsv_adamw_step_f32(const void* tensors, const void* chunks, int ntensors, int nchunks, float lr, float beta1, float beta2,
                        float eps, float weight_decay, int step, float grad_scale, void* stream);

/* =====================================================================================================
 * IFBlock / IFNet engine (a3, a4, a5) — Flow-2D/model/IFNet.py:16-27,34-122,144-276 ; Flow-3D/model/IFNet.py.
 * Activations are channels-last: [N][D][H][W][Cs] with Cs the channel count rounded up to a multiple of 16
 * (2-D tensors use D = 1).  `act_dtype`: OFSV_F32 (exact-arithmetic validation path) or OFSV_BF16 (tensor cores).
 * ===================================================================================================== */
#define OFSV_F32 0
#define OFSV_BF16 1

/* IFBlock input builder: F.interpolate(x, 1/scale) ‖ F.interpolate(flow, 1/scale)*(1/scale) ‖ torch.cat
 * (IFNet.py:84-93 2-D, :82-90 3-D; the cat of IFNet.py:174 / :166).  Sources are the fp32 NC(D)HW tensors of the
 * reference (any of warped0/warped1/mask/flow may be NULL for block0: then only img0,img1 are packed).
 * nd = 2|3, scale ∈ {1,2,4}; dst [N][D/s][H/s][W/s][Cs], channel order img0,img1,warped0,warped1,mask,flow[0..2nd). */
int ofsv_pack_block_input(const float* img0, const float* img1, const float* warped0, const float* warped1,
                          const float* mask, const float* flow, void* dst, int act_dtype, int nd, int N, int D, int H,
                          int W, int scale, int Cs, int s2d, void* stream);
/* s2d = 1: dst is the shifted space-to-depth tensor [N][Dn/2+1][Hn/2+1][Wn/2+1][2^nd][Cs] of the resized grid
 * (Dn,Hn,Wn) = (D,H,W)/scale (see ofsv_conv_desc.out_s2d); only interior sub-cells are written. */

/* Layout changes between the reference's fp32 NC(D)HW tensors and the engine's channels-last bf16 activations, P = D*H*W pixels:
 *   ofsv_pack_nhwc_bf16 : dst [N][P][Cs] bf16 = the channel concatenation of 1..8 sources [N][channels[i]][P] fp32 (host arrays of
 *                         device pointers / channel counts), zero-padded to Cs — torch.cat + pad + cast in one pass, e.g. the
 *                         estimator input torch.cat([corr, x_1x1, flow], 1) of UPFlow/model/upflow.py:657;
 *   ofsv_unpack_nhwc_f32: dst [N][C][P] fp32 = the first C channels of src [N][P][Cs] bf16. */
int ofsv_pack_nhwc_bf16(const float* const* srcs, const int* channels, int nsrc, void* dst, int N, int64_t P, int Cs, void* stream);
int ofsv_unpack_nhwc_f32(const void* src, float* dst, int N, int64_t P, int Cs, int C, void* stream);

/* Data edge: uint8 volume -> fp32, dst[i] = (float)src[i] / div with IEEE division (div = 255 reproduces
 * the reference loaders' `/ 255.`: Datasets/read_data.py, Flow-3D/load_datasets.py), so only bytes cross PCIe. */
int ofsv_u8_to_f32(const uint8_t* src, float* dst, int64_t n, float div, void* stream);

/* ... and back: dst[i] = (uint8)clamp(src[i] * mul, 0, 255), truncated toward zero — the reference's export
 * `(img * 255).byte()` (Flow-3D/inference_img.py:105) done before the download. */
int ofsv_f32_to_u8(const float* src, uint8_t* dst, int64_t n, float mul, void* stream);

/* Evaluation metrics on the device, float64 like the numpy reference, reduced in a fixed order.
 * `partials` is caller-provided scratch of N * OFSV_METRIC_BLOCKS doubles; `out` receives N doubles.
 *   ofsv_sq_err_f64: out[n] = sum_i (((a[n][i] - b[n][i]) * scale)^2) over `count` elements per sample, all in fp64:
 *     MSE = out/count feeds `calculate_psnr` (error.py:27-34, scale = 255 for [0,1] data) and
 *     `-10*log10(mean((gt-pred)^2))` (Flow-3D/train.py:385-388, scale = 1).
 *   ofsv_ssim2d_f64: out[n] = mean SSIM map of the N image pairs x[n], y[n] ([H][W] fp32 planes) with the 11x11 Gaussian
 *     window (sigma 1.5) over the 'valid' region and C1 = (0.01 L)^2, C2 = (0.03 L)^2, L = data_range (error.py:36-56). */
#define OFSV_METRIC_BLOCKS 64
int ofsv_sq_err_f64(const float* a, const float* b, double* partials, double* out, int N, int64_t count, float scale,
                    void* stream);
int ofsv_ssim2d_f64(const float* x, const float* y, double* partials, double* out, int N, int H, int W, double data_range,
                    void* stream);

#define OFSV_MAX_TAPS 64
/* One convolution layer in "tap" form.  For every phase ph, every virtual output position o = (oz,oy,ox) in
 * [0,Do)x[0,Ho)x[0,Wo):
 *   y[n, o*out_stride + parity(ph), co] = act( bias[co] + sum_t sum_ci x[n, o*in_stride + tap_off[ph*ntaps+t], ci]
 *                                                                   * w[ph][t][ci][co] ) (+ residual)
 * with x read as zero outside [0,Di)x[0,Hi)x[0,Wi).
 *   Conv(k,s,p):            nphase = 1, ntaps = k^nd, tap_off = k_idx - p, in_stride = s, out_stride = 1.
 *   ConvTranspose(4,2,1):   nphase = 2^nd (one per output parity; parity bits of ph = (z,y,x), x lowest), ntaps = 2^nd,
 *                           in_stride = 1, out_stride = 2; per axis parity 0 uses kernel taps {1 @ 0, 3 @ -1},
 *                           parity 1 uses {2 @ 0, 0 @ +1}. */
typedef struct ofsv_conv_desc {
  int32_t nd;                         /* 2 or 3 (2-D uses D = 1 and tap_off[.][0] = 0) */
  int32_t N, Di, Hi, Wi, Cin_s;       /* input  [N][Di][Hi][Wi][Cin_s], Cin_s % 16 == 0 */
  int32_t Do, Ho, Wo;                 /* virtual output grid (per phase) */
  int32_t Dy, Hy, Wy, Cout_s;         /* physical output tensor [N][Dy][Hy][Wy][Cout_s], Cout_s % 8 == 0 */
  int32_t Cout_w;                     /* padded Cout of w / bias / prelu, Cout_w % 16 == 0, Cout_s <= Cout_w */
  int32_t in_stride, out_stride;
  int32_t nphase, ntaps;              /* nphase * ntaps <= OFSV_MAX_TAPS */
  int8_t tap_off[OFSV_MAX_TAPS][4];   /* (z,y,x,unused) input offset of tap [ph*ntaps + t] */
  int32_t has_prelu, has_residual;
  int32_t in_dtype, out_dtype;        /* OFSV_F32 | OFSV_BF16 */
  int32_t out_shuffle;                /* 0, or 8: depth-to-space heads (ofsv_conv_halo only) — nphase = 1, the Cout_w = 2^nd * 8
                                       * columns are [output parity (z,y,x)][8 channels], row o is written to the 2^nd positions
                                       * 2*o + parity of a [N][2Do][2Ho][2Wo][8] tensor (ConvTranspose(4,2,1) with all parities
                                       * evaluated as ONE 3^nd-tap conv whose weights are zero where a parity does not use a tap) */
  int32_t out_s2d;                    /* 1: y is written in the SHIFTED SPACE-TO-DEPTH layout (ofsv_conv_halo only, nphase = 1,
                                       * out_stride = 1, no residual): logical output (z,y,x) of the [Dy][Hy][Wy] grid goes to cell
                                       * ((i+1)>>1) / sub-cell ((i+1)&1) per axis of a [N][Dy/2+1][Hy/2+1][Wy/2+1][2^nd][Cout_s]
                                       * tensor whose border sub-cells the caller keeps zero.  A Conv(k<=4, s=2, p=1) over the
                                       * logical tensor is then a stride-1 conv with tap offsets in {0,1}^nd over 2^nd*Cout_s
                                       * channels: kernel index k = 2*offset + sub-cell. */
  int32_t out_shuffle_hfast;          /* with out_shuffle (fp32 output): write the depth-to-space tensor H-FASTEST, [N][2Do][2Wo][2Ho][8]
                                       * (= OFSV_STATE_DWH8 of ofsv_block_stage_3d; the residual state is read in the same layout) */
} ofsv_conv_desc;

/* SIMT (CUDA-core, fp32 accumulate) engine: exact-order validation path and small-channel layers.
 * w: fp32 [nphase][ntaps][Cin_s][Cout_w]; bias, prelu: fp32 [Cout_w]; residual: same layout/dtype as y (or NULL). */
int ofsv_conv_simt(const ofsv_conv_desc* d, const void* x, const float* w, const float* bias, const float* prelu,
                   const void* residual, void* y, void* stream);

/* tcgen05/TMEM implicit-GEMM engine (bf16 operands, fp32 accumulate in tensor memory).
 * w: bf16 [nphase][ntaps][Cin_s/KC][Cout_w][KC] (K-major B tiles, KC = largest of 64/32/16 dividing Cin_s). */
int ofsv_conv_tc(const ofsv_conv_desc* d, const void* x, const void* w, const float* bias, const float* prelu,
                 const void* residual, void* y, void* stream);

/* Same contract as ofsv_conv_tc for the stride-1 layers (3^d convs, ConvTranspose phases, depth-to-space heads, the
 * space-to-depth conv0 layers: every tap offset in {-1,0,1}): each input halo plane is loaded into shared memory once per
 * super-tile and every tap is a shifted UMMA descriptor; output slices that read the same plane are STACKED along N of one
 * tcgen05.mma (csrc/conv_stack.cu); persistent CTAs, double-buffered TMEM accumulators, TMA-store epilogue.
 * `w` must be in the layout ofsv_conv_halo_weight_layout(d) names.  Returns OFSV_ENOSUP (nothing launched) for other layers. */
int ofsv_conv_halo(const ofsv_conv_desc* d, const void* x, const void* w, const float* bias, const float* prelu,
                   const void* residual, void* y, void* stream);

/* Weight re-packing (SURVEY.md section 8b `ofsv_pack_*`): ONE launch per layer and parameter version.
 *   w_tap : fp32 [nphase*ntaps][Cin_s][Cout_w] (the layout ofsv_conv_simt consumes)
 *   w_out : bf16, nphase*ntaps*Cin_s*Cout_w elements, K-major [Cout_w][KC] blocks in
 *     OFSV_WL_TAP   : [nphase][ntaps][Cin_s/KC] order, KC = largest of 64/32/16 dividing Cin_s (ofsv_conv_tc, the plane-ring kernel)
 *     OFSV_WL_STACK : [pass][in-plane tap offset (dy,dx)][Cin_s/KC][slot = (phase asc, dz desc)] order, KC = 32 or 16: the
 *                     slots of one in-plane offset are consecutive rows, so ONE MMA reads the weights of every output slice
 *                     that shares an input plane (csrc/conv_stack.cu).
 * Only nd, nphase, ntaps, tap_off, Cin_s, Cout_w of the descriptor are read (no shapes): a layer is packed once.
 * ofsv_conv_halo_weight_layout: which of the two ofsv_conv_halo expects for this descriptor (>= 0), or a negative error. */
#define OFSV_WL_TAP 0
#define OFSV_WL_STACK 1
int ofsv_conv_pack_weights(const ofsv_conv_desc* d, const float* w_tap, void* w_out, int layout, void* stream);
int ofsv_conv_halo_weight_layout(const ofsv_conv_desc* d);

/* Batched re-packing for the training step (all layers of a block change every step): ofsv_conv_pack_record fills one HOST record per
 * (layer, layout) — same arguments and result as ofsv_conv_pack_weights —, the caller uploads the array of records once and
 * ofsv_conv_pack_weights_batched re-packs all of them in ONE launch whenever the fp32 tap forms changed. */
typedef struct ofsv_pack_rec {
  const float* w_tap;
  void* w_out;
  int32_t nblocks, Cin_s, Cout_w, KC;
  uint16_t blk[OFSV_MAX_TAPS * 8];
} ofsv_pack_rec;
int ofsv_conv_pack_record(const ofsv_conv_desc* d, int layout, const float* w_tap, void* w_out, ofsv_pack_rec* rec);
int ofsv_conv_pack_weights_batched(const ofsv_pack_rec* recs_dev, int nrec, void* stream);

/* One-line description of the launch configuration ofsv_conv_halo would pick for `d` (kernel, super-tile depth, ring depths,
 * epilogue mode, shared memory, modelled tensor-pipe fraction of the MMA list); host only, nothing is launched. */
int ofsv_conv_halo_describe(const ofsv_conv_desc* d, char* buf, int buflen);

/* Host-only self-check of the stacked kernel's MMA list for super-tile depth td (runs without a GPU): every (phase, tap,
 * output slice) term covered exactly once, first-touch flags consistent, runs contiguous.  Optionally returns the number of
 * MMAs-per-K-step and the modelled tensor cycles of one super-tile. */
int ofsv_conv_stack_selfcheck(const ofsv_conv_desc* d, int td, int* nops_out, double* mma_cycles_out);

/* Process-wide tuning / A-B switches between EQUIVALENT code paths (every setting computes the same results; there are no
 * environment variables in the launch path).  Keys: "stack_epilogue" (-1 auto, 0 = per-thread stores instead of the TMA-store
 * epilogue), "stack_td" (0 auto, 1|2|4 = super-tile depth when feasible), "warp_slab" (1 default, 0 = gather kernel only), "wgrad_brick" (-1 default: the
 * brick-window weight-gradient kernel where it is the fastest, 1 = wherever its windows fit, 0 = per-tap / tap-group kernels only), "tc_pair" (-1 default: ofsv_conv_tc gives a CTA two output tiles that share every weight tile
 * when the layer has many tiles and a long K loop; 0 never, 1 whenever there are two tiles), "tc_stages" (0 default: ofsv_conv_tc's
 * pipeline depth such that two CTAs fit an SM; 2..4 forces it). */
int ofsv_set_tuning(const char* key, int value);

/* ---- training tier: backward of conv()/deconv() + PReLU — what `loss_G.backward()` runs under every layer of
 * IFBlock (Flow-3D/model/RIFE.py:255-259, Flow-2D/model/RIFE.py:315-317 through Flow-{2D,3D}/model/IFNet.py:16-27).
 * The input gradient of a tap-form layer is another tap-form layer (conv <-> transposed conv) and runs on ofsv_conv_halo /
 * ofsv_conv_tc; these two entries are the rest:
 *
 * ofsv_prelu_bias_bwd_bf16: gy, y, gpre bf16 [P][Cs] (y = the layer's output AFTER PReLU; gpre may alias gy); slope fp32 [Cs] or
 *   NULL (layer without activation: gpre = gy).  gpre = gy * (y > 0 ? 1 : slope); dbias[c] = sum_P gpre; dslope[c] = sum_P gy * pre
 *   over pre < 0, pre = y / slope (requires slope > 0).  `work`: 2 * Cs * ofsv_prelu_bias_bwd_blocks() floats of scratch; the
 *   reduction order is fixed (deterministic).
 * ofsv_conv_wgrad_bf16: dw fp32 [nphase*ntaps][Cin_s][Cout_w] (the tap form ofsv_conv_simt consumes),
 *     dw[ph*ntaps+t][ci][co] = sum_{n,o} x[n, o*in_stride + tap_off[ph*ntaps+t]][ci] * gy[n, o*out_stride + parity(ph)][co],
 *   x bf16 [N][Di][Hi][Wi][Cin_s] (the layer's input), gy bf16 [N][Dy][Hy][Wy][gy_cs] (gradient w.r.t. the pre-activation,
 *   gy_cs >= Cout_w).  Only the geometry fields of the descriptor are read.  bf16 mma.sync GEMM over the output positions, split
 *   along K over ofsv_conv_wgrad_splits(d) CTAs per tile; `work` = splits * nphase*ntaps*Cin_s*Cout_w floats (may be NULL when
 *   splits == 1), summed in a fixed order. */
/* Per-step refresh of the fp32 tap-form weights (and padded bias / slope vectors) of many layers from the reference's parameter
 * tensors in ONE launch (the optimizer changes every parameter every step; opticalflowscivis_b200/train.py).  `recs_dev` is a DEVICE
 * array of records:
 *   kind 0: dst[(t*Cin_s + ci0 + i)*Cout_w + co0 + o] = kidx[t] >= 0 ? src[(a*B + b)*K + kidx[t]] : 0 for t < T, with the parameter
 *           viewed [A][B][K] (K = kernel volume) and (i, o) = (a, b) (swap = 0: ConvTranspose weights forward, Conv weights as the
 *           matrices of an input-gradient layer) or (b, a) (swap = 1);
 *   kind 1: dst[i] = src[i], i < n. */
typedef struct ofsv_refresh_rec {
  const float* src;
  float* dst;
  int32_t kind, A, B, K, swap, T, Cin_s, Cout_w, ci0, co0, n, pad_;
  int16_t kidx[OFSV_MAX_TAPS];
} ofsv_refresh_rec;
int ofsv_conv_refresh_tapform(const ofsv_refresh_rec* recs_dev, int nrec, void* stream);
int ofsv_prelu_bias_bwd_blocks(void);
int ofsv_prelu_bias_bwd_bf16(const void* gy, const void* y, const float* slope, void* gpre, float* dbias, float* dslope, float* work,
                             int64_t P, int Cs, void* stream);
int ofsv_conv_wgrad_splits(const ofsv_conv_desc* d);
int ofsv_conv_wgrad_bf16(const ofsv_conv_desc* d, const void* x, const void* gy, int gy_cs, float* dw, float* work, void* stream);

/* IFBlock output stage: flow/mask deltas at block resolution -> full resolution (IFNet.py:115-116 / :118-119,
 * F.interpolate(.., scale) and flow*scale), then flow += flow_d ; mask += mask_d (IFNet.py:177-178 / :169-170).
 * head [N][D/s][H/s][W/s][Cs] channels-last fp32 (channels 0..2nd-1 flow, 2nd mask); flow_prev/mask_prev may be NULL
 * (block0).  flow_out (N,2nd,·), mask_out (N,1,·) fp32 NC(D)HW. */
int ofsv_head_upsample_add(const float* head, int Cs, const float* flow_prev, const float* mask_prev, float* flow_out,
                           float* mask_out, int nd, int N, int D, int H, int W, int scale, void* stream);

/* Backward of ofsv_head_upsample_add and ofsv_pack_block_input for the training step (autograd of
 * IFNet.py:84-93,115-119 / :82-90,118-119 under Model.update).
 *   ofsv_head_upsample_add_bwd: ghead [N][D/s][H/s][W/s][8] fp32 = adjoint of the trilinear / bilinear up-sampling applied to
 *     (gflow (N,2nd,.) * s, gmask (N,1,.)); the gradients w.r.t. flow_prev / mask_prev are gflow / gmask themselves.  Gather form,
 *     deterministic.
 *   ofsv_pack_block_input_bwd: gx [N][D/s][H/s][W/s][16] bf16 (plain layout, Cs = 16) -> gradients of warped0, warped1, mask (N,1,.)
 *     and flow (N,2nd,.) fp32; img0 / img1 take none. */
int ofsv_head_upsample_add_bwd(const float* gflow, const float* gmask, float* ghead, int nd, int N, int D, int H, int W, int scale,
                               void* stream);
int ofsv_pack_block_input_bwd(const void* gx, float* gw0, float* gw1, float* gmask, float* gflow, int nd, int N, int D, int H, int W,
                              int scale, void* stream);

/* Fused 3-D IFBlock output stage: ofsv_head_upsample_add + ofsv_warp_blend_3d_f32 (+ the next block's
 * ofsv_pack_block_input) in one pass over the full-resolution voxels — Flow-3D/model/IFNet.py:118-119 (resize, *scale),
 * :169-170 (flow/mask accumulate), :186-191 (sigmoid, warp x2), :242 (blend), :82-90,166 (next block's resized concat).
 * It works on the CHANNELS-LAST flow/mask state fm[N][D][H][W][8] fp32 = (flow 0..5, mask logit, 0): one 32 B sector
 * per voxel, read and written with coalesced 16 B accesses; it is also the layout the depth-to-space head conv produces.
 * head [N][D/sh][H/sh][W/sh][8] fp32; fm_prev NULL for block0; merged / mask_sig (N,1,D,H,W) optional;
 * scale_next in {0: no packed output, 1, 2}: pack_out = the next block's conv0 input, bf16 [N][D/sn][H/sn][W/sn][16], or with
 * pack_s2d = 1 its shifted space-to-depth form (ofsv_conv_desc.out_s2d) so that conv0 (k=4,s=2,p=1) runs as a stride-1
 * conv on ofsv_conv_halo.
 *
 * state_layout: how head / fm_prev / fm_out are stored.
 *   OFSV_STATE_DHW8 : [N][D][H][W][8]  (W fastest; csrc/block_stage.cu: state tiles staged through shared memory)
 *   OFSV_STATE_DWH8 : [N][D][W][H][8]  (H fastest; csrc/block_stage_hfast.cu).  The reference warp rotates axes, so its gathers are
 *                     only coalesced with lanes along H; with H as the state's fastest spatial axis one thread owns one voxel for
 *                     the whole stage and reads / writes its 32 B of state with single 256-bit coalesced accesses.  The depth-to-
 *                     space head conv writes this layout when ofsv_conv_desc.out_shuffle_hfast is set. */
#define OFSV_STATE_DHW8 0
#define OFSV_STATE_DWH8 1
int ofsv_block_stage_3d(const float* head, const float* fm_prev, const float* img0, const float* img1, const float* lin_h,
                        const float* lin_d, const float* lin_w, float* fm_out, float* merged, float* mask_sig,
                        void* pack_out, int N, int D, int H, int W, int scale_head, int scale_next, int pack_s2d,
                        int ref_mode, int state_layout, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OFSV_H_ */
