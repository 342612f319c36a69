// Evaluation metrics of the data edge on the device, so that interpolated volumes need not leave the GPU to
// be scored:
//   * squared-error sums for PSNR            error.py:27-34 (`calculate_psnr`), Flow-3D/train.py:385-388 (`-10 log10(mean(d*d))`)
//   * SSIM of 2-D images                     error.py:36-56 (`ssim`: 11x11 Gaussian window, sigma 1.5, 'valid' region, float64)
// Both are computed in float64 like the numpy reference and reduced in a FIXED order (per-block partial sums in a caller-provided
// workspace, then one warp per sample), so results are bit-reproducible run to run.
#include "ofsv_common.cuh"

namespace ofsv {

constexpr int MET_BLOCKS = 64;     // partial sums per sample (OFSV_METRIC_BLOCKS in ofsv.h)
constexpr int MET_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < MET_THREADS / 32; ++w) t += s_red[w];
  return t;   // valid in thread 0
}

// partial[n][b] = sum over the b-th slice of sample n of ((a - b) * scale)^2, everything in float64 (img.astype(np.float64))
__global__ void __launch_bounds__(MET_THREADS) sq_err_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                     double* __restrict__ partial, int64_t count, float scale) {
  __shared__ double s_red[MET_THREADS / 32];
  const int n = blockIdx.y;
  const float* pa = a + (int64_t)n * count;
  const float* pb = b + (int64_t)n * count;
  const int64_t per = (count + MET_BLOCKS - 1) / MET_BLOCKS;
  const int64_t beg = (int64_t)blockIdx.x * per, end = min(count, beg + per);
  double acc = 0.0;
  for (int64_t i = beg + threadIdx.x; i < end; i += MET_THREADS) {
    const double d = ((double)__ldg(pa + i) - (double)__ldg(pb + i)) * (double)scale;
    acc += d * d;
  }
  const double t = block_sum(acc, s_red);
  if (threadIdx.x == 0) partial[(int64_t)n * MET_BLOCKS + blockIdx.x] = t;
}

// out[n] = (sum_b partial[n][b]) * mul, summed in index order
__global__ void final_sum_kernel(const double* __restrict__ partial, double* __restrict__ out, int N, double mul) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double t = 0.0;
  for (int b = 0; b < MET_BLOCKS; ++b) t += partial[(int64_t)n * MET_BLOCKS + b];
  out[n] = t * mul;
}

struct SsimWin { double w[11]; };

// one thread per pixel of the 'valid' region; float64 throughout (img.astype(np.float64), error.py:40-41)
__global__ void __launch_bounds__(MET_THREADS) ssim2d_partial_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                     double* __restrict__ partial, int H, int W, double C1,
                                                                     double C2, SsimWin g) {
  __shared__ double s_red[MET_THREADS / 32];
  const int n = blockIdx.y;
  const int Hv = H - 10, Wv = W - 10;
  const int64_t npix = (int64_t)Hv * Wv;
  const float* px = x + (int64_t)n * H * W;
  const float* py = y + (int64_t)n * H * W;
  const int64_t per = (npix + MET_BLOCKS - 1) / MET_BLOCKS;
  const int64_t beg = (int64_t)blockIdx.x * per, end = min(npix, beg + per);
  double acc = 0.0;
  for (int64_t i = beg + threadIdx.x; i < end; i += MET_THREADS) {
    const int oy = (int)(i / Wv), ox = (int)(i - (int64_t)oy * Wv);
    double mu1 = 0.0, mu2 = 0.0, s11 = 0.0, s22 = 0.0, s12 = 0.0;
    for (int dy = 0; dy < 11; ++dy) {
      const float* rx = px + (int64_t)(oy + dy) * W + ox;
      const float* ry = py + (int64_t)(oy + dy) * W + ox;
#pragma unroll
      for (int dx = 0; dx < 11; ++dx) {
        const double wgt = g.w[dy] * g.w[dx];          // window = outer(kernel, kernel) (error.py:43)
        const double a = (double)__ldg(rx + dx), b = (double)__ldg(ry + dx);
        mu1 += wgt * a; mu2 += wgt * b;
        s11 += wgt * (a * a); s22 += wgt * (b * b); s12 += wgt * (a * b);
      }
    }
    const double mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
    const double sigma1_sq = s11 - mu1_sq, sigma2_sq = s22 - mu2_sq, sigma12 = s12 - mu1_mu2;
    acc += ((2.0 * mu1_mu2 + C1) * (2.0 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2));
  }
  const double t = block_sum(acc, s_red);
  if (threadIdx.x == 0) partial[(int64_t)n * MET_BLOCKS + blockIdx.x] = t;
}

}  // namespace ofsv

using namespace ofsv;

extern "C" int ofsv_sq_err_f64(const float* a, const float* b, double* partials, double* out, int N, int64_t count, float scale,
                               void* stream) {
  OFSV_REQUIRE(N >= 0 && count >= 1, "ofsv_sq_err_f64: bad shape");
  OFSV_REQUIRE(N <= 65535, "ofsv_sq_err_f64: N exceeds grid.y");
  if (N == 0) return OFSV_OK;
  OFSV_REQUIRE(a && b && partials && out, "ofsv_sq_err_f64: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  sq_err_partial_kernel<<<dim3(MET_BLOCKS, (unsigned)N), MET_THREADS, 0, st>>>(a, b, partials, count, scale);
  int rc = check_launch("sq_err_partial_kernel");
  if (rc != OFSV_OK) return rc;
  final_sum_kernel<<<(unsigned)cdiv(N, 128), 128, 0, st>>>(partials, out, N, 1.0);
  return check_launch("final_sum_kernel");
}

extern "C" int ofsv_ssim2d_f64(const float* x, const float* y, double* partials, double* out, int N, int H, int W,
                               double data_range, void* stream) {
  OFSV_REQUIRE(N >= 0 && H >= 11 && W >= 11, "ofsv_ssim2d_f64: images must be at least 11x11 (the window of error.py:42)");
  OFSV_REQUIRE(N <= 65535, "ofsv_ssim2d_f64: N exceeds grid.y");
  OFSV_REQUIRE(data_range > 0.0, "ofsv_ssim2d_f64: data_range must be positive");
  if (N == 0) return OFSV_OK;
  OFSV_REQUIRE(x && y && partials && out, "ofsv_ssim2d_f64: null pointer");
  // cv2.getGaussianKernel(11, 1.5): exp(-(i - 5)^2 / (2 sigma^2)) normalised to sum 1, float64
  SsimWin g;
  double sum = 0.0;
  for (int i = 0; i < 11; ++i) { g.w[i] = exp(-((double)(i - 5) * (double)(i - 5)) / (2.0 * 1.5 * 1.5)); sum += g.w[i]; }
  for (int i = 0; i < 11; ++i) g.w[i] /= sum;
  const double C1 = (0.01 * data_range) * (0.01 * data_range), C2 = (0.03 * data_range) * (0.03 * data_range);
  cudaStream_t st = (cudaStream_t)stream;
  ssim2d_partial_kernel<<<dim3(MET_BLOCKS, (unsigned)N), MET_THREADS, 0, st>>>(x, y, partials, H, W, C1, C2, g);
  int rc = check_launch("ssim2d_partial_kernel");
  if (rc != OFSV_OK) return rc;
  final_sum_kernel<<<(unsigned)cdiv(N, 128), 128, 0, st>>>(partials, out, N, 1.0 / ((double)(H - 10) * (double)(W - 10)));
  return check_launch("final_sum_kernel");
}
