// Fused multi-tensor AdamW step (row f.1: the optimizer of Model.update — `AdamW(self.flownet.parameters(), lr=1e-6,
// weight_decay=1e-3)`, Flow-2D/model/RIFE.py:26,81-82,317 ; Flow-3D/model/RIFE.py:29,86-87,259).  The reference calls
// torch.optim.AdamW, i.e. per parameter tensor (amsgrad = False, maximize = False):
//     p   *= 1 - lr * wd
//     m    = m + (g - m) * (1 - beta1)                      (lerp_)
//     v    = v * beta2 + (1 - beta2) * g * g                (mul_, addcmul_)
//     p   -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// ONE launch updates every tensor of the model (≈150 tensors, 9-36 MB): a chunk table maps 4096-element chunks to (tensor,
// offset), so the kernel streams p, g, m, v once (16 B read + 12 B written per element) instead of ~10 launches per tensor.
#include "ofsv_common.cuh"

namespace ofsv {

constexpr int ADAMW_CHUNK = 4096;

struct AdamWTensor {
  float* p; const float* g; float* m; float* v;
  int64_t n;
};

__global__ void __launch_bounds__(256)
    adamw_kernel(const AdamWTensor* __restrict__ tensors, const int2* __restrict__ chunks, int nchunks, float lr, float beta1,
                 float beta2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale) {
  const float decay = 1.0f - lr * wd, w1 = 1.0f - beta1, w2 = 1.0f - beta2, step_size = lr / bc1;
  for (int ck = blockIdx.x; ck < nchunks; ck += gridDim.x) {
    const int2 c = chunks[ck];                                   // (tensor index, chunk index inside the tensor)
    const AdamWTensor t = tensors[c.x];
    const int64_t beg = (int64_t)c.y * ADAMW_CHUNK;
    const int64_t end = beg + ADAMW_CHUNK < t.n ? beg + ADAMW_CHUNK : t.n;
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const float g = t.g[i] * grad_scale;                       // grad_scale = 1 / world size after a SUM allreduce, else 1
      float p = t.p[i] * decay;
      float m = t.m[i], v = t.v[i];
      m = m + (g - m) * w1;
      v = v * beta2 + w2 * g * g;
      const float denom = sqrtf(v) / bc2_sqrt + eps;
      p = p - step_size * (m / denom);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
    }
  }
}

}  // namespace ofsv

using namespace ofsv;

// `tensors`: device array of ntensors {p, g, m, v, n} records (5 x 8 bytes each); `chunks`: device array of nchunks int2.
extern "C" int ofsv_adamw_step_f32(const void* tensors, const void* chunks, int ntensors, int nchunks, float lr, float beta1,
                                   float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
  OFSV_REQUIRE(ntensors >= 0 && nchunks >= 0 && step >= 1, "ofsv_adamw_step_f32: bad arguments (ntensors=%d nchunks=%d step=%d)", ntensors, nchunks, step);
  OFSV_REQUIRE(beta1 >= 0.0f && beta1 < 1.0f && beta2 >= 0.0f && beta2 < 1.0f && eps >= 0.0f, "ofsv_adamw_step_f32: bad hyper-parameters");
  if (ntensors == 0 || nchunks == 0) return OFSV_OK;
  OFSV_REQUIRE(tensors && chunks, "ofsv_adamw_step_f32: null pointer");
  // bias corrections in double like the python scalars of torch.optim.adamw._single_tensor_adamw
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const int grid = nchunks < device_num_sms() * 8 ? nchunks : device_num_sms() * 8;
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const AdamWTensor*>(tensors), reinterpret_cast<const int2*>(chunks),
                                                      nchunks, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale);
  return check_launch("adamw_kernel");
}
