// K4 — fused multi-tensor Adam over the flat parameter arena.
//
// Replaces gluon.Trainer('adam', {'learning_rate', 'clip_gradient'}).step(batch_size)
// (/root/reference/music_style_transfer/VarAutoEncoder/trainer.py:94-101,177), i.e. MXNet 1.3
// optimizer.Adam.update + the adam_update operator, one launch per parameter tensor (~46 per step):
//   g = clip(rescale * g + wd * w, +-clip);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
//   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  w -= lr_t * m / (sqrt(v) + eps)
// Here parameters, gradients and both moments are four flat fp32 arenas, so the whole update is ONE
// HBM-bound pass (28 B per parameter).  The step count lives on the device so the launch is
// CUDA-graph replayable; the gradient arena is optionally zeroed in the same pass for the next step.
#include "msx_common.cuh"

namespace {

__global__ void adam_tick_kernel(float* __restrict__ state, float lr, float b1, float b2) {
  pdl_entry();
  // state[0] = t (as float, exact up to 2^24 steps), state[1] = lr_t
  const float t = state[0] + 1.f;
  state[0] = t;
  const double c1 = 1.0 - pow((double)b1, (double)t), c2 = 1.0 - pow((double)b2, (double)t);
  state[1] = (float)((double)lr * sqrt(c2) / c1);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n4, long long n,
                                                   const float* __restrict__ state, float b1, float b2, float eps,
                                                   float wd, float rescale, float clip, int zero_grad) {
  pdl_entry();
  const float lr_t = state[1];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 wv = reinterpret_cast<float4*>(w)[i], gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* wp = &wv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gg = gp[j] * rescale + wd * wp[j];
      if (clip > 0.f) gg = fminf(fmaxf(gg, -clip), clip);
      mp[j] = b1 * mp[j] + (1.f - b1) * gg;
      vp[j] = b2 * vp[j] + (1.f - b2) * gg * gg;
      wp[j] -= lr_t * mp[j] / (sqrtf(vp[j]) + eps);
    }
    reinterpret_cast<float4*>(w)[i] = wv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // scalar tail
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float gg = g[i] * rescale + wd * w[i];
    if (clip > 0.f) gg = fminf(fmaxf(gg, -clip), clip);
    m[i] = b1 * m[i] + (1.f - b1) * gg;
    v[i] = b2 * v[i] + (1.f - b2) * gg * gg;
    w[i] -= lr_t * m[i] / (sqrtf(v[i]) + eps);
    if (zero_grad) g[i] = 0.f;
  }
}

}  // namespace

// shared with adam_nvlink.cu
extern "C" void msx_adam_tick_launch(float* state, float lr, float b1, float b2, void* stream) {
  (void)msx_launch(adam_tick_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, state, lr, b1, b2);
}

extern "C" int msx_adam_step(float* w, float* g, float* m, float* v, long long n, float* state, float lr, float beta1,
                             float beta2, float eps, float wd, float rescale, float clip, int zero_grad, void* stream) {
  MSX_REQUIRE(w && g && m && v && state, "msx_adam_step: null pointer");
  MSX_REQUIRE((((uintptr_t)w | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "msx_adam_step: arenas must be 16-byte aligned");
  if (n == 0) return MSX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  MSX_CUDA(msx_launch(adam_tick_kernel, dim3(1), dim3(1), 0, st, state, lr, beta1, beta2));
  const long long n4 = n / 4;
  const int grid = (int)min((long long)msx_num_sms() * 8, (n4 + 255) / 256 + 1);
  MSX_CUDA(msx_launch(adam_kernel, dim3(grid), dim3(256), 0, st, w, g, m, v, n4, n, state, beta1, beta2, eps, wd, rescale, clip, zero_grad));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
