// A12 — multinomial sampling of the next event token from decoder logits.
//
// Replaces `mx.nd.random.multinomial(probs)` + `-log(pick(probs, next))` in Sampling.sample
// (/root/reference/music_style_transfer/VarAutoEncoder/sampler.py:181-184) on the probabilities
// `softmax(output_layer(h))` (model.py:198 / :271).  One warp per batch row: softmax statistics, an
// inclusive warp scan of the probabilities in vocabulary order, and the first index whose cumulative
// probability exceeds u (u from the caller for parity runs, else Philox keyed by (seed, step, row)).
#include "msx_common.cuh"

namespace {

__global__ void __launch_bounds__(256) sample_kernel(const float* __restrict__ logits, int ld, int V,
                                                     const float* __restrict__ uniforms, unsigned long long seed,
                                                     unsigned long long step, int* __restrict__ next,
                                                     float* __restrict__ score, int* __restrict__ out_seq, int out_ld,
                                                     int out_col, int B) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* x = logits + (size_t)row * ld;
  float mx = -INFINITY;
  for (int v = lane; v < V; v += 32) mx = fmaxf(mx, x[v]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int v = lane; v < V; v += 32) sum += expf(x[v] - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  float u;
  if (uniforms) u = uniforms[row];
  else u = u32_to_unit(Philox::gen(seed, (unsigned long long)row, step).x);
  // contiguous chunks per lane so that the scan runs in vocabulary order
  const int per = (V + 31) / 32;
  const int v0 = lane * per, v1 = min(V, v0 + per);
  float local = 0.f;
  for (int v = v0; v < v1; ++v) local += expf(x[v] - mx) * inv;
  float incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(MSX_FULL, incl, o);
    if (lane >= o) incl += t;
  }
  float run = incl - local;
  int pick = V;   // first v with cdf(v) > u
  for (int v = v0; v < v1; ++v) {
    run += expf(x[v] - mx) * inv;
    if (run > u) { pick = v; break; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pick = min(pick, __shfl_xor_sync(MSX_FULL, pick, o));
  if (pick >= V) pick = V - 1;
  if (lane == 0) {
    next[row] = pick;
    if (score) score[row] += -(x[pick] - mx - logf(sum));
    if (out_seq) out_seq[(size_t)row * out_ld + out_col] = pick;
  }
}

}  // namespace

extern "C" int msx_sample_multinomial(const float* logits, int ld, int V, const float* uniforms, unsigned long long seed,
                                      unsigned long long step, int32_t* next, float* score, int32_t* out_seq,
                                      int out_ld, int out_col, int B, void* stream) {
  MSX_REQUIRE(logits && next, "msx_sample_multinomial: null pointer");
  MSX_REQUIRE(V > 0 && ld >= V, "msx_sample_multinomial: bad vocabulary size");
  if (B == 0) return MSX_OK;
  sample_kernel<<<msx_ceil_div(B, 8), 256, 0, (cudaStream_t)stream>>>(logits, ld, V, uniforms, seed, step, next, score,
                                                                    out_seq, out_ld, out_col, B);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
