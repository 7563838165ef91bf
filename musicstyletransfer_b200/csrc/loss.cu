// K3 — fused losses: reparameterisation + KL, softmax cross-entropy from logits (+ metrics), sigmoid-BCE.
//
// Replaces (reference, /root/reference/music_style_transfer/VarAutoEncoder):
//   model.py:292            z = means + N(0,1) * vars
//   loss.py:8-12            VariationalKLLoss    kl_b = sum_z 0.5 (s^2 + m^2 - 1 - log s^2)
//   model.py:182/256 + loss.py:16-23   softmax(output_layer(x)) -> log -> pick -> mask(label != 0) -> mean over T
//   loss.py:38-81           BinaryCrossEntropy (sigmoid, label smoothing, eps 1e-12, negative-label
//                           down-weighting with the (w*bce)*bce quirk, mean over T*P)
//   trainer.py:107-120,181-186 + metrics.py   Perplexity / Accuracy / TopKAccuracy inputs, accumulated on device
// All HBM-bound: logits are read once in forward (row statistics) and once in backward (in-place gradient).
#include "msx_common.cuh"

namespace {

// ---------------------------------------------------------------- reparameterisation + KL
// lat [B, 2Z] = [means | stds];  z = m + eps * s;  kl_b = sum 0.5 (s^2 + m^2 - 1 - log(s^2))
__global__ void __launch_bounds__(128) reparam_kl_fwd_kernel(const float* __restrict__ lat, const float* __restrict__ eps,
                                                             float* __restrict__ z, float* __restrict__ kl, int Z) {
  pdl_entry();
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < Z; i += blockDim.x) {
    const float m = lat[(size_t)b * 2 * Z + i], s = lat[(size_t)b * 2 * Z + Z + i];
    z[(size_t)b * Z + i] = m + eps[(size_t)b * Z + i] * s;
    acc += 0.5f * (s * s + m * m - 1.f - logf(s * s));
  }
  __shared__ float red[4];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    kl[b] = t;
  }
}

// dlat = [dz + gkl*m | dz*eps + gkl*(s - 1/s)]
__global__ void reparam_kl_bwd_kernel(const float* __restrict__ lat, const float* __restrict__ eps,
                                      const float* __restrict__ dz, const float* __restrict__ gkl, float kl_weight,
                                      float* __restrict__ dlat, int B, int Z) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * Z) return;
  const int b = (int)(i / Z), j = (int)(i % Z);
  const float m = lat[(size_t)b * 2 * Z + j], s = lat[(size_t)b * 2 * Z + Z + j];
  const float g = (gkl ? gkl[b] : 1.f) * kl_weight;
  const float d = dz ? dz[i] : 0.f;
  dlat[(size_t)b * 2 * Z + j] = d + g * m;
  dlat[(size_t)b * 2 * Z + Z + j] = d * eps[i] + g * (s - 1.f / s);
}

__global__ void normal_fill_kernel(float* __restrict__ out, long long n, unsigned long long seed,
                                   const unsigned long long* seed_ctr, unsigned long long offset) {
  pdl_entry();
  seed = msx_eff_seed(seed, seed_ctr);
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n) return;
  const uint4 r = Philox::gen(seed, (unsigned long long)i4, offset);
  const float u0 = 1.f - u32_to_unit(r.x), u1 = u32_to_unit(r.y), u2 = 1.f - u32_to_unit(r.z), u3 = u32_to_unit(r.w);
  const float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
  float v[4];
  sincospif(2.f * u1, &v[1], &v[0]);
  sincospif(2.f * u3, &v[3], &v[2]);
  v[0] *= r0; v[1] *= r0; v[2] *= r1; v[3] *= r1;
  for (int j = 0; j < 4 && i4 * 4 + j < n; ++j) out[i4 * 4 + j] = v[j];
}

// ---------------------------------------------------------------- softmax CE from logits
// One warp per row, the row held in registers (one HBM pass, float4 loads when ld % 4 == 0): log-sum-exp, picked
// logit, rank of the picked logit (accuracy / top-k).  ce[b] = (1/denom) sum_t mask * (lse - logit[label]) is
// accumulated with one atomic per row into a zeroed ce; metrics[0..3] += {sum of min(nll, -log 1e-10), #non-pad
// labels, #argmax hits, #top-k hits} are reduced per CTA first.
constexpr int kCeMaxPerLane = 16;   // V <= 512 on the register path

template <bool VEC4>
__device__ __forceinline__ void ce_load_row(const float* __restrict__ x, int V, int ld, int lane, float (&r)[kCeMaxPerLane]) {
  if (VEC4) {
#pragma unroll
    for (int i = 0; i < kCeMaxPerLane / 4; ++i) {
      const int c = (lane + 32 * i) * 4;
      float4 q = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (c < ld) q = *reinterpret_cast<const float4*>(x + c);
      r[4 * i] = c < V ? q.x : -INFINITY;
      r[4 * i + 1] = c + 1 < V ? q.y : -INFINITY;
      r[4 * i + 2] = c + 2 < V ? q.z : -INFINITY;
      r[4 * i + 3] = c + 3 < V ? q.w : -INFINITY;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kCeMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      r[i] = c < V ? x[c] : -INFINITY;
    }
  }
}
// column of register slot i of `lane`
template <bool VEC4>
__device__ __forceinline__ int ce_col(int lane, int i) { return VEC4 ? (lane + 32 * (i >> 2)) * 4 + (i & 3) : lane + 32 * i; }

template <bool VEC4>
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ logits, int ld,
                                                     const int* __restrict__ labels, float* __restrict__ ce,
                                                     float* __restrict__ lse_out, float* __restrict__ metrics, int rows,
                                                     int T, int V, int top_k, float inv_denom) {
  pdl_entry();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float m_nll = 0.f, m_tok = 0.f, m_hit = 0.f, m_topk = 0.f;
  for (int row = blockIdx.x * nw + warp; row < rows; row += gridDim.x * nw) {
    const int label = __ldg(labels + row);
    float r[kCeMaxPerLane];
    ce_load_row<VEC4>(logits + (size_t)row * ld, V, ld, lane, r);
    float mx = r[0];
#pragma unroll
    for (int i = 1; i < kCeMaxPerLane; ++i) mx = fmaxf(mx, r[i]);
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kCeMaxPerLane; ++i) sum += expf(r[i] - mx);      // exp(-inf) = 0 for the padding slots
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    if (lane == 0) lse_out[row] = lse;
    if (label != 0) {
      const int lc = min(max(label, 0), V - 1);
      float mine = -INFINITY;
#pragma unroll
      for (int i = 0; i < kCeMaxPerLane; ++i) mine = ce_col<VEC4>(lane, i) == lc ? r[i] : mine;
      const float picked = warp_max(mine);
      int greater = 0;
#pragma unroll
      for (int i = 0; i < kCeMaxPerLane; ++i) greater += r[i] > picked ? 1 : 0;
      greater = __reduce_add_sync(MSX_FULL, greater);
      const float nll = lse - picked;
      if (lane == 0) {
        atomicAdd(ce + row / T, nll * inv_denom);
        m_nll += fminf(nll, 23.02585093f);
        m_tok += 1.f;
        m_hit += greater == 0 ? 1.f : 0.f;
        m_topk += greater < top_k ? 1.f : 0.f;
      }
    }
  }
  if (!metrics) return;
  __shared__ float red[8][4];
  if (lane == 0) { red[warp][0] = m_nll; red[warp][1] = m_tok; red[warp][2] = m_hit; red[warp][3] = m_topk; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float a = 0.f;
    for (int w = 0; w < nw; ++w) a += red[w][threadIdx.x];
    if (a != 0.f) atomicAdd(metrics + threadIdx.x, a);
  }
}

// Generic fallback for vocabularies beyond the register path (V > 512): three passes over the row.
__global__ void __launch_bounds__(256) ce_fwd_big_kernel(const float* __restrict__ logits, int ld,
                                                         const int* __restrict__ labels, float* __restrict__ ce,
                                                         float* __restrict__ lse_out, float* __restrict__ metrics, int rows,
                                                         int T, int V, int top_k, float inv_denom) {
  pdl_entry();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int row = blockIdx.x * nw + warp; row < rows; row += gridDim.x * nw) {
    const float* x = logits + (size_t)row * ld;
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, x[v]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int v = lane; v < V; v += 32) sum += expf(x[v] - mx);
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    if (lane == 0) lse_out[row] = lse;
    const int label = labels[row];
    if (label != 0) {
      const float picked = x[min(max(label, 0), V - 1)];
      int greater = 0;
      for (int v = lane; v < V; v += 32) greater += x[v] > picked ? 1 : 0;
      greater = __reduce_add_sync(MSX_FULL, greater);
      const float nll = lse - picked;
      if (lane == 0) {
        atomicAdd(ce + row / T, nll * inv_denom);
        if (metrics) {
          atomicAdd(metrics + 0, fminf(nll, 23.02585093f));
          atomicAdd(metrics + 1, 1.f);
          if (greater == 0) atomicAdd(metrics + 2, 1.f);
          if (greater < top_k) atomicAdd(metrics + 3, 1.f);
        }
      }
    }
  }
}

// in place: logits <- d ce_b / d logits * gout[b] = (softmax - onehot) * mask * gout[b] / denom ; pad columns zeroed.
// dbias (optional, V <= 512): column sums of the gradient = bias gradient of the output layer, accumulated per
// warp in registers over its rows, reduced over the CTA's warps in shared memory, one atomic per column per CTA.
template <bool VEC4>
__global__ void __launch_bounds__(256) ce_bwd_kernel(float* __restrict__ logits, int ld, const int* __restrict__ labels,
                                                     const float* __restrict__ lse, const float* __restrict__ gout,
                                                     int rows, int T, int V, float inv_denom, float* __restrict__ dbias) {
  pdl_entry();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float bs[kCeMaxPerLane];
#pragma unroll
  for (int i = 0; i < kCeMaxPerLane; ++i) bs[i] = 0.f;
  for (int row = blockIdx.x * nw + warp; row < rows; row += gridDim.x * nw) {
    float* x = logits + (size_t)row * ld;
    const int label = __ldg(labels + row);
    const float l = __ldg(lse + row);
    const float g = label != 0 ? (gout ? __ldg(gout + row / T) : 1.f) * inv_denom : 0.f;
    float r[kCeMaxPerLane];
    ce_load_row<VEC4>(x, V, ld, lane, r);
#pragma unroll
    for (int i = 0; i < kCeMaxPerLane; ++i) {
      const int c = ce_col<VEC4>(lane, i);
      r[i] = (c < V && label != 0) ? (expf(r[i] - l) - (c == label ? 1.f : 0.f)) * g : 0.f;
      bs[i] += r[i];
    }
    if (VEC4) {
#pragma unroll
      for (int i = 0; i < kCeMaxPerLane / 4; ++i) {
        const int c = (lane + 32 * i) * 4;
        if (c < ld) *reinterpret_cast<float4*>(x + c) = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kCeMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < ld) x[c] = r[i];
      }
    }
  }
  if (!dbias) return;
  __shared__ float red[8][kCeMaxPerLane * 32];
#pragma unroll
  for (int i = 0; i < kCeMaxPerLane; ++i) red[warp][i * 32 + lane] = bs[i];
  __syncthreads();
  for (int j = threadIdx.x; j < kCeMaxPerLane * 32; j += blockDim.x) {
    float a = 0.f;
    for (int w = 0; w < nw; ++w) a += red[w][j];
    const int i = j >> 5, ln = j & 31;
    const int c = ce_col<VEC4>(ln, i);
    if (c < V && a != 0.f) atomicAdd(dbias + c, a);
  }
}

// Training path: forward and backward of the cross-entropy in ONE pass over the logits (the row is in registers anyway):
// ce / lse / metrics as ce_fwd_kernel, then the row is overwritten with d(sum_b ce_b) / d logits as ce_bwd_kernel with
// head gradient 1 — the logits are read once and written once instead of read twice and written once.
__global__ void __launch_bounds__(256) ce_fwd_bwd_kernel(float* __restrict__ logits, int ld, const int* __restrict__ labels,
                                                         float* __restrict__ ce, float* __restrict__ lse_out,
                                                         float* __restrict__ metrics, int rows, int T, int V, int top_k,
                                                         float inv_denom, float* __restrict__ dbias) {
  pdl_entry();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float m_nll = 0.f, m_tok = 0.f, m_hit = 0.f, m_topk = 0.f;
  float bs[kCeMaxPerLane];
#pragma unroll
  for (int i = 0; i < kCeMaxPerLane; ++i) bs[i] = 0.f;
  for (int row = blockIdx.x * nw + warp; row < rows; row += gridDim.x * nw) {
    float* x = logits + (size_t)row * ld;
    const int label = __ldg(labels + row);
    float r[kCeMaxPerLane];
    ce_load_row<true>(x, V, ld, lane, r);
    float mx = r[0];
#pragma unroll
    for (int i = 1; i < kCeMaxPerLane; ++i) mx = fmaxf(mx, r[i]);
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kCeMaxPerLane; ++i) sum += expf(r[i] - mx);
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    if (lane == 0) lse_out[row] = lse;
    if (label != 0) {
      const int lc = min(max(label, 0), V - 1);
      float mine = -INFINITY;
#pragma unroll
      for (int i = 0; i < kCeMaxPerLane; ++i) mine = ce_col<true>(lane, i) == lc ? r[i] : mine;
      const float picked = warp_max(mine);
      int greater = 0;
#pragma unroll
      for (int i = 0; i < kCeMaxPerLane; ++i) greater += r[i] > picked ? 1 : 0;
      greater = __reduce_add_sync(MSX_FULL, greater);
      const float nll = lse - picked;
      if (lane == 0) {
        atomicAdd(ce + row / T, nll * inv_denom);
        m_nll += fminf(nll, 23.02585093f);
        m_tok += 1.f;
        m_hit += greater == 0 ? 1.f : 0.f;
        m_topk += greater < top_k ? 1.f : 0.f;
      }
    }
    const float g = label != 0 ? inv_denom : 0.f;
#pragma unroll
    for (int i = 0; i < kCeMaxPerLane; ++i) {
      const int c = ce_col<true>(lane, i);
      r[i] = (c < V && label != 0) ? (expf(r[i] - lse) - (c == label ? 1.f : 0.f)) * g : 0.f;
      bs[i] += r[i];
    }
#pragma unroll
    for (int i = 0; i < kCeMaxPerLane / 4; ++i) {
      const int c = (lane + 32 * i) * 4;
      if (c < ld) *reinterpret_cast<float4*>(x + c) = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
    }
  }
  __shared__ float red[8][kCeMaxPerLane * 32];
  if (dbias) {
#pragma unroll
    for (int i = 0; i < kCeMaxPerLane; ++i) red[warp][i * 32 + lane] = bs[i];
    __syncthreads();
    for (int j = threadIdx.x; j < kCeMaxPerLane * 32; j += blockDim.x) {
      float a = 0.f;
      for (int w = 0; w < nw; ++w) a += red[w][j];
      const int c = ce_col<true>(j & 31, j >> 5);
      if (c < V && a != 0.f) atomicAdd(dbias + c, a);
    }
    __syncthreads();
  }
  if (!metrics) return;
  if (lane == 0) { red[warp][0] = m_nll; red[warp][1] = m_tok; red[warp][2] = m_hit; red[warp][3] = m_topk; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float a = 0.f;
    for (int w = 0; w < nw; ++w) a += red[w][threadIdx.x];
    if (a != 0.f) atomicAdd(metrics + threadIdx.x, a);
  }
}

__global__ void __launch_bounds__(256) ce_bwd_big_kernel(float* __restrict__ logits, int ld, const int* __restrict__ labels,
                                                         const float* __restrict__ lse, const float* __restrict__ gout,
                                                         int rows, int T, int V, float inv_denom) {
  pdl_entry();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int row = blockIdx.x * nw + warp; row < rows; row += gridDim.x * nw) {
    float* x = logits + (size_t)row * ld;
    const int label = labels[row];
    const float g = label != 0 ? (gout ? gout[row / T] : 1.f) * inv_denom : 0.f;
    const float l = lse[row];
    for (int v = lane; v < ld; v += 32) {
      float d = 0.f;
      if (v < V && label != 0) d = (expf(x[v] - l) - (v == label ? 1.f : 0.f)) * g;
      x[v] = d;
    }
  }
}

// probs = softmax(logits) (API parity with Model.hybrid_forward's first return value)
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ logits, int ld,
                                                           float* __restrict__ probs, long long rows, int V) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* x = logits + row * ld;
  float mx = -INFINITY;
  for (int v = lane; v < V; v += 32) mx = fmaxf(mx, x[v]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int v = lane; v < V; v += 32) sum += expf(x[v] - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int v = lane; v < V; v += 32) probs[row * V + v] = expf(x[v] - mx) * inv;
}

// SoftmaxCrossEntropy on PROBABILITIES (loss.py:16-23), forward only (API parity; training uses ce_fwd on logits)
__global__ void __launch_bounds__(128) ce_from_probs_kernel(const float* __restrict__ probs,
                                                            const int* __restrict__ labels, float* __restrict__ ce,
                                                            int T, int V) {
  pdl_entry();
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int label = labels[(size_t)b * T + t];
    if (label != 0) acc -= logf(probs[((size_t)b * T + t) * V + min(max(label, 0), V - 1)]);
  }
  __shared__ float red[4];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) ce[b] = (red[0] + red[1] + red[2] + red[3]) / T;
}

// ---------------------------------------------------------------- sigmoid BCE on piano rolls
struct BceCfg {
  int from_sigmoid;
  float smoothing;
  int downweight;
};

__device__ __forceinline__ float bce_elem(float x, float y, const BceCfg& c, float w, float* dldx) {
  const float p = c.from_sigmoid ? x : 1.f / (1.f + expf(-x));
  const float ys = (1.f - c.smoothing) * y + c.smoothing * 0.5f;
  const float lp = logf(1e-12f + p), lq = logf(1e-12f + (1.f - p));
  float bce = -(ys * lp + (1.f - ys) * lq);
  float dbdp = -(ys / (1e-12f + p) - (1.f - ys) / (1e-12f + (1.f - p)));
  float dl = dbdp;
  if (c.downweight && y == 0.f) {
    dl = 2.f * w * bce * dbdp;
    bce = (w * bce) * bce;
  }
  if (dldx) *dldx = c.from_sigmoid ? dl : dl * p * (1.f - p);
  return bce;
}

// one CTA per sample: count positives, then mean over the S*P cells; optionally write the gradient
__global__ void __launch_bounds__(256) bce_kernel(const float* __restrict__ pred, const uint8_t* __restrict__ label,
                                                  float* __restrict__ out, const float* __restrict__ gout,
                                                  float* __restrict__ dpred, int n, BceCfg cfg) {
  pdl_entry();
  const int b = blockIdx.x;
  const float* x = pred + (size_t)b * n;
  const uint8_t* y = label + (size_t)b * n;
  __shared__ float red[8];
  __shared__ float w_sh;
  float npos = 0.f;
  if (cfg.downweight) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) npos += y[i] != 0 ? 1.f : 0.f;
    npos = warp_sum(npos);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = npos;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
      w_sh = t / ((float)n - t + 1e-12f);
    }
    __syncthreads();
  }
  const float w = cfg.downweight ? w_sh : 0.f;
  const float g = dpred ? (gout ? gout[b] : 1.f) / n : 0.f;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float d;
    acc += bce_elem(x[i], y[i] != 0 ? 1.f : 0.f, cfg, w, dpred ? &d : nullptr);   // label = note sounding (velocity rolls too)
    if (dpred) dpred[(size_t)b * n + i] = d * g;
  }
  acc = warp_sum(acc);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && out) {
    float t = 0.f;
    for (int wv = 0; wv < (blockDim.x >> 5); ++wv) t += red[wv];
    out[b] = t / n;
  }
}

// running sums of the step metrics (trainer.py:107-120 CustomMetric means): sums += {sum kl, sum (ce + w kl), B}
__global__ void __launch_bounds__(256) loss_sums_kernel(const float* __restrict__ ce, const float* __restrict__ kl,
                                                        float kl_weight, float* __restrict__ sums, int B) {
  pdl_entry();
  __shared__ float red[2][8];
  float a = 0.f, t = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float k = kl[i];
    a += k;
    t += ce[i] + kl_weight * k;
  }
  a = warp_sum(a);
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = t; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sa = 0.f, stt = 0.f;
    for (int w = 0; w < 8; ++w) { sa += red[0][w]; stt += red[1][w]; }
    sums[0] += sa;
    sums[1] += stt;
    sums[2] += (float)B;
  }
}

}  // namespace

extern "C" int msx_loss_sums(const float* ce, const float* kl, float kl_weight, float* sums, int B, void* stream) {
  MSX_REQUIRE(B >= 0, "msx_loss_sums: B < 0");
  if (B == 0) return MSX_OK;
  MSX_REQUIRE(ce && kl && sums, "msx_loss_sums: null pointer");
  MSX_CUDA(msx_launch(loss_sums_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, ce, kl, kl_weight, sums, B));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_reparam_kl_fwd(const float* lat, const float* eps, float* z, float* kl, int B, int Z, void* stream) {
  MSX_REQUIRE(lat && eps && z && kl, "msx_reparam_kl_fwd: null pointer");
  if (B == 0) return MSX_OK;
  MSX_CUDA(msx_launch(reparam_kl_fwd_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, lat, eps, z, kl, Z));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_reparam_kl_bwd(const float* lat, const float* eps, const float* dz, const float* gkl, float kl_weight,
                                  float* dlat, int B, int Z, void* stream) {
  MSX_REQUIRE(lat && eps && dlat, "msx_reparam_kl_bwd: null pointer");
  if (B == 0) return MSX_OK;
  const long long n = (long long)B * Z;
  MSX_CUDA(msx_launch(reparam_kl_bwd_kernel, dim3(msx_ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, lat, eps, dz, gkl, kl_weight, dlat, B, Z));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_normal_fill(float* out, long long n, unsigned long long seed, unsigned long long offset,
                               void* stream) {
  MSX_REQUIRE(out || n == 0, "msx_normal_fill: null pointer");
  if (n == 0) return MSX_OK;
  MSX_CUDA(msx_launch(normal_fill_kernel, dim3(msx_ceil_div((n + 3) / 4, 256)), dim3(256), 0, (cudaStream_t)stream, out, n, seed, msx_step_counter(), offset));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_ce_fwd(const float* logits, int ld, const int32_t* labels, float* ce, float* lse, float* metrics,
                          int B, int T, int V, int denom, int top_k, void* stream) {
  MSX_REQUIRE(logits && labels && ce && lse, "msx_ce_fwd: null pointer");
  MSX_REQUIRE(ld >= V && V > 0, "msx_ce_fwd: bad leading dimension");
  if (B == 0) return MSX_OK;
  MSX_REQUIRE(denom > 0, "msx_ce_fwd: denom must be > 0");
  MSX_REQUIRE((long long)B * T < (1ll << 31), "msx_ce_fwd: B * T must fit 31 bits");
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = B * T;
  MSX_CUDA(cudaMemsetAsync(ce, 0, (size_t)B * sizeof(float), st));      // rows add their share atomically
  const int grid = min(msx_num_sms() * 8, (rows + 7) / 8);
  const bool vec4 = (ld & 3) == 0 && ((uintptr_t)logits & 15) == 0;
  if (V > 32 * kCeMaxPerLane)
    MSX_CUDA(msx_launch(ce_fwd_big_kernel, dim3(grid), dim3(256), 0, st, logits, ld, labels, ce, lse, metrics, rows, T, V, top_k, 1.f / denom));
  else if (vec4)
    MSX_CUDA(msx_launch(ce_fwd_kernel<true>, dim3(grid), dim3(256), 0, st, logits, ld, labels, ce, lse, metrics, rows, T, V, top_k, 1.f / denom));
  else
    MSX_CUDA(msx_launch(ce_fwd_kernel<false>, dim3(grid), dim3(256), 0, st, logits, ld, labels, ce, lse, metrics, rows, T, V, top_k, 1.f / denom));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_ce_bwd(float* logits_inout, int ld, const int32_t* labels, const float* lse, const float* gout, int B,
                          int T, int V, int denom, float* dbias, void* stream) {
  MSX_REQUIRE(logits_inout && labels && lse, "msx_ce_bwd: null pointer");
  if (B == 0) return MSX_OK;
  MSX_REQUIRE((long long)B * T < (1ll << 31), "msx_ce_bwd: B * T must fit 31 bits");
  const int rows = B * T;
  MSX_REQUIRE(dbias == nullptr || V <= 32 * kCeMaxPerLane, "msx_ce_bwd: fused bias gradient supports V <= 512");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = min(msx_num_sms() * 8, (rows + 7) / 8);
  const bool vec4 = (ld & 3) == 0 && ((uintptr_t)logits_inout & 15) == 0;
  if (V > 32 * kCeMaxPerLane)
    MSX_CUDA(msx_launch(ce_bwd_big_kernel, dim3(grid), dim3(256), 0, st, logits_inout, ld, labels, lse, gout, rows, T, V, 1.f / denom));
  else if (vec4)
    MSX_CUDA(msx_launch(ce_bwd_kernel<true>, dim3(grid), dim3(256), 0, st, logits_inout, ld, labels, lse, gout, rows, T, V, 1.f / denom, dbias));
  else
    MSX_CUDA(msx_launch(ce_bwd_kernel<false>, dim3(grid), dim3(256), 0, st, logits_inout, ld, labels, lse, gout, rows, T, V, 1.f / denom, dbias));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

// Fused training path (head gradient 1): see ce_fwd_bwd_kernel.  Needs the register path (V <= 512) and 16-byte aligned rows;
// returns MSX_ERR_UNSUPPORTED otherwise (callers then use msx_ce_fwd + msx_ce_bwd).
extern "C" int msx_ce_fwd_bwd(float* logits_inout, int ld, const int32_t* labels, float* ce, float* lse, float* metrics, int B,
                              int T, int V, int denom, int top_k, float* dbias, void* stream) {
  MSX_REQUIRE(logits_inout && labels && ce && lse, "msx_ce_fwd_bwd: null pointer");
  MSX_REQUIRE(ld >= V && V > 0, "msx_ce_fwd_bwd: bad leading dimension");
  if (B == 0) return MSX_OK;
  MSX_REQUIRE(denom > 0, "msx_ce_fwd_bwd: denom must be > 0");
  MSX_REQUIRE((long long)B * T < (1ll << 31), "msx_ce_fwd_bwd: B * T must fit 31 bits");
  if (V > 32 * kCeMaxPerLane || (ld & 3) != 0 || ((uintptr_t)logits_inout & 15) != 0) {
    msx_set_error("msx_ce_fwd_bwd: needs V <= 512 and 16-byte aligned rows");
    return MSX_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = B * T;
  MSX_CUDA(cudaMemsetAsync(ce, 0, (size_t)B * sizeof(float), st));
  // one resident wave of CTAs (persistent row loop): no partial last wave, and the per-CTA bias-gradient flush is paid once
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ce_fwd_bwd_kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
  const int grid = min(msx_num_sms() * per_sm, (rows + 7) / 8);
  MSX_CUDA(msx_launch(ce_fwd_bwd_kernel, dim3(grid), dim3(256), 0, st, logits_inout, ld, labels, ce, lse, metrics, rows, T, V, top_k, 1.f / denom, dbias));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_softmax_rows(const float* logits, int ld, float* probs, long long rows, int V, void* stream) {
  MSX_REQUIRE(logits && probs, "msx_softmax_rows: null pointer");
  if (rows == 0) return MSX_OK;
  MSX_CUDA(msx_launch(softmax_rows_kernel, dim3(msx_ceil_div(rows, 8)), dim3(256), 0, (cudaStream_t)stream, logits, ld, probs, rows, V));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_ce_from_probs(const float* probs, const int32_t* labels, float* ce, int B, int T, int V,
                                 void* stream) {
  MSX_REQUIRE(probs && labels && ce, "msx_ce_from_probs: null pointer");
  if (B == 0) return MSX_OK;
  MSX_CUDA(msx_launch(ce_from_probs_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, probs, labels, ce, T, V));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_bce(const float* pred, const uint8_t* label, float* out, const float* gout, float* dpred, int B,
                       int n_per_sample, int from_sigmoid, float label_smoothing, int downweight, void* stream) {
  MSX_REQUIRE(pred && label && (out || dpred), "msx_bce: null pointer");
  if (B == 0) return MSX_OK;
  BceCfg cfg{from_sigmoid, label_smoothing, downweight};
  MSX_CUDA(msx_launch(bce_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, pred, label, out, gout, dpred, n_per_sample, cfg));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
