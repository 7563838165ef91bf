// K2c (any row length) — tiled FFMA attention in the reference's convention, forward + backward.
//
// Same contract as attention.cu (MultiHeadDotAttention.hybrid_forward lines 91-103 and _mask_logits,
// /root/reference/music_style_transfer/VarAutoEncoder/transformer.py:91-126):
//   S[k][q] = K_k . Q_q / sqrt(d_h) + (key k padded ? -1e9 : 0);  P = softmax over the QUERY axis;  O[q] = sum_k P[k][q] V[k]
// attention.cu keeps the whole T x T score matrix of a (batch, head) in shared memory (T <= ~180), the tcgen05 kernels
// stop at T = 384; --max-seq-len is a free flag (VarAutoEncoder/config.py:36) and the positional table is 10 000 long
// (transformer.py:217,255), so this kernel takes every other length: exact fp32, nothing larger than a 32 x 32 tile on chip.
//
// The reference's softmax normalises every KEY row over the queries, so key tiles are independent: one CTA owns 32 keys of
// one (batch, head) and walks the queries in chunks of 32.
//   forward   pass 1: running row max / row sum over all query chunks;  pass 2: P = exp(S - max) / sum, O[q] += P^T V —
//             the 32 x d_h partial of a chunk is added to the context with atomics (the T / 32 key tiles of a (batch, head)
//             all contribute to every query row; the host zero-fills the context first)
//   backward  pass 1 as above;  pass 2: dP = V dO^T, delta_k = sum_q P dP, dV = P dO (registers, plain stores);
//             pass 3: dS = P (dP - delta_k) / sqrt(d_h), dK = dS Q (registers, plain stores), dQ[q] += dS^T K (atomics)
#include "msx_common.cuh"

namespace {

constexpr int KT = 32, QC = 32, kThreadsT = 256;
constexpr int kMaxDh = 64;

struct TiledSmem {
  float Ks[KT][kMaxDh + 1], Vs[KT][kMaxDh + 1], Xs[QC][kMaxDh + 1], Ys[QC][kMaxDh + 1];   // Xs: Q chunk, Ys: dO chunk
  float Ps[KT][QC + 1];
  float rowmask[KT];
};

// cooperative load of `rows` rows of one head slice (row stride ld) into a [..][kMaxDh + 1] tile; rows >= n_valid are zeros
__device__ __forceinline__ void load_tile(float (*dst)[kMaxDh + 1], const float* src, long long ld, int n_valid, int rows,
                                          int dh) {
  for (int i = threadIdx.x; i < rows * dh; i += kThreadsT) {
    const int r = i / dh, d = i % dh;
    dst[r][d] = r < n_valid ? __ldg(src + (long long)r * ld + d) : 0.f;
  }
}

// score of (key kk, query lane) of the current chunk; -inf for queries beyond T
__device__ __forceinline__ float score(const TiledSmem& sm, int kk, int lane, int dh, float inv_scale, bool q_valid) {
  float acc = 0.f;
  for (int d = 0; d < dh; ++d) acc = fmaf(sm.Ks[kk][d], sm.Xs[lane][d], acc);
  return q_valid ? fmaf(acc, inv_scale, sm.rowmask[kk]) : -INFINITY;
}

// pass 1 for the four keys of this warp: running max m[i] and sum l[i] over all query chunks (replicated on every lane)
__device__ __forceinline__ void row_stats(TiledSmem& sm, const float* Qg, long long ld, int T, int dh, float inv_scale,
                                          int warp, int lane, float (&m)[4], float (&l)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f; }
  for (int q0 = 0; q0 < T; q0 += QC) {
    __syncthreads();
    load_tile(sm.Xs, Qg + (long long)q0 * ld, ld, T - q0, QC, dh);
    __syncthreads();
    const bool qv = q0 + lane < T;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float s = score(sm, warp * 4 + i, lane, dh, inv_scale, qv);
      const float nm = fmaxf(m[i], warp_max(s));
      l[i] = l[i] * __expf(m[i] - nm) + warp_sum(qv ? __expf(s - nm) : 0.f);
      m[i] = nm;
    }
  }
}

__global__ void __launch_bounds__(kThreadsT) attn_tiled_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ mask,
                                                                   float* __restrict__ ctx, int T, int H, int dh,
                                                                   float inv_scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TiledSmem& sm = *reinterpret_cast<TiledSmem*>(smem_raw);
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int k0 = blockIdx.x * KT;
  const int D = H * dh;
  const long long ld = 3LL * D;
  const float* Kg = qkv + ((long long)b * T) * ld + h * dh;
  const float* Qg = Kg + D;
  const float* Vg = Kg + 2 * D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk = min(KT, T - k0);
  load_tile(sm.Ks, Kg + (long long)k0 * ld, ld, nk, KT, dh);
  load_tile(sm.Vs, Vg + (long long)k0 * ld, ld, nk, KT, dh);
  if (tid < KT) sm.rowmask[tid] = (tid < nk && __ldg(mask + (long long)b * T + k0 + tid) > 0.f) ? 0.f : -1e9f;
  float m[4], l[4];
  row_stats(sm, Qg, ld, T, dh, inv_scale, warp, lane, m, l);
  const int oq = tid >> 3, og = tid & 7;                 // output mapping of pass 2: query row, column group
  for (int q0 = 0; q0 < T; q0 += QC) {
    __syncthreads();
    load_tile(sm.Xs, Qg + (long long)q0 * ld, ld, T - q0, QC, dh);
    __syncthreads();
    const bool qv = q0 + lane < T;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = warp * 4 + i;
      const float s = score(sm, kk, lane, dh, inv_scale, qv);
      sm.Ps[kk][lane] = (qv && kk < nk) ? __expf(s - m[i]) / l[i] : 0.f;
    }
    __syncthreads();
    if (q0 + oq < T) {
      float* out = ctx + ((long long)b * T + q0 + oq) * D + h * dh;
      for (int d = og; d < dh; d += 8) {
        float acc = 0.f;
#pragma unroll 8
        for (int kk = 0; kk < KT; ++kk) acc = fmaf(sm.Ps[kk][oq], sm.Vs[kk][d], acc);
        atomicAdd(out + d, acc);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreadsT) attn_tiled_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ mask,
                                                                   const float* __restrict__ dctx, float* __restrict__ dqkv,
                                                                   int T, int H, int dh, float inv_scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TiledSmem& sm = *reinterpret_cast<TiledSmem*>(smem_raw);
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int k0 = blockIdx.x * KT;
  const int D = H * dh;
  const long long ld = 3LL * D;
  const float* Kg = qkv + ((long long)b * T) * ld + h * dh;
  const float* Qg = Kg + D;
  const float* Vg = Kg + 2 * D;
  const float* dOg = dctx + ((long long)b * T) * D + h * dh;
  float* dKg = dqkv + ((long long)b * T) * ld + h * dh;
  float* dQg = dKg + D;
  float* dVg = dKg + 2 * D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk = min(KT, T - k0);
  load_tile(sm.Ks, Kg + (long long)k0 * ld, ld, nk, KT, dh);
  load_tile(sm.Vs, Vg + (long long)k0 * ld, ld, nk, KT, dh);
  if (tid < KT) sm.rowmask[tid] = (tid < nk && __ldg(mask + (long long)b * T + k0 + tid) > 0.f) ? 0.f : -1e9f;
  float m[4], l[4];
  row_stats(sm, Qg, ld, T, dh, inv_scale, warp, lane, m, l);
  const int ok = tid >> 3, og = tid & 7;                 // accumulator mapping: key row (or query row for dQ), column group
  float delta[4] = {0.f, 0.f, 0.f, 0.f};
  float dv[kMaxDh / 8], dk[kMaxDh / 8];
#pragma unroll
  for (int j = 0; j < kMaxDh / 8; ++j) dv[j] = dk[j] = 0.f;
  // ---- pass 2: delta_k = sum_q P dP, dV = P dO
  for (int q0 = 0; q0 < T; q0 += QC) {
    __syncthreads();
    load_tile(sm.Xs, Qg + (long long)q0 * ld, ld, T - q0, QC, dh);
    load_tile(sm.Ys, dOg + (long long)q0 * D, D, T - q0, QC, dh);
    __syncthreads();
    const bool qv = q0 + lane < T;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = warp * 4 + i;
      const float s = score(sm, kk, lane, dh, inv_scale, qv);
      const float p = (qv && kk < nk) ? __expf(s - m[i]) / l[i] : 0.f;
      float dp = 0.f;
      for (int d = 0; d < dh; ++d) dp = fmaf(sm.Vs[kk][d], sm.Ys[lane][d], dp);
      delta[i] += warp_sum(p * dp);
      sm.Ps[kk][lane] = p;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kMaxDh / 8; ++j) {
      const int d = og + 8 * j;
      if (d < dh) {
        float acc = 0.f;
#pragma unroll 8
        for (int q = 0; q < QC; ++q) acc = fmaf(sm.Ps[ok][q], sm.Ys[q][d], acc);
        dv[j] += acc;
      }
    }
  }
  // ---- pass 3: dS = P (dP - delta) / sqrt(d_h); dK = dS Q; dQ += dS^T K
  for (int q0 = 0; q0 < T; q0 += QC) {
    __syncthreads();
    load_tile(sm.Xs, Qg + (long long)q0 * ld, ld, T - q0, QC, dh);
    load_tile(sm.Ys, dOg + (long long)q0 * D, D, T - q0, QC, dh);
    __syncthreads();
    const bool qv = q0 + lane < T;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = warp * 4 + i;
      const float s = score(sm, kk, lane, dh, inv_scale, qv);
      const float p = (qv && kk < nk) ? __expf(s - m[i]) / l[i] : 0.f;
      float dp = 0.f;
      for (int d = 0; d < dh; ++d) dp = fmaf(sm.Vs[kk][d], sm.Ys[lane][d], dp);
      sm.Ps[kk][lane] = p * (dp - delta[i]) * inv_scale;             // dS
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kMaxDh / 8; ++j) {
      const int d = og + 8 * j;
      if (d < dh) {
        float acc = 0.f;
#pragma unroll 8
        for (int q = 0; q < QC; ++q) acc = fmaf(sm.Ps[ok][q], sm.Xs[q][d], acc);
        dk[j] += acc;
        if (q0 + ok < T) {                                           // dQ row q0 + ok: sum over this tile's keys
          float aq = 0.f;
#pragma unroll 8
          for (int kk = 0; kk < KT; ++kk) aq = fmaf(sm.Ps[kk][ok], sm.Ks[kk][d], aq);
          atomicAdd(dQg + (long long)(q0 + ok) * ld + d, aq);
        }
      }
    }
  }
  if (ok < nk) {
#pragma unroll
    for (int j = 0; j < kMaxDh / 8; ++j) {
      const int d = og + 8 * j;
      if (d < dh) {
        dKg[(long long)(k0 + ok) * ld + d] = dk[j];
        dVg[(long long)(k0 + ok) * ld + d] = dv[j];
      }
    }
  }
}

}  // namespace

// Any T >= 1, d_h <= 64.  ctx [B*T, H*dh] is zero-filled here (cudaMemsetAsync) before the key tiles accumulate into it.
extern "C" int msx_attention_tiled_fwd(const float* qkv, const float* mask, float* ctx, int B, int T, int H, int dh,
                                       void* stream) {
  MSX_REQUIRE(qkv && mask && ctx, "msx_attention_tiled_fwd: null pointer");
  MSX_REQUIRE(B > 0 && T > 0 && H > 0 && dh > 0 && dh <= kMaxDh, "msx_attention_tiled_fwd: bad shape (d_h <= 64)");
  MSX_REQUIRE((long long)B * H <= 65535, "msx_attention_tiled_fwd: B * H must not exceed 65535");
  cudaStream_t st = (cudaStream_t)stream;
  MSX_CUDA(cudaMemsetAsync(ctx, 0, (size_t)B * T * H * dh * sizeof(float), st));
  MSX_CUDA(cudaFuncSetAttribute(attn_tiled_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TiledSmem)));
  attn_tiled_fwd_kernel<<<dim3(msx_ceil_div(T, KT), B * H), kThreadsT, sizeof(TiledSmem), st>>>(qkv, mask, ctx, T, H, dh,
                                                                                               1.f / sqrtf((float)dh));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_attention_tiled_bwd(const float* qkv, const float* mask, const float* dctx, float* dqkv, int B, int T,
                                       int H, int dh, void* stream) {
  MSX_REQUIRE(qkv && mask && dctx && dqkv, "msx_attention_tiled_bwd: null pointer");
  MSX_REQUIRE(B > 0 && T > 0 && H > 0 && dh > 0 && dh <= kMaxDh, "msx_attention_tiled_bwd: bad shape (d_h <= 64)");
  MSX_REQUIRE((long long)B * H <= 65535, "msx_attention_tiled_bwd: B * H must not exceed 65535");
  cudaStream_t st = (cudaStream_t)stream;
  MSX_CUDA(cudaMemsetAsync(dqkv, 0, (size_t)B * T * 3 * H * dh * sizeof(float), st));
  MSX_CUDA(cudaFuncSetAttribute(attn_tiled_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TiledSmem)));
  attn_tiled_bwd_kernel<<<dim3(msx_ceil_div(T, KT), B * H), kThreadsT, sizeof(TiledSmem), st>>>(qkv, mask, dctx, dqkv, T, H, dh,
                                                                                               1.f / sqrtf((float)dh));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
