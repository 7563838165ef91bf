// K2c — fused multi-head attention in the reference's (non-standard) convention, forward + backward.
//
// Replaces MultiHeadDotAttention.hybrid_forward lines 91-103 and _mask_logits
// (/root/reference/music_style_transfer/VarAutoEncoder/transformer.py:91-126):
//   S[k][q] = K_k . Q_q / sqrt(d_h) + (key k padded ? -1e9 : 0)        (:96-99, :106-116)
//   P       = softmax(S, axis = q)   -- normalised over the QUERY axis  (:100)
//   O[q]    = sum_k P[k][q] V[k]                                        (:102)
// In fp32 a padded key row is exactly -1e9 everywhere, so it softmaxes to the uniform 1/T_Q and adds
// V[k]/T_Q to every query — reproduced here by doing the same arithmetic in the same order.
//
// One CTA per (batch, head); K, Q, V (and dO) tiles and the full T x T score matrix live in shared
// memory (T <= ~180 at d_h = 32), every product is a 4x4 register-blocked FFMA loop.  The backward
// recomputes S and P, uses delta_k = V_k . dV_k (= sum_q P dP) so that only one T x T buffer is needed.
#include "msx_common.cuh"

namespace {

constexpr int kThreads = 128;

template <int DH>
struct AttnSmem {
  float *K, *Q, *V, *dO, *S, *rowmask;
  int ld, lds;
  __device__ AttnSmem(float* base, int T, bool bwd) {
    ld = DH + 1;
    lds = T + 1;
    K = base;
    Q = K + T * ld;
    V = Q + T * ld;
    dO = V + T * ld;
    S = dO + (bwd ? T * ld : 0);
    rowmask = S + T * lds;
  }
};

__host__ __device__ inline size_t attn_smem_floats(int T, int DH, bool bwd) {
  return (size_t)(bwd ? 4 : 3) * T * (DH + 1) + (size_t)T * (T + 1) + T;
}

// S = scale * K Q^T + mask, then row softmax over q (in place -> P)
template <int DH>
__device__ void scores_softmax(const AttnSmem<DH>& sm, int T, float scale) {
  const int tid = threadIdx.x;
  const int nb = (T + 3) / 4;
  for (int blk = tid; blk < nb * nb; blk += kThreads) {
    const int k0 = blk / nb, q0 = blk % nb;   // rows k0 + i*nb, q0 + j*nb: consecutive lanes -> consecutive rows
    float acc[4][4] = {};
#pragma unroll 8
    for (int d = 0; d < DH; ++d) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = sm.K[min(k0 + i * nb, T - 1) * sm.ld + d];
        b[i] = sm.Q[min(q0 + i * nb, T - 1) * sm.ld + d];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k0 + i * nb < T && q0 + j * nb < T)
          sm.S[(k0 + i * nb) * sm.lds + q0 + j * nb] = acc[i][j] / scale + sm.rowmask[k0 + i * nb];
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int k = warp; k < T; k += kThreads / 32) {
    float* row = sm.S + k * sm.lds;
    float mx = -INFINITY;
    for (int q = lane; q < T; q += 32) mx = fmaxf(mx, row[q]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int q = lane; q < T; q += 32) {
      const float e = expf(row[q] - mx);
      row[q] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int q = lane; q < T; q += 32) row[q] *= inv;
  }
  __syncthreads();
}

template <int DH>
__device__ void load_tile(float* dst, int ld, const float* src, int row_stride, int T) {
  for (int i = threadIdx.x; i < T * DH; i += kThreads) {
    const int t = i / DH, d = i % DH;
    dst[t * ld + d] = __ldg(src + (size_t)t * row_stride + d);
  }
}

// qkv [B*T, 3*D] rows = [K | Q | V]; mask [B*T] (1 = real key, 0 = padded); ctx [B*T, D]
template <int DH>
__global__ void __launch_bounds__(kThreads) attn_fwd_kernel(const float* __restrict__ qkv,
                                                            const float* __restrict__ mask, float* __restrict__ ctx,
                                                            int T, int H, float scale) {
  extern __shared__ __align__(16) float smem_f[];
  AttnSmem<DH> sm(smem_f, T, false);
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * DH;
  const float* base = qkv + (size_t)b * T * 3 * D + h * DH;
  load_tile<DH>(sm.K, sm.ld, base, 3 * D, T);
  load_tile<DH>(sm.Q, sm.ld, base + D, 3 * D, T);
  load_tile<DH>(sm.V, sm.ld, base + 2 * D, 3 * D, T);
  for (int k = threadIdx.x; k < T; k += kThreads) sm.rowmask[k] = __ldg(mask + (size_t)b * T + k) > 0.f ? 0.f : -1e9f;
  __syncthreads();
  scores_softmax<DH>(sm, T, scale);
  // O[q][d] = sum_k P[k][q] V[k][d]   -- 4 queries x 4 dims per work item
  const int nq = (T + 3) / 4, nd = DH / 4;
  for (int blk = threadIdx.x; blk < nq * nd; blk += kThreads) {
    const int q0 = blk % nq, d0 = (blk / nq) * 4;   // queries q0 + i*nq
    float acc[4][4] = {};
    for (int k = 0; k < T; ++k) {
      float pq[4], vd[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pq[i] = sm.S[k * sm.lds + min(q0 + i * nq, T - 1)];
        vd[i] = sm.V[k * sm.ld + d0 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pq[i], vd[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (q0 + i * nq < T) {
        float* dst = ctx + ((size_t)b * T + q0 + i * nq) * D + h * DH + d0;
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = acc[i][j];
      }
  }
}

// dctx [B*T, D] -> dqkv [B*T, 3*D]
template <int DH>
__global__ void __launch_bounds__(kThreads) attn_bwd_kernel(const float* __restrict__ qkv,
                                                            const float* __restrict__ mask,
                                                            const float* __restrict__ dctx, float* __restrict__ dqkv,
                                                            int T, int H, float scale) {
  extern __shared__ __align__(16) float smem_f[];
  AttnSmem<DH> sm(smem_f, T, true);
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * DH;
  const float* base = qkv + (size_t)b * T * 3 * D + h * DH;
  load_tile<DH>(sm.K, sm.ld, base, 3 * D, T);
  load_tile<DH>(sm.Q, sm.ld, base + D, 3 * D, T);
  load_tile<DH>(sm.V, sm.ld, base + 2 * D, 3 * D, T);
  load_tile<DH>(sm.dO, sm.ld, dctx + (size_t)b * T * D + h * DH, D, T);
  for (int k = threadIdx.x; k < T; k += kThreads) sm.rowmask[k] = __ldg(mask + (size_t)b * T + k) > 0.f ? 0.f : -1e9f;
  __syncthreads();
  scores_softmax<DH>(sm, T, scale);
  float* dbase = dqkv + (size_t)b * T * 3 * D + h * DH;
  const int nt = (T + 3) / 4, nd = DH / 4;
  // dV[k][d] = sum_q P[k][q] dO[q][d];  delta_k = V_k . dV_k  (kept in rowmask, no longer needed)
  for (int blk = threadIdx.x; blk < nt * nd; blk += kThreads) {
    const int k0 = blk % nt, d0 = (blk / nt) * 4;   // keys k0 + i*nt
    float acc[4][4] = {};
    for (int q = 0; q < T; ++q) {
      float pk[4], od[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pk[i] = sm.S[min(k0 + i * nt, T - 1) * sm.lds + q];
        od[i] = sm.dO[q * sm.ld + d0 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pk[i], od[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (k0 + i * nt < T) {
        float* dst = dbase + (size_t)(k0 + i * nt) * 3 * D + 2 * D + d0;
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = acc[i][j];
      }
  }
  __syncthreads();  // dV globally written by this CTA; re-read below through L1/L2 (same thread block -> visible after sync)
  for (int k = threadIdx.x; k < T; k += kThreads) {
    const float* dv = dbase + (size_t)k * 3 * D + 2 * D;
    float s = 0.f;
    for (int d = 0; d < DH; ++d) s = fmaf(sm.V[k * sm.ld + d], dv[d], s);
    sm.rowmask[k] = s;
  }
  __syncthreads();
  // dS[k][q] = P[k][q] * (dO_q . V_k - delta_k) / scale   (in place over P)
  for (int blk = threadIdx.x; blk < nt * nt; blk += kThreads) {
    const int k0 = blk / nt, q0 = blk % nt;
    float acc[4][4] = {};
#pragma unroll 8
    for (int d = 0; d < DH; ++d) {
      float a[4], c[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = sm.V[min(k0 + i * nt, T - 1) * sm.ld + d];
        c[i] = sm.dO[min(q0 + i * nt, T - 1) * sm.ld + d];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], c[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k0 + i * nt < T && q0 + j * nt < T) {
          float* s = sm.S + (k0 + i * nt) * sm.lds + q0 + j * nt;
          *s = *s * (acc[i][j] - sm.rowmask[k0 + i * nt]) / scale;
        }
  }
  __syncthreads();
  // dK[k][d] = sum_q dS[k][q] Q[q][d]
  for (int blk = threadIdx.x; blk < nt * nd; blk += kThreads) {
    const int k0 = blk % nt, d0 = (blk / nt) * 4;
    float acc[4][4] = {};
    for (int q = 0; q < T; ++q) {
      float sk[4], qd[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sk[i] = sm.S[min(k0 + i * nt, T - 1) * sm.lds + q];
        qd[i] = sm.Q[q * sm.ld + d0 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(sk[i], qd[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (k0 + i * nt < T) {
        float* dst = dbase + (size_t)(k0 + i * nt) * 3 * D + d0;
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = acc[i][j];
      }
  }
  // dQ[q][d] = sum_k dS[k][q] K[k][d]
  for (int blk = threadIdx.x; blk < nt * nd; blk += kThreads) {
    const int q0 = blk % nt, d0 = (blk / nt) * 4;
    float acc[4][4] = {};
    for (int k = 0; k < T; ++k) {
      float sq[4], kd[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sq[i] = sm.S[k * sm.lds + min(q0 + i * nt, T - 1)];
        kd[i] = sm.K[k * sm.ld + d0 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(sq[i], kd[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (q0 + i * nt < T) {
        float* dst = dbase + (size_t)(q0 + i * nt) * 3 * D + D + d0;
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = acc[i][j];
      }
  }
}

template <int DH>
int launch_attn(bool bwd, const float* qkv, const float* mask, const float* dctx, float* out, int B, int T, int H,
                cudaStream_t st) {
  const size_t smem = attn_smem_floats(T, DH, bwd) * sizeof(float);
  MSX_REQUIRE(smem <= 227 * 1024, "msx_attention: T=%d does not fit in shared memory (d_h=%d)", T, DH);
  const float scale = sqrtf((float)DH);
  if (!bwd) {
    MSX_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_kernel<DH><<<B * H, kThreads, smem, st>>>(qkv, mask, out, T, H, scale);
  } else {
    MSX_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_kernel<DH><<<B * H, kThreads, smem, st>>>(qkv, mask, dctx, out, T, H, scale);
  }
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

int dispatch_attn(bool bwd, const float* qkv, const float* mask, const float* dctx, float* out, int B, int T, int H,
                  int dh, cudaStream_t st) {
  switch (dh) {
    case 4: return launch_attn<4>(bwd, qkv, mask, dctx, out, B, T, H, st);
    case 8: return launch_attn<8>(bwd, qkv, mask, dctx, out, B, T, H, st);
    case 16: return launch_attn<16>(bwd, qkv, mask, dctx, out, B, T, H, st);
    case 32: return launch_attn<32>(bwd, qkv, mask, dctx, out, B, T, H, st);
    case 64: return launch_attn<64>(bwd, qkv, mask, dctx, out, B, T, H, st);
    default: msx_set_error("msx_attention: head dim %d unsupported (4,8,16,32,64)", dh); return MSX_ERR_UNSUPPORTED;
  }
}

}  // namespace

extern "C" int msx_attention_tiled_fwd(const float* qkv, const float* mask, float* ctx, int B, int T, int H, int dh,
                                       void* stream);
extern "C" int msx_attention_tiled_bwd(const float* qkv, const float* mask, const float* dctx, float* dqkv, int B, int T,
                                       int H, int dh, void* stream);

// rows whose T x T score matrix does not fit one SM's shared memory go to the key-tiled kernel (attention_tiled.cu)
static bool fits_one_cta(int T, int dh, bool bwd) { return attn_smem_floats(T, dh, bwd) * sizeof(float) <= 227 * 1024; }

extern "C" int msx_attention_fwd(const float* qkv, const float* mask, float* ctx, int B, int T, int H, int dh,
                                 void* stream) {
  MSX_REQUIRE(qkv && mask && ctx, "msx_attention_fwd: null pointer");
  MSX_REQUIRE(B > 0 && T > 0 && H > 0, "msx_attention_fwd: bad shape");
  if (!fits_one_cta(T, dh, false)) return msx_attention_tiled_fwd(qkv, mask, ctx, B, T, H, dh, stream);
  return dispatch_attn(false, qkv, mask, nullptr, ctx, B, T, H, dh, (cudaStream_t)stream);
}

extern "C" int msx_attention_bwd(const float* qkv, const float* mask, const float* dctx, float* dqkv, int B, int T,
                                 int H, int dh, void* stream) {
  MSX_REQUIRE(qkv && mask && dctx && dqkv, "msx_attention_bwd: null pointer");
  MSX_REQUIRE(B > 0 && T > 0 && H > 0, "msx_attention_bwd: bad shape");
  if (!fits_one_cta(T, dh, true)) return msx_attention_tiled_bwd(qkv, mask, dctx, dqkv, B, T, H, dh, stream);
  return dispatch_attn(true, qkv, mask, dctx, dqkv, B, T, H, dh, (cudaStream_t)stream);
}
