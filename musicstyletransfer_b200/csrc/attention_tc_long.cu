// K2c (tensor-core path, long rows) — attention in the reference's convention on tcgen05 for 128 < T <= 384, d_h = 32:
// the L = 128 / 256 points of the BASELINE config 3 sweep (T = 129 / 257), which attention_tc.cu (one key tile on the
// 128 TMEM lanes) does not take.
//
// Same contract as attention.cu / attention_tc.cu (replaces MultiHeadDotAttention.hybrid_forward lines 91-103 and
// _mask_logits, /root/reference/music_style_transfer/VarAutoEncoder/transformer.py:91-126):
//   S[k][q] = K_k . Q_q / sqrt(d_h) + (key k padded ? -1e9 : 0);  P = softmax over the QUERY axis;  O[q] = sum_k P[k][q] V[k]
//
// The softmax normalises each KEY row over all queries, so key tiles are independent: one CTA (128 threads) per
// (batch, head) walks the key tiles of 128; for a key tile the scores of ALL queries accumulate in TMEM with the keys
// on the lanes (S[128 x TQ], up to 384 columns), the thread that owns a key row does max / exp2 / sum over its TMEM
// columns, and the normalised row goes out in query chunks of 128 as the MN-major A operand of O[chunk] += P^T V
// (double-buffered in shared memory), O accumulating over key tiles in TMEM.  The forward saves (max * log2 e,
// 1 / sum) per key row; the backward rebuilds P from them chunk by chunk (flash-attention style), so it never needs
// more than 128 score columns at a time:
//   phase 1 (per key tile, over query chunks):  P -> TMEM,  dV += P dO            (A operand from TMEM)
//   delta_k = V_k . dV_k                        (= sum_q P dP, thread-local: dV has the keys on the lanes)
//   phase 2 (over query chunks):  dP = V dO^T,  dS = P (dP - delta) / sqrt(d_h) -> TMEM and, transposed, shared memory
//                                 dK += dS Q (A from TMEM),  dQ[chunk] += dS^T K (A MN-major from shared memory)
// dQ accumulates over key tiles in TMEM (32 columns per query chunk).
#include "attention_tc_common.cuh"

namespace {

constexpr int kTile = 128;                  // keys per key tile = queries per query chunk
constexpr int kTileBytes = kTile * 128;     // one [128 rows][128 B] operand tile (K-major or MN-major)

struct AttnLongParams {
  const float* mask;   // [B*T]
  float* stats;        // [B*H*T, 2]: (row max * log2 e, 1 / row sum) of every key row; written by fwd, read by bwd
  void* out;           // fwd: ctx [B*T, H*32]; bwd: dqkv [B*T, 3*H*32]; fp32, or bf16 when out_bf16
  int out_bf16;
  float* dbias;        // bwd, optional [3*H*32]: += column sums of dqkv
  int T, H, TQ;        // TQ = roundup16(T)
  float inv_scale;
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ forward
template <int NT>
__global__ void __launch_bounds__(128, 1)
    attn_tcl_fwd_kernel(const __grid_constant__ CUtensorMap tmKm /* K-major {32, 128} box over qkv */,
                        const __grid_constant__ CUtensorMap tmMn /* MN-major {32, 128} box over qkv */,
                        const AttnLongParams p) {
  pdl_entry();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sQ = base;                          // NT K-major tiles: all queries of the (batch, head)
  unsigned char* sK = sQ + NT * kTileBytes;          // K-major key tile
  unsigned char* sV = sK + kTileBytes;               // MN-major value tile (d contiguous, 128 key rows)
  unsigned char* sP = sV + kTileBytes;               // 2 x 4 slabs: P chunk, q contiguous, 128 key rows per slab
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sP + 2 * 4 * kTileBytes);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);
  unsigned long long* bar_q = &bars[0];
  unsigned long long* bar_k = &bars[1];
  unsigned long long* bar_v = &bars[2];
  unsigned long long* bar_s = &bars[3];
  unsigned long long* bar_o = &bars[4];              // [2]

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T, TQ = p.TQ, D = p.H * DH;

  // P staging must hold finite values everywhere the MMAs read (short last query chunk): zero it once
  for (int i = tid * 16; i < 2 * 4 * kTileBytes; i += 128 * 16) *reinterpret_cast<float4*>(sP + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_async_smem();
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = *tmem_slot;
  const unsigned tm_O = tmem, tm_S = tmem + 128;     // O: NT x 32 columns; S: up to 384 columns
  const unsigned lane_off = (unsigned)(warp * 32) << 16;
  // MMA 2: A = P^T MN-major (queries contiguous), B = V MN-major (d contiguous), M = 128 queries, N = 32
  const unsigned idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(DH >> 3) << 17) |
                          ((unsigned)(128 >> 4) << 24);

  if (tid == 0) {
    mbar_expect_tx(bar_q, (unsigned)(NT * kTileBytes));
    for (int qc = 0; qc < NT; ++qc) tma_load_2d(sQ + qc * kTileBytes, &tmKm, bar_q, D + h * DH, b * T + qc * kTile);
    mbar_expect_tx(bar_k, (unsigned)kTileBytes);
    tma_load_2d(sK, &tmKm, bar_k, h * DH, b * T);
    mbar_expect_tx(bar_v, (unsigned)kTileBytes);
    tma_load_2d(sV, &tmMn, bar_v, 2 * D + h * DH, b * T);
  }
  int n = 0;                                          // running (key tile, query chunk) counter: P buffer = n & 1
  for (int kt = 0; kt < NT; ++kt) {
    const unsigned par = (unsigned)(kt & 1);
    if (tid == 0) {
      if (kt == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_k, par);
      tc_fence_after();
      // MMA 1: S[128 keys x TQ] = K Q^T in query chunks of <= 128 columns (both operands K-major, +32 B per k-step)
      for (int qc = 0; qc < NT; ++qc) {
        const int nq = min(kTile, TQ - qc * kTile);
        const unsigned idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(nq >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
#pragma unroll
        for (int k = 0; k < DH / 8; ++k)
          umma_tf32(tm_S + qc * kTile, make_desc(smem_u32(sK) + k * 32, 16, 1024, 2),
                    make_desc(smem_u32(sQ) + qc * kTileBytes + k * 32, 16, 1024, 2), idesc1, k > 0 ? 1u : 0u);
      }
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, par);
    tc_fence_after();
    if (tid == 0 && kt + 1 < NT) {                    // the K tile is free once MMA 1 has retired
      mbar_expect_tx(bar_k, (unsigned)kTileBytes);
      tma_load_2d(sK, &tmKm, bar_k, h * DH, b * T + (kt + 1) * kTile);
    }
    const int kg = kt * kTile + tid;                  // this thread's key row
    const bool valid = kg < T;
    const float rowmask = (valid && __ldg(p.mask + (size_t)b * T + kg) > 0.f) ? 0.f : -1e9f;
    // pass 1: row maximum over the T queries
    float mx = -INFINITY;
    for (int c = 0; c < TQ; c += 16) {
      float v[16];
      tmem_ld16(tm_S + lane_off + c, v);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c + j < T) mx = fmaxf(mx, fmaf(v[j], p.inv_scale, rowmask));
    }
    const float mxl = mx * kLog2e;
    // pass 2: e = exp2(s * log2 e - max * log2 e) back into TMEM, row sum
    float sum = 0.f;
    for (int c = 0; c < TQ; c += 16) {
      float v[16];
      tmem_ld16(tm_S + lane_off + c, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] = (c + j < T) ? exp2f(fmaf(fmaf(v[j], p.inv_scale, rowmask), kLog2e, -mxl)) : 0.f;
        sum += v[j];
      }
      tmem_st16(tm_S + lane_off + c, v);
    }
    tmem_st_wait();
    const float inv = valid ? 1.f / sum : 0.f;        // rows beyond T (neighbouring sequence / zeros) contribute nothing
    if (valid) {
      float2* st = reinterpret_cast<float2*>(p.stats) + ((size_t)(b * p.H + h) * T + kg);
      *st = make_float2(mxl, inv);
    }
    // pass 3: normalised row -> shared memory, one query chunk at a time, each chunk followed by its MMA 2
    for (int qc = 0; qc < NT; ++qc, ++n) {
      const int buf = n & 1;
      unsigned char* pb = sP + buf * 4 * kTileBytes;
      if (n >= 2) {                                    // the MMA that read this buffer two chunks ago has retired
        mbar_wait(&bar_o[buf], (unsigned)(((n >> 1) - 1) & 1));
        tc_fence_after();
      }
      const int nq = min(kTile, TQ - qc * kTile);
      for (int c = 0; c < nq; c += 16) {
        float v[16];
        tmem_ld16(tm_S + lane_off + qc * kTile + c, v);
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(pb + mn_major_off(c + j, tid, kTile)) =
              make_float4(to_tf32(v[j] * inv), to_tf32(v[j + 1] * inv), to_tf32(v[j + 2] * inv), to_tf32(v[j + 3] * inv));
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        if (qc == 0) mbar_wait(bar_v, par);
        // MMA 2: O[chunk][128 queries x 32] += P^T[128 q x 128 keys] V[128 keys x 32], 16 k-steps of 8 key rows
        for (int j = 0; j < kTile / 8; ++j)
          umma_tf32(tm_O + qc * DH, make_desc(smem_u32(pb) + j * 1024, kTileBytes, 512, 1),
                    make_desc(smem_u32(sV) + j * 1024, kTileBytes, 512, 1), idesc2, (kt > 0 || j > 0) ? 1u : 0u);
        umma_commit(&bar_o[buf]);
      }
      __syncwarp();
    }
    if (tid == 0 && kt + 1 < NT) {                    // V tile is free once the last MMA 2 of this key tile has retired
      const int last = n - 1;
      mbar_wait(&bar_o[last & 1], (unsigned)((last >> 1) & 1));
      mbar_expect_tx(bar_v, (unsigned)kTileBytes);
      tma_load_2d(sV, &tmMn, bar_v, 2 * D + h * DH, b * T + (kt + 1) * kTile);
    }
    __syncwarp();
  }
  {
    const int last = n - 1;                            // commits complete in order: the last one covers every MMA
    mbar_wait(&bar_o[last & 1], (unsigned)((last >> 1) & 1));
    tc_fence_after();
  }
  for (int qc = 0; qc < NT; ++qc) {
    const int q = qc * kTile + tid;                    // lanes = queries
    float o[32];
    tmem_ld16(tm_O + lane_off + qc * DH, o);
    tmem_ld16(tm_O + lane_off + qc * DH + 16, o + 16);
    if (q < T) store_row32(p.out, p.out_bf16 != 0, ((size_t)b * T + q) * D + h * DH, o);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------ backward
template <int NT>
__global__ void __launch_bounds__(128, 1)
    attn_tcl_bwd_kernel(const __grid_constant__ CUtensorMap tmKm /* K-major {32,128} over qkv */,
                        const __grid_constant__ CUtensorMap tmMn /* MN-major {32,128} over qkv */,
                        const __grid_constant__ CUtensorMap tmDOk /* K-major {32,128} over dctx */,
                        const __grid_constant__ CUtensorMap tmDOm /* MN-major {32,128} over dctx */,
                        const AttnLongParams p) {
  pdl_entry();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sKk = base;                         // per key tile: K K-major, V K-major, K MN-major
  unsigned char* sVk = sKk + kTileBytes;
  unsigned char* sKm = sVk + kTileBytes;
  unsigned char* sQk = sKm + kTileBytes;             // per query chunk: Q K-major, Q MN-major, dO K-major, dO MN-major
  unsigned char* sQm = sQk + kTileBytes;
  unsigned char* sDOk = sQm + kTileBytes;
  unsigned char* sDOm = sDOk + kTileBytes;
  unsigned char* sY = sDOm + kTileBytes;             // dS^T chunk, q contiguous: 4 slabs x 128 key rows
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sY + 4 * kTileBytes);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);
  unsigned long long* bar_kt = &bars[0];
  unsigned long long* bar_qc = &bars[1];
  unsigned long long* bar_m1 = &bars[2];
  unsigned long long* bar_m2 = &bars[3];
  __shared__ float red[3 * 128];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T, TQ = p.TQ, D = p.H * DH;

  for (int i = tid * 16; i < 4 * kTileBytes; i += 128 * 16) *reinterpret_cast<float4*>(sY + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_async_smem();
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = *tmem_slot;
  const unsigned tm_dV = tmem, tm_dK = tmem + 32, tm_dQ = tmem + 64, tm_S = tmem + 256, tm_dP = tmem + 384;
  const unsigned lane_off = (unsigned)(warp * 32) << 16;
  const unsigned idesc_ts = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((unsigned)(DH >> 3) << 17) |
                            ((unsigned)(128 >> 4) << 24);   // A from TMEM, B MN-major, N = 32
  const unsigned idesc_mm = idesc_ts | (1u << 15);          // A MN-major from shared memory, B MN-major

  float acc_k[32], acc_v[32], acc_q[32];                    // column sums of dK / dV / dQ rows (bias gradient)
#pragma unroll
  for (int j = 0; j < 32; ++j) acc_k[j] = acc_v[j] = acc_q[j] = 0.f;

  unsigned step = 0;                                        // bar_qc / bar_m1 / bar_m2 complete once per (key tile, phase, chunk)
  for (int kt = 0; kt < NT; ++kt) {
    const int kg = kt * kTile + tid;
    const bool valid = kg < T;
    const float rowmask = (valid && __ldg(p.mask + (size_t)b * T + kg) > 0.f) ? 0.f : -1e9f;
    float mxl = 0.f, inv = 0.f;
    if (valid) {
      const float2 st = __ldg(reinterpret_cast<const float2*>(p.stats) + ((size_t)(b * p.H + h) * T + kg));
      mxl = st.x;
      inv = st.y;
    }
    if (tid == 0) {                                         // every MMA that read the previous key tile has retired (bar_m2 waits)
      mbar_expect_tx(bar_kt, (unsigned)(3 * kTileBytes));
      tma_load_2d(sKk, &tmKm, bar_kt, h * DH, b * T + kt * kTile);
      tma_load_2d(sVk, &tmKm, bar_kt, 2 * D + h * DH, b * T + kt * kTile);
      tma_load_2d(sKm, &tmMn, bar_kt, h * DH, b * T + kt * kTile);
    }
    // ---------------- phase 1: dV[keys x 32] = sum over query chunks of P[keys x q] dO[q x 32]
    for (int qc = 0; qc < NT; ++qc, ++step) {
      const unsigned par = step & 1;
      const int nq = min(kTile, TQ - qc * kTile);
      const unsigned idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(nq >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
      if (tid == 0) {
        mbar_expect_tx(bar_qc, (unsigned)(2 * kTileBytes));
        tma_load_2d(sQk, &tmKm, bar_qc, D + h * DH, b * T + qc * kTile);
        tma_load_2d(sDOm, &tmDOm, bar_qc, h * DH, b * T + qc * kTile);
        if (qc == 0) mbar_wait(bar_kt, (unsigned)(kt & 1));
        mbar_wait(bar_qc, par);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < DH / 8; ++k)       // S = K Q^T
          umma_tf32(tm_S, make_desc(smem_u32(sKk) + k * 32, 16, 1024, 2), make_desc(smem_u32(sQk) + k * 32, 16, 1024, 2),
                    idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_m1);
      }
      __syncwarp();
      mbar_wait(bar_m1, par);
      tc_fence_after();
      for (int c = 0; c < nq; c += 16) {
        float v[16];
        tmem_ld16(tm_S + lane_off + c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          v[j] = (qc * kTile + c + j < T)
                     ? to_tf32(exp2f(fmaf(fmaf(v[j], p.inv_scale, rowmask), kLog2e, -mxl)) * inv) : 0.f;
        tmem_st16(tm_S + lane_off + c, v);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        for (int j = 0; j < nq / 8; ++j)       // A = P from TMEM (8 query columns per k-step), B = dO MN-major (+8 query rows)
          umma_tf32_ts(tm_dV, tm_S + j * 8, make_desc(smem_u32(sDOm) + j * 1024, kTileBytes, 512, 1), idesc_ts,
                       (qc > 0 || j > 0) ? 1u : 0u);
        umma_commit(bar_m2);
      }
      __syncwarp();
      mbar_wait(bar_m2, par);                  // S and the chunk tiles are free again
      tc_fence_after();
    }
    // ---------------- delta_k = V_k . dV_k; dV row out
    float delta = 0.f;
    {
      float o[32];
      tmem_ld16(tm_dV + lane_off, o);
      tmem_ld16(tm_dV + lane_off + 16, o + 16);
      // V row of this key from the K-major SWIZZLE_128B tile (as the MMAs saw it: TF32-rounded by TMA)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v4 = *reinterpret_cast<const float4*>(sVk + tid * 128 + ((c ^ (tid & 7)) << 4));
        delta = fmaf(v4.x, o[4 * c], delta);
        delta = fmaf(v4.y, o[4 * c + 1], delta);
        delta = fmaf(v4.z, o[4 * c + 2], delta);
        delta = fmaf(v4.w, o[4 * c + 3], delta);
      }
      if (valid) {
        store_row32(p.out, p.out_bf16 != 0, ((size_t)b * T + kg) * 3 * D + 2 * D + h * DH, o);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc_v[j] += o[j];
      }
    }
    // ---------------- phase 2: dK += dS Q, dQ[chunk] += dS^T K
    for (int qc = 0; qc < NT; ++qc, ++step) {
      const unsigned par = step & 1;
      const int nq = min(kTile, TQ - qc * kTile);
      const unsigned idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(nq >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
      if (tid == 0) {
        mbar_expect_tx(bar_qc, (unsigned)(3 * kTileBytes));
        tma_load_2d(sQk, &tmKm, bar_qc, D + h * DH, b * T + qc * kTile);
        tma_load_2d(sQm, &tmMn, bar_qc, D + h * DH, b * T + qc * kTile);
        tma_load_2d(sDOk, &tmDOk, bar_qc, h * DH, b * T + qc * kTile);
        mbar_wait(bar_qc, par);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < DH / 8; ++k) {     // S = K Q^T and dP = V dO^T, two independent chains
          umma_tf32(tm_S, make_desc(smem_u32(sKk) + k * 32, 16, 1024, 2), make_desc(smem_u32(sQk) + k * 32, 16, 1024, 2),
                    idesc_s, k > 0 ? 1u : 0u);
          umma_tf32(tm_dP, make_desc(smem_u32(sVk) + k * 32, 16, 1024, 2), make_desc(smem_u32(sDOk) + k * 32, 16, 1024, 2),
                    idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(bar_m1);
      }
      __syncwarp();
      mbar_wait(bar_m1, par);
      tc_fence_after();
      for (int c = 0; c < nq; c += 16) {
        float v[16], g[16];
        tmem_ld16(tm_S + lane_off + c, v);
        tmem_ld16(tm_dP + lane_off + c, g);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float pr = (qc * kTile + c + j < T) ? exp2f(fmaf(fmaf(v[j], p.inv_scale, rowmask), kLog2e, -mxl)) * inv : 0.f;
          g[j] = to_tf32(pr * ((g[j] - delta) * p.inv_scale));      // dS
        }
        tmem_st16(tm_dP + lane_off + c, g);
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(sY + mn_major_off(c + j, tid, kTile)) = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
      }
      tmem_st_wait();
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        for (int j = 0; j < nq / 8; ++j)       // dK[keys x 32] += dS[keys x q] Q[q x 32]: A from TMEM, B = Q MN-major
          umma_tf32_ts(tm_dK, tm_dP + j * 8, make_desc(smem_u32(sQm) + j * 1024, kTileBytes, 512, 1), idesc_ts,
                       (qc > 0 || j > 0) ? 1u : 0u);
        for (int j = 0; j < kTile / 8; ++j)    // dQ[chunk][q x 32] += dS^T[q x keys] K[keys x 32]: A = Y, B = K MN-major
          umma_tf32(tm_dQ + qc * DH, make_desc(smem_u32(sY) + j * 1024, kTileBytes, 512, 1),
                    make_desc(smem_u32(sKm) + j * 1024, kTileBytes, 512, 1), idesc_mm, (kt > 0 || j > 0) ? 1u : 0u);
        umma_commit(bar_m2);
      }
      __syncwarp();
      mbar_wait(bar_m2, par);
      tc_fence_after();
    }
    {
      float o[32];
      tmem_ld16(tm_dK + lane_off, o);
      tmem_ld16(tm_dK + lane_off + 16, o + 16);
      if (valid) {
        store_row32(p.out, p.out_bf16 != 0, ((size_t)b * T + kg) * 3 * D + h * DH, o);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc_k[j] += o[j];
      }
    }
    tc_fence_before();
    __syncthreads();                           // every thread has read dV / dK before the next key tile overwrites them
    tc_fence_after();
  }
  for (int qc = 0; qc < NT; ++qc) {
    const int q = qc * kTile + tid;            // lanes = queries
    float o[32];
    tmem_ld16(tm_dQ + lane_off + qc * DH, o);
    tmem_ld16(tm_dQ + lane_off + qc * DH + 16, o + 16);
    if (q < T) {
      store_row32(p.out, p.out_bf16 != 0, ((size_t)b * T + q) * 3 * D + D + h * DH, o);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc_q[j] += o[j];
    }
  }
  if (p.dbias) {
    red[0 * 128 + tid] = warp_colsum32(acc_k, tid & 31);
    red[1 * 128 + tid] = warp_colsum32(acc_q, tid & 31);
    red[2 * 128 + tid] = warp_colsum32(acc_v, tid & 31);
    __syncthreads();
    if (tid < 96) {
      const int m = tid >> 5, c = tid & 31;
      atomicAdd(p.dbias + m * D + h * DH + c, red[m * 128 + c] + red[m * 128 + 32 + c] + red[m * 128 + 64 + c] + red[m * 128 + 96 + c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

constexpr size_t fwd_smem(int nt) { return 1024 + (size_t)(nt + 2 + 8) * kTileBytes + 128; }
constexpr size_t kBwdSmem = 1024 + (size_t)(7 + 4) * kTileBytes + 128;

}  // namespace

extern "C" int msx_attention_tcl_supported(const float* qkv, int T, int dh) {
  return (qkv && dh == 32 && T > 128 && T <= 3 * kTile && ((uintptr_t)qkv & 15) == 0) ? 1 : 0;
}

extern "C" int msx_attention_tcl_fwd(const float* qkv, const float* mask, void* ctx, int ctx_bf16, float* stats, int B, int T,
                                     int H, int dh, void* stream) {
  MSX_REQUIRE(qkv && mask && ctx && stats, "msx_attention_tcl_fwd: null pointer");
  MSX_REQUIRE(msx_attention_tcl_supported(qkv, T, dh) && ((uintptr_t)ctx & 15) == 0 && ((uintptr_t)stats & 7) == 0,
              "msx_attention_tcl_fwd: needs d_h == 32, 128 < T <= 384, 16-byte aligned buffers");
  if (B == 0) return MSX_OK;
  const int D = H * DH;
  AttnLongParams p;
  p.mask = mask; p.stats = stats; p.out = ctx; p.out_bf16 = ctx_bf16 ? 1 : 0; p.dbias = nullptr;
  p.T = T; p.H = H; p.TQ = (T + 15) / 16 * 16; p.inv_scale = 1.f / sqrtf((float)DH);
  const long long rows = (long long)B * T;
  CUtensorMap tk, tm;
  int rc;
  if ((rc = make_map(&tk, qkv, rows, 3 * D, 3 * D, DH, kTile, false))) return rc;
  if ((rc = make_map(&tm, qkv, rows, 3 * D, 3 * D, DH, kTile, true))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (T <= 2 * kTile) {
    MSX_CUDA(cudaFuncSetAttribute(attn_tcl_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem(2)));
    MSX_CUDA(msx_launch(attn_tcl_fwd_kernel<2>, dim3(B * H), dim3(128), fwd_smem(2), st, tk, tm, p));
  } else {
    MSX_CUDA(cudaFuncSetAttribute(attn_tcl_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem(3)));
    MSX_CUDA(msx_launch(attn_tcl_fwd_kernel<3>, dim3(B * H), dim3(128), fwd_smem(3), st, tk, tm, p));
  }
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_attention_tcl_bwd(const float* qkv, const float* mask, const float* dctx, const float* stats, void* dqkv,
                                     int dqkv_bf16, float* dbias, int B, int T, int H, int dh, void* stream) {
  MSX_REQUIRE(qkv && mask && dctx && stats && dqkv, "msx_attention_tcl_bwd: null pointer");
  MSX_REQUIRE(msx_attention_tcl_supported(qkv, T, dh) && ((uintptr_t)dctx & 15) == 0 && ((uintptr_t)dqkv & 15) == 0 &&
                  ((uintptr_t)stats & 7) == 0,
              "msx_attention_tcl_bwd: needs d_h == 32, 128 < T <= 384, 16-byte aligned buffers");
  if (B == 0) return MSX_OK;
  const int D = H * DH;
  AttnLongParams p;
  p.mask = mask; p.stats = const_cast<float*>(stats); p.out = dqkv; p.out_bf16 = dqkv_bf16 ? 1 : 0; p.dbias = dbias;
  p.T = T; p.H = H; p.TQ = (T + 15) / 16 * 16; p.inv_scale = 1.f / sqrtf((float)DH);
  const long long rows = (long long)B * T;
  CUtensorMap tk, tm, tdk, tdm;
  int rc;
  if ((rc = make_map(&tk, qkv, rows, 3 * D, 3 * D, DH, kTile, false))) return rc;
  if ((rc = make_map(&tm, qkv, rows, 3 * D, 3 * D, DH, kTile, true))) return rc;
  if ((rc = make_map(&tdk, dctx, rows, D, D, DH, kTile, false))) return rc;
  if ((rc = make_map(&tdm, dctx, rows, D, D, DH, kTile, true))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (T <= 2 * kTile) {
    MSX_CUDA(cudaFuncSetAttribute(attn_tcl_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
    MSX_CUDA(msx_launch(attn_tcl_bwd_kernel<2>, dim3(B * H), dim3(128), kBwdSmem, st, tk, tm, tdk, tdm, p));
  } else {
    MSX_CUDA(cudaFuncSetAttribute(attn_tcl_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
    MSX_CUDA(msx_launch(attn_tcl_bwd_kernel<3>, dim3(B * H), dim3(128), kBwdSmem, st, tk, tm, tdk, tdm, p));
  }
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
