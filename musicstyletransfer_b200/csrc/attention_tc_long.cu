// K2c (tensor-core path, long rows) — attention in the reference's convention on tcgen05 for 128 < T <= 768, d_h = 32:
// the L = 128 / 256 points of the BASELINE config 3 sweep (T = 129 / 257), which attention_tc.cu (one key tile on the
// 128 TMEM lanes) does not take.
//
// Same contract as attention.cu / attention_tc.cu (replaces MultiHeadDotAttention.hybrid_forward lines 91-103 and
// _mask_logits, /root/reference/music_style_transfer/VarAutoEncoder/transformer.py:91-126):
//   S[k][q] = K_k . Q_q / sqrt(d_h) + (key k padded ? -1e9 : 0);  P = softmax over the QUERY axis;  O[q] = sum_k P[k][q] V[k]
//
// The softmax normalises each KEY row over all queries, so key tiles are independent: one CTA (256 threads) per
// (batch, head) walks the key tiles of 128; for a key tile the scores of ALL queries accumulate in TMEM with the keys
// on the lanes (S[128 x TQ], up to 384 columns), the TWO threads that own a key row (warps w and w + 4 reach the same
// TMEM lane quarter) do max / exp2 / sum over their halves of its TMEM columns, and the normalised row goes out in
// query chunks of 128 as the MN-major A operand of O[chunk] += P^T V (double-buffered in shared memory), O accumulating
// over key tiles in TMEM.  Both kernels are bound by the instruction stream of the row owners (ncu: 18 % issue utilisation
// with 128 threads, stall samples spread evenly over a straight-line stream), hence two threads per row, the single-tile
// specialisations and the q0_only variants below; measurements in profiles/micro/attn_long_rate_r2.txt.  The forward saves (max * log2 e,
// 1 / sum) per key row; the backward rebuilds P from them chunk by chunk (flash-attention style), so it never needs
// more than 128 score columns at a time:
//   phase 1 (per key tile, over query chunks):  P -> TMEM,  dV += P dO            (A operand from TMEM)
//   delta_k = V_k . dV_k                        (= sum_q P dP, thread-local: dV has the keys on the lanes)
//   phase 2 (over query chunks):  dP = V dO^T,  dS = P (dP - delta) / sqrt(d_h) -> TMEM and, transposed, shared memory
//                                 dK += dS Q (A from TMEM),  dQ[chunk] += dS^T K (A MN-major from shared memory)
// dQ accumulates over key tiles in TMEM (32 columns per query chunk).
#include <cuda_bf16.h>
#include <cstdlib>
#include "attention_tc_common.cuh"

namespace {

constexpr int kTile = 128;                  // keys per key tile = queries per query chunk
constexpr int kTileBytes = kTile * 128;     // one [128 rows][128 B] operand tile (K-major or MN-major)

// Trailing key rows.  The reference's rows are T = L + 1 long (SOS / EOS, VarAutoEncoder/data.py:150-170), so the sweep's
// row lengths are 128 n + 1: a key tile of their own would cost a full tile pass for ONE key.  Up to kTailMax trailing keys
// (T = 128 n + tail) are taken off the tensor path instead: a key row's scores against every query are 32-long dot products,
// one or two per thread (thread = query), its softmax over the query axis two block reductions, and it enters the outputs
// as rank-1 updates in the epilogue (forward: O[q] += P[k*][q] V[k*]; backward: dQ[q] += dS[k*][q] K[k*], with dV[k*] / dK[k*]
// block column sums).  Operands are rounded to TF32 exactly where the tensor path rounds them (K, Q of the scores), so the
// forward's saved statistics and the backward's recomputed P agree bit for bit.
//
// The same trailing positions are taken off the tensor path as QUERIES: the thread that owns key row k adds the scores
// K_k . Q_q* of the trailing queries (32-long dot products, thread-local) to its row maximum / sum, and the trailing
// queries' outputs (forward: O[q*] = sum_k P[k][q*] V[k]; backward: dQ[q*] = sum_k dS[k][q*] K[k]) are block column sums,
// while dV_k += P[k][q*] dO[q*] and dK_k += dS[k][q*] Q[q*] are thread-local rank-1 updates of the key thread's rows.  So
// T = 128 n + tail runs n x n tiles instead of (n + 1) x (n + 1); T = 129 ... 132 is ONE tile: the forward then aliases
// the P staging onto the (dead) Q / K tiles and allocates 256 TMEM columns, so two CTAs share an SM, and the backward
// loads its seven operand tiles once and issues S = K Q^T and dP = V dO^T together.
constexpr int kTailMax = 4;
__host__ __device__ inline int tail_keys(int T) {
  const int t = T % 128;
  return (t >= 1 && t <= kTailMax) ? t : 0;
}

struct AttnLongParams {
  const float* qkv;    // [B*T, 3*H*32]: the tail-key path reads rows directly
  const float* dctx;   // bwd: [B*T, H*32]
  const float* mask;   // [B*T]
  float* stats;        // [B*H*T, 2]: (row max * log2 e, 1 / row sum) of every key row; written by fwd, read by bwd
  void* out;           // fwd: ctx [B*T, H*32]; bwd: dqkv [B*T, 3*H*32]; fp32, or bf16 when out_bf16
  void* out_lo;        // fwd, optional (out_bf16): the context as bf16 hi (out) / lo planes, the operands of the p3 W_proj GEMM
  int out_bf16;
  float* dbias;        // bwd, optional [3*H*32]: += column sums of dqkv
  int T, H, TQ;        // TQ = score columns on the tensor path: roundup16(T), or T - tail when trailing positions leave it
  float inv_scale;
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32-long dot product of a K row (shared memory, the same address for every thread: a broadcast; rounded to TF32 here) with a query row
__device__ __forceinline__ float dot32_tf32(const float* __restrict__ krow, const float (&q)[32]) {
  float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 k4 = *(reinterpret_cast<const float4*>(krow) + c);
    d4[0] = fmaf(to_tf32(k4.x), q[4 * c], d4[0]);
    d4[1] = fmaf(to_tf32(k4.y), q[4 * c + 1], d4[1]);
    d4[2] = fmaf(to_tf32(k4.z), q[4 * c + 2], d4[2]);
    d4[3] = fmaf(to_tf32(k4.w), q[4 * c + 3], d4[3]);
  }
  return (d4[0] + d4[1]) + (d4[2] + d4[3]);
}

// same accumulation order as dot32_tf32, both rows in registers
__device__ __forceinline__ float dot32(const float (&a)[32], const float (&b)[32]) {
  float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    d4[0] = fmaf(a[4 * c], b[4 * c], d4[0]);
    d4[1] = fmaf(a[4 * c + 1], b[4 * c + 1], d4[1]);
    d4[2] = fmaf(a[4 * c + 2], b[4 * c + 2], d4[2]);
    d4[3] = fmaf(a[4 * c + 3], b[4 * c + 3], d4[3]);
  }
  return (d4[0] + d4[1]) + (d4[2] + d4[3]);
}
// row r of a K-major SWIZZLE_128B tile [128 rows][32 floats] (as TMA wrote it: TF32-rounded)
__device__ __forceinline__ void load_row_km(const unsigned char* tile, int r, float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 x = *reinterpret_cast<const float4*>(tile + r * 128 + ((c ^ (r & 7)) << 4));
    v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
  }
}
// row r (one key's 32 head columns) of an MN-major SWIZZLE_128B_ATOM_32B tile (mn_major_off with mn = column, k = r)
__device__ __forceinline__ void load_row_mn(const unsigned char* tile, int r, float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 x = *reinterpret_cast<const float4*>(tile + r * 128 + (((c >> 1) ^ (r & 3)) << 5) + ((c & 1) << 4));
    v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
  }
}
// a 32-float row held in shared memory (the same address for every thread: a broadcast)
__device__ __forceinline__ void load_row32(const float* row, float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 x = *(reinterpret_cast<const float4*>(row) + c);
    v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
  }
}
// The rows of the trailing positions, fetched once per (batch, head) into shared memory when the kernel starts (their
// DRAM / L2 latency hides behind the first TMA loads): [0] Q rounded to TF32, [1] Q, [2] dO (backward), [3] K, [4] V
enum { kTrQt = 0, kTrQ = 1, kTrDO = 2, kTrK = 3, kTrV = 4 };
template <int NWARPS>
__device__ __forceinline__ void fetch_tail_rows(float (*tr)[kTailMax][32], const float* qkv, const float* dctx, size_t row0, int ntail,
                                                int D, int h, int warp, int lane) {
  for (int r = warp; r < 5 * ntail; r += NWARPS) {
    const int ty = r / ntail, i = r - ty * ntail;
    const size_t pos = row0 + i;
    float v = 0.f;
    if (ty == kTrQt) v = to_tf32(__ldg(qkv + pos * 3 * D + D + h * DH + lane));
    else if (ty == kTrQ) v = __ldg(qkv + pos * 3 * D + D + h * DH + lane);
    else if (ty == kTrDO) v = dctx ? __ldg(dctx + pos * D + h * DH + lane) : 0.f;
    else if (ty == kTrK) v = __ldg(qkv + pos * 3 * D + h * DH + lane);
    else v = __ldg(qkv + pos * 3 * D + 2 * D + h * DH + lane);
    tr[ty][i][lane] = v;
  }
}

constexpr int kBwdThreads = 256, kFwdThreads = 256;

__device__ __forceinline__ float block_sum256(float v, float* red8, int tid) {
  v = warp_sum(v);
  if ((tid & 31) == 0) red8[tid >> 5] = v;
  __syncthreads();
  v = ((red8[0] + red8[1]) + (red8[2] + red8[3])) + ((red8[4] + red8[5]) + (red8[6] + red8[7]));
  __syncthreads();
  return v;
}
// column sums of a [256 threads x 32] register matrix: afterwards every thread holds the total of column (tid & 31)
__device__ __forceinline__ float block_colsum256(float (&v)[32], float* red /* [256] */, int tid) {
  red[tid] = warp_colsum32(v, tid & 31);
  __syncthreads();
  const int c = tid & 31;
  const float t = ((red[c] + red[32 + c]) + (red[64 + c] + red[96 + c])) + ((red[128 + c] + red[160 + c]) + (red[192 + c] + red[224 + c]));
  __syncthreads();
  return t;
}
// block_colsum256 and block_sum256 of a second value behind one pair of barriers
__device__ __forceinline__ float block_colsum256_sum(float (&v)[32], float x, float* red /* [256] */, float* red8, int tid, float* xsum) {
  red[tid] = warp_colsum32(v, tid & 31);
  x = warp_sum(x);
  if ((tid & 31) == 0) red8[tid >> 5] = x;
  __syncthreads();
  const int c = tid & 31;
  const float t = ((red[c] + red[32 + c]) + (red[64 + c] + red[96 + c])) + ((red[128 + c] + red[160 + c]) + (red[192 + c] + red[224 + c]));
  *xsum = ((red8[0] + red8[1]) + (red8[2] + red8[3])) + ((red8[4] + red8[5]) + (red8[6] + red8[7]));
  __syncthreads();
  return t;
}
// the same over the 128 threads of each half of the CTA separately (threads 0 .. 127 and 128 .. 255 hold different matrices)
__device__ __forceinline__ float half_colsum128(float (&v)[32], float* red /* [256] */, int tid) {
  red[tid] = warp_colsum32(v, tid & 31);
  __syncthreads();
  const int c = (tid & 128) + (tid & 31);
  const float t = (red[c] + red[32 + c]) + (red[64 + c] + red[96 + c]);
  __syncthreads();
  return t;
}
// every lane holds v[0..15]; afterwards lane L holds the sum over all lanes of v[L & 15]
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float recv = __shfl_xor_sync(0xffffffffu, send, s);
      v[i] = (up ? v[i + s] : v[i]) + recv;
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}
// o += a * (16 consecutive floats of a row held in shared memory; the same address for the threads of a warp: a broadcast)
__device__ __forceinline__ void axpy16(float a, const float* row16, float (&o)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 x = *(reinterpret_cast<const float4*>(row16) + c);
    o[4 * c] = fmaf(a, x.x, o[4 * c]); o[4 * c + 1] = fmaf(a, x.y, o[4 * c + 1]);
    o[4 * c + 2] = fmaf(a, x.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(a, x.w, o[4 * c + 3]);
  }
}

__device__ __forceinline__ float block_max256(float v, float* red8, int tid) {
  v = warp_max(v);
  if ((tid & 31) == 0) red8[tid >> 5] = v;
  __syncthreads();
  v = fmaxf(fmaxf(fmaxf(red8[0], red8[1]), fmaxf(red8[2], red8[3])), fmaxf(fmaxf(red8[4], red8[5]), fmaxf(red8[6], red8[7])));
  __syncthreads();
  return v;
}

// 16 columns of a context row: fp32, bf16, or bf16 hi / lo planes (hi = rn_bf16(o), lo = rn_bf16(o - hi))
__device__ __forceinline__ void store_ctx16(const AttnLongParams& p, size_t elem, const float* o) {
  if (p.out_lo) store_row32_planes(p.out, p.out_lo, elem, o, 16);
  else store_row32(p.out, p.out_bf16 != 0, elem, o, 16);
}
__device__ __forceinline__ void store_ctx1(const AttnLongParams& p, size_t elem, float o) {
  if (p.out_bf16) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(o);
    reinterpret_cast<unsigned short*>(p.out)[elem] = __bfloat16_as_ushort(hi);
    if (p.out_lo) reinterpret_cast<unsigned short*>(p.out_lo)[elem] = __bfloat16_as_ushort(__float2bfloat16_rn(o - __bfloat162float(hi)));
  } else {
    reinterpret_cast<float*>(p.out)[elem] = o;
  }
}

// ------------------------------------------------------------------------------------------------ forward
// 256 threads: warps w and w + 4 share the TMEM lane quarter w & 3, i.e. two threads own one key row (one query row in the
// epilogue) and split its columns — the two 64-column halves of every 128-wide query chunk in the three softmax passes
// (partial row maxima / sums meet in shared memory), 16 + 16 head columns in the output rows.  As in the backward, the
// kernel is bound by the instruction stream of the row owners, not by the tensor pipe or HBM.
// NKT = key tiles = query chunks on the tensor path (NT, or NT - 1 with trailing positions).
// ALIAS (NT = 2, NKT = 1: T = 129 ... 132, one key tile x one query chunk on the tensor path): P staging (single buffer)
// lies over the Q / K tiles, 256 TMEM columns, 81 KB of shared memory: two CTAs per SM.
// Q0 (the encoder's top layer under SOS-rows-only, model.py:97-100 reads position 0): scores and the query-axis softmax
// cover every query as before (each key row's normaliser), then O[0] = sum_k P[k][0] V[k] is a column sum over the key
// threads — no P staging, no MMA 2, one 128-byte row written per (batch, head); the other rows of ctx stay untouched.
// Q0 needs neither the P staging nor the O accumulators: with at most 256 score columns (NKT <= 2) it allocates 256 TMEM
// columns and (NT + 2) tiles of shared memory, so two CTAs share an SM (T = 257: 628 -> see attn_long_rate_r2.txt).
template <int NT, int NKT, bool ALIAS, bool Q0>
__global__ void __launch_bounds__(kFwdThreads, (ALIAS || (Q0 && NKT <= 2)) ? 2 : 1)
    attn_tcl_fwd_kernel(const __grid_constant__ CUtensorMap tmKm /* K-major {32, 128} box over qkv */,
                        const __grid_constant__ CUtensorMap tmMn /* MN-major {32, 128} box over qkv */,
                        const AttnLongParams p) {
  pdl_entry();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sQ = base;                          // NT K-major tiles: all queries of the (batch, head)
  unsigned char* sK = sQ + NT * kTileBytes;          // K-major key tile
  unsigned char* sV = ALIAS ? base + 4 * kTileBytes : sK + kTileBytes;   // MN-major value tile (d contiguous, 128 key rows)
  // P chunk, q contiguous, 4 slabs of 128 key rows; two buffers, or ONE lying over sQ / sK (dead once MMA 1 has retired)
  unsigned char* sP = ALIAS ? base : sV + kTileBytes;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>((ALIAS || Q0) ? sV + kTileBytes : sP + 2 * 4 * kTileBytes);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);
  unsigned long long* bar_q = &bars[0];
  unsigned long long* bar_k = &bars[1];
  unsigned long long* bar_v = &bars[2];
  unsigned long long* bar_s = &bars[3];
  unsigned long long* bar_o = &bars[4];              // [2]
  constexpr bool Q0SMALL = Q0 && !ALIAS && NKT <= 2;   // S (<= 256 columns) is all the TMEM the kernel needs
  constexpr int kTmemCols = (ALIAS || Q0SMALL) ? 256 : 512;
  static_assert(NKT == NT || NKT == NT - 1, "at most one chunk of trailing positions");
  static_assert(!ALIAS || (NT == 2 && NKT == 1), "ALIAS: one full tile + trailing positions");
  constexpr bool TAIL = NKT < NT;
  constexpr int NQQ = (NT * kTile + kFwdThreads - 1) / kFwdThreads;    // queries per thread in the trailing-key section

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int row = quarter * 32 + lane;               // key row of the tile (query row of the chunk in the epilogue)
  const int c0 = half * 64, h16 = half * 16;         // this thread's columns of a query chunk / of a head row
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T, TQ = p.TQ, D = p.H * DH;
  const int ntail = TAIL ? tail_keys(T) : 0;         // trailing positions handled off the tensor path (see kTailMax)
  __shared__ float xs[2][kTile];                     // partial row maxima, then partial row sums, of the two column halves
  __shared__ float red8[8];
  __shared__ float red[TAIL ? 256 : 1];
  __shared__ float red_q[Q0 ? 256 : 1];
  __shared__ float ot_s[kTailMax][32];               // output rows of the trailing queries
  __shared__ float pt_s[TAIL ? kTailMax : 1][TAIL ? NT * kTile : 1];   // P[trailing key][query]
  __shared__ __align__(16) float tr_s[TAIL ? 5 : 1][kTailMax][32];     // rows of the trailing positions (fetch_tail_rows)
  if (TAIL) fetch_tail_rows<kFwdThreads / 32>(tr_s, p.qkv, nullptr, (size_t)b * T + (size_t)(NT - 1) * kTile, ntail, D, h, warp, lane);

  if (!TAIL && !Q0) {
    // P staging must hold finite values everywhere the MMAs read (short last query chunk): zero it once
    for (int i = tid * 16; i < 2 * 4 * kTileBytes; i += kFwdThreads * 16) *reinterpret_cast<float4*>(sP + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_async_smem();
  }
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = *tmem_slot;
  const unsigned tm_O = tmem, tm_S = Q0SMALL ? tmem : tmem + 128;   // O: 32 columns per query chunk; S: up to 384 columns (ALIAS: 128)
  const unsigned lane_off = (unsigned)(quarter * 32) << 16;
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
  float o0 = 0.f, e0 = 0.f;                           // Q0: O[0][column lane] summed over the key tiles; exp of (key row, query 0)
  float ot[kTailMax / 2];                             // O[trailing query 2 ii + half][column lane], summed over the key tiles
#pragma unroll
  for (int i = 0; i < kTailMax / 2; ++i) ot[i] = 0.f;
  int n = 0;                                          // running (key tile, query chunk) counter: P buffer = n & 1
#pragma unroll 1
  for (int kt = 0; kt < NKT; ++kt) {
    const unsigned par = (unsigned)(kt & 1);
    if (tid == 0) {
      if (kt == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_k, par);
      tc_fence_after();
      // MMA 1: S[128 keys x TQ] = K Q^T in query chunks of <= 128 columns (both operands K-major, +32 B per k-step)
      for (int qc = 0; qc < NKT; ++qc) {
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
    const int kg = kt * kTile + row;                  // this thread's key row
    const bool valid = kg < T;
    const float rowmask = (valid && __ldg(p.mask + (size_t)b * T + kg) > 0.f) ? 0.f : -1e9f;
    // while MMA 1 runs: this key row's scores against the trailing queries (rows 0 .. ntail-1 of the last Q tile; both
    // threads of the row)
    float st[kTailMax];
#pragma unroll
    for (int i = 0; i < kTailMax; ++i) st[i] = 0.f;
    if (TAIL) {
      if (kt == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_k, par);
      {
        float kr[32];
        load_row_km(sK, row, kr);
#pragma unroll
        for (int i = 0; i < kTailMax; ++i) {
          if (i < ntail) {
            float qr[32];
            load_row_km(sQ + (NT - 1) * kTileBytes, i, qr);
            st[i] = fmaf(dot32(kr, qr), p.inv_scale, rowmask);
          }
        }
      }
      if (kt == 0) {
        // once per item, the trailing KEY rows (thread = query; independent of the tensor path): scores of key k* against
        // this thread's queries (rows of the resident, TF32-rounded Q tiles; K[k*] from the fetch at kernel start), softmax
        // over the query axis across the CTA, P[k*][q] kept for the rank-1 update of the output rows
#pragma unroll 1
        for (int i = 0; i < ntail; ++i) {
          const int ks = (NT - 1) * kTile + i;
          const float rmask = __ldg(p.mask + (size_t)b * T + ks) > 0.f ? 0.f : -1e9f;
          float sc[NQQ];
          float mx = -INFINITY;
#pragma unroll
          for (int qq = 0; qq < NQQ; ++qq) {
            const int q = qq * kFwdThreads + tid;
            sc[qq] = 0.f;
            if (q < T) {
              float qr[32];
              load_row_km(sQ + (q >> 7) * kTileBytes, q & 127, qr);
              sc[qq] = fmaf(dot32_tf32(tr_s[kTrK][i], qr), p.inv_scale, rmask);
              mx = fmaxf(mx, sc[qq]);
            }
          }
          const float mxl = block_max256(mx, red8, tid) * kLog2e;
          float sum = 0.f;
#pragma unroll
          for (int qq = 0; qq < NQQ; ++qq) {
            sc[qq] = (qq * kFwdThreads + tid < T) ? exp2f(fmaf(sc[qq], kLog2e, -mxl)) : 0.f;
            sum += sc[qq];
          }
          const float inv = 1.f / block_sum256(sum, red8, tid);
#pragma unroll
          for (int qq = 0; qq < NQQ; ++qq)
            if (qq * kFwdThreads + tid < T) pt_s[i][qq * kFwdThreads + tid] = sc[qq] * inv;
          if (tid == 0) reinterpret_cast<float2*>(p.stats)[(size_t)(b * p.H + h) * T + ks] = make_float2(mxl, inv);
        }
      }
      __syncthreads();                                // every thread has read its K / Q rows: the tiles may be reused
    }
    mbar_wait(bar_s, par);
    tc_fence_after();
    if (tid == 0 && kt + 1 < NKT) {                   // the K tile is free once MMA 1 has retired
      mbar_expect_tx(bar_k, (unsigned)kTileBytes);
      tma_load_2d(sK, &tmKm, bar_k, h * DH, b * T + (kt + 1) * kTile);
    }
    // pass 1: row maximum over this thread's half of every query chunk
    float mx = -INFINITY;
#pragma unroll 1
    for (int qc = 0; qc < NKT; ++qc) {
      const int cend = qc * kTile + min(min(kTile, TQ - qc * kTile), c0 + 64);
      for (int c = qc * kTile + c0; c < cend; c += 16) {
        float v[16];
        tmem_ld16(tm_S + lane_off + c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (TAIL || c + j < T) mx = fmaxf(mx, fmaf(v[j], p.inv_scale, rowmask));
      }
    }
    if (TAIL && half == 0) {
#pragma unroll
      for (int i = 0; i < kTailMax; ++i)
        if (i < ntail) mx = fmaxf(mx, st[i]);
    }
    xs[half][row] = mx;
    __syncthreads();
    const float mxl = fmaxf(xs[0][row], xs[1][row]) * kLog2e;
    __syncthreads();
    // pass 2: e = exp2(s * log2 e - max * log2 e) back into TMEM, row sum
    float sum = 0.f;
#pragma unroll 1
    for (int qc = 0; qc < NKT; ++qc) {
      const int cend = qc * kTile + min(min(kTile, TQ - qc * kTile), c0 + 64);
      for (int c = qc * kTile + c0; c < cend; c += 16) {
        float v[16];
        tmem_ld16(tm_S + lane_off + c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] = (TAIL || c + j < T) ? exp2f(fmaf(fmaf(v[j], p.inv_scale, rowmask), kLog2e, -mxl)) : 0.f;
          sum += v[j];
        }
        if (Q0) {
          if (c == 0) e0 = v[0];                        // query 0: first column of the first thread of the row
        } else {
          tmem_st16(tm_S + lane_off + c, v);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kTailMax; ++i) {
      if (i < ntail) {
        st[i] = exp2f(fmaf(st[i], kLog2e, -mxl));
        if (half == 0) sum += st[i];
      }
    }
    tmem_st_wait();
    xs[half][row] = sum;
    __syncthreads();
    const float inv = valid ? 1.f / (xs[0][row] + xs[1][row]) : 0.f;   // rows beyond T (neighbouring sequence / zeros) contribute nothing
    if (valid && half == 0) {
      float2* stp = reinterpret_cast<float2*>(p.stats) + ((size_t)(b * p.H + h) * T + kg);
      *stp = make_float2(mxl, inv);
    }
    if (Q0) {
      // O[0] += sum over this tile's keys of P[k][0] V[k] (the second threads of the rows add zeros)
      mbar_wait(bar_v, par);
      float c[32];
      load_row_mn(sV, row, c);
      const float p0 = (half == 0) ? e0 * inv : 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) c[j] *= p0;
      o0 += half_colsum128(c, red_q, tid);              // ends with a barrier: every thread has read its V row
      if (tid == 0 && kt + 1 < NKT) {
        mbar_expect_tx(bar_v, (unsigned)kTileBytes);
        tma_load_2d(sV, &tmMn, bar_v, 2 * D + h * DH, b * T + (kt + 1) * kTile);
      }
    }
    // pass 3: normalised row -> shared memory, one query chunk at a time, each chunk followed by its MMA 2
#pragma unroll 1
    for (int qc = 0; qc < (Q0 ? 0 : NKT); ++qc, ++n) {
      const int buf = ALIAS ? 0 : (n & 1);
      unsigned char* pb = sP + buf * 4 * kTileBytes;
      if (n >= 2) {                                    // the MMA that read this buffer two chunks ago has retired
        mbar_wait(&bar_o[buf], (unsigned)(((n >> 1) - 1) & 1));
        tc_fence_after();
      }
      const int cend = min(min(kTile, TQ - qc * kTile), c0 + 64);
      for (int c = c0; c < cend; c += 16) {
        float v[16];
        tmem_ld16(tm_S + lane_off + qc * kTile + c, v);
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(pb + mn_major_off(c + j, row, kTile)) =
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
    if (TAIL && !Q0) {
      // while the MMA 2s run: O[q*] += sum over this tile's keys of P[k][q*] V[k] (V row as the MMA sees it); the first
      // threads of the rows take the even trailing queries, the second threads the odd ones
      mbar_wait(bar_v, par);
      float vr[32];
      load_row_mn(sV, row, vr);
#pragma unroll
      for (int ii = 0; ii < kTailMax / 2; ++ii) {
        if (2 * ii < ntail) {                          // CTA-uniform
          const float pt = ((half == 0) ? st[2 * ii] : st[2 * ii + 1]) * inv;      // 0 beyond ntail
          float c[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) c[j] = pt * vr[j];
          ot[ii] += half_colsum128(c, red, tid);        // ends with a barrier: every thread has read its V row
        }
      }
    }
    if (!Q0 && tid == 0 && kt + 1 < NKT) {            // V tile is free once the last MMA 2 of this key tile has retired
      const int last = n - 1;
      mbar_wait(&bar_o[last & 1], (unsigned)((last >> 1) & 1));
      mbar_expect_tx(bar_v, (unsigned)kTileBytes);
      tma_load_2d(sV, &tmMn, bar_v, 2 * D + h * DH, b * T + (kt + 1) * kTile);
    }
    __syncthreads();                                  // xs is rewritten by the next key tile
  }
  if (Q0) {
    if (warp == 0) {                                   // the row of query 0: column `lane`
      if (TAIL) {
#pragma unroll
        for (int i = 0; i < kTailMax; ++i)             // O[0] += P[k*][0] V[k*]
          if (i < ntail) o0 = fmaf(pt_s[i][0], tr_s[kTrV][i][lane], o0);
      }
      const size_t e = (size_t)b * T * D + h * DH + lane;
      store_ctx1(p, e, o0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
    return;
  }
  {
    const int last = n - 1;                            // commits complete in order: the last one covers every MMA
    mbar_wait(&bar_o[ALIAS ? 0 : (last & 1)], (unsigned)((last >> 1) & 1));
    tc_fence_after();
  }
  if (TAIL) {
    if (quarter == 0) {
#pragma unroll
      for (int ii = 0; ii < kTailMax / 2; ++ii) ot_s[2 * ii + half][lane] = ot[ii];
    }
    __syncthreads();
  }
#pragma unroll
  for (int qc = 0; qc < NT; ++qc) {
    const int q = qc * kTile + row;                    // lanes = queries
    float o[16];
    if (qc < NKT) {
      tmem_ld16(tm_O + lane_off + qc * DH + h16, o);
    } else {                                           // the trailing queries' rows: the threads of rows 0 .. ntail-1
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = ot_s[row & (kTailMax - 1)][h16 + j];
    }
    if (TAIL && q < T) {
#pragma unroll
      for (int i = 0; i < kTailMax; ++i)               // O[q] += P[k*][q] V[k*]
        if (i < ntail) axpy16(pt_s[i][q], tr_s[kTrV][i] + h16, o);
    }
    if (q < T) store_ctx16(p, ((size_t)b * T + q) * D + h * DH + h16, o);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

// ------------------------------------------------------------------------------------------------ forward, two sweeps
// 384 < T <= 768 (NT = 4 ... 6 query chunks): the scores of a key tile against ALL queries no longer fit TMEM and the Q
// tiles no longer fit shared memory next to the P staging.  Per key tile the query chunks are walked twice, one chunk of
// 128 score columns at a time (Q chunks through a two-deep ring, the next one in flight while a chunk is processed;
// MMA 1 is issued again in the second sweep — four small MMAs per chunk):
//   sweep 1: running row maximum / sum over the chunks (online softmax) -> the statistics of the key rows
//   sweep 2: P = exp2(s - max) / sum -> shared memory -> O[chunk] += P^T V      (Q0: one column sum for query 0 instead)
// Thread layout, trailing positions and epilogue as in attn_tcl_fwd_kernel.
// Q0: no P staging and no O accumulators — four tiles of shared memory and 128 score columns, two CTAs per SM.
// CP (compact, up to four query chunks): ONE P buffer with the Q ring lying in its first two slabs, O in TMEM columns 0 .. 127
// and S in 128 .. 255 — six tiles and 256 columns, two CTAs per SM.  The second sweep then runs without prefetch (a Q chunk
// may only land once the MMA 2 that read the P buffer has retired); the other CTA on the SM fills those gaps.
template <int NT, int NKT, bool Q0, bool CP>
__global__ void __launch_bounds__(kFwdThreads, (Q0 || CP) ? 2 : 1)
    attn_tcl_fwd2_kernel(const __grid_constant__ CUtensorMap tmKm /* K-major {32, 128} box over qkv */,
                         const __grid_constant__ CUtensorMap tmMn /* MN-major {32, 128} box over qkv */,
                         const AttnLongParams p) {
  pdl_entry();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  static_assert(!(Q0 && CP), "Q0 stages no P");
  static_assert(!CP || NT <= 4, "CP: O accumulators in TMEM columns 0 .. 127");
  unsigned char* sQ = base;                          // two K-major query-chunk tiles (ring); CP: slabs 0 / 1 of the P buffer
  unsigned char* sK = CP ? base + 4 * kTileBytes : sQ + 2 * kTileBytes;   // K-major key tile
  unsigned char* sV = sK + kTileBytes;               // MN-major value tile
  unsigned char* sP = CP ? base : sV + kTileBytes;   // (CP: 1, else 2) x 4 slabs: P chunk, q contiguous, 128 key rows per slab
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(Q0 ? sP : (CP ? base + 6 * kTileBytes : sP + 2 * 4 * kTileBytes));
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);
  unsigned long long* bar_q = &bars[0];              // [2]
  unsigned long long* bar_k = &bars[2];
  unsigned long long* bar_v = &bars[3];
  unsigned long long* bar_s = &bars[4];
  unsigned long long* bar_o = &bars[5];              // [2]
  static_assert(NKT == NT || NKT == NT - 1, "at most one chunk of trailing positions");
  static_assert(NT * DH <= 256, "O accumulators: TMEM columns 0 .. 255, S in 256 .. 383");
  constexpr bool TAIL = NKT < NT;
  constexpr int NQQ = (NT * kTile + kFwdThreads - 1) / kFwdThreads;
  constexpr int kTmemCols = (Q0 || CP) ? 256 : 512;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int row = quarter * 32 + lane;
  const int c0 = half * 64, h16 = half * 16;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T, TQ = p.TQ, D = p.H * DH;
  const int ntail = TAIL ? tail_keys(T) : 0;
  __shared__ float xs_m[2][kTile], xs_l[2][kTile];   // running (max, sum) of the two column halves of every key row
  __shared__ float red8[8];
  __shared__ float red[TAIL ? 256 : 1];
  __shared__ float red_q[Q0 ? 256 : 1];
  __shared__ float ot_s[kTailMax][32];
  __shared__ float pt_s[TAIL ? kTailMax : 1][TAIL ? NT * kTile : 1];
  __shared__ __align__(16) float tr_s[TAIL ? 5 : 1][kTailMax][32];
  if (TAIL) fetch_tail_rows<kFwdThreads / 32>(tr_s, p.qkv, nullptr, (size_t)b * T + (size_t)(NT - 1) * kTile, ntail, D, h, warp, lane);

  if (!TAIL && !Q0) {                                // short last query chunk: the MMAs read the whole staging tile
    for (int i = tid * 16; i < (CP ? 1 : 2) * 4 * kTileBytes; i += kFwdThreads * 16) *reinterpret_cast<float4*>(sP + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_async_smem();
  }
  if (tid == 0) {
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = *tmem_slot;
  const unsigned tm_O = tmem, tm_S = Q0 ? tmem : (CP ? tmem + 128 : tmem + 256);
  const unsigned lane_off = (unsigned)(quarter * 32) << 16;
  const unsigned idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(DH >> 3) << 17) |
                          ((unsigned)(128 >> 4) << 24);

  // Q chunk qc into the ring slot of global step gg (thread 0; the slot's previous reader, step gg - 2, has retired)
  auto issue_q = [&](int qc, unsigned gg) {
    const int slot = (int)(gg & 1);
    mbar_expect_tx(&bar_q[slot], (unsigned)kTileBytes);
    tma_load_2d(sQ + slot * kTileBytes, &tmKm, &bar_q[slot], D + h * DH, b * T + qc * kTile);
  };
  if (tid == 0) {
    mbar_expect_tx(bar_k, (unsigned)kTileBytes);
    tma_load_2d(sK, &tmKm, bar_k, h * DH, b * T);
    mbar_expect_tx(bar_v, (unsigned)kTileBytes);
    tma_load_2d(sV, &tmMn, bar_v, 2 * D + h * DH, b * T);
    issue_q(0, 0);
  }
  if (TAIL) {
    // trailing key rows (thread = query, Q rows straight from global memory, rounded like the TMA unit rounds them; K[k*] from
    // the fetch at kernel start): softmax over the query axis across the CTA, P[k*][q] kept for the output rows
#pragma unroll 1
    for (int i = 0; i < ntail; ++i) {
      const int ks = (NT - 1) * kTile + i;
      const float rmask = __ldg(p.mask + (size_t)b * T + ks) > 0.f ? 0.f : -1e9f;
      float sc[NQQ];
      float mx = -INFINITY;
#pragma unroll
      for (int qq = 0; qq < NQQ; ++qq) {
        const int q = qq * kFwdThreads + tid;
        sc[qq] = 0.f;
        if (q < T) {
          float qr[32];
          const float4* qp = (q >= NKT * kTile) ? reinterpret_cast<const float4*>(tr_s[kTrQ][q - NKT * kTile])
                                                : reinterpret_cast<const float4*>(p.qkv + ((size_t)b * T + q) * 3 * D + D + h * DH);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 q4 = qp[c];
            qr[4 * c] = to_tf32(q4.x); qr[4 * c + 1] = to_tf32(q4.y); qr[4 * c + 2] = to_tf32(q4.z); qr[4 * c + 3] = to_tf32(q4.w);
          }
          sc[qq] = fmaf(dot32_tf32(tr_s[kTrK][i], qr), p.inv_scale, rmask);
          mx = fmaxf(mx, sc[qq]);
        }
      }
      const float mxl = block_max256(mx, red8, tid) * kLog2e;
      float sum = 0.f;
#pragma unroll
      for (int qq = 0; qq < NQQ; ++qq) {
        sc[qq] = (qq * kFwdThreads + tid < T) ? exp2f(fmaf(sc[qq], kLog2e, -mxl)) : 0.f;
        sum += sc[qq];
      }
      const float inv = 1.f / block_sum256(sum, red8, tid);
#pragma unroll
      for (int qq = 0; qq < NQQ; ++qq)
        if (qq * kFwdThreads + tid < T) pt_s[i][qq * kFwdThreads + tid] = sc[qq] * inv;
      if (tid == 0) reinterpret_cast<float2*>(p.stats)[(size_t)(b * p.H + h) * T + ks] = make_float2(mxl, inv);
    }
  }
  float o0 = 0.f;                                     // Q0: O[0][column lane]
  float ot[kTailMax / 2];
#pragma unroll
  for (int i = 0; i < kTailMax / 2; ++i) ot[i] = 0.f;
  unsigned g = 0;                                     // global step: Q ring slot g & 1, bar_s parity g & 1
  int n = 0;                                          // MMA 2 counter: P buffer n & 1
  constexpr int kSteps = (Q0 ? 1 : 2) * NKT;          // steps per key tile
#pragma unroll 1
  for (int kt = 0; kt < NKT; ++kt) {
    const unsigned par = (unsigned)(kt & 1);
    const int kg = kt * kTile + row;
    const bool valid = kg < T;
    const float rowmask = (valid && __ldg(p.mask + (size_t)b * T + kg) > 0.f) ? 0.f : -1e9f;
    float st[kTailMax];
#pragma unroll
    for (int i = 0; i < kTailMax; ++i) st[i] = 0.f;
    if (TAIL) {                                       // this key row's scores against the trailing queries (both threads)
      mbar_wait(bar_k, par);
      float kr[32];
      load_row_km(sK, row, kr);
#pragma unroll
      for (int i = 0; i < kTailMax; ++i) {
        if (i < ntail) {
          float qr[32];
          load_row32(tr_s[kTrQt][i], qr);
          st[i] = fmaf(dot32(kr, qr), p.inv_scale, rowmask);
        }
      }
    }
    float m = -INFINITY, l = 0.f, s0 = 0.f, mxl = 0.f, inv = 0.f;
#pragma unroll 1
    for (int u = 0; u < kSteps; ++u, ++g) {
      const int qc = u < NKT ? u : u - NKT;
      const bool sweep2 = u >= NKT;
      const int nq = min(kTile, TQ - qc * kTile);
      if (tid == 0) {
        if (CP && u > NKT) {                           // a later step of sweep 2: its Q chunk lands in the P buffer, free once
          mbar_wait(&bar_o[0], (unsigned)((n - 1) & 1));   // the previous MMA 2 has retired
          issue_q(qc, g);
        }
        mbar_wait(&bar_q[g & 1], (g >> 1) & 1);
        if (u == 0) mbar_wait(bar_k, par);
        tc_fence_after();
        const unsigned idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(nq >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
#pragma unroll
        for (int k = 0; k < DH / 8; ++k)
          umma_tf32(tm_S, make_desc(smem_u32(sK) + k * 32, 16, 1024, 2),
                    make_desc(smem_u32(sQ) + (g & 1) * kTileBytes + k * 32, 16, 1024, 2), idesc1, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
        // the next chunk (this key tile's next step, or the first of the next key tile) into the other ring slot
        if (CP) {
          if (u + 1 <= NKT) issue_q(u + 1 < NKT ? u + 1 : 0, g + 1);      // sweep 1 (P idle): the next chunk / the first of sweep 2
        } else {
          if (u + 1 < kSteps) issue_q(u + 1 < NKT ? u + 1 : u + 1 - NKT, g + 1);
          else if (kt + 1 < NKT) issue_q(0, g + 1);
        }
      }
      __syncwarp();
      mbar_wait(bar_s, g & 1);
      tc_fence_after();
      const int cend = min(nq, c0 + 64);
      if (!sweep2) {
        // sweep 1: chunk maximum, then the running sum rescaled to the new maximum
        float cm = -INFINITY;
        for (int c = c0; c < cend; c += 16) {
          float v[16];
          tmem_ld16(tm_S + lane_off + c, v);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (TAIL || qc * kTile + c + j < T) cm = fmaxf(cm, fmaf(v[j], p.inv_scale, rowmask));
          if (Q0 && qc == 0 && c == 0) s0 = fmaf(v[0], p.inv_scale, rowmask);
        }
        // Rescale by the difference of the two ROUNDED products max * log2 e, the very numbers the element exponents
        // fmaf(s, log2 e, -ml) are taken against: for a padded key row s = -1e9 + x collapses to -1e9, max * log2 e carries a
        // rounding error of up to 2^6 and that error must enter sum and elements as the same factor (it cancels in P).
        if (cm > m) {                                  // exp2(-inf) = 0 covers the first chunk
          l *= exp2f(__fmul_rn(m, kLog2e) - __fmul_rn(cm, kLog2e));      // __fmul_rn: never contracted into an FMA
          m = cm;
        }
        const float ml = __fmul_rn(m, kLog2e);
        for (int c = c0; c < cend; c += 16) {
          float v[16];
          tmem_ld16(tm_S + lane_off + c, v);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (TAIL || qc * kTile + c + j < T) l += exp2f(fmaf(fmaf(v[j], p.inv_scale, rowmask), kLog2e, -ml));
        }
        tc_fence_before();
        __syncthreads();                               // every thread has read S: the next MMA 1 may overwrite it
        if (u == NKT - 1) {
          // the row's statistics: both column halves and the trailing queries (both threads compute the same values)
          xs_m[half][row] = m;
          xs_l[half][row] = l;
          __syncthreads();
          float M = fmaxf(xs_m[0][row], xs_m[1][row]);
#pragma unroll
          for (int i = 0; i < kTailMax; ++i)
            if (i < ntail) M = fmaxf(M, st[i]);
          mxl = __fmul_rn(M, kLog2e);
          float L = xs_l[0][row] * exp2f(__fmul_rn(xs_m[0][row], kLog2e) - mxl) + xs_l[1][row] * exp2f(__fmul_rn(xs_m[1][row], kLog2e) - mxl);
#pragma unroll
          for (int i = 0; i < kTailMax; ++i) {
            if (i < ntail) {
              st[i] = exp2f(fmaf(st[i], kLog2e, -mxl));
              L += st[i];
            }
          }
          inv = valid ? 1.f / L : 0.f;
          if (valid && half == 0) {
            float2* stp = reinterpret_cast<float2*>(p.stats) + ((size_t)(b * p.H + h) * T + kg);
            *stp = make_float2(mxl, inv);
          }
        }
      } else {
        // sweep 2: normalised chunk -> shared memory -> MMA 2
        const int buf = CP ? 0 : (n & 1);
        unsigned char* pb = sP + buf * 4 * kTileBytes;
        if (CP) {
          if (n >= 1) {
            mbar_wait(&bar_o[0], (unsigned)((n - 1) & 1));
            tc_fence_after();
          }
        } else if (n >= 2) {
          mbar_wait(&bar_o[buf], (unsigned)(((n >> 1) - 1) & 1));
          tc_fence_after();
        }
        for (int c = c0; c < cend; c += 16) {
          float v[16];
          tmem_ld16(tm_S + lane_off + c, v);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = (TAIL || qc * kTile + c + j < T) ? to_tf32(exp2f(fmaf(fmaf(v[j], p.inv_scale, rowmask), kLog2e, -mxl)) * inv) : 0.f;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(pb + mn_major_off(c + j, row, kTile)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();                               // P staged, S read by every thread
        if (tid == 0) {
          tc_fence_after();
          if (qc == 0) mbar_wait(bar_v, par);
          for (int j = 0; j < kTile / 8; ++j)
            umma_tf32(tm_O + qc * DH, make_desc(smem_u32(pb) + j * 1024, kTileBytes, 512, 1),
                      make_desc(smem_u32(sV) + j * 1024, kTileBytes, 512, 1), idesc2, (kt > 0 || j > 0) ? 1u : 0u);
          umma_commit(&bar_o[buf]);
        }
        __syncwarp();
        ++n;
      }
    }
    if (Q0) {
      // O[0] += sum over this tile's keys of P[k][0] V[k] (the second threads of the rows add zeros)
      mbar_wait(bar_v, par);
      float c[32];
      load_row_mn(sV, row, c);
      const float p0 = (half == 0) ? exp2f(fmaf(s0, kLog2e, -mxl)) * inv : 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) c[j] *= p0;
      o0 += half_colsum128(c, red_q, tid);
    } else if (TAIL) {
      // while the MMA 2s run: O[q*] += sum over this tile's keys of P[k][q*] V[k]
      mbar_wait(bar_v, par);
      float vr[32];
      load_row_mn(sV, row, vr);
#pragma unroll
      for (int ii = 0; ii < kTailMax / 2; ++ii) {
        if (2 * ii < ntail) {
          const float pt = ((half == 0) ? st[2 * ii] : st[2 * ii + 1]) * inv;
          float c[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) c[j] = pt * vr[j];
          ot[ii] += half_colsum128(c, red, tid);
        }
      }
    }
    if (tid == 0 && kt + 1 < NKT) {
      // the key tile's K and V are free: every MMA 1 has retired (bar_s waits), the last MMA 2 is waited for here
      if (!Q0) {
        const int last = n - 1;
        if (CP) mbar_wait(&bar_o[0], (unsigned)(last & 1));
        else mbar_wait(&bar_o[last & 1], (unsigned)((last >> 1) & 1));
      }
      mbar_expect_tx(bar_k, (unsigned)kTileBytes);
      tma_load_2d(sK, &tmKm, bar_k, h * DH, b * T + (kt + 1) * kTile);
      mbar_expect_tx(bar_v, (unsigned)kTileBytes);
      tma_load_2d(sV, &tmMn, bar_v, 2 * D + h * DH, b * T + (kt + 1) * kTile);
      if (CP) issue_q(0, g);                           // the P buffer is free: first chunk of the next key tile
    }
    __syncthreads();                                  // xs_* / V row reads are done before the next key tile
  }
  if (Q0) {
    if (warp == 0) {
      if (TAIL) {
#pragma unroll
        for (int i = 0; i < kTailMax; ++i)
          if (i < ntail) o0 = fmaf(pt_s[i][0], tr_s[kTrV][i][lane], o0);
      }
      const size_t e = (size_t)b * T * D + h * DH + lane;
      store_ctx1(p, e, o0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
    return;
  }
  {
    const int last = n - 1;
    if (CP) mbar_wait(&bar_o[0], (unsigned)(last & 1));
    else mbar_wait(&bar_o[last & 1], (unsigned)((last >> 1) & 1));
    tc_fence_after();
  }
  if (TAIL) {
    if (quarter == 0) {
#pragma unroll
      for (int ii = 0; ii < kTailMax / 2; ++ii) ot_s[2 * ii + half][lane] = ot[ii];
    }
    __syncthreads();
  }
#pragma unroll
  for (int qc = 0; qc < NT; ++qc) {
    const int q = qc * kTile + row;
    float o[16];
    if (qc < NKT) {
      tmem_ld16(tm_O + lane_off + qc * DH + h16, o);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = ot_s[row & (kTailMax - 1)][h16 + j];
    }
    if (TAIL && q < T) {
#pragma unroll
      for (int i = 0; i < kTailMax; ++i)
        if (i < ntail) axpy16(pt_s[i][q], tr_s[kTrV][i] + h16, o);
    }
    if (q < T) store_ctx16(p, ((size_t)b * T + q) * D + h * DH + h16, o);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

// ------------------------------------------------------------------------------------------------ backward
// 256 threads: warps w and w + 4 share the TMEM lane quarter w & 3 (a warp reaches lanes 32 (w % 4) ... + 31), i.e. two
// threads own one key row (one query row in the dQ epilogue) and split its columns — 64 + 64 score columns in the P and dS
// passes, 16 + 16 head columns in the row epilogues.  The kernel is bound by the instruction stream of the row owners
// (ncu, 128-thread version: 6.8 k instructions per warp and item at 18 % issue utilisation, one warp per scheduler), so
// halving the columns per thread halves the critical path.
//
// NKT = key tiles = query chunks on the tensor path (NT, or NT - 1 when T = 128 (NT - 1) + 1 ... 4: trailing positions).
// SINGLE (NT = 2, NKT = 1: T = 129 ... 132): the seven operand tiles are loaded once, S = K Q^T and dP = V dO^T are issued
// together, P goes to its own TMEM columns so that S survives for the dS pass: one TMA round trip and three MMA round trips
// per (batch, head) instead of four and eight.
// Q0 (the encoder's top layer under SOS-rows-only): dctx is non-zero in the row of query 0 only.  With g = dO[0],
// a_k = V_k . g, c_k = a_k P[k][0]:  dV_k = P[k][0] g  and  dS[k][q] = P[k][q] ([q = 0] a_k - c_k) / sqrt(d_h)  are thread-local,
// so phase 1 (P pass, dV MMA), the dP MMA, the dO tiles and the second TMEM read of the dS pass all disappear.
// SINGLE && Q0 (the top layer's backward at T = 129 ... 132): 224 TMEM columns are enough (dK | dQ | S, dS written in place of
// S) and the dS^T staging lies over the K-major K / Q / V tiles, dead once S = K Q^T has retired and the threads have read
// their K / V rows: six tiles of shared memory, 256 TMEM columns, two CTAs per SM.
template <int NT, int NKT, bool SINGLE, bool Q0>
__global__ void __launch_bounds__(kBwdThreads, SINGLE ? 2 : 1)
    attn_tcl_bwd_kernel(const __grid_constant__ CUtensorMap tmKm /* K-major {32,128} over qkv */,
                        const __grid_constant__ CUtensorMap tmMn /* MN-major {32,128} over qkv */,
                        const __grid_constant__ CUtensorMap tmDOk /* K-major {32,128} over dctx */,
                        const __grid_constant__ CUtensorMap tmDOm /* MN-major {32,128} over dctx */,
                        const AttnLongParams p) {
  pdl_entry();
  constexpr bool TAIL = NKT < NT;
  static_assert(NKT == NT || NKT == NT - 1, "at most one chunk of trailing positions");
  static_assert(!SINGLE || (NT == 2 && NKT == 1), "SINGLE: one full tile + trailing positions");
  constexpr int NQQ = (NT * kTile + kBwdThreads - 1) / kBwdThreads;    // queries per thread in the trailing-key section
  constexpr int kTmemCols = SINGLE ? 256 : 512;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr bool SQ = SINGLE && Q0;                  // tiles: [Kk | Qk | Vk | -] = dS^T staging later, then Km, Qm
  // S2 = SINGLE && !Q0, also two CTAs per SM — tiles: [Qk | dOm | Vk | dOk] = dS^T staging later, then Kk (reloaded as Km once
  // S = K Q^T has retired), Qm; TMEM: columns 128 .. 255 hold S, then P in place, then dK | dQ; columns 0 .. 127 hold dV and,
  // once every thread has read dV, dP (its MMA is issued then: one MMA round trip more than the 512-column layout) and dS.
  constexpr bool S2 = SINGLE && !Q0;
  unsigned char* sKk = S2 ? base + 4 * kTileBytes : base;              // per key tile: K K-major, V K-major, K MN-major
  unsigned char* sVk = SINGLE ? base + 2 * kTileBytes : sKk + kTileBytes;
  unsigned char* sKm = SINGLE ? base + 4 * kTileBytes : sVk + kTileBytes;
  // Per query chunk, slots of three tiles: {Q K-major, dO MN-major, -} in phase 1 and {Q K-major, Q MN-major, dO K-major}
  // in phase 2.  DB (rows without trailing positions): two slots, slot = step & 1, and the loads of step s + 1 are issued
  // when step s starts (its slot was last read by step s - 1, whose MMAs have retired), so their latency hides behind a
  // whole step: T = 193 bwd 2.95 -> 2.71 ms, T = 384 3.33 -> 3.07 ms.  With trailing positions the extra 48 KB of shared memory
  // cost the trailing-key section its L1 hits (230 KB of shared memory leave 28 KB of L1: T = 260 3.25 -> 3.77 ms), so those
  // variants keep one slot and load at the start of the step.  SINGLE: four tiles, loaded once.
  // Q0 needs two tiles per slot (Q K-major, Q MN-major) and its trailing-key section reads no dO rows: double-buffered too
  // (193 KB of shared memory keep the 196 KB carve-out, i.e. 60 KB of L1).
  constexpr bool DB = !TAIL || Q0;
  constexpr int kSlotTiles = Q0 ? 2 : 3;
  constexpr int kChunkTiles = SINGLE ? 4 : (DB ? 2 * kSlotTiles : kSlotTiles);
  unsigned char* sC = SQ ? base + kTileBytes : (S2 ? base : sKm + kTileBytes);   // SINGLE: Q K-major (S2: and dO MN-major) of "slot 0"
  unsigned char* sY = SINGLE ? base : sC + kChunkTiles * kTileBytes;  // dS^T chunk, q contiguous: 4 slabs x 128 key rows
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(SINGLE ? base + 6 * kTileBytes : sY + 4 * kTileBytes);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);
  unsigned long long* bar_kt = &bars[0];
  unsigned long long* bar_ld = &bars[4];             // [2]: chunk tiles of a slot have landed
  unsigned long long* bar_m1 = &bars[2];
  unsigned long long* bar_m2 = &bars[3];
  __shared__ float red[256];
  __shared__ float red8[8];
  __shared__ float red_b[3][8][16];
  __shared__ float tp_s[Q0 ? 1 : 2][Q0 ? 1 : kTailMax][Q0 ? 1 : kTile];   // P / dP of (key row, trailing query), exchanged between the row's two threads
  __shared__ float dq_s[kTailMax][32];               // dQ rows of the trailing queries
  __shared__ __align__(16) float g_s[32];            // Q0: dO[query 0]
  __shared__ __align__(16) float tr_s[TAIL ? 5 : 1][kTailMax][32];     // rows of the trailing positions (fetch_tail_rows)
  __shared__ float ds_s[TAIL ? kTailMax : 1][TAIL ? (SINGLE ? kTile + 8 : NT * kTile) : 1];   // dS[trailing key][query]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int row = quarter * 32 + lane;               // key row of the tile (query row of the chunk in the dQ epilogue)
  const int c0 = half * 64, h16 = half * 16;         // this thread's score columns [c0, c0 + 64) and head columns [h16, h16 + 16)
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T, TQ = p.TQ, D = p.H * DH;
  const int ntail = TAIL ? tail_keys(T) : 0;         // trailing positions handled off the tensor path (see kTailMax)

  const size_t tail_row0 = (size_t)b * T + (size_t)(NT - 1) * kTile;   // first trailing position (TAIL)
  if (TAIL) fetch_tail_rows<kBwdThreads / 32>(tr_s, p.qkv, Q0 ? nullptr : p.dctx, tail_row0, ntail, D, h, warp, lane);
  if (Q0 && warp == 7) g_s[lane] = __ldg(p.dctx + (size_t)b * T * D + h * DH + lane);
  if (!TAIL) {                                       // a short last chunk leaves part of the dS^T staging unwritten
    for (int i = tid * 16; i < 4 * kTileBytes; i += kBwdThreads * 16) *reinterpret_cast<float4*>(sY + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_async_smem();
  }
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = *tmem_slot;
  const unsigned tm_dV = tmem, tm_dK = SQ ? tmem : (S2 ? tmem + 128 : tmem + 32), tm_dQ = SQ ? tmem + 32 : (S2 ? tmem + 160 : tmem + 64);
  const unsigned tm_S = SINGLE ? tmem + 128 : tmem + 256;              // SQ: dS in place of S
  const unsigned tm_dP = SQ ? tmem + 128 : (S2 ? tmem : tmem + 384);   // S2: over dV, issued once dV has been read
  const unsigned tm_P = tm_S;                               // P in place of S
  const unsigned lane_off = (unsigned)(quarter * 32) << 16;
  const unsigned idesc_ts = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((unsigned)(DH >> 3) << 17) |
                            ((unsigned)(128 >> 4) << 24);   // A from TMEM, B MN-major, N = 32
  const unsigned idesc_mm = idesc_ts | (1u << 15);          // A MN-major from shared memory, B MN-major

  float acc_k[16], acc_v[16], acc_q[16];                    // column sums of this thread's dK / dV / dQ half rows (bias gradient)
#pragma unroll
  for (int j = 0; j < 16; ++j) acc_k[j] = acc_v[j] = acc_q[j] = 0.f;
  float dqt[2] = {0.f, 0.f};                                // dQ[trailing query 2 ii + half][column lane], summed over the key tiles

  // chunk loads of step s of a key tile (s < NKT: phase 1, else phase 2) into the slot of global step g; thread 0 only
  auto issue_chunk = [&](int s, unsigned g) {
    const int slot = DB ? (int)(g & 1) : 0;
    unsigned char* t0 = sC + slot * kSlotTiles * kTileBytes;
    const int r0 = b * T + (s < NKT ? s : s - NKT) * kTile;
    if (s < NKT) {
      mbar_expect_tx(&bar_ld[slot], (unsigned)(2 * kTileBytes));
      tma_load_2d(t0, &tmKm, &bar_ld[slot], D + h * DH, r0);
      tma_load_2d(t0 + kTileBytes, &tmDOm, &bar_ld[slot], h * DH, r0);
    } else {
      mbar_expect_tx(&bar_ld[slot], (unsigned)((Q0 ? 2 : 3) * kTileBytes));
      tma_load_2d(t0, &tmKm, &bar_ld[slot], D + h * DH, r0);
      tma_load_2d(t0 + kTileBytes, &tmMn, &bar_ld[slot], D + h * DH, r0);
      if (!Q0) tma_load_2d(t0 + 2 * kTileBytes, &tmDOk, &bar_ld[slot], h * DH, r0);
    }
  };
  unsigned char* const sDOk1 = base + 3 * kTileBytes;       // S2: dO K-major (read by the dP MMA)
  unsigned char* const sQm1 = base + 5 * kTileBytes;        // SINGLE: Q MN-major
  unsigned step = 0;                                        // bar_ld / bar_m1 / bar_m2 complete once per (key tile, phase, chunk)
#pragma unroll 1
  for (int kt = 0; kt < NKT; ++kt) {
    const int kg = kt * kTile + row;
    const bool valid = kg < T;
    const float rowmask = (valid && __ldg(p.mask + (size_t)b * T + kg) > 0.f) ? 0.f : -1e9f;
    float mxl = 0.f, inv = 0.f;
    if (valid) {
      const float2 st = __ldg(reinterpret_cast<const float2*>(p.stats) + ((size_t)(b * p.H + h) * T + kg));
      mxl = st.x;
      inv = st.y;
    }
    if (tid == 0) {                                         // every MMA that read the previous key tile has retired (bar_m2 waits)
      mbar_expect_tx(bar_kt, (unsigned)((SINGLE ? (Q0 ? 5 : 6) : 3) * kTileBytes));
      tma_load_2d(sKk, &tmKm, bar_kt, h * DH, b * T + kt * kTile);
      tma_load_2d(sVk, &tmKm, bar_kt, 2 * D + h * DH, b * T + kt * kTile);
      if (SINGLE) {
        tma_load_2d(sC, &tmKm, bar_kt, D + h * DH, b * T);                       // phase 1 = step 0 = slot 0: Q K-major, dO MN-major
        if (!Q0) {
          tma_load_2d(sDOk1, &tmDOk, bar_kt, h * DH, b * T);
          tma_load_2d(sC + kTileBytes, &tmDOm, bar_kt, h * DH, b * T);
        }
        tma_load_2d(sQm1, &tmMn, bar_kt, D + h * DH, b * T);
      }
      if (!S2) tma_load_2d(sKm, &tmMn, bar_kt, h * DH, b * T + kt * kTile);   // S2: lands in the K-major tile's place later
      if (DB && !SINGLE) issue_chunk(Q0 ? NKT : 0, step);
    }
    if (TAIL && kt == 0) {
      // while the first TMA loads are in flight — trailing key rows (thread = query, Q / dO rows straight from global
      // memory, K[k*] / V[k*] from the rows fetched at kernel start): P from the forward's statistics, dP = V[k*] . dO[q], delta = sum_q P dP, dS = P (dP - delta) / sqrt(d_h);
      // dV[k*] = sum_q P dO[q] and dK[k*] = sum_q dS Q[q] are block column sums, dQ[q] += dS K[k*] joins the last epilogue
#pragma unroll 1
      for (int i = 0; i < ntail; ++i) {                    // CTA-uniform
        const int ks = (NT - 1) * kTile + i;
        const float rmask = __ldg(p.mask + (size_t)b * T + ks) > 0.f ? 0.f : -1e9f;
        const float2 stt = __ldg(reinterpret_cast<const float2*>(p.stats) + ((size_t)(b * p.H + h) * T + ks));
        const float* krow = tr_s[kTrK][i];
        const float4* vrow = reinterpret_cast<const float4*>(tr_s[kTrV][i]);
        float pv[NQQ], dpv[NQQ];
        float col[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) col[j] = 0.f;
        float dl = 0.f;
#pragma unroll
        for (int qq = 0; qq < NQQ; ++qq) {
          const int q = qq * kBwdThreads + tid;
          pv[qq] = dpv[qq] = 0.f;
          if (q < T) {
            float qr[32], dor[32];
            // rows of the trailing queries come from the fetch at kernel start (shared memory), the others from global memory
            const bool tq = q >= NKT * kTile;
            const float4* qp = tq ? reinterpret_cast<const float4*>(tr_s[kTrQ][q - NKT * kTile])
                                  : reinterpret_cast<const float4*>(p.qkv + ((size_t)b * T + q) * 3 * D + D + h * DH);
            const float4* dp4 = (tq && !Q0) ? reinterpret_cast<const float4*>(tr_s[kTrDO][q - NKT * kTile])
                                            : reinterpret_cast<const float4*>(p.dctx + ((size_t)b * T + q) * D + h * DH);
            float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 q4 = qp[c], v4 = vrow[c];
              const float4 g4 = Q0 ? (q == 0 ? reinterpret_cast<const float4*>(g_s)[c] : make_float4(0.f, 0.f, 0.f, 0.f)) : dp4[c];
              qr[4 * c] = to_tf32(q4.x); qr[4 * c + 1] = to_tf32(q4.y); qr[4 * c + 2] = to_tf32(q4.z); qr[4 * c + 3] = to_tf32(q4.w);
              dor[4 * c] = g4.x; dor[4 * c + 1] = g4.y; dor[4 * c + 2] = g4.z; dor[4 * c + 3] = g4.w;
              d4[0] = fmaf(v4.x, g4.x, d4[0]); d4[1] = fmaf(v4.y, g4.y, d4[1]);
              d4[2] = fmaf(v4.z, g4.z, d4[2]); d4[3] = fmaf(v4.w, g4.w, d4[3]);
            }
            const float sc = fmaf(dot32_tf32(krow, qr), p.inv_scale, rmask);
            const float pr = exp2f(fmaf(sc, kLog2e, -stt.x)) * stt.y;
            const float dp = (d4[0] + d4[1]) + (d4[2] + d4[3]);
            pv[qq] = pr;
            dpv[qq] = dp;
            dl = fmaf(pr, dp, dl);
#pragma unroll
            for (int j = 0; j < 32; ++j) col[j] = fmaf(pr, dor[j], col[j]);
          }
        }
        float delta;
        const float dv = block_colsum256_sum(col, dl, red, red8, tid, &delta);   // dV[k*][lane] in every thread
#pragma unroll
        for (int j = 0; j < 32; ++j) col[j] = 0.f;
#pragma unroll
        for (int qq = 0; qq < NQQ; ++qq) {
          const int q = qq * kBwdThreads + tid;
          const float ds = pv[qq] * ((dpv[qq] - delta) * p.inv_scale);
          if (q < T) {
            ds_s[i][q] = ds;
            const float4* qp = (q >= NKT * kTile) ? reinterpret_cast<const float4*>(tr_s[kTrQ][q - NKT * kTile])
                                                  : reinterpret_cast<const float4*>(p.qkv + ((size_t)b * T + q) * 3 * D + D + h * DH);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 q4 = qp[c];
              col[4 * c] = fmaf(ds, q4.x, col[4 * c]); col[4 * c + 1] = fmaf(ds, q4.y, col[4 * c + 1]);
              col[4 * c + 2] = fmaf(ds, q4.z, col[4 * c + 2]); col[4 * c + 3] = fmaf(ds, q4.w, col[4 * c + 3]);
            }
          }
        }
        const float dk = block_colsum256(col, red, tid);        // dK[k*][lane]
        if (tid < 32) {
          const size_t e = ((size_t)b * T + ks) * 3 * D + h * DH + tid;
          if (p.out_bf16) {
            unsigned short* o16 = reinterpret_cast<unsigned short*>(p.out);
            o16[e] = __bfloat16_as_ushort(__float2bfloat16_rn(dk));
            o16[e + 2 * D] = __bfloat16_as_ushort(__float2bfloat16_rn(dv));
          } else {
            float* o32 = reinterpret_cast<float*>(p.out);
            o32[e] = dk;
            o32[e + 2 * D] = dv;
          }
          if (p.dbias) {
            atomicAdd(p.dbias + 0 * D + h * DH + tid, dk);
            atomicAdd(p.dbias + 2 * D + h * DH + tid, dv);
          }
        }
      }
    }
    // ---------------- phase 1: dV[keys x 32] = sum over query chunks of P[keys x q] dO[q x 32]
#pragma unroll 1
    for (int qc = 0; qc < (Q0 ? 0 : NKT); ++qc, ++step) {
      const unsigned par = step & 1;
      const int nq = min(kTile, TQ - qc * kTile);
      const unsigned idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(nq >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
      const int slot = DB ? (int)(step & 1) : 0;
      unsigned char* const sQk = sC + slot * kSlotTiles * kTileBytes;
      unsigned char* const sDOm = sQk + kTileBytes;
      if (tid == 0) {
        if (!SINGLE) {
          if (DB) issue_chunk(qc + 1, step + 1);           // the next step of this key tile (phase 1 or the first of phase 2)
          else issue_chunk(qc, step);
        }
        if (qc == 0) mbar_wait(bar_kt, (unsigned)(kt & 1));
        if (!SINGLE) mbar_wait(&bar_ld[slot], DB ? ((step >> 1) & 1) : (step & 1));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < DH / 8; ++k)       // S = K Q^T
          umma_tf32(tm_S, make_desc(smem_u32(sKk) + k * 32, 16, 1024, 2), make_desc(smem_u32(sQk) + k * 32, 16, 1024, 2),
                    idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_m1);
      }
      __syncwarp();
      if (TAIL && qc == 0) {
        // while the MMA runs: P (first thread of the row) and dP (second thread) of this key row against the trailing
        // queries — thread-local dot products; K / V rows as the MMAs see them, Q / dO rows from the fetch at kernel start
        // (Q rounded like the TMA unit rounds it so that P matches the forward's statistics)
        mbar_wait(bar_kt, (unsigned)(kt & 1));
        float kv[32];
        load_row_km(half ? sVk : sKk, row, kv);
#pragma unroll
        for (int i = 0; i < kTailMax; ++i) {
          if (i < ntail) {
            float r[32];
            float val;
            if (half == 0) {
              load_row32(tr_s[kTrQt][i], r);
              const float sc = fmaf(dot32(kv, r), p.inv_scale, rowmask);
              val = valid ? exp2f(fmaf(sc, kLog2e, -mxl)) * inv : 0.f;
            } else {
              load_row32(tr_s[kTrDO][i], r);
              val = dot32(kv, r);
            }
            if (!Q0) tp_s[half][i][row] = val;   // read after the barrier that follows the P pass
          }
        }
      }
      mbar_wait(bar_m1, par);
      tc_fence_after();
      {
        const int cend = min(nq, c0 + 64);
        for (int c = c0; c < cend; c += 16) {
          float v[16];
          tmem_ld16(tm_S + lane_off + c, v);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = (TAIL || qc * kTile + c + j < T)
                       ? to_tf32(exp2f(fmaf(fmaf(v[j], p.inv_scale, rowmask), kLog2e, -mxl)) * inv) : 0.f;
          tmem_st16(tm_P + lane_off + c, v);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        for (int j = 0; j < nq / 8; ++j)       // A = P from TMEM (8 query columns per k-step), B = dO MN-major (+8 query rows)
          umma_tf32_ts(tm_dV, tm_P + j * 8, make_desc(smem_u32(sDOm) + j * 1024, kTileBytes, 512, 1), idesc_ts,
                       (qc > 0 || j > 0) ? 1u : 0u);
        umma_commit(bar_m2);
        if (S2) {                              // S = K Q^T has retired and every thread has read its K row: K MN-major over K K-major
          mbar_expect_tx(&bar_ld[0], (unsigned)kTileBytes);
          tma_load_2d(sKm, &tmMn, &bar_ld[0], h * DH, b * T);
        }
      }
      __syncwarp();
      mbar_wait(bar_m2, par);                  // S and the chunk tiles are free again
      tc_fence_after();
    }
    // ---------------- delta_k = V_k . dV_k (both threads of the row, over all 32 columns); dV half row out
    float pt[kTailMax], dpt[kTailMax];         // P / dP of (this key row, trailing query i); below pt = dS
    float delta = 0.f;
    float a_k = 0.f;                           // Q0: V_k . dO[0]
    if (Q0) {
      mbar_wait(bar_kt, (unsigned)(kt & 1));
      float r[32], g[32];
      load_row_km(sVk, row, r);
      load_row32(g_s, g);
      a_k = dot32(r, g);
#pragma unroll
      for (int i = 0; i < kTailMax; ++i) pt[i] = dpt[i] = 0.f;
      if (TAIL) {                              // P of this key row against the trailing queries (both threads of the row)
        load_row_km(sKk, row, r);
#pragma unroll
        for (int i = 0; i < kTailMax; ++i) {
          if (i < ntail) {
            load_row32(tr_s[kTrQt][i], g);
            const float sc = fmaf(dot32(r, g), p.inv_scale, rowmask);
            pt[i] = valid ? exp2f(fmaf(sc, kLog2e, -mxl)) * inv : 0.f;
          }
        }
      }
    } else {
      float o[32];
      tmem_ld16(tm_dV + lane_off, o);
      tmem_ld16(tm_dV + lane_off + 16, o + 16);
#pragma unroll
      for (int i = 0; i < kTailMax; ++i) {
        pt[i] = dpt[i] = 0.f;
        if (i < ntail) {                       // dV_k += P[k][q*] dO[q*]
          pt[i] = tp_s[0][Q0 ? 0 : i][Q0 ? 0 : row];
          dpt[i] = tp_s[Q0 ? 0 : 1][Q0 ? 0 : i][Q0 ? 0 : row];
          const float4* gp = reinterpret_cast<const float4*>(tr_s[kTrDO][i]);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 g4 = gp[c];
            o[4 * c] = fmaf(pt[i], g4.x, o[4 * c]); o[4 * c + 1] = fmaf(pt[i], g4.y, o[4 * c + 1]);
            o[4 * c + 2] = fmaf(pt[i], g4.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(pt[i], g4.w, o[4 * c + 3]);
          }
        }
      }
      // V row of this key from the K-major SWIZZLE_128B tile (as the MMAs saw it: TF32-rounded by TMA)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v4 = *reinterpret_cast<const float4*>(sVk + row * 128 + ((c ^ (row & 7)) << 4));
        delta = fmaf(v4.x, o[4 * c], delta);
        delta = fmaf(v4.y, o[4 * c + 1], delta);
        delta = fmaf(v4.z, o[4 * c + 2], delta);
        delta = fmaf(v4.w, o[4 * c + 3], delta);
      }
      if (valid) {
        if (half == 0) {
          store_row32(p.out, p.out_bf16 != 0, ((size_t)b * T + kg) * 3 * D + 2 * D + h * DH, o, 16);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc_v[j] += o[j];
        } else {
          store_row32(p.out, p.out_bf16 != 0, ((size_t)b * T + kg) * 3 * D + 2 * D + h * DH + 16, o + 16, 16);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc_v[j] += o[16 + j];
        }
      }
    }
    if (!Q0) {
#pragma unroll
      for (int i = 0; i < kTailMax; ++i) pt[i] = pt[i] * ((dpt[i] - delta) * p.inv_scale);   // dS[k][q*]
    }
    if (S2) {
      // every thread has read dV (and its V row): dP = V dO^T goes over the dV columns
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const unsigned idesc_p = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(kTile >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
#pragma unroll
        for (int k = 0; k < DH / 8; ++k)
          umma_tf32(tm_dP, make_desc(smem_u32(sVk) + k * 32, 16, 1024, 2), make_desc(smem_u32(sDOk1) + k * 32, 16, 1024, 2),
                    idesc_p, k > 0 ? 1u : 0u);
        umma_commit(bar_m1);
      }
      __syncwarp();
      mbar_wait(bar_m1, 1u);                   // second completion of bar_m1 (the first was S = K Q^T)
      tc_fence_after();
    }
    // ---------------- phase 2: dK += dS Q, dQ[chunk] += dS^T K
#pragma unroll 1
    for (int qc = 0; qc < NKT; ++qc, ++step) {
      const unsigned par = step & 1;
      const int nq = min(kTile, TQ - qc * kTile);
      const unsigned idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(nq >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
      const int slot = DB ? (int)(step & 1) : 0;
      unsigned char* const sQk = sC + slot * kSlotTiles * kTileBytes;
      unsigned char* const sQm = SINGLE ? sQm1 : sQk + kTileBytes;
      unsigned char* const sDOk = sQk + 2 * kTileBytes;
      if (!SINGLE || Q0) {
        if (tid == 0) {
          if (!SINGLE) {
            if (!DB) issue_chunk(NKT + qc, step);
            else if (qc + 1 < NKT) issue_chunk(NKT + qc + 1, step + 1);
            mbar_wait(&bar_ld[slot], DB ? ((step >> 1) & 1) : (step & 1));
          }
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < DH / 8; ++k) {     // S = K Q^T and dP = V dO^T, two independent chains
            umma_tf32(tm_S, make_desc(smem_u32(sKk) + k * 32, 16, 1024, 2), make_desc(smem_u32(sQk) + k * 32, 16, 1024, 2),
                      idesc_s, k > 0 ? 1u : 0u);
            if (!Q0)
              umma_tf32(tm_dP, make_desc(smem_u32(sVk) + k * 32, 16, 1024, 2), make_desc(smem_u32(sDOk) + k * 32, 16, 1024, 2),
                        idesc_s, k > 0 ? 1u : 0u);
          }
          umma_commit(bar_m1);
        }
        __syncwarp();
        mbar_wait(bar_m1, par);
        tc_fence_after();
      }
      if (Q0 && qc == 0) {
        // P[k][0] from the first score column (read by both threads of the row): delta, the dV half row, dS of the trailing queries
        float v[16];
        tmem_ld16(tm_S + lane_off, v);
        const float p0 = valid ? exp2f(fmaf(fmaf(v[0], p.inv_scale, rowmask), kLog2e, -mxl)) * inv : 0.f;
        delta = p0 * a_k;
        if (valid) {
          float o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = p0 * g_s[h16 + j];
          store_row32(p.out, p.out_bf16 != 0, ((size_t)b * T + kg) * 3 * D + 2 * D + h * DH + h16, o, 16);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc_v[j] += o[j];
        }
#pragma unroll
        for (int i = 0; i < kTailMax; ++i) pt[i] = pt[i] * ((0.f - delta) * p.inv_scale);   // dS[k][q*]: dP = 0 there
        if (SQ) {
          // dS goes over S and the dS^T staging over the K / Q / V tiles: both threads of every row have read S[k][0],
          // every thread its K / V rows
          tc_fence_before();
          __syncthreads();
          tc_fence_after();
        }
      }
      {
        const int cend = min(nq, c0 + 64);
        for (int c = c0; c < cend; c += 16) {
          float v[16], g[16];
          if (Q0) {
            tmem_ld16(tm_S + lane_off + c, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] = 0.f;
            if (qc == 0 && c == 0) g[0] = a_k;                          // dP[k][q] = [q = 0] V_k . dO[0]
          } else {
            tmem_ld16_issue(tm_S + lane_off + c, v);
            tmem_ld16_issue(tm_dP + lane_off + c, g);
            tmem_ld16_wait(v);
            tmem_ld16_wait(g);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            // S2: the S columns hold P (rounded to TF32 for the dV MMA) — taken as it is
            const float pr = S2 ? v[j]
                                : ((TAIL || qc * kTile + c + j < T) ? exp2f(fmaf(fmaf(v[j], p.inv_scale, rowmask), kLog2e, -mxl)) * inv : 0.f);
            g[j] = to_tf32(pr * ((g[j] - delta) * p.inv_scale));      // dS
          }
          tmem_st16(tm_dP + lane_off + c, g);
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(sY + mn_major_off(c + j, row, kTile)) = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
        }
      }
      tmem_st_wait();
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        if (S2) mbar_wait(&bar_ld[0], 0u);
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
      if (TAIL && qc == NKT - 1) {
        // while the last MMAs of the key tile run: dQ[q*] += sum over this tile's keys of dS[k][q*] K[k]; the first
        // threads of the rows take the even trailing queries, the second threads the odd ones
        float kr[32];
        if (S2) mbar_wait(&bar_ld[0], 0u);
        if (SINGLE) load_row_mn(sKm, row, kr);  // the K-major tile lies under the dS^T staging / was replaced by this copy
        else load_row_km(sKk, row, kr);
#pragma unroll
        for (int ii = 0; ii < kTailMax / 2; ++ii) {
          if (2 * ii < ntail) {                // CTA-uniform
            const float a = (half == 0) ? pt[2 * ii] : pt[2 * ii + 1];     // 0 beyond ntail
            float c[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) c[j] = a * kr[j];
            dqt[ii] += half_colsum128(c, red, tid);
          }
        }
      }
      mbar_wait(bar_m2, par);
      tc_fence_after();
    }
    {
      float o[16];
      tmem_ld16(tm_dK + lane_off + h16, o);
#pragma unroll
      for (int i = 0; i < kTailMax; ++i)
        if (i < ntail) axpy16(pt[i], tr_s[kTrQ][i] + h16, o);   // dK_k += dS[k][q*] Q[q*]
      if (valid) {
        store_row32(p.out, p.out_bf16 != 0, ((size_t)b * T + kg) * 3 * D + h * DH + h16, o, 16);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc_k[j] += o[j];
      }
    }
    tc_fence_before();
    __syncthreads();                           // every thread has read dV / dK before the next key tile overwrites them
    tc_fence_after();
  }
  if (TAIL) {
    if (quarter == 0) {
#pragma unroll
      for (int ii = 0; ii < kTailMax / 2; ++ii) dq_s[2 * ii + half][lane] = dqt[ii];
    }
    __syncthreads();
  }
#pragma unroll
  for (int qc = 0; qc < NT; ++qc) {
    const int q = qc * kTile + row;            // lanes = queries
    float o[16];
    if (qc < NKT) {
      tmem_ld16(tm_dQ + lane_off + qc * DH + h16, o);
    } else {                                   // the trailing queries' rows: the threads of rows 0 .. ntail-1
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = dq_s[row & (kTailMax - 1)][h16 + j];
    }
    if (TAIL && q < T) {
#pragma unroll
      for (int i = 0; i < kTailMax; ++i)       // dQ[q] += dS[k*][q] K[k*]
        if (i < ntail) axpy16(ds_s[i][q], tr_s[kTrK][i] + h16, o);
    }
    if (q < T) {
      store_row32(p.out, p.out_bf16 != 0, ((size_t)b * T + q) * 3 * D + D + h * DH + h16, o, 16);
#pragma unroll
      for (int j = 0; j < 16; ++j) acc_q[j] += o[j];
    }
  }
  if (p.dbias) {
    const float sk = warp_colsum16(acc_k, lane), sq = warp_colsum16(acc_q, lane), sv = warp_colsum16(acc_v, lane);
    if (lane < 16) {
      red_b[0][warp][lane] = sk;
      red_b[1][warp][lane] = sq;
      red_b[2][warp][lane] = sv;
    }
    __syncthreads();
    if (tid < 96) {
      const int m = tid >> 5, c = tid & 31, w0 = (c >> 4) * 4;
      atomicAdd(p.dbias + m * D + h * DH + c,
                (red_b[m][w0][c & 15] + red_b[m][w0 + 1][c & 15]) + (red_b[m][w0 + 2][c & 15] + red_b[m][w0 + 3][c & 15]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

constexpr size_t fwd_smem(int nt) { return 1024 + (size_t)(nt + 2 + 8) * kTileBytes + 128; }
constexpr size_t kFwdAliasSmem = 1024 + (size_t)5 * kTileBytes + 128;
constexpr size_t fwd_q0_smem(int nt) { return 1024 + (size_t)(nt + 2) * kTileBytes + 128; }   // no P staging
constexpr size_t kFwd2Smem = 1024 + (size_t)(2 + 2 + 8) * kTileBytes + 128;
constexpr size_t kFwd2Q0Smem = 1024 + (size_t)(2 + 2) * kTileBytes + 128;
constexpr size_t kFwd2CpSmem = 1024 + (size_t)(4 + 2) * kTileBytes + 128;
// score columns on the tensor path
inline int tensor_queries(int T) { const int t = tail_keys(T); return t ? T - t : (T + 15) / 16 * 16; }
constexpr size_t bwd_smem(bool single, bool tail, bool q0 = false) {
  return 1024 + (size_t)(single ? 6 : 3 + (q0 ? 4 : (tail ? 3 : 6)) + 4) * kTileBytes + 128;
}

}  // namespace

extern "C" int msx_attention_tcl_supported(const float* qkv, int T, int dh) {
  return (qkv && dh == 32 && T > 128 && T <= 6 * kTile && ((uintptr_t)qkv & 15) == 0) ? 1 : 0;
}

extern "C" int msx_attention_tcl_fwd_q0(const float* qkv, const float* mask, void* ctx, int ctx_bf16, float* stats, int q0_only,
                                        int B, int T, int H, int dh, void* stream);
extern "C" int msx_attention_tcl_bwd_q0(const float* qkv, const float* mask, const float* dctx, const float* stats, void* dqkv,
                                        int dqkv_bf16, float* dbias, int q0_only, int B, int T, int H, int dh, void* stream);
extern "C" int msx_attention_tcl_fwd_p(const float* qkv, const float* mask, void* ctx, void* ctx_lo, int ctx_bf16, float* stats,
                                       int q0_only, int B, int T, int H, int dh, void* stream);

extern "C" int msx_attention_tcl_fwd(const float* qkv, const float* mask, void* ctx, int ctx_bf16, float* stats, int B, int T,
                                     int H, int dh, void* stream) {
  return msx_attention_tcl_fwd_q0(qkv, mask, ctx, ctx_bf16, stats, 0, B, T, H, dh, stream);
}

extern "C" int msx_attention_tcl_fwd_q0(const float* qkv, const float* mask, void* ctx, int ctx_bf16, float* stats, int q0_only,
                                        int B, int T, int H, int dh, void* stream) {
  return msx_attention_tcl_fwd_p(qkv, mask, ctx, nullptr, ctx_bf16, stats, q0_only, B, T, H, dh, stream);
}

extern "C" int msx_attention_tcl_fwd_p(const float* qkv, const float* mask, void* ctx, void* ctx_lo, int ctx_bf16, float* stats,
                                       int q0_only, int B, int T, int H, int dh, void* stream) {
  MSX_REQUIRE(qkv && mask && ctx && stats, "msx_attention_tcl_fwd: null pointer");
  MSX_REQUIRE(!ctx_lo || (ctx_bf16 && ((uintptr_t)ctx_lo & 15) == 0), "msx_attention_tcl_fwd: ctx_lo needs a bfloat16 ctx and 16-byte alignment");
  MSX_REQUIRE(msx_attention_tcl_supported(qkv, T, dh) && ((uintptr_t)ctx & 15) == 0 && ((uintptr_t)stats & 7) == 0,
              "msx_attention_tcl_fwd: needs d_h == 32, 128 < T <= 768, 16-byte aligned buffers");
  if (B == 0) return MSX_OK;
  const int D = H * DH;
  AttnLongParams p;
  p.qkv = qkv; p.dctx = nullptr;
  p.mask = mask; p.stats = stats; p.out = ctx; p.out_lo = ctx_lo; p.out_bf16 = ctx_bf16 ? 1 : 0; p.dbias = nullptr;
  p.T = T; p.H = H; p.TQ = tensor_queries(T); p.inv_scale = 1.f / sqrtf((float)DH);
  const long long rows = (long long)B * T;
  CUtensorMap tk, tm;
  int rc;
  if ((rc = make_map(&tk, qkv, rows, 3 * D, 3 * D, DH, kTile, false))) return rc;
  if ((rc = make_map(&tm, qkv, rows, 3 * D, 3 * D, DH, kTile, true))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
#define MSX_TCL_FWD(NT_, NKT_, ALIAS_, SMEM_)                                                                              \
  do {                                                                                                                   \
    if (q0_only) {                                                                                                       \
      const size_t q0smem = (ALIAS_) ? (size_t)(SMEM_) : fwd_q0_smem(NT_);                                               \
      MSX_CUDA(cudaFuncSetAttribute(attn_tcl_fwd_kernel<NT_, NKT_, ALIAS_, true>,                                        \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q0smem));                          \
      MSX_CUDA(msx_launch(attn_tcl_fwd_kernel<NT_, NKT_, ALIAS_, true>, dim3(B * H), dim3(kFwdThreads), q0smem, st, tk,  \
                          tm, p));                                                                                       \
    } else {                                                                                                             \
      MSX_CUDA(cudaFuncSetAttribute(attn_tcl_fwd_kernel<NT_, NKT_, ALIAS_, false>,                                       \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_)));                         \
      MSX_CUDA(msx_launch(attn_tcl_fwd_kernel<NT_, NKT_, ALIAS_, false>, dim3(B * H), dim3(kFwdThreads), (SMEM_), st,    \
                          tk, tm, p));                                                                                   \
    }                                                                                                                    \
  } while (0)
#define MSX_TCL_FWD2(NT_, NKT_)                                                                                            \
  do {                                                                                                                   \
    if (q0_only) {                                                                                                       \
      MSX_CUDA(cudaFuncSetAttribute(attn_tcl_fwd2_kernel<NT_, NKT_, true, false>,                                        \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwd2Q0Smem));                     \
      MSX_CUDA(msx_launch(attn_tcl_fwd2_kernel<NT_, NKT_, true, false>, dim3(B * H), dim3(kFwdThreads), kFwd2Q0Smem, st, \
                          tk, tm, p));                                                                                   \
    } else {                                                                                                             \
      constexpr bool kCp = (NT_) <= 4;                                                                                   \
      const size_t sm2 = kCp ? kFwd2CpSmem : kFwd2Smem;                                                                  \
      MSX_CUDA(cudaFuncSetAttribute(attn_tcl_fwd2_kernel<NT_, NKT_, false, kCp>,                                         \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));                             \
      MSX_CUDA(msx_launch(attn_tcl_fwd2_kernel<NT_, NKT_, false, kCp>, dim3(B * H), dim3(kFwdThreads), sm2, st, tk, tm,  \
                          p));                                                                                           \
    }                                                                                                                    \
  } while (0)
  const bool tail = tail_keys(T) != 0;
  const int nt = (T + kTile - 1) / kTile;
  // Full forward, one-pass kernel (one CTA per SM, everything prefetched) vs compact chunked kernel (two CTAs per SM, serial
  // second sweep), per call of 264 k rows: T = 257 1032 / 928 us, T = 260 1298 / 1348, T = 193 1194 / 1243, T = 384 1220 / 1455:
  // the compact kernel wins with one trailing position and few chunks; MSX_TCL_COMPACT=1 / 0 forces it on / off for experiments.
  static const char* cp_env = getenv("MSX_TCL_COMPACT");
  const bool compact3 = cp_env ? cp_env[0] == '1' : tail_keys(T) == 1;
  if (nt == 2) {
    if (tail) MSX_TCL_FWD(2, 1, true, kFwdAliasSmem);  // T = 129 ... 132: one tile on the tensor path, two CTAs per SM
    else if (!q0_only && cp_env && cp_env[0] == '1') MSX_TCL_FWD2(2, 2);
    else MSX_TCL_FWD(2, 2, false, fwd_smem(2));
  } else if (nt == 3) {
    if (tail && !q0_only && compact3) MSX_TCL_FWD2(3, 2);                    // T = 257
    else if (tail) MSX_TCL_FWD(3, 2, false, fwd_smem(3));                     // T = 258 ... 260, q0_only at T = 257 ... 260
    else if (q0_only || (cp_env && cp_env[0] == '1')) MSX_TCL_FWD2(3, 3);     // 384 score columns: chunked, two CTAs per SM
    else MSX_TCL_FWD(3, 3, false, fwd_smem(3));
  } else if (nt == 4) {                                // 384 < T <= 768: two sweeps over the query chunks
    if (tail) MSX_TCL_FWD2(4, 3);
    else MSX_TCL_FWD2(4, 4);
  } else if (nt == 5) {
    if (tail) MSX_TCL_FWD2(5, 4);
    else MSX_TCL_FWD2(5, 5);
  } else {
    if (tail) MSX_TCL_FWD2(6, 5);
    else MSX_TCL_FWD2(6, 6);
  }
#undef MSX_TCL_FWD2
#undef MSX_TCL_FWD
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_attention_tcl_bwd(const float* qkv, const float* mask, const float* dctx, const float* stats, void* dqkv,
                                     int dqkv_bf16, float* dbias, int B, int T, int H, int dh, void* stream) {
  return msx_attention_tcl_bwd_q0(qkv, mask, dctx, stats, dqkv, dqkv_bf16, dbias, 0, B, T, H, dh, stream);
}

extern "C" int msx_attention_tcl_bwd_q0(const float* qkv, const float* mask, const float* dctx, const float* stats, void* dqkv,
                                        int dqkv_bf16, float* dbias, int q0_only, int B, int T, int H, int dh, void* stream) {
  MSX_REQUIRE(qkv && mask && dctx && stats && dqkv, "msx_attention_tcl_bwd: null pointer");
  MSX_REQUIRE(msx_attention_tcl_supported(qkv, T, dh) && ((uintptr_t)dctx & 15) == 0 && ((uintptr_t)dqkv & 15) == 0 &&
                  ((uintptr_t)stats & 7) == 0,
              "msx_attention_tcl_bwd: needs d_h == 32, 128 < T <= 768, 16-byte aligned buffers");
  if (B == 0) return MSX_OK;
  const int D = H * DH;
  AttnLongParams p;
  p.qkv = qkv; p.dctx = dctx;
  p.mask = mask; p.stats = const_cast<float*>(stats); p.out = dqkv; p.out_lo = nullptr; p.out_bf16 = dqkv_bf16 ? 1 : 0; p.dbias = dbias;
  p.T = T; p.H = H; p.TQ = tensor_queries(T); p.inv_scale = 1.f / sqrtf((float)DH);
  const long long rows = (long long)B * T;
  CUtensorMap tk, tm, tdk, tdm;
  int rc;
  if ((rc = make_map(&tk, qkv, rows, 3 * D, 3 * D, DH, kTile, false))) return rc;
  if ((rc = make_map(&tm, qkv, rows, 3 * D, 3 * D, DH, kTile, true))) return rc;
  if ((rc = make_map(&tdk, dctx, rows, D, D, DH, kTile, false))) return rc;
  if ((rc = make_map(&tdm, dctx, rows, D, D, DH, kTile, true))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
#define MSX_TCL_BWD(NT_, NKT_, SINGLE_)                                                                                     \
  do {                                                                                                                   \
    const size_t kBwdSmem = bwd_smem(SINGLE_, NKT_ < NT_, q0_only != 0);                                                 \
    if (q0_only) {                                                                                                       \
      MSX_CUDA(cudaFuncSetAttribute(attn_tcl_bwd_kernel<NT_, NKT_, SINGLE_, true>,                                       \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));                        \
      MSX_CUDA(msx_launch(attn_tcl_bwd_kernel<NT_, NKT_, SINGLE_, true>, dim3(B * H), dim3(kBwdThreads), kBwdSmem, st,   \
                          tk, tm, tdk, tdm, p));                                                                         \
    } else {                                                                                                             \
      MSX_CUDA(cudaFuncSetAttribute(attn_tcl_bwd_kernel<NT_, NKT_, SINGLE_, false>,                                      \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));                        \
      MSX_CUDA(msx_launch(attn_tcl_bwd_kernel<NT_, NKT_, SINGLE_, false>, dim3(B * H), dim3(kBwdThreads), kBwdSmem, st,  \
                          tk, tm, tdk, tdm, p));                                                                         \
    }                                                                                                                    \
  } while (0)
  const bool tail = tail_keys(T) != 0;
  const int nt = (T + kTile - 1) / kTile;
  if (nt == 2) {
    if (tail) MSX_TCL_BWD(2, 1, true);                 // T = 129 ... 132: one tile on the tensor path
    else MSX_TCL_BWD(2, 2, false);
  } else if (nt == 3) {
    if (tail) MSX_TCL_BWD(3, 2, false);                // T = 257 ... 260
    else MSX_TCL_BWD(3, 3, false);
  } else if (nt == 4) {                                // 384 < T <= 768: dQ accumulators fill TMEM columns 64 .. 255
    if (tail) MSX_TCL_BWD(4, 3, false);
    else MSX_TCL_BWD(4, 4, false);
  } else if (nt == 5) {
    if (tail) MSX_TCL_BWD(5, 4, false);
    else MSX_TCL_BWD(5, 5, false);
  } else {
    if (tail) MSX_TCL_BWD(6, 5, false);
    else MSX_TCL_BWD(6, 6, false);
  }
#undef MSX_TCL_BWD
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
