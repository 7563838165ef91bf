// K2c (tensor-core path) — attention in the reference's convention on tcgen05, T <= 128, head width 32 or 16
// (template parameter kDh; 16-wide heads still move 32-wide tiles and reduce over / store the first 16 columns).
//
// Same contract as attention.cu (replaces MultiHeadDotAttention.hybrid_forward lines 91-103 and _mask_logits,
// /root/reference/music_style_transfer/VarAutoEncoder/transformer.py:91-126):
//   S[k][q] = K_k . Q_q / sqrt(d_h) + (key k padded ? -1e9 : 0);  P = softmax over the QUERY axis;  O[q] = sum_k P[k][q] V[k]
//
// One CTA (128 threads) per (batch, head).  TMA loads the K, Q, V head slices straight out of the fused
// [B*T, 3D] projection (TFLOAT32 maps round to nearest).  MMA 1: S[128 keys x TQ queries] = K Q^T accumulates in
// TMEM with the KEYS on the 128 TMEM lanes, so each thread owns one key row and the reference's softmax over
// the query axis is a thread-local loop over TMEM columns (no shuffles, no shared-memory score matrix).  The
// thread writes its normalised row as the MN-major A operand (SWIZZLE_128B_BASE32B, the only layout tcgen05
// takes for MN-major 32-bit data) of MMA 2: O[128 queries x 32] = P^T V, whose accumulator has the QUERIES on
// the lanes: every thread stores one 128-byte row of the context.  Rows beyond T (neighbouring sequence / OOB
// zeros) only ever meet P == 0 or land in output rows that are not stored.
#include "attention_tc_common.cuh"

namespace {

struct AttnTcParams {
  const float* mask;   // [B*T]
  float* ctx;          // [B*T, H*32] fp32, or bf16 when out_bf16 (the operand of the bf16 W_proj GEMM)
  void* ctx_lo;        // optional (out_bf16 only): lo plane, ctx then holds the hi plane (operands of the p3 W_proj GEMM)
  const float* qkv;    // Q0 kernels only: the V rows are read straight from the projection
  int out_bf16;
  int T, H, TQ, TK;    // TQ = roundup16(T) (MMA N), TK = roundup8(T) (reduction length of MMA 2)
  int dh;              // head width: 32, or 16 ("half heads": 32-wide tiles are still loaded, the score MMAs reduce over the
                       // first 16 columns only and only 16 output columns are stored; the other half belongs to the next head)
  int items;           // B * H
  int group_bytes;     // shared memory of one pipeline group
  int smem_bytes;      // dynamic shared memory behind the 1024-byte aligned base
  float inv_scale;
};

// lo part of an fp32 word for the compensated score product: x - (x with the 13 low mantissa bits cleared), rounded to TF32
__device__ __forceinline__ unsigned lo_tf32_bits(unsigned xb) {
  const float x = __uint_as_float(xb);
  const float hi = __uint_as_float(xb & 0xFFFFE000u);
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x - hi));
  return r;
}

__device__ __forceinline__ void group_sync(int g) {       // named barrier of one 128-thread pipeline group
  asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
}

// One lane of a converged warp.  The tcgen05 / TMA instructions take uniform-register operands: issuing them from
// `if (lane == 0)` code makes the compiler wrap every one in an ELECT + R2UR "waterfall" loop (~90 cycles each, measured
// with the clock64 trace), issuing them under elect.sync from warp-uniform values does not.
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Forward.  Persistent CTAs (one per SM), each running G independent 128-thread pipeline groups; a group walks its
// (batch, head) items with its own TMA ring, TMEM slice and mbarriers, so the latency chain of one item
// (TMA -> MMA 1 -> softmax -> MMA 2 -> store) is hidden by the other groups and by the prefetch of the next item.
// Roles inside a group: lane 0 of warp 3 is the ISSUER (all TMA loads, both MMAs; for T <= 96 warp 3 owns no valid
// key row, so the issuer runs ahead of the softmax warps: K, Q of item n+1 are fetched as soon as MMA 1 of item n
// has retired, V is double-buffered); every warp that owns valid rows does softmax + the output rows.
// NCH = TQ / 16: the key row's scores are held in registers (one tcgen05.ld pass, one exp per score).
// X3 (compensated scores): the softmax turns an ABSOLUTE score error into a RELATIVE error of P, so single-pass TF32
// scores (|S| 2^-11) dominate the forward error of the whole step (measured: latent means 6e-4 with TF32 scores vs 1e-5
// with exact ones).  With X3 the K / Q tiles arrive as raw fp32 (FLOAT32 maps), the issuer warp writes their lo parts
// (x - trunc_tf32(x), element-wise, so the TMA swizzle is preserved) to a second pair of tiles and MMA 1 accumulates
// K_lo Q_hi + K_hi Q_lo + K_hi Q_hi (kind::tf32 reads the raw words truncated = hi): fp32-equivalent scores.
// Q0 (the encoder's top layer, whose output is read at position 0 only, model.py:97-100): only the context row of
// query 0 is produced.  The scores and the query-axis softmax still cover every query (each key row's normaliser), but
// O[0] = sum_k P[k][0] V[k] is a reduction over the key rows: no P tile, no MMA 2, no V tile, one 128-byte row stored per
// (batch, head) instead of T rows.
template <int NCH, int G, bool STAGE, int kDh, bool X3, bool Q0 = false>
__global__ void __launch_bounds__(128 * G, 1)
    attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmQ,
                       const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  pdl_entry();
  constexpr int TQ = NCH * 16;
  constexpr int kTmemStride = (TQ + 2 * DH + 31) / 32 * 32;  // per group: O even [0,32) | O odd [32,64) | S [64, 64+TQ)
  constexpr int kTmemCols = G * kTmemStride <= 128 ? 128 : G * kTmemStride <= 256 ? 256 : 512;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ unsigned long long bars_all[G][6];            // kq, v0, v1, s_full, o_full, p_ready
  __shared__ unsigned tmem_slot;
  __shared__ float red_q0[Q0 ? G : 1][2][4][32];           // Q0: per-warp partial sums of P[k][0] V[k], double-buffered

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform as far as the compiler can tell
  const int g = warp >> 2, gt = tid & 127;                 // pipeline group, thread within the group (= key row / query row)
  const int T = p.T, TK = p.TK, D = p.H * kDh;
  const int slab = TK * 128;                               // bytes of a [TK rows][128 B] tile
  unsigned char* sK = base + (size_t)g * p.group_bytes;    // [TK rows][128 B] K-major SW128; MMA 1 addresses 128 rows, the
                                                           // rows beyond TK alias the tiles behind it (their S rows are unused)
  unsigned char* sQ = sK + slab;                           // [TQ rows][128 B] K-major SW128
  unsigned char* sV = sQ + TQ * 128;                       // 2 x [TK rows][128 B] MN-major (d contiguous), SW128_32B
  unsigned char* sP = sV + 2 * slab;                       // ceil(TQ/32) slabs x [TK rows][128 B] MN-major (q contiguous)
  // X3: lo parts of K | Q, same layout as [sK, sQ + TQ rows).  They live only from the conversion to the retirement of MMA 1,
  // P only from the softmax to the retirement of MMA 2 (and the staged output store): the two share the P region when it
  // is large enough (TQ >= 2 * 32 rows of slabs: always for the rows this kernel is used on), so the compensated scores
  // cost no pipeline group.
  unsigned char* sKQlo = (((TQ + 31) / 32) * slab >= (TK + TQ) * 128) ? sP : sP + ((TQ + 31) / 32) * slab;
  unsigned long long* bar_kq = &bars_all[g][0];
  unsigned long long* bar_v = &bars_all[g][1];
  unsigned long long* bar_s = &bars_all[g][3];
  unsigned long long* bar_o = &bars_all[g][4];
  unsigned long long* bar_p = &bars_all[g][5];

  // zero-fill the dynamic shared memory once: MMA descriptors over-address tiles (128 K rows, 4 P / dS^T slabs), and
  // whatever they read must be finite so that the unused accumulator rows never hold NaN / Inf
  for (int i = tid * 16; i < p.smem_bytes; i += blockDim.x * 16) *reinterpret_cast<float4*>(base + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    for (int i = 0; i < G; ++i) {
      for (int j = 0; j < 5; ++j) mbar_init(&bars_all[i][j], 1);
      mbar_init(&bars_all[i][5], 4);                       // p_ready: lane 0 of each of the group's 4 warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_O = tmem_slot + g * kTmemStride, tmem_S = tmem_O + 2 * DH;
  const unsigned lane_off = (unsigned)((warp & 3) * 32) << 16;

  const int first = blockIdx.x * G + g, stride = gridDim.x * G;
  const unsigned idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(TQ >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
  const unsigned idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(DH >> 3) << 17) |
                          ((unsigned)(128 >> 4) << 24);
  const bool issuer = (warp & 3) == 3;                     // the group's warp 3 (warp-uniform), one elected lane issues
  // keys beyond T hold a neighbouring sequence or zeros: a warp whose 32 rows are all >= T (T = 65: warp 3 of every
  // group) skips the softmax arithmetic and the output rows
  const bool warp_live = (warp & 3) * 32 < T;
  // descriptors are fixed per buffer: built once, advanced by constant increments per MMA (a dependent 64-bit
  // shift/or chain per MMA costs the single issuing thread ~90 cycles)
  const unsigned long long dK = make_desc(smem_u32(sK), 16, 1024, 2), dQ = make_desc(smem_u32(sQ), 16, 1024, 2);
  const unsigned long long dP = make_desc(smem_u32(sP), slab, 512, 1), dV = make_desc(smem_u32(sV), slab, 512, 1);
  const unsigned long long kqlo = (unsigned long long)((sKQlo - sK) >> 4);   // descriptor offset of the lo tiles

  if (issuer && first < p.items) {                         // prologue: loads of the group's first item
    const int b = first / p.H, h = first % p.H;
    if (elect_one()) {
      mbar_expect_tx(bar_kq, (unsigned)((TK + TQ) * 128));
      tma_load_2d(sK, &tmK, bar_kq, h * kDh, b * T);
      tma_load_2d(sQ, &tmQ, bar_kq, D + h * kDh, b * T);
      if (!Q0) {
        mbar_expect_tx(&bar_v[0], (unsigned)slab);
        tma_load_2d(sV, &tmV, &bar_v[0], 2 * D + h * kDh, b * T);
      }
    }
    __syncwarp();
  }
  float mraw_next = (gt < T && first < p.items) ? __ldg(p.mask + (size_t)(first / p.H) * T + gt) : 0.f;
  // coalesced context store through the dead P tile (STAGE, chosen by the host): pays off for fp32 rows (16-byte pieces
  // of 32 different lines per store instruction otherwise; measured 167 -> 152 us); bf16 rows are 64 contiguous bytes per
  // lane already and the extra group barrier costs more than the staging saves (138 -> 143 us), so they keep the
  // row-per-lane store.  A template parameter: the staged store costs registers the other variant must not pay for.
  constexpr bool stage_ok = STAGE;
  int n = 0;
  for (int item = first; item < p.items; item += stride, ++n) {
    const int b = item / p.H, h = item % p.H;
    const int nxt = item + stride;
    const unsigned par = (unsigned)(n & 1);
    const float mraw = mraw_next;                          // key-padding mask of this item (prefetched one item ahead)
    if (gt < T && nxt < p.items) mraw_next = __ldg(p.mask + (size_t)(nxt / p.H) * T + gt);
    if (X3) {
      // lo tiles of K | Q (contiguous [TK + TQ rows][128 B]), written by all four warps of the group: one warp alone took
      // ~1.4 us per item for the 19 KB (38 dependent LDS -> cvt -> STS rounds), which sat on every item's critical path
      mbar_wait(bar_kq, par);
      const uint4* src = reinterpret_cast<const uint4*>(sK);
      uint4* dst = reinterpret_cast<uint4*>(sKQlo);
      const int n16 = (TK + TQ) * 8;
#pragma unroll 2
      for (int i = gt; i < n16; i += 128) {
        const uint4 v = src[i];
        dst[i] = make_uint4(lo_tf32_bits(v.x), lo_tf32_bits(v.y), lo_tf32_bits(v.z), lo_tf32_bits(v.w));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      group_sync(g);
    }
    if (issuer) {                                          // whole warp, warp-uniform control flow
      const int b2 = nxt / p.H, h2 = nxt % p.H;
      if (!X3) mbar_wait(bar_kq, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        // MMA 1: S[128 keys x TQ] = K[128 x 32] * Q[TQ x 32]^T, both K-major (+32 B per K = 8 step)
#pragma unroll
        for (int k = 0; k < kDh / 8; ++k) {
          if (X3) {
            umma_tf32(tmem_S, dK + kqlo + 2 * k, dQ + 2 * k, idesc1, k > 0 ? 1u : 0u);     // K_lo Q_hi
            umma_tf32(tmem_S, dK + 2 * k, dQ + kqlo + 2 * k, idesc1, 1u);                  // K_hi Q_lo
            umma_tf32(tmem_S, dK + 2 * k, dQ + 2 * k, idesc1, 1u);                         // K_hi Q_hi
          } else {
            umma_tf32(tmem_S, dK + 2 * k, dQ + 2 * k, idesc1, k > 0 ? 1u : 0u);
          }
        }
        umma_commit(bar_s);
        if (!Q0 && nxt < p.items) {                        // V of the next item into the other V buffer (free since o_full(n-1))
          mbar_expect_tx(&bar_v[(n + 1) & 1], (unsigned)slab);
          tma_load_2d(sV + ((n + 1) & 1) * slab, &tmV, &bar_v[(n + 1) & 1], 2 * D + h2 * kDh, b2 * T);
        }
      }
      __syncwarp();
      mbar_wait(bar_s, par);
      if (nxt < p.items && elect_one()) {                  // K, Q tiles are free once MMA 1 has retired
        mbar_expect_tx(bar_kq, (unsigned)((TK + TQ) * 128));
        tma_load_2d(sK, &tmK, bar_kq, h2 * kDh, b2 * T);
        tma_load_2d(sQ, &tmQ, bar_kq, D + h2 * kDh, b2 * T);
      }
    }
    __syncwarp();
    if (warp_live) {
      const int k = gt;                                    // this thread's key row
      const bool valid = k < T;
      const float rowmask = (valid && mraw > 0.f) ? 0.f : -1e9f;
      float vrow[Q0 ? 32 : 1];
      if (Q0) {                                            // this key's V row, in flight while MMA 1 runs (rows >= T: any finite row)
        const float4* vp = reinterpret_cast<const float4*>(p.qkv + ((size_t)b * T + min(k, T - 1)) * 3 * D + 2 * D + h * kDh);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v4 = __ldg(vp + j);
          vrow[(4 * j) % (Q0 ? 32 : 1)] = v4.x; vrow[(4 * j + 1) % (Q0 ? 32 : 1)] = v4.y;
          vrow[(4 * j + 2) % (Q0 ? 32 : 1)] = v4.z; vrow[(4 * j + 3) % (Q0 ? 32 : 1)] = v4.w;
        }
      }
      mbar_wait(bar_s, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // the key row's TQ scores live in registers: softmax over the QUERY axis is thread-local
      float sc[TQ];
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld16_issue(tmem_S + lane_off + c * 16, sc + c * 16);
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld16_wait(sc + c * 16);
      // every score this loop can see is finite (shared memory is zero-filled at kernel start, so even the rows and
      // columns beyond T, which hold a neighbouring sequence, are), hence rows >= T need no per-element guard: their
      // normaliser is forced to 0.  Only the last 16 columns can lie beyond T.
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < TQ; ++j) {
        sc[j] = fmaf(sc[j], p.inv_scale, rowmask);
        if (j < TQ - 16 || j < T) mx4[j & 3] = fmaxf(mx4[j & 3], sc[j]);
      }
      const float mxl = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * kLog2e;
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < TQ; ++j) {
        sc[j] = exp2f(fmaf(sc[j], kLog2e, -mxl));
        if (j >= TQ - 16 && j >= T) sc[j] = 0.f;
        sum4[j & 3] += sc[j];
      }
      const float inv = valid ? 1.f / ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3])) : 0.f;
      if (Q0) {                                            // this warp's share of O[0] = sum_k P[k][0] V[k]: lane j <- column j
        const float p0 = sc[0] * inv;
        float t[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = vrow[j % (Q0 ? 32 : 1)] * p0;
        red_q0[Q0 ? g : 0][n & 1][warp & 3][lane] = warp_colsum32(t, lane);
      }
      if (!Q0 && k < TK) {                                 // rows T..TK-1 are the zero padding of MMA 2's reduction dimension
#pragma unroll
        for (int j = 0; j < TQ; j += 4)
          *reinterpret_cast<float4*>(sP + mn_major_off(j, k, TK)) =
              make_float4(to_tf32(sc[j] * inv), to_tf32(sc[j + 1] * inv), to_tf32(sc[j + 2] * inv), to_tf32(sc[j + 3] * inv));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (Q0) {
      red_q0[Q0 ? g : 0][n & 1][warp & 3][lane] = 0.f;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);                     // this warp's rows of P are staged (and its reads of S, O(n-1) are done)
    if (Q0) {
      // the issuer may overwrite S (MMA 1 of the next item) once every warp has read its rows: p_ready says so
      if (issuer) mbar_wait(bar_p, par);
      group_sync(g);                                       // the four partial sums are in place
      if ((warp & 3) == 0 && lane < kDh) {
        const float (*r)[32] = red_q0[Q0 ? g : 0][n & 1];
        const float o = (r[0][lane] + r[1][lane]) + (r[2][lane] + r[3][lane]);
        const size_t e = (size_t)b * T * D + h * kDh + lane;
        if (p.out_bf16) {
          unsigned hi, lo;
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(0.f), "f"(o));
          reinterpret_cast<unsigned short*>(p.ctx)[e] = (unsigned short)(hi & 0xFFFFu);
          if (p.ctx_lo) {
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(0.f), "f"(o - __uint_as_float(hi << 16)));
            reinterpret_cast<unsigned short*>(p.ctx_lo)[e] = (unsigned short)(lo & 0xFFFFu);
          }
        } else {
          p.ctx[e] = o;
        }
      }
      continue;                                            // next item: red_q0[(n + 1) & 1]; this buffer is rewritten at n + 2,
                                                           // behind the group barrier of item n + 1
    }
    if (issuer) {
      mbar_wait(bar_p, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mbar_wait(&bar_v[n & 1], (unsigned)((n >> 1) & 1));
      // MMA 2: O[128 queries x 32] = P^T[128 x TK] * V[TK x 32]; A MN-major (queries contiguous), B MN-major (d contiguous)
      const unsigned long long dv = dV + (unsigned long long)((n & 1) * (slab >> 4));
      if (elect_one()) {
        // even and odd k-steps accumulate into two TMEM tiles (two independent chains); the epilogue adds them
#pragma unroll 4
        for (int j = 0; j < TK / 8; ++j)
          umma_tf32(tmem_O + (j & 1) * DH, dP + 64 * j, dv + 64 * j, idesc2, j > 1 ? 1u : 0u);
        umma_commit(bar_o);
      }
    }
    __syncwarp();
    // every warp waits (also the ones without valid rows): it keeps the group in lockstep, one p_ready arrival per
    // warp and phase; for the issuer it also means P and V(n) are free again
    mbar_wait(bar_o, par);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp_live) {
      const int q = gt;
      float o[32], o2[32];
      tmem_ld16_issue(tmem_O + lane_off, o);
      tmem_ld16_issue(tmem_O + lane_off + 16, o + 16);
      if (TK > 8) {                                        // the odd accumulator exists only with >= 2 k-steps
        tmem_ld16_issue(tmem_O + lane_off + DH, o2);
        tmem_ld16_issue(tmem_O + lane_off + DH + 16, o2 + 16);
      }
      tmem_ld16_wait(o);
      tmem_ld16_wait(o + 16);
      if (TK > 8) {
        tmem_ld16_wait(o2);
        tmem_ld16_wait(o2 + 16);
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] += o2[j];
      }
      // P staging (this group's sP) is dead once MMA 2 has retired: the warp's 4 KB slice of it stages the context rows
      // for a coalesced store; very short rows (the slice would not fit) keep the row-per-lane store
      const int w4 = warp & 3;
      if (stage_ok) {
        store_tile32_coalesced(sP + w4 * 4096, p.ctx, p.out_bf16 != 0, ((size_t)b * T + w4 * 32) * D + h * kDh, (size_t)D,
                               T - w4 * 32, o, lane);
      } else if (q < T) {
        if (p.ctx_lo) store_row32_planes(p.ctx, p.ctx_lo, ((size_t)b * T + q) * D + h * kDh, o, kDh);
        else store_row32(p.ctx, p.out_bf16 != 0, ((size_t)b * T + q) * D + h * kDh, o, kDh);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    // the staging slices overlap P rows that OTHER warps of the group write in the next item's softmax
    if (stage_ok) group_sync(g);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "n"(kTmemCols) : "memory");
  }
}


// ------------------------------------------------------------------------------------------------ backward
// S = K Q^T and dP = V dO^T accumulate with the KEYS on the TMEM lanes, so P, delta_k = sum_q P dP and
// dS = P (dP - delta_k) / sqrt(d_h) are thread-local.  P / dS are then written to shared memory as MN-major A
// operands: X[k contiguous][q rows] feeds dV = P dO and dK = dS Q (lanes = keys), Y[q contiguous][k rows] feeds
// dQ = dS^T K (lanes = queries).  The K-major operand tiles of the first two MMAs are dead by then and X aliases them.
struct AttnTcBwdParams {
  const float* mask;   // [B*T]
  float* dqkv;         // [B*T, 3*H*32] fp32, or bf16 when out_bf16 (operand of the bf16 K|Q|V dgrad / wgrad GEMMs)
  const float* qkv;    // Q0 kernels only: V rows and the context gradient of query 0 are read straight from global memory
  const float* dctx;
  int out_bf16;
  float* dbias;        // optional [3*H*32]: += column sums of dqkv (bias gradient of the fused K|Q|V projection)
  int T, H, TQ, TK;    // TQ = roundup16(T), TK = roundup8(T)
  int dh;              // head width 32 or 16 (see AttnTcParams)
  float inv_scale;
  long long* trace;    // optional profiling hook (msx_attention_tc_set_trace): clock64 stamps of block 0's groups
};

template <int kTmemCols, int kDh>
__global__ void __launch_bounds__(128)
    attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmK128, const __grid_constant__ CUtensorMap tmQk,
                       const __grid_constant__ CUtensorMap tmDOk, const __grid_constant__ CUtensorMap tmDOm,
                       const __grid_constant__ CUtensorMap tmQKVm, const AttnTcBwdParams p) {
  pdl_entry();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int T = p.T, TQ = p.TQ, TK = p.TK;
  const int slab = TK * 128;                         // bytes of one 32-wide MN slab with TK reduction rows
  unsigned char* sY = base;                          // dS^T, q contiguous: ceil(TQ/32) slabs written (4 addressed)
  unsigned char* sDOm = sY + ((TQ + 31) / 32) * slab;  // dO  MN-major [TK rows][128 B]
  unsigned char* sQm = sDOm + slab;                  // Q   MN-major
  unsigned char* sKm = sQm + slab;                   // K   MN-major
  unsigned char* sR1 = sKm + slab;                   // phase A: K-major K | V | Q | dO ; phase B: X (4 slabs)
  unsigned char* sKk = sR1;
  unsigned char* sVk = sKk + 128 * 128;
  unsigned char* sQk = sVk + 128 * 128;
  unsigned char* sDOk = sQk + TQ * 128;
  unsigned char* sX = sR1;
  const int r1_bytes = max(2 * 128 * 128 + 2 * TQ * 128, 4 * slab);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sR1 + r1_bytes);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 4);

  __shared__ float red[3 * 128];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int D = p.H * kDh;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = *tmem_slot;
  const unsigned tm_dV = tmem, tm_dK = tmem + 32, tm_dQ = tmem + 64, tm_S = tmem + 96, tm_dP = tmem + 96 + TQ;
  const unsigned idesc_kk = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(TQ >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
  const unsigned idesc_mm = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(DH >> 3) << 17) |
                            ((unsigned)(128 >> 4) << 24);

  if (tid == 0) {
    mbar_expect_tx(&bars[0], (unsigned)((2 * 128 + 2 * TQ + 3 * TK) * 128));
    tma_load_2d(sKk, &tmK128, &bars[0], h * kDh, b * T);
    tma_load_2d(sVk, &tmK128, &bars[0], 2 * D + h * kDh, b * T);
    tma_load_2d(sQk, &tmQk, &bars[0], D + h * kDh, b * T);
    tma_load_2d(sDOk, &tmDOk, &bars[0], h * kDh, b * T);
    tma_load_2d(sDOm, &tmDOm, &bars[0], h * kDh, b * T);
    tma_load_2d(sQm, &tmQKVm, &bars[0], D + h * kDh, b * T);
    tma_load_2d(sKm, &tmQKVm, &bars[0], h * kDh, b * T);
    mbar_wait(&bars[0], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int k = 0; k < kDh / 8; ++k)       // S = K Q^T
      umma_tf32(tm_S, make_desc(smem_u32(sKk) + k * 32, 16, 1024, 2), make_desc(smem_u32(sQk) + k * 32, 16, 1024, 2),
                idesc_kk, k > 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < kDh / 8; ++k)       // dP = V dO^T
      umma_tf32(tm_dP, make_desc(smem_u32(sVk) + k * 32, 16, 1024, 2), make_desc(smem_u32(sDOk) + k * 32, 16, 1024, 2),
                idesc_kk, k > 0 ? 1u : 0u);
    umma_commit(&bars[1]);
  }
  const int k = tid;
  const bool valid = k < T;
  const float rowmask = (valid && __ldg(p.mask + (size_t)b * T + k) > 0.f) ? 0.f : -1e9f;
  __syncwarp();
  mbar_wait(&bars[1], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned lane_off = (unsigned)(warp * 32) << 16;
  float mx = -INFINITY, sum = 0.f, delta = 0.f;
  for (int c = 0; c < TQ; c += 16) {
    float v[16];
    tmem_ld16(tm_S + lane_off + c, v);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (valid && c + j < T) mx = fmaxf(mx, v[j] * p.inv_scale + rowmask);
  }
  for (int c = 0; c < TQ; c += 16) {
    float v[16];
    tmem_ld16(tm_S + lane_off + c, v);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (valid && c + j < T) sum += expf(v[j] * p.inv_scale + rowmask - mx);
  }
  const float inv = valid ? 1.f / sum : 0.f;
  // pass 3: P -> X (k contiguous) for dV, delta_k = sum_q P dP
  for (int c = 0; c < TQ; c += 16) {
    float v[16], g[16];
    tmem_ld16(tm_S + lane_off + c, v);
    tmem_ld16(tm_dP + lane_off + c, g);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float pr = (valid && c + j < T) ? expf(v[j] * p.inv_scale + rowmask - mx) * inv : 0.f;
      delta = fmaf(pr, g[j], delta);
      if (c + j < TK) *reinterpret_cast<float*>(sX + mn_major_off(k, c + j, TK)) = to_tf32(pr);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int j = 0; j < TK / 8; ++j)       // dV[keys x 32] = P[keys x q] dO[q x 32]
      umma_tf32(tm_dV, make_desc(smem_u32(sX) + j * 1024, slab, 512, 1), make_desc(smem_u32(sDOm) + j * 1024, slab, 512, 1),
                idesc_mm, j > 0 ? 1u : 0u);
    umma_commit(&bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);                  // X may be overwritten once dV's MMAs have read it
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // pass 4: dS = P (dP - delta) / sqrt(d_h) -> X (k contiguous, for dK) and Y (q contiguous, for dQ)
  for (int c = 0; c < TQ; c += 16) {
    float v[16], g[16];
    tmem_ld16(tm_S + lane_off + c, v);
    tmem_ld16(tm_dP + lane_off + c, g);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float pr = (valid && c + j < T) ? expf(v[j] * p.inv_scale + rowmask - mx) * inv : 0.f;
      v[j] = to_tf32(pr * (g[j] - delta) * p.inv_scale);
      if (c + j < TK) *reinterpret_cast<float*>(sX + mn_major_off(k, c + j, TK)) = v[j];
    }
    if (k < TK) {
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(sY + mn_major_off(c + j, k, TK)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int j = 0; j < TK / 8; ++j)       // dK[keys x 32] = dS[keys x q] Q[q x 32]
      umma_tf32(tm_dK, make_desc(smem_u32(sX) + j * 1024, slab, 512, 1), make_desc(smem_u32(sQm) + j * 1024, slab, 512, 1),
                idesc_mm, j > 0 ? 1u : 0u);
    for (int j = 0; j < TK / 8; ++j)       // dQ[queries x 32] = dS^T[q x keys] K[keys x 32]
      umma_tf32(tm_dQ, make_desc(smem_u32(sY) + j * 1024, slab, 512, 1), make_desc(smem_u32(sKm) + j * 1024, slab, 512, 1),
                idesc_mm, j > 0 ? 1u : 0u);
    umma_commit(&bars[3]);
  }
  __syncwarp();
  mbar_wait(&bars[3], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    // lanes = keys for dK / dV, lanes = queries for dQ: each thread stores three 128-byte rows
    float o[32];
    const size_t row_elem = ((size_t)b * T + tid) * 3 * D + h * kDh;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const unsigned src = m == 0 ? tm_dK : m == 1 ? tm_dQ : tm_dV;
      tmem_ld16(src + lane_off, o);
      tmem_ld16(src + lane_off + 16, o + 16);
      if (tid < T) store_row32(p.dqkv, p.out_bf16 != 0, row_elem + m * D, o, kDh);
      if (p.dbias) {
        if (tid >= T) {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = 0.f;
        }
        red[m * 128 + warp * 32 + (tid & 31)] = warp_colsum32(o, tid & 31);
      }
    }
    if (p.dbias) {
      __syncthreads();
      if (tid < 96 && (tid & 31) < kDh) {
        const int m = tid >> 5, c = tid & 31;
        atomicAdd(p.dbias + m * D + h * kDh + c, red[m * 128 + c] + red[m * 128 + 32 + c] + red[m * 128 + 64 + c] + red[m * 128 + 96 + c]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------ backward, pipelined (T <= 80)
// Persistent CTAs with G = 2 independent 128-thread pipeline groups (see the forward kernel).  Differences from the
// one-shot kernel above:
//   * S and dP rows are held in registers (one TMEM pass, one exp per score);
//   * P and dS go back into TMEM (tcgen05.st, in place of S and dP) and feed dV = P dO and dK = dS Q as TMEM A operands
//     (lanes = keys, columns = queries is exactly the layout they were computed in), so only dS^T for dQ = dS^T K is
//     staged through shared memory, and all three output MMAs are issued in one batch;
//   * the K-major tiles of the first two MMAs are double-buffered and prefetched one item ahead; the dS^T staging
//     tile aliases the current stage once its MMAs have retired;
//   * the bias-gradient column sums accumulate in registers across the items of a group (flushed when the head changes).
// Q0 (the encoder's top layer under sos_rows_only): dO is non-zero for query 0 only, g = dO[0].  Then dP = V dO^T has one
// non-zero column a_k = V[k] . g, and with c_k = a_k P[k][0]
//     dV[k] = P[k][0] g            dS[k][q] = (-c_k P[k][q] + [q == 0] c_k) / sqrt(d_h)
// are thread-local: no dP MMA, no dV MMA, no dO tiles, no P back in TMEM; two output MMAs (dK, dQ) instead of three.
template <int NCH, bool STAGE, int kDh, bool Q0 = false>
__global__ void __launch_bounds__(256, 1)
    attn_tc_bwd_pipe_kernel(const __grid_constant__ CUtensorMap tmKk, const __grid_constant__ CUtensorMap tmQk,
                            const __grid_constant__ CUtensorMap tmDOk, const __grid_constant__ CUtensorMap tmDOm,
                            const __grid_constant__ CUtensorMap tmQKVm, const AttnTcBwdParams p, const int items,
                            const int group_bytes, const int smem_bytes) {
  pdl_entry();
  constexpr int G = 2;
  constexpr int TQ = NCH * 16;
  constexpr int kTmemStride = 256;                         // dV | dK | dQ | S/P [TQ] | dP/dS [TQ]
  static_assert(96 + 2 * TQ <= kTmemStride, "TMEM budget");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ unsigned long long bars_all[G][6];            // a0, a1, b, mma_a, mma_b, p_ready
  __shared__ float red[G][3 * 128];
  __shared__ unsigned tmem_slot;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform as far as the compiler can tell
  const int g = warp >> 2, gt = tid & 127;
  const int T = p.T, TK = p.TK, D = p.H * kDh;
  const int slab = TK * 128;
  const int stage_bytes = 2 * slab + 2 * TQ * 128;         // Kk | Vk | Qk | dOk (K-major, SW128)
  unsigned char* gbase = base + (size_t)g * group_bytes;
  unsigned char* sBm = gbase + 2 * stage_bytes;            // dOm | Qm | Km (MN-major, SW128_32B), TK rows each
  unsigned long long* bar_a = &bars_all[g][0];
  unsigned long long* bar_b = &bars_all[g][2];
  unsigned long long* bar_ma = &bars_all[g][3];
  unsigned long long* bar_mb = &bars_all[g][4];
  unsigned long long* bar_p = &bars_all[g][5];

  // zero-fill the dynamic shared memory once: MMA descriptors over-address tiles (128 K rows, 4 P / dS^T slabs), and
  // whatever they read must be finite so that the unused accumulator rows never hold NaN / Inf
  for (int i = tid * 16; i < smem_bytes; i += blockDim.x * 16) *reinterpret_cast<float4*>(base + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    for (int i = 0; i < G; ++i) {
      for (int j = 0; j < 5; ++j) mbar_init(&bars_all[i][j], 1);
      mbar_init(&bars_all[i][5], 4);                       // p_ready: lane 0 of each of the group's 4 warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKk) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQk) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDOk) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDOm) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKVm) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "n"(G * kTmemStride)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = tmem_slot + g * kTmemStride;
  const unsigned tm_dV = tmem, tm_dK = tmem + 32, tm_dQ = tmem + 64, tm_S = tmem + 96, tm_dP = tmem + 96 + TQ;
  const unsigned lane_off = (unsigned)((warp & 3) * 32) << 16;
  const unsigned idesc_kk = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(TQ >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
  const unsigned idesc_ts = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((unsigned)(DH >> 3) << 17) |
                            ((unsigned)(128 >> 4) << 24);   // A from TMEM (K-major), B MN-major
  const unsigned idesc_mm = idesc_ts | (1u << 15);          // A MN-major from shared memory, B MN-major

  const int first = blockIdx.x * G + g, stride = gridDim.x * G;
  const bool issuer = (warp & 3) == 3;                      // the group's warp 3, one elected lane issues (see the forward kernel)
  const bool warp_live = (warp & 3) * 32 < T;
  // profiling hook: lane 0 of the issuer warp (stamps 0-3, 6) and thread 0 (stamps 4, 7, 8) of each group of block 0
  long long* trc = (p.trace && blockIdx.x == 0 && (gt == 0 || gt == 96)) ? p.trace + g * 16 * 9 : nullptr;
#define MSX_STAMP(i) do { if (trc && n < 16) trc[n * 9 + (i)] = clock64(); } while (0)

  auto load_a = [&](int item, int stage) {                  // K-major tiles of MMA S and MMA dP
    const int b = item / p.H, h = item % p.H;
    unsigned char* st = gbase + stage * stage_bytes;
    mbar_expect_tx(&bar_a[stage], Q0 ? (unsigned)(slab + TQ * 128) : (unsigned)stage_bytes);
    tma_load_2d(st, &tmKk, &bar_a[stage], h * kDh, b * T);
    if (!Q0) tma_load_2d(st + slab, &tmKk, &bar_a[stage], 2 * D + h * kDh, b * T);
    tma_load_2d(st + 2 * slab, &tmQk, &bar_a[stage], D + h * kDh, b * T);
    if (!Q0) tma_load_2d(st + 2 * slab + TQ * 128, &tmDOk, &bar_a[stage], h * kDh, b * T);
  };
  // descriptors of stage 0 / the MN-major tiles, advanced by constant increments (see the forward kernel)
  const unsigned long long dKk = make_desc(smem_u32(gbase), 16, 1024, 2);
  const unsigned long long dYm = make_desc(smem_u32(gbase), slab, 512, 1);
  const unsigned long long dBm = make_desc(smem_u32(sBm), slab, 512, 1);
  const unsigned slab16 = (unsigned)(slab >> 4), tq16 = (unsigned)(TQ * 128) >> 4, stage16 = (unsigned)(stage_bytes >> 4);

  // bias gradient: every thread sums its own dK / dQ / dV rows over the items of the group (plain adds); the
  // cross-lane column sums (31 shuffles per 32 columns) run once per head change instead of once per item
  float acc[96];
#pragma unroll
  for (int j = 0; j < 96; ++j) acc[j] = 0.f;
  int acc_h = -1;
  auto flush_bias = [&]() {                                 // group-uniform
    if (!p.dbias || acc_h < 0) return;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      float t[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) { t[j] = acc[m * 32 + j]; acc[m * 32 + j] = 0.f; }
      red[g][m * 128 + (warp & 3) * 32 + lane] = warp_colsum32(t, lane);
    }
    group_sync(g);
    if (gt < 96 && (gt & 31) < kDh) {
      const int m = gt >> 5, c = gt & 31;
      atomicAdd(p.dbias + m * D + acc_h * kDh + c,
                red[g][m * 128 + c] + red[g][m * 128 + 32 + c] + red[g][m * 128 + 64 + c] + red[g][m * 128 + 96 + c]);
    }
    group_sync(g);
  };

  if (issuer && first < items) {
    if (elect_one()) load_a(first, 0);
    __syncwarp();
  }
  float mraw_next = (gt < T && first < items) ? __ldg(p.mask + (size_t)(first / p.H) * T + gt) : 0.f;
  // coalesced output store through the dead stage, fp32 outputs only (414 -> 397 us; bf16 outputs: 268 -> 348 us with
  // it, the end-of-item group barrier stalls the pipeline more than the 64-byte-per-lane stores cost)
  // (STAGE is a template parameter: the staged store costs registers the row-per-lane variant must not pay for)
  constexpr bool stage_ok = STAGE;
  int n = 0;
  for (int item = first; item < items; item += stride, ++n) {
    const int b = item / p.H, h = item % p.H;
    const int nxt = item + stride;
    const unsigned par = (unsigned)(n & 1);
    const int stg = n & 1;
    unsigned char* sY = gbase + stg * stage_bytes;          // dS^T (q contiguous) aliases the stage after MMA S / dP retired
    const float mraw = mraw_next;
    if (gt < T && nxt < items) mraw_next = __ldg(p.mask + (size_t)(nxt / p.H) * T + gt);
    if (p.dbias && h != acc_h) {
      flush_bias();
      acc_h = h;
    }
    if (issuer) {                                           // whole warp, warp-uniform control flow
      MSX_STAMP(0);
      if (elect_one()) {
        // MN-major tiles of the output MMAs (free: the previous item's output MMAs have retired)
        mbar_expect_tx(bar_b, (unsigned)((Q0 ? 2 : 3) * slab));
        if (!Q0) tma_load_2d(sBm, &tmDOm, bar_b, h * kDh, b * T);
        tma_load_2d(sBm + slab, &tmQKVm, bar_b, D + h * kDh, b * T);
        tma_load_2d(sBm + 2 * slab, &tmQKVm, bar_b, h * kDh, b * T);
        if (nxt < items) load_a(nxt, stg ^ 1);              // the other stage held item n-1 (its MMAs and dS^T are consumed)
      }
      __syncwarp();
      mbar_wait(&bar_a[stg], (unsigned)((n >> 1) & 1));
      MSX_STAMP(1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned long long kk = dKk + (unsigned long long)(stg * stage16), vk = kk + slab16, qk = vk + slab16, dok = qk + tq16;
      if (elect_one()) {
        // S = K Q^T and dP = V dO^T: two independent accumulate chains, interleaved
#pragma unroll
        for (int k = 0; k < kDh / 8; ++k) {
          umma_tf32(tm_S, kk + 2 * k, qk + 2 * k, idesc_kk, k > 0 ? 1u : 0u);
          if (!Q0) umma_tf32(tm_dP, vk + 2 * k, dok + 2 * k, idesc_kk, k > 0 ? 1u : 0u);
        }
        umma_commit(bar_ma);
      }
      __syncwarp();
      MSX_STAMP(2);
    }
    __syncwarp();
    float p0_keep = 0.f;
    if (warp_live) {
      const int k = gt;
      const bool valid = k < T;
      const float rowmask = (valid && mraw > 0.f) ? 0.f : -1e9f;
      float a_k = 0.f;
      if (Q0) {                                             // g = dO[0] of this (batch, head) and a_k = V[k] . g, in flight during MMA S
        const float4* gp4 = reinterpret_cast<const float4*>(p.dctx + (size_t)b * T * D + h * kDh);
        const float4* vp4 = reinterpret_cast<const float4*>(p.qkv + ((size_t)b * T + min(k, T - 1)) * 3 * D + 2 * D + h * kDh);
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g4 = __ldg(gp4 + j), v4 = __ldg(vp4 + j);
          d4[0] = fmaf(g4.x, v4.x, d4[0]); d4[1] = fmaf(g4.y, v4.y, d4[1]);
          d4[2] = fmaf(g4.z, v4.z, d4[2]); d4[3] = fmaf(g4.w, v4.w, d4[3]);
        }
        a_k = (d4[0] + d4[1]) + (d4[2] + d4[3]);
      }
      mbar_wait(bar_ma, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // the key row's scores stay in registers; dP is re-read from TMEM in 16-column chunks (two cheap passes)
      float sc[TQ];
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld16_issue(tm_S + lane_off + c * 16, sc + c * 16);
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld16_wait(sc + c * 16);
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < TQ; ++j) {                        // finite inputs, tail-only guards: see the forward kernel
        sc[j] = fmaf(sc[j], p.inv_scale, rowmask);
        if (j < TQ - 16 || j < T) mx4[j & 3] = fmaxf(mx4[j & 3], sc[j]);
      }
      const float mxl = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * kLog2e;
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < TQ; ++j) {
        sc[j] = exp2f(fmaf(sc[j], kLog2e, -mxl));
        if (j >= TQ - 16 && j >= T) sc[j] = 0.f;
        sum4[j & 3] += sc[j];
      }
      const float inv = valid ? 1.f / ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3])) : 0.f;
      if (Q0) {
        const float p0 = sc[0] * inv, ck = a_k * p0;
        p0_keep = p0;                                       // dV[k] = P[k][0] g leaves with the other two rows in the epilogue
        const float cs = -ck * inv * p.inv_scale;           // dS[k][q] = cs * e[q] (+ ck / sqrt(d_h) for q == 0)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          float gp[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) gp[j] = to_tf32(fmaf(cs, sc[c * 16 + j], (c == 0 && j == 0) ? ck * p.inv_scale : 0.f));
          tmem_st16(tm_dP + lane_off + c * 16, gp);
          if (k < TK) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(sY + mn_major_off(c * 16 + j, k, TK)) = make_float4(gp[j], gp[j + 1], gp[j + 2], gp[j + 3]);
          }
        }
      }
      float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < (Q0 ? 0 : NCH); ++c) {
        float gp[16];
        tmem_ld16_issue(tm_dP + lane_off + c * 16, gp);
        tmem_ld16_wait(gp);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          sc[c * 16 + j] *= inv;                            // P
          d4[j & 3] = fmaf(sc[c * 16 + j], gp[j], d4[j & 3]);
        }
      }
      const float delta = (d4[0] + d4[1]) + (d4[2] + d4[3]);
#pragma unroll
      for (int c = 0; c < (Q0 ? 0 : NCH); ++c) {
        float gp[16];
        tmem_ld16_issue(tm_dP + lane_off + c * 16, gp);
        tmem_ld16_wait(gp);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          gp[j] = to_tf32(sc[c * 16 + j] * ((gp[j] - delta) * p.inv_scale));   // dS
          sc[c * 16 + j] = to_tf32(sc[c * 16 + j]);
        }
        tmem_st16(tm_dP + lane_off + c * 16, gp);
        tmem_st16(tm_S + lane_off + c * 16, sc + c * 16);
        if (k < TK) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(sY + mn_major_off(c * 16 + j, k, TK)) = make_float4(gp[j], gp[j + 1], gp[j + 2], gp[j + 3]);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      if (gt == 0) MSX_STAMP(4);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);                      // P, dS (TMEM) and dS^T (shared) rows of this warp are staged
    if (issuer) {
      mbar_wait(bar_p, par);
      MSX_STAMP(3);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mbar_wait(bar_b, par);
      const unsigned long long ym = dYm + (unsigned long long)(stg * stage16);
      const unsigned long long dom = dBm, qm = dBm + slab16, km = qm + slab16;
      const int nk = TK / 8;
      if (elect_one()) {
        // three independent accumulate chains, interleaved:
        //   dV[keys x 32] = P[keys x q] dO[q x 32],  dK[keys x 32] = dS[keys x q] Q[q x 32]  (A from TMEM)
        //   dQ[queries x 32] = dS^T[q x keys] K[keys x 32]                                     (A from shared memory)
#pragma unroll 3
        for (int j = 0; j < nk; ++j) {
          if (!Q0) umma_tf32_ts(tm_dV, tm_S + j * 8, dom + 64 * j, idesc_ts, j > 0 ? 1u : 0u);
          umma_tf32_ts(tm_dK, tm_dP + j * 8, qm + 64 * j, idesc_ts, j > 0 ? 1u : 0u);
          umma_tf32(tm_dQ, ym + 64 * j, km + 64 * j, idesc_mm, j > 0 ? 1u : 0u);
        }
        umma_commit(bar_mb);
      }
      __syncwarp();
      MSX_STAMP(6);
    }
    __syncwarp();
    mbar_wait(bar_mb, par);                                 // every warp: keeps the group in lockstep (see the forward kernel)
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (gt == 0) MSX_STAMP(7);
    if (warp_live) {
      // lanes = keys for dK / dV, lanes = queries for dQ: each thread stores three 128-byte rows
      const size_t row_elem = ((size_t)b * T + gt) * 3 * D + h * kDh;
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        float o[32];
        const unsigned src = m == 0 ? tm_dK : m == 1 ? tm_dQ : tm_dV;
        if (Q0 && m == 2) {                                 // dV[k] = P[k][0] g (g: a 128-byte row every thread of the item reads)
          const float4* gp4 = reinterpret_cast<const float4*>(p.dctx + (size_t)b * T * D + h * kDh);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g4 = __ldg(gp4 + j);
            o[4 * j] = p0_keep * g4.x; o[4 * j + 1] = p0_keep * g4.y; o[4 * j + 2] = p0_keep * g4.z; o[4 * j + 3] = p0_keep * g4.w;
          }
        } else {
          tmem_ld16_issue(src + lane_off, o);
          tmem_ld16_issue(src + lane_off + 16, o + 16);
          tmem_ld16_wait(o);
          tmem_ld16_wait(o + 16);
        }
        // the stage (K-major tiles / dS^T) is dead once the output MMAs have retired: the warp's 4 KB slice of it stages
        // the rows for a coalesced store
        if (stage_ok) {
          const int w4 = warp & 3;
          store_tile32_coalesced(sY + w4 * 4096, p.dqkv, p.out_bf16 != 0, ((size_t)b * T + w4 * 32) * 3 * D + h * kDh + m * D,
                                 (size_t)3 * D, T - w4 * 32, o, lane);
        } else if (gt < T) {
          store_row32(p.dqkv, p.out_bf16 != 0, row_elem + m * D, o, kDh);
        }
        if (gt < T && p.dbias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[m * 32 + j] += o[j];
        }
      }
      if (stage_ok) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // staging (generic proxy) before the next TMA fill
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (gt == 0) MSX_STAMP(8);
    // the issuer refills this stage (item n + 2) at the top of the next item: every warp must be done staging in it
    if (stage_ok) group_sync(g);
  }
#undef MSX_STAMP
  flush_bias();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "n"(G * kTmemStride) : "memory");
  }
}

}  // namespace

extern "C" int msx_attention_tc_supported(const float* qkv, int T, int dh) {
  // dh == 16 ("half heads", e.g. the 128-wide Transformer decoder with 8 heads): the kernels still move 32-wide tiles
  // and reduce / store only the first 16 columns, see AttnTcParams::dh
  return (qkv && (dh == 32 || dh == 16) && T >= 1 && T <= 128 && ((uintptr_t)qkv & 15) == 0) ? 1 : 0;
}

namespace {
template <int NCH, int G, bool X3>
int launch_fwd(const CUtensorMap& tk, const CUtensorMap& tq, const CUtensorMap& tv, AttnTcParams p, cudaStream_t st, bool q0 = false) {
  // the last group's MMA descriptors address 128 K rows / 4 P slabs: keep the tail inside the allocation
  const size_t smem = 1024 + (size_t)G * p.group_bytes + 16 * 1024;
  p.smem_bytes = (int)smem - 1024;
  MSX_REQUIRE(smem <= 227 * 1024, "msx_attention_tc_fwd: shared memory budget exceeded (T=%d)", p.T);
  const int want = (p.items + G - 1) / G;
  const int grid = want < msx_num_sms() ? want : msx_num_sms();
  // fp32 context rows leave through the dead P tile (coalesced) when the per-warp staging slices fit into it
  const bool stage = !p.out_bf16 && p.dh == 32 && ((p.T + 31) / 32) * 4096 <= ((p.TQ + 31) / 32) * p.TK * 128;
  if (q0) {
    MSX_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<NCH, G, false, 32, X3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSX_CUDA(msx_launch(attn_tc_fwd_kernel<NCH, G, false, 32, X3, true>, dim3(grid), dim3(128 * G), smem, st, tk, tq, tv, p));
  } else if (p.dh == 16) {
    MSX_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<NCH, G, false, 16, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSX_CUDA(msx_launch(attn_tc_fwd_kernel<NCH, G, false, 16, X3>, dim3(grid), dim3(128 * G), smem, st, tk, tq, tv, p));
  } else if (stage) {
    MSX_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<NCH, G, true, 32, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSX_CUDA(msx_launch(attn_tc_fwd_kernel<NCH, G, true, 32, X3>, dim3(grid), dim3(128 * G), smem, st, tk, tq, tv, p));
  } else {
    MSX_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<NCH, G, false, 32, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSX_CUDA(msx_launch(attn_tc_fwd_kernel<NCH, G, false, 32, X3>, dim3(grid), dim3(128 * G), smem, st, tk, tq, tv, p));
  }
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
}  // namespace

extern "C" int msx_attention_tc_fwd_p(const float* qkv, const float* mask, void* ctx, void* ctx_lo, int ctx_bf16, int x3_scores,
                                      int q0_only, int B, int T, int H, int dh, void* stream);
extern "C" int msx_attention_tc_fwd_ex2(const float* qkv, const float* mask, void* ctx, int ctx_bf16, int x3_scores, int B,
                                        int T, int H, int dh, void* stream) {
  return msx_attention_tc_fwd_p(qkv, mask, ctx, nullptr, ctx_bf16, x3_scores, 0, B, T, H, dh, stream);
}
extern "C" int msx_attention_tc_fwd_ex(const float* qkv, const float* mask, void* ctx, int ctx_bf16, int B, int T, int H,
                                       int dh, void* stream) {
  return msx_attention_tc_fwd_ex2(qkv, mask, ctx, ctx_bf16, 0, B, T, H, dh, stream);
}
extern "C" int msx_attention_tc_fwd(const float* qkv, const float* mask, float* ctx, int B, int T, int H, int dh,
                                    void* stream) {
  return msx_attention_tc_fwd_ex2(qkv, mask, ctx, 0, 0, B, T, H, dh, stream);
}

extern "C" int msx_attention_tc_fwd_p(const float* qkv, const float* mask, void* ctx, void* ctx_lo, int ctx_bf16, int x3_scores,
                                      int q0_only, int B, int T, int H, int dh, void* stream) {
  MSX_REQUIRE(qkv && mask && ctx, "msx_attention_tc_fwd: null pointer");
  MSX_REQUIRE(!q0_only || dh == 32, "msx_attention_tc_fwd_p: q0_only needs d_h == 32");
  MSX_REQUIRE(!ctx_lo || (ctx_bf16 && ((uintptr_t)ctx_lo & 15) == 0), "msx_attention_tc_fwd_p: the lo plane needs a bf16 ctx and 16-byte alignment");
  MSX_REQUIRE(((uintptr_t)ctx & 15) == 0, "msx_attention_tc_fwd: ctx must be 16-byte aligned");
  MSX_REQUIRE(msx_attention_tc_supported(qkv, T, dh), "msx_attention_tc_fwd: needs d_h == 32, T <= 128, 16-byte aligned qkv");
  if (B == 0) return MSX_OK;
  const int D = H * dh;
  AttnTcParams p;
  p.mask = mask; p.ctx = reinterpret_cast<float*>(ctx); p.ctx_lo = ctx_lo; p.qkv = qkv; p.out_bf16 = ctx_bf16 ? 1 : 0; p.T = T; p.H = H; p.dh = dh;
  p.TQ = (T + 15) / 16 * 16;
  p.TK = (T + 7) / 8 * 8;
  p.items = B * H;
  p.inv_scale = 1.f / sqrtf((float)dh);
  // per group: K [TK] + Q [TQ] + 2 x V [TK] + P [ceil(TQ/32) slabs x TK] rows of 128 B, every tile 1024-byte aligned
  //            (+ the lo parts of K and Q for compensated scores when they do not fit into the P region, which they share)
  const bool lo_in_p = ((p.TQ + 31) / 32) * p.TK >= p.TK + p.TQ;
  p.group_bytes = (3 * p.TK + p.TQ + ((p.TQ + 31) / 32) * p.TK + ((x3_scores && !lo_in_p) ? p.TK + p.TQ : 0)) * 128;
  const long long rows = (long long)B * T;
  CUtensorMap tk, tq, tv;
  int rc;
  if ((rc = make_map(&tk, qkv, rows, 3 * D, 3 * D, DH, p.TK, false, x3_scores != 0))) return rc;
  if ((rc = make_map(&tq, qkv, rows, 3 * D, 3 * D, DH, p.TQ, false, x3_scores != 0))) return rc;
  if ((rc = make_map(&tv, qkv, rows, 3 * D, 3 * D, DH, p.TK, true))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (x3_scores) {                           // same group counts: the lo tiles share the P region (short rows: they are tiny)
    switch (p.TQ / 16) {
      case 1: return launch_fwd<1, 3, true>(tk, tq, tv, p, st, q0_only != 0);
      case 2: return launch_fwd<2, 3, true>(tk, tq, tv, p, st, q0_only != 0);
      case 3: return launch_fwd<3, 3, true>(tk, tq, tv, p, st, q0_only != 0);
      case 4: return launch_fwd<4, 3, true>(tk, tq, tv, p, st, q0_only != 0);
      case 5: return launch_fwd<5, 3, true>(tk, tq, tv, p, st, q0_only != 0);
      case 6: return launch_fwd<6, 2, true>(tk, tq, tv, p, st, q0_only != 0);
      case 7: return launch_fwd<7, 1, true>(tk, tq, tv, p, st, q0_only != 0);
      default: return launch_fwd<8, 1, true>(tk, tq, tv, p, st, q0_only != 0);
    }
  }
  switch (p.TQ / 16) {
    case 1: return launch_fwd<1, 3, false>(tk, tq, tv, p, st, q0_only != 0);
    case 2: return launch_fwd<2, 3, false>(tk, tq, tv, p, st, q0_only != 0);
    case 3: return launch_fwd<3, 3, false>(tk, tq, tv, p, st, q0_only != 0);
    case 4: return launch_fwd<4, 3, false>(tk, tq, tv, p, st, q0_only != 0);
    case 5: return launch_fwd<5, 3, false>(tk, tq, tv, p, st, q0_only != 0);
    case 6: return launch_fwd<6, 2, false>(tk, tq, tv, p, st, q0_only != 0);
    case 7: return launch_fwd<7, 1, false>(tk, tq, tv, p, st, q0_only != 0);
    default: return launch_fwd<8, 1, false>(tk, tq, tv, p, st, q0_only != 0);
  }
}

namespace { long long* g_attn_trace = nullptr; }
// Profiling hook: device buffer of 2 x 16 x 9 int64 receiving clock64 stamps of the pipelined backward kernel
// (block 0, per group, first 16 items: loop top, tiles landed, MMA S/dP issued, S/dP ready, softmax done,
// group barrier, output MMAs issued, outputs ready, stores issued).  nullptr disables.
extern "C" int msx_attention_tc_set_trace(long long* buf) {
  g_attn_trace = buf;
  return MSX_OK;
}

extern "C" int msx_attention_tc_bwd_ex(const float* qkv, const float* mask, const float* dctx, void* dqkv, int dqkv_bf16,
                                       float* dbias, int B, int T, int H, int dh, void* stream);
extern "C" int msx_attention_tc_bwd_q0(const float* qkv, const float* mask, const float* dctx, void* dqkv, int dqkv_bf16,
                                       float* dbias, int q0_only, int B, int T, int H, int dh, void* stream);
extern "C" int msx_attention_tc_bwd(const float* qkv, const float* mask, const float* dctx, float* dqkv, float* dbias,
                                    int B, int T, int H, int dh, void* stream) {
  return msx_attention_tc_bwd_ex(qkv, mask, dctx, dqkv, 0, dbias, B, T, H, dh, stream);
}

extern "C" int msx_attention_tc_bwd_ex(const float* qkv, const float* mask, const float* dctx, void* dqkv, int dqkv_bf16,
                                       float* dbias, int B, int T, int H, int dh, void* stream) {
  return msx_attention_tc_bwd_q0(qkv, mask, dctx, dqkv, dqkv_bf16, dbias, 0, B, T, H, dh, stream);
}

extern "C" int msx_attention_tc_bwd_q0(const float* qkv, const float* mask, const float* dctx, void* dqkv, int dqkv_bf16,
                                       float* dbias, int q0_only, int B, int T, int H, int dh, void* stream) {
  MSX_REQUIRE(qkv && mask && dctx && dqkv, "msx_attention_tc_bwd: null pointer");
  MSX_REQUIRE(msx_attention_tc_supported(qkv, T, dh) && ((uintptr_t)dctx & 15) == 0 && ((uintptr_t)dqkv & 15) == 0,
              "msx_attention_tc_bwd: needs d_h == 32, T <= 128, 16-byte aligned buffers");
  if (B == 0) return MSX_OK;
  const int D = H * dh;
  AttnTcBwdParams p;
  p.mask = mask; p.dqkv = reinterpret_cast<float*>(dqkv); p.out_bf16 = dqkv_bf16 ? 1 : 0; p.dbias = dbias; p.T = T; p.H = H;
  p.dh = dh; p.qkv = qkv; p.dctx = dctx;
  p.trace = g_attn_trace;
  p.TQ = (T + 15) / 16 * 16;
  p.TK = (T + 7) / 8 * 8;
  p.inv_scale = 1.f / sqrtf((float)dh);
  const long long rows = (long long)B * T;
  CUtensorMap tK128, tQk, tDOk, tDOm, tQKVm;
  int rc;
  if ((rc = make_map(&tK128, qkv, rows, 3 * D, 3 * D, DH, 128, false))) return rc;
  if ((rc = make_map(&tQk, qkv, rows, 3 * D, 3 * D, DH, p.TQ, false))) return rc;
  if ((rc = make_map(&tDOk, dctx, rows, D, D, DH, p.TQ, false))) return rc;
  if ((rc = make_map(&tDOm, dctx, rows, D, D, DH, p.TK, true))) return rc;
  if ((rc = make_map(&tQKVm, qkv, rows, 3 * D, 3 * D, DH, p.TK, true))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.TQ <= 80) {
    // pipelined kernel: K-major tiles carry TK rows only (MMA rows beyond TK alias the following tiles, their outputs are unused)
    CUtensorMap tKk;
    if ((rc = make_map(&tKk, qkv, rows, 3 * D, 3 * D, DH, p.TK, false))) return rc;
    const int slab = p.TK * 128;
    const int group_bytes = 2 * (2 * slab + 2 * p.TQ * 128) + 3 * slab;
    // MMA descriptors over-address: 128 rows of the K / V tiles, 4 slabs of dS^T.  For T = 65 that stays inside the
    // stage; for short sequences the last group's last stage needs the allocation extended.
    const int stage_bytes = 2 * slab + 2 * p.TQ * 128;
    const int extent = max(slab + 128 * 128, 4 * slab);
    const size_t smem = 1024 + (size_t)max(2 * group_bytes, group_bytes + stage_bytes + extent);
    MSX_REQUIRE(smem <= 220 * 1024, "msx_attention_tc_bwd: shared memory budget exceeded (T=%d)", T);
    const int items = B * H;
    const int want = (items + 1) / 2;
    const int grid = want < msx_num_sms() ? want : msx_num_sms();
    // fp32 outputs leave through a shared-memory staging tile (coalesced rows) when it fits into a pipeline stage
    const bool stage = !p.out_bf16 && dh == 32 && ((T + 31) / 32) * 4096 <= stage_bytes;
#define MSX_BWD_PIPE(NCH)                                                                                              \
  case NCH:                                                                                                            \
    if (q0_only && dh == 32 && stage) {                                                                                \
      MSX_CUDA(cudaFuncSetAttribute(attn_tc_bwd_pipe_kernel<NCH, true, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      MSX_CUDA(msx_launch(attn_tc_bwd_pipe_kernel<NCH, true, 32, true>, dim3(grid), dim3(256), smem, st, tKk, tQk, tDOk, tDOm, tQKVm, p, items, group_bytes, (int)smem - 1024)); \
    } else if (q0_only && dh == 32) {                                                                                  \
      MSX_CUDA(cudaFuncSetAttribute(attn_tc_bwd_pipe_kernel<NCH, false, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      MSX_CUDA(msx_launch(attn_tc_bwd_pipe_kernel<NCH, false, 32, true>, dim3(grid), dim3(256), smem, st, tKk, tQk, tDOk, tDOm, tQKVm, p, items, group_bytes, (int)smem - 1024)); \
    } else if (dh == 16) {                                                                                                    \
      MSX_CUDA(cudaFuncSetAttribute(attn_tc_bwd_pipe_kernel<NCH, false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      MSX_CUDA(msx_launch(attn_tc_bwd_pipe_kernel<NCH, false, 16>, dim3(grid), dim3(256), smem, st, tKk, tQk, tDOk, tDOm, tQKVm, p, items, group_bytes, (int)smem - 1024)); \
    } else if (stage) {                                                                                                \
      MSX_CUDA(cudaFuncSetAttribute(attn_tc_bwd_pipe_kernel<NCH, true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      MSX_CUDA(msx_launch(attn_tc_bwd_pipe_kernel<NCH, true, 32>, dim3(grid), dim3(256), smem, st, tKk, tQk, tDOk, tDOm, tQKVm, p, items, group_bytes, (int)smem - 1024)); \
    } else {                                                                                                           \
      MSX_CUDA(cudaFuncSetAttribute(attn_tc_bwd_pipe_kernel<NCH, false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      MSX_CUDA(msx_launch(attn_tc_bwd_pipe_kernel<NCH, false, 32>, dim3(grid), dim3(256), smem, st, tKk, tQk, tDOk, tDOm, tQKVm, p, items, group_bytes, (int)smem - 1024)); \
    }                                                                                                                  \
    break;
    switch (p.TQ / 16) {
      MSX_BWD_PIPE(1)
      MSX_BWD_PIPE(2)
      MSX_BWD_PIPE(3)
      MSX_BWD_PIPE(4)
      default:
      MSX_BWD_PIPE(5)
    }
#undef MSX_BWD_PIPE
    MSX_LAUNCH_CHECK();
    return MSX_OK;
  }
  const int slab = p.TK * 128;
  const int r1 = max(2 * 128 * 128 + 2 * p.TQ * 128, 4 * slab);
  const size_t smem = 1024 + (size_t)((p.TQ + 31) / 32) * slab + 3 * (size_t)slab + r1 + 64;
  MSX_REQUIRE(smem <= 227 * 1024, "msx_attention_tc_bwd: shared memory budget exceeded (T=%d)", T);
  if (96 + 2 * p.TQ <= 256) {
    if (dh == 16) {
      MSX_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<256, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MSX_CUDA(msx_launch(attn_tc_bwd_kernel<256, 16>, dim3(B * H), dim3(128), smem, st, tK128, tQk, tDOk, tDOm, tQKVm, p));
    } else {
      MSX_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<256, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MSX_CUDA(msx_launch(attn_tc_bwd_kernel<256, 32>, dim3(B * H), dim3(128), smem, st, tK128, tQk, tDOk, tDOm, tQKVm, p));
    }
  } else {
    if (dh == 16) {
      MSX_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<512, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MSX_CUDA(msx_launch(attn_tc_bwd_kernel<512, 16>, dim3(B * H), dim3(128), smem, st, tK128, tQk, tDOk, tDOm, tQKVm, p));
    } else {
      MSX_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<512, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MSX_CUDA(msx_launch(attn_tc_bwd_kernel<512, 32>, dim3(B * H), dim3(128), smem, st, tK128, tQk, tDOk, tDOm, tQKVm, p));
    }
  }
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
