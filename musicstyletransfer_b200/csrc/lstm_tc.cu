// K2g (tensor path) — persistent LSTM recurrence with the h W_h2h^T products on the tensor cores.
//
// Same contract as lstm.cu (replaces the fused gluon.rnn.LSTM of LSTMDecoder.forward_train,
// /root/reference/music_style_transfer/VarAutoEncoder/model.py:148-153,179), H = 128 only.  The recurrence is T
// sequential [R x 128] x [128 x 512] products; at R = 32 rows per cluster and 65 steps it is latency- and
// FFMA-bound in lstm.cu (~45 % of the fp32 pipe).  Here a 2-CTA cluster still owns R = 32 batch rows and each CTA
// the four gates of half of the hidden units, but
//   * the CTA's 256 x 128 slice of W_h2h lives in REGISTERS for all T steps as TF32 B fragments of
//     mma.sync.m16n8k8 (128 registers per thread), so a step reads no weights at all;
//   * h_{t-1} (forward) / d(pre-activation) (backward) sit in shared memory in A-fragment order: one LDS.128
//     yields the four A registers of an MMA;
//   * a warp's accumulator tile holds all four gates of the same (row, unit) cells, so the cell update is
//     thread-local; the new h goes to both CTAs' fragment buffers: local st.shared + st.async into the peer, completion
//     counted on an mbarrier in the receiver (no cluster barrier on the recurrence, see the helpers below).
// Why mma.sync and not tcgen05: UMMA needs M >= 64 batch rows (or the gates on the M axis with a 4-CTA weight
// split); with B = 2048 that leaves 16-32 CTAs busy.  mma.sync.m16n8k8.tf32 measured 480 FMA/clk/SM on B200
// (profiles/micro/mma_sync_rate.cu), 3.75x the FFMA pipe, and lets 128 CTAs share the work.
#include <cooperative_groups.h>

#include "msx_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int H = 128, UH = 64, R = 32;   // hidden units, units per CTA, batch rows per cluster
constexpr int kThreads = 256;
constexpr int kRowPitch = 4 * UH + 8;     // floats per row of a staged [R x 4 x 64] gate slab (+8: conflict-free float2 reads)
constexpr int kColPitch = UH + 8;         // floats per row of a staged [R x 64] state slab

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float tf32r(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const float4& a, float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
                 "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// Per-step exchange between the two CTAs of a cluster WITHOUT a cluster barrier: the halves of the new state travel as
// st.async stores that complete transaction bytes on an mbarrier in the RECEIVER's shared memory, the local half is
// covered by one mbarrier.arrive per warp.  A barrier.cluster.arrive.release (the round-1 scheme) also has to make the
// step's GLOBAL stores (56 KB per CTA) visible at cluster scope before it completes, which put their drain (0.9 us of a
// 3.5 us step) on the critical path of the recurrence; an mbarrier.arrive releases at CTA scope only and st.async carries
// its own completion, so the global stores drain behind the next steps' MMAs.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned peer_addr(unsigned local, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_v4(unsigned addr, const float4& v, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void st_async_v2(unsigned addr, const float2& v, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(addr), "f"(v.x),
               "f"(v.y), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}

// A operand [32 rows x K] in m16n8k8 fragment order: [m-tile][k-step][lane][4].  The eight k of a step are assigned to
// the fragment slots as k % 8 = 2 * (lane % 4) + (register / 2) (the MMA sums over k, so any assignment used for both
// operands is valid); that way the two units a thread owns (accumulator columns 2t, 2t+1) for rows g and g+8 form
// exactly one float4 of the fragment buffer.
__device__ __forceinline__ int afrag_index(int row, int k, int ksteps) {
  const int m = row >> 4, rr = row & 15, s = k >> 3, kk = k & 7;
  return ((m * ksteps + s) * 32 + (rr & 7) * 4 + (kk >> 1)) * 4 + (rr >> 3) + 2 * (kk & 1);
}

// ------------------------------------------------------------------------------------ forward
// gx [B,T,4H] in: x W_i2h^T + b_i2h ; out: gate activations (i,f,g,o).  hs / hprev / cs [B,T,H].
//
// Schedule.  The R = 32 rows of a cluster are two independent SUB-BLOCKS of 16 rows (one m16 tile each), and the eight
// warps two GROUPS (warps 0-3 / 4-7: one warp of each group per SM sub-partition).  Every warp walks
//     M(A,t) G(A,t) M(B,t) G(B,t) M(A,t+1) ...          M = 64 MMAs of a sub-block, G = gate math + exchange + stores
// and the second group starts one phase late, so on each sub-partition one warp is in its MMA phase (tensor pipe) while
// the other is in its gate phase (MUFU, LSU): the phases of a step no longer add up, and the state of a sub-block has the
// whole next phase to reach the peer CTA before anyone waits for it.  (With all warps in one phase the step cost
// MMA 0.9 + MUFU 0.3 + stores 0.7 + exchange / waits 1.0 us = 3.1 us, measured with the phases switched off one by one.)
constexpr int RB = 16;                   // rows of a sub-block
struct FwdSmem {
  float hfrag[2][2][RB * H];             // [sub-block][parity]: h_{t-1} in A-fragment order (4 x 8 KB)
  float gxs[2][3][RB * kRowPitch];       // [sub-block][t % 3]: this CTA's slice of gx (cp.async, two steps ahead: the loads
                                         // come from HBM and a phase is shorter than their latency)
  unsigned long long full[2][2];         // hfrag [sub-block][parity] (+ the gx slab of that step) complete: 8 local warps + 4 KB
                                         // from the peer
};

// NSB = 2: the schedule above (32 rows per cluster).  NSB = 1: 16 rows per cluster, every warp in the same phase: twice as many
// clusters for small batches, where SMs are idle and only the length of the step chain matters (B = 32: 2 clusters).
template <int NSB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    lstm_tc_fwd_kernel(float* __restrict__ gx, const int* __restrict__ tokens, const float* __restrict__ table,
                       const float* __restrict__ w_h2h, const float* __restrict__ b_h2h,
                       const float* __restrict__ h0, const float* __restrict__ c0, int ld0, float* __restrict__ hs,
                       float* __restrict__ hprev, float* __restrict__ cs, int B, int T) {
  pdl_entry();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int ul0 = warp * 8 + 2 * t4;                        // this thread's two local units (accumulator columns 2t, 2t+1)
  const int u0 = crank * UH + ul0;
  const int un = crank * UH + warp * 8 + g;                 // the unit whose W rows this thread holds as B fragments (n = lane / 4)
  const bool late = NSB == 2 && warp >= 4;                  // second group: one phase behind
  constexpr int RC = RB * NSB;                              // rows per cluster

  // A cluster is persistent over blocks of R batch rows (grid = min(#row blocks, resident clusters)): W_h2h goes into
  // registers once per CTA.  With T = 1 (one decode step over a large batch: style transfer / beam search) the launch
  // used to be 3-4 waves of CTAs that each re-loaded their 128 KB of W for a single step.
  int b0 = 0;
  // gx slab of (sub-block sb, step t) -> gxs[sb][t % 3]: 16 rows x 4 gates x 64 units = 1024 16-byte chunks, 4 per thread.
  // Always commits a group (possibly empty) so that the wait_group arithmetic below is the same in every step.
  auto prefetch_gx = [&](int sb, int t) {
    float* dst = sm.gxs[sb][t % 3];
    if (t < T) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ch = tid + i * kThreads, row = ch >> 6, rem = ch & 63, gate = rem >> 4, c4 = (rem & 15) * 4;
        const int b = min(b0 + sb * RB + row, B - 1);
        // table mode: the input pre-activations of row (b, t) are row tokens[b, t] of the [V, 4H] table emb W_i2h^T + b_i2h
        // (2.4 KB rows of an L2-resident 600 KB table): the [B*T, 4H] input never exists, gx only receives the activations
        const float* src = tokens ? table + (size_t)__ldg(tokens + (size_t)b * T + t) * 4 * H
                                  : gx + ((size_t)b * T + t) * 4 * H;
        cp_async16(dst + row * kRowPitch + gate * UH + c4, src + gate * H + crank * UH + c4);
      }
    }
    cp_async_commit();
  };
  // W_h2h slice as B fragments: breg[gate][k-step] = W[gate*H + un][8s + 2t], W[gate*H + un][8s + 2t + 1]
  float breg[4][16][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float* wr = w_h2h + (size_t)(j * H + un) * H + 2 * t4;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      const float2 w2 = __ldg(reinterpret_cast<const float2*>(wr + 8 * s));
      breg[j][s][0] = tf32r(w2.x);
      breg[j][s][1] = tf32r(w2.y);
    }
  }
  float bias[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bias[j][0] = __ldg(b_h2h + j * H + u0);
    bias[j][1] = __ldg(b_h2h + j * H + u0 + 1);
  }
  const unsigned hfrag_peer = peer_addr(smem_addr(&sm.hfrag[0][0][0]), crank ^ 1);   // shared::cluster addresses in the peer CTA
  const unsigned full_peer = peer_addr(smem_addr(&sm.full[0][0]), crank ^ 1);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&sm.full[0][0] + i, kThreads / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned phase = 0;                                       // bit 2 sb + i: parity of the next completion of full[sb][i]
  const int n_blocks = (B + RC - 1) / RC, n_clusters = gridDim.x >> 1;
  for (int blk = blockIdx.x >> 1; blk < n_blocks; blk += n_clusters) {
    b0 = blk * RC;
    cluster.sync();                                         // both CTAs are done with the previous row block's buffers
    // always four groups in flight before step 0 (NSB = 1: two of them empty), so that the wait_group counts below hold
    prefetch_gx(0, 0);
    if (NSB == 2) prefetch_gx(1, 0); else cp_async_commit();
    prefetch_gx(0, 1);
    if (NSB == 2) prefetch_gx(1, 1); else cp_async_commit();
    for (int i = tid; i < RC * H; i += kThreads) {
      const int r = i / H, k = i % H;
      sm.hfrag[r >> 4][0][afrag_index(r & 15, k, 16)] = (b0 + r < B) ? tf32r(__ldg(h0 + (size_t)(b0 + r) * ld0 + k)) : 0.f;
    }
    // cells of this thread: sub-block sb, rows 8 hi + g, units u0, u0 + 1
    float c[NSB][2][2], hp[NSB][2][2];
#pragma unroll
    for (int sb = 0; sb < NSB; ++sb)
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        const int b = min(b0 + sb * RB + 8 * hi + g, B - 1);
        c[sb][hi][0] = __ldg(c0 + (size_t)b * ld0 + u0);
        c[sb][hi][1] = __ldg(c0 + (size_t)b * ld0 + u0 + 1);
        hp[sb][hi][0] = __ldg(h0 + (size_t)b * ld0 + u0);
        hp[sb][hi][1] = __ldg(h0 + (size_t)b * ld0 + u0 + 1);
      }
    asm volatile("cp.async.wait_group 2;" ::: "memory");    // the slabs of step 0
    cluster.sync();
    // one-time skew: the second group starts when the first has issued its first MMA phase
    if (late) asm volatile("bar.sync 1, 256;" ::: "memory");

    for (int t = 0; t < T; ++t) {
#pragma unroll
      for (int sb = 0; sb < NSB; ++sb) {
        unsigned long long* bar_cur = &sm.full[sb][t & 1];
        if (t > 0) {                                        // h_{t-1} (both halves) and the gx slab of step t are in place
          mbar_wait(bar_cur, (phase >> (2 * sb + (t & 1))) & 1u);
          phase ^= 1u << (2 * sb + (t & 1));
        }
        prefetch_gx(sb, t + 2);                             // its buffer was last read in step t-1 (every warp arrived since)
        if (NSB == 1) cp_async_commit();                    // the group the second sub-block would have committed
        const float4* hcur = reinterpret_cast<const float4*>(sm.hfrag[sb][t & 1]);
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[j][0] = acc[j][2] = bias[j][0];
          acc[j][1] = acc[j][3] = bias[j][1];
        }
#pragma unroll
        for (int s = 0; s < 16; ++s) {
          const float4 a = hcur[s * 32 + lane];
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_tf32(acc[j], a, breg[j][s][0], breg[j][s][1]);
        }
        if (NSB == 2 && t == 0 && sb == 0 && !late) asm volatile("bar.arrive 1, 256;" ::: "memory");
        const float* gxc = sm.gxs[sb][t % 3];
        float2 iv[2], fv[2], gv[2], ov[2], cv[2], hv[2];
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
          const float* gp = gxc + (8 * hi + g) * kRowPitch + ul0;
          const float2 xi = *reinterpret_cast<const float2*>(gp), xf = *reinterpret_cast<const float2*>(gp + UH);
          const float2 xg = *reinterpret_cast<const float2*>(gp + 2 * UH), xo = *reinterpret_cast<const float2*>(gp + 3 * UH);
          iv[hi] = make_float2(sigmoidf_(acc[0][2 * hi] + xi.x), sigmoidf_(acc[0][2 * hi + 1] + xi.y));
          fv[hi] = make_float2(sigmoidf_(acc[1][2 * hi] + xf.x), sigmoidf_(acc[1][2 * hi + 1] + xf.y));
          gv[hi] = make_float2(tanhf(acc[2][2 * hi] + xg.x), tanhf(acc[2][2 * hi + 1] + xg.y));
          ov[hi] = make_float2(sigmoidf_(acc[3][2 * hi] + xo.x), sigmoidf_(acc[3][2 * hi + 1] + xo.y));
          c[sb][hi][0] = fv[hi].x * c[sb][hi][0] + iv[hi].x * gv[hi].x;
          c[sb][hi][1] = fv[hi].y * c[sb][hi][1] + iv[hi].y * gv[hi].y;
          cv[hi] = make_float2(c[sb][hi][0], c[sb][hi][1]);
          hv[hi] = make_float2(ov[hi].x * tanhf(c[sb][hi][0]), ov[hi].y * tanhf(c[sb][hi][1]));
        }
        // new h of (rows g, g+8) x (units u0, u0+1) is one float4 of the fragment buffer (see afrag_index).  The peer's copy
        // completes 16 bytes on the peer's full[sb][(t+1) & 1]; the buffer it lands in was last read by the peer in step
        // t-1, and the peer has sent its step-(t-1) halves (which this CTA waited for above) after those reads.
        if (t + 1 < T) {
          const float4 hf = make_float4(tf32r(hv[0].x), tf32r(hv[1].x), tf32r(hv[0].y), tf32r(hv[1].y));
          const int fi = (((u0 >> 3)) * 32 + g * 4 + t4) * 4;
          const int slot = 2 * sb + ((t + 1) & 1);
          st_async_v4(hfrag_peer + (slot * RB * H + fi) * 4, hf, full_peer + slot * 8);
          *reinterpret_cast<float4*>(&sm.hfrag[0][0][0] + slot * RB * H + fi) = hf;
          // this thread's chunks of the slab (sb, t+1) have landed: issued one step ago, two younger groups may be in flight
          asm volatile("cp.async.wait_group 2;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (warp == 0) mbar_arrive_expect_tx(&sm.full[0][0] + slot, RB * UH * 4);   // the peer's half: 4 KB
            else mbar_arrive(&sm.full[0][0] + slot);
          }
        }
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
          const int brow = b0 + sb * RB + 8 * hi + g;
          if (brow < B) {
            const size_t o = (size_t)brow * T + t;
            float* gp = gx + o * 4 * H + u0;
            *reinterpret_cast<float2*>(gp) = iv[hi];
            *reinterpret_cast<float2*>(gp + H) = fv[hi];
            *reinterpret_cast<float2*>(gp + 2 * H) = gv[hi];
            *reinterpret_cast<float2*>(gp + 3 * H) = ov[hi];
            *reinterpret_cast<float2*>(hs + o * H + u0) = hv[hi];
            *reinterpret_cast<float2*>(hprev + o * H + u0) = make_float2(hp[sb][hi][0], hp[sb][hi][1]);
            *reinterpret_cast<float2*>(cs + o * H + u0) = cv[hi];
          }
          hp[sb][hi][0] = hv[hi].x;
          hp[sb][hi][1] = hv[hi].y;
        }
      }
    }
  }                                                         // row blocks
  cluster.sync();                                           // no CTA leaves while its peer could still address its shared memory
}

// ------------------------------------------------------------------------------------ backward
// gates [B,T,4H] in: saved activations ; out: d(pre-activation gates).  dhs [B,T,H]; dh0 / dc0 rows of a [B,ld0] buffer.
// Per step: thread-local cell backward for the CTA's 64 units -> dg (A-fragment order) -> partial
// dh_rec[R x 128] = dg[R x 256 local gate columns] Wn[256 x 128] on mma.sync (warp w owns hidden columns [16w, 16w+16))
// -> the partial of this CTA's own units stays in dhrec, the partial of the peer's units is stored into the peer's
// dhin through DSMEM; both are summed when the next step reads them.
struct BwdSmem {
  float dgfrag[R * 4 * UH];              // d(pre-activation) [32 x 256] in A-fragment order (32 KB)
  float dhrec[2][R * kColPitch];         // recurrent dh of the local units: own partial   (step t writes [t & 1], step t-1
  float dhin[2][R * kColPitch];          //                                   peer's partial  reads it: one barrier per step)
  float gts[2][R * kRowPitch];           // saved gate activations of steps t, t-1 (cp.async, one step ahead)
  float cst[2][R * kColPitch];           // c_{t-1}
  float dht[2][R * kColPitch];           // dhs
  unsigned long long full[2];            // partials of step t in dhrec / dhin[t & 1] and the slabs of step t-1: 8 local warps + 8 KB
};

// MT = 2: 32 rows per cluster (two m16 tiles per warp).  MT = 1: 16 rows per cluster for small batches (see the forward kernel).
template <int MT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    lstm_tc_bwd_kernel(float* __restrict__ gates, const float* __restrict__ w_h2h, const float* __restrict__ cs,
                       const float* __restrict__ c0, int ld0, const float* __restrict__ dhs, float* __restrict__ dh0,
                       float* __restrict__ dc0, float* __restrict__ db_i2h, float* __restrict__ db_h2h, int B, int T) {
  pdl_entry();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  constexpr int RR = 16 * MT;                               // rows per cluster
  const int b0 = (blockIdx.x >> 1) * RR;
  const int ul0 = warp * 8 + 2 * t4;                        // cell mapping: local units ul0, ul0 + 1 (as in the forward kernel)
  const int u0 = crank * UH + ul0;

  // slabs of step t -> buffer t & 1: gates 2048 chunks (8 per thread), c_{t-1} and dhs 512 chunks each (2 per thread)
  auto prefetch = [&](int t) {
    const int buf = t & 1;
#pragma unroll
    for (int i = 0; i < 4 * MT; ++i) {
      const int ch = tid + i * kThreads, row = ch >> 6, rem = ch & 63, gate = rem >> 4, c4 = (rem & 15) * 4;
      const int b = min(b0 + row, B - 1);
      cp_async16(sm.gts[buf] + row * kRowPitch + gate * UH + c4, gates + ((size_t)b * T + t) * 4 * H + gate * H + crank * UH + c4);
    }
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const int ch = tid + i * kThreads, row = ch >> 4, c4 = (ch & 15) * 4;
      const int b = min(b0 + row, B - 1);
      cp_async16(sm.dht[buf] + row * kColPitch + c4, dhs + ((size_t)b * T + t) * H + crank * UH + c4);
      // c_{t-1}; for t == 0 it is c0, a [B, ld0] buffer whose rows are only 8-byte aligned: two 8-byte copies
      if (t > 0) {
        cp_async16(sm.cst[buf] + row * kColPitch + c4, cs + ((size_t)b * T + t - 1) * H + crank * UH + c4);
      } else {
        const float* src = c0 + (size_t)b * ld0 + crank * UH + c4;
        float* dst = sm.cst[buf] + row * kColPitch + c4;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst + 2)), "l"(src + 2) : "memory");
      }
    }
    cp_async_commit();
  };
  prefetch(T - 1);

  // Wn[j][k] = W_h2h[gate(j) * H + crank * 64 + ul(j)][k], j = gate * 64 + ul (local gate column).
  // breg[n-tile][k-step] = Wn[8s + 2t][16 warp + 8 nn + g], Wn[8s + 2t + 1][...]
  float breg[2][32][2];
#pragma unroll
  for (int s = 0; s < 32; ++s) {
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int j = 8 * s + 2 * t4 + h2;
      const float* wr = w_h2h + (size_t)((j >> 6) * H + crank * UH + (j & 63)) * H + 16 * warp + g;
      breg[0][s][h2] = tf32r(__ldg(wr));
      breg[1][s][h2] = tf32r(__ldg(wr + 8));
    }
  }
  for (int i = tid; i < 2 * R * kColPitch; i += kThreads) { (&sm.dhrec[0][0])[i] = 0.f; (&sm.dhin[0][0])[i] = 0.f; }
  float dc[2 * MT][2], bsum[4][2], cnext[2 * MT][2];                  // cnext: c_t of the step about to be processed
  int brow[2 * MT];
#pragma unroll
  for (int q = 0; q < 2 * MT; ++q) {
    brow[q] = b0 + 16 * (q >> 1) + 8 * (q & 1) + g;
    const int b = min(brow[q], B - 1);
    const float2 cv = *reinterpret_cast<const float2*>(cs + ((size_t)b * T + (T - 1)) * H + u0);
    cnext[q][0] = cv.x; cnext[q][1] = cv.y;
    dc[q][0] = dc[q][1] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) bsum[j][0] = bsum[j][1] = 0.f;
  // hidden columns [16 warp, 16 warp + 16): warps 0-3 produce CTA 0's units, warps 4-7 CTA 1's
  const bool mine = (warp >> 2) == crank;
  const unsigned dhin_peer = peer_addr(smem_addr(&sm.dhin[0][0]), crank ^ 1);   // shared::cluster addresses in the peer CTA
  const unsigned full_peer = peer_addr(smem_addr(&sm.full[0]), crank ^ 1);
  if (tid == 0) {
    mbar_init(&sm.full[0], kThreads / 32);
    mbar_init(&sm.full[1], kThreads / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned phase = 0;                                         // bit i: parity of the next completion of full[i]
  cp_async_wait_all();
  cluster.sync();

  for (int t = T - 1; t >= 0; --t) {
    const int buf = t & 1;
    if (t < T - 1) {                                        // step t+1 is complete in both CTAs, the slabs of step t have landed
      mbar_wait(&sm.full[buf ^ 1], (phase >> (buf ^ 1)) & 1u);
      phase ^= 1u << (buf ^ 1);
    }
    if (t > 0) prefetch(t - 1);                             // its buffers were last read in step t+1 (every warp arrived since)
    // ---- cell backward (thread-local), operands from the staged slabs
    float2 di[2 * MT], df[2 * MT], dg2[2 * MT], dou[2 * MT];
#pragma unroll
    for (int q = 0; q < 2 * MT; ++q) {
      const int row = 16 * (q >> 1) + 8 * (q & 1) + g;
      const float* gp = sm.gts[buf] + row * kRowPitch + ul0;
      const float2 iv = *reinterpret_cast<const float2*>(gp), fv = *reinterpret_cast<const float2*>(gp + UH);
      const float2 gv = *reinterpret_cast<const float2*>(gp + 2 * UH), ov = *reinterpret_cast<const float2*>(gp + 3 * UH);
      const float2 cp = *reinterpret_cast<const float2*>(sm.cst[buf] + row * kColPitch + ul0);
      const float2 dhv = *reinterpret_cast<const float2*>(sm.dht[buf] + row * kColPitch + ul0);
      const float2 r1 = *reinterpret_cast<const float2*>(sm.dhrec[buf ^ 1] + row * kColPitch + ul0);   // written by step t+1
      const float2 r2 = *reinterpret_cast<const float2*>(sm.dhin[buf ^ 1] + row * kColPitch + ul0);
      const bool live = brow[q] < B;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float ig = e ? iv.y : iv.x, fg = e ? fv.y : fv.x, gg = e ? gv.y : gv.x, og = e ? ov.y : ov.x;
        const float cprev = e ? cp.y : cp.x;
        const float dh = (e ? dhv.y : dhv.x) + (e ? r1.y : r1.x) + (e ? r2.y : r2.x);
        const float tc = tanhf(cnext[q][e]);
        const float dout = live ? dh * tc * og * (1.f - og) : 0.f;
        const float dct = dc[q][e] + dh * og * (1.f - tc * tc);
        const float dii = live ? dct * gg * ig * (1.f - ig) : 0.f;
        const float dgg = live ? dct * ig * (1.f - gg * gg) : 0.f;
        const float dff = live ? dct * cprev * fg * (1.f - fg) : 0.f;
        dc[q][e] = dct * fg;
        cnext[q][e] = cprev;
        bsum[0][e] += dii; bsum[1][e] += dff; bsum[2][e] += dgg; bsum[3][e] += dout;
        if (e == 0) { di[q].x = dii; df[q].x = dff; dg2[q].x = dgg; dou[q].x = dout; }
        else        { di[q].y = dii; df[q].y = dff; dg2[q].y = dgg; dou[q].y = dout; }
      }
    }
    // dg of (rows g, g+8) x (units ul0, ul0+1) of m-tile m and one gate is one float4 of the fragment buffer
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const int fb = (m * 32 + (ul0 >> 3)) * 32 + g * 4 + t4;
      float4* dstf = reinterpret_cast<float4*>(sm.dgfrag);
      dstf[fb + 0 * 8 * 32] = make_float4(tf32r(di[2 * m].x), tf32r(di[2 * m + 1].x), tf32r(di[2 * m].y), tf32r(di[2 * m + 1].y));
      dstf[fb + 1 * 8 * 32] = make_float4(tf32r(df[2 * m].x), tf32r(df[2 * m + 1].x), tf32r(df[2 * m].y), tf32r(df[2 * m + 1].y));
      dstf[fb + 2 * 8 * 32] = make_float4(tf32r(dg2[2 * m].x), tf32r(dg2[2 * m + 1].x), tf32r(dg2[2 * m].y), tf32r(dg2[2 * m + 1].y));
      dstf[fb + 3 * 8 * 32] = make_float4(tf32r(dou[2 * m].x), tf32r(dou[2 * m + 1].x), tf32r(dou[2 * m].y), tf32r(dou[2 * m + 1].y));
    }
    __syncthreads();                                        // dg complete
#pragma unroll
    for (int q = 0; q < 2 * MT; ++q) {                           // global stores drain during the MMAs
      if (brow[q] < B) {
        float* gp = gates + ((size_t)brow[q] * T + t) * 4 * H + u0;
        *reinterpret_cast<float2*>(gp) = di[q];
        *reinterpret_cast<float2*>(gp + H) = df[q];
        *reinterpret_cast<float2*>(gp + 2 * H) = dg2[q];
        *reinterpret_cast<float2*>(gp + 3 * H) = dou[q];
      }
    }
    // ---- partial dh_rec = dg Wn over all 256 local gate columns for this warp's 16 hidden columns
    float acc[MT][2][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) acc[m][nn][0] = acc[m][nn][1] = acc[m][nn][2] = acc[m][nn][3] = 0.f;
    const float4* af = reinterpret_cast<const float4*>(sm.dgfrag);
#pragma unroll
    for (int s = 0; s < 32; ++s) {
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const float4 a = af[(m * 32 + s) * 32 + lane];
        mma_tf32(acc[m][0], a, breg[0][s][0], breg[0][s][1]);
        mma_tf32(acc[m][1], a, breg[1][s][0], breg[1][s][1]);
      }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) {
        const int col = ((16 * warp + 8 * nn) & 63) + 2 * t4;   // local unit in the owning CTA
        const int o0 = buf * R * kColPitch + (16 * m + g) * kColPitch + col, o1 = o0 + 8 * kColPitch;
        if (mine) {
          *reinterpret_cast<float2*>(&sm.dhrec[0][0] + o0) = make_float2(acc[m][nn][0], acc[m][nn][1]);
          *reinterpret_cast<float2*>(&sm.dhrec[0][0] + o1) = make_float2(acc[m][nn][2], acc[m][nn][3]);
        } else {                                             // the peer's units: 8 bytes each on the peer's full[buf]
          st_async_v2(dhin_peer + o0 * 4, make_float2(acc[m][nn][0], acc[m][nn][1]), full_peer + buf * 8);
          st_async_v2(dhin_peer + o1 * 4, make_float2(acc[m][nn][2], acc[m][nn][3]), full_peer + buf * 8);
        }
      }
    cp_async_wait_all();                                    // this thread's chunks of the slabs of step t-1 have landed
    __syncwarp();
    if (lane == 0) {
      if (warp == 0) mbar_arrive_expect_tx(&sm.full[buf], RR * UH * 4);   // the peer's partial of this CTA's units: 8 KB per step
      else mbar_arrive(&sm.full[buf]);
    }
  }
  mbar_wait(&sm.full[0], phase & 1u);                       // step 0 complete in both CTAs
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      // the 8 lanes that share t4 hold the same units for different rows: fold them before the atomics
      float v = bsum[j][e];
      v += __shfl_xor_sync(MSX_FULL, v, 4);
      v += __shfl_xor_sync(MSX_FULL, v, 8);
      v += __shfl_xor_sync(MSX_FULL, v, 16);
      if (g == 0) {
        if (db_i2h) atomicAdd(db_i2h + j * H + u0 + e, v);
        if (db_h2h) atomicAdd(db_h2h + j * H + u0 + e, v);
      }
    }
#pragma unroll
  for (int q = 0; q < 2 * MT; ++q) {
    const int row = 16 * (q >> 1) + 8 * (q & 1) + g;
    if (brow[q] < B) {
      const float* r1 = sm.dhrec[0] + row * kColPitch + ul0;   // written by step t = 0
      const float* r2 = sm.dhin[0] + row * kColPitch + ul0;
      *reinterpret_cast<float2*>(dh0 + (size_t)brow[q] * ld0 + u0) = make_float2(r1[0] + r2[0], r1[1] + r2[1]);
      *reinterpret_cast<float2*>(dc0 + (size_t)brow[q] * ld0 + u0) = make_float2(dc[q][0], dc[q][1]);
    }
  }
  cluster.sync();                                           // no CTA leaves while its peer could still address its shared memory
}

}  // namespace

// 1 when the tensor LSTM kernels take this problem: H == 128 and 8-byte aligned rows (float2 accesses)
extern "C" int msx_lstm_tc_supported(int H_, int ld0, const float* h0, const float* c0) {
  return (H_ == H && (ld0 & 1) == 0 && h0 && c0 && ((uintptr_t)h0 & 7) == 0 && ((uintptr_t)c0 & 7) == 0) ? 1 : 0;
}

extern "C" int msx_lstm_tc_fwd_tab(float* gates_out, const int32_t* tokens, const float* table, const float* w_h2h,
                                   const float* b_h2h, const float* h0, const float* c0, int ld0, float* hs, float* hprev,
                                   float* cs, int B, int T, int H_, void* stream);
extern "C" int msx_lstm_tc_fwd(float* gx_inout, const float* w_h2h, const float* b_h2h, const float* h0, const float* c0,
                               int ld0, float* hs, float* hprev, float* cs, int B, int T, int H_, void* stream) {
  return msx_lstm_tc_fwd_tab(gx_inout, nullptr, nullptr, w_h2h, b_h2h, h0, c0, ld0, hs, hprev, cs, B, T, H_, stream);
}

// Table mode (tokens != NULL): the layer's input is an embedding lookup, so x_t W_i2h^T + b_i2h is row tokens[b, t] of
// table [V, 4H] = emb W_i2h^T + b_i2h (one tiny GEMM per step).  The kernel fetches those rows itself; gates_out receives the
// gate activations as before.  Token ids must lie in [0, V): the caller's table has V rows.
extern "C" int msx_lstm_tc_fwd_tab(float* gx_inout, const int32_t* tokens, const float* table, const float* w_h2h,
                                   const float* b_h2h, const float* h0, const float* c0, int ld0, float* hs, float* hprev,
                                   float* cs, int B, int T, int H_, void* stream) {
  MSX_REQUIRE(gx_inout && w_h2h && b_h2h && h0 && c0 && hs && hprev && cs, "msx_lstm_tc_fwd: null pointer");
  MSX_REQUIRE((tokens == nullptr) == (table == nullptr) && ((uintptr_t)table & 15) == 0,
              "msx_lstm_tc_fwd_tab: tokens and a 16-byte aligned table go together");
  MSX_REQUIRE(msx_lstm_tc_supported(H_, ld0, h0, c0), "msx_lstm_tc_fwd: needs H == 128, even ld0, 8-byte aligned h0 / c0");
  if (B == 0 || T == 0) return MSX_OK;
  // persistent clusters: one 2-CTA cluster per SM pair at most, each walks its row blocks with W_h2h kept in registers;
  // 16-row clusters while they all fit at once (small batches: more SMs on the same chain of T steps)
  const int resident = msx_num_sms() / 2;
  if ((B + RB - 1) / RB <= resident) {
    const int clusters = (B + RB - 1) / RB;
    MSX_CUDA(cudaFuncSetAttribute(lstm_tc_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem)));
    MSX_CUDA(msx_launch(lstm_tc_fwd_kernel<1>, dim3(clusters * 2), dim3(kThreads), sizeof(FwdSmem), (cudaStream_t)stream, gx_inout, tokens, table, w_h2h, b_h2h,
                                                                                            h0, c0, ld0, hs, hprev, cs, B, T));
  } else {
    const int blocks = (B + R - 1) / R;
    const int clusters = blocks < resident ? blocks : resident;
    MSX_CUDA(cudaFuncSetAttribute(lstm_tc_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem)));
    MSX_CUDA(msx_launch(lstm_tc_fwd_kernel<2>, dim3(clusters * 2), dim3(kThreads), sizeof(FwdSmem), (cudaStream_t)stream, gx_inout, tokens, table, w_h2h, b_h2h,
                                                                                            h0, c0, ld0, hs, hprev, cs, B, T));
  }
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_lstm_tc_bwd(float* gates_inout, const float* w_h2h, const float* cs, const float* c0, int ld0,
                               const float* dhs, float* dh0, float* dc0, float* db_i2h, float* db_h2h, int B, int T,
                               int H_, void* stream) {
  MSX_REQUIRE(gates_inout && w_h2h && cs && c0 && dhs && dh0 && dc0, "msx_lstm_tc_bwd: null pointer");
  MSX_REQUIRE(msx_lstm_tc_supported(H_, ld0, dh0, c0) && ((uintptr_t)dc0 & 7) == 0,
              "msx_lstm_tc_bwd: needs H == 128, even ld0, 8-byte aligned c0 / dh0 / dc0");
  if (B == 0 || T == 0) return MSX_OK;
  if ((B + 15) / 16 <= msx_num_sms() / 2) {                 // small batches: 16-row clusters, all resident at once
    const int clusters = (B + 15) / 16;
    MSX_CUDA(cudaFuncSetAttribute(lstm_tc_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
    MSX_CUDA(msx_launch(lstm_tc_bwd_kernel<1>, dim3(clusters * 2), dim3(kThreads), sizeof(BwdSmem), (cudaStream_t)stream, gates_inout, w_h2h, cs, c0, ld0, dhs,
                                                                                            dh0, dc0, db_i2h, db_h2h, B, T));
  } else {
    const int clusters = (B + R - 1) / R;
    MSX_CUDA(cudaFuncSetAttribute(lstm_tc_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
    MSX_CUDA(msx_launch(lstm_tc_bwd_kernel<2>, dim3(clusters * 2), dim3(kThreads), sizeof(BwdSmem), (cudaStream_t)stream, gates_inout, w_h2h, cs, c0, ld0, dhs,
                                                                                            dh0, dc0, db_i2h, db_h2h, B, T));
  }
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
