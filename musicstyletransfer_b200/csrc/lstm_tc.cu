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
//     thread-local; the new h goes to both CTAs' fragment buffers through DSMEM, one cluster barrier per step.
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
// split cluster barrier: the release of `arrive` only has to cover the shared-memory / DSMEM writes issued before it,
// so the global stores of a step are issued between arrive and wait and drain during the next step's MMAs
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

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
struct FwdSmem {
  float hfrag[2][R * H];                 // h_{t-1} in A-fragment order, double-buffered (2 x 16 KB)
  float gxs[2][R * kRowPitch];           // this CTA's slice of gx for steps t, t+1 (cp.async, one step ahead)
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    lstm_tc_fwd_kernel(float* __restrict__ gx, const float* __restrict__ w_h2h, const float* __restrict__ b_h2h,
                       const float* __restrict__ h0, const float* __restrict__ c0, int ld0, float* __restrict__ hs,
                       float* __restrict__ hprev, float* __restrict__ cs, int B, int T) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int ul0 = warp * 8 + 2 * t4;                        // this thread's two local units (accumulator columns 2t, 2t+1)
  const int u0 = crank * UH + ul0;
  const int un = crank * UH + warp * 8 + g;                 // the unit whose W rows this thread holds as B fragments (n = lane / 4)

  // A cluster is persistent over blocks of R batch rows (grid = min(#row blocks, resident clusters)): W_h2h goes into
  // registers once per CTA.  With T = 1 (one decode step over a large batch: style transfer / beam search) the launch
  // used to be 3-4 waves of CTAs that each re-loaded their 128 KB of W for a single step.
  int b0 = 0;
  // gx slab of step t -> gxs[t & 1]: 32 rows x 4 gates x 64 units = 2048 16-byte chunks, 8 per thread
  auto prefetch_gx = [&](int t) {
    float* dst = sm.gxs[t & 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = tid + i * kThreads, row = ch >> 6, rem = ch & 63, gate = rem >> 4, c4 = (rem & 15) * 4;
      const int b = min(b0 + row, B - 1);
      cp_async16(dst + row * kRowPitch + gate * UH + c4, gx + ((size_t)b * T + t) * 4 * H + gate * H + crank * UH + c4);
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
  float* hfrag_peer = cluster.map_shared_rank(&sm.hfrag[0][0], crank ^ 1);
  const int n_blocks = (B + R - 1) / R, n_clusters = gridDim.x >> 1;
  for (int blk = blockIdx.x >> 1; blk < n_blocks; blk += n_clusters) {
  b0 = blk * R;
  cluster.sync();                                           // both CTAs are done with the previous row block's buffers
  prefetch_gx(0);
  for (int i = tid; i < R * H; i += kThreads) {
    const int r = i / H, k = i % H;
    sm.hfrag[0][afrag_index(r, k, 16)] = (b0 + r < B) ? tf32r(__ldg(h0 + (size_t)(b0 + r) * ld0 + k)) : 0.f;
  }
  // cells of this thread: rows 16m + 8hi + g (index q = 2m + hi), units u0, u0 + 1
  float c[4][2], hp[4][2];
  int brow[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    brow[q] = b0 + 16 * (q >> 1) + 8 * (q & 1) + g;
    const int b = min(brow[q], B - 1);
    c[q][0] = __ldg(c0 + (size_t)b * ld0 + u0);
    c[q][1] = __ldg(c0 + (size_t)b * ld0 + u0 + 1);
    hp[q][0] = __ldg(h0 + (size_t)b * ld0 + u0);
    hp[q][1] = __ldg(h0 + (size_t)b * ld0 + u0 + 1);
  }
  cp_async_wait_all();
  cluster.sync();

  for (int t = 0; t < T; ++t) {
    if (t + 1 < T) prefetch_gx(t + 1);                      // its buffer was last read in step t-1
    const float4* hcur = reinterpret_cast<const float4*>(sm.hfrag[t & 1]);
    float acc[2][4][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[m][j][0] = acc[m][j][2] = bias[j][0];
        acc[m][j][1] = acc[m][j][3] = bias[j][1];
      }
#pragma unroll
    for (int s = 0; s < 16; ++s) {
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const float4 a = hcur[(m * 16 + s) * 32 + lane];
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_tf32(acc[m][j], a, breg[j][s][0], breg[j][s][1]);
      }
    }
    const float* gxc = sm.gxs[t & 1];
    float* hnext = sm.hfrag[(t + 1) & 1];
    float* hnext_peer = hfrag_peer + ((t + 1) & 1) * R * H;
    float2 iv[4], fv[4], gv[4], ov[4], cv[4], hv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int m = q >> 1, hi = q & 1, row = 16 * m + 8 * hi + g;
      const float* gp = gxc + row * kRowPitch + ul0;
      const float2 xi = *reinterpret_cast<const float2*>(gp), xf = *reinterpret_cast<const float2*>(gp + UH);
      const float2 xg = *reinterpret_cast<const float2*>(gp + 2 * UH), xo = *reinterpret_cast<const float2*>(gp + 3 * UH);
      iv[q] = make_float2(sigmoidf_(acc[m][0][2 * hi] + xi.x), sigmoidf_(acc[m][0][2 * hi + 1] + xi.y));
      fv[q] = make_float2(sigmoidf_(acc[m][1][2 * hi] + xf.x), sigmoidf_(acc[m][1][2 * hi + 1] + xf.y));
      gv[q] = make_float2(tanhf(acc[m][2][2 * hi] + xg.x), tanhf(acc[m][2][2 * hi + 1] + xg.y));
      ov[q] = make_float2(sigmoidf_(acc[m][3][2 * hi] + xo.x), sigmoidf_(acc[m][3][2 * hi + 1] + xo.y));
      c[q][0] = fv[q].x * c[q][0] + iv[q].x * gv[q].x;
      c[q][1] = fv[q].y * c[q][1] + iv[q].y * gv[q].y;
      cv[q] = make_float2(c[q][0], c[q][1]);
      hv[q] = make_float2(ov[q].x * tanhf(c[q][0]), ov[q].y * tanhf(c[q][1]));
    }
    // new h of (rows g, g+8) x (units u0, u0+1) of m-tile m is one float4 of the fragment buffer (see afrag_index)
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const float4 hf = make_float4(tf32r(hv[2 * m].x), tf32r(hv[2 * m + 1].x), tf32r(hv[2 * m].y), tf32r(hv[2 * m + 1].y));
      const int fi = ((m * 16 + (u0 >> 3)) * 32 + g * 4 + t4) * 4;
      *reinterpret_cast<float4*>(hnext + fi) = hf;
      *reinterpret_cast<float4*>(hnext_peer + fi) = hf;
    }
    cp_async_wait_all();                                    // gx of step t+1 has landed (visible to all after the barrier)
    cluster_arrive();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (brow[q] < B) {
        const size_t o = (size_t)brow[q] * T + t;
        float* gp = gx + o * 4 * H + u0;
        *reinterpret_cast<float2*>(gp) = iv[q];
        *reinterpret_cast<float2*>(gp + H) = fv[q];
        *reinterpret_cast<float2*>(gp + 2 * H) = gv[q];
        *reinterpret_cast<float2*>(gp + 3 * H) = ov[q];
        *reinterpret_cast<float2*>(hs + o * H + u0) = hv[q];
        *reinterpret_cast<float2*>(hprev + o * H + u0) = make_float2(hp[q][0], hp[q][1]);
        *reinterpret_cast<float2*>(cs + o * H + u0) = cv[q];
      }
      hp[q][0] = hv[q].x;
      hp[q][1] = hv[q].y;
    }
    cluster_wait();
  }
  }                                                         // row blocks
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
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    lstm_tc_bwd_kernel(float* __restrict__ gates, const float* __restrict__ w_h2h, const float* __restrict__ cs,
                       const float* __restrict__ c0, int ld0, const float* __restrict__ dhs, float* __restrict__ dh0,
                       float* __restrict__ dc0, float* __restrict__ db_i2h, float* __restrict__ db_h2h, int B, int T) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int b0 = (blockIdx.x >> 1) * R;
  const int ul0 = warp * 8 + 2 * t4;                        // cell mapping: local units ul0, ul0 + 1 (as in the forward kernel)
  const int u0 = crank * UH + ul0;

  // slabs of step t -> buffer t & 1: gates 2048 chunks (8 per thread), c_{t-1} and dhs 512 chunks each (2 per thread)
  auto prefetch = [&](int t) {
    const int buf = t & 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = tid + i * kThreads, row = ch >> 6, rem = ch & 63, gate = rem >> 4, c4 = (rem & 15) * 4;
      const int b = min(b0 + row, B - 1);
      cp_async16(sm.gts[buf] + row * kRowPitch + gate * UH + c4, gates + ((size_t)b * T + t) * 4 * H + gate * H + crank * UH + c4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
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
  float dc[4][2], bsum[4][2], cnext[4][2];                  // cnext: c_t of the step about to be processed
  int brow[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
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
  float* dst_base = mine ? &sm.dhrec[0][0] : cluster.map_shared_rank(&sm.dhin[0][0], crank ^ 1);
  cp_async_wait_all();
  cluster.sync();

  for (int t = T - 1; t >= 0; --t) {
    const int buf = t & 1;
    if (t > 0) prefetch(t - 1);                             // its buffers were last read in step t+1
    // ---- cell backward (thread-local), operands from the staged slabs
    float2 di[4], df[4], dg2[4], dou[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
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
    for (int m = 0; m < 2; ++m) {
      const int fb = (m * 32 + (ul0 >> 3)) * 32 + g * 4 + t4;
      float4* dstf = reinterpret_cast<float4*>(sm.dgfrag);
      dstf[fb + 0 * 8 * 32] = make_float4(tf32r(di[2 * m].x), tf32r(di[2 * m + 1].x), tf32r(di[2 * m].y), tf32r(di[2 * m + 1].y));
      dstf[fb + 1 * 8 * 32] = make_float4(tf32r(df[2 * m].x), tf32r(df[2 * m + 1].x), tf32r(df[2 * m].y), tf32r(df[2 * m + 1].y));
      dstf[fb + 2 * 8 * 32] = make_float4(tf32r(dg2[2 * m].x), tf32r(dg2[2 * m + 1].x), tf32r(dg2[2 * m].y), tf32r(dg2[2 * m + 1].y));
      dstf[fb + 3 * 8 * 32] = make_float4(tf32r(dou[2 * m].x), tf32r(dou[2 * m + 1].x), tf32r(dou[2 * m].y), tf32r(dou[2 * m + 1].y));
    }
    __syncthreads();                                        // dg complete
#pragma unroll
    for (int q = 0; q < 4; ++q) {                           // global stores drain during the MMAs
      if (brow[q] < B) {
        float* gp = gates + ((size_t)brow[q] * T + t) * 4 * H + u0;
        *reinterpret_cast<float2*>(gp) = di[q];
        *reinterpret_cast<float2*>(gp + H) = df[q];
        *reinterpret_cast<float2*>(gp + 2 * H) = dg2[q];
        *reinterpret_cast<float2*>(gp + 3 * H) = dou[q];
      }
    }
    // ---- partial dh_rec = dg Wn over all 256 local gate columns for this warp's 16 hidden columns
    float acc[2][2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) acc[m][nn][0] = acc[m][nn][1] = acc[m][nn][2] = acc[m][nn][3] = 0.f;
    const float4* af = reinterpret_cast<const float4*>(sm.dgfrag);
#pragma unroll
    for (int s = 0; s < 32; ++s) {
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const float4 a = af[(m * 32 + s) * 32 + lane];
        mma_tf32(acc[m][0], a, breg[0][s][0], breg[0][s][1]);
        mma_tf32(acc[m][1], a, breg[1][s][0], breg[1][s][1]);
      }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) {
        const int col = ((16 * warp + 8 * nn) & 63) + 2 * t4;   // local unit in the owning CTA
        float* d = dst_base + buf * R * kColPitch;
        *reinterpret_cast<float2*>(d + (16 * m + g) * kColPitch + col) = make_float2(acc[m][nn][0], acc[m][nn][1]);
        *reinterpret_cast<float2*>(d + (16 * m + g + 8) * kColPitch + col) = make_float2(acc[m][nn][2], acc[m][nn][3]);
      }
    cp_async_wait_all();                                    // slabs of step t-1 have landed
    cluster.sync();                                         // dhrec / dhin of step t-1 complete in both CTAs
  }
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
  for (int q = 0; q < 4; ++q) {
    const int row = 16 * (q >> 1) + 8 * (q & 1) + g;
    if (brow[q] < B) {
      const float* r1 = sm.dhrec[0] + row * kColPitch + ul0;   // written by step t = 0
      const float* r2 = sm.dhin[0] + row * kColPitch + ul0;
      *reinterpret_cast<float2*>(dh0 + (size_t)brow[q] * ld0 + u0) = make_float2(r1[0] + r2[0], r1[1] + r2[1]);
      *reinterpret_cast<float2*>(dc0 + (size_t)brow[q] * ld0 + u0) = make_float2(dc[q][0], dc[q][1]);
    }
  }
}

}  // namespace

// 1 when the tensor LSTM kernels take this problem: H == 128 and 8-byte aligned rows (float2 accesses)
extern "C" int msx_lstm_tc_supported(int H_, int ld0, const float* h0, const float* c0) {
  return (H_ == H && (ld0 & 1) == 0 && h0 && c0 && ((uintptr_t)h0 & 7) == 0 && ((uintptr_t)c0 & 7) == 0) ? 1 : 0;
}

extern "C" int msx_lstm_tc_fwd(float* gx_inout, const float* w_h2h, const float* b_h2h, const float* h0, const float* c0,
                               int ld0, float* hs, float* hprev, float* cs, int B, int T, int H_, void* stream) {
  MSX_REQUIRE(gx_inout && w_h2h && b_h2h && h0 && c0 && hs && hprev && cs, "msx_lstm_tc_fwd: null pointer");
  MSX_REQUIRE(msx_lstm_tc_supported(H_, ld0, h0, c0), "msx_lstm_tc_fwd: needs H == 128, even ld0, 8-byte aligned h0 / c0");
  if (B == 0 || T == 0) return MSX_OK;
  // persistent clusters: one 2-CTA cluster per SM pair at most, each walks its row blocks with W_h2h kept in registers
  const int blocks = (B + R - 1) / R;
  const int resident = msx_num_sms() / 2;
  const int clusters = blocks < resident ? blocks : resident;
  MSX_CUDA(cudaFuncSetAttribute(lstm_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem)));
  lstm_tc_fwd_kernel<<<clusters * 2, kThreads, sizeof(FwdSmem), (cudaStream_t)stream>>>(gx_inout, w_h2h, b_h2h, h0, c0, ld0, hs,
                                                                                       hprev, cs, B, T);
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
  const int clusters = (B + R - 1) / R;
  MSX_CUDA(cudaFuncSetAttribute(lstm_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
  lstm_tc_bwd_kernel<<<clusters * 2, kThreads, sizeof(BwdSmem), (cudaStream_t)stream>>>(gates_inout, w_h2h, cs, c0, ld0, dhs, dh0,
                                                                                       dc0, db_i2h, db_h2h, B, T);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
