// K2g — persistent LSTM recurrence (forward and backward through time), fp32-exact.
//
// Replaces the fused gluon.rnn.LSTM call of LSTMDecoder.forward_train
// (/root/reference/music_style_transfer/VarAutoEncoder/model.py:148-153,179): T sequential steps of
//   g = x_t W_i2h^T + b_i2h + h W_h2h^T + b_h2h ; i,f,o = sigmoid, g~ = tanh ; c = f c + i g~ ; h = o tanh(c)
// The input projection x W_i2h^T + b_i2h for all T is one GEMM done before (msx_gemm_*), so this
// kernel only carries the h W_h2h^T recurrence.  W_h2h (4H x H fp32 = 256 KB at H=128) does not fit
// one SM's shared memory, so a CLUSTER OF TWO CTAs owns a tile of R batch rows: each CTA keeps the
// 4 gate rows of half of the hidden units resident in its shared memory for all T steps, computes
// the new h for its units and pushes it into BOTH CTAs' h buffers through distributed shared memory;
// one cluster barrier per time step.  Each thread owns all four gates of one unit for R/4 rows, so
// the cell update needs no cross-thread traffic.  The backward kernel walks t = T-1..0 with the same
// residency (natural W layout), exchanging the partial dh of the other CTA's units through DSMEM.
#include <cooperative_groups.h>

#include <stdlib.h>

#include "msx_common.cuh"

namespace cg = cooperative_groups;

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ------------------------------------------------------------------------------------ forward
// gx     [B,T,4H]  in: x W_i2h^T + b_i2h ; out: gate activations (i,f,g,o) saved for backward
// hs     [B,T,H]   h_t ;  hprev [B,T,H] h_{t-1} (h0 at t=0) ;  cs [B,T,H] c_t
template <int H, int RPT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(2 * H)
    lstm_fwd_kernel(float* __restrict__ gx, const float* __restrict__ w_h2h, const float* __restrict__ b_h2h,
                    const float* __restrict__ h0, const float* __restrict__ c0, int ld0, float* __restrict__ hs,
                    float* __restrict__ hprev, float* __restrict__ cs, int B, int T) {
  constexpr int UH = H / 2;        // hidden units owned by this CTA
  constexpr int R = 4 * RPT;       // batch rows per cluster
  extern __shared__ __align__(16) float smem_f[];
  float* Wt = smem_f;              // [H][UH*4]   Wt[k][4*ul+g] = W_h2h[g*H + u][k]
  float* hbuf = Wt + H * UH * 4;   // [2][R][H]
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const int ul = tid % UH, rg = tid / UH;       // local unit, row group (4 groups)
  const int u = crank * UH + ul;
  const int b0 = (blockIdx.x / 2) * R;

  for (int i = tid; i < H * UH * 4; i += blockDim.x) {
    const int k = i / (UH * 4), c = i % (UH * 4);
    const int g = c & 3, uu = crank * UH + (c >> 2);
    Wt[i] = __ldg(w_h2h + (size_t)(g * H + uu) * H + k);
  }
  for (int i = tid; i < R * H; i += blockDim.x) {
    const int r = i / H, k = i % H;
    hbuf[i] = (b0 + r < B) ? __ldg(h0 + (size_t)(b0 + r) * ld0 + k) : 0.f;
  }
  float c[RPT];
  float bias[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) bias[g] = __ldg(b_h2h + g * H + u);
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int b = b0 + rg * RPT + r;
    c[r] = b < B ? __ldg(c0 + (size_t)b * ld0 + u) : 0.f;
  }
  float* hbuf_peer = cluster.map_shared_rank(hbuf, crank ^ 1);
  cluster.sync();

  for (int t = 0; t < T; ++t) {
    const float* hcur = hbuf + (t & 1) * R * H;
    float acc[RPT][4], gxv[RPT][4];   // gx loads are issued now and consumed after the recurrence FMAs
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int b = min(b0 + rg * RPT + r, B - 1);
      const float* gp = gx + ((size_t)b * T + t) * 4 * H + u;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        gxv[r][g] = gp[g * H];
        acc[r][g] = bias[g];
      }
    }
#pragma unroll 2
    for (int k = 0; k < H; k += 4) {
      float4 w[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(Wt + (k + kk) * UH * 4 + ul * 4);
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const float4 hv = *reinterpret_cast<const float4*>(hcur + (rg * RPT + r) * H + k);
        acc[r][0] = fmaf(hv.x, w[0].x, acc[r][0]); acc[r][1] = fmaf(hv.x, w[0].y, acc[r][1]);
        acc[r][2] = fmaf(hv.x, w[0].z, acc[r][2]); acc[r][3] = fmaf(hv.x, w[0].w, acc[r][3]);
        acc[r][0] = fmaf(hv.y, w[1].x, acc[r][0]); acc[r][1] = fmaf(hv.y, w[1].y, acc[r][1]);
        acc[r][2] = fmaf(hv.y, w[1].z, acc[r][2]); acc[r][3] = fmaf(hv.y, w[1].w, acc[r][3]);
        acc[r][0] = fmaf(hv.z, w[2].x, acc[r][0]); acc[r][1] = fmaf(hv.z, w[2].y, acc[r][1]);
        acc[r][2] = fmaf(hv.z, w[2].z, acc[r][2]); acc[r][3] = fmaf(hv.z, w[2].w, acc[r][3]);
        acc[r][0] = fmaf(hv.w, w[3].x, acc[r][0]); acc[r][1] = fmaf(hv.w, w[3].y, acc[r][1]);
        acc[r][2] = fmaf(hv.w, w[3].z, acc[r][2]); acc[r][3] = fmaf(hv.w, w[3].w, acc[r][3]);
      }
    }
    float* hnext = hbuf + ((t + 1) & 1) * R * H;
    float* hnext_peer = hbuf_peer + ((t + 1) & 1) * R * H;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int row = rg * RPT + r, b = b0 + row;
      const float ig = sigmoidf_(acc[r][0] + gxv[r][0]), fg = sigmoidf_(acc[r][1] + gxv[r][1]);
      const float gg = tanhf(acc[r][2] + gxv[r][2]), og = sigmoidf_(acc[r][3] + gxv[r][3]);
      const float hp = hcur[row * H + u];
      c[r] = fg * c[r] + ig * gg;
      const float hn = og * tanhf(c[r]);
      hnext[row * H + u] = hn;
      hnext_peer[row * H + u] = hn;
      if (b < B) {
        const size_t o = (size_t)b * T + t;
        float* gp = gx + o * 4 * H + u;
        gp[0] = ig; gp[H] = fg; gp[2 * H] = gg; gp[3 * H] = og;
        hs[o * H + u] = hn;
        hprev[o * H + u] = hp;
        cs[o * H + u] = c[r];
      }
    }
    cluster.sync();
  }
}

// ------------------------------------------------------------------------------------ backward
// gates [B,T,4H] in: saved activations ; out: d(pre-activation gates)  (feeds the wgrad / dgrad GEMMs)
// dhs   [B,T,H]  gradient wrt h_t from the output layer ;  dh0/dc0 rows of a [B,ld0] buffer
template <int H, int RPT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(2 * H)
    lstm_bwd_kernel(float* __restrict__ gates, const float* __restrict__ w_h2h, const float* __restrict__ cs,
                    const float* __restrict__ c0, int ld0, const float* __restrict__ dhs, float* __restrict__ dh0,
                    float* __restrict__ dc0, float* __restrict__ db_i2h, float* __restrict__ db_h2h, int B, int T) {
  constexpr int UH = H / 2;
  constexpr int R = 8 * RPT;       // batch rows per cluster (8 row groups in the reduction mapping)
  constexpr int RC = R / 4;        // rows per thread in the cell mapping (4 row groups)
  constexpr int J = 4 * UH;        // local gate columns
  extern __shared__ __align__(16) float smem_f[];
  float* Wn = smem_f;              // [J][H]      Wn[jl][k] = W_h2h[row(jl)][k], jl = g*UH + ul
  float* dg = Wn + J * H;          // [R][J]      d(pre-activation) of the local gate columns
  float* dhrec = dg + R * J;       // [R][UH]     recurrent dh for the local units (own partial)
  float* dhin = dhrec + R * UH;    // [R][UH]     partial pushed by the peer CTA
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const int b0 = (blockIdx.x / 2) * R;
  // cell mapping: thread -> (local unit, 4 row groups)
  const int ul = tid % UH, rgc = tid / UH;
  const int u = crank * UH + ul;
  // reduction mapping: thread -> (4 consecutive k, 8 row groups)
  const int k4 = (tid % (H / 4)) * 4, rgr = tid / (H / 4);

  for (int i = tid; i < J * H; i += blockDim.x) {
    const int jl = i / H, k = i % H;
    const int g = jl / UH, uu = crank * UH + jl % UH;
    Wn[i] = __ldg(w_h2h + (size_t)(g * H + uu) * H + k);
  }
  for (int i = tid; i < R * UH; i += blockDim.x) { dhrec[i] = 0.f; dhin[i] = 0.f; }
  float dc[RC];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};   // bias gradient of this thread's unit (sum over its rows and all t)
#pragma unroll
  for (int r = 0; r < RC; ++r) dc[r] = 0.f;
  float* dhin_peer = cluster.map_shared_rank(dhin, crank ^ 1);
  cluster.sync();

  for (int t = T - 1; t >= 0; --t) {
    // ---- cell backward for the local units
#pragma unroll
    for (int r = 0; r < RC; ++r) {
      const int row = rgc * RC + r, b = b0 + row;
      float di = 0.f, df = 0.f, dgg = 0.f, dout = 0.f;
      if (b < B) {
        const size_t o = (size_t)b * T + t;
        float* gp = gates + o * 4 * H + u;
        const float ig = gp[0], fg = gp[H], gg = gp[2 * H], og = gp[3 * H];
        const float ct = cs[o * H + u];
        const float cp = t > 0 ? cs[(o - 1) * H + u] : __ldg(c0 + (size_t)b * ld0 + u);
        const float dh = dhs[o * H + u] + dhrec[row * UH + ul] + dhin[row * UH + ul];
        const float tc = tanhf(ct);
        dout = dh * tc * og * (1.f - og);
        float dct = dc[r] + dh * og * (1.f - tc * tc);
        di = dct * gg * ig * (1.f - ig);
        dgg = dct * ig * (1.f - gg * gg);
        df = dct * cp * fg * (1.f - fg);
        dc[r] = dct * fg;
        gp[0] = di; gp[H] = df; gp[2 * H] = dgg; gp[3 * H] = dout;
        bsum[0] += di; bsum[1] += df; bsum[2] += dgg; bsum[3] += dout;
      }
      dg[row * J + 0 * UH + ul] = di;
      dg[row * J + 1 * UH + ul] = df;
      dg[row * J + 2 * UH + ul] = dgg;
      dg[row * J + 3 * UH + ul] = dout;
    }
    cluster.sync();   // dg complete; both CTAs have consumed dhin of the previous step
    // ---- partial dh_rec[r][k] = sum_j dg[r][j] Wn[j][k] over the local columns, all k
    float acc[RPT][4];
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
#pragma unroll 2
    for (int j = 0; j < J; j += 4) {
      float4 w[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) w[jj] = *reinterpret_cast<const float4*>(Wn + (j + jj) * H + k4);
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const float4 d = *reinterpret_cast<const float4*>(dg + (rgr * RPT + r) * J + j);
        acc[r][0] = fmaf(d.x, w[0].x, acc[r][0]); acc[r][1] = fmaf(d.x, w[0].y, acc[r][1]);
        acc[r][2] = fmaf(d.x, w[0].z, acc[r][2]); acc[r][3] = fmaf(d.x, w[0].w, acc[r][3]);
        acc[r][0] = fmaf(d.y, w[1].x, acc[r][0]); acc[r][1] = fmaf(d.y, w[1].y, acc[r][1]);
        acc[r][2] = fmaf(d.y, w[1].z, acc[r][2]); acc[r][3] = fmaf(d.y, w[1].w, acc[r][3]);
        acc[r][0] = fmaf(d.z, w[2].x, acc[r][0]); acc[r][1] = fmaf(d.z, w[2].y, acc[r][1]);
        acc[r][2] = fmaf(d.z, w[2].z, acc[r][2]); acc[r][3] = fmaf(d.z, w[2].w, acc[r][3]);
        acc[r][0] = fmaf(d.w, w[3].x, acc[r][0]); acc[r][1] = fmaf(d.w, w[3].y, acc[r][1]);
        acc[r][2] = fmaf(d.w, w[3].z, acc[r][2]); acc[r][3] = fmaf(d.w, w[3].w, acc[r][3]);
      }
    }
    // k in the local half stays here, the other half goes to the peer's dhin
    const bool mine = (k4 / UH) == crank;
    float* dst = mine ? dhrec : dhin_peer;
#pragma unroll
    for (int r = 0; r < RPT; ++r)
      *reinterpret_cast<float4*>(dst + (rgr * RPT + r) * UH + (k4 % UH)) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    cluster.sync();   // dhrec / dhin ready for step t-1
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (db_i2h) atomicAdd(db_i2h + g * H + u, bsum[g]);
    if (db_h2h) atomicAdd(db_h2h + g * H + u, bsum[g]);
  }
#pragma unroll
  for (int r = 0; r < RC; ++r) {
    const int row = rgc * RC + r, b = b0 + row;
    if (b < B) {
      dh0[(size_t)b * ld0 + u] = dhrec[row * UH + ul] + dhin[row * UH + ul];
      dc0[(size_t)b * ld0 + u] = dc[r];
    }
  }
}

template <int H, int RPT>
int launch_fwd(float* gx, const float* w, const float* bh, const float* h0, const float* c0, int ld0, float* hs,
               float* hprev, float* cs, int B, int T, cudaStream_t st) {
  constexpr int R = 4 * RPT;
  const size_t smem = ((size_t)H * (H / 2) * 4 + 2 * R * H) * sizeof(float);
  MSX_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel<H, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int clusters = (B + R - 1) / R;
  lstm_fwd_kernel<H, RPT><<<clusters * 2, 2 * H, smem, st>>>(gx, w, bh, h0, c0, ld0, hs, hprev, cs, B, T);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

template <int H, int RPT>
int launch_bwd(float* gates, const float* w, const float* cs, const float* c0, int ld0, const float* dhs, float* dh0,
               float* dc0, float* dbi, float* dbh, int B, int T, cudaStream_t st) {
  constexpr int R = 8 * RPT;
  const size_t smem = ((size_t)2 * H * H + (size_t)R * 2 * H + 2 * R * (H / 2)) * sizeof(float);
  MSX_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel<H, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int clusters = (B + R - 1) / R;
  lstm_bwd_kernel<H, RPT><<<clusters * 2, 2 * H, smem, st>>>(gates, w, cs, c0, ld0, dhs, dh0, dc0, dbi, dbh, B, T);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}


// ------------------------------------------------------------------------------------ any hidden size
// --d-hidden is a free flag of the reference's CLI (VarAutoEncoder/config.py, LSTMConfig.hidden_dim); the resident
// kernels above cover the sizes whose W_h2h halves fit a CTA pair (32 / 64 / 128).  For every other H (1 <= H <= 512)
// these kernels run the same recurrence with W_h2h streamed from L2 every step: one CTA per 8 batch rows, h double-buffered
// in shared memory, 64 units x 32 reduction indices of W staged per tile.  Exact fp32 FMA arithmetic like the resident
// kernels; a functional path (the W re-reads make it L2-bound), not a tuned one.
constexpr int kGenR = 8, kGenUT = 64, kGenKT = 32, kGenThreads = 256, kGenMaxH = 512;

__global__ void __launch_bounds__(kGenThreads)
    lstm_gen_fwd_kernel(float* __restrict__ gx, const float* __restrict__ w_h2h, const float* __restrict__ b_h2h,
                        const float* __restrict__ h0, const float* __restrict__ c0, int ld0, float* __restrict__ hs,
                        float* __restrict__ hprev, float* __restrict__ cs, int B, int T, int H) {
  extern __shared__ __align__(16) float smem_f[];
  float* hbuf = smem_f;                                   // [2][R][H]
  float* Ws = hbuf + 2 * kGenR * H;                       // [4 * UT][KT + 1]
  const int tid = threadIdx.x;
  const int ul = tid % kGenUT, rg = tid / kGenUT;         // local unit, row group (4 groups x 2 rows)
  const int b0 = blockIdx.x * kGenR;
  for (int i = tid; i < kGenR * H; i += kGenThreads) {
    const int r = i / H, k = i % H;
    hbuf[i] = (b0 + r < B) ? __ldg(h0 + (size_t)(b0 + r) * ld0 + k) : 0.f;
  }
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    const float* hcur = hbuf + (t & 1) * kGenR * H;
    float* hnext = hbuf + ((t + 1) & 1) * kGenR * H;
    for (int u0 = 0; u0 < H; u0 += kGenUT) {
      const int u = u0 + ul;
      float acc[2][4];
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[0][g] = acc[1][g] = u < H ? __ldg(b_h2h + g * H + u) : 0.f;
      for (int k0 = 0; k0 < H; k0 += kGenKT) {
        __syncthreads();                                  // the previous tile has been consumed
        for (int i = tid; i < 4 * kGenUT * kGenKT; i += kGenThreads) {
          const int row = i / kGenKT, kk = i % kGenKT;    // row = g * UT + local unit
          const int uu = u0 + row % kGenUT, k = k0 + kk;
          Ws[row * (kGenKT + 1) + kk] = (uu < H && k < H) ? __ldg(w_h2h + (size_t)((row / kGenUT) * H + uu) * H + k) : 0.f;
        }
        __syncthreads();
        const int kn = min(kGenKT, H - k0);
        for (int kk = 0; kk < kn; ++kk) {
          const float hv0 = hcur[(2 * rg) * H + k0 + kk], hv1 = hcur[(2 * rg + 1) * H + k0 + kk];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float w = Ws[(g * kGenUT + ul) * (kGenKT + 1) + kk];
            acc[0][g] = fmaf(hv0, w, acc[0][g]);
            acc[1][g] = fmaf(hv1, w, acc[1][g]);
          }
        }
      }
      if (u < H) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int row = 2 * rg + r, b = b0 + row;
          float hn = 0.f;
          if (b < B) {
            const size_t o = (size_t)b * T + t;
            float* gp = gx + o * 4 * H + u;
            const float ig = sigmoidf_(acc[r][0] + gp[0]), fg = sigmoidf_(acc[r][1] + gp[H]);
            const float gg = tanhf(acc[r][2] + gp[2 * H]), og = sigmoidf_(acc[r][3] + gp[3 * H]);
            const float cp = t > 0 ? cs[(o - 1) * H + u] : __ldg(c0 + (size_t)b * ld0 + u);   // this thread's own store of step t-1
            const float cn = fg * cp + ig * gg;
            hn = og * tanhf(cn);
            gp[0] = ig; gp[H] = fg; gp[2 * H] = gg; gp[3 * H] = og;
            hs[o * H + u] = hn;
            hprev[o * H + u] = hcur[row * H + u];
            cs[o * H + u] = cn;
          }
          hnext[row * H + u] = hn;
        }
      }
    }
    __syncthreads();                                      // h_t complete before step t + 1 reads it
  }
}

__global__ void __launch_bounds__(kGenThreads)
    lstm_gen_bwd_kernel(float* __restrict__ gates, const float* __restrict__ w_h2h, const float* __restrict__ cs,
                        const float* __restrict__ c0, int ld0, const float* __restrict__ dhs, float* __restrict__ dh0,
                        float* __restrict__ dc0, float* __restrict__ db_i2h, float* __restrict__ db_h2h, int B, int T, int H) {
  extern __shared__ __align__(16) float smem_f[];
  float* dg = smem_f;                                     // [R][4H]  d(pre-activation) of step t
  float* dhrec = dg + kGenR * 4 * H;                      // [R][H]   recurrent dh arriving from step t + 1
  float* dcs = dhrec + kGenR * H;                         // [R][H]   dc carried to step t - 1
  float* bacc = dcs + kGenR * H;                          // [4H]     bias gradient of this CTA's rows
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * kGenR;
  for (int i = tid; i < kGenR * H; i += kGenThreads) { dhrec[i] = 0.f; dcs[i] = 0.f; }
  for (int i = tid; i < 4 * H; i += kGenThreads) bacc[i] = 0.f;
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    // ---- cell backward: a thread owns unit u for all rows (so its bias sums need no atomics)
    for (int u = tid; u < H; u += kGenThreads) {
      float bs[4] = {0.f, 0.f, 0.f, 0.f};
      for (int row = 0; row < kGenR; ++row) {
        const int b = b0 + row;
        float di = 0.f, df = 0.f, dgg = 0.f, dout = 0.f;
        if (b < B) {
          const size_t o = (size_t)b * T + t;
          float* gp = gates + o * 4 * H + u;
          const float ig = gp[0], fg = gp[H], gg = gp[2 * H], og = gp[3 * H];
          const float ct = cs[o * H + u];
          const float cp = t > 0 ? cs[(o - 1) * H + u] : __ldg(c0 + (size_t)b * ld0 + u);
          const float dh = dhs[o * H + u] + dhrec[row * H + u];
          const float tc = tanhf(ct);
          dout = dh * tc * og * (1.f - og);
          const float dct = dcs[row * H + u] + dh * og * (1.f - tc * tc);
          di = dct * gg * ig * (1.f - ig);
          dgg = dct * ig * (1.f - gg * gg);
          df = dct * cp * fg * (1.f - fg);
          dcs[row * H + u] = dct * fg;
          gp[0] = di; gp[H] = df; gp[2 * H] = dgg; gp[3 * H] = dout;
          bs[0] += di; bs[1] += df; bs[2] += dgg; bs[3] += dout;
        }
        float* d = dg + row * 4 * H + u;
        d[0] = di; d[H] = df; d[2 * H] = dgg; d[3 * H] = dout;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) bacc[g * H + u] += bs[g];
    }
    __syncthreads();
    // ---- dh_rec[r][k] = sum_j dg[r][j] W[j][k]: a thread owns column k (coalesced W rows out of L2)
    for (int k = tid; k < H; k += kGenThreads) {
      float acc[kGenR];
#pragma unroll
      for (int r = 0; r < kGenR; ++r) acc[r] = 0.f;
      for (int j = 0; j < 4 * H; ++j) {
        const float w = __ldg(w_h2h + (size_t)j * H + k);
#pragma unroll
        for (int r = 0; r < kGenR; ++r) acc[r] = fmaf(dg[r * 4 * H + j], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < kGenR; ++r) dhrec[r * H + k] = acc[r];
    }
    __syncthreads();
  }
  for (int i = tid; i < 4 * H; i += kGenThreads) {
    if (db_i2h) atomicAdd(db_i2h + i, bacc[i]);
    if (db_h2h) atomicAdd(db_h2h + i, bacc[i]);
  }
  for (int i = tid; i < kGenR * H; i += kGenThreads) {
    const int row = i / H, u = i % H, b = b0 + row;
    if (b < B) {
      dh0[(size_t)b * ld0 + u] = dhrec[i];
      dc0[(size_t)b * ld0 + u] = dcs[i];
    }
  }
}

int launch_gen_fwd(float* gx, const float* w, const float* bh, const float* h0, const float* c0, int ld0, float* hs,
                   float* hprev, float* cs, int B, int T, int H, cudaStream_t st) {
  const size_t smem = ((size_t)2 * kGenR * H + 4 * kGenUT * (kGenKT + 1)) * sizeof(float);
  MSX_CUDA(cudaFuncSetAttribute(lstm_gen_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lstm_gen_fwd_kernel<<<(B + kGenR - 1) / kGenR, kGenThreads, smem, st>>>(gx, w, bh, h0, c0, ld0, hs, hprev, cs, B, T, H);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

int launch_gen_bwd(float* gates, const float* w, const float* cs, const float* c0, int ld0, const float* dhs, float* dh0,
                   float* dc0, float* dbi, float* dbh, int B, int T, int H, cudaStream_t st) {
  const size_t smem = ((size_t)kGenR * 4 * H + 2 * kGenR * H + 4 * H) * sizeof(float);
  MSX_CUDA(cudaFuncSetAttribute(lstm_gen_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lstm_gen_bwd_kernel<<<(B + kGenR - 1) / kGenR, kGenThreads, smem, st>>>(gates, w, cs, c0, ld0, dhs, dh0, dc0, dbi, dbh, B, T, H);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

}  // namespace

// MSX_LSTM_GENERIC=1 (tests): the any-size kernels also take H = 32 / 64 / 128
static bool lstm_force_generic() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MSX_LSTM_GENERIC");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

extern "C" int msx_lstm_fwd(float* gx_inout, const float* w_h2h, const float* b_h2h, const float* h0, const float* c0,
                            int ld0, float* hs, float* hprev, float* cs, int B, int T, int H, void* stream) {
  MSX_REQUIRE(gx_inout && w_h2h && b_h2h && h0 && c0 && hs && hprev && cs, "msx_lstm_fwd: null pointer");
  if (B == 0 || T == 0) return MSX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool small = B <= msx_num_sms() * 4;   // few rows: spread them over more SMs with 8-row tiles
  MSX_REQUIRE(H >= 1 && H <= kGenMaxH, "msx_lstm_fwd: hidden size %d unsupported (1 .. %d)", H, kGenMaxH);
  if (lstm_force_generic()) return launch_gen_fwd(gx_inout, w_h2h, b_h2h, h0, c0, ld0, hs, hprev, cs, B, T, H, st);
  switch (H) {
    case 32: return small ? launch_fwd<32, 2>(gx_inout, w_h2h, b_h2h, h0, c0, ld0, hs, hprev, cs, B, T, st)
                          : launch_fwd<32, 8>(gx_inout, w_h2h, b_h2h, h0, c0, ld0, hs, hprev, cs, B, T, st);
    case 64: return small ? launch_fwd<64, 2>(gx_inout, w_h2h, b_h2h, h0, c0, ld0, hs, hprev, cs, B, T, st)
                          : launch_fwd<64, 8>(gx_inout, w_h2h, b_h2h, h0, c0, ld0, hs, hprev, cs, B, T, st);
    case 128: return small ? launch_fwd<128, 2>(gx_inout, w_h2h, b_h2h, h0, c0, ld0, hs, hprev, cs, B, T, st)
                           : launch_fwd<128, 8>(gx_inout, w_h2h, b_h2h, h0, c0, ld0, hs, hprev, cs, B, T, st);
    default: return launch_gen_fwd(gx_inout, w_h2h, b_h2h, h0, c0, ld0, hs, hprev, cs, B, T, H, st);
  }
}

extern "C" int msx_lstm_bwd(float* gates_inout, const float* w_h2h, const float* cs, const float* c0, int ld0,
                            const float* dhs, float* dh0, float* dc0, float* db_i2h, float* db_h2h, int B, int T, int H,
                            void* stream) {
  MSX_REQUIRE(gates_inout && w_h2h && cs && c0 && dhs && dh0 && dc0, "msx_lstm_bwd: null pointer");
  if (B == 0 || T == 0) return MSX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool small = B <= msx_num_sms() * 4;
  MSX_REQUIRE(H >= 1 && H <= kGenMaxH, "msx_lstm_bwd: hidden size %d unsupported (1 .. %d)", H, kGenMaxH);
  if (lstm_force_generic()) return launch_gen_bwd(gates_inout, w_h2h, cs, c0, ld0, dhs, dh0, dc0, db_i2h, db_h2h, B, T, H, st);
  switch (H) {
    case 32: return small ? launch_bwd<32, 1>(gates_inout, w_h2h, cs, c0, ld0, dhs, dh0, dc0, db_i2h, db_h2h, B, T, st)
                          : launch_bwd<32, 4>(gates_inout, w_h2h, cs, c0, ld0, dhs, dh0, dc0, db_i2h, db_h2h, B, T, st);
    case 64: return small ? launch_bwd<64, 1>(gates_inout, w_h2h, cs, c0, ld0, dhs, dh0, dc0, db_i2h, db_h2h, B, T, st)
                          : launch_bwd<64, 4>(gates_inout, w_h2h, cs, c0, ld0, dhs, dh0, dc0, db_i2h, db_h2h, B, T, st);
    case 128: return small ? launch_bwd<128, 1>(gates_inout, w_h2h, cs, c0, ld0, dhs, dh0, dc0, db_i2h, db_h2h, B, T, st)
                           : launch_bwd<128, 4>(gates_inout, w_h2h, cs, c0, ld0, dhs, dh0, dc0, db_i2h, db_h2h, B, T, st);
    default: return launch_gen_bwd(gates_inout, w_h2h, cs, c0, ld0, dhs, dh0, dc0, db_i2h, db_h2h, B, T, H, st);
  }
}
