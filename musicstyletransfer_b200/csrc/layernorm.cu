// K2d — fused (dropout) + residual + LayerNorm, forward and backward.
//
// Replaces `self.ln1(x + self.dropout(x_att))` / `self.ln2(x + self.dropout(x_att))`
// (/root/reference/music_style_transfer/VarAutoEncoder/transformer.py:155,158) and the decoder's
// `self.ln3(x_att + self.dropout(x_att))` (:200), i.e. gluon.nn.Dropout + broadcast add +
// gluon.nn.LayerNorm (last axis, biased variance, eps 1e-5), plus their autograd backward.
//
// One warp per row, row kept in registers (D <= 1024), two-pass mean / variance exactly as
// (x-mean)^2 averaged, so fp32 results track the un-fused oracle closely.  HBM-bound: reads x and y,
// writes out (+ 8 B of statistics per row).  Backward re-creates s = x + drop(y) from the saved
// inputs and the counter-hash mask, reduces dgamma / dbeta per CTA before one atomic per column; for D <= 256 the
// rows it will process next are prefetched into per-warp shared-memory slots with cp.async (kAsync).  Grids are one
// resident wave of CTAs.  The _ex entry points add the bf16-variant options: a bfloat16 copy of the output / of the
// y-gradient for the bf16 GEMMs, and a bfloat16 y input.
#include "msx_common.cuh"

namespace {

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  uint2 r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.x) : "f"(b), "f"(a));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.y) : "f"(d), "f"(c));
  return r;
}

constexpr int kMaxPer = 8;      // max groups per lane -> D <= 1024 (vector path), D <= 256 (scalar path)
constexpr int kWarps = 8;

template <int VEC>
__device__ __forceinline__ int elem_index(int i, int lane, int j) {
  return VEC == 4 ? (i * 32 + lane) * 4 + j : i * 32 + lane;
}

// s = x + drop(y) for this lane's elements of row `row`
// y_bf16: y is bfloat16 (bf16 variant: the Dense output that only this LayerNorm reads), vector path only
template <int VEC, int kPer>
__device__ __forceinline__ void load_sum(const float* x, const float* y, bool y_bf16, size_t row, size_t rng_row,
                                         int D, int nper, int lane, float p, float inv_keep, unsigned long long seed,
                                         unsigned site, float s[kPer][VEC], float keep[kPer][VEC]) {
  // row addresses x / y (0 when they point at this warp's shared-memory copy of the row), rng_row is the row's index in
  // the tensor (dropout counter)
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    if (i >= nper) break;
    if (VEC == 4) {
      const int e = (i * 32 + lane) * 4;
      const float4 xv = *reinterpret_cast<const float4*>(x + row * D + e);
      float4 yv;
      if (y_bf16) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(y) + row * D + e);
        yv = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                         __uint_as_float(u.y & 0xFFFF0000u));
      } else {
        yv = *reinterpret_cast<const float4*>(y + row * D + e);
      }
      float k4[4] = {1.f, 1.f, 1.f, 1.f};
      if (p > 0.f) dropout_scale4(seed, site, (rng_row * D + e) >> 2, p, inv_keep, k4);
      s[i][0] = xv.x + yv.x * k4[0];
      s[i][1 % VEC] = xv.y + yv.y * k4[1];
      s[i][2 % VEC] = xv.z + yv.z * k4[2];
      s[i][3 % VEC] = xv.w + yv.w * k4[3];
#pragma unroll
      for (int j = 0; j < VEC; ++j) keep[i][j] = k4[j];
    } else {
      const int e = i * 32 + lane;
      float k4[4] = {1.f, 1.f, 1.f, 1.f};
      if (p > 0.f) dropout_scale4(seed, site, (rng_row * D + e) >> 2, p, inv_keep, k4);
      const float kk = k4[(rng_row * D + e) & 3];
      s[i][0] = x[row * D + e] + y[row * D + e] * kk;
      keep[i][0] = kk;
    }
  }
}

template <int VEC, int kPer>
__global__ void __launch_bounds__(kWarps * 32) add_ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int y_bf16,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float* __restrict__ out,
                                                                 unsigned short* __restrict__ out16,
                                                                 unsigned short* __restrict__ out16lo,
                                                                 float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                 long long M, int D, float eps, float p, float inv_keep,
                                                                 unsigned long long seed, const unsigned long long* seed_ctr,
                                                                 unsigned site) {
  pdl_entry();
  seed = msx_eff_seed(seed, seed_ctr);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nper = D / (32 * VEC);
  const float invD = 1.f / D;
  for (long long row = (long long)blockIdx.x * kWarps + warp; row < M; row += (long long)gridDim.x * kWarps) {
    float s[kPer][VEC], keep[kPer][VEC];
    load_sum<VEC, kPer>(x, y, y_bf16 != 0, (size_t)row, (size_t)row, D, nper, lane, p, inv_keep, seed, site, s, keep);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kPer; ++i)
      if (i < nper)
#pragma unroll
        for (int j = 0; j < VEC; ++j) sum += s[i][j];
    const float mean = warp_sum(sum) * invD;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < kPer; ++i)
      if (i < nper)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float d = s[i][j] - mean;
          var = fmaf(d, d, var);
        }
    var = warp_sum(var) * invD;
    const float rstd = 1.f / sqrtf(var + eps);
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      if (i >= nper) break;
      if (VEC == 4) {
        const int e = (i * 32 + lane) * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + e));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + e));
        float4 o;
        o.x = (s[i][0] - mean) * rstd * g.x + b.x;
        o.y = (s[i][1 % VEC] - mean) * rstd * g.y + b.y;
        o.z = (s[i][2 % VEC] - mean) * rstd * g.z + b.z;
        o.w = (s[i][3 % VEC] - mean) * rstd * g.w + b.w;
        *reinterpret_cast<float4*>(out + (size_t)row * D + e) = o;
        if (out16lo) {                                   // hi / lo planes for the p3 GEMMs
          uint2 hi, lo;
          split4_bf16(o.x, o.y, o.z, o.w, hi, lo);
          *reinterpret_cast<uint2*>(out16 + (size_t)row * D + e) = hi;
          *reinterpret_cast<uint2*>(out16lo + (size_t)row * D + e) = lo;
        } else if (out16) {
          *reinterpret_cast<uint2*>(out16 + (size_t)row * D + e) = pack4_bf16(o.x, o.y, o.z, o.w);
        }
      } else {
        const int e = i * 32 + lane;
        out[(size_t)row * D + e] = (s[i][0] - mean) * rstd * __ldg(gamma + e) + __ldg(beta + e);
      }
    }
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
  }
}

// ds = rstd * (g*dout - mean(g*dout) - xhat * mean(g*dout*xhat));  dres = ds;  dy = ds * keep
// kAsync (vector path, kPer <= 2): the three input rows of the NEXT row this warp will process are already on their way
// into a per-warp shared-memory slot (cp.async, each lane copies exactly the elements it will read back, so a
// cp.async.wait_group is the only synchronisation) while the current row is reduced and stored.  The kernel is
// latency-bound (ncu: 0.3 eligible warps per scheduler, 72 % of the stall cycles on the row loads): this keeps one row
// per warp in flight all the time instead of only between the load and the first reduction.
__device__ __forceinline__ void ln_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ln_cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
constexpr int kLnStages = 3;   // rows in flight per warp: the one being reduced + two on their way (2 CTAs / SM, no spills)

template <int VEC, int kPer, bool kAsync>
__global__ void __launch_bounds__(kWarps * 32, (kAsync ? 2 : (kPer <= 2 ? 3 : 1))) add_ln_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ y, int y_bf16, const float* __restrict__ gamma,
    const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ dout,
    float* __restrict__ dres, float* __restrict__ dy, unsigned short* __restrict__ dy16, float* __restrict__ dgamma,
    float* __restrict__ dbeta, float* __restrict__ dybias, long long M, int D, float p, float inv_keep, unsigned long long seed,
    const unsigned long long* seed_ctr, unsigned site, int accumulate_dres, int fuse_xy) {
  pdl_entry();
  seed = msx_eff_seed(seed, seed_ctr);
  __shared__ float red[kWarps][32 * VEC + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nper = D / (32 * VEC);
  const float invD = 1.f / D;
  float dg[kPer][VEC], db[kPer][VEC], dyb[kPer][VEC];   // dyb: column sums of the y-gradient = bias grad of the Dense that made y
#pragma unroll
  for (int i = 0; i < kPer; ++i)
#pragma unroll
    for (int j = 0; j < VEC; ++j) dg[i][j] = db[i][j] = dyb[i][j] = 0.f;

  extern __shared__ __align__(16) unsigned char ln_dyn[];
  // per warp and stage: x | y | dout slots of kPer * 32 float4 (a bf16 y uses the first half of its slot)
  float* wbuf = reinterpret_cast<float*>(ln_dyn) + (size_t)warp * kLnStages * 3 * kPer * 128;
  const long long row_first = (long long)blockIdx.x * kWarps + warp, row_stride = (long long)gridDim.x * kWarps;
  auto issue = [&](long long r, int stage) {
    if (r < M) {
      float* sb = wbuf + (size_t)stage * 3 * kPer * 128;
#pragma unroll
      for (int i = 0; i < kPer; ++i) {
        if (i >= nper) break;
        const int e = (i * 32 + lane) * 4;
        ln_cp_async16(sb + e, x + (size_t)r * D + e);
        if (y_bf16) ln_cp_async8(reinterpret_cast<unsigned short*>(sb + kPer * 128) + e, reinterpret_cast<const unsigned short*>(y) + (size_t)r * D + e);
        else ln_cp_async16(sb + kPer * 128 + e, y + (size_t)r * D + e);
        ln_cp_async16(sb + 2 * kPer * 128 + e, dout + (size_t)r * D + e);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (kAsync) {
#pragma unroll
    for (int st = 0; st < kLnStages - 1; ++st) issue(row_first + st * row_stride, st);
  }
  int it = 0;
  for (long long row = row_first; row < M; row += row_stride, ++it) {
    // s = x + drop(y); the keep mask is held as one bit per element
    float s[kPer][VEC];
    unsigned keepbits = 0u;
    const float* xsrc = x;
    const float* ysrc = y;
    const float* dsrc = dout;
    size_t arow = (size_t)row;
    if (kAsync) {
      issue(row + (kLnStages - 1) * row_stride, (it + kLnStages - 1) % kLnStages);   // keep kLnStages - 1 rows on their way
      asm volatile("cp.async.wait_group %0;" ::"n"(kLnStages - 1) : "memory");
      xsrc = wbuf + (size_t)(it % kLnStages) * 3 * kPer * 128;
      ysrc = xsrc + kPer * 128;
      dsrc = xsrc + 2 * kPer * 128;
      arow = 0;
    }
    {
      float keep[kPer][VEC];
      load_sum<VEC, kPer>(xsrc, ysrc, y_bf16 != 0, arow, (size_t)row, D, nper, lane, p, inv_keep, seed, site, s, keep);
#pragma unroll
      for (int i = 0; i < kPer; ++i)
        if (i < nper)
#pragma unroll
          for (int j = 0; j < VEC; ++j) keepbits |= (keep[i][j] > 0.f ? 1u : 0u) << (i * VEC + j);
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    float go[kPer][VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      if (i >= nper) break;
      float dv[VEC], gv[VEC];
      if (VEC == 4) {
        const int e = (i * 32 + lane) * 4;
        const float4 d4 = *reinterpret_cast<const float4*>(dsrc + arow * D + e);
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + e));
        dv[0] = d4.x; dv[1 % VEC] = d4.y; dv[2 % VEC] = d4.z; dv[3 % VEC] = d4.w;
        gv[0] = g4.x; gv[1 % VEC] = g4.y; gv[2 % VEC] = g4.z; gv[3 % VEC] = g4.w;
      } else {
        const int e = i * 32 + lane;
        dv[0] = dout[(size_t)row * D + e];
        gv[0] = __ldg(gamma + e);
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float d = dv[j];
        const float xh = (s[i][j] - mean) * rstd;
        s[i][j] = xh;
        dg[i][j] = fmaf(d, xh, dg[i][j]);
        db[i][j] += d;
        const float gd = d * gv[j];
        go[i][j] = gd;
        s1 += gd;
        s2 = fmaf(gd, xh, s2);
      }
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      if (i >= nper) break;
      float dr[VEC], dyv[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float ds = rstd * (go[i][j] - s1 - s[i][j] * s2);
        // fuse_xy: x and y are the same tensor (decoder's ln3(f + drop(f))) -> one gradient ds*(1+keep)
        const float kp = ((keepbits >> (i * VEC + j)) & 1u) ? inv_keep : 0.f;
        dyv[j] = ds * kp;
        dr[j] = fuse_xy ? ds * (1.f + kp) : ds;
        dyb[i][j] += fuse_xy ? dr[j] : dyv[j];
      }
      if (VEC == 4) {
        const size_t o = (size_t)row * D + (i * 32 + lane) * 4;
        if (dy && !fuse_xy)
          *reinterpret_cast<float4*>(dy + o) = make_float4(dyv[0], dyv[1 % VEC], dyv[2 % VEC], dyv[3 % VEC]);
        float4 r4 = make_float4(dr[0], dr[1 % VEC], dr[2 % VEC], dr[3 % VEC]);
        if (dy16)      // bf16 copy of the gradient of the Dense output y (what its dgrad / wgrad GEMMs read)
          *reinterpret_cast<uint2*>(dy16 + o) = fuse_xy ? pack4_bf16(r4.x, r4.y, r4.z, r4.w)
                                                        : pack4_bf16(dyv[0], dyv[1 % VEC], dyv[2 % VEC], dyv[3 % VEC]);
        if (accumulate_dres) {
          const float4 o4 = *reinterpret_cast<const float4*>(dres + o);
          r4.x += o4.x; r4.y += o4.y; r4.z += o4.z; r4.w += o4.w;
        }
        *reinterpret_cast<float4*>(dres + o) = r4;
      } else {
        const size_t o = (size_t)row * D + i * 32 + lane;
        if (dy && !fuse_xy) dy[o] = dyv[0];
        dres[o] = accumulate_dres ? dres[o] + dr[0] : dr[0];
      }
    }
  }
  // CTA-level reduction of dgamma/dbeta, then one atomic per column
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    if (i >= nper) break;
    for (int pass = 0; pass < (dybias ? 3 : 2); ++pass) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < VEC; ++j) red[warp][lane * VEC + j] = pass == 0 ? dg[i][j] : pass == 1 ? db[i][j] : dyb[i][j];
      __syncthreads();
      for (int c = threadIdx.x; c < 32 * VEC; c += kWarps * 32) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += red[w][c];
        const int l = c / VEC, j = c % VEC;
        const int e = elem_index<VEC>(i, l, j);
        atomicAdd((pass == 0 ? dgamma : pass == 1 ? dbeta : dybias) + e, t);
      }
    }
  }
}

}  // namespace

// Grid = exactly one wave of resident CTAs (persistent row loop): with the former 8 CTAs per SM the last partial wave
// ran at a fraction of the machine and every CTA paid the dgamma / dbeta atomics.
template <typename K>
static int ln_wave_grid(K kernel, int dyn_smem, long long M) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarps * 32, dyn_smem) != cudaSuccess || per_sm < 1) per_sm = 2;
  const long long want = (M + kWarps - 1) / kWarps;
  const long long wave = (long long)msx_num_sms() * per_sm;
  return (int)(want < wave ? want : wave);
}

extern "C" int msx_add_ln_fwd_p(const float* x, const void* y_any, int y_bf16, const float* gamma, const float* beta,
                                float* out, void* out_bf16, void* out_bf16_lo, float* mean, float* rstd, long long M, int D,
                                float eps, float drop_p, unsigned long long seed, unsigned site, void* stream) {
  const float* y = reinterpret_cast<const float*>(y_any);
  MSX_REQUIRE(!out_bf16_lo || out_bf16, "msx_add_ln_fwd_p: a lo plane needs the hi plane");
  MSX_REQUIRE(((uintptr_t)out_bf16_lo & 7) == 0, "msx_add_ln_fwd_p: the lo plane must be 8-byte aligned");
  unsigned short* out16lo = reinterpret_cast<unsigned short*>(out_bf16_lo);
  MSX_REQUIRE(x && y && gamma && beta && out && mean && rstd, "msx_add_ln_fwd: null pointer");
  MSX_REQUIRE(D % 32 == 0 && D >= 32, "msx_add_ln_fwd: D must be a multiple of 32");
  MSX_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "msx_add_ln_fwd: bad dropout");
  if (M == 0) return MSX_OK;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const bool vec = (D % 128 == 0) && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)out | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0 &&
                   ((uintptr_t)out_bf16 & 7) == 0;
  MSX_REQUIRE(!(out_bf16 || y_bf16) || vec, "msx_add_ln_fwd: bf16 tensors need D %% 128 == 0 and 16-byte aligned tensors");
  unsigned short* out16 = reinterpret_cast<unsigned short*>(out_bf16);
  cudaStream_t st = (cudaStream_t)stream;
#define LN_FWD(V, P) MSX_CUDA(msx_launch(add_ln_fwd_kernel<V, P>, dim3(ln_wave_grid(add_ln_fwd_kernel<V, P>, 0, M)), dim3(kWarps * 32), 0, st, x, y, y_bf16, gamma, beta, out, out16, out16lo, mean, rstd, M, D, eps, drop_p, inv_keep, seed, msx_step_counter(), site))
  if (vec) {
    MSX_REQUIRE(D <= 128 * kMaxPer, "msx_add_ln_fwd: D too large");
    const int nper = D / 128;
    if (nper <= 1) LN_FWD(4, 1); else if (nper <= 2) LN_FWD(4, 2); else if (nper <= 4) LN_FWD(4, 4); else LN_FWD(4, 8);
  } else {
    MSX_REQUIRE(D <= 32 * kMaxPer, "msx_add_ln_fwd: D=%d unsupported (need D%%128==0 or D<=256)", D);
    const int nper = D / 32;
    if (nper <= 2) LN_FWD(1, 2); else LN_FWD(1, 8);
  }
#undef LN_FWD
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_add_ln_fwd_ex(const float* x, const void* y_any, int y_bf16, const float* gamma, const float* beta,
                                 float* out, void* out_bf16, float* mean, float* rstd, long long M, int D, float eps,
                                 float drop_p, unsigned long long seed, unsigned site, void* stream) {
  return msx_add_ln_fwd_p(x, y_any, y_bf16, gamma, beta, out, out_bf16, nullptr, mean, rstd, M, D, eps, drop_p, seed, site, stream);
}

extern "C" int msx_add_ln_fwd(const float* x, const float* y, const float* gamma, const float* beta, float* out,
                              float* mean, float* rstd, long long M, int D, float eps, float drop_p,
                              unsigned long long seed, unsigned site, void* stream) {
  return msx_add_ln_fwd_ex(x, y, 0, gamma, beta, out, nullptr, mean, rstd, M, D, eps, drop_p, seed, site, stream);
}

extern "C" int msx_add_ln_bwd_ex(const float* x, const void* y_any, int y_bf16, const float* gamma, const float* mean,
                                 const float* rstd, const float* dout, float* dres, float* dy, void* dy_bf16, float* dgamma,
                                 float* dbeta, float* dybias, long long M, int D, float drop_p, unsigned long long seed,
                                 unsigned site, int accumulate_dres, int fuse_xy, void* stream) {
  const float* y = reinterpret_cast<const float*>(y_any);
  MSX_REQUIRE(!(y_bf16 && fuse_xy), "msx_add_ln_bwd: fuse_xy (x and y alias) needs an fp32 y");
  MSX_REQUIRE(x && y && gamma && mean && rstd && dout && dres && dgamma && dbeta, "msx_add_ln_bwd: null pointer");
  MSX_REQUIRE(D % 32 == 0 && D >= 32, "msx_add_ln_bwd: D must be a multiple of 32");
  if (M == 0) return MSX_OK;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const bool vec = (D % 128 == 0) && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)dout | (uintptr_t)dres | (uintptr_t)dy |
                                       (uintptr_t)gamma) & 15) == 0 && ((uintptr_t)dy_bf16 & 7) == 0;
  MSX_REQUIRE(!(dy_bf16 || y_bf16) || vec, "msx_add_ln_bwd: bf16 tensors need D %% 128 == 0 and 16-byte aligned tensors");
  unsigned short* dy16 = reinterpret_cast<unsigned short*>(dy_bf16);
  cudaStream_t st = (cudaStream_t)stream;
#define LN_BWD(V, P) MSX_CUDA(msx_launch(add_ln_bwd_kernel<V, P, false>, dim3(ln_wave_grid(add_ln_bwd_kernel<V, P, false>, 0, M)), dim3(kWarps * 32), 0, st, x, y, y_bf16, gamma, mean, rstd, dout, dres, dy, dy16, dgamma, dbeta, dybias, M, D, drop_p, inv_keep, seed, msx_step_counter(), site, accumulate_dres, fuse_xy))
#define LN_BWD_ASYNC(P)                                                                                                   \
  do {                                                                                                                    \
    const int dyn = kWarps * kLnStages * 3 * P * 128 * 4;                                                                 \
    MSX_CUDA(cudaFuncSetAttribute(add_ln_bwd_kernel<4, P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));      \
    MSX_CUDA(msx_launch(add_ln_bwd_kernel<4, P, true>, dim3(ln_wave_grid(add_ln_bwd_kernel<4, P, true>, dyn, M)), dim3(kWarps * 32), dyn, st, x, y, y_bf16, gamma, mean, rstd, dout, dres, dy, dy16,  \
        dgamma, dbeta, dybias, M, D, drop_p, inv_keep, seed, msx_step_counter(), site, accumulate_dres, fuse_xy));         \
  } while (0)
  if (vec) {
    MSX_REQUIRE(D <= 128 * kMaxPer, "msx_add_ln_bwd: D too large");
    const int nper = D / 128;
    // accumulate_dres reads dres in the same pass and fuse_xy aliases x and y: both keep the direct-load kernel
    const bool async_ok = !accumulate_dres && !fuse_xy;
    if (nper <= 1) { if (async_ok) LN_BWD_ASYNC(1); else LN_BWD(4, 1); }
    else if (nper <= 2) { if (async_ok) LN_BWD_ASYNC(2); else LN_BWD(4, 2); }
    else if (nper <= 4) LN_BWD(4, 4); else LN_BWD(4, 8);
  } else {
    MSX_REQUIRE(D <= 32 * kMaxPer, "msx_add_ln_bwd: D=%d unsupported", D);
    const int nper = D / 32;
    if (nper <= 2) LN_BWD(1, 2); else LN_BWD(1, 8);
  }
#undef LN_BWD
#undef LN_BWD_ASYNC
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_add_ln_bwd(const float* x, const float* y, const float* gamma, const float* mean, const float* rstd,
                              const float* dout, float* dres, float* dy, float* dgamma, float* dbeta, float* dybias,
                              long long M, int D, float drop_p, unsigned long long seed, unsigned site, int accumulate_dres, int fuse_xy,
                              void* stream) {
  return msx_add_ln_bwd_ex(x, y, 0, gamma, mean, rstd, dout, dres, dy, nullptr, dgamma, dbeta, dybias, M, D, drop_p, seed, site,
                           accumulate_dres, fuse_xy, stream);
}
