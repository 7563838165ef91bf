// K2 (exact-fp32 path) — register-tiled FFMA GEMM with the fused epilogues the VarAutoEncoder step needs.
//
// Replaces the MXNet operators behind every gluon.nn.Dense on the hot path
// (/root/reference/music_style_transfer/VarAutoEncoder/transformer.py:36-40,65-68,88-93,104;
//  model.py:70-71,139-157,214-227) and their autograd backward (trainer.py:176):
//   forward  Y  = X W^T + b            transA=0 transB=1 (+ReLU, +dropout)
//   dgrad    dX = dY W                 transA=0 transB=0 (+ReLU'/dropout mask from the saved activation, +accumulate)
//   wgrad    dW += dY^T X, db += 1^T dY transA=1 transB=0, split-K with red.global.add, bias grad folded in
//
// 128x128x8 tiles, 256 threads, 8x8 outputs per thread laid out as four 4x4 blocks so that every
// shared-memory read is a conflict-free LDS.128; global->register->shared double buffering.
// The tcgen05/TMA tensor-core path lives in gemm_tc.cu; this kernel is the fp32-exact mode and the
// fallback for shapes the tensor path does not take (tiny M, unaligned leading dimensions).
#include "msx_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 8, NT = 256;

struct GemmParams {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  int lda, ldb, ldc;
  const float* bias;     // [N] added to every row (forward)
  int relu;              // max(x,0) after bias
  float drop_p;          // dropout after relu (0 = off)
  float drop_inv_keep;
  unsigned long long seed;
  const unsigned long long* seed_ctr;
  unsigned site;
  const float* aux;      // dgrad: multiply by (aux[m,n] > 0 ? aux_scale : 0)
  int ldaux;
  float aux_scale;
  int accumulate;        // C += result (plain read-modify-write; no split-K)
  int splitk;            // >1: grid.z slices of K, results added with atomics into a pre-zeroed C
  float* colsum;         // wgrad: colsum[m] += sum_k opA[m,k] (bias gradient), done by the n-tile-0 CTAs
  int vecA, vecB, vecC;  // 16-byte vector access allowed
};

template <int TA, int TB>
__global__ void __launch_bounds__(NT) sgemm_kernel(const GemmParams p) {
  pdl_entry();
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kchunk = ((p.K + p.splitk - 1) / p.splitk + BK - 1) / BK * BK;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(p.K, kbeg + kchunk);
  if (kbeg >= kend) return;

  // ---- global -> register staging (one float4 of A and one of B per thread per k-tile)
  float4 ra, rb;
  auto load_tiles = [&](int k0) {
    if (TA == 0) {  // A[m][k], k contiguous: thread -> row tid/2, k offset (tid&1)*4
      const int r = m0 + (tid >> 1), kk = k0 + (tid & 1) * 4;
      const float* src = p.A + (size_t)r * p.lda + kk;
      if (r < p.M && kk + 3 < kend && p.vecA) {
        ra = __ldg(reinterpret_cast<const float4*>(src));
      } else {
        float t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = (r < p.M && kk + i < kend) ? __ldg(src + i) : 0.f;
        ra = make_float4(t[0], t[1], t[2], t[3]);
      }
    } else {        // A stored [k][m], m contiguous: thread -> k tid/32, m offset (tid&31)*4
      const int kk = k0 + (tid >> 5), r = m0 + (tid & 31) * 4;
      const float* src = p.A + (size_t)kk * p.lda + r;
      if (kk < kend && r + 3 < p.M && p.vecA) {
        ra = __ldg(reinterpret_cast<const float4*>(src));
      } else {
        float t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = (kk < kend && r + i < p.M) ? __ldg(src + i) : 0.f;
        ra = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
    if (TB == 1) {  // B stored [n][k], k contiguous
      const int c = n0 + (tid >> 1), kk = k0 + (tid & 1) * 4;
      const float* src = p.B + (size_t)c * p.ldb + kk;
      if (c < p.N && kk + 3 < kend && p.vecB) {
        rb = __ldg(reinterpret_cast<const float4*>(src));
      } else {
        float t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = (c < p.N && kk + i < kend) ? __ldg(src + i) : 0.f;
        rb = make_float4(t[0], t[1], t[2], t[3]);
      }
    } else {        // B stored [k][n], n contiguous
      const int kk = k0 + (tid >> 5), c = n0 + (tid & 31) * 4;
      const float* src = p.B + (size_t)kk * p.ldb + c;
      if (kk < kend && c + 3 < p.N && p.vecB) {
        rb = __ldg(reinterpret_cast<const float4*>(src));
      } else {
        float t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = (kk < kend && c + i < p.N) ? __ldg(src + i) : 0.f;
        rb = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
  };
  auto store_tiles = [&](int buf) {
    if (TA == 0) {
      const int r = tid >> 1, kk = (tid & 1) * 4;
      As[buf][kk + 0][r] = ra.x; As[buf][kk + 1][r] = ra.y; As[buf][kk + 2][r] = ra.z; As[buf][kk + 3][r] = ra.w;
    } else {
      *reinterpret_cast<float4*>(&As[buf][tid >> 5][(tid & 31) * 4]) = ra;
    }
    if (TB == 1) {
      const int c = tid >> 1, kk = (tid & 1) * 4;
      Bs[buf][kk + 0][c] = rb.x; Bs[buf][kk + 1][c] = rb.y; Bs[buf][kk + 2][c] = rb.z; Bs[buf][kk + 3][c] = rb.w;
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][tid >> 5][(tid & 31) * 4]) = rb;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float csum = 0.f;
  const bool do_colsum = p.colsum != nullptr && blockIdx.x == 0 && tid < BM;

  load_tiles(kbeg);
  store_tiles(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool has_next = k0 + BK < kend;
    if (has_next) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (do_colsum) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) csum += As[buf][kk][tid];
    }
    if (has_next) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
  if (do_colsum && m0 + tid < p.M) atomicAdd(p.colsum + m0 + tid, csum);

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= p.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int c = n0 + jh * 64 + tx * 4;
      if (c >= p.N) continue;
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += (c + j < p.N) ? __ldg(p.bias + c + j) : 0.f;
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.drop_p > 0.f) {
        float s[4];
        dropout_scale4(msx_eff_seed(p.seed, p.seed_ctr), p.site, ((unsigned long long)r * p.N + c) >> 2, p.drop_p, p.drop_inv_keep, s);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] *= s[j];
      }
      if (p.aux) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[j] *= (c + j < p.N && __ldg(p.aux + (size_t)r * p.ldaux + c + j) > 0.f) ? p.aux_scale : 0.f;
      }
      float* dst = p.C + (size_t)r * p.ldc + c;
      if (p.splitk > 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < p.N) atomicAdd(dst + j, v[j]);
      } else if (p.vecC && c + 3 < p.N) {
        if (p.accumulate) {
          const float4 o = *reinterpret_cast<const float4*>(dst);
          v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
        }
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < p.N) dst[j] = p.accumulate ? dst[j] + v[j] : v[j];
      }
    }
  }
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

// C[M,N] = epilogue(opA(A)[M,K] * opB(B)[K,N]);  see include/msx.h for the argument contract.
extern "C" int msx_gemm_f32(const float* A, int lda, int transA, const float* B, int ldb, int transB, float* C, int ldc,
                            int M, int N, int K, const float* bias, int relu, float drop_p, unsigned long long seed,
                            unsigned site, const float* aux, int ldaux, float aux_scale, int accumulate, int splitk,
                            float* colsum, void* stream) {
  MSX_REQUIRE(M >= 0 && N >= 0 && K >= 0, "msx_gemm_f32: negative dimension");
  if (M == 0 || N == 0) return MSX_OK;
  MSX_REQUIRE(A && B && C, "msx_gemm_f32: null operand");
  MSX_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "msx_gemm_f32: dropout probability must be in [0,1)");
  MSX_REQUIRE(!(splitk > 1 && (bias || relu || drop_p > 0.f || aux || accumulate)),
              "msx_gemm_f32: split-K only supports the plain atomic-add epilogue");
  MSX_REQUIRE(colsum == nullptr || transA == 1, "msx_gemm_f32: colsum (bias gradient) needs transA=1");
  if (K == 0) {
    MSX_REQUIRE(!accumulate && splitk <= 1, "msx_gemm_f32: K == 0 with accumulate");
  }
  GemmParams p;
  p.A = A; p.B = B; p.C = C; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.bias = bias; p.relu = relu; p.drop_p = drop_p; p.drop_inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  p.seed = seed; p.seed_ctr = msx_step_counter(); p.site = site; p.aux = aux; p.ldaux = ldaux; p.aux_scale = aux_scale; p.accumulate = accumulate;
  p.splitk = splitk < 1 ? 1 : splitk;
  p.colsum = colsum;
  p.vecA = aligned16(A) && (lda % 4 == 0);
  p.vecB = aligned16(B) && (ldb % 4 == 0);
  p.vecC = aligned16(C) && (ldc % 4 == 0);
  dim3 grid(msx_ceil_div(N, BN), msx_ceil_div(M, BM), p.splitk);
  MSX_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "msx_gemm_f32: grid too large (M=%d)", M);
  cudaStream_t st = (cudaStream_t)stream;
  if (transA == 0 && transB == 1) MSX_CUDA(msx_launch(sgemm_kernel<0, 1>, dim3(grid), dim3(NT), 0, st, p));
  else if (transA == 0 && transB == 0) MSX_CUDA(msx_launch(sgemm_kernel<0, 0>, dim3(grid), dim3(NT), 0, st, p));
  else if (transA == 1 && transB == 0) MSX_CUDA(msx_launch(sgemm_kernel<1, 0>, dim3(grid), dim3(NT), 0, st, p));
  else MSX_CUDA(msx_launch(sgemm_kernel<1, 1>, dim3(grid), dim3(NT), 0, st, p));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

// ---------------------------------------------------------------------------------------------
// Bias gradient for the tensor-core path: out[n] += sum_m X[m, n]   (db of a Dense layer, trainer.py:176)
namespace {
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int ld, long long M, int N,
                                                     float* __restrict__ out) {
  pdl_entry();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (c < N)
    for (long long r = (long long)blockIdx.y * 8 + ty; r < M; r += (long long)gridDim.y * 8) acc += X[r * ld + c];
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(out + c, t);
  }
}
}  // namespace

extern "C" int msx_colsum(const float* X, int ld, long long M, int N, float* out, void* stream) {
  MSX_REQUIRE(X && out, "msx_colsum: null pointer");
  if (M == 0 || N == 0) return MSX_OK;
  const int gx = msx_ceil_div(N, 32);
  int gy = (int)min((long long)msx_ceil_div(msx_num_sms() * 8, gx), (M + 63) / 64);
  if (gy < 1) gy = 1;
  MSX_CUDA(msx_launch(colsum_kernel, dim3(dim3(gx, gy)), dim3(256), 0, (cudaStream_t)stream, X, ld, M, N, out));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
