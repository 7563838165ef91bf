// Per-token row sums: out[v, :] += sum of X[r, :] over the rows r with tokens[r] == v.
//
// This is the backward of "row r of a table indexed by tokens[r]": the embedding gradient (Embedding backward of
// /root/reference/music_style_transfer/VarAutoEncoder/model.py:141,175), and — in the LSTM decoder's table mode — the
// gradient of the [V, 4H] table emb W_i2h^T + b_i2h that feeds the first LSTM layer (engine._lstm_decoder_bwd): the sums of
// d(pre-activations) per token replace the [B*T, 4H] x [4H, H] dgrad, the [4H, B*T] x [B*T, H] wgrad and the embedding
// scatter of the untabled path.
//
// Two steps, both HBM-shaped:
//   msx_token_sort         counting sort of the row indices by token (histogram -> exclusive scan -> scatter; the order of
//                          equal tokens is whatever the scatter's atomics produce: like every atomic reduction of the step
//                          the result is defined up to fp32 summation order);
//   msx_rows_sum_by_token  CTA c walks rows perm[128 c .. 128 c + 127] of the SORTED order: the tokens of a slice are
//                          non-decreasing, so a thread keeps the running sum of its 4 columns in registers and flushes it
//                          (one 16-byte red.global.add) only when the token changes: ~(M / 128 + V) flushes per column
//                          group instead of M atomics; every row is one coalesced D * 4-byte read.
#include "msx_common.cuh"

namespace {

constexpr int kSliceRows = 128;
constexpr int kMaxV = 4096;

__device__ __forceinline__ int clamp_tok(int t, int V) { return min(max(t, 0), V - 1); }

__global__ void __launch_bounds__(256) tok_hist_kernel(const int* __restrict__ tokens, long long M, int V, int* __restrict__ counts) {
  pdl_entry();
  extern __shared__ int bins[];
  for (int i = threadIdx.x; i < V; i += blockDim.x) bins[i] = 0;
  __syncthreads();
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < M; r += (long long)gridDim.x * blockDim.x)
    atomicAdd(&bins[clamp_tok(__ldg(tokens + r), V)], 1);
  __syncthreads();
  for (int i = threadIdx.x; i < V; i += blockDim.x)
    if (bins[i]) atomicAdd(counts + i, bins[i]);
}

// counts[V] -> offsets[V] (exclusive), cursors[V] = offsets; one block
__global__ void __launch_bounds__(1024) tok_scan_kernel(const int* __restrict__ counts, int V, int* __restrict__ offsets,
                                                        int* __restrict__ cursors) {
  pdl_entry();
  __shared__ int part[1024];
  const int per = (V + blockDim.x - 1) / blockDim.x;
  const int lo = threadIdx.x * per, hi = min(V, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += counts[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) { const int t = part[i]; part[i] = run; run += t; }
  }
  __syncthreads();
  int run = part[threadIdx.x];
  for (int i = lo; i < hi; ++i) { offsets[i] = run; cursors[i] = run; run += counts[i]; }
}

// Each block sorts a contiguous chunk of rows locally (shared-memory histogram), reserves its range of every token's
// segment with ONE global atomic per token present, then places the rows with shared-memory cursors: the synthetic 4/4 rows
// (and real melodies) concentrate on a few dozen tokens, per-row global atomics on their cursors would serialise.
constexpr int kScatterChunk = 2048;
__global__ void __launch_bounds__(256) tok_scatter_kernel(const int* __restrict__ tokens, long long M, int V, int* __restrict__ cursors,
                                                          int* __restrict__ perm, int* __restrict__ sorted_tok) {
  pdl_entry();
  extern __shared__ int sm[];                               // bins[V] | base[V]
  int* bins = sm;
  int* base = sm + V;
  const long long r0 = (long long)blockIdx.x * kScatterChunk;
  const int n = (int)min((long long)kScatterChunk, M - r0);
  for (int i = threadIdx.x; i < V; i += blockDim.x) bins[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&bins[clamp_tok(__ldg(tokens + r0 + i), V)], 1);
  __syncthreads();
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    base[i] = bins[i] ? atomicAdd(cursors + i, bins[i]) : 0;
    bins[i] = 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int v = clamp_tok(__ldg(tokens + r0 + i), V);
    const int pos = base[v] + atomicAdd(&bins[v], 1);
    perm[pos] = (int)(r0 + i);
    sorted_tok[pos] = v;
  }
}

// blockDim.x = D / 4 threads, each owning 4 consecutive columns
__global__ void __launch_bounds__(256) rows_sum_by_token_kernel(const float* __restrict__ X, int ld, int D, const int* __restrict__ perm,
                                                                const int* __restrict__ sorted_tok, long long M, float scale,
                                                                float* __restrict__ out) {
  pdl_entry();
  __shared__ int s_row[kSliceRows], s_tok[kSliceRows];
  const long long i0 = (long long)blockIdx.x * kSliceRows;
  const int n = (int)min((long long)kSliceRows, M - i0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s_row[i] = __ldg(perm + i0 + i);
    s_tok[i] = __ldg(sorted_tok + i0 + i);
  }
  __syncthreads();
  const int c = threadIdx.x * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur = s_tok[0];
  auto flush = [&](int v) {
    atomicAdd(reinterpret_cast<float4*>(out + (size_t)v * D + c),                 // red.global.add.v4.f32
              make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale));
    acc = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  constexpr int U = 8;                                      // independent row loads in flight per thread
  for (int i = 0; i < n; i += U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u < n) v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)s_row[i + u] * ld + c));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u >= n) break;
      const int t = s_tok[i + u];
      if (t != cur) {                                       // block-uniform branch: the slice's tokens are the same for every thread
        flush(cur);
        cur = t;
      }
      acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
    }
  }
  flush(cur);
}

}  // namespace

// workspace: int32 [3 * V] (counts | offsets | cursors).  perm / sorted_tok: int32 [M].
extern "C" int msx_token_sort(const int32_t* tokens, long long M, int V, int32_t* perm, int32_t* sorted_tok, int32_t* workspace,
                              void* stream) {
  MSX_REQUIRE(tokens && perm && sorted_tok && workspace, "msx_token_sort: null pointer");
  MSX_REQUIRE(V >= 1 && V <= kMaxV && M >= 0 && M < (1ll << 31), "msx_token_sort: 1 <= V <= 4096, M < 2^31");
  if (M == 0) return MSX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  MSX_CUDA(cudaMemsetAsync(workspace, 0, (size_t)V * sizeof(int), st));
  const long long want = (M + 255) / 256;
  const int grid = (int)(want < 2ll * msx_num_sms() ? want : 2ll * msx_num_sms());
  MSX_CUDA(msx_launch(tok_hist_kernel, dim3(grid), dim3(256), (size_t)V * sizeof(int), st, tokens, M, V, workspace));
  MSX_CUDA(msx_launch(tok_scan_kernel, dim3(1), dim3(1024), 0, st, workspace, V, workspace + V, workspace + 2 * V));
  MSX_CUDA(msx_launch(tok_scatter_kernel, dim3((int)((M + kScatterChunk - 1) / kScatterChunk)), dim3(256), 2 * (size_t)V * sizeof(int), st, 
      tokens, M, V, workspace + 2 * V, perm, sorted_tok));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

// out [V, D] += scale * per-token sums of the rows of X [M, ld] (perm / sorted_tok from msx_token_sort).
extern "C" int msx_rows_sum_by_token(const float* X, int ld, int D, const int32_t* perm, const int32_t* sorted_tok, long long M,
                                     float scale, float* out, void* stream) {
  MSX_REQUIRE(X && perm && sorted_tok && out, "msx_rows_sum_by_token: null pointer");
  MSX_REQUIRE(D >= 4 && D <= 1024 && (D & 3) == 0 && (ld & 3) == 0 && (((uintptr_t)X | (uintptr_t)out) & 15) == 0,
              "msx_rows_sum_by_token: D %% 4 == 0, D <= 1024, 16-byte aligned rows");
  if (M == 0) return MSX_OK;
  const int grid = (int)((M + kSliceRows - 1) / kSliceRows);
  MSX_CUDA(msx_launch(rows_sum_by_token_kernel, dim3(grid), dim3(D / 4), 0, (cudaStream_t)stream, X, ld, D, perm, sorted_tok, M, scale, out));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
