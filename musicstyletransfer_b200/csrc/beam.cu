// A12 (second sampling mode) — one step of batched beam search on the LSTM decoder.
//
// Replaces the per-step host/NDArray code of BeamSearchSampler.sample
// (/root/reference/music_style_transfer/VarAutoEncoder/sampler.py:216-246): expansion scores -log p added to the
// hypothesis scores, finished hypotheses (last token EOS or PAD) frozen, top-k over the beam_size * V candidates of a
// batch row, and the reordering (`take`) of sequences, scores and recurrent states by the winning hypotheses.  The
// reference's loop is LSTM-era and inconsistent at HEAD (it re-takes the PREVIOUS states, adds the kept score twice,
// its inner loop clobbers the step index); this implements the evident intent and the oracle
// (oracle/model.py:beam_search_lstm) states the same rules:
//   * candidate score = score[hyp] + (-log softmax(logits[hyp])[v]); a finished hypothesis has exactly one
//     candidate, (hyp, PAD) at its own score (the reference zeroes its whole expansion row);
//   * at step 1 all beams of a row are identical copies, so only beam 0 is expanded (otherwise the beam fills
//     with duplicates);
//   * the beam_size smallest candidates win, ties broken by the smaller flat index hyp * V + v.
// One CTA per batch row; the selection is beam_size rounds of a block-wide arg-min (beam_size <= 16).
#include "msx_common.cuh"

namespace {

constexpr int kMaxBeam = 16;
constexpr int PAD_ID = 0, EOS_ID = 2;

__global__ void __launch_bounds__(256) beam_step_kernel(const float* __restrict__ logits, int ld, int V, int K,
                                                        const int* __restrict__ seq_in, int* __restrict__ seq_out,
                                                        int seq_ld, int step, const float* __restrict__ score_in,
                                                        float* __restrict__ score_out, int* __restrict__ parent,
                                                        int* __restrict__ next_tok, int* __restrict__ unfinished) {
  extern __shared__ float cand[];                    // [K][V] candidate scores
  __shared__ float lse[kMaxBeam];
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  __shared__ int win_idx[kMaxBeam];
  __shared__ float win_val[kMaxBeam];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // log-sum-exp per hypothesis: warp w handles hypotheses w, w + 8, ...
  for (int k = warp; k < K; k += 8) {
    const float* x = logits + (size_t)(b * K + k) * ld;
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, x[v]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(x[v] - mx);
    s = warp_sum(s);
    if (lane == 0) lse[k] = mx + logf(s);
  }
  __syncthreads();
  for (int i = tid; i < K * V; i += blockDim.x) {
    const int k = i / V, v = i % V;
    const int hyp = b * K + k;
    const int last = seq_in[(size_t)hyp * seq_ld + step - 1];
    float c;
    if (step == 1 && k > 0) c = INFINITY;                                  // identical copies of beam 0
    else if (last == EOS_ID || last == PAD_ID) c = v == PAD_ID ? score_in[hyp] : INFINITY;   // finished: frozen
    else c = score_in[hyp] + (lse[k] - logits[(size_t)hyp * ld + v]);
    cand[i] = c;
  }
  __syncthreads();
  for (int r = 0; r < K; ++r) {
    float best = INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < K * V; i += blockDim.x) {
      const float c = cand[i];
      if (c < best || (c == best && i < bi)) { best = c; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(MSX_FULL, best, o);
      const int oi = __shfl_xor_sync(MSX_FULL, bi, o);
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = best; red_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (red_v[w] < best || (red_v[w] == best && red_i[w] < bi)) { best = red_v[w]; bi = red_i[w]; }
      if (bi == 0x7fffffff) bi = 0;                  // fewer live candidates than beams (all +inf): repeat candidate 0
      win_idx[r] = bi;
      win_val[r] = best;
      cand[bi] = INFINITY;
    }
    __syncthreads();
  }
  // reorder: row r of the new beam continues hypothesis win_idx[r] / V with token win_idx[r] % V
  for (int r = warp; r < K; r += 8) {
    const int src = b * K + win_idx[r] / V, dst = b * K + r, tok = win_idx[r] % V;
    for (int j = lane; j < step; j += 32) seq_out[(size_t)dst * seq_ld + j] = seq_in[(size_t)src * seq_ld + j];
    if (lane == 0) {
      seq_out[(size_t)dst * seq_ld + step] = tok;
      score_out[dst] = isinf(win_val[r]) ? score_in[src] : win_val[r];
      parent[dst] = src;
      next_tok[dst] = tok;
      if (unfinished && tok != EOS_ID && tok != PAD_ID) atomicAdd(unfinished + step, 1);
    }
  }
}

// out[r, :] = in[parent[r], :] for the recurrent states
__global__ void gather_rows_kernel(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ parent,
                                   int rows, int width) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * width) return;
  const int r = (int)(i / width), c = (int)(i % width);
  out[i] = in[(size_t)parent[r] * width + c];
}

}  // namespace

extern "C" int msx_beam_step(const float* logits, int ld, int V, int B, int beam, const int32_t* seq_in, int32_t* seq_out,
                             int seq_ld, int step, const float* score_in, float* score_out, int32_t* parent,
                             int32_t* next_tok, int32_t* unfinished, void* stream) {
  MSX_REQUIRE(logits && seq_in && seq_out && score_in && score_out && parent && next_tok, "msx_beam_step: null pointer");
  MSX_REQUIRE(beam >= 1 && beam <= kMaxBeam, "msx_beam_step: beam size must be in [1, %d]", kMaxBeam);
  MSX_REQUIRE(V > 0 && ld >= V && step >= 1 && step < seq_ld, "msx_beam_step: bad sizes");
  if (B == 0) return MSX_OK;
  const size_t smem = (size_t)beam * V * sizeof(float);
  MSX_REQUIRE(smem <= 200 * 1024, "msx_beam_step: beam * V too large for shared memory");
  MSX_CUDA(cudaFuncSetAttribute(beam_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  beam_step_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(logits, ld, V, beam, seq_in, seq_out, seq_ld, step, score_in,
                                                          score_out, parent, next_tok, unfinished);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_gather_rows(const float* in, float* out, const int32_t* parent, int rows, int width, void* stream) {
  MSX_REQUIRE(in && out && parent, "msx_gather_rows: null pointer");
  if (rows == 0 || width == 0) return MSX_OK;
  const long long n = (long long)rows * width;
  gather_rows_kernel<<<msx_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, parent, rows, width);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
