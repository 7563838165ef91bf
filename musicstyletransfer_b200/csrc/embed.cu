// K2a — fused embedding front-end: gather + class embedding + sqrt(D) scale + positional encoding + pad mask.
//
// Replaces (reference, /root/reference/music_style_transfer/VarAutoEncoder):
//   model.py:81-91      Encoder: mask = tokens != 0; tok_emb[tokens] + class_emb[classes][:,None,:]
//   transformer.py:270  TransformerEncoder: sqrt(D) * x + pos_embeddings[:T]
//   model.py:241-247    Decoder: concat(initial_state, emb[tokens]); SequenceMask(len+1)
//   transformer.py:237  TransformerDecoder: sqrt(D) * x + pos_embeddings[:T+1]
//   model.py:176        LSTMDecoder: emb[tokens]                       (scale 1, no PE, no class term)
// One warp per output row; rows are D floats, float4 when D % 4 == 0.  Backward scatters with
// red.global.add (embedding rows collide by construction) and reduces the class term per CTA.
#include "msx_common.cuh"

namespace {

// out[b, pos, :] = scale * ((pos < prefix ? prefix_vec[b] : tok_emb[tokens[b, pos-prefix]]) + cls_emb[classes[b]]) + pe[pos]
__global__ void __launch_bounds__(256) embed_fwd_kernel(const int* __restrict__ tokens, const int* __restrict__ classes,
                                                        const int* __restrict__ seq_lens,
                                                        const float* __restrict__ tok_emb,
                                                        const float* __restrict__ cls_emb,
                                                        const float* __restrict__ prefix_vec,
                                                        const float* __restrict__ pe, float* __restrict__ out,
                                                        unsigned short* __restrict__ out16,
                                                        float* __restrict__ mask, int B, int T, int D, int prefix,
                                                        float scale, int vocab) {
  const int TP = T + prefix;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)B * TP) return;
  const int b = (int)(row / TP), pos = (int)(row % TP);
  const float* src;
  int tok = 1;
  if (pos < prefix) {
    src = prefix_vec + (size_t)b * D;
  } else {
    tok = __ldg(tokens + (size_t)b * T + pos - prefix);
    tok = min(max(tok, 0), vocab - 1);
    src = tok_emb + (size_t)tok * D;
  }
  const float* cls = cls_emb ? cls_emb + (size_t)__ldg(classes + b) * D : nullptr;
  const float* pr = pe ? pe + (size_t)pos * D : nullptr;
  float* dst = out ? out + (size_t)row * D : nullptr;
  unsigned short* dst16 = out16 ? out16 + (size_t)row * D : nullptr;
  for (int e = lane; e < D; e += 32) {
    float v = __ldg(src + e);
    if (cls) v += __ldg(cls + e);
    v *= scale;
    if (pr) v += __ldg(pr + e);
    if (dst) dst[e] = v;
    if (dst16) {                            // bf16 copy (round to nearest even): operand of the bf16 GEMM variant
      unsigned r;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(v));
      dst16[e] = (unsigned short)(r & 0xFFFFu);
    }
  }
  if (mask && lane == 0) {
    // encoder: tokens != 0 (model.py:81-83); decoder: position < seq_len + 1 (model.py:246-247)
    mask[row] = seq_lens ? (pos < __ldg(seq_lens + b) + prefix ? 1.f : 0.f) : (tok != 0 ? 1.f : 0.f);
  }
}

// d tok_emb[tok] += scale*dout ; d cls_emb[cls[b]] += scale*sum_t dout[b,t] ; d prefix_vec[b] = scale*dout[b,0]
__global__ void __launch_bounds__(256) embed_bwd_kernel(const int* __restrict__ tokens, const int* __restrict__ classes,
                                                        const float* __restrict__ dout, float* __restrict__ d_tok_emb,
                                                        float* __restrict__ d_cls_emb, float* __restrict__ d_prefix,
                                                        int B, int T, int D, int prefix, float scale, int vocab) {
  const int b = blockIdx.x;
  const int TP = T + prefix;
  const int c = d_cls_emb ? __ldg(classes + b) : 0;
  for (int e = threadIdx.x; e < D; e += blockDim.x) {
    float csum = 0.f;
    for (int pos = 0; pos < TP; ++pos) {
      const float g = scale * __ldg(dout + ((size_t)b * TP + pos) * D + e);
      csum += g;
      if (pos < prefix) {
        d_prefix[(size_t)b * D + e] = g;
      } else {
        int tok = __ldg(tokens + (size_t)b * T + pos - prefix);
        tok = min(max(tok, 0), vocab - 1);
        atomicAdd(d_tok_emb + (size_t)tok * D + e, g);
      }
    }
    if (d_cls_emb) atomicAdd(d_cls_emb + (size_t)c * D + e, csum);
  }
}

// Backward without per-element atomics (prefix == 0, at most two classes).  The scatter d_tok_emb[tok] += dout[row] has few
// distinct targets (V = 293 rows, and on 4/4 material a third of all positions hit the four TIMESHIFT rows), so
// red.global.add from every position serialises on a handful of L2 lines (88-150 us for a 136 MB read).  Here a CTA owns
// ONE vocabulary row and one segment of the positions: it scans the segment's token ids (L2-resident), lists the
// matching positions in shared memory, sums their dout rows in registers (coalesced row reads, every dout row is read
// exactly once overall) and issues one atomic per column at the end.  The class-embedding gradient rides along: the sum
// is kept per class of the position's sequence and added to d_cls_emb as well.
constexpr int kGatherThreads = 256, kGatherWarps = 8, kGatherChunk = 2048, kGatherMaxD = 256;

__global__ void __launch_bounds__(kGatherThreads) embed_bwd_gather_kernel(
    const int* __restrict__ tokens, const int* __restrict__ classes, const float* __restrict__ dout,
    float* __restrict__ d_tok_emb, float* __restrict__ d_cls_emb, long long rows, int T, int D, float scale, int vocab,
    int segments) {
  __shared__ int list[kGatherChunk];
  __shared__ int cnt;
  __shared__ __align__(16) float red[kGatherWarps][kGatherMaxD];
  const int v = blockIdx.x / segments, sgm = blockIdx.x % segments;
  const long long seg = (rows + segments - 1) / segments;
  const long long r0 = (long long)sgm * seg, r1 = min(rows, r0 + seg);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // a warp sums whole rows: lane owns the float4 groups lane and lane + 32 (D <= 256), one partial sum per class
  float4 acc[2][2];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[c][j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int nv4 = D >> 2;
  bool any = false;
  for (long long base = r0; base < r1; base += kGatherChunk) {
    if (tid == 0) cnt = 0;
    __syncthreads();
    const long long end = min(r1, base + (long long)kGatherChunk);
    for (long long i = base + tid; i < end; i += kGatherThreads) {
      int tok = __ldg(tokens + i);
      tok = min(max(tok, 0), vocab - 1);
      if (tok == v) {
        // position within the chunk and, when the class gradient is wanted, the class of its sequence in bit 30
        const int cls = d_cls_emb ? (__ldg(classes + i / T) & 1) : 0;
        list[atomicAdd(&cnt, 1)] = (int)(i - base) | (cls << 30);
      }
    }
    __syncthreads();
    const int n = cnt;
    any |= n > 0;
#pragma unroll 2
    for (int k = warp; k < n; k += kGatherWarps) {
      const int e = list[k];
      const float4* src = reinterpret_cast<const float4*>(dout + (size_t)(base + (e & 0x3FFFFFFF)) * D);
      const bool c1 = (e >> 30) != 0;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c4 = lane + 32 * j;
        if (c4 < nv4) {
          const float4 g = __ldg(src + c4);
          acc[0][j].x += c1 ? 0.f : g.x; acc[0][j].y += c1 ? 0.f : g.y; acc[0][j].z += c1 ? 0.f : g.z; acc[0][j].w += c1 ? 0.f : g.w;
          acc[1][j].x += c1 ? g.x : 0.f; acc[1][j].y += c1 ? g.y : 0.f; acc[1][j].z += c1 ? g.z : 0.f; acc[1][j].w += c1 ? g.w : 0.f;
        }
      }
    }
    __syncthreads();
  }
  if (!any) return;                          // block-uniform
  for (int c = 0; c < (d_cls_emb ? 2 : 1); ++c) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < nv4) {
        // without a class table everything sits in acc[0]; with one, class c's sum goes out on round c
        *reinterpret_cast<float4*>(&red[warp][4 * c4]) = acc[c][j];
      }
    }
    __syncthreads();
    if (tid < D) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kGatherWarps; ++w) t += red[w][tid];
      t *= scale;
      if (t != 0.f) {
        atomicAdd(d_tok_emb + (size_t)v * D + tid, t);
        if (d_cls_emb) atomicAdd(d_cls_emb + (size_t)c * D + tid, t);
      }
    }
  }
}

}  // namespace

extern "C" int msx_embed_fwd_ex(const int32_t* tokens, const int32_t* classes, const int32_t* seq_lens,
                                const float* tok_emb, const float* cls_emb, const float* prefix_vec, const float* pe,
                                float* out, void* out_bf16, float* mask, int B, int T, int D, int prefix, float scale,
                                int vocab, void* stream);
extern "C" int msx_embed_fwd(const int32_t* tokens, const int32_t* classes, const int32_t* seq_lens,
                             const float* tok_emb, const float* cls_emb, const float* prefix_vec, const float* pe,
                             float* out, float* mask, int B, int T, int D, int prefix, float scale, int vocab,
                             void* stream) {
  MSX_REQUIRE(out, "msx_embed_fwd: null pointer");
  return msx_embed_fwd_ex(tokens, classes, seq_lens, tok_emb, cls_emb, prefix_vec, pe, out, nullptr, mask, B, T, D, prefix,
                          scale, vocab, stream);
}

extern "C" int msx_embed_fwd_ex(const int32_t* tokens, const int32_t* classes, const int32_t* seq_lens,
                                const float* tok_emb, const float* cls_emb, const float* prefix_vec, const float* pe,
                                float* out, void* out_bf16, float* mask, int B, int T, int D, int prefix, float scale,
                                int vocab, void* stream) {
  MSX_REQUIRE(tokens && tok_emb && (out || out_bf16), "msx_embed_fwd: null pointer");
  MSX_REQUIRE(prefix == 0 || prefix_vec, "msx_embed_fwd: prefix rows need prefix_vec");
  MSX_REQUIRE(!cls_emb || classes, "msx_embed_fwd: class embedding needs classes");
  if (B == 0 || T + prefix == 0) return MSX_OK;
  const long long rows = (long long)B * (T + prefix);
  const int wpb = 8;
  embed_fwd_kernel<<<msx_ceil_div(rows, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      tokens, classes, seq_lens, tok_emb, cls_emb, prefix_vec, pe, out, reinterpret_cast<unsigned short*>(out_bf16), mask, B, T,
      D, prefix, scale, vocab);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_embed_bwd_ex(const int32_t* tokens, const int32_t* classes, const float* dout, float* d_tok_emb,
                                float* d_cls_emb, float* d_prefix, int B, int T, int D, int prefix, float scale, int vocab,
                                int num_classes, void* stream);
extern "C" int msx_embed_bwd(const int32_t* tokens, const int32_t* classes, const float* dout, float* d_tok_emb,
                             float* d_cls_emb, float* d_prefix, int B, int T, int D, int prefix, float scale, int vocab,
                             void* stream) {
  return msx_embed_bwd_ex(tokens, classes, dout, d_tok_emb, d_cls_emb, d_prefix, B, T, D, prefix, scale, vocab, 0, stream);
}

// num_classes: rows of the class table (0 = unknown); the gather kernel keeps one partial sum per class and takes <= 2.
extern "C" int msx_embed_bwd_ex(const int32_t* tokens, const int32_t* classes, const float* dout, float* d_tok_emb,
                                float* d_cls_emb, float* d_prefix, int B, int T, int D, int prefix, float scale, int vocab,
                                int num_classes, void* stream) {
  MSX_REQUIRE(tokens && dout && d_tok_emb, "msx_embed_bwd: null pointer");
  MSX_REQUIRE(prefix == 0 || d_prefix, "msx_embed_bwd: prefix rows need d_prefix");
  MSX_REQUIRE(!d_cls_emb || classes, "msx_embed_bwd: class embedding needs classes");
  if (B == 0) return MSX_OK;
  const long long rows = (long long)B * T;
  // large problems without prefix rows and with at most two classes: gather kernel (one CTA per vocabulary row and
  // position segment, no per-element atomics); the class term needs class ids in {0, 1}, which the caller guarantees
  // through num_classes <= 2 (msx_embed_bwd_ex) — other shapes keep the scatter kernel below
  if (prefix == 0 && rows >= 8192 && D <= kGatherMaxD && (D & 3) == 0 && (((uintptr_t)dout) & 15) == 0 && (!d_cls_emb || (num_classes >= 1 && num_classes <= 2))) {
    int segments = (int)((rows + 4095) / 4096);
    if (segments > 64) segments = 64;
    if (segments < 1) segments = 1;
    embed_bwd_gather_kernel<<<vocab * segments, kGatherThreads, 0, (cudaStream_t)stream>>>(
        tokens, classes, dout, d_tok_emb, d_cls_emb, rows, T, D, scale, vocab, segments);
    MSX_LAUNCH_CHECK();
    return MSX_OK;
  }
  // (a variant that accumulated a 64-column slice of the table with shared-memory atomics was 6x SLOWER: float
  // atomicAdd on shared memory is a CAS loop and the synthetic 4/4 rows hit few distinct tokens)
  embed_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(tokens, classes, dout, d_tok_emb, d_cls_emb, d_prefix, B, T, D,
                                                        prefix, scale, vocab);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
