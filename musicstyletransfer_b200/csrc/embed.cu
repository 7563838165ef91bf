// K2a — fused embedding front-end: gather + class embedding + sqrt(D) scale + positional encoding + pad mask.
//
// Replaces (reference, /root/reference/music_style_transfer/VarAutoEncoder):
//   model.py:81-91      Encoder: mask = tokens != 0; tok_emb[tokens] + class_emb[classes][:,None,:]
//   transformer.py:270  TransformerEncoder: sqrt(D) * x + pos_embeddings[:T]
//   model.py:241-247    Decoder: concat(initial_state, emb[tokens]); SequenceMask(len+1)
//   transformer.py:237  TransformerDecoder: sqrt(D) * x + pos_embeddings[:T+1]
//   model.py:176        LSTMDecoder: emb[tokens]                       (scale 1, no PE, no class term)
// One warp per output row; rows are D floats.  Backward: large problems sort the positions of a 512-position segment
// by token id in shared memory and sum whole rows per token in registers (embed_bwd_sorted_kernel); small problems and
// prefix rows scatter with red.global.add (embedding rows collide by construction) and reduce the class term per CTA.
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
                                                        unsigned short* __restrict__ out16lo,
                                                        float* __restrict__ mask, int B, int T, int D, int prefix,
                                                        float scale, int vocab, int vec) {
  pdl_entry();
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
  if (vec) {                                // vector path (host: D % 4 == 0 and 16-byte aligned tables / outputs)
    for (int e = lane * 4; e < D; e += 128) {
      float4 v = __ldg(reinterpret_cast<const float4*>(src + e));
      if (cls) {
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(cls + e));
        v.x += c4.x; v.y += c4.y; v.z += c4.z; v.w += c4.w;
      }
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      if (pr) {
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(pr + e));
        v.x += p4.x; v.y += p4.y; v.z += p4.z; v.w += p4.w;
      }
      if (dst) *reinterpret_cast<float4*>(dst + e) = v;
      if (out16lo) {
        uint2 hi, lo;
        split4_bf16(v.x, v.y, v.z, v.w, hi, lo);
        *reinterpret_cast<uint2*>(dst16 + e) = hi;
        *reinterpret_cast<uint2*>(out16lo + (size_t)row * D + e) = lo;
      } else if (dst16) {
        uint2 hi;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi.x) : "f"(v.y), "f"(v.x));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi.y) : "f"(v.w), "f"(v.z));
        *reinterpret_cast<uint2*>(dst16 + e) = hi;
      }
    }
  } else
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
      if (out16lo) {                        // lo plane of the p3 GEMM operand: rn_bf16(v - hi)
        unsigned l;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(0.f), "f"(v - __uint_as_float(r << 16)));
        out16lo[(size_t)row * D + e] = (unsigned short)(l & 0xFFFFu);
      }
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
  pdl_entry();
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

// Backward without per-element atomics (prefix == 0, at most two classes, vocab <= 512, D <= 256).  The scatter
// d_tok_emb[tok] += dout[row] has few distinct targets (V = 293 rows, and on 4/4 material a third of all positions hit the
// four TIMESHIFT rows), so red.global.add from every position serialises on a handful of L2 lines (80-150 us for a
// 136 MB read).  Here a CTA takes a segment of 512 positions, counting-sorts them by token id in shared memory
// (histogram with integer atomics, warp-scan prefix, scatter), and its warps then walk the vocabulary: a warp sums the
// dout rows of one token in registers (coalesced float4 row reads, every row read exactly once overall) and issues one
// 16-byte vector atomic per lane for that token.  The class-embedding gradient is the per-class total of the same rows:
// it accumulates in registers across all tokens of the warp and leaves once per CTA.
constexpr int kSortThreads = 256, kSortWarps = 8, kSortSeg = 512, kSortMaxV = 512, kSortMaxD = 256;

__global__ void __launch_bounds__(kSortThreads) embed_bwd_sorted_kernel(
    const int* __restrict__ tokens, const int* __restrict__ classes, const float* __restrict__ dout,
    float* __restrict__ d_tok_emb, float* __restrict__ d_cls_emb, long long rows, int T, int D, float scale, int vocab) {
  pdl_entry();
  __shared__ int cnt[kSortMaxV];
  __shared__ int start[kSortMaxV];
  __shared__ int sorted[kSortSeg];
  __shared__ int wsum[kSortWarps];
  __shared__ __align__(16) float red[kSortWarps][kSortMaxD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long r0 = (long long)blockIdx.x * kSortSeg;
  const int n = (int)min((long long)kSortSeg, rows - r0);
  for (int i = tid; i < kSortMaxV; i += kSortThreads) cnt[i] = 0;
  __syncthreads();
  // histogram: every position takes a rank within its token's bucket
  int tok[2], rank[2], cls[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = tid + j * kSortThreads;
    tok[j] = -1;
    if (i < n) {
      int t = __ldg(tokens + r0 + i);
      t = min(max(t, 0), vocab - 1);
      tok[j] = t;
      rank[j] = atomicAdd(&cnt[t], 1);
      cls[j] = d_cls_emb ? (__ldg(classes + (r0 + i) / T) & 1) : 0;
    }
  }
  __syncthreads();
  // exclusive prefix sum of the 512 bucket sizes: 2 per thread, warp scan, then the warp totals
  {
    const int a = cnt[2 * tid], b2 = cnt[2 * tid + 1];
    int v = a + b2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(MSX_FULL, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    const int excl = base + v - (a + b2);
    start[2 * tid] = excl;
    start[2 * tid + 1] = excl + a;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 2; ++j)
    if (tok[j] >= 0) sorted[start[tok[j]] + rank[j]] = (tid + j * kSortThreads) | (cls[j] << 30);
  __syncthreads();
  // warps walk the vocabulary; lane owns the float4 groups lane and lane + 32 of a row
  const int nv4 = D >> 2;
  float4 cacc[2][2];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int j = 0; j < 2; ++j) cacc[c][j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int v = warp; v < vocab; v += kSortWarps) {
    const int m = cnt[v];
    if (m == 0) continue;                                  // warp-uniform
    const int s0 = start[v];
    float4 acc[2][2];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int j = 0; j < 2; ++j) acc[c][j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < m; ++k) {
      const int e = sorted[s0 + k];
      const float4* src = reinterpret_cast<const float4*>(dout + (size_t)(r0 + (e & 0x3FFFFFFF)) * D);
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
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < nv4) {
        const float4 t = make_float4(scale * (acc[0][j].x + acc[1][j].x), scale * (acc[0][j].y + acc[1][j].y),
                                     scale * (acc[0][j].z + acc[1][j].z), scale * (acc[0][j].w + acc[1][j].w));
        atomicAdd(reinterpret_cast<float4*>(d_tok_emb + (size_t)v * D) + c4, t);      // red.global.add.v4.f32
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          cacc[c][j].x += acc[c][j].x; cacc[c][j].y += acc[c][j].y; cacc[c][j].z += acc[c][j].z; cacc[c][j].w += acc[c][j].w;
        }
      }
    }
  }
  if (!d_cls_emb) return;
  for (int c = 0; c < 2; ++c) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < nv4) *reinterpret_cast<float4*>(&red[warp][4 * c4]) = cacc[c][j];
    }
    __syncthreads();
    if (tid < nv4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        const float4 r4 = *reinterpret_cast<const float4*>(&red[w][4 * tid]);
        t.x += r4.x; t.y += r4.y; t.z += r4.z; t.w += r4.w;
      }
      t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale;
      if (t.x != 0.f || t.y != 0.f || t.z != 0.f || t.w != 0.f)
        atomicAdd(reinterpret_cast<float4*>(d_cls_emb + (size_t)c * D) + tid, t);
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

extern "C" int msx_embed_fwd_p(const int32_t* tokens, const int32_t* classes, const int32_t* seq_lens,
                               const float* tok_emb, const float* cls_emb, const float* prefix_vec, const float* pe,
                               float* out, void* out_bf16, void* out_bf16_lo, float* mask, int B, int T, int D, int prefix,
                               float scale, int vocab, void* stream);
extern "C" int msx_embed_fwd_ex(const int32_t* tokens, const int32_t* classes, const int32_t* seq_lens,
                                const float* tok_emb, const float* cls_emb, const float* prefix_vec, const float* pe,
                                float* out, void* out_bf16, float* mask, int B, int T, int D, int prefix, float scale,
                                int vocab, void* stream) {
  return msx_embed_fwd_p(tokens, classes, seq_lens, tok_emb, cls_emb, prefix_vec, pe, out, out_bf16, nullptr, mask, B, T, D,
                         prefix, scale, vocab, stream);
}

extern "C" int msx_embed_fwd_p(const int32_t* tokens, const int32_t* classes, const int32_t* seq_lens,
                               const float* tok_emb, const float* cls_emb, const float* prefix_vec, const float* pe,
                               float* out, void* out_bf16, void* out_bf16_lo, float* mask, int B, int T, int D, int prefix,
                               float scale, int vocab, void* stream) {
  MSX_REQUIRE(tokens && tok_emb && (out || out_bf16), "msx_embed_fwd: null pointer");
  MSX_REQUIRE(!out_bf16_lo || out_bf16, "msx_embed_fwd_p: a lo plane needs the hi plane");
  MSX_REQUIRE(prefix == 0 || prefix_vec, "msx_embed_fwd: prefix rows need prefix_vec");
  MSX_REQUIRE(!cls_emb || classes, "msx_embed_fwd: class embedding needs classes");
  if (B == 0 || T + prefix == 0) return MSX_OK;
  const long long rows = (long long)B * (T + prefix);
  const int wpb = 8;
  const int vec = (D & 3) == 0 && (((uintptr_t)tok_emb | (uintptr_t)cls_emb | (uintptr_t)prefix_vec | (uintptr_t)pe | (uintptr_t)out |
                                    (uintptr_t)out_bf16 | (uintptr_t)out_bf16_lo) & 15) == 0;
  MSX_CUDA(msx_launch(embed_fwd_kernel, dim3(msx_ceil_div(rows, wpb)), dim3(wpb * 32), 0, (cudaStream_t)stream, 
      tokens, classes, seq_lens, tok_emb, cls_emb, prefix_vec, pe, out, reinterpret_cast<unsigned short*>(out_bf16),
      reinterpret_cast<unsigned short*>(out_bf16_lo), mask, B, T, D, prefix, scale, vocab, vec));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

namespace {
// out[b*T + t, :] = scale * (tab[C + tokens[b, t], :] + tab[classes[b], :]) + postab[t, :]; blockDim.x = N / 4.
// A CTA owns ONE position t and walks a slice of the batch: the position row stays in registers, the C class rows stay in
// L1, so per output row only the token row travels from L2 (three table rows per output row made the kernel L2-bound).
__global__ void __launch_bounds__(256) rows_from_tables_kernel(const int* __restrict__ tokens, const int* __restrict__ classes,
                                                               const float* __restrict__ tab, const float* __restrict__ postab,
                                                               float* __restrict__ out, int B, int T, int N, int C, int V,
                                                               float scale, int bchunks) {
  pdl_entry();
  const int c = threadIdx.x * 4;
  const int t = blockIdx.x % T, chunk = blockIdx.x / T;
  const float4 p = __ldg(reinterpret_cast<const float4*>(postab + (size_t)t * N + c));
  constexpr int U = 4;                                      // batch rows per iteration
  for (int b0 = chunk * U; b0 < B; b0 += bchunks * U) {
    float4 a[U], k[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int b = min(b0 + u, B - 1);
      const int tok = min(max(__ldg(tokens + (size_t)b * T + t), 0), V - 1), cls = min(max(__ldg(classes + b), 0), C - 1);
      a[u] = __ldg(reinterpret_cast<const float4*>(tab + (size_t)(C + tok) * N + c));
      k[u] = __ldg(reinterpret_cast<const float4*>(tab + (size_t)cls * N + c));
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (b0 + u < B)
        *reinterpret_cast<float4*>(out + ((size_t)(b0 + u) * T + t) * N + c) =
            make_float4(fmaf(scale, a[u].x + k[u].x, p.x), fmaf(scale, a[u].y + k[u].y, p.y), fmaf(scale, a[u].z + k[u].z, p.z),
                        fmaf(scale, a[u].w + k[u].w, p.w));
  }
}
}  // namespace

// The first encoder layer's K|Q|V projection without a [B*T, D] x [D, 3D] GEMM.  Its input is
// x0 = scale * (emb[token] + class2hid[class]) + pe[position] (Encoder.hybrid_forward, model.py:84-93, PositionalEmbeddings,
// transformer.py:204-231), so x0 W^T + b = scale * (tab[C + token] + tab[class]) + postab[position] with the per-step tables
// tab [C + V, N] = [class2hid; emb] W^T and postab [T, N] = pe[:T] W^T + b (two GEMMs over 295 and T rows).  One 3 KB row
// written per position, three L2-resident table rows read.
extern "C" int msx_rows_from_tables(const int32_t* tokens, const int32_t* classes, const float* tab, const float* postab,
                                    float* out, int B, int T, int N, int C, int V, float scale, void* stream) {
  MSX_REQUIRE(tokens && classes && tab && postab && out, "msx_rows_from_tables: null pointer");
  MSX_REQUIRE(N >= 4 && N <= 1024 && (N & 3) == 0 && (((uintptr_t)tab | (uintptr_t)postab | (uintptr_t)out) & 15) == 0,
              "msx_rows_from_tables: N %% 4 == 0, N <= 1024, 16-byte aligned tables and output");
  MSX_REQUIRE(C >= 1 && V >= 1, "msx_rows_from_tables: empty table");
  if (B == 0 || T == 0) return MSX_OK;
  // T x bchunks CTAs: about 8 resident CTAs per SM
  int bchunks = (msx_num_sms() * 8 + T - 1) / T;
  if (bchunks > (B + 3) / 4) bchunks = (B + 3) / 4;
  MSX_CUDA(msx_launch(rows_from_tables_kernel, dim3(T * bchunks), dim3(N / 4), 0, (cudaStream_t)stream, tokens, classes, tab, postab, out, B, T, N, C, V, scale,
                                                                          bchunks));
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
  // large problems without prefix rows and with at most two classes: sort-and-reduce kernel (see above); the class term
  // needs class ids in {0, 1} (num_classes <= 2, msx_embed_bwd_ex) — other shapes keep the scatter kernel below
  if (prefix == 0 && rows >= 8192 && vocab <= kSortMaxV && D <= kSortMaxD && (D & 3) == 0 &&
      ((((uintptr_t)dout) | ((uintptr_t)d_tok_emb) | ((uintptr_t)d_cls_emb)) & 15) == 0 &&
      (!d_cls_emb || (num_classes >= 1 && num_classes <= 2))) {
    const int grid = (int)((rows + kSortSeg - 1) / kSortSeg);
    MSX_CUDA(msx_launch(embed_bwd_sorted_kernel, dim3(grid), dim3(kSortThreads), 0, (cudaStream_t)stream, tokens, classes, dout, d_tok_emb, d_cls_emb, rows,
                                                                             T, D, scale, vocab));
    MSX_LAUNCH_CHECK();
    return MSX_OK;
  }
  // (a variant that accumulated a 64-column slice of the table with shared-memory atomics was 6x SLOWER: float
  // atomicAdd on shared memory is a CAS loop and the synthetic 4/4 rows hit few distinct tokens)
  MSX_CUDA(msx_launch(embed_bwd_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, tokens, classes, dout, d_tok_emb, d_cls_emb, d_prefix, B, T, D,
                                                        prefix, scale, vocab));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
