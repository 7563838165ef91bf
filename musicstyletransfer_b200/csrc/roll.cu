// Piano-roll front end of the VarAutoEncoder step (`--featurisation roll`): the training path over K1's piano-roll tensors
// with the sigmoid-BCE reconstruction loss the north star names (BinaryCrossEntropy,
// /root/reference/music_style_transfer/VarAutoEncoder/loss.py:27-81; roll layout [B, S, P] time-major as loss.py:56,76-79
// and utils.py:52-61 expect).  HEAD no longer has a roll model, so the model side is a derived specification
// (oracle/roll_model.py, "parity unpinned"): the token embedding of Encoder / LSTMDecoder (model.py:86-91,176) becomes a
// Dense layer over the multi-hot pitch vector of a slice, everything else (class embedding, sqrt(D) scale + positional
// encodings, Transformer encoder, reparameterisation, LSTM decoder, output Dense) is the token model's.
//
//   msx_roll_features    uint8 roll [B, S, 128] -> fp32 GEMM operands  Renc [B, S+1, 132], Rdec [B, S, 132]:
//                        column p < 128 = (roll > 0), column 128 = start-of-sequence flag, 129..131 zero (16-byte rows);
//                        Renc row 0 = SOS, rows 1..S = the slices; Rdec row t = Renc row t (teacher forcing: slice t-1)
//   msx_embed_dense_fwd  out[b,t,:] = scale * (E[b,t,:] + cls_emb[classes[b]]) + pe[t]      (model.py:89-91, transformer.py:270)
//   msx_embed_dense_bwd  dE = scale * dout;  d_cls_emb[classes[b]] += scale * sum_t dout[b,t,:]
#include "msx_common.cuh"

namespace {

constexpr int kP = 128, kW = 132;

__global__ void __launch_bounds__(256) roll_features_kernel(const uint8_t* __restrict__ roll, float* __restrict__ renc,
                                                            float* __restrict__ rdec, int B, int S) {
  // one warp per (b, t) row of Renc, t in [0, S]; lane handles 4 pitches + (lane 0) the 4 extra columns
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * (S + 1)) return;
  const int b = (int)(row / (S + 1)), t = (int)(row % (S + 1));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t > 0) {
    const uchar4 u = *reinterpret_cast<const uchar4*>(roll + ((size_t)b * S + (t - 1)) * kP + lane * 4);
    v = make_float4(u.x ? 1.f : 0.f, u.y ? 1.f : 0.f, u.z ? 1.f : 0.f, u.w ? 1.f : 0.f);
  }
  const float4 tail = make_float4(t == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f);
  float* e = renc + (size_t)row * kW;
  *reinterpret_cast<float4*>(e + lane * 4) = v;
  if (lane == 0) *reinterpret_cast<float4*>(e + kP) = tail;
  if (t < S) {
    float* d = rdec + ((size_t)b * S + t) * kW;
    *reinterpret_cast<float4*>(d + lane * 4) = v;
    if (lane == 0) *reinterpret_cast<float4*>(d + kP) = tail;
  }
}

__global__ void __launch_bounds__(256) embed_dense_fwd_kernel(const float* __restrict__ E, const int* __restrict__ classes,
                                                              const float* __restrict__ cls_emb, const float* __restrict__ pe,
                                                              float* __restrict__ out, int B, int T, int D4, float scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // float4 index
  if (i >= (long long)B * T * D4) return;
  const int c4 = (int)(i % D4);
  const long long row = i / D4;
  const int t = (int)(row % T), b = (int)(row / T);
  const float4 e = reinterpret_cast<const float4*>(E)[i];
  const float4 c = cls_emb ? __ldg(reinterpret_cast<const float4*>(cls_emb) + (size_t)__ldg(classes + b) * D4 + c4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 p = pe ? __ldg(reinterpret_cast<const float4*>(pe) + (size_t)t * D4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  reinterpret_cast<float4*>(out)[i] = make_float4(fmaf(scale, e.x + c.x, p.x), fmaf(scale, e.y + c.y, p.y),
                                                  fmaf(scale, e.z + c.z, p.z), fmaf(scale, e.w + c.w, p.w));
}

// one CTA per sequence: dE rows = scale * dout rows, class-embedding gradient = scale * column sums over the T rows
__global__ void __launch_bounds__(128) embed_dense_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ classes,
                                                              float* __restrict__ dE, float* __restrict__ d_cls_emb, int T,
                                                              int D, float scale) {
  const int b = blockIdx.x;
  const float* src = dout + (size_t)b * T * D;
  float* dst = dE + (size_t)b * T * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      const float v = scale * src[(size_t)t * D + c];
      dst[(size_t)t * D + c] = v;
      acc += v;
    }
    if (d_cls_emb) atomicAdd(d_cls_emb + (size_t)classes[b] * D + c, acc);
  }
}

}  // namespace

extern "C" int msx_roll_features(const uint8_t* roll, float* renc, float* rdec, int B, int S, void* stream) {
  MSX_REQUIRE(B >= 0 && S >= 1, "msx_roll_features: bad sizes");
  if (B == 0) return MSX_OK;
  MSX_REQUIRE(roll && renc && rdec, "msx_roll_features: null pointer");
  MSX_REQUIRE((((uintptr_t)roll & 3) | ((uintptr_t)renc & 15) | ((uintptr_t)rdec & 15)) == 0, "msx_roll_features: misaligned buffer");
  const long long rows = (long long)B * (S + 1);
  roll_features_kernel<<<msx_ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(roll, renc, rdec, B, S);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_embed_dense_fwd(const float* E, const int32_t* classes, const float* cls_emb, const float* pe, float* out,
                                   int B, int T, int D, float scale, void* stream) {
  MSX_REQUIRE(B >= 0 && T >= 1 && D >= 4 && (D & 3) == 0, "msx_embed_dense_fwd: D must be a positive multiple of 4");
  if (B == 0) return MSX_OK;
  MSX_REQUIRE(E && out && (!cls_emb || classes), "msx_embed_dense_fwd: null pointer");
  const long long n4 = (long long)B * T * (D / 4);
  embed_dense_fwd_kernel<<<msx_ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>(E, classes, cls_emb, pe, out, B, T, D / 4, scale);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_embed_dense_bwd(const float* dout, const int32_t* classes, float* dE, float* d_cls_emb, int B, int T, int D,
                                   float scale, void* stream) {
  MSX_REQUIRE(B >= 0 && T >= 1 && D >= 1, "msx_embed_dense_bwd: bad sizes");
  if (B == 0) return MSX_OK;
  MSX_REQUIRE(dout && dE && (!d_cls_emb || classes), "msx_embed_dense_bwd: null pointer");
  embed_dense_bwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(dout, classes, dE, d_cls_emb, T, D, scale);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
