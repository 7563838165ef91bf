// A2 on the device + row utilities of the step.
//
// msx_rows_plan / msx_rows_build / msx_rows_gather_batch replace MelodyDataset._get_token_arrays, _count_sequence_length
// and the per-batch part of _preprocess_batch (/root/reference/music_style_transfer/VarAutoEncoder/data.py:133-198): K1's
// per-track token streams become the `tokens [R, L+1]`, `labels [R, L+1]`, `classes [R]`, `seq_lens [R]` rows the step
// consumes, with the reference's quirks (the possibly empty remainder row of every melody, the duplicated last row per
// class, EOS written by NumPy advanced indexing into every column that is SOME row's length).
//
// Strided row copy / add:  The encoder's output is read at the SOS position only
// (/root/reference/music_style_transfer/VarAutoEncoder/model.py:97-100: `last = out[:, 0, :]`), so the top encoder layer
// works on one row per sequence after its attention; these move those rows between the [B*T, D] and [B, D] layouts.
#include "msx_common.cuh"

namespace {

// ADD = false: out[r, c] = in[r, c];  ADD = true: out[r, c] += in[r, c]   (row r of X starts at X + r * ldX), float4 columns
template <bool ADD>
__global__ void __launch_bounds__(256) rows_strided_kernel(const float* __restrict__ in, long long ld_in, float* __restrict__ out,
                                                           long long ld_out, int rows, int width4) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * width4) return;
  const long long r = i / width4;
  const int c = (int)(i % width4);
  const float4 v = *(reinterpret_cast<const float4*>(in + r * ld_in) + c);
  float4* o = reinterpret_cast<float4*>(out + r * ld_out) + c;
  if (ADD) {
    const float4 w = *o;
    *o = make_float4(w.x + v.x, w.y + v.y, w.z + v.z, w.w + v.w);
  } else {
    *o = v;
  }
}

constexpr int kPad = 0, kSos = 1, kEos = 2;      // MIDIUtil/defaults.py:44-46
constexpr int kPlanThreads = 1024;

// One CTA.  Tracks arrive grouped by class in the reference's iteration order (sorted class names, melodies in file order).
// Track t yields n_t / L full rows and one remainder row (data.py:139-150); after the last track of class c the most recent
// remainder row is appended once more when it is non-empty (data.py:152-155; for a class without tracks `tokens` still
// holds the previous class's buffer, so the previous class's remainder is the one that is tested and copied).
//   row_start[t]      first row of track t            (row_start[N] = total rows R)
//   dup_row[c]        row index of class c's duplicate row, or -1;  dup_src[c] = the track whose remainder row it copies
//   len_present[j]    1 when some row has exactly j non-PAD data tokens, j in [0, L]   (labels[:, seq_lens] = EOS)
__global__ void __launch_bounds__(kPlanThreads) rows_plan_kernel(const int* __restrict__ n_tokens, const int* __restrict__ class_start,
                                                                int n_tracks, int n_classes, int L, int* __restrict__ row_start,
                                                                int* __restrict__ dup_row, int* __restrict__ dup_src,
                                                                int* __restrict__ len_present) {
  pdl_entry();
  __shared__ int scan[kPlanThreads];
  __shared__ int carry;
  const int tid = threadIdx.x;
  for (int j = tid; j <= L; j += blockDim.x) len_present[j] = 0;
  if (tid == 0) carry = 0;
  __syncthreads();
  // rows per track, plus one for the track that closes a class with a non-empty remainder row (the duplicate follows it)
  for (int base = 0; base < n_tracks; base += kPlanThreads) {
    const int t = base + tid;
    int rows = 0;
    if (t < n_tracks) {
      const int n = n_tokens[t];
      rows = n / L + 1;
      len_present[n % L] = 1;                        // remainder row length (0 = all PAD)
      if (n >= L) len_present[L] = 1;                // a full row
    }
    scan[tid] = rows;
    __syncthreads();
    for (int o = 1; o < kPlanThreads; o <<= 1) {     // Hillis-Steele inclusive scan of this chunk
      const int v = tid >= o ? scan[tid - o] : 0;
      __syncthreads();
      scan[tid] += v;
      __syncthreads();
    }
    if (t < n_tracks) row_start[t] = carry + scan[tid] - rows;     // exclusive, duplicates not yet counted
    __syncthreads();
    if (tid == kPlanThreads - 1) carry += scan[tid];
    __syncthreads();
  }
  if (tid == 0) row_start[n_tracks] = carry;        // rows of all tracks, duplicates not yet counted
  __syncthreads();
  // duplicates: sequential over the classes.  Class c's tracks shift by the duplicates of the classes before it; its own
  // duplicate (if any) sits right behind its last track's rows.
  if (tid == 0) {
    int added = 0, last_track = -1;
    for (int c = 0; c < n_classes; ++c) {
      const int t0 = class_start[c], t1 = class_start[c + 1];
      if (t1 > t0) last_track = t1 - 1;
      for (int t = t0; t < t1; ++t) row_start[t] += added;
      const bool has_dup = last_track >= 0 && n_tokens[last_track] % L != 0;
      dup_row[c] = has_dup ? row_start[t1] + added : -1;      // row_start[t1] is still the un-shifted count of rows before t1
      dup_src[c] = last_track;
      if (has_dup) ++added;
    }
    row_start[n_tracks] += added;
  }
}

// writes one output row: data = L ids (PAD beyond n_valid)
__device__ __forceinline__ void rows_write(const int* __restrict__ src, int n_valid, int L, int cls,
                                           const int* __restrict__ len_present, int* __restrict__ tok, int* __restrict__ lab,
                                           int* __restrict__ classes, int* __restrict__ seq_lens, long long row, int lane_tid,
                                           int nthreads) {
  int* trow = tok + row * (L + 1);
  int* lrow = lab + row * (L + 1);
  for (int j = lane_tid; j <= L; j += nthreads) {
    const int d_prev = (j >= 1 && j - 1 < n_valid) ? src[j - 1] : kPad;      // tokens = [SOS | data]
    trow[j] = j == 0 ? kSos : d_prev;
    const int d = (j < L && j < n_valid) ? src[j] : kPad;                     // labels = [data | PAD], then EOS columns
    lrow[j] = len_present[j] ? kEos : d;
  }
  if (lane_tid == 0) {
    classes[row] = cls;
    seq_lens[row] = n_valid + 1;                                              // non-PAD count including SOS (data.py:175-179,189)
  }
}

// one CTA per track (blockIdx.x < n_tracks) or per class duplicate (blockIdx.x >= n_tracks)
__global__ void __launch_bounds__(128) rows_build_kernel(const int* __restrict__ ids, long long ld, int col0,
                                                         const int* __restrict__ n_tokens, const int* __restrict__ track_class,
                                                         int n_tracks, int n_classes, int L, const int* __restrict__ row_start,
                                                         const int* __restrict__ dup_row, const int* __restrict__ dup_src,
                                                         const int* __restrict__ len_present, int* __restrict__ tok,
                                                         int* __restrict__ lab, int* __restrict__ classes,
                                                         int* __restrict__ seq_lens) {
  pdl_entry();
  const int b = blockIdx.x;
  if (b < n_tracks) {
    const int n = n_tokens[b];
    const int* src = ids + (long long)b * ld + col0;
    const int rows = n / L + 1;
    const long long r0 = row_start[b];
    for (int r = 0; r < rows; ++r) {
      const int valid = min(L, n - r * L);
      rows_write(src + (long long)r * L, valid, L, track_class[b], len_present, tok, lab, classes, seq_lens, r0 + r,
                 threadIdx.x, blockDim.x);
    }
  } else {
    const int c = b - n_tracks;
    const int row = dup_row[c];
    if (row < 0) return;
    const int t = dup_src[c];
    const int n = n_tokens[t];
    const int* src = ids + (long long)t * ld + col0 + (long long)(n / L) * L;
    rows_write(src, n % L, L, c, len_present, tok, lab, classes, seq_lens, row, threadIdx.x, blockDim.x);
  }
}

// one CTA per batch row: out rows are trimmed to T_out columns
__global__ void __launch_bounds__(128) rows_gather_kernel(const int* __restrict__ tok, const int* __restrict__ lab,
                                                          const int* __restrict__ classes, const int* __restrict__ seq_lens,
                                                          const int* __restrict__ index, int ld, int T_out,
                                                          int* __restrict__ btok, int* __restrict__ blab, int* __restrict__ bcls,
                                                          int* __restrict__ blen) {
  pdl_entry();
  const int b = blockIdx.x;
  const long long r = index[b];
  for (int j = threadIdx.x; j < T_out; j += blockDim.x) {
    btok[(long long)b * T_out + j] = tok[r * ld + j];
    blab[(long long)b * T_out + j] = lab[r * ld + j];
  }
  if (threadIdx.x == 0) {
    bcls[b] = classes[r];
    blen[b] = seq_lens[r];
  }
}

// out[e] = x[e] * keep(e) / (1 - p): the inter-layer dropout of the stacked LSTM decoder (gluon.rnn.LSTM(dropout=...),
// model.py:148-153); the backward applies the same mask (same seed / site) to the gradient.  In place when out == x.
__global__ void __launch_bounds__(256) dropout_kernel(const float* x, float* out, long long n4, float p, float inv_keep,
                                                      unsigned long long seed, const unsigned long long* ctr, unsigned site) {
  pdl_entry();
  const unsigned long long eff = msx_eff_seed(seed, ctr);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
    float k[4];
    dropout_scale4(eff, site, (uint64_t)q, p, inv_keep, k);
    const float4 v = reinterpret_cast<const float4*>(x)[q];
    reinterpret_cast<float4*>(out)[q] = make_float4(v.x * k[0], v.y * k[1], v.z * k[2], v.w * k[3]);
  }
}

// labels of the Transformer decoder: its rows carry the latent prefix position in front (model.py:244-253 drops it again
// after the layers), which never has a target: out[b, 0] = PAD, out[b, 1 + t] = labels[b, t]
__global__ void __launch_bounds__(256) prefix_labels_kernel(const int* __restrict__ labels, int* __restrict__ out, int B, int T) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * (T + 1)) return;
  const int b = (int)(i / (T + 1)), t = (int)(i % (T + 1));
  out[i] = t == 0 ? kPad : labels[(long long)b * T + t - 1];
}

}  // namespace

extern "C" int msx_prefix_labels(const int32_t* labels, int32_t* out, int B, int T, void* stream) {
  MSX_REQUIRE(B >= 0 && T >= 1, "msx_prefix_labels: bad sizes");
  if (B == 0) return MSX_OK;
  MSX_REQUIRE(labels && out, "msx_prefix_labels: null pointer");
  MSX_CUDA(msx_launch(prefix_labels_kernel, dim3(msx_ceil_div((long long)B * (T + 1), 256)), dim3(256), 0, (cudaStream_t)stream, labels, out, B, T));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_dropout(const float* x, float* out, long long n, float drop_p, unsigned long long seed, unsigned site,
                           void* stream) {
  MSX_REQUIRE(n >= 0 && (n & 3) == 0, "msx_dropout: n must be a non-negative multiple of 4");
  if (n == 0) return MSX_OK;
  MSX_REQUIRE(x && out && ((((uintptr_t)x) | ((uintptr_t)out)) & 15) == 0, "msx_dropout: null or misaligned pointer");
  MSX_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "msx_dropout: dropout probability must be in [0,1)");
  const long long n4 = n / 4;
  const long long want = (n4 + 255) / 256;
  const int grid = (int)(want < (long long)msx_num_sms() * 8 ? want : (long long)msx_num_sms() * 8);
  MSX_CUDA(msx_launch(dropout_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, out, n4, drop_p, 1.f / (1.f - drop_p), seed, msx_step_counter(), site));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_rows_plan(const int32_t* n_tokens, const int32_t* class_start, int n_tracks, int n_classes, int max_seq_len,
                             int32_t* row_start, int32_t* dup_row, int32_t* dup_src, int32_t* len_present, void* stream) {
  MSX_REQUIRE(n_tracks >= 0 && n_classes >= 1 && max_seq_len >= 1, "msx_rows_plan: bad sizes");
  MSX_REQUIRE(n_tokens && class_start && row_start && dup_row && dup_src && len_present, "msx_rows_plan: null pointer");
  MSX_CUDA(msx_launch(rows_plan_kernel, dim3(1), dim3(kPlanThreads), 0, (cudaStream_t)stream, n_tokens, class_start, n_tracks, n_classes, max_seq_len,
                                                                 row_start, dup_row, dup_src, len_present));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_rows_build(const int32_t* ids, long long ld, int col0, const int32_t* n_tokens, const int32_t* track_class,
                              int n_tracks, int n_classes, int max_seq_len, const int32_t* row_start, const int32_t* dup_row,
                              const int32_t* dup_src, const int32_t* len_present, int32_t* tokens, int32_t* labels,
                              int32_t* classes, int32_t* seq_lens, void* stream) {
  MSX_REQUIRE(n_tracks >= 0 && n_classes >= 1 && max_seq_len >= 1 && col0 >= 0, "msx_rows_build: bad sizes");
  if (n_tracks == 0) return MSX_OK;
  MSX_REQUIRE(ids && n_tokens && track_class && row_start && dup_row && dup_src && len_present && tokens && labels && classes &&
                  seq_lens, "msx_rows_build: null pointer");
  MSX_CUDA(msx_launch(rows_build_kernel, dim3(n_tracks + n_classes), dim3(128), 0, (cudaStream_t)stream, ids, ld, col0, n_tokens, track_class, n_tracks,
                                                                            n_classes, max_seq_len, row_start, dup_row, dup_src,
                                                                            len_present, tokens, labels, classes, seq_lens));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_rows_gather_batch(const int32_t* tokens, const int32_t* labels, const int32_t* classes, const int32_t* seq_lens,
                                     const int32_t* index, int batch, int ld, int t_out, int32_t* b_tokens, int32_t* b_labels,
                                     int32_t* b_classes, int32_t* b_seq_lens, void* stream) {
  MSX_REQUIRE(batch >= 0 && t_out >= 1 && t_out <= ld, "msx_rows_gather_batch: bad sizes");
  if (batch == 0) return MSX_OK;
  MSX_REQUIRE(tokens && labels && classes && seq_lens && index && b_tokens && b_labels && b_classes && b_seq_lens,
              "msx_rows_gather_batch: null pointer");
  MSX_CUDA(msx_launch(rows_gather_kernel, dim3(batch), dim3(128), 0, (cudaStream_t)stream, tokens, labels, classes, seq_lens, index, ld, t_out, b_tokens,
                                                              b_labels, b_classes, b_seq_lens));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

extern "C" int msx_rows_strided(const float* in, long long ld_in, float* out, long long ld_out, int rows, int width, int add,
                                void* stream) {
  MSX_REQUIRE(rows >= 0 && width >= 0, "msx_rows_strided: negative size");
  if (rows == 0 || width == 0) return MSX_OK;
  MSX_REQUIRE(in && out, "msx_rows_strided: null pointer");
  MSX_REQUIRE((width & 3) == 0 && (ld_in & 3) == 0 && (ld_out & 3) == 0 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0,
              "msx_rows_strided: width and leading dimensions must be multiples of 4, pointers 16-byte aligned");
  const long long n = (long long)rows * (width / 4);
  if (add)
    MSX_CUDA(msx_launch(rows_strided_kernel<true>, dim3(msx_ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, in, ld_in, out, ld_out, rows, width / 4));
  else
    MSX_CUDA(msx_launch(rows_strided_kernel<false>, dim3(msx_ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, in, ld_in, out, ld_out, rows, width / 4));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
